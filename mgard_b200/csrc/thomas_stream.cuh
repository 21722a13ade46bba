// Thomas solves along the CONTIGUOUS axis, streaming formulation (reference Ipk1Reo3D,
// IterativeProcessingKernel3D.hpp:27-420; tridiag_forward2 / tridiag_backward2,
// IPKFunctor.h:14-51).
//
// The recurrences are sequential per line (bit-exact contract: evaluated in the
// reference's order), so throughput = lines in flight / latency of one step.  Keeping whole
// lines resident between the two sweeps (thomas_tma.cuh) caps the lines in flight at what
// shared memory holds - 48 lines of 1025 nodes per SM.  Here a line is never resident: a
// block owns 128 lines (one per thread) and moves 64-column tiles of them through shared
// memory, left to right for the forward sweep (results written back to the array), right to
// left for the backward sweep.  Twice the traffic of the resident formulation (the forward
// results make a round trip, mostly through the L2), but 512 lines per SM make progress at
// once and the kernel runs at memory speed instead of at the speed of 48 dependent chains.
//   * tile in: warp w loads its 32 lines row by row, lane l reads columns l and l + 32 - 256
//     contiguous bytes per row; the values of the NEXT tile are already in registers while the current
//     one is solved (software pipeline);
//   * solve: thread t walks line t through the tile in shared memory (pitch 65: no bank
//     conflicts), carrying the recurrence in a register;
//   * tile out: the same rows, coalesced.
#pragma once
#include <cuda_runtime.h>

namespace thomas_stream {

// x / b of the backward sweep.  fp32: the division written out as in thomas_tma.cuh
// (mgb_tma::div_by: the same instructions on the same operands as the compiler's
// division, bit-identical inside the guarded range) with the part that depends on b
// alone - reciprocal and its refinement - off the dependent chain of the recurrence, and
// the range guard as a plain branch to the exact division (never taken on sane data).
__device__ __forceinline__ float div_chain(float x, float b) {
  const float y = mgb_tma::refined_rcp(b);
  const bool b_ok = mgb_tma::rcp_in_range(b);
  const float ax = fabsf(x);
  const bool oor = !b_ok || !(ax < 1.1529215e18f) || (ax <= 8.6736174e-19f && ax != 0.0f);
  const float q0 = __fmul_rn(x, y);
  const float r = __fmaf_rn(-b, q0, x);
  float q = __fmaf_rn(y, r, q0);
  q = ax == 0.0f ? x : q;
  if (oor)
    q = x / b;
  return q;
}
__device__ __forceinline__ double div_chain(double x, double b) { return x / b; }

constexpr int LINES = 128; // lines (= threads) per block
// columns per tile: 256 bytes of a line per access
template <typename T> struct Tile { static constexpr int W = 256 / (int)sizeof(T); };

// x: `lines` lines of n contiguous elements, solved in place.
template <typename T>
__global__ void __launch_bounds__(LINES, 4)
thomas_stream_kernel(T *__restrict__ x, int n, long long lines, const T *__restrict__ fw,
                     const T *__restrict__ am, const T *__restrict__ bm) {
  constexpr int TW = Tile<T>::W;
  __shared__ T tile[LINES][TW + 1];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const long long line0 = (long long)blockIdx.x * LINES;
  const int nl = (int)min((long long)LINES, lines - line0); // lines of this block
  // warp w moves lines w*32 .. w*32+31 of the block; thread tid solves line tid
  T *wbase = x + (line0 + w * 32) * (long long)n;
  const int wl = max(0, min(32, nl - w * 32)); // lines this warp moves
  const int ntiles = (n + TW - 1) / TW;
  T(*wt)[TW + 1] = tile + w * 32;
  T reg[TW / 32][32];
  auto load_tile = [&](int j) {
#pragma unroll
    for (int h = 0; h < TW / 32; h++) {
      const int c = j * TW + h * 32 + lane;
#pragma unroll
      for (int r = 0; r < 32; r++)
        reg[h][r] = (r < wl && c < n) ? wbase[(long long)r * n + c] : (T)0;
    }
  };
  auto tile_from_regs = [&]() {
#pragma unroll
    for (int h = 0; h < TW / 32; h++)
#pragma unroll
      for (int r = 0; r < 32; r++)
        wt[r][h * 32 + lane] = reg[h][r];
  };
  auto store_tile = [&](int j) {
#pragma unroll
    for (int h = 0; h < TW / 32; h++) {
      const int c = j * TW + h * 32 + lane;
      if (c < n) {
#pragma unroll
        for (int r = 0; r < 32; r++)
          if (r < wl)
            wbase[(long long)r * n + c] = wt[r][h * 32 + lane];
      }
    }
  };
  // a thread only touches the rows its own warp moves, so warp-level synchronisation is
  // all that is needed between the phases of a tile
  // ---- forward sweep -------------------------------------------------------------------
  T prev = (T)0;
  load_tile(0);
  for (int j = 0; j < ntiles; j++) {
    tile_from_regs();
    if (j + 1 < ntiles)
      load_tile(j + 1);
    __syncwarp();
    const int c0 = j * TW, cn = min(TW, n - c0);
    T *row = wt[lane];
    if (cn == TW) {
#pragma unroll
      for (int k = 0; k < TW; k++) {
        prev = row[k] - prev * __ldg(fw + c0 + k);
        row[k] = prev;
      }
    } else {
      for (int k = 0; k < cn; k++) {
        prev = row[k] - prev * __ldg(fw + c0 + k);
        row[k] = prev;
      }
    }
    __syncwarp();
    store_tile(j);
    __syncwarp();
  }
  // the forward results of this block's lines are read back by the same warps below
  __threadfence_block();
  // ---- backward sweep ------------------------------------------------------------------
  prev = (T)0;
  load_tile(ntiles - 1);
  for (int j = ntiles - 1; j >= 0; j--) {
    tile_from_regs();
    if (j > 0)
      load_tile(j - 1);
    __syncwarp();
    const int c0 = j * TW, cn = min(TW, n - c0);
    T *row = wt[lane];
    if (cn == TW) {
#pragma unroll
      for (int k = TW - 1; k >= 0; k--) {
        prev = div_chain(row[k] - __ldg(am + c0 + k + 1) * prev, __ldg(bm + c0 + k + 1));
        row[k] = prev;
      }
    } else {
      for (int k = cn - 1; k >= 0; k--) {
        prev = div_chain(row[k] - __ldg(am + c0 + k + 1) * prev, __ldg(bm + c0 + k + 1));
        row[k] = prev;
      }
    }
    __syncwarp();
    store_tile(j);
    __syncwarp();
  }
}

} // namespace thomas_stream
