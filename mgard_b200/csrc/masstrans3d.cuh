// Fused 3-D load-vector kernel (sm_100a): mass matrix x restriction along f, c
// and r of the level-l coefficient function, read straight from the
// coefficient array in its coarse-first layout.
//
// Replaces, for D == 3, the reference's Lpk1Reo3D -> Lpk2Reo3D -> Lpk3Reo3D
// chain (Correction/LinearProcessingKernel3D.hpp:27-1090, mass_trans:
// Correction/LPKFunctor.h:47-66) with the same arithmetic in the same order, so
// results stay bit-identical, in one pass: n_l elements read, n_l/8 written.
//
// In the coarse-first layout the even (E) and odd (O) padded positions of a
// line are two contiguous vectors, and the five inputs of mass_trans at coarse
// index i are E[i-1], O[i-1], E[i], O[i], E[i+1].  So
//   f pass: a warp takes one row; every lane loads E[kf] and O[kf] (two
//           coalesced loads) and gets its neighbours' values with shuffles (the
//           warp spans 32 columns and owns the inner 30);
//   c pass: from the f-pass rows of the tile in shared memory;
//   r pass: from a five-deep register ring while the block sweeps r.
// A thread block owns a TC x TF tile of coarse (c, f) columns and a segment of
// coarse r indices; the loads of the next plane are issued before the current
// one is worked on.
#pragma once

namespace masstrans3d {

typedef long long i64;

// A warp covers 32 consecutive coarse f columns kf0-1 .. kf0+30 and OWNS the
// middle TF = 30: the two outer lanes only supply their neighbours' inputs, so
// no separate halo loads are needed.
constexpr int TC = 8, TF = 30, NT = 256, NW = NT / 32;
constexpr int NROW = 2 * TC + 3;                  // E rows kc0-1..kc0+TC, O rows kc0-1..kc0+TC-1
constexpr int RPW = (NROW + NW - 1) / NW;         // rows per warp and plane
constexpr int NST = 4;                            // raw planes in flight (power of two)

template <typename T> struct Params {
  int n[3], nc[3];  // fine / coarse level shape (r, c, f)
  i64 sin[3];       // coefficient array strides
  i64 sw[3];        // dense load-vector strides
  const T *mt[3];   // 9 x nc mass_trans tables
  int rsegs, ctiles, ftiles;
};

template <typename T>
__device__ __forceinline__ T mass_trans_k(T a, T b, T c, T d, T e, const T (&k)[9]) {
  T tb = a * k[0] + b * k[1] + c * k[2];
  T tc = b * k[2] + c * k[3] + d * k[4];
  T td = c * k[4] + d * k[5] + e * k[6];
  tc += tb * k[7] + td * k[8];
  return tc;
}

// position of the m-th even (odd = false) / odd (odd = true) padded node of a
// dimension in the coarse-first layout; -1: no such node (its value is zero)
__device__ __forceinline__ int pos(int m, bool odd, int n, int nc) {
  if (m < 0)
    return -1;
  const int p = odd ? nc + m : m;
  return (odd ? p < n : m < nc) ? p : -1;
}

template <typename T>
__global__ void __launch_bounds__(NT, 3)
masstrans3d_kernel(const Params<T> P, const T *__restrict__ in, T *__restrict__ w_out) {
  __shared__ T s_a1[2][NROW][32];
  __shared__ T s_raw[NST][NROW][2][32]; // [stage][row][E / O][lane] // f-pass rows of the current plane (double buffered)
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int bid = blockIdx.x;
  const int ft = bid % P.ftiles;
  bid /= P.ftiles;
  const int ct = bid % P.ctiles;
  const int rs = bid / P.ctiles;
  const int kc0 = ct * TC, kf0 = ft * TF;
  const int nr = P.n[0], ncn = P.n[1], nf = P.n[2];
  const int rr = P.nc[0], cc = P.nc[1], ff = P.nc[2];
  const int per = (rr + P.rsegs - 1) / P.rsegs;
  const int rk0 = rs * per, rk1 = min(rr, rk0 + per);
  if (rk0 >= rk1)
    return;

  // ---- per-thread constants -------------------------------------------------
  // f: this lane's column, offsets of its E and O element (-1: none -> zero)
  const int kf = kf0 - 1 + lane;
  const bool own_f = lane >= 1 && lane <= TF && kf < ff;
  const int pe = pos(kf, false, nf, ff), po = pos(kf, true, nf, ff);
  // rows of this warp: per row the offsets of the lane's E and O element from the
  // plane base (32-bit; -1: nothing to load)
  // (offsets of missing elements are 0 with a copy size of 0 = zero fill)
  int roff_e[RPW], roff_o[RPW];
  int se_even[RPW], se_odd[RPW], so_any[RPW]; // copy sizes on even-r / odd-r planes
#pragma unroll
  for (int q = 0; q < RPW; q++) {
    const int row = wid + q * NW;
    roff_e[q] = roff_o[q] = 0;
    se_even[q] = se_odd[q] = so_any[q] = 0;
    if (row < NROW) {
      const bool odd = row >= TC + 2;
      const int m = kc0 - 1 + (odd ? row - (TC + 2) : row);
      const int pc = pos(m, odd, ncn, cc);
      if (pc >= 0) {
        if (pe >= 0) {
          roff_e[q] = (int)((i64)pc * P.sin[1] + (i64)pe * P.sin[2]);
          se_odd[q] = (int)sizeof(T);
          // the all-coarse block (even r, even c, even f) counts as zero
          se_even[q] = odd ? (int)sizeof(T) : 0;
        }
        if (po >= 0) {
          roff_o[q] = (int)((i64)pc * P.sin[1] + (i64)po * P.sin[2]);
          so_any[q] = (int)sizeof(T);
        }
      }
    }
  }
  T kfc[9], kcc[9];
#pragma unroll
  for (int m = 0; m < 9; m++) {
    kfc[m] = (kf >= 0 && kf < ff) ? P.mt[2][m * ff + kf] : (T)0;
    kcc[m] = (kc0 + wid < cc) ? P.mt[1][m * cc + kc0 + wid] : (T)0;
  }
  const bool col_ok = (kc0 + wid < cc) && own_f;
  const i64 w_col = (i64)(kc0 + wid) * P.sw[1] + (i64)kf * P.sw[2];

  // ---- plane sequence: (E,O) of k = rk0-1 .. rk1-1, then E of rk1 ------------
  // plane index t: k = rk0 - 1 + t / 2, odd-r plane iff t & 1
  const int nplanes = 2 * (rk1 - rk0 + 1) + 1;
  // raw planes: ring of NST stages in shared memory filled with asynchronous
  // copies NST-1 planes ahead (a warp only ever reads the rows it requested
  // itself, so completion needs no block-wide barrier)
  const unsigned raw_addr = (unsigned)__cvta_generic_to_shared(&s_raw[0][0][0][0]);
  auto issue = [&](int t) {
    const int k = rk0 - 1 + (t >> 1);
    const bool rodd = t & 1;
    const int pr = pos(k, rodd, nr, rr);
    const T *base = in + (i64)(pr >= 0 ? pr : 0) * P.sin[0];
    const unsigned stage = raw_addr + (unsigned)((t & (NST - 1)) * (NROW * 64) * (int)sizeof(T));
#pragma unroll
    for (int q = 0; q < RPW; q++) {
      const int row = wid + q * NW;
      if (row < NROW) { // warp uniform
        const unsigned d = stage + (unsigned)((row * 64 + lane) * (int)sizeof(T));
        const T *ge = base + roff_e[q], *go = base + roff_o[q];
        const int se = pr >= 0 ? (rodd ? se_odd[q] : se_even[q]) : 0; // 0: zero fill
        const int so = pr >= 0 ? so_any[q] : 0;
        if (sizeof(T) == 4) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(ge), "r"(se));
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d + 32 * 4), "l"(go), "r"(so));
        } else {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(ge), "r"(se));
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d + 32 * 8), "l"(go), "r"(so));
        }
      }
    }
  };
  T ring[5] = {(T)0, (T)0, (T)0, (T)0, (T)0};
#pragma unroll
  for (int t = 0; t < NST - 1; t++) {
    if (t < nplanes)
      issue(t);
    asm volatile("cp.async.commit_group;\n" ::);
  }
  const int lm = lane > 0 ? lane - 1 : 0, lp = lane < 31 ? lane + 1 : 31;
  for (int t = 0; t < nplanes; t++) {
    if (t + NST - 1 < nplanes)
      issue(t + NST - 1);
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group %0;\n" ::"n"(NST - 1));
    __syncwarp();
    // f pass of plane t
    T a1[RPW];
    const T(*raw)[2][32] = s_raw[t & (NST - 1)];
#pragma unroll
    for (int q = 0; q < RPW; q++) {
      const int row = wid + q * NW;
      a1[q] = (T)0;
      if (row < NROW) // warp uniform; lanes 0 and 31 produce values nobody uses
        a1[q] = mass_trans_k<T>(raw[row][0][lm], raw[row][1][lm], raw[row][0][lane],
                                raw[row][1][lane], raw[row][0][lp], kfc);
    }
    T(*sa)[32] = s_a1[t & 1];
#pragma unroll
    for (int q = 0; q < RPW; q++) {
      const int row = wid + q * NW;
      if (row < NROW)
        sa[row][lane] = a1[q];
    }
    __syncthreads();
    // c pass: coarse row kc0 + wid from E rows wid, wid+1, wid+2 and O rows wid, wid+1
    const T a2 = mass_trans_k<T>(sa[wid][lane], sa[TC + 2 + wid][lane], sa[wid + 1][lane],
                                 sa[TC + 2 + wid + 1][lane], sa[wid + 2][lane], kcc);
    ring[0] = ring[1];
    ring[1] = ring[2];
    ring[2] = ring[3];
    ring[3] = ring[4];
    ring[4] = a2;
    // r pass: after plane E of k+1 (t even, t >= 4) the ring holds
    // E[k-1], O[k-1], E[k], O[k], E[k+1] for k = rk0 - 1 + t/2 - 1
    if (!(t & 1) && t >= 4) {
      const int k = rk0 - 2 + (t >> 1);
      if (col_ok && k >= rk0 && k < rk1) {
        T kr[9];
#pragma unroll
        for (int m = 0; m < 9; m++)
          kr[m] = P.mt[0][m * rr + k];
        w_out[(i64)k * P.sw[0] + w_col] =
            mass_trans_k<T>(ring[0], ring[1], ring[2], ring[3], ring[4], kr);
      }
    }
  }
}

} // namespace masstrans3d
