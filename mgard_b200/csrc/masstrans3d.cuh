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
//           coalesced loads) and gets its neighbours' values with shuffles;
//   c pass: from the f-pass rows of the tile in shared memory;
//   r pass: from a five-deep register ring while the block sweeps r.
// A thread block owns a TC x TF tile of coarse (c, f) columns and a segment of
// coarse r indices; the loads of the next plane are issued before the current
// one is worked on.
#pragma once

namespace masstrans3d {

typedef long long i64;

constexpr int TC = 8, TF = 32, NT = 256, NW = NT / 32;
constexpr int NROW = 2 * TC + 3;                  // E rows kc0-1..kc0+TC, O rows kc0-1..kc0+TC-1
constexpr int RPW = (NROW + NW - 1) / NW;         // rows per warp and plane

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
  __shared__ T s_a1[2][NROW][TF]; // f-pass rows of the current plane (double buffered)
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
  // f: offsets of this lane's E and O element and of the halo element it fetches
  // (lane 0: E[kf0-1], lane 1: O[kf0-1], lane 2: E[kf0+TF])
  const int kf = kf0 + lane;
  const int pe = pos(kf, false, nf, ff), po = pos(kf, true, nf, ff);
  int px = -1;
  bool x_odd = false;
  if (lane == 0)
    px = pos(kf0 - 1, false, nf, ff);
  else if (lane == 1) {
    px = pos(kf0 - 1, true, nf, ff);
    x_odd = true;
  } else if (lane == 2)
    px = pos(kf0 + TF, false, nf, ff);
  const i64 off_e = (i64)pe * P.sin[2], off_o = (i64)po * P.sin[2], off_x = (i64)px * P.sin[2];
  // rows of this warp: row index q -> (c position, is odd-c row)
  i64 row_off[RPW];
  bool row_ok[RPW], row_odd[RPW];
#pragma unroll
  for (int q = 0; q < RPW; q++) {
    const int row = wid + q * NW;
    row_ok[q] = false;
    row_odd[q] = false;
    row_off[q] = 0;
    if (row < NROW) {
      const bool odd = row >= TC + 2;
      const int m = kc0 - 1 + (odd ? row - (TC + 2) : row);
      const int pc = pos(m, odd, ncn, cc);
      row_odd[q] = odd;
      if (pc >= 0) {
        row_ok[q] = true;
        row_off[q] = (i64)pc * P.sin[1];
      }
    }
  }
  T kfc[9], kcc[9];
#pragma unroll
  for (int m = 0; m < 9; m++) {
    kfc[m] = kf < ff ? P.mt[2][m * ff + kf] : (T)0;
    kcc[m] = (kc0 + wid < cc) ? P.mt[1][m * cc + kc0 + wid] : (T)0;
  }
  const bool col_ok = (kc0 + wid < cc) && (kf < ff);
  const i64 w_col = (i64)(kc0 + wid) * P.sw[1] + (i64)kf * P.sw[2];

  // ---- plane sequence: (E,O) of k = rk0-1 .. rk1-1, then E of rk1 ------------
  // plane index t: k = rk0 - 1 + t / 2, odd-r plane iff t & 1
  const int nplanes = 2 * (rk1 - rk0 + 1) + 1;
  T ve[RPW], vo[RPW], vx[RPW]; // prefetched raw values of the next plane
  auto fetch = [&](int t) {
    const int k = rk0 - 1 + (t >> 1);
    const bool rodd = t & 1;
    const int pr = pos(k, rodd, nr, rr);
    const T *base = in + (i64)(pr >= 0 ? pr : 0) * P.sin[0];
#pragma unroll
    for (int q = 0; q < RPW; q++) {
      ve[q] = vo[q] = vx[q] = (T)0;
      if (pr >= 0 && row_ok[q]) {
        const T *rp = base + row_off[q];
        // the all-coarse block (even r, even c, even f) counts as zero
        const bool ezero = !rodd && !row_odd[q];
        if (pe >= 0 && !ezero)
          ve[q] = rp[off_e];
        if (po >= 0)
          vo[q] = rp[off_o];
        if (px >= 0 && !(ezero && !x_odd))
          vx[q] = rp[off_x];
      }
    }
  };
  T ring[5] = {(T)0, (T)0, (T)0, (T)0, (T)0};
  fetch(0);
  for (int t = 0; t < nplanes; t++) {
    // f pass of plane t from the prefetched registers
    T a1[RPW];
#pragma unroll
    for (int q = 0; q < RPW; q++) {
      const T e = ve[q], o = vo[q], x = vx[q];
      T em1 = __shfl_up_sync(0xffffffffu, e, 1), om1 = __shfl_up_sync(0xffffffffu, o, 1);
      T ep1 = __shfl_down_sync(0xffffffffu, e, 1);
      const T x0 = __shfl_sync(0xffffffffu, x, 0), x1 = __shfl_sync(0xffffffffu, x, 1),
              x2 = __shfl_sync(0xffffffffu, x, 2);
      if (lane == 0) {
        em1 = x0;
        om1 = x1;
      }
      if (lane == 31)
        ep1 = x2;
      a1[q] = mass_trans_k<T>(em1, om1, e, o, ep1, kfc);
    }
    if (t + 1 < nplanes)
      fetch(t + 1);
    T(*sa)[TF] = s_a1[t & 1];
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
