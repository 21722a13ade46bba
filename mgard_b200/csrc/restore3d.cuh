// Tiled 3-D restore kernel (sm_100a): nodal values of level l from the coarse
// nodes of level l-1 plus the level-l coefficients.
//
// Replaces, for D == 3, the reference's GpkRev3D
// (Coefficient/GridProcessingKernel3D.hpp:1231-2400) with the same arithmetic
// in the same order (lerp along f, then c, then r, GPKFunctor.h:13-26; value =
// coefficient + interpolant), so results stay bit-identical.
//
// A thread block owns TR x TC x TF coarse cells.  It stages the
// (TR+1) x (TC+1) x (TF+1) coarse corner values in shared memory; a thread
// then walks TR cells of one (c, f) column and produces the eight nodal values
// of each 2 x 2 x 2 cell from its eight corners (the upper four become the
// lower four of the next cell) and seven coefficient loads, which are
// contiguous along f for a warp.  All index work (ghost node of an even-sized
// dimension, hole in front of it, coarse-first positions) is per thread and per
// dimension, outside the cell loop.
#pragma once

namespace restore3d {

typedef long long i64;

constexpr int TR = 4, TC = 8, TF = 32, NT = 256;
constexpr int CF = TF + 1, CC = TC + 1, CR = TR + 1;

template <typename T> struct Params {
  int n[3], nc[3], np[3]; // fine / coarse sizes, padded nodal sizes 2*nc-1
  i64 sc[3];              // dense coarse strides
  i64 sb[3];              // coefficient array strides (coarse-first layout)
  i64 so[3];              // dense nodal output strides
  const T *ratio[3];      // level-l ratio tables
  int tiles_c, tiles_f;
};

template <typename T> __device__ __forceinline__ T lerp_ref(T v0, T v1, T t) {
  T r = v0 + v0 * t * (T)-1;
  r = r + t * v1;
  return r;
}

// padded nodal index -> actual nodal index; -1: hole / out of range
__device__ __forceinline__ int src_index(int j, int n, int np) {
  if (j < 0 || j >= np)
    return -1;
  if ((n & 1) == 0) {
    if (j == n)
      return n - 1;
    if (j == n - 1)
      return -1;
  }
  return j;
}

template <typename T>
__global__ void __launch_bounds__(NT)
restore3d_kernel(const Params<T> P, const T *__restrict__ coarse, const T *__restrict__ coef,
       T *__restrict__ out) {
  __shared__ T s[CR * CC * CF];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  int bid = blockIdx.x;
  const int tf = bid % P.tiles_f;
  bid /= P.tiles_f;
  const int tc = bid % P.tiles_c;
  const int tr = bid / P.tiles_c;
  const int kr0 = tr * TR, kc0 = tc * TC, kf0 = tf * TF;
  const int rr = P.nc[0], cc = P.nc[1], ff = P.nc[2];
  // corner values (zero outside the coarse box)
  for (int e = tid; e < CR * CC * CF; e += NT) {
    const int a = e / (CC * CF), rem = e - a * (CC * CF);
    const int b = rem / CF, c = rem - b * CF;
    const int kr = kr0 + a, kc = kc0 + b, kf = kf0 + c;
    T v = (T)0;
    if (kr < rr && kc < cc && kf < ff)
      v = coarse[(i64)kr * P.sc[0] + (i64)kc * P.sc[1] + (i64)kf * P.sc[2]];
    s[e] = v;
  }
  __syncthreads();
  const int kc = kc0 + ty, kf = kf0 + tx;
  if (kc >= cc || kf >= ff)
    return;
  // f and c: actual indices of the even / odd node of this column, coefficient
  // positions, interpolation ratios
  const int fe = src_index(2 * kf, P.n[2], P.np[2]), fo = src_index(2 * kf + 1, P.n[2], P.np[2]);
  const int ce = src_index(2 * kc, P.n[1], P.np[1]), co = src_index(2 * kc + 1, P.n[1], P.np[1]);
  const T rf = fo >= 0 ? P.ratio[2][2 * kf] : (T)0;
  const T rc = co >= 0 ? P.ratio[1][2 * kc] : (T)0;
  const i64 b_fe = (i64)kf * P.sb[2], b_fo = (i64)(ff + kf) * P.sb[2];
  const i64 b_ce = (i64)kc * P.sb[1], b_co = (i64)(cc + kc) * P.sb[1];
  const i64 o_fe = (i64)fe * P.so[2], o_fo = (i64)fo * P.so[2];
  const i64 o_ce = (i64)ce * P.so[1], o_co = (i64)co * P.so[1];
  const T *sp = s + ty * CF + tx;
  T c00 = sp[0], c01 = sp[1], c10 = sp[CF], c11 = sp[CF + 1]; // lower plane (c, f corners)
#pragma unroll
  for (int lr = 0; lr < TR; lr++) {
    const int kr = kr0 + lr;
    if (kr >= rr)
      break;
    const T *hp = sp + (lr + 1) * (CC * CF);
    const T h00 = hp[0], h01 = hp[1], h10 = hp[CF], h11 = hp[CF + 1];
    const int re = src_index(2 * kr, P.n[0], P.np[0]), ro = src_index(2 * kr + 1, P.n[0], P.np[0]);
    const i64 b_re = (i64)kr * P.sb[0], b_ro = (i64)(rr + kr) * P.sb[0];
    // coefficient loads first (independent), then the arithmetic
    T q001 = (T)0, q010 = (T)0, q011 = (T)0, q100 = (T)0, q101 = (T)0, q110 = (T)0, q111 = (T)0;
    if (fo >= 0)
      q001 = coef[b_re + b_ce + b_fo];
    if (co >= 0) {
      q010 = coef[b_re + b_co + b_fe];
      if (fo >= 0)
        q011 = coef[b_re + b_co + b_fo];
    }
    if (ro >= 0) {
      q100 = coef[b_ro + b_ce + b_fe];
      if (fo >= 0)
        q101 = coef[b_ro + b_ce + b_fo];
      if (co >= 0) {
        q110 = coef[b_ro + b_co + b_fe];
        if (fo >= 0)
          q111 = coef[b_ro + b_co + b_fo];
      }
    }
    // interpolation along f at the four (r, c) corners
    const T lf0 = lerp_ref(c00, c01, rf), lf1 = lerp_ref(c10, c11, rf);
    const T hf0 = lerp_ref(h00, h01, rf), hf1 = lerp_ref(h10, h11, rf);
    T *o_e = out + (i64)re * P.so[0];
    // even plane
    o_e[o_ce + o_fe] = c00;
    if (fo >= 0)
      o_e[o_ce + o_fo] = q001 + lf0;
    if (co >= 0) {
      o_e[o_co + o_fe] = q010 + lerp_ref(c00, c10, rc);
      if (fo >= 0)
        o_e[o_co + o_fo] = q011 + lerp_ref(lf0, lf1, rc);
    }
    if (ro >= 0) {
      const T rt = P.ratio[0][2 * kr];
      T *o_o = out + (i64)ro * P.so[0];
      o_o[o_ce + o_fe] = q100 + lerp_ref(c00, h00, rt);
      if (fo >= 0)
        o_o[o_ce + o_fo] = q101 + lerp_ref(lf0, hf0, rt);
      if (co >= 0) {
        o_o[o_co + o_fe] = q110 + lerp_ref(lerp_ref(c00, c10, rc), lerp_ref(h00, h10, rc), rt);
        if (fo >= 0)
          o_o[o_co + o_fo] = q111 + lerp_ref(lerp_ref(lf0, lf1, rc), lerp_ref(hf0, hf1, rc), rt);
      }
    }
    c00 = h00;
    c01 = h01;
    c10 = h10;
    c11 = h11;
  }
}

} // namespace restore3d
