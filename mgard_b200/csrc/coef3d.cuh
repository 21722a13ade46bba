// Tiled 3-D coefficient kernel (sm_100a): multilevel coefficients of level l
// (nodal value minus the multilinear interpolant of the coarse neighbours),
// written to their final place in the coarse-first layout, plus the coarse
// nodes as the next level's dense input.
//
// Replaces, for D == 3, the reference's CopyND + GpkReo3D
// (Coefficient/GridProcessingKernel3D.hpp:21-1229; lerp: GPKFunctor.h:13-26)
// with the same arithmetic in the same order (interpolate along f, then c,
// then r; coefficient = value - interpolant), so results stay bit-identical.
//
// Mirror image of restore3d.cuh without shared memory: a thread walks TR cells
// of one (c, f) column and loads the nodes it needs itself - its own eight per
// cell plus the even nodes of the next row / column / plane, which its
// neighbours load as their own at about the same time (L1 hits).  The even
// plane 2kr+2 of one cell is the plane 2kr of the next, so a cell costs eleven
// independent loads, seven coefficients and one coarse node; each of the eight
// output streams is contiguous along f for a warp.
#pragma once

namespace coef3d {

typedef long long i64;

constexpr int TR = 8, TC = 8, TF = 32, NT = 256;

template <typename T> struct Params {
  int n[3], nc[3], np[3]; // fine / coarse sizes, padded nodal sizes 2*nc-1
  i64 si[3];              // dense nodal input strides
  i64 sb[3];              // coefficient array strides (coarse-first layout)
  i64 sc[3];              // dense coarse output strides
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

// ABSMAX (fp32 only): max |x| over the input as a by-product -- every node is the
// "own" node of exactly one thread -- so that a relative L-infinity bound needs no
// separate pass over the input (norm_calculator, NormCalculator.hpp:13-83).
template <typename T, bool ABSMAX>
__global__ void __launch_bounds__(NT)
coef3d_kernel(const Params<T> P, const T *__restrict__ in, T *__restrict__ coef,
       T *__restrict__ coarse, unsigned *__restrict__ absmax_bits) {
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  int bid = blockIdx.x;
  const int tf = bid % P.tiles_f;
  bid /= P.tiles_f;
  const int tc = bid % P.tiles_c;
  const int tr = bid / P.tiles_c;
  const int kr0 = tr * TR, kc = tc * TC + ty, kf = tf * TF + tx;
  const int rr = P.nc[0], cc = P.nc[1], ff = P.nc[2];
  if (kc >= cc || kf >= ff)
    return;
  // f and c: offsets of the even node, the odd node and the next even node of
  // this column (an odd node exists only together with the even node after it)
  const int fe = src_index(2 * kf, P.n[2], P.np[2]), fo = src_index(2 * kf + 1, P.n[2], P.np[2]);
  const int ce = src_index(2 * kc, P.n[1], P.np[1]), co = src_index(2 * kc + 1, P.n[1], P.np[1]);
  const i64 i_fe = (i64)fe * P.si[2], i_fo = (i64)fo * P.si[2],
            i_f2 = (i64)src_index(2 * kf + 2, P.n[2], P.np[2]) * P.si[2];
  const i64 i_ce = (i64)ce * P.si[1], i_co = (i64)co * P.si[1],
            i_c2 = (i64)src_index(2 * kc + 2, P.n[1], P.np[1]) * P.si[1];
  const T rf = fo >= 0 ? P.ratio[2][2 * kf] : (T)0;
  const T rc = co >= 0 ? P.ratio[1][2 * kc] : (T)0;
  const i64 b_fe = (i64)kf * P.sb[2], b_fo = (i64)(ff + kf) * P.sb[2];
  const i64 b_ce = (i64)kc * P.sb[1], b_co = (i64)(cc + kc) * P.sb[1];
  const i64 c_col = (i64)kc * P.sc[1] + (i64)kf * P.sc[2];
  // the seven nodes of an even plane this thread uses: 00 01 02 / 10 11 / 20 22
  struct Even {
    T v00, v01, v02, v10, v11, v20, v22;
  };
  auto load_even = [&](int jr) {
    Even e;
    e.v00 = e.v01 = e.v02 = e.v10 = e.v11 = e.v20 = e.v22 = (T)0;
    const T *p = in + (i64)jr * P.si[0];
    e.v00 = p[i_ce + i_fe];
    if (fo >= 0) {
      e.v01 = p[i_ce + i_fo];
      e.v02 = p[i_ce + i_f2];
    }
    if (co >= 0) {
      e.v10 = p[i_co + i_fe];
      e.v20 = p[i_c2 + i_fe];
      if (fo >= 0) {
        e.v11 = p[i_co + i_fo];
        e.v22 = p[i_c2 + i_f2];
      }
    }
    return e;
  };
  Even lo = load_even(src_index(2 * kr0, P.n[0], P.np[0]));
  float amax = 0.0f;
#pragma unroll
  for (int lr = 0; lr < TR; lr++) {
    const int kr = kr0 + lr;
    if (kr >= rr)
      break;
    const int ro = src_index(2 * kr + 1, P.n[0], P.np[0]);
    // loads of this cell first: odd plane (own four nodes) and the next even plane
    T o00 = (T)0, o01 = (T)0, o10 = (T)0, o11 = (T)0;
    Even hi = lo;
    if (ro >= 0) {
      const T *p = in + (i64)ro * P.si[0];
      o00 = p[i_ce + i_fe];
      if (fo >= 0)
        o01 = p[i_ce + i_fo];
      if (co >= 0) {
        o10 = p[i_co + i_fe];
        if (fo >= 0)
          o11 = p[i_co + i_fo];
      }
    }
    // the next even plane exists whenever there is another coarse plane (for an
    // even-sized dimension the odd plane in front of the ghost plane is a hole)
    if (kr + 1 < rr)
      hi = load_even(src_index(2 * kr + 2, P.n[0], P.np[0]));
    if (ABSMAX)
      amax = fmaxf(fmaxf(fmaxf(amax, fabsf((float)lo.v00)), fmaxf(fabsf((float)lo.v01), fabsf((float)lo.v10))),
                   fmaxf(fmaxf(fabsf((float)lo.v11), fabsf((float)o00)),
                         fmaxf(fabsf((float)o01), fmaxf(fabsf((float)o10), fabsf((float)o11)))));
    coarse[(i64)kr * P.sc[0] + c_col] = lo.v00;
    const i64 b_re = (i64)kr * P.sb[0], b_ro = (i64)(rr + kr) * P.sb[0];
    const T lf0 = lerp_ref(lo.v00, lo.v02, rf), lf1 = lerp_ref(lo.v20, lo.v22, rf);
    if (fo >= 0)
      coef[b_re + b_ce + b_fo] = lo.v01 - lf0;
    if (co >= 0) {
      coef[b_re + b_co + b_fe] = lo.v10 - lerp_ref(lo.v00, lo.v20, rc);
      if (fo >= 0)
        coef[b_re + b_co + b_fo] = lo.v11 - lerp_ref(lf0, lf1, rc);
    }
    if (ro >= 0) {
      const T rt = P.ratio[0][2 * kr];
      const T hf0 = lerp_ref(hi.v00, hi.v02, rf), hf1 = lerp_ref(hi.v20, hi.v22, rf);
      coef[b_ro + b_ce + b_fe] = o00 - lerp_ref(lo.v00, hi.v00, rt);
      if (fo >= 0)
        coef[b_ro + b_ce + b_fo] = o01 - lerp_ref(lf0, hf0, rt);
      if (co >= 0) {
        coef[b_ro + b_co + b_fe] =
            o10 - lerp_ref(lerp_ref(lo.v00, lo.v20, rc), lerp_ref(hi.v00, hi.v20, rc), rt);
        if (fo >= 0)
          coef[b_ro + b_co + b_fo] =
              o11 - lerp_ref(lerp_ref(lf0, lf1, rc), lerp_ref(hf0, hf1, rc), rt);
      }
    }
    lo = hi;
  }
  if (ABSMAX) {
    // non-negative floats order like their bit patterns
    const unsigned mask = __activemask();
    const unsigned m = __reduce_max_sync(mask, __float_as_uint(amax));
    if ((threadIdx.x & 31) == __ffs(mask) - 1)
      atomicMax(absmax_bits, m);
  }
}

} // namespace coef3d
