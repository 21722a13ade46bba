// Multilevel decomposition / recomposition kernels (sm_100a).
//
// What is computed (bit-exact with the reference's non-FMA arithmetic; this
// file must be compiled with -fmad=false):
//   coefficients   reference GpkReo3D / GpkReo<D>   (Coefficient/GridProcessingKernel3D.hpp:21-1229,
//                  lerp: Coefficient/GPKFunctor.h:13-26)
//   restore        reference GpkRev3D / GpkRev<D>   (GridProcessingKernel3D.hpp:1231-2400)
//   mass x restr.  reference Lpk{1,2,3}Reo3D        (Correction/LinearProcessingKernel3D.hpp:27-1090,
//                  mass_trans: Correction/LPKFunctor.h:47-66)
//   tridiagonal    reference Ipk{1,2,3}Reo3D        (Correction/IterativeProcessingKernel3D.hpp,
//                  Correction/IPKFunctor.h:14-51)
//   level loop     reference multi_dimension::decompose / recompose
//                  (DataRefactoring.hpp:25-177,180-317)
//
// How it is laid out here (not the reference's way): the input field is never
// modified.  Level l reads a DENSE box (the input itself for the finest level,
// a dense coarse buffer afterwards), writes its coefficients straight to their
// final position in the output array (coarse-first layout along every
// dimension) and the coarse nodes to the next dense buffer; no CopyND /
// in-place permutation passes.  The kernels in this file are dimension-generic
// (D = 1..5) through a (rows x fastest-dim) mapping: a thread block owns rows,
// threads sweep the contiguous dimension, so every global access is coalesced.
// D == 3 takes the tiled kernels of coef3d.cuh / masstrans3d.cuh / restore3d.cuh.
#include <algorithm>
#include <cstdint>
#include <limits>

#include "coef3d.cuh"
#include "masstrans3d.cuh"
#include "thomas_tma.cuh"
#include "thomas_stream.cuh"
#include "restore3d.cuh"
#include "plan.h"

// quantize.cu
int mgb_quantize_range(mgb_plan *plan, const void *d_coef, int ebtype, double tol, double s,
                       double norm, uint16_t *d_sym, uint32_t *d_hist,
                       unsigned long long *d_ocount, uint64_t *d_oidx, int64_t *d_oval,
                       uint64_t outlier_cap, uint64_t first, uint64_t count, int zero,
                       unsigned max_blocks, cudaStream_t st, const void *d_qtab);
int mgb_prepare_quantizers(mgb_plan *plan, int ebtype, double tol, double s, int src, const void *d_src,
                           uint64_t n_total, uint64_t nsub, void *d_qtab, double *d_norm_out,
                           cudaStream_t st);

namespace {

typedef long long i64;

struct Geom {
  int D;          // number of dims
  int n[5];       // fine level shape
  int nc[5];      // coarse level shape
  i64 sa[5];      // strides of array A (meaning per kernel)
  i64 sb[5];      // strides of array B
  i64 sc[5];      // strides of array C
  const void *t0[5]; // per-dim table 0 (ratio / coefficient table)
  int axis, zero_block;
  unsigned rows;  // product of n[0..D-2] (or kernel specific)
  int grid_rows;  // slower-dim indices come from blockIdx.{x,y,z} (no division)
};

template <typename T> __device__ __forceinline__ T lerp_ref(T v0, T v1, T t) {
  // GPKFunctor.h:13-26, non-FMA branch
  T r = v0 + v0 * t * (T)-1;
  r = r + t * v1;
  return r;
}

// ---------------------------------------------------------------------------
// Coefficient kernel.  A = dense nodal input (level box), B = output array
// (full-array strides, coarse-first layout), C = dense coarse output.
// DS = number of slower dims (D-1); MASK bit (DS-1-d) set <=> slower dim d is
// an odd (new) node, so bit 0 is the fastest of the slower dims.
// ---------------------------------------------------------------------------
constexpr int popc(int x) { return x == 0 ? 0 : (x & 1) + popc(x >> 1); }

template <typename T, int DS, int MASK> struct Corner {
  static constexpr int K = popc(MASK);
  // interpolate the value at column index `col` of the corner rows
  // (fastest-dim first reduction order: slower dims from fastest to slowest)
  template <typename F>
  static __device__ __forceinline__ T eval(const i64 (&dstride)[5],
                                           const T (&rat)[5], F rowval) {
    T vals[1 << K];
#pragma unroll
    for (int c = 0; c < (1 << K); c++) {
      i64 off = 0;
      int bit = 0;
#pragma unroll
      for (int d = DS - 1; d >= 0; d--) {
        if (MASK & (1 << (DS - 1 - d))) {
          off += ((c >> bit) & 1) ? dstride[d] : -dstride[d];
          bit++;
        }
      }
      vals[c] = rowval(off);
    }
    int bit = 0;
#pragma unroll
    for (int d = DS - 1; d >= 0; d--) {
      if (MASK & (1 << (DS - 1 - d))) {
#pragma unroll
        for (int m = 0; m < (1 << (K - 1 - bit)); m++)
          vals[m] = lerp_ref(vals[2 * m], vals[2 * m + 1], rat[d]);
        bit++;
      }
    }
    return vals[0];
  }
};

template <typename T, int DS, int MASK>
__device__ __forceinline__ void
coef_row(const Geom &g, const T *__restrict__ in_row, T *__restrict__ out_row,
         T *__restrict__ coarse_row, const i64 (&dstride)[5],
         const T (&rat)[5]) {
  const int D = DS + 1;
  const int nf = g.n[D - 1], ncf = g.nc[D - 1];
  const bool f_even_n = (nf & 1) == 0;
  const T *__restrict__ ratio_f = (const T *)g.t0[D - 1];
  const i64 osf = g.sb[D - 1];
  const int npairs = (nf + 1) >> 1;
  for (int j = threadIdx.x; j < npairs; j += blockDim.x) {
    const int f0 = 2 * j, f1 = 2 * j + 1;
    // even node f0 -> position j
    {
      T v = in_row[f0];
      if (MASK == 0) {
        coarse_row[j] = v;
      } else {
        T it = Corner<T, DS, MASK>::eval(
            dstride, rat, [&](i64 off) { return in_row[off + f0]; });
        out_row[(i64)j * osf] = v - it;
      }
    }
    if (f1 < nf) {
      T v = in_row[f1];
      if (f_even_n && f1 == nf - 1) {
        // ghost node: behaves as the last coarse node along f
        if (MASK == 0) {
          coarse_row[ncf - 1] = v;
        } else {
          T it = Corner<T, DS, MASK>::eval(
              dstride, rat, [&](i64 off) { return in_row[off + f1]; });
          out_row[(i64)(ncf - 1) * osf] = v - it;
        }
      } else {
        const T rf = ratio_f[f0];
        T it = Corner<T, DS, MASK>::eval(dstride, rat, [&](i64 off) {
          return lerp_ref(in_row[off + f0], in_row[off + f0 + 2], rf);
        });
        out_row[(i64)(ncf + j) * osf] = v - it;
      }
    }
  }
}

template <typename T, int DS, int MASK> struct CoefDispatch {
  static __device__ __forceinline__ void
  run(int mask, const Geom &g, const T *in_row, T *out_row, T *coarse_row,
      const i64 (&dstride)[5], const T (&rat)[5]) {
    if (mask == MASK)
      coef_row<T, DS, MASK>(g, in_row, out_row, coarse_row, dstride, rat);
    else
      CoefDispatch<T, DS, MASK - 1>::run(mask, g, in_row, out_row, coarse_row,
                                         dstride, rat);
  }
};
template <typename T, int DS> struct CoefDispatch<T, DS, -1> {
  static __device__ __forceinline__ void run(int, const Geom &, const T *, T *,
                                             T *, const i64 (&)[5],
                                             const T (&)[5]) {}
};

// Indices of the slower dims of the row a thread works on.  grid_rows: the
// fastest slower dim comes from blockIdx.x / threadIdx.y, the next from
// blockIdx.y and the rest from blockIdx.z, so that the common 2-D / 3-D case
// needs no integer division; otherwise the linear row index is decomposed.
template <int DS>
__device__ __forceinline__ bool row_indices(const Geom &g, unsigned (&idx)[5]) {
  if (DS == 0)
    return blockIdx.x == 0 && threadIdx.y == 0;
  if (g.grid_rows) {
    idx[DS - 1] = blockIdx.x * blockDim.y + threadIdx.y;
    if (idx[DS - 1] >= (unsigned)g.n[DS - 1])
      return false;
    if (DS >= 2)
      idx[DS - 2] = blockIdx.y;
    if (DS >= 3) {
      unsigned rem = blockIdx.z;
#pragma unroll
      for (int d = DS - 3; d >= 0; d--) {
        unsigned nd = (unsigned)g.n[d];
        idx[d] = rem % nd;
        rem /= nd;
      }
    }
    return true;
  }
  unsigned row = blockIdx.x * blockDim.y + threadIdx.y;
  if (row >= g.rows)
    return false;
#pragma unroll
  for (int d = DS - 1; d >= 0; d--) {
    unsigned nd = (unsigned)g.n[d];
    idx[d] = row % nd;
    row /= nd;
  }
  return true;
}

template <typename T, int DS>
__global__ void __launch_bounds__(512) coef_kernel(const Geom g, const T *__restrict__ in,
                                                   T *__restrict__ out,
                                                   T *__restrict__ coarse) {
  unsigned idx[5];
  if (!row_indices<DS>(g, idx))
    return;
  i64 in_off = 0, out_off = 0, c_off = 0;
  int mask = 0;
  i64 dstride[5];
  T rat[5];
#pragma unroll
  for (int d = DS - 1; d >= 0; d--) {
    unsigned nd = (unsigned)g.n[d];
    unsigned i = idx[d];
    bool ghost = ((nd & 1) == 0) && (i == nd - 1);
    bool odd = (i & 1) && !ghost;
    unsigned p = odd ? g.nc[d] + (i >> 1) : (ghost ? g.nc[d] - 1 : (i >> 1));
    in_off += (i64)i * g.sa[d];
    out_off += (i64)p * g.sb[d];
    c_off += (i64)p * g.sc[d]; // only meaningful when mask == 0
    dstride[d] = g.sa[d];
    rat[d] = (T)0;
    if (odd) {
      mask |= 1 << (DS - 1 - d);
      rat[d] = ((const T *)g.t0[d])[i - 1];
    }
  }
  CoefDispatch<T, DS, (1 << DS) - 1>::run(mask, g, in + in_off, out + out_off,
                                          coarse + c_off, dstride, rat);
}

// ---------------------------------------------------------------------------
// Restore kernel (recomposition).  A = dense coarse input, B = coefficient
// array (full strides, coarse-first layout), C = dense nodal output.
// ---------------------------------------------------------------------------
template <typename T, int DS, int MASK>
__device__ __forceinline__ void
restore_row(const Geom &g, const T *__restrict__ crow, const T *__restrict__ coef_row_,
            T *__restrict__ out_row, const i64 (&dstride)[5],
            const i64 (&dlo)[5], const T (&rat)[5]) {
  const int D = DS + 1;
  const int nf = g.n[D - 1], ncf = g.nc[D - 1];
  const bool f_even_n = (nf & 1) == 0;
  const T *__restrict__ ratio_f = (const T *)g.t0[D - 1];
  const i64 bsf = g.sb[D - 1];
  const int npairs = (nf + 1) >> 1;
  // corner rows: offsets relative to crow: for odd dims "low" coarse index is
  // already folded into crow; +stride selects the high neighbour.
  auto corner = [&](auto rowval) {
    constexpr int K = popc(MASK);
    T vals[1 << K];
#pragma unroll
    for (int c = 0; c < (1 << K); c++) {
      i64 off = 0;
      int bit = 0;
#pragma unroll
      for (int d = DS - 1; d >= 0; d--) {
        if (MASK & (1 << (DS - 1 - d))) {
          off += ((c >> bit) & 1) ? dstride[d] : 0;
          bit++;
        }
      }
      vals[c] = rowval(off);
    }
    int bit = 0;
#pragma unroll
    for (int d = DS - 1; d >= 0; d--) {
      if (MASK & (1 << (DS - 1 - d))) {
#pragma unroll
        for (int m = 0; m < (1 << (K - 1 - bit)); m++)
          vals[m] = lerp_ref(vals[2 * m], vals[2 * m + 1], rat[d]);
        bit++;
      }
    }
    return vals[0];
  };
  for (int j = threadIdx.x; j < npairs; j += blockDim.x) {
    const int f0 = 2 * j, f1 = 2 * j + 1;
    {
      T r;
      if (MASK == 0) {
        r = crow[j];
      } else {
        T it = corner([&](i64 off) { return crow[off + j]; });
        r = coef_row_[(i64)j * bsf] + it;
      }
      out_row[f0] = r;
    }
    if (f1 < nf) {
      T r;
      if (f_even_n && f1 == nf - 1) {
        if (MASK == 0) {
          r = crow[ncf - 1];
        } else {
          T it = corner([&](i64 off) { return crow[off + ncf - 1]; });
          r = coef_row_[(i64)(ncf - 1) * bsf] + it;
        }
      } else {
        const T rf = ratio_f[f0];
        T it = corner([&](i64 off) {
          return lerp_ref(crow[off + j], crow[off + j + 1], rf);
        });
        r = coef_row_[(i64)(ncf + j) * bsf] + it;
      }
      out_row[f1] = r;
    }
  }
}

template <typename T, int DS, int MASK> struct RestoreDispatch {
  static __device__ __forceinline__ void
  run(int mask, const Geom &g, const T *crow, const T *coef_row_, T *out_row,
      const i64 (&dstride)[5], const i64 (&dlo)[5], const T (&rat)[5]) {
    if (mask == MASK)
      restore_row<T, DS, MASK>(g, crow, coef_row_, out_row, dstride, dlo, rat);
    else
      RestoreDispatch<T, DS, MASK - 1>::run(mask, g, crow, coef_row_, out_row,
                                            dstride, dlo, rat);
  }
};
template <typename T, int DS> struct RestoreDispatch<T, DS, -1> {
  static __device__ __forceinline__ void run(int, const Geom &, const T *,
                                             const T *, T *, const i64 (&)[5],
                                             const i64 (&)[5], const T (&)[5]) {}
};

template <typename T, int DS>
__global__ void __launch_bounds__(512) restore_kernel(const Geom g, const T *__restrict__ coarse,
                                                      const T *__restrict__ coef,
                                                      T *__restrict__ out) {
  unsigned idx[5];
  if (!row_indices<DS>(g, idx))
    return;
  i64 c_off = 0, b_off = 0, o_off = 0;
  int mask = 0;
  i64 dstride[5], dlo[5];
  T rat[5];
#pragma unroll
  for (int d = DS - 1; d >= 0; d--) {
    unsigned nd = (unsigned)g.n[d];
    unsigned i = idx[d];
    bool ghost = ((nd & 1) == 0) && (i == nd - 1);
    bool odd = (i & 1) && !ghost;
    unsigned p = odd ? g.nc[d] + (i >> 1) : (ghost ? g.nc[d] - 1 : (i >> 1));
    unsigned clo = ghost ? g.nc[d] - 1 : (i >> 1); // low coarse neighbour
    c_off += (i64)clo * g.sa[d];
    b_off += (i64)p * g.sb[d];
    o_off += (i64)i * g.sc[d];
    dstride[d] = g.sa[d];
    dlo[d] = 0;
    rat[d] = (T)0;
    if (odd) {
      mask |= 1 << (DS - 1 - d);
      rat[d] = ((const T *)g.t0[d])[i - 1];
    }
  }
  RestoreDispatch<T, DS, (1 << DS) - 1>::run(mask, g, coarse + c_off,
                                             coef + b_off, out + o_off,
                                             dstride, dlo, rat);
}

// ---------------------------------------------------------------------------
// Mass matrix x restriction along one axis on the coarse-first layout.
// A = input (strided, shape n[] with axis already-reduced dims holding nc),
// B = dense output.  t0[axis] = 9-row coefficient table [9][nc_axis]:
// h1/6,(h1+h2)/3,h2/6,(h2+h3)/3,h3/6,(h3+h4)/3,h4/6,r1,r4 (LPKFunctor.h:47-66).
// g.n[]  : input shape;  output shape = n[] with n[axis] -> nc[axis].
// zero_block: treat input elements with all indices < nc[] as zero
// (LinearProcessingKernel3D.hpp:99-133).
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(512) mass_trans_kernel(const Geom g, const T *__restrict__ in,
                                                         T *__restrict__ out) {
  const int D = g.D;
  const int a = g.axis;
  unsigned row = blockIdx.x * blockDim.y + threadIdx.y;
  if (row >= g.rows)
    return;
  // rows enumerate output indices of dims 0..D-2 (slower dims)
  i64 in_off = 0, out_off = 0;
  bool in_block = g.zero_block != 0;
  int ia = 0; // index along axis if axis is a slower dim
  unsigned rem = row;
  for (int d = D - 2; d >= 0; d--) {
    unsigned nd = (unsigned)(d == a ? g.nc[d] : g.n[d]);
    unsigned i = rem % nd;
    rem /= nd;
    out_off += (i64)i * g.sb[d];
    if (d == a) {
      ia = (int)i;
    } else {
      in_off += (i64)i * g.sa[d];
      if ((int)i >= g.nc[d])
        in_block = false;
    }
  }
  const T *__restrict__ tab = (const T *)g.t0[a];
  const int na = g.n[a], nca = g.nc[a];
  const int ncoef = na - nca;
  const int nf_out = (a == D - 1) ? nca : g.n[D - 1];
  const i64 sa_a = g.sa[a];
  const i64 sa_f = g.sa[D - 1];
  for (int f = threadIdx.x; f < nf_out; f += blockDim.x) {
    const int i = (a == D - 1) ? f : ia;
    const T *__restrict__ p = in + in_off + ((a == D - 1) ? 0 : (i64)f * sa_f);
    bool zb = in_block && (a == D - 1 || f < g.nc[D - 1]);
    T va = (T)0, vb = (T)0, vc = (T)0, vd = (T)0, ve = (T)0;
    if (!zb) {
      vc = p[(i64)i * sa_a];
      if (i >= 1)
        va = p[(i64)(i - 1) * sa_a];
      if (i + 1 < nca)
        ve = p[(i64)(i + 1) * sa_a];
    }
    if (i >= 1 && i - 1 < ncoef)
      vb = p[(i64)(nca + i - 1) * sa_a];
    if (i < ncoef)
      vd = p[(i64)(nca + i) * sa_a];
    const T c16 = tab[i], c13 = tab[nca + i], c26 = tab[2 * nca + i],
            c23 = tab[3 * nca + i], c36 = tab[4 * nca + i],
            c34 = tab[5 * nca + i], c46 = tab[6 * nca + i],
            r1 = tab[7 * nca + i], r4 = tab[8 * nca + i];
    T tb = va * c16 + vb * c13 + vc * c26;
    T tc = vb * c26 + vc * c23 + vd * c36;
    T td = vc * c36 + vd * c34 + ve * c46;
    tc += tb * r1 + td * r4;
    out[out_off + (i64)f * g.sb[D - 1]] = tc;
  }
}

// ---------------------------------------------------------------------------
// Thomas solve along a strided axis of a dense array viewed as
// (outer, n, inner), inner > 1: one thread per line, coalesced across inner.
// mode: 0 plain, 1 add result to `acc`, 2 subtract result from `acc`
// (AddND / SubtractND of DataRefactoring.hpp:99,241 fused into the last solve).
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) thomas_strided_kernel(T *__restrict__ x, int n, i64 inner,
                                                             i64 lines,
                                                             const T *__restrict__ fw,
                                                             const T *__restrict__ am,
                                                             const T *__restrict__ bm,
                                                             T *__restrict__ acc, int mode) {
  i64 line = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= lines)
    return;
  i64 o = line / inner, in = line - o * inner;
  const i64 off0 = o * (i64)n * inner + in;
  T *p = x + off0;
  T prev = (T)0;
  constexpr int U = 8;
  int i = 0;
  for (; i + U <= n; i += U) {
    T v[U];
#pragma unroll
    for (int k = 0; k < U; k++)
      v[k] = p[(i64)(i + k) * inner];
#pragma unroll
    for (int k = 0; k < U; k++) {
      prev = v[k] - prev * fw[i + k];
      p[(i64)(i + k) * inner] = prev;
    }
  }
  for (; i < n; i++) {
    prev = p[(i64)i * inner] - prev * fw[i];
    p[(i64)i * inner] = prev;
  }
  prev = (T)0;
  i = n - 1;
  if (mode == 0) {
    for (; i - U + 1 >= 0; i -= U) {
      T v[U];
#pragma unroll
      for (int k = 0; k < U; k++)
        v[k] = p[(i64)(i - k) * inner];
#pragma unroll
      for (int k = 0; k < U; k++) {
        prev = (v[k] - am[i - k + 1] * prev) / bm[i - k + 1];
        p[(i64)(i - k) * inner] = prev;
      }
    }
    for (; i >= 0; i--) {
      prev = (p[(i64)i * inner] - am[i + 1] * prev) / bm[i + 1];
      p[(i64)i * inner] = prev;
    }
  } else {
    // last solve of a correction: the result is added to / subtracted from acc
    T *q = acc + off0;
    for (; i - U + 1 >= 0; i -= U) {
      T v[U], a[U];
#pragma unroll
      for (int k = 0; k < U; k++) {
        v[k] = p[(i64)(i - k) * inner];
        a[k] = q[(i64)(i - k) * inner];
      }
#pragma unroll
      for (int k = 0; k < U; k++) {
        prev = (v[k] - am[i - k + 1] * prev) / bm[i - k + 1];
        q[(i64)(i - k) * inner] = mode == 1 ? a[k] + prev : a[k] - prev;
      }
    }
    for (; i >= 0; i--) {
      prev = (p[(i64)i * inner] - am[i + 1] * prev) / bm[i + 1];
      const T a = q[(i64)i * inner];
      q[(i64)i * inner] = mode == 1 ? a + prev : a - prev;
    }
  }
}

// Thomas solve along the contiguous axis: a warp owns 32 lines and streams
// 32-column tiles through shared memory (transpose so that global accesses
// stay coalesced while each lane walks its own line).
template <typename T>
__global__ void __launch_bounds__(128) thomas_contig_kernel(T *__restrict__ x, int n, i64 lines,
                                                            const T *__restrict__ fw,
                                                            const T *__restrict__ am,
                                                            const T *__restrict__ bm) {
  __shared__ T tile[4][32][33];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  i64 line0 = ((i64)blockIdx.x * 4 + w) * 32;
  if (line0 >= lines)
    return;
  int nl = (int)min((i64)32, lines - line0);
  T(*t)[33] = tile[w];
  T prev = (T)0;
  for (int c0 = 0; c0 < n; c0 += 32) {
    int ncol = min(32, n - c0);
    for (int r = 0; r < nl; r++)
      if (lane < ncol)
        t[r][lane] = x[(line0 + r) * n + c0 + lane];
    __syncwarp();
    if (lane < nl) {
      for (int c = 0; c < ncol; c++) {
        prev = t[lane][c] - prev * fw[c0 + c];
        t[lane][c] = prev;
      }
    }
    __syncwarp();
    for (int r = 0; r < nl; r++)
      if (lane < ncol)
        x[(line0 + r) * n + c0 + lane] = t[r][lane];
    __syncwarp();
  }
  prev = (T)0;
  for (int c1 = n; c1 > 0; c1 -= 32) {
    int c0 = max(0, c1 - 32);
    int ncol = c1 - c0;
    for (int r = 0; r < nl; r++)
      if (lane < ncol)
        t[r][lane] = x[(line0 + r) * n + c0 + lane];
    __syncwarp();
    if (lane < nl) {
      for (int c = ncol - 1; c >= 0; c--) {
        prev = (t[lane][c] - am[c0 + c + 1] * prev) / bm[c0 + c + 1];
        t[lane][c] = prev;
      }
    }
    __syncwarp();
    for (int r = 0; r < nl; r++)
      if (lane < ncol)
        x[(line0 + r) * n + c0 + lane] = t[r][lane];
    __syncwarp();
  }
}

// Thomas solve with the whole line set of a thread block staged in shared
// memory.  A block owns W lines of the (outer, n, inner) view: W consecutive
// lines along the contiguous axis (inner == 1) or W neighbouring columns of one
// outer slice (inner > 1).  Every line is brought in with asynchronous copies
// (all in flight at once, coalesced), the two sequential sweeps then run out
// of shared memory (the recurrence, not memory latency, bounds them), and the
// result goes back coalesced.  mode 1 / 2: the result is added to / subtracted
// from `acc` instead of being stored (AddND / SubtractND of
// DataRefactoring.hpp:99,241 fused into the last solve).
template <typename T, int W>
__global__ void __launch_bounds__(W)
thomas_smem_kernel(T *__restrict__ x, int n, i64 inner, i64 lines,
                   const T *__restrict__ fw, const T *__restrict__ am,
                   const T *__restrict__ bm, T *__restrict__ acc, int mode) {
  extern __shared__ __align__(16) unsigned char thomas_smem[];
  T *s = reinterpret_cast<T *>(thomas_smem);
  constexpr int P = W + 1;
  const int lane = threadIdx.x;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(s);
  bool mine;      // this thread owns a line
  i64 gbase;      // strided: element 0 of the thread's line; contiguous: of the block
  int nl = W;     // contiguous: lines of this block
  if (inner == 1) {
    const i64 line0 = (i64)blockIdx.x * W;
    nl = (int)min((i64)W, lines - line0);
    mine = lane < nl;
    gbase = line0 * n;
    // line t, elements i0 + lane: coalesced segments, transposed into s[i][t]
    for (int t = 0; t < nl; t++) {
      const T *g = x + gbase + (i64)t * n;
      unsigned d = sbase + (unsigned)((lane * P + t) * sizeof(T));
      for (int i = lane; i < n; i += W) {
        if (sizeof(T) == 4)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(g + i));
        else
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(g + i));
        d += (unsigned)(W * P * sizeof(T));
      }
    }
  } else {
    const i64 chunks = (inner + W - 1) / W;
    const i64 o = blockIdx.x / chunks;
    const i64 in0 = (blockIdx.x - o * chunks) * W;
    mine = in0 + lane < inner;
    gbase = o * (i64)n * inner + in0 + lane;
    if (mine) {
      const T *g = x + gbase;
      unsigned d = sbase + (unsigned)(lane * sizeof(T));
#pragma unroll 4
      for (int i = 0; i < n; i++) {
        if (sizeof(T) == 4)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(g));
        else
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(g));
        d += (unsigned)(P * sizeof(T));
        g += inner;
      }
    }
  }
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();
  if (mine) {
    T *c = s + lane;
    T prev = (T)0;
    int i = 0;
#pragma unroll 1
    for (; i + 8 <= n; i += 8) {
      T v[8], f[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        v[k] = c[(i + k) * P];
        f[k] = __ldg(fw + i + k);
      }
#pragma unroll
      for (int k = 0; k < 8; k++) {
        prev = v[k] - prev * f[k];
        c[(i + k) * P] = prev;
      }
    }
    for (; i < n; i++) {
      prev = c[i * P] - prev * __ldg(fw + i);
      c[i * P] = prev;
    }
    prev = (T)0;
    i = n - 1;
#pragma unroll 1
    for (; i >= 7; i -= 8) {
      T v[8], a[8], b[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        v[k] = c[(i - k) * P];
        a[k] = __ldg(am + i - k + 1);
        b[k] = __ldg(bm + i - k + 1);
      }
#pragma unroll
      for (int k = 0; k < 8; k++) {
        prev = (v[k] - a[k] * prev) / b[k];
        c[(i - k) * P] = prev;
      }
    }
    for (; i >= 0; i--) {
      prev = (c[i * P] - __ldg(am + i + 1) * prev) / __ldg(bm + i + 1);
      c[i * P] = prev;
    }
  }
  __syncthreads();
  if (inner == 1) {
    for (int t = 0; t < nl; t++) {
      const i64 g0 = gbase + (i64)t * n;
      const T *sp = s + lane * P + t;
      for (int i = lane; i < n; i += W) {
        const T r = *sp;
        sp += W * P;
        const i64 g = g0 + i;
        if (mode == 0)
          x[g] = r;
        else
          acc[g] = mode == 1 ? acc[g] + r : acc[g] - r;
      }
    }
  } else if (mine) {
    const T *sp = s + lane;
    i64 g = gbase;
    if (mode == 0) {
#pragma unroll 4
      for (int i = 0; i < n; i++, sp += P, g += inner)
        x[g] = *sp;
    } else if (mode == 1) {
#pragma unroll 4
      for (int i = 0; i < n; i++, sp += P, g += inner)
        acc[g] = acc[g] + *sp;
    } else {
#pragma unroll 4
      for (int i = 0; i < n; i++, sp += P, g += inner)
        acc[g] = acc[g] - *sp;
    }
  }
}

// All solves of a small correction in ONE launch: when the whole coarse array
// fits in shared memory a single thread block loads it, runs the Thomas solves
// along every dimension (fastest first, one thread per line, a barrier between
// dimensions) and adds / subtracts the result to / from the coarse nodes.  The
// levels this applies to are latency bound, so one launch instead of D (+ the
// traffic between them) is what counts.
struct SmallSolve {
  int D;
  int n[5];            // coarse shape
  const void *fw[5], *am[5], *bm[5];
};
template <typename T>
__global__ void __launch_bounds__(1024)
thomas_small_kernel(const SmallSolve g, const T *__restrict__ w, T *__restrict__ acc, int mode,
                    int total) {
  extern __shared__ __align__(16) unsigned char small_smem[];
  T *s = reinterpret_cast<T *>(small_smem);
  for (int i = threadIdx.x; i < total; i += blockDim.x)
    s[i] = w[i];
  __syncthreads();
  int inner = 1;
  for (int a = g.D - 1; a >= 0; a--) {
    const int n = g.n[a];
    const int lines = total / n;
    const T *__restrict__ fw = (const T *)g.fw[a];
    const T *__restrict__ am = (const T *)g.am[a];
    const T *__restrict__ bm = (const T *)g.bm[a];
    for (int line = threadIdx.x; line < lines; line += blockDim.x) {
      const int o = line / inner, in = line - o * inner;
      T *p = s + (size_t)o * n * inner + in;
      T prev = (T)0;
      for (int i = 0; i < n; i++) {
        prev = p[i * inner] - prev * __ldg(fw + i);
        p[i * inner] = prev;
      }
      prev = (T)0;
      for (int i = n - 1; i >= 0; i--) {
        prev = (p[i * inner] - __ldg(am + i + 1) * prev) / __ldg(bm + i + 1);
        p[i * inner] = prev;
      }
    }
    inner *= n;
    __syncthreads();
  }
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const T r = s[i];
    acc[i] = mode == 1 ? acc[i] + r : acc[i] - r;
  }
}

// acc[i] += sign * w[i] on dense arrays (AddND / SubtractND)
template <typename T>
__global__ void axpy_kernel(T *__restrict__ acc, const T *__restrict__ w, i64 n, int subtract) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  i64 stride = (i64)gridDim.x * blockDim.x;
  for (; i < n; i += stride)
    acc[i] = subtract ? acc[i] - w[i] : acc[i] + w[i];
}

// copy between a dense box and a strided box (level-0 coarse nodes)
template <typename T>
__global__ void box_copy_kernel(const Geom g, const T *__restrict__ src, T *__restrict__ dst,
                                i64 total) {
  i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total)
    return;
  i64 rem = idx, so = 0, dof = 0;
  for (int d = g.D - 1; d >= 0; d--) {
    i64 i = rem % g.n[d];
    rem /= g.n[d];
    so += i * g.sa[d];
    dof += i * g.sb[d];
  }
  dst[dof] = src[so];
}

// ------------------------------ host drivers -------------------------------

void dense_strides(const uint64_t *shape, int D, i64 *s) {
  i64 acc = 1;
  for (int d = D - 1; d >= 0; d--) {
    s[d] = acc;
    acc *= (i64)shape[d];
  }
}

void row_launch_dims(int nf_threads, unsigned rows, dim3 &grid, dim3 &block) {
  // threads along the fastest dimension: a multiple of 32 that wastes the fewest
  // slots in the last sweep (sizes 2^k + 1 give 2^(k-1) + 1 pairs: a power of two
  // would leave its second sweep almost empty)
  int bx = 32, best = 1 << 30;
  for (int cand = 32; cand <= 512; cand += 32) {
    int sweeps = (nf_threads + cand - 1) / cand;
    int waste = sweeps * cand - nf_threads;
    // prefer few sweeps when the waste ties
    int cost = waste * 8 + sweeps;
    if (cost < best) {
      best = cost;
      bx = cand;
    }
    if (cand >= nf_threads)
      break;
  }
  int by = std::max(1, 256 / bx);
  block = dim3(bx, by, 1);
  grid = dim3((rows + by - 1) / by, 1, 1);
}

template <typename T>
void fill_tables(const mgb_plan *p, int l, Geom &g, bool masstrans) {
  for (int d = 0; d < p->D; d++) {
    const mgb_dim_tables &m = p->tab[l][d];
    g.t0[d] = p->dtab(masstrans ? m.mt : m.ratio);
  }
}

template <typename T>
void thomas_all(mgb_plan *p, int l, T *w, T *acc, int mode, cudaStream_t st);
template <typename T>
void axpy(T *acc, const T *w, i64 n, int subtract, cudaStream_t st);

// correction = Thomas_{f,c,r..}( MassTrans_{f,c,r..}( coefficient function ) )
// (CalcCorrection3D.hpp:30-196), result dense with the coarse shape of level l-1
// in *result (either d_wA or d_wB).
template <typename T>
int correction(mgb_plan *p, int l, const T *coef, T **result, T *acc, int mode,
               cudaStream_t st) {
  const int D = p->D;
  i64 full[5];
  dense_strides(p->shape, D, full);
  const T *src = coef;
  T *bufs[2] = {(T *)p->d_wA, (T *)p->d_wB};
  int which = 0;
  uint64_t cur_shape[5];
  for (int d = 0; d < D; d++)
    cur_shape[d] = p->lshape[l][d];
  i64 src_stride[5];
  for (int d = 0; d < D; d++)
    src_stride[d] = full[d];
  for (int a = D - 1; a >= 0; a--) {
    Geom g = {};
    g.D = D;
    g.axis = a;
    g.zero_block = (a == D - 1);
    for (int d = 0; d < D; d++) {
      g.n[d] = (int)cur_shape[d];
      g.nc[d] = (int)p->lshape[l - 1][d];
      g.sa[d] = src_stride[d];
    }
    g.n[a] = (int)p->lshape[l][a];
    uint64_t out_shape[5];
    for (int d = 0; d < D; d++)
      out_shape[d] = d == a ? p->lshape[l - 1][d] : cur_shape[d];
    dense_strides(out_shape, D, g.sb);
    fill_tables<T>(p, l, g, true);
    unsigned rows = 1;
    for (int d = 0; d < D - 1; d++)
      rows *= (unsigned)out_shape[d];
    g.rows = rows;
    dim3 grid, block;
    row_launch_dims((int)out_shape[D - 1], rows, grid, block);
    T *dst = bufs[which];
    MGB_LAUNCH(MGB_K_MASSTRANS, st, (mass_trans_kernel<T><<<grid, block, 0, st>>>(g, src, dst)));
    src = dst;
    which ^= 1;
    for (int d = 0; d < D; d++) {
      cur_shape[d] = out_shape[d];
      src_stride[d] = g.sb[d];
    }
  }
  T *w = (T *)src; // dense, coarse shape
  thomas_all<T>(p, l, w, acc, mode, st);
  *result = w;
  return MGB_SUCCESS;
}

// grid over the slower dims without a linear row index when the sizes fit
template <int DS> void row_grid(Geom &g, dim3 &grid, const dim3 &block) {
  g.grid_rows = 0;
  if (DS == 0)
    return;
  unsigned long long gy = DS >= 2 ? (unsigned long long)g.n[DS - 2] : 1, gz = 1;
  for (int d = 0; d + 3 <= DS; d++)
    gz *= (unsigned long long)g.n[d];
  if (gy > 65535 || gz > 65535)
    return;
  g.grid_rows = 1;
  grid = dim3((g.n[DS - 1] + block.y - 1) / block.y, (unsigned)gy, (unsigned)gz);
}
template <typename T, int DS>
void launch_coef(const Geom &g0, const T *in, T *out, T *coarse, cudaStream_t st) {
  Geom g = g0;
  dim3 grid, block;
  row_launch_dims((g.n[DS] + 1) / 2, g.rows, grid, block);
  row_grid<DS>(g, grid, block);
  MGB_LAUNCH(MGB_K_COEF, st, (coef_kernel<T, DS><<<grid, block, 0, st>>>(g, in, out, coarse)));
}
template <typename T, int DS>
void launch_restore(const Geom &g0, const T *coarse, const T *coef, T *out,
                    cudaStream_t st) {
  Geom g = g0;
  dim3 grid, block;
  row_launch_dims((g.n[DS] + 1) / 2, g.rows, grid, block);
  row_grid<DS>(g, grid, block);
  MGB_LAUNCH(MGB_K_RESTORE, st, (restore_kernel<T, DS><<<grid, block, 0, st>>>(g, coarse, coef, out)));
}

template <typename T>
void axpy(T *acc, const T *w, i64 n, int subtract, cudaStream_t st) {
  unsigned blocks = (unsigned)std::min<i64>((n + 255) / 256, 148 * 16);
  MGB_LAUNCH(MGB_K_AXPY, st, (axpy_kernel<T><<<blocks, 256, 0, st>>>(acc, w, n, subtract)));
}

// contiguous axis through the TMA-staged kernel (thomas_tma.cuh): G lines per warp, as
// many warps as shared memory holds
template <typename T, int G>
bool launch_thomas_tma(T *w, int n, i64 lines, int nwarp, const T *fw, const T *am, const T *bm, T *acc,
                       int mode, cudaStream_t st) {
  static bool configured[64] = {};
  if (mgb_first_use_on_device(configured)) {
    if (cudaFuncSetAttribute(thomas_tma_kernel<T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
  }
  const size_t stage = (((size_t)G * n * sizeof(T)) + 127) & ~(size_t)127;
  const size_t tables = (((size_t)4 * n * sizeof(T)) + 127) & ~(size_t)127;
  const size_t smem = (((size_t)nwarp * 8 + 127) & ~(size_t)127) + tables + (size_t)nwarp * stage;
  const i64 ngroups = (lines + G - 1) / G;
  const unsigned blocks = (unsigned)std::min<i64>(148, (ngroups + nwarp - 1) / nwarp);
  MGB_LAUNCH(MGB_K_THOMAS_CONTIG, st,
             (thomas_tma_kernel<T, G><<<blocks, nwarp * 32, smem, st>>>(w, n, (long long)lines, fw, am, bm, acc,
                                                                      mode)));
  return cudaGetLastError() == cudaSuccess;
}

template <typename T, int W>
bool launch_thomas_smem(T *w, int n, i64 inner, i64 outer, const T *fw, const T *am,
                        const T *bm, T *acc, int mode, cudaStream_t st) {
  const size_t smem = (size_t)n * (W + 1) * sizeof(T);
  static bool configured[64] = {};
  if (mgb_first_use_on_device(configured)) {
    if (cudaFuncSetAttribute(thomas_smem_kernel<T, W>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return false;
    cudaFuncSetAttribute(thomas_smem_kernel<T, W>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
  }
  i64 blocks = inner == 1 ? (outer + W - 1) / W : outer * ((inner + W - 1) / W);
  const i64 lines = outer * inner;
  MGB_LAUNCH(inner == 1 ? MGB_K_THOMAS_CONTIG : MGB_K_THOMAS_STRIDED, st,
             (thomas_smem_kernel<T, W><<<(unsigned)blocks, W, smem, st>>>(w, n, inner, lines, fw, am,
                                                                          bm, acc, mode)));
  return true;
}

// D == 3: coefficients (coef3d.cuh) and load vector (masstrans3d.cuh)
template <typename T>
void launch_coef3d(mgb_plan *p, int l, const T *in, T *coef, T *coarse, cudaStream_t st,
                   unsigned *absmax = nullptr) {
  coef3d::Params<T> P;
  i64 full[5], dc[5], dn[5];
  dense_strides(p->shape, 3, full);
  dense_strides(p->lshape[l - 1], 3, dc);
  dense_strides(p->lshape[l], 3, dn);
  for (int d = 0; d < 3; d++) {
    P.n[d] = (int)p->lshape[l][d];
    P.nc[d] = (int)p->lshape[l - 1][d];
    P.np[d] = 2 * P.nc[d] - 1;
    P.si[d] = dn[d];
    P.sb[d] = full[d];
    P.sc[d] = dc[d];
    P.ratio[d] = (const T *)p->dtab(p->tab[l][d].ratio);
  }
  const int tiles_r = (P.nc[0] + coef3d::TR - 1) / coef3d::TR;
  P.tiles_c = (P.nc[1] + coef3d::TC - 1) / coef3d::TC;
  P.tiles_f = (P.nc[2] + coef3d::TF - 1) / coef3d::TF;
  unsigned grid = (unsigned)(tiles_r * P.tiles_c * P.tiles_f);
  if (absmax && sizeof(T) == 4)
    MGB_LAUNCH(MGB_K_COEF, st,
               (coef3d::coef3d_kernel<T, true><<<grid, coef3d::NT, 0, st>>>(P, in, coef, coarse, absmax)));
  else
    MGB_LAUNCH(MGB_K_COEF, st,
               (coef3d::coef3d_kernel<T, false><<<grid, coef3d::NT, 0, st>>>(P, in, coef, coarse, nullptr)));
}
// warp-per-tile formulation with TMA-staged rows (masstrans3d.cuh, second half); false:
// the level is too small to fill the GPU with independent warps (the block formulation
// splits finer) or the staging does not fit
template <typename T>
bool launch_masstrans3d_warp(mgb_plan *p, int l, const T *coef, T *w_out, cudaStream_t st) {
  namespace mt = masstrans3d;
  mt::WParams<T> P;
  i64 full[5], dc[5];
  dense_strides(p->shape, 3, full);
  dense_strides(p->lshape[l - 1], 3, dc);
  for (int d = 0; d < 3; d++) {
    P.n[d] = (int)p->lshape[l][d];
    P.nc[d] = (int)p->lshape[l - 1][d];
    P.sin[d] = full[d];
    P.sw[d] = dc[d];
    P.mt[d] = (const T *)p->dtab(p->tab[l][d].mt);
  }
  if (full[2] != 1 || sizeof(T) != 4)
    return false;
  P.total = (i64)p->N;
  P.ctiles = (P.nc[1] + mt::TC - 1) / mt::TC;
  P.ftiles = (P.nc[2] + mt::TF - 1) / mt::TF;
  const int tiles = P.ctiles * P.ftiles;
  // 16 warps per SM; every r segment re-reads three warm-up planes, so segments of at
  // least 8 coarse planes, at most 32 (the r constants of a segment sit in shared memory)
  const long long want = 148ll * 2 * mt::W_NW;
  const int max_segs = std::max(1, P.nc[0] / 8);
  if ((long long)tiles * max_segs < want)
    return false;
  int rsegs = (int)std::min<long long>(max_segs, (want * 4 + tiles - 1) / tiles);
  int per = std::min(32, (P.nc[0] + rsegs - 1) / rsegs);
  P.per = per;
  P.rsegs = (P.nc[0] + per - 1) / per;
  const size_t smem = mt::warp_smem_base<T>() + (size_t)mt::W_NW * mt::warp_smem_bytes<T>(per);
  if (smem > 113 * 1024)
    return false;
  static bool configured[64] = {};
  if (mgb_first_use_on_device(configured)) {
    if (cudaFuncSetAttribute(mt::masstrans3d_warp_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             113 * 1024) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
  }
  const long long warps = (long long)tiles * P.rsegs;
  const unsigned grid = (unsigned)((warps + mt::W_NW - 1) / mt::W_NW);
  MGB_LAUNCH(MGB_K_MASSTRANS, st,
             (mt::masstrans3d_warp_kernel<T><<<grid, mt::W_NW * 32, smem, st>>>(P, coef, w_out)));
  return true;
}

template <typename T>
void launch_masstrans3d(mgb_plan *p, int l, const T *coef, T *w_out, cudaStream_t st) {
  static const bool v1 = getenv("MGB_MASSTRANS_V1") != nullptr;
  if (!v1 && launch_masstrans3d_warp<T>(p, l, coef, w_out, st))
    return;
  masstrans3d::Params<T> P;
  i64 full[5], dc[5];
  dense_strides(p->shape, 3, full);
  dense_strides(p->lshape[l - 1], 3, dc);
  for (int d = 0; d < 3; d++) {
    P.n[d] = (int)p->lshape[l][d];
    P.nc[d] = (int)p->lshape[l - 1][d];
    P.sin[d] = full[d];
    P.sw[d] = dc[d];
    P.mt[d] = (const T *)p->dtab(p->tab[l][d].mt);
  }
  P.ctiles = (P.nc[1] + masstrans3d::TC - 1) / masstrans3d::TC;
  P.ftiles = (P.nc[2] + masstrans3d::TF - 1) / masstrans3d::TF;
  const int tiles = P.ctiles * P.ftiles;
  // ~16 blocks per SM; every r segment re-reads two warm-up plane pairs, so keep
  // segments >= 8 coarse planes unless the level is too small to fill the GPU
  int rsegs = (148 * 16 + tiles - 1) / tiles;
  int seg_cap = std::max(1, P.nc[0] / 8);
  if (tiles * seg_cap < 148 * 3)
    seg_cap = std::max(1, P.nc[0] / 2);
  P.rsegs = std::max(1, std::min(rsegs, seg_cap));
  unsigned grid = (unsigned)(tiles * P.rsegs);
  MGB_LAUNCH(MGB_K_MASSTRANS, st,
             (masstrans3d::masstrans3d_kernel<T><<<grid, masstrans3d::NT, 0, st>>>(P, coef, w_out)));
}

// D == 3: tiled restore (restore3d.cuh) instead of the row-based restore_kernel
template <typename T>
void launch_restore3d(mgb_plan *p, int l, const T *coarse, const T *coef, T *out,
                      cudaStream_t st) {
  restore3d::Params<T> P;
  i64 full[5], dc[5], dn[5];
  dense_strides(p->shape, 3, full);
  dense_strides(p->lshape[l - 1], 3, dc);
  dense_strides(p->lshape[l], 3, dn);
  for (int d = 0; d < 3; d++) {
    P.n[d] = (int)p->lshape[l][d];
    P.nc[d] = (int)p->lshape[l - 1][d];
    P.np[d] = 2 * P.nc[d] - 1;
    P.sc[d] = dc[d];
    P.sb[d] = full[d];
    P.so[d] = dn[d];
    P.ratio[d] = (const T *)p->dtab(p->tab[l][d].ratio);
  }
  const int tiles_r = (P.nc[0] + restore3d::TR - 1) / restore3d::TR;
  P.tiles_c = (P.nc[1] + restore3d::TC - 1) / restore3d::TC;
  P.tiles_f = (P.nc[2] + restore3d::TF - 1) / restore3d::TF;
  unsigned grid = (unsigned)(tiles_r * P.tiles_c * P.tiles_f);
  MGB_LAUNCH(MGB_K_RESTORE, st,
             (restore3d::restore3d_kernel<T><<<grid, restore3d::NT, 0, st>>>(P, coarse, coef, out)));
}

// Thomas solves (all dims) in place on the dense coarse-shaped array w; the
// last solve adds (mode 1) / subtracts (mode 2) its result to / from `acc`
// when acc != nullptr, otherwise w holds the correction afterwards.
template <typename T>
void thomas_all(mgb_plan *p, int l, T *w, T *acc, int mode, cudaStream_t st) {
  const int D = p->D;
  // small correction: every solve and the final add / subtract in one launch
  {
    const uint64_t total = mgb_level_elems(p, l - 1);
    const size_t smem = total * sizeof(T);
    // (measured on B200: one block wins up to ~17^3 nodes; at 33^3 three multi-block
    // launches are faster than one block walking 3 x 1089 lines)
    if (acc && total <= 8192 && smem <= 200 * 1024) {
      SmallSolve g;
      g.D = D;
      for (int d = 0; d < D; d++) {
        const mgb_dim_tables &m = p->tab[l - 1][d];
        g.n[d] = (int)p->lshape[l - 1][d];
        g.fw[d] = p->dtab(m.fw);
        g.am[d] = p->dtab(m.am);
        g.bm[d] = p->dtab(m.bm);
      }
      static bool configured[64] = {};
      if (mgb_first_use_on_device(configured)) {
        cudaFuncSetAttribute(thomas_small_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             200 * 1024);
      }
      int threads = 1024;
      MGB_LAUNCH(MGB_K_THOMAS_CONTIG, st,
                 (thomas_small_kernel<T><<<1, threads, smem, st>>>(g, w, acc, mode, (int)total)));
      return;
    }
  }
  const size_t smem_max = 227 * 1024, smem_sm = 228 * 1024 - 1024;
  for (int a = D - 1; a >= 0; a--) {
    const mgb_dim_tables &m = p->tab[l - 1][a];
    const T *fw = (const T *)p->dtab(m.fw);
    const T *am = (const T *)p->dtab(m.am);
    const T *bm = (const T *)p->dtab(m.bm);
    int n = (int)p->lshape[l - 1][a];
    i64 inner = 1, outer = 1;
    for (int d = a + 1; d < D; d++)
      inner *= (i64)p->lshape[l - 1][d];
    for (int d = 0; d < a; d++)
      outer *= (i64)p->lshape[l - 1][d];
    const bool lastax = (a == 0);
    T *accp = (lastax && acc) ? acc : nullptr;
    const int md = accp ? mode : 0;
    if (inner > 1) {
      // strided axis: one thread per line, coalesced across the inner dimension
      // (all lines resident at once hides more latency than staging here)
      i64 lines = outer * inner;
      unsigned blocks = (unsigned)((lines + 127) / 128);
      MGB_LAUNCH(MGB_K_THOMAS_STRIDED, st,
                 (thomas_strided_kernel<T><<<blocks, 128, 0, st>>>(w, n, inner, lines, fw, am, bm,
                                                                  accp, md)));
      continue;
    }
    // contiguous axis, plain solve, enough lines for 128-line blocks on every SM: the
    // streaming kernel (thomas_stream.cuh)
    static const int stream_mode = getenv("MGB_THOMAS_STREAM") ? atoi(getenv("MGB_THOMAS_STREAM")) : 1;
    // (lines of up to ~400 nodes fit in shared memory by the hundred per SM: the resident
    // formulation below is faster there - 0.077 against 0.097 ms at 257 x 257 lines of 257)
    if (stream_mode && md == 0 && outer >= 148ll * thomas_stream::LINES && (n >= 400 || stream_mode == 2)) {
      const unsigned blocks = (unsigned)((outer + thomas_stream::LINES - 1) / thomas_stream::LINES);
      MGB_LAUNCH(MGB_K_THOMAS_CONTIG, st,
                 (thomas_stream::thomas_stream_kernel<T><<<blocks, thomas_stream::LINES, 0, st>>>(
                     w, n, (long long)outer, fw, am, bm)));
      continue;
    }
    // contiguous axis: whole lines staged in shared memory by the TMA engine when the
    // natural pitch is free of bank conflicts (odd n: the 2^k + 1 sizes)
    if (getenv("MGB_NO_TMA_THOMAS") == nullptr) {
      const int banks = sizeof(T) == 4 ? 32 : 16;
      int gcd = n, b = banks;
      while (b) {
        const int t = gcd % b;
        gcd = b;
        b = t;
      }
      // lines per warp (G) against warps per block (S): a warp instruction costs the same
      // with 4 or 32 active lanes, so small G multiplies the issue work; large G leaves
      // few lines in flight to hide the dependent chain.  Crude cycle model per SM:
      const size_t avail = 220 * 1024;
      int bestG = 0, bestS = 0;
      double best_cost = 0;
      const int Gs[4] = {32, 16, 8, 4};
      const double chain = sizeof(T) == 4 ? 40.0 : 110.0, instr = sizeof(T) == 4 ? 22.0 : 40.0;
      for (int k = 0; k < 4; k++) {
        const size_t stage = (((size_t)Gs[k] * n * sizeof(T)) + 127) & ~(size_t)127;
        const size_t tables = (((size_t)4 * n * sizeof(T)) + 127) & ~(size_t)127;
        if (avail < 256 + tables + stage)
          continue;
        int S = (int)std::min<size_t>(32, (avail - 256 - tables) / stage);
        // no more warps than there are groups for one block per SM
        const i64 ngroups = (outer + Gs[k] - 1) / Gs[k];
        S = (int)std::min<i64>(S, std::max<i64>(1, (ngroups + 147) / 148));
        if (S < 1)
          continue;
        const double groups_sm = (double)ngroups / 148.0;
        const double t_chain = std::ceil(groups_sm / S) * n * chain;
        const double t_issue = groups_sm * n * instr / 4.0;
        const double cost = std::max(t_chain, t_issue);
        if (!bestG || cost < best_cost) {
          bestG = Gs[k];
          bestS = S;
          best_cost = cost;
        }
      }
      bool ok = false;
      if (gcd <= 2 && bestG) {
        if (bestG == 32)
          ok = launch_thomas_tma<T, 32>(w, n, outer, bestS, fw, am, bm, accp, md, st);
        else if (bestG == 16)
          ok = launch_thomas_tma<T, 16>(w, n, outer, bestS, fw, am, bm, accp, md, st);
        else if (bestG == 8)
          ok = launch_thomas_tma<T, 8>(w, n, outer, bestS, fw, am, bm, accp, md, st);
        else
          ok = launch_thomas_tma<T, 4>(w, n, outer, bestS, fw, am, bm, accp, md, st);
      }
      if (ok)
        continue;
    }
    int bestW = 0;
    i64 best = 0;
    const int Ws[3] = {32, 16, 8};
    for (int k = 0; k < 3; k++) {
      size_t sm = (size_t)n * (Ws[k] + 1) * sizeof(T);
      if (sm > smem_max)
        continue;
      i64 nb = std::min<i64>(32, (i64)(smem_sm / (sm + 1024)));
      i64 eff = std::min<i64>(Ws[k], outer);
      if (nb * eff > best) {
        best = nb * eff;
        bestW = Ws[k];
      }
    }
    bool done = false;
    if (bestW == 32)
      done = launch_thomas_smem<T, 32>(w, n, inner, outer, fw, am, bm, accp, md, st);
    else if (bestW == 16)
      done = launch_thomas_smem<T, 16>(w, n, inner, outer, fw, am, bm, accp, md, st);
    else if (bestW == 8)
      done = launch_thomas_smem<T, 8>(w, n, inner, outer, fw, am, bm, accp, md, st);
    if (done)
      continue;
    {
      i64 lines = outer;
      unsigned blocks = (unsigned)((lines + 127) / 128);
      MGB_LAUNCH(MGB_K_THOMAS_CONTIG, st,
                 (thomas_contig_kernel<T><<<blocks, 128, 0, st>>>(w, n, lines, fw, am, bm)));
    }
    if (accp)
      axpy<T>(accp, w, (i64)mgb_level_elems(p, l - 1), mode == 2, st);
  }
}

// the tiled 3-D kernels keep per-plane offsets in 32 bits
static inline bool tiled3d(const mgb_plan *p) {
  return p->D == 3 && !p->force_generic && p->shape[1] * p->shape[2] < (1ull << 31);
}

template <typename T>
int decompose_t(mgb_plan *p, const T *d_in, T *d_out, cudaStream_t st) {
  int rc = mgb_plan_ensure_workspace(p);
  if (rc)
    return rc;
  const int D = p->D;
  i64 full[5];
  dense_strides(p->shape, D, full);
  const T *cur = d_in;
  T *cbuf = (T *)p->d_cbuf;
  for (int l = p->L; l >= 1; l--) {
    Geom g = {};
    g.D = D;
    unsigned rows = 1;
    for (int d = 0; d < D; d++) {
      g.n[d] = (int)p->lshape[l][d];
      g.nc[d] = (int)p->lshape[l - 1][d];
      g.sb[d] = full[d];
      if (d < D - 1)
        rows *= (unsigned)g.n[d];
    }
    g.rows = rows;
    dense_strides(p->lshape[l], D, g.sa);
    dense_strides(p->lshape[l - 1], D, g.sc);
    fill_tables<T>(p, l, g, false);
    T *coarse = cbuf + p->cbuf_off[l - 1];
    if (tiled3d(p)) {
      T *w = (T *)p->d_wA;
      const bool with_norm = l == p->L && p->fused_norm.armed && sizeof(T) == 4;
      launch_coef3d<T>(p, l, cur, d_out, coarse, st, with_norm ? p->d_absmax : nullptr);
      if (with_norm) {
        // max |x| -> norm -> quantizer table, all in device memory (read by both
        // quantizer launches; the host sees the norm with the block size at the end)
        rc = mgb_prepare_quantizers(p, MGB_REL, p->fused_norm.tol, p->fused_norm.s, 0, p->d_absmax, p->N, 0,
                                    p->d_qtab, (double *)(p->d_scalars + 9), st);
        if (rc)
          return rc;
      }
      launch_masstrans3d<T>(p, l, d_out, w, st);
      thomas_all<T>(p, l, w, coarse, 1, st);
      if (l == p->L && p->early_q.armed && (void *)d_out == (void *)p->d_coef) {
        // planes r >= coarse size hold level-L coefficients only (final).  Quantize
        // them now, two blocks per SM, while the coarse levels - a chain of small
        // latency-bound launches - use a fraction of the machine.  Start at a
        // multiple of 8 elements so that the 128-bit path applies.
        const uint64_t plane = (uint64_t)p->shape[1] * p->shape[2];
        const uint64_t first = (p->lshape[l - 1][0] * plane + 7) / 8 * 8;
        if (first < p->N) {
          if (!p->side_q) {
            MGB_CUDA_CHECK(cudaStreamCreateWithFlags(&p->side_q, cudaStreamNonBlocking));
            MGB_CUDA_CHECK(cudaEventCreateWithFlags(&p->ev_qfork, cudaEventDisableTiming));
            MGB_CUDA_CHECK(cudaEventCreateWithFlags(&p->ev_qjoin, cudaEventDisableTiming));
          }
          MGB_CUDA_CHECK(cudaEventRecord(p->ev_qfork, st));
          MGB_CUDA_CHECK(cudaStreamWaitEvent(p->side_q, p->ev_qfork, 0));
          rc = mgb_quantize_range(p, p->d_coef, p->early_q.ebtype, p->early_q.tol, p->early_q.s,
                                  p->early_q.norm, p->d_sym, p->d_hist, p->d_scalars, p->d_oidx,
                                  p->d_oval, p->outlier_cap, first, p->N - first, 1, 148 * 2,
                                  p->side_q, p->early_q.d_qtab);
          if (rc)
            return rc;
          MGB_CUDA_CHECK(cudaEventRecord(p->ev_qjoin, p->side_q));
          p->early_q.first = first;
          p->early_q.done = true;
        }
      }
      cur = coarse;
      continue;
    }
    switch (D) {
    case 1: launch_coef<T, 0>(g, cur, d_out, coarse, st); break;
    case 2: launch_coef<T, 1>(g, cur, d_out, coarse, st); break;
    case 3: launch_coef<T, 2>(g, cur, d_out, coarse, st); break;
    case 4: launch_coef<T, 3>(g, cur, d_out, coarse, st); break;
    case 5: launch_coef<T, 4>(g, cur, d_out, coarse, st); break;
    }
    T *w = nullptr;
    rc = correction<T>(p, l, d_out, &w, coarse, 1, st);
    if (rc)
      return rc;
    cur = coarse;
  }
  // level-0 nodes into the corner of the output
  {
    Geom g = {};
    g.D = D;
    for (int d = 0; d < D; d++) {
      g.n[d] = (int)p->lshape[0][d];
      g.sb[d] = full[d];
    }
    dense_strides(p->lshape[0], D, g.sa);
    i64 total = (i64)mgb_level_elems(p, 0);
    MGB_LAUNCH(MGB_K_BOXCOPY, st,
               (box_copy_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, cur, d_out, total)));
  }
  MGB_CUDA_CHECK(cudaGetLastError());
  return MGB_SUCCESS;
}

template <typename T>
int recompose_t(mgb_plan *p, const T *d_in, T *d_out, cudaStream_t st) {
  int rc = mgb_plan_ensure_workspace(p);
  if (rc)
    return rc;
  const int D = p->D;
  i64 full[5];
  dense_strides(p->shape, D, full);
  T *cbuf = (T *)p->d_cbuf;
  if (p->L == 0) {
    MGB_CUDA_CHECK(cudaMemcpyAsync(d_out, d_in, p->N * sizeof(T),
                                   cudaMemcpyDeviceToDevice, st));
    return MGB_SUCCESS;
  }
  {
    Geom g = {};
    g.D = D;
    for (int d = 0; d < D; d++) {
      g.n[d] = (int)p->lshape[0][d];
      g.sa[d] = full[d];
    }
    dense_strides(p->lshape[0], D, g.sb);
    i64 total = (i64)mgb_level_elems(p, 0);
    MGB_LAUNCH(MGB_K_BOXCOPY, st,
               (box_copy_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
                   g, d_in, cbuf + p->cbuf_off[0], total)));
  }
  // The load vector of a level depends only on the coefficients, not on the level
  // recursion: the finest level's (throughput bound, most of the work) runs on
  // the caller's stream while the latency-bound chain of the coarse levels runs
  // next to it on a high-priority side stream.
  const bool fork = tiled3d(p) && p->L >= 2;
  cudaStream_t sc = st; // stream of the coarse-level chain
  if (fork) {
    if (!p->side) {
      int least = 0, greatest = 0;
      cudaDeviceGetStreamPriorityRange(&least, &greatest);
      MGB_CUDA_CHECK(cudaStreamCreateWithPriority(&p->side, cudaStreamNonBlocking, greatest));
      MGB_CUDA_CHECK(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
      MGB_CUDA_CHECK(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
    }
    sc = p->side;
    MGB_CUDA_CHECK(cudaEventRecord(p->ev_fork, st));
    MGB_CUDA_CHECK(cudaStreamWaitEvent(sc, p->ev_fork, 0));
    launch_masstrans3d<T>(p, p->L, d_in, (T *)p->d_wA, st);
  }
  for (int l = 1; l <= p->L; l++) {
    T *coarse = cbuf + p->cbuf_off[l - 1];
    T *w = nullptr;
    cudaStream_t sl = (fork && l < p->L) ? sc : st;
    if (tiled3d(p)) {
      if (fork && l == p->L) {
        // the finest level's solves do not need the coarse values either: they run
        // beside the coarse-level chain too, and only the subtraction waits for it
        // (same arithmetic as the subtraction fused into the last solve)
        static const bool early_solve = getenv("MGB_NO_EARLY_SOLVE") == nullptr;
        w = (T *)p->d_wA;
        if (early_solve)
          thomas_all<T>(p, l, w, (T *)nullptr, 0, st);
        MGB_CUDA_CHECK(cudaEventRecord(p->ev_join, sc));
        MGB_CUDA_CHECK(cudaStreamWaitEvent(st, p->ev_join, 0));
        if (early_solve)
          axpy<T>(coarse, w, (i64)mgb_level_elems(p, l - 1), 1, st);
        else
          thomas_all<T>(p, l, w, coarse, 2, sl);
      } else {
        w = (T *)(fork ? p->d_wB : p->d_wA);
        launch_masstrans3d<T>(p, l, d_in, w, sl);
        thomas_all<T>(p, l, w, coarse, 2, sl);
      }
    } else {
      rc = correction<T>(p, l, d_in, &w, coarse, 2, st);
    }
    if (rc)
      return rc;
    Geom g = {};
    g.D = D;
    unsigned rows = 1;
    for (int d = 0; d < D; d++) {
      g.n[d] = (int)p->lshape[l][d];
      g.nc[d] = (int)p->lshape[l - 1][d];
      g.sb[d] = full[d];
      if (d < D - 1)
        rows *= (unsigned)g.n[d];
    }
    g.rows = rows;
    dense_strides(p->lshape[l - 1], D, g.sa);
    dense_strides(p->lshape[l], D, g.sc);
    fill_tables<T>(p, l, g, false);
    T *dst = l == p->L ? d_out : cbuf + p->cbuf_off[l];
    if (tiled3d(p)) {
      launch_restore3d<T>(p, l, coarse, d_in, dst, sl);
      continue;
    }
    switch (D) {
    case 1: launch_restore<T, 0>(g, coarse, d_in, dst, st); break;
    case 2: launch_restore<T, 1>(g, coarse, d_in, dst, st); break;
    case 3: launch_restore<T, 2>(g, coarse, d_in, dst, st); break;
    case 4: launch_restore<T, 3>(g, coarse, d_in, dst, st); break;
    case 5: launch_restore<T, 4>(g, coarse, d_in, dst, st); break;
    }
  }
  MGB_CUDA_CHECK(cudaGetLastError());
  return MGB_SUCCESS;
}


// ---------------------------------------------------------------------------
// decomposition_type::SingleDim (reference DataRefactoring/SingleDimension/*.hpp):
// per level, one dimension after another (fastest first): coefficients by 1-D
// interpolation along that dimension into the coarse-first layout
// (CoefficientKernel.hpp:44-134), load vector of the coefficients
// (MassTransKernel.hpp:36-101) with the 13-argument mass_trans
// (LPKFunctor.h:14-66), Thomas solve with the level l-1 tables, added to the
// coarse part.  D <= 3 (the reference's own D >= 4 variant does not round-trip).
// Generic kernels over a strided 3-D box; this mode is not a bench line.
// ---------------------------------------------------------------------------
struct SdGeom {
  int n[3];      // box extents, with the active axis holding the COARSE size
  int ax;        // active axis (0..2, dimensions left-padded to 3)
  int nfine, nc; // fine / coarse size along the axis
  i64 sa[3];     // strides of array A
  i64 sb[3];     // strides of array B
};

__device__ __forceinline__ bool sd_index(const SdGeom &g, i64 t, int (&i)[3]) {
  const i64 total = (i64)g.n[0] * g.n[1] * g.n[2];
  if (t >= total)
    return false;
  i[2] = (int)(t % g.n[2]);
  t /= g.n[2];
  i[1] = (int)(t % g.n[1]);
  i[0] = (int)(t / g.n[1]);
  return true;
}

// A = interleaved source, B = destination in coarse-first layout
template <typename T>
__global__ void __launch_bounds__(256) sd_coef_kernel(const SdGeom g, const T *__restrict__ A, T *__restrict__ B,
                                                      const T *__restrict__ ratio) {
  int i[3];
  if (!sd_index(g, (i64)blockIdx.x * blockDim.x + threadIdx.x, i))
    return;
  const int j = i[g.ax], ncoef = g.nfine - g.nc;
  i64 a = 0, b = 0;
  for (int d = 0; d < 3; d++)
    if (d != g.ax) {
      a += i[d] * g.sa[d];
      b += i[d] * g.sb[d];
    }
  const i64 sa = g.sa[g.ax], sb = g.sb[g.ax];
  const int sj = j <= ncoef ? 2 * j : g.nfine - 1; // even nodes, then the last node of an even size
  const T left = A[a + sj * sa];
  B[b + j * sb] = left;
  if (j < ncoef) {
    const T mid = A[a + (2 * j + 1) * sa], right = A[a + (2 * j + 2) * sa];
    B[b + (g.nc + j) * sb] = mid - lerp_ref(left, right, ratio[2 * j]);
  }
}

// A = coarse-first source (coarse part and coefficients), B = interleaved destination
template <typename T>
__global__ void __launch_bounds__(256) sd_restore_kernel(const SdGeom g, const T *__restrict__ A, T *__restrict__ B,
                                                         const T *__restrict__ ratio) {
  int i[3];
  if (!sd_index(g, (i64)blockIdx.x * blockDim.x + threadIdx.x, i))
    return;
  const int j = i[g.ax], ncoef = g.nfine - g.nc;
  i64 a = 0, b = 0;
  for (int d = 0; d < 3; d++)
    if (d != g.ax) {
      a += i[d] * g.sa[d];
      b += i[d] * g.sb[d];
    }
  const i64 sa = g.sa[g.ax], sb = g.sb[g.ax];
  const int sj = j <= ncoef ? 2 * j : g.nfine - 1;
  const T left = A[a + j * sa];
  B[b + sj * sb] = left;
  if (j < ncoef) {
    const T right = A[a + (j + 1) * sa];
    B[b + (2 * j + 1) * sb] = A[a + (g.nc + j) * sa] + lerp_ref(left, right, ratio[2 * j]);
  }
}

// A = array holding the coefficients (coarse-first), B = dense load vector (coarse shape)
template <typename T>
__global__ void __launch_bounds__(256) sd_masstrans_kernel(const SdGeom g, const T *__restrict__ A,
                                                           T *__restrict__ B, const T *__restrict__ dist) {
  int i[3];
  if (!sd_index(g, (i64)blockIdx.x * blockDim.x + threadIdx.x, i))
    return;
  const int j = i[g.ax], ncoef = g.nfine - g.nc;
  i64 a = 0, bo = 0;
  for (int d = 0; d < 3; d++) {
    if (d != g.ax)
      a += i[d] * g.sa[d];
    bo += i[d] * g.sb[d];
  }
  const i64 sa = g.sa[g.ax];
  const T va = 0, vc = 0, ve = 0;
  T vb = 0, vd = 0;
  if (j > 0 && j < ncoef)
    vb = A[a + (g.nc + j - 1) * sa];
  if (j < ncoef)
    vd = A[a + (g.nc + j) * sa];
  T h1 = 0, h2 = 0, h3 = 0, h4 = 0;
  if (j > 0 && 2 * j < g.nfine - 1) {
    h1 = dist[2 * j - 2];
    h2 = dist[2 * j - 1];
  }
  if (2 * j < g.nfine - 1) {
    h3 = dist[2 * j];
    h4 = dist[2 * j + 1];
  }
  T r1 = 0, r4 = 0;
  if (h1 + h2 != 0)
    r1 = h1 / (h1 + h2);
  if (h3 + h4 != 0)
    r4 = h4 / (h3 + h4);
  const T tb = va * (h1 / 6) + vb * ((h1 + h2) / 3) + vc * (h2 / 6);
  T tc = vb * (h2 / 6) + vc * ((h2 + h3) / 3) + vd * (h3 / 6);
  const T td = vc * (h3 / 6) + vd * ((h3 + h4) / 3) + ve * (h4 / 6);
  tc += tb * r1 + td * r4;
  B[bo] = tc;
}

// B (strided) += / -= A (dense), or plain copies between a strided box and a dense one
template <typename T, int MODE>
__global__ void __launch_bounds__(256) sd_box_kernel(const SdGeom g, const T *__restrict__ A, T *__restrict__ B) {
  int i[3];
  if (!sd_index(g, (i64)blockIdx.x * blockDim.x + threadIdx.x, i))
    return;
  i64 a = 0, b = 0;
  for (int d = 0; d < 3; d++) {
    a += i[d] * g.sa[d];
    b += i[d] * g.sb[d];
  }
  if (MODE == 0)
    B[b] = A[a];
  else if (MODE == 1)
    B[b] = B[b] + A[a];
  else
    B[b] = B[b] - A[a];
}

struct SdLevelDim {
  int fine[3], nc, ncoef;
};

inline void sd_dense_strides(const int (&n)[3], i64 (&s)[3]) {
  s[2] = 1;
  s[1] = n[2];
  s[0] = (i64)n[1] * n[2];
}

template <typename T> int sd_prepare(mgb_plan *p, int (&full)[3], i64 (&sfull)[3], int &pad) {
  if (p->D > 3)
    return MGB_TOO_MANY_DIMS;
  int rc = mgb_plan_ensure_workspace(p);
  if (rc)
    return rc;
  if (!p->d_sd)
    MGB_CUDA_CHECK(cudaMalloc(&p->d_sd, p->N * p->tsize));
  pad = 3 - p->D;
  for (int d = 0; d < 3; d++)
    full[d] = d < pad ? 1 : (int)p->shape[d - pad];
  sd_dense_strides(full, sfull);
  return MGB_SUCCESS;
}

// load vector of the coefficients held in V (coarse-first along ax) -> solved
// correction in the dense buffer `corr`
template <typename T>
void sd_correction(mgb_plan *p, const T *V, const i64 (&sfull)[3], const int (&fine)[3], int ax, int pad, int l,
                   T *corr, cudaStream_t st) {
  const int dim = ax - pad;
  SdGeom g;
  g.ax = ax;
  g.nfine = fine[ax];
  g.nc = (int)p->lshape[l - 1][dim];
  for (int d = 0; d < 3; d++) {
    g.n[d] = d == ax ? g.nc : fine[d];
    g.sa[d] = sfull[d];
  }
  sd_dense_strides(g.n, g.sb);
  const i64 total = (i64)g.n[0] * g.n[1] * g.n[2];
  const T *dist = (const T *)p->dtab(p->tab[l][dim].dist);
  MGB_LAUNCH(MGB_K_MASSTRANS, st,
             (sd_masstrans_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, V, corr, dist)));
  // Thomas along ax with the level l-1 tables (Ipk kernels, CalcCorrection.hpp:44-93)
  i64 outer = 1, inner = 1;
  for (int d = 0; d < ax; d++)
    outer *= g.n[d];
  for (int d = ax + 1; d < 3; d++)
    inner *= g.n[d];
  const mgb_dim_tables &tc = p->tab[l - 1][dim];
  const i64 lines = outer * inner;
  MGB_LAUNCH(MGB_K_THOMAS_STRIDED, st,
             (thomas_strided_kernel<T><<<(unsigned)((lines + 127) / 128), 128, 0, st>>>(
                 corr, g.nc, inner, lines, (const T *)p->dtab(tc.fw), (const T *)p->dtab(tc.am),
                 (const T *)p->dtab(tc.bm), nullptr, 0)));
}

template <typename T>
void sd_accumulate(const i64 (&sfull)[3], const int (&coarse)[3], const T *corr, T *V, bool subtract,
                   cudaStream_t st) {
  SdGeom g;
  g.ax = 0;
  g.nfine = g.nc = 0;
  for (int d = 0; d < 3; d++) {
    g.n[d] = coarse[d];
    g.sb[d] = sfull[d];
  }
  sd_dense_strides(g.n, g.sa);
  const i64 total = (i64)g.n[0] * g.n[1] * g.n[2];
  if (subtract)
    MGB_LAUNCH(MGB_K_AXPY, st, (sd_box_kernel<T, 2><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, corr, V)));
  else
    MGB_LAUNCH(MGB_K_AXPY, st, (sd_box_kernel<T, 1><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, corr, V)));
}

// single_dimension::decompose (SingleDimension/DataRefactoring.hpp:25-108)
template <typename T> int decompose_single_t(mgb_plan *p, const T *d_in, T *d_out, cudaStream_t st) {
  int full[3], pad;
  i64 sfull[3];
  int rc = sd_prepare<T>(p, full, sfull, pad);
  if (rc)
    return rc;
  if (p->L == 0)
    MGB_CUDA_CHECK(cudaMemcpyAsync(d_out, d_in, p->N * sizeof(T), cudaMemcpyDeviceToDevice, st));
  T *tmp = (T *)p->d_sd, *corr = (T *)p->d_wB;
  bool first = true;
  for (int l = p->L; l > 0; l--) {
    for (int ax = 2; ax >= pad; ax--) {
      const int dim = ax - pad;
      int fine[3];
      for (int d = 0; d < 3; d++)
        fine[d] = d < pad ? 1 : (int)(d > ax ? p->lshape[l - 1][d - pad] : p->lshape[l][d - pad]);
      // source: the input itself for the very first step, else a dense copy of the box
      SdGeom g;
      g.ax = ax;
      g.nfine = fine[ax];
      g.nc = (int)p->lshape[l - 1][dim];
      for (int d = 0; d < 3; d++) {
        g.n[d] = fine[d];
        g.sb[d] = sfull[d];
      }
      const T *src = d_in;
      if (first) {
        for (int d = 0; d < 3; d++)
          g.sa[d] = sfull[d];
      } else {
        SdGeom c = g;
        for (int d = 0; d < 3; d++)
          c.sa[d] = sfull[d];
        sd_dense_strides(c.n, c.sb);
        const i64 tot = (i64)c.n[0] * c.n[1] * c.n[2];
        MGB_LAUNCH(MGB_K_BOXCOPY, st,
                   (sd_box_kernel<T, 0><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(c, d_out, tmp)));
        sd_dense_strides(g.n, g.sa);
        src = tmp;
      }
      first = false;
      g.n[ax] = g.nc;
      const i64 total = (i64)g.n[0] * g.n[1] * g.n[2];
      const T *ratio = (const T *)p->dtab(p->tab[l][dim].ratio);
      MGB_LAUNCH(MGB_K_COEF, st,
                 (sd_coef_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, src, d_out, ratio)));
      sd_correction<T>(p, d_out, sfull, fine, ax, pad, l, corr, st);
      int coarse[3];
      for (int d = 0; d < 3; d++)
        coarse[d] = d == ax ? g.nc : fine[d];
      sd_accumulate<T>(sfull, coarse, corr, d_out, false, st);
    }
  }
  MGB_CUDA_CHECK(cudaGetLastError());
  return MGB_SUCCESS;
}

// single_dimension::recompose (SingleDimension/DataRefactoring.hpp:110-194)
template <typename T> int recompose_single_t(mgb_plan *p, const T *d_in, T *d_out, cudaStream_t st) {
  int full[3], pad;
  i64 sfull[3];
  int rc = sd_prepare<T>(p, full, sfull, pad);
  if (rc)
    return rc;
  T *V = (T *)p->d_sd, *corr = (T *)p->d_wB;
  if (p->L == 0) {
    MGB_CUDA_CHECK(cudaMemcpyAsync(d_out, d_in, p->N * sizeof(T), cudaMemcpyDeviceToDevice, st));
    return MGB_SUCCESS;
  }
  MGB_CUDA_CHECK(cudaMemcpyAsync(V, d_in, p->N * sizeof(T), cudaMemcpyDeviceToDevice, st));
  for (int l = 0; l < p->L; l++) {
    for (int ax = pad; ax < 3; ax++) {
      const int dim = ax - pad;
      int fine[3], coarse[3];
      for (int d = 0; d < 3; d++)
        fine[d] = d < pad ? 1 : (int)(d > ax ? p->lshape[l][d - pad] : p->lshape[l + 1][d - pad]);
      const int nc = (int)p->lshape[l][dim];
      for (int d = 0; d < 3; d++)
        coarse[d] = d == ax ? nc : fine[d];
      sd_correction<T>(p, V, sfull, fine, ax, pad, l + 1, corr, st);
      sd_accumulate<T>(sfull, coarse, corr, V, true, st);
      const bool last = l == p->L - 1 && ax == 2;
      SdGeom g;
      g.ax = ax;
      g.nfine = fine[ax];
      g.nc = nc;
      for (int d = 0; d < 3; d++) {
        g.n[d] = coarse[d];
        g.sa[d] = sfull[d];
      }
      if (last) {
        for (int d = 0; d < 3; d++)
          g.sb[d] = sfull[d];
      } else {
        sd_dense_strides(fine, g.sb);
      }
      const i64 total = (i64)g.n[0] * g.n[1] * g.n[2];
      const T *ratio = (const T *)p->dtab(p->tab[l + 1][dim].ratio);
      MGB_LAUNCH(MGB_K_RESTORE, st,
                 (sd_restore_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, V, d_out, ratio)));
      if (!last) { // the output buffer served as the dense temporary: back into V
        SdGeom c;
        c.ax = 0;
        c.nfine = c.nc = 0;
        for (int d = 0; d < 3; d++) {
          c.n[d] = fine[d];
          c.sb[d] = sfull[d];
        }
        sd_dense_strides(c.n, c.sa);
        const i64 tot = (i64)c.n[0] * c.n[1] * c.n[2];
        MGB_LAUNCH(MGB_K_BOXCOPY, st,
                   (sd_box_kernel<T, 0><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(c, d_out, V)));
      }
    }
  }
  MGB_CUDA_CHECK(cudaGetLastError());
  return MGB_SUCCESS;
}

} // namespace

int mgb_decompose_impl(mgb_plan *p, const void *d_in, void *d_out, cudaStream_t st) {
  if (p->cfg.decomposition == 1) {
    if (p->dtype == MGB_F32)
      return decompose_single_t<float>(p, (const float *)d_in, (float *)d_out, st);
    return decompose_single_t<double>(p, (const double *)d_in, (double *)d_out, st);
  }
  if (p->dtype == MGB_F32)
    return decompose_t<float>(p, (const float *)d_in, (float *)d_out, st);
  return decompose_t<double>(p, (const double *)d_in, (double *)d_out, st);
}
int mgb_recompose_impl(mgb_plan *p, const void *d_in, void *d_out, cudaStream_t st) {
  if (p->cfg.decomposition == 1) {
    if (p->dtype == MGB_F32)
      return recompose_single_t<float>(p, (const float *)d_in, (float *)d_out, st);
    return recompose_single_t<double>(p, (const double *)d_in, (double *)d_out, st);
  }
  if (p->dtype == MGB_F32)
    return recompose_t<float>(p, (const float *)d_in, (float *)d_out, st);
  return recompose_t<double>(p, (const double *)d_in, (double *)d_out, st);
}

extern "C" int mgb_decompose(mgb_plan *plan, const void *d_in, void *d_out, void *stream) {
  if (!plan || !d_in || !d_out)
    return MGB_BAD_ARGUMENT;
  return mgb_decompose_impl(plan, d_in, d_out, (cudaStream_t)stream);
}
extern "C" int mgb_recompose(mgb_plan *plan, const void *d_in, void *d_out, void *stream) {
  if (!plan || !d_in || !d_out)
    return MGB_BAD_ARGUMENT;
  return mgb_recompose_impl(plan, d_in, d_out, (cudaStream_t)stream);
}
