// MGARD-CPU convention on the GPU: mgard::compress / mgard::decompress
// (reference include/compress.tpp:35-83) -- TensorMeshHierarchy
// (include/TensorMeshHierarchy.tpp:40-139), shuffle (include/shuffle.tpp:8-37),
// decompose / recompose (include/decompose.tpp:129-219) with their line operators
// (TensorProlongation.tpp:22-69, TensorMassMatrix.tpp:15-90,178-290,
// TensorRestriction.tpp:24-71), the multilevel coefficient quantizer
// (TensorMultilevelCoefficientQuantizer.tpp:13-77, LinearQuantizer.tpp:8-53), the
// zlib payload (src/compressors.cpp:552-629) and the header (src/format.cpp:102-140,
// 219-233).
//
// Bit-exact by construction: every value is produced by the reference's expression
// in the reference's order (no FMA contraction, IEEE division and square root).
// The reference addresses nodes through the hierarchy, so values do not depend on
// the memory layout: here the array stays NODAL (row-major) during the level
// recursion -- each level is a strided box described by per-dimension index lists,
// which also covers the non-dyadic top level and non-uniform coordinates -- and the
// level ("shuffled") order is produced only once, fused with the quantizer.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <vector>

#include <zlib.h>

#include "format.h"
#include "plan.h"

namespace {

constexpr int CD = MGB_MAX_DIMS; // dimensions are left-padded with size-1 ("flat") ones

template <typename T> struct DimTables { // one (level, dimension)
  uint32_t n = 1, nnew = 0, nold = 1;
  // offsets into the device pools
  uint64_t pos = 0, info = 0, newl = 0, oldl = 0; // uint32 pool
  uint64_t x = 0, w = 0, dv = 0, cc = 0, vw = 0;  // T pool
};

enum Op {
  OP_COPY_OLD_ZERO_NEW = 0, // decompose.tpp:78-91
  OP_PROLONG,               // TensorProlongation.tpp:22-69
  OP_SUB_NEW,               // decompose.tpp:110-126
  OP_MASS,                  // TensorMassMatrix.tpp:15-90
  OP_RESTRICT,              // TensorRestriction.tpp:24-71
  OP_ADD_OLD,               // decompose.tpp:33-41
  OP_ZERO_OLD_COPY_NEW,     // decompose.tpp:93-107
  OP_SUB_OLD_ZERO_NEW,      // decompose.tpp:43-57
  OP_NEG_OLD_SUB_NEW,       // decompose.tpp:59-76
  OP_SHUFFLE,               // shuffle.tpp:8-21
  OP_UNSHUFFLE,             // shuffle.tpp:23-37
  OP_QUANT_NODAL,           // nodal coefficients -> shuffled int64
  OP_DEQUANT_NODAL,         // shuffled int64 -> nodal coefficients
  OP_QUANT_SHUFFLED,        // shuffled coefficients -> shuffled int64
  OP_DEQUANT_SHUFFLED
};

template <typename T> struct LevelArgs {
  uint32_t cnt[CD];         // iteration extent per dimension
  const uint32_t *sel[CD];  // optional sub-list of level-l positions (null: all)
  const uint32_t *pos[CD];  // level-l position -> index in the finest grid
  const uint32_t *info[CD]; // (#level-(l-1) nodes before this one) << 1 | is_new
  const T *x[CD];           // coordinates of the level-l nodes
  const T *vw[CD];          // (x_succ - x_pred) / 2 in the level-l mesh
  uint64_t stride[CD];      // nodal strides in elements
  uint32_t n[CD];           // level-l sizes
  uint64_t csuffix[CD + 1]; // products of the level-(l-1) sizes of dims >= d
  uint32_t flat;            // bit d: dimension d has size 1
  int d;                    // dimension the operator acts along
  int level0;               // level 0 introduces all of its nodes
  uint64_t total;           // product of cnt
  uint64_t base;            // ndof(l - 1): first shuffled slot of the level
  T *v;
  T *buf;
  const T *src;
  T *dst;
  const T *sin; // shuffled input
  T *sout;      // shuffled output
  long long *q;
  const long long *qin;
  // quantizer
  int s_inf;
  T quantum; // s = inf
  T two_tol, exp2sl, ndof;
  int *flag;
};

template <typename T> __device__ __forceinline__ T quantum_of(const LevelArgs<T> &a, const uint32_t *j) {
  if (a.s_inf)
    return a.quantum;
  // s_quantum (TensorMultilevelCoefficientQuantizer.tpp:38-58)
  T vf = 1;
#pragma unroll
  for (int d = 0; d < CD; d++)
    if (!((a.flat >> d) & 1))
      vf *= a.vw[d][j[d]];
  return a.two_tol / (a.exp2sl * sqrt(a.ndof * vf));
}

template <typename T> __device__ __forceinline__ long long quantize_one(const LevelArgs<T> &a, T x, T quantum) {
  // LinearQuantizer (LinearQuantizer.tpp:8-26)
  const T minimum = (T)((double)quantum * ((double)std::numeric_limits<long long>::min() - 0.5));
  const T maximum = (T)((double)quantum * ((double)std::numeric_limits<long long>::max() + 0.5));
  if (x <= minimum || x >= maximum) {
    *a.flag = 1; // the reference throws std::domain_error
    return 0;
  }
  const double mag = 0.5 + (double)fabs(x / quantum);
  return (long long)copysign(mag, (double)x);
}

template <typename T, int OP> __global__ void __launch_bounds__(256) cpu_level_kernel(const LevelArgs<T> a) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.total)
    return;
  uint32_t j[CD];
  uint64_t rem = t, off = 0;
  bool allold = true;
#pragma unroll
  for (int d = CD - 1; d >= 0; d--) {
    const uint32_t c = a.cnt[d];
    uint32_t i = 0;
    if (c > 1) {
      i = (uint32_t)(rem % c);
      rem /= c;
    }
    j[d] = a.sel[d] ? a.sel[d][i] : i;
    off += (uint64_t)a.pos[d][j[d]] * a.stride[d];
    allold = allold && !(a.info[d][j[d]] & 1u);
  }
  if (a.level0)
    allold = false;
  const int D = a.d;

  if (OP == OP_COPY_OLD_ZERO_NEW) {
    a.buf[off] = allold ? a.v[off] : (T)0;
  } else if (OP == OP_ZERO_OLD_COPY_NEW) {
    a.buf[off] = allold ? (T)0 : a.v[off];
  } else if (OP == OP_SUB_NEW) {
    if (allold) {
      a.buf[off] = 0;
    } else {
      const T r = a.v[off] - a.buf[off];
      a.v[off] = r;
      a.buf[off] = r;
    }
  } else if (OP == OP_SUB_OLD_ZERO_NEW) {
    a.buf[off] = allold ? a.buf[off] + (T)(-1) * a.v[off] : (T)0;
  } else if (OP == OP_NEG_OLD_SUB_NEW) {
    a.v[off] = allold ? -a.buf[off] : a.v[off] + (T)(-1) * a.buf[off];
  } else if (OP == OP_ADD_OLD) {
    a.v[off] = a.v[off] + (T)1 * a.buf[off];
  } else if (OP == OP_PROLONG) {
    // j[D] is a new node; its level-l neighbours are the enclosing old nodes
    const uint32_t jm = j[D];
    const uint64_t sd = a.stride[D];
    const uint64_t base = off - (uint64_t)a.pos[D][jm] * sd;
    const T xl = a.x[D][jm - 1], xm = a.x[D][jm], xr = a.x[D][jm + 1];
    const T vl = a.buf[base + (uint64_t)a.pos[D][jm - 1] * sd];
    const T vr = a.buf[base + (uint64_t)a.pos[D][jm + 1] * sd];
    const T wr = (T)1 / (xr - xl);
    a.buf[off] = a.buf[off] + (vl * (xr - xm) + vr * (xm - xl)) * wr;
  } else if (OP == OP_MASS) {
    const uint32_t jm = j[D], n = a.n[D];
    const uint64_t sd = a.stride[D];
    const uint64_t base = off - (uint64_t)a.pos[D][jm] * sd;
    const T vm = a.src[off];
    T r;
    if (jm == 0) {
      const T hr = a.x[D][1] - a.x[D][0];
      const T vr = a.src[base + (uint64_t)a.pos[D][1] * sd];
      r = hr / 3 * vm + hr / 6 * vr;
    } else if (jm == n - 1) {
      const T hl = a.x[D][jm] - a.x[D][jm - 1];
      const T vl = a.src[base + (uint64_t)a.pos[D][jm - 1] * sd];
      r = hl / 6 * vl + hl / 3 * vm;
    } else {
      const T hl = a.x[D][jm] - a.x[D][jm - 1];
      const T hr = a.x[D][jm + 1] - a.x[D][jm];
      const T vl = a.src[base + (uint64_t)a.pos[D][jm - 1] * sd];
      const T vr = a.src[base + (uint64_t)a.pos[D][jm + 1] * sd];
      r = hl / 6 * vl + (hl + hr) / 3 * vm + hr / 6 * vr;
    }
    a.dst[off] = r;
  } else if (OP == OP_RESTRICT) {
    // j[D] is an old node: first the interval on its left, then the one on its right
    const uint32_t jc = j[D], n = a.n[D];
    const uint64_t sd = a.stride[D];
    const uint64_t base = off - (uint64_t)a.pos[D][jc] * sd;
    T c = a.buf[off];
    if (jc >= 2 && (a.info[D][jc - 1] & 1u)) {
      const T xl = a.x[D][jc - 2], xm = a.x[D][jc - 1], xr = a.x[D][jc];
      const T vm = a.buf[base + (uint64_t)a.pos[D][jc - 1] * sd];
      const T wr = (T)1 / (xr - xl);
      c = c + vm * (xm - xl) * wr;
    }
    if (jc + 2 < n && (a.info[D][jc + 1] & 1u)) {
      const T xl = a.x[D][jc], xm = a.x[D][jc + 1], xr = a.x[D][jc + 2];
      const T vm = a.buf[base + (uint64_t)a.pos[D][jc + 1] * sd];
      const T wr = (T)1 / (xr - xl);
      c = c + vm * (xr - xm) * wr;
    }
    a.buf[off] = c;
  } else {
    // level-order ("shuffled") slot of a node introduced by this level: its
    // row-major rank in the level-l mesh minus the level-(l-1) nodes before it
    if (allold)
      return;
    uint64_t lin = 0, before = 0;
    bool tight = true;
#pragma unroll
    for (int d = 0; d < CD; d++) {
      lin = lin * a.n[d] + j[d];
      if (tight) {
        const uint32_t inf = a.info[d][j[d]];
        before += (uint64_t)(inf >> 1) * a.csuffix[d + 1];
        tight = !(inf & 1u);
      }
    }
    const uint64_t sp = a.level0 ? lin : a.base + lin - before;
    if (OP == OP_SHUFFLE) {
      a.sout[sp] = a.v[off];
    } else if (OP == OP_UNSHUFFLE) {
      a.v[off] = a.sin[sp];
    } else if (OP == OP_QUANT_NODAL) {
      a.q[sp] = quantize_one(a, a.v[off], quantum_of(a, j));
    } else if (OP == OP_QUANT_SHUFFLED) {
      a.q[sp] = quantize_one(a, a.sin[sp], quantum_of(a, j));
    } else if (OP == OP_DEQUANT_NODAL) {
      a.v[off] = quantum_of(a, j) * (T)a.qin[sp]; // LinearDequantizer (LinearQuantizer.tpp:41-53)
    } else if (OP == OP_DEQUANT_SHUFFLED) {
      a.sout[sp] = quantum_of(a, j) * (T)a.qin[sp];
    }
  }
}

// ConstituentMassMatrixInverse (TensorMassMatrix.tpp:178-290): one thread per line
// of the level box, w[j] = (h_{j-1}/6) / divisors[j-1], cc[j] = h_j / 6.
template <typename T>
__global__ void __launch_bounds__(128) cpu_thomas_kernel(const LevelArgs<T> a, const T *__restrict__ w,
                                                         const T *__restrict__ dv, const T *__restrict__ cc) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.total)
    return;
  uint64_t rem = t, off = 0;
  const int D = a.d;
#pragma unroll
  for (int d = CD - 1; d >= 0; d--) {
    const uint32_t c = a.cnt[d];
    uint32_t i = 0;
    if (c > 1) {
      i = (uint32_t)(rem % c);
      rem /= c;
    }
    if (d != D)
      off += (uint64_t)a.pos[d][i] * a.stride[d];
  }
  T *p = a.buf + off;
  const uint32_t *pos = a.pos[D];
  const uint64_t sd = a.stride[D];
  const uint32_t n = a.n[D];
  T prev = p[(uint64_t)pos[0] * sd];
  for (uint32_t j = 1; j + 1 < n; j++) {
    T *e = p + (uint64_t)pos[j] * sd;
    prev = *e - w[j] * prev;
    *e = prev;
  }
  T *last = p + (uint64_t)pos[n - 1] * sd;
  T nxt = (*last - w[n - 1] * prev) / dv[n - 1];
  *last = nxt;
  for (uint32_t j = n - 1; j-- > 0;) {
    T *e = p + (uint64_t)pos[j] * sd;
    nxt = (*e - cc[j] * nxt) / dv[j];
    *e = nxt;
  }
}

} // namespace

// mgard::TensorMeshHierarchy<N, Real> + device tables + workspaces
struct mgb_cpu_plan {
  int ndim = 0, dtype = MGB_F64;
  size_t tsize = 8;
  uint64_t shape[CD] = {1, 1, 1, 1, 1};  // left-padded
  uint64_t user_shape[CD] = {1, 1, 1, 1, 1};
  uint64_t stride[CD];
  uint32_t flat = 0;
  int L = 0;
  bool uniform = true;
  uint64_t N = 0;
  std::vector<std::vector<uint64_t>> shapes; // [l][d]
  std::vector<uint64_t> ndof;                // [l]
  std::vector<std::vector<double>> coords;   // user dims, as doubles (header)
  std::vector<std::vector<DimTables<double>>> tab; // [l][d] (offsets are type-agnostic)
  uint32_t *d_u32 = nullptr;
  unsigned char *d_real = nullptr;
  std::vector<unsigned char> h_real; // host copy of the T pool (coordinates)
  unsigned char *d_v = nullptr, *d_b0 = nullptr, *d_b1 = nullptr; // N * T each
  long long *d_q = nullptr;
  int *d_flag = nullptr;
  ~mgb_cpu_plan() {
    cudaFree(d_u32);
    cudaFree(d_real);
    cudaFree(d_v);
    cudaFree(d_b0);
    cudaFree(d_b1);
    cudaFree(d_q);
    cudaFree(d_flag);
  }
};

namespace {

inline int floor_log2(uint64_t n) { // log2 of TensorMeshHierarchy.tpp:11-20
  int e = -1;
  for (; n; ++e, n >>= 1)
    ;
  return e;
}

template <typename T>
int build_plan(mgb_cpu_plan *p, const void *const *coords_in) {
  const int pad = CD - p->ndim;
  // coordinates in T (TensorMeshHierarchy.tpp:145-157 for the uniform constructor)
  std::vector<std::vector<T>> xs(CD);
  for (int d = 0; d < CD; d++) {
    const uint64_t n = p->shape[d];
    xs[d].resize(n);
    if (d >= pad && coords_in) {
      memcpy(xs[d].data(), coords_in[d - pad], n * sizeof(T));
    } else {
      const T h = n > 1 ? static_cast<T>(1) / (n - 1) : 0;
      for (uint64_t j = 0; j < n; j++)
        xs[d][j] = j * h;
    }
    if (d >= pad) {
      p->coords[d - pad].assign(xs[d].begin(), xs[d].end());
      for (uint64_t j = 1; j < n; j++)
        if (!(xs[d][j] > xs[d][j - 1]))
          return MGB_BAD_ARGUMENT;
    }
  }
  // levels (TensorMeshHierarchy.tpp:52-97)
  bool any_nonflat = false, any_nondyadic = false;
  uint64_t rounded[CD];
  int L_dyadic = std::numeric_limits<int>::max();
  for (int d = 0; d < CD; d++) {
    const uint64_t size = p->shape[d];
    if (size == 0)
      return MGB_BAD_ARGUMENT;
    if (size == 1) {
      rounded[d] = 1;
      continue;
    }
    any_nonflat = true;
    const int l = floor_log2(size - 1);
    L_dyadic = std::min(L_dyadic, l);
    rounded[d] = (1ull << l) + 1;
    any_nondyadic = any_nondyadic || rounded[d] != size;
  }
  if (!any_nonflat)
    return MGB_BAD_ARGUMENT;
  p->L = any_nondyadic ? L_dyadic + 1 : L_dyadic;
  const int L = p->L;
  p->shapes.assign(L + 1, std::vector<uint64_t>(CD, 1));
  {
    uint64_t cur[CD];
    for (int d = 0; d < CD; d++)
      cur[d] = ((rounded[d] - 1) >> L_dyadic) + 1;
    for (int l = 0; l < L; l++)
      for (int d = 0; d < CD; d++) {
        p->shapes[l][d] = cur[d];
        cur[d] = ((cur[d] - 1) << 1) + 1;
      }
    for (int d = 0; d < CD; d++)
      p->shapes[L][d] = p->shape[d];
  }
  p->ndof.resize(L + 1);
  for (int l = 0; l <= L; l++) {
    uint64_t m = 1;
    for (int d = 0; d < CD; d++)
      m *= p->shapes[l][d];
    p->ndof[l] = m;
  }
  // per (level, dimension) tables
  std::vector<uint32_t> u32;
  std::vector<T> real;
  p->tab.assign(L + 1, std::vector<DimTables<double>>(CD));
  for (int d = 0; d < CD; d++) {
    const uint64_t ntop = p->shape[d];
    std::vector<int> dob(ntop, 0);
    std::vector<std::vector<uint32_t>> idx(L + 1);
    for (int l = 0; l <= L; l++) {
      const uint64_t n = p->shapes[l][d];
      idx[l].resize(n);
      for (uint64_t j = 0; j < n; j++) // TensorMeshHierarchy.tpp:103-113
        idx[l][j] = ntop == 1 ? 0 : (uint32_t)((j * (ntop - 1)) / (n - 1));
    }
    for (int l = L; l >= 0; l--)
      for (uint32_t i : idx[l])
        dob[i] = l;
    for (int l = 0; l <= L; l++) {
      DimTables<double> &t = p->tab[l][d];
      const uint32_t n = (uint32_t)idx[l].size();
      t.n = n;
      t.pos = u32.size();
      u32.insert(u32.end(), idx[l].begin(), idx[l].end());
      std::vector<uint32_t> info(n), newl, oldl;
      uint32_t before = 0;
      for (uint32_t j = 0; j < n; j++) {
        const bool is_new = l > 0 && dob[idx[l][j]] == l;
        info[j] = (before << 1) | (is_new ? 1u : 0u);
        if (is_new) {
          newl.push_back(j);
        } else {
          oldl.push_back(j);
          before++;
        }
      }
      t.nnew = (uint32_t)newl.size();
      t.nold = (uint32_t)oldl.size();
      t.info = u32.size();
      u32.insert(u32.end(), info.begin(), info.end());
      t.newl = u32.size();
      u32.insert(u32.end(), newl.begin(), newl.end());
      t.oldl = u32.size();
      u32.insert(u32.end(), oldl.begin(), oldl.end());
      // a new node must sit strictly between two old ones that are adjacent in the
      // level-l list (guaranteed by n_{l-1} - 1 >= (n_l - 1) / 2)
      for (uint32_t j : newl)
        if (j == 0 || j + 1 >= n || (info[j - 1] & 1u) || (info[j + 1] & 1u))
          return MGB_FAILURE;
      std::vector<T> x(n), w(n, 0), dv(n, 1), cc(n, 0), vw(n, 0);
      for (uint32_t j = 0; j < n; j++)
        x[j] = xs[d][idx[l][j]];
      if (n >= 2) {
        // divisors (TensorMassMatrix.tpp:123-176), evaluated in T
        T h_right = x[1] - x[0], h_left = 0;
        dv[0] = 2 * h_right / 6;
        for (uint32_t j = 1; j + 1 < n; j++) {
          h_left = h_right;
          h_right = x[j + 1] - x[j];
          const T a_j = h_left / 6;
          const T wj = a_j / dv[j - 1];
          w[j] = wj;
          dv[j] = 2 * (h_left + h_right) / 6 - wj * a_j;
        }
        {
          h_left = h_right;
          const T a_j = h_left / 6;
          const T wj = a_j / dv[n - 2];
          w[n - 1] = wj;
          dv[n - 1] = 2 * h_left / 6 - wj * a_j;
        }
        for (uint32_t j = 0; j + 1 < n; j++) {
          const T h = x[j + 1] - x[j];
          cc[j] = h / 6;
        }
        // quantizer volume factor per dimension (TensorMultilevelCoefficientQuantizer.tpp:44-52;
        // predecessor / successor saturate at the ends, utilities.tpp:297-317)
        for (uint32_t j = 0; j < n; j++)
          vw[j] = (x[j + 1 < n ? j + 1 : j] - x[j ? j - 1 : 0]) / 2;
      }
      t.x = real.size();
      real.insert(real.end(), x.begin(), x.end());
      t.w = real.size();
      real.insert(real.end(), w.begin(), w.end());
      t.dv = real.size();
      real.insert(real.end(), dv.begin(), dv.end());
      t.cc = real.size();
      real.insert(real.end(), cc.begin(), cc.end());
      t.vw = real.size();
      real.insert(real.end(), vw.begin(), vw.end());
    }
  }
  if (cudaMalloc(&p->d_u32, u32.size() * 4 + 16) != cudaSuccess ||
      cudaMalloc(&p->d_real, real.size() * sizeof(T) + 16) != cudaSuccess ||
      cudaMalloc(&p->d_flag, sizeof(int)) != cudaSuccess)
    return MGB_CUDA_ERROR;
  MGB_CUDA_CHECK(cudaMemcpy(p->d_u32, u32.data(), u32.size() * 4, cudaMemcpyHostToDevice));
  MGB_CUDA_CHECK(cudaMemcpy(p->d_real, real.data(), real.size() * sizeof(T), cudaMemcpyHostToDevice));
  MGB_CUDA_CHECK(cudaMemset(p->d_flag, 0, sizeof(int)));
  return MGB_SUCCESS;
}

int ensure_workspace(mgb_cpu_plan *p, bool need_q) {
  const size_t bytes = p->N * p->tsize;
  if (!p->d_v) {
    if (cudaMalloc(&p->d_v, bytes) != cudaSuccess || cudaMalloc(&p->d_b0, bytes) != cudaSuccess ||
        cudaMalloc(&p->d_b1, bytes) != cudaSuccess) {
      cudaGetLastError();
      return MGB_CUDA_ERROR;
    }
  }
  if (need_q && !p->d_q) {
    if (cudaMalloc(&p->d_q, p->N * sizeof(long long)) != cudaSuccess) {
      cudaGetLastError();
      return MGB_CUDA_ERROR;
    }
  }
  return MGB_SUCCESS;
}

enum Select { SEL_ALL = 0, SEL_NEW, SEL_OLD, SEL_ONE };

// Level-l box with one dimension optionally restricted to its new / old nodes.
template <typename T> LevelArgs<T> level_args(const mgb_cpu_plan *p, int l, int d_act, Select sel_act) {
  LevelArgs<T> a;
  memset(&a, 0, sizeof(a));
  const T *real = reinterpret_cast<const T *>(p->d_real);
  a.total = 1;
  a.csuffix[CD] = 1;
  for (int d = CD - 1; d >= 0; d--) {
    const DimTables<double> &t = p->tab[l][d];
    a.n[d] = t.n;
    a.cnt[d] = t.n;
    a.sel[d] = nullptr;
    if (d == d_act) {
      if (sel_act == SEL_NEW) {
        a.cnt[d] = t.nnew;
        a.sel[d] = p->d_u32 + t.newl;
      } else if (sel_act == SEL_OLD) {
        a.cnt[d] = t.nold;
        a.sel[d] = p->d_u32 + t.oldl;
      } else if (sel_act == SEL_ONE) {
        a.cnt[d] = 1;
      }
    }
    a.pos[d] = p->d_u32 + t.pos;
    a.info[d] = p->d_u32 + t.info;
    a.x[d] = real + t.x;
    a.vw[d] = real + t.vw;
    a.stride[d] = p->stride[d];
    a.total *= a.cnt[d];
    a.csuffix[d] = a.csuffix[d + 1] * (l > 0 ? p->shapes[l - 1][d] : 1);
  }
  a.flat = p->flat;
  a.d = d_act < 0 ? 0 : d_act;
  a.level0 = l == 0;
  a.base = l > 0 ? p->ndof[l - 1] : 0;
  a.flag = p->d_flag;
  return a;
}

template <typename T, int OP> void launch_level(const LevelArgs<T> &a, cudaStream_t st) {
  if (a.total == 0)
    return;
  const uint64_t blocks = (a.total + 255) / 256;
  MGB_LAUNCH(MGB_K_AXPY, st, (cpu_level_kernel<T, OP><<<(unsigned)blocks, 256, 0, st>>>(a)));
}

// M, R on level l, M^-1 on level l - 1 (decompose.tpp:156-163): the projection of
// the level-l coefficient function in `b0`; returns the buffer holding the result
template <typename T> T *project(mgb_cpu_plan *p, int l, T *b0, T *b1, cudaStream_t st) {
  T *cur = b0, *other = b1;
  for (int d = 0; d < CD; d++) {
    if ((p->flat >> d) & 1)
      continue;
    LevelArgs<T> a = level_args<T>(p, l, d, SEL_ALL);
    a.src = cur;
    a.dst = other;
    launch_level<T, OP_MASS>(a, st);
    std::swap(cur, other);
  }
  for (int d = 0; d < CD; d++) {
    if ((p->flat >> d) & 1)
      continue;
    LevelArgs<T> a = level_args<T>(p, l, d, SEL_OLD);
    a.buf = cur;
    launch_level<T, OP_RESTRICT>(a, st);
  }
  const T *real = reinterpret_cast<const T *>(p->d_real);
  for (int d = 0; d < CD; d++) {
    if ((p->flat >> d) & 1)
      continue;
    LevelArgs<T> a = level_args<T>(p, l - 1, d, SEL_ONE);
    a.buf = cur;
    const DimTables<double> &t = p->tab[l - 1][d];
    const uint64_t blocks = (a.total + 127) / 128;
    MGB_LAUNCH(MGB_K_THOMAS_STRIDED, st,
               (cpu_thomas_kernel<T><<<(unsigned)blocks, 128, 0, st>>>(a, real + t.w, real + t.dv, real + t.cc)));
  }
  return cur;
}

// mgard::decompose on the nodal array `v` (decompose.tpp:129-174)
template <typename T> void decompose_nodal(mgb_cpu_plan *p, T *v, cudaStream_t st) {
  T *b0 = reinterpret_cast<T *>(p->d_b0), *b1 = reinterpret_cast<T *>(p->d_b1);
  for (int l = p->L; l > 0; l--) {
    LevelArgs<T> a = level_args<T>(p, l, -1, SEL_ALL);
    a.v = v;
    a.buf = b0;
    launch_level<T, OP_COPY_OLD_ZERO_NEW>(a, st);
    for (int d = 0; d < CD; d++) {
      if ((p->flat >> d) & 1)
        continue;
      LevelArgs<T> pa = level_args<T>(p, l, d, SEL_NEW);
      pa.buf = b0;
      launch_level<T, OP_PROLONG>(pa, st);
    }
    launch_level<T, OP_SUB_NEW>(a, st);
    T *corr = project<T>(p, l, b0, b1, st);
    LevelArgs<T> c = level_args<T>(p, l - 1, -1, SEL_ALL);
    c.v = v;
    c.buf = corr;
    // the all-old test of the level kernel is irrelevant to OP_ADD_OLD
    launch_level<T, OP_ADD_OLD>(c, st);
  }
}

// mgard::recompose on the nodal array `v` (decompose.tpp:177-219)
template <typename T> void recompose_nodal(mgb_cpu_plan *p, T *v, cudaStream_t st) {
  T *b0 = reinterpret_cast<T *>(p->d_b0), *b1 = reinterpret_cast<T *>(p->d_b1);
  for (int l = 1; l <= p->L; l++) {
    LevelArgs<T> a = level_args<T>(p, l, -1, SEL_ALL);
    a.v = v;
    a.buf = b0;
    launch_level<T, OP_ZERO_OLD_COPY_NEW>(a, st);
    T *corr = project<T>(p, l, b0, b1, st);
    a.buf = corr;
    launch_level<T, OP_SUB_OLD_ZERO_NEW>(a, st);
    for (int d = 0; d < CD; d++) {
      if ((p->flat >> d) & 1)
        continue;
      LevelArgs<T> pa = level_args<T>(p, l, d, SEL_NEW);
      pa.buf = corr;
      launch_level<T, OP_PROLONG>(pa, st);
    }
    launch_level<T, OP_NEG_OLD_SUB_NEW>(a, st);
  }
}

template <typename T> void set_quantizer(const mgb_cpu_plan *p, LevelArgs<T> &a, int l, double s_in, double tol_in) {
  const T s = (T)s_in, tol = (T)tol_in;
  a.s_inf = std::isinf(s) && s > 0;
  if (a.s_inf) {
    // supremum_quantum (TensorMultilevelCoefficientQuantizer.tpp:13-27)
    std::size_t dims = 0;
    for (int d = 0; d < CD; d++)
      if (p->shape[d] > 1)
        ++dims;
    a.quantum = (2 * tol) / ((static_cast<std::size_t>(p->L) + 1) * (1 + std::pow(3, dims)));
  } else {
    a.two_tol = 2 * tol;
    a.exp2sl = std::exp2(s * static_cast<std::size_t>(l));
    a.ndof = static_cast<T>(static_cast<std::size_t>(p->N)); // ndof * volume_factor is evaluated in T
  }
}

// one pass per level over the nodes that level introduces
template <typename T, int OP>
void level_map(mgb_cpu_plan *p, T *v, const T *sin, T *sout, long long *q, const long long *qin, double s,
               double tol, cudaStream_t st) {
  for (int l = 0; l <= p->L; l++) {
    LevelArgs<T> a = level_args<T>(p, l, -1, SEL_ALL);
    a.v = v;
    a.sin = sin;
    a.sout = sout;
    a.q = q;
    a.qin = qin;
    if (OP >= OP_QUANT_NODAL)
      set_quantizer<T>(p, a, l, s, tol);
    launch_level<T, OP>(a, st);
  }
}

int check_flag(mgb_cpu_plan *p, cudaStream_t st) {
  int flag = 0;
  MGB_CUDA_CHECK(cudaMemcpyAsync(&flag, p->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  MGB_CUDA_CHECK(cudaStreamSynchronize(st));
  if (flag) {
    cudaMemsetAsync(p->d_flag, 0, sizeof(int), st);
    return MGB_FAILURE; // "number too large to be quantized" (LinearQuantizer.tpp:21-23)
  }
  return MGB_SUCCESS;
}

bool device_pointer(const void *ptr) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

bool have_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return false;
  }
  return true;
}

} // namespace

extern "C" {

int mgb_cpu_plan_create(int ndim, const uint64_t *shape, int dtype, const void *const *coords,
                        mgb_cpu_plan **plan) {
  if (!plan || !shape)
    return MGB_BAD_ARGUMENT;
  *plan = nullptr;
  if (ndim < 1 || ndim > CD)
    return MGB_TOO_MANY_DIMS;
  if (dtype != MGB_F32 && dtype != MGB_F64)
    return MGB_BAD_DTYPE;
  if (!have_device())
    return MGB_BACKEND_NOT_AVAILABLE; // no CPU fallback
  mgb_cpu_plan *p = new mgb_cpu_plan();
  p->ndim = ndim;
  p->dtype = dtype;
  p->tsize = dtype == MGB_F32 ? 4 : 8;
  p->uniform = coords == nullptr;
  p->coords.resize(ndim);
  const int pad = CD - ndim;
  p->N = 1;
  for (int d = 0; d < ndim; d++) {
    p->shape[pad + d] = shape[d];
    p->user_shape[d] = shape[d];
    if (shape[d] == 0 || shape[d] >= (1ull << 32)) {
      delete p;
      return MGB_BAD_ARGUMENT;
    }
    p->N *= shape[d];
  }
  uint64_t s = 1;
  for (int d = CD - 1; d >= 0; d--) {
    p->stride[d] = s;
    s *= p->shape[d];
    if (p->shape[d] == 1)
      p->flat |= 1u << d;
  }
  const int rc = dtype == MGB_F32 ? build_plan<float>(p, coords) : build_plan<double>(p, coords);
  if (rc) {
    delete p;
    return rc;
  }
  *plan = p;
  return MGB_SUCCESS;
}

void mgb_cpu_plan_destroy(mgb_cpu_plan *plan) { delete plan; }

int mgb_cpu_plan_levels(const mgb_cpu_plan *plan) { return plan ? plan->L : -1; }

uint64_t mgb_cpu_plan_ndof(const mgb_cpu_plan *plan, int level) {
  if (!plan || level < 0 || level > plan->L)
    return 0;
  return plan->ndof[level];
}

uint64_t mgb_cpu_plan_level_shape(const mgb_cpu_plan *plan, int level, int dim) {
  if (!plan || level < 0 || level > plan->L || dim < 0 || dim >= plan->ndim)
    return 0;
  return plan->shapes[level][CD - plan->ndim + dim];
}

#define CPU_DISPATCH(p, expr_f, expr_d)                                                                          \
  do {                                                                                                           \
    if ((p)->dtype == MGB_F32) {                                                                                 \
      typedef float T;                                                                                           \
      expr_f;                                                                                                    \
    } else {                                                                                                     \
      typedef double T;                                                                                          \
      expr_d;                                                                                                    \
    }                                                                                                            \
  } while (0)

int mgb_cpu_shuffle(mgb_cpu_plan *p, const void *d_in, void *d_out, void *stream) {
  if (!p || !d_in || !d_out)
    return MGB_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  CPU_DISPATCH(p, (level_map<T, OP_SHUFFLE>(p, (T *)d_in, nullptr, (T *)d_out, nullptr, nullptr, 0, 0, st)),
               (level_map<T, OP_SHUFFLE>(p, (T *)d_in, nullptr, (T *)d_out, nullptr, nullptr, 0, 0, st)));
  MGB_CUDA_CHECK(cudaGetLastError());
  return MGB_SUCCESS;
}

int mgb_cpu_unshuffle(mgb_cpu_plan *p, const void *d_in, void *d_out, void *stream) {
  if (!p || !d_in || !d_out)
    return MGB_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  CPU_DISPATCH(p, (level_map<T, OP_UNSHUFFLE>(p, (T *)d_out, (const T *)d_in, nullptr, nullptr, nullptr, 0, 0, st)),
               (level_map<T, OP_UNSHUFFLE>(p, (T *)d_out, (const T *)d_in, nullptr, nullptr, nullptr, 0, 0, st)));
  MGB_CUDA_CHECK(cudaGetLastError());
  return MGB_SUCCESS;
}

int mgb_cpu_decompose(mgb_cpu_plan *p, const void *d_in, void *d_out, void *stream) {
  if (!p || !d_in || !d_out)
    return MGB_BAD_ARGUMENT;
  int rc = ensure_workspace(p, false);
  if (rc)
    return rc;
  cudaStream_t st = (cudaStream_t)stream;
  MGB_CUDA_CHECK(cudaMemcpyAsync(p->d_v, d_in, p->N * p->tsize, cudaMemcpyDeviceToDevice, st));
  CPU_DISPATCH(p, (decompose_nodal<T>(p, (T *)p->d_v, st), level_map<T, OP_SHUFFLE>(p, (T *)p->d_v, nullptr, (T *)d_out, nullptr, nullptr, 0, 0, st)),
               (decompose_nodal<T>(p, (T *)p->d_v, st), level_map<T, OP_SHUFFLE>(p, (T *)p->d_v, nullptr, (T *)d_out, nullptr, nullptr, 0, 0, st)));
  MGB_CUDA_CHECK(cudaGetLastError());
  return MGB_SUCCESS;
}

int mgb_cpu_recompose(mgb_cpu_plan *p, const void *d_in, void *d_out, void *stream) {
  if (!p || !d_in || !d_out)
    return MGB_BAD_ARGUMENT;
  int rc = ensure_workspace(p, false);
  if (rc)
    return rc;
  cudaStream_t st = (cudaStream_t)stream;
  CPU_DISPATCH(p, (level_map<T, OP_UNSHUFFLE>(p, (T *)d_out, (const T *)d_in, nullptr, nullptr, nullptr, 0, 0, st), recompose_nodal<T>(p, (T *)d_out, st)),
               (level_map<T, OP_UNSHUFFLE>(p, (T *)d_out, (const T *)d_in, nullptr, nullptr, nullptr, 0, 0, st), recompose_nodal<T>(p, (T *)d_out, st)));
  MGB_CUDA_CHECK(cudaGetLastError());
  return MGB_SUCCESS;
}

int mgb_cpu_quantize(mgb_cpu_plan *p, const void *d_coef, double s, double tol, int64_t *d_q, void *stream) {
  if (!p || !d_coef || !d_q || !(tol > 0))
    return MGB_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  CPU_DISPATCH(p, (level_map<T, OP_QUANT_SHUFFLED>(p, nullptr, (const T *)d_coef, nullptr, (long long *)d_q, nullptr, s, tol, st)),
               (level_map<T, OP_QUANT_SHUFFLED>(p, nullptr, (const T *)d_coef, nullptr, (long long *)d_q, nullptr, s, tol, st)));
  MGB_CUDA_CHECK(cudaGetLastError());
  return check_flag(p, st);
}

int mgb_cpu_dequantize(mgb_cpu_plan *p, const int64_t *d_q, double s, double tol, void *d_coef, void *stream) {
  if (!p || !d_coef || !d_q || !(tol > 0))
    return MGB_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  CPU_DISPATCH(p, (level_map<T, OP_DEQUANT_SHUFFLED>(p, nullptr, nullptr, (T *)d_coef, nullptr, (const long long *)d_q, s, tol, st)),
               (level_map<T, OP_DEQUANT_SHUFFLED>(p, nullptr, nullptr, (T *)d_coef, nullptr, (const long long *)d_q, s, tol, st)));
  MGB_CUDA_CHECK(cudaGetLastError());
  return MGB_SUCCESS;
}

// mgard::compress + CompressedDataset::write (compress.tpp:35-67, CompressedDataset.tpp:26-29).
int mgb_cpu_compress(int ndim, int dtype, const uint64_t *shape, const void *const *coords, double s, double tol,
                     const void *in, void **out, size_t *out_size) {
  if (!in || !out || !out_size || !(tol > 0))
    return MGB_BAD_ARGUMENT;
  mgb_cpu_plan *p = nullptr;
  int rc = mgb_cpu_plan_create(ndim, shape, dtype, coords, &p);
  if (rc)
    return rc;
  struct Guard {
    mgb_cpu_plan *p;
    ~Guard() { delete p; }
  } guard{p};
  // compress_memory_z feeds the whole buffer through a 32-bit avail_in
  // (compressors.cpp:560): larger inputs are not representable in this format
  if (p->N * sizeof(long long) > 0xffffffffull)
    return MGB_OUTPUT_TOO_LARGE;
  rc = ensure_workspace(p, true);
  if (rc)
    return rc;
  cudaStream_t st = 0;
  const size_t bytes = p->N * p->tsize;
  MGB_CUDA_CHECK(cudaMemcpyAsync(p->d_v, in, bytes, device_pointer(in) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  CPU_DISPATCH(p, (decompose_nodal<T>(p, (T *)p->d_v, st), level_map<T, OP_QUANT_NODAL>(p, (T *)p->d_v, nullptr, nullptr, p->d_q, nullptr, s, tol, st)),
               (decompose_nodal<T>(p, (T *)p->d_v, st), level_map<T, OP_QUANT_NODAL>(p, (T *)p->d_v, nullptr, nullptr, p->d_q, nullptr, s, tol, st)));
  MGB_CUDA_CHECK(cudaGetLastError());
  rc = check_flag(p, st);
  if (rc)
    return rc;
  std::vector<long long> q(p->N);
  MGB_CUDA_CHECK(cudaMemcpy(q.data(), p->d_q, p->N * sizeof(long long), cudaMemcpyDeviceToHost));

  mgb_header h;
  h.convention = 1;
  h.ndim = ndim;
  h.dtype = dtype;
  for (int d = 0; d < ndim; d++)
    h.shape[d] = shape[d];
  h.ebtype = MGB_ABS;
  // the reference stores the Real-typed arguments widened to double
  h.s = dtype == MGB_F32 ? (double)(float)s : s;
  h.tol = dtype == MGB_F32 ? (double)(float)tol : tol;
  if (!p->uniform)
    h.coords = p->coords;
  const std::vector<uint8_t> head = mgb_encode_stream_header(h);

  // compress_memory_z (compressors.cpp:552-606): one deflate stream, level 9
  z_stream strm;
  memset(&strm, 0, sizeof(strm));
  if (deflateInit(&strm, Z_BEST_COMPRESSION) != Z_OK)
    return MGB_FAILURE;
  const size_t src_bytes = q.size() * sizeof(long long);
  const size_t bound = deflateBound(&strm, (uLong)src_bytes);
  uint8_t *buf = (uint8_t *)malloc(head.size() + bound);
  if (!buf) {
    deflateEnd(&strm);
    return MGB_FAILURE;
  }
  memcpy(buf, head.data(), head.size());
  strm.next_in = reinterpret_cast<Bytef *>(q.data());
  strm.avail_in = (uInt)src_bytes;
  strm.next_out = buf + head.size();
  strm.avail_out = (uInt)std::min<size_t>(bound, 0xffffffffu);
  const int zr = deflate(&strm, Z_FINISH);
  const size_t payload = bound - strm.avail_out;
  deflateEnd(&strm);
  if (zr != Z_STREAM_END) {
    free(buf);
    return MGB_FAILURE;
  }
  *out = buf;
  *out_size = head.size() + payload;
  return MGB_SUCCESS;
}

// mgard::decompress(void const *, size_t) (compress.tpp:69-83 behind the header
// dispatch of include/compress.hpp:62-72).  *out is malloc'ed (host).
int mgb_cpu_decompress(const void *in, size_t in_size, void **out, int *ndim, uint64_t *shape, int *dtype) {
  if (!in || !out)
    return MGB_BAD_ARGUMENT;
  if (!have_device())
    return MGB_BACKEND_NOT_AVAILABLE;
  mgb_header h;
  uint64_t hb = 0;
  int rc = mgb_parse_stream_header((const uint8_t *)in, in_size, h, hb);
  if (rc)
    return rc;
  if (h.convention != 1)
    return MGB_BAD_STREAM;
  std::vector<std::vector<float>> cf;
  std::vector<const void *> cptr;
  if (!h.coords.empty()) {
    for (int d = 0; d < h.ndim; d++) {
      if (h.dtype == MGB_F32) {
        cf.emplace_back(h.coords[d].begin(), h.coords[d].end());
        cptr.push_back(cf.back().data());
      } else {
        cptr.push_back(h.coords[d].data());
      }
    }
    if (h.dtype == MGB_F32) // emplace_back may have moved the vectors
      for (int d = 0; d < h.ndim; d++)
        cptr[d] = cf[d].data();
  }
  mgb_cpu_plan *p = nullptr;
  rc = mgb_cpu_plan_create(h.ndim, h.shape, h.dtype, cptr.empty() ? nullptr : cptr.data(), &p);
  if (rc)
    return rc;
  struct Guard {
    mgb_cpu_plan *p;
    ~Guard() { delete p; }
  } guard{p};
  rc = ensure_workspace(p, true);
  if (rc)
    return rc;
  // decompress_memory_z (compressors.cpp:608-629)
  std::vector<long long> q(p->N);
  {
    z_stream strm;
    memset(&strm, 0, sizeof(strm));
    strm.next_in = const_cast<Bytef *>((const Bytef *)in + hb);
    strm.avail_in = (uInt)(in_size - hb);
    strm.next_out = reinterpret_cast<Bytef *>(q.data());
    strm.avail_out = (uInt)(q.size() * sizeof(long long));
    if (inflateInit2(&strm, 15 + 32) != Z_OK)
      return MGB_BAD_STREAM;
    const int zr = inflate(&strm, Z_FINISH);
    const bool full = strm.avail_out == 0;
    inflateEnd(&strm);
    if (zr != Z_STREAM_END || !full)
      return MGB_BAD_STREAM;
  }
  cudaStream_t st = 0;
  MGB_CUDA_CHECK(cudaMemcpyAsync(p->d_q, q.data(), q.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
  CPU_DISPATCH(p, (level_map<T, OP_DEQUANT_NODAL>(p, (T *)p->d_v, nullptr, nullptr, nullptr, p->d_q, h.s, h.tol, st), recompose_nodal<T>(p, (T *)p->d_v, st)),
               (level_map<T, OP_DEQUANT_NODAL>(p, (T *)p->d_v, nullptr, nullptr, nullptr, p->d_q, h.s, h.tol, st), recompose_nodal<T>(p, (T *)p->d_v, st)));
  MGB_CUDA_CHECK(cudaGetLastError());
  const size_t bytes = p->N * p->tsize;
  void *host = malloc(bytes);
  if (!host)
    return MGB_FAILURE;
  if (cudaMemcpy(host, p->d_v, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) {
    free(host);
    return MGB_CUDA_ERROR;
  }
  *out = host;
  if (ndim)
    *ndim = h.ndim;
  if (shape)
    for (int d = 0; d < h.ndim; d++)
      shape[d] = h.shape[d];
  if (dtype)
    *dtype = h.dtype;
  return MGB_SUCCESS;
}

} // extern "C"
