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
#include <queue>
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
  OP_COEF_FUSED,            // copy_on_old_zero_on_new + prolongation + subtract in one pass
  OP_RECOMP_OLD,            // subtract_on_old + copy_negation_on_old (old nodes)
  OP_RECOMP_NEW,            // prolongation + subtract_on_new (new nodes)
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

// Value the dimension-by-dimension TensorProlongationAddition (TensorProlongation.tpp:22-69,
// TensorLinearOperator.tpp:71-109) leaves on a new node of the level-l mesh when the new
// nodes start from zero: nested 1-D interpolations from the surrounding old nodes, the
// lowest new dimension innermost (it is the first pass to run), each stage added to zero
// exactly as the passes do.  `src` holds the values on the old nodes.
template <typename T>
__device__ __forceinline__ T nested_interpolation(const LevelArgs<T> &a, const T *__restrict__ src,
                                                  const uint32_t *j, uint64_t off) {
  int dims[CD], nd = 0;
#pragma unroll
  for (int d = 0; d < CD; d++)
    if (a.info[d][j[d]] & 1u)
      dims[nd++] = d;
  T val[1 << CD];
  for (int c = 0; c < (1 << nd); c++) {
    long long o = (long long)off;
    for (int k = 0; k < nd; k++) {
      const int d = dims[k];
      const uint32_t jj = ((c >> k) & 1) ? j[d] + 1 : j[d] - 1;
      o += ((long long)a.pos[d][jj] - (long long)a.pos[d][j[d]]) * (long long)a.stride[d];
    }
    val[c] = src[o];
  }
  for (int k = 0; k < nd; k++) {
    const int d = dims[k];
    const T xl = a.x[d][j[d] - 1], xm = a.x[d][j[d]], xr = a.x[d][j[d] + 1];
    const T wr = (T)1 / (xr - xl);
    for (int c = 0; c < (1 << (nd - k - 1)); c++)
      val[c] = (T)0 + (val[2 * c] * (xr - xm) + val[2 * c + 1] * (xm - xl)) * wr;
  }
  return val[0];
}

// IDX: type of the linear thread index (uint32_t whenever the box has < 2^32 nodes: the
// index decoding is a chain of divisions)
template <typename T, int OP, typename IDX>
__global__ void __launch_bounds__(256) cpu_level_kernel(const LevelArgs<T> a) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.total)
    return;
  uint32_t j[CD];
  IDX rem = (IDX)t;
  uint64_t off = 0;
  bool allold = true;
  // operations that never look at the "introduced by an earlier level" flag skip its loads
  constexpr bool NEED_OLD = !(OP == OP_PROLONG || OP == OP_MASS || OP == OP_RESTRICT || OP == OP_ADD_OLD ||
                              OP == OP_RECOMP_OLD);
#pragma unroll
  for (int d = CD - 1; d >= 0; d--) {
    const uint32_t c = a.cnt[d];
    uint32_t i = 0;
    if (c > 1) {
      const IDX q = rem / c;
      i = (uint32_t)(rem - q * c);
      rem = q;
    }
    j[d] = a.sel[d] ? a.sel[d][i] : i;
    off += (uint64_t)a.pos[d][j[d]] * a.stride[d];
    if (NEED_OLD)
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
  } else if (OP == OP_COEF_FUSED) {
    // old nodes are only read, new nodes only written: in place
    if (allold) {
      a.buf[off] = 0;
    } else {
      const T r = a.v[off] - nested_interpolation<T>(a, a.v, j, off);
      a.v[off] = r;
      a.buf[off] = r;
    }
  } else if (OP == OP_RECOMP_OLD) {
    // level l-1 box: buf = projection - v (kept for the interpolation), v = -buf
    const T t = a.buf[off] + (T)(-1) * a.v[off];
    a.buf[off] = t;
    a.v[off] = -t;
  } else if (OP == OP_RECOMP_NEW) {
    if (!allold)
      a.v[off] = a.v[off] + (T)(-1) * nested_interpolation<T>(a, a.buf, j, off);
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

// ---- CPU_HUFFMAN_ZSTD payload (reference src/compressors.cpp:70-115,140-181,316-512) ----
// Symbols: q + nql/2 when that lies in (0, nql), else 0 with the value appended to
// the "miss" list; a Huffman tree over the 131072-bin frequency table; codes packed
// MSB-first into 32-bit words.  The frequency table and the packed words are built
// on the GPU; the tree (a few hundred nodes) on the host with the reference's own
// container (std::priority_queue) so that ties break identically.
constexpr int NQL = 32768 * 4;
constexpr int HCH = 4096; // symbols per block in the packing kernels
constexpr int HTH = 256;
constexpr int HPER = HCH / HTH;

__device__ __forceinline__ int shifted_symbol(long long q, bool *hit) {
  // build_ft (:140-160) tests the 64-bit value, huffman_encoding (:343-360) the
  // value truncated to int; both are kept
  const int qi = (int)(q + NQL / 2);
  *hit = qi > 0 && qi < NQL;
  return qi;
}

__global__ void __launch_bounds__(256) cpuhuff_hist_kernel(const long long *__restrict__ q, uint64_t n,
                                                           unsigned *__restrict__ hist) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < n; base += stride) {
    const uint64_t i = base + threadIdx.x;
    const bool active = i < n;
    int bin = 0;
    if (active) {
      const long long v = q[i] + NQL / 2;
      bin = (v > 0 && v < NQL) ? (int)v : 0;
    }
    // one atomic per distinct bin in the warp (the data are mostly one value)
    const unsigned act = __ballot_sync(0xffffffffu, active);
    if (active) {
      const unsigned peers = __match_any_sync(act, bin);
      if ((threadIdx.x & 31) == __ffs(peers) - 1)
        atomicAdd(&hist[bin], (unsigned)__popc(peers));
    }
  }
}

// per block of HCH symbols: total code bits and number of misses
__global__ void __launch_bounds__(HTH) cpuhuff_count_kernel(const long long *__restrict__ q, uint64_t n,
                                                            const unsigned *__restrict__ code_len,
                                                            unsigned long long *__restrict__ blk_bits,
                                                            unsigned *__restrict__ blk_miss) {
  __shared__ unsigned s_bits[HTH / 32], s_miss[HTH / 32];
  const uint64_t first = (uint64_t)blockIdx.x * HCH + (uint64_t)threadIdx.x * HPER;
  unsigned bits = 0, miss = 0;
  for (int k = 0; k < HPER; k++) {
    const uint64_t i = first + k;
    if (i < n) {
      bool hit;
      const int qi = shifted_symbol(q[i], &hit);
      bits += code_len[hit ? qi : 0];
      miss += hit ? 0u : 1u;
    }
  }
  for (int o = 16; o; o >>= 1) {
    bits += __shfl_down_sync(0xffffffffu, bits, o);
    miss += __shfl_down_sync(0xffffffffu, miss, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_bits[threadIdx.x >> 5] = bits;
    s_miss[threadIdx.x >> 5] = miss;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long b = 0;
    unsigned m = 0;
    for (int w = 0; w < HTH / 32; w++) {
      b += s_bits[w];
      m += s_miss[w];
    }
    blk_bits[blockIdx.x] = b;
    blk_miss[blockIdx.x] = m;
  }
}

// exclusive scan of the per-block totals (one block; the arrays are N / 4096 long)
__global__ void __launch_bounds__(1024) cpuhuff_scan_kernel(unsigned long long *blk_bits, unsigned *blk_miss,
                                                            unsigned long long *blk_miss_off, uint64_t nblk,
                                                            unsigned long long *totals) {
  __shared__ unsigned long long s_b[1024], s_m[1024];
  unsigned long long carry_b = 0, carry_m = 0;
  for (uint64_t base = 0; base < nblk; base += 1024) {
    const uint64_t i = base + threadIdx.x;
    const unsigned long long b = i < nblk ? blk_bits[i] : 0, m = i < nblk ? blk_miss[i] : 0;
    s_b[threadIdx.x] = b;
    s_m[threadIdx.x] = m;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      unsigned long long tb = 0, tm = 0;
      if ((int)threadIdx.x >= o) {
        tb = s_b[threadIdx.x - o];
        tm = s_m[threadIdx.x - o];
      }
      __syncthreads();
      s_b[threadIdx.x] += tb;
      s_m[threadIdx.x] += tm;
      __syncthreads();
    }
    if (i < nblk) {
      blk_bits[i] = carry_b + s_b[threadIdx.x] - b;
      blk_miss_off[i] = carry_m + s_m[threadIdx.x] - m;
    }
    carry_b += s_b[1023];
    carry_m += s_m[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    totals[0] = carry_b;
    totals[1] = carry_m;
  }
}

// huffman_encoding's packing loop (:362-384): MSB-first into zeroed 32-bit words
__global__ void __launch_bounds__(HTH) cpuhuff_pack_kernel(const long long *__restrict__ q, uint64_t n,
                                                           const unsigned *__restrict__ code_len,
                                                           const unsigned *__restrict__ code_bits,
                                                           const unsigned long long *__restrict__ blk_bits,
                                                           const unsigned long long *__restrict__ blk_miss_off,
                                                           unsigned *__restrict__ words, int *__restrict__ misses, int *flag) {
  __shared__ unsigned s_bits[HTH / 32], s_miss[HTH / 32];
  const uint64_t first = (uint64_t)blockIdx.x * HCH + (uint64_t)threadIdx.x * HPER;
  unsigned len[HPER], code[HPER];
  int missv[HPER];
  unsigned bits = 0, miss = 0, missmask = 0;
  for (int k = 0; k < HPER; k++) {
    const uint64_t i = first + k;
    len[k] = 0;
    code[k] = 0;
    if (i < n) {
      bool hit;
      const int qi = shifted_symbol(q[i], &hit);
      // the reference keeps misses as int (:343,356): a wider value cannot be stored
      if ((long long)qi != q[i] + NQL / 2)
        *flag = 1;
      const int sym = hit ? qi : 0;
      len[k] = code_len[sym];
      code[k] = code_bits[sym];
      if (!hit) {
        missv[k] = qi;
        missmask |= 1u << k;
        miss++;
      }
    }
    bits += len[k];
  }
  // exclusive scan over the threads of the block
  unsigned ib = bits, im = miss;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned tb = __shfl_up_sync(0xffffffffu, ib, o), tm = __shfl_up_sync(0xffffffffu, im, o);
    if (lane >= o) {
      ib += tb;
      im += tm;
    }
  }
  if (lane == 31) {
    s_bits[warp] = ib;
    s_miss[warp] = im;
  }
  __syncthreads();
  unsigned wb = 0, wm = 0;
  for (int w = 0; w < warp; w++) {
    wb += s_bits[w];
    wm += s_miss[w];
  }
  unsigned long long pos = blk_bits[blockIdx.x] + wb + (ib - bits);
  unsigned long long mpos = blk_miss_off[blockIdx.x] + wm + (im - miss);
  for (int k = 0; k < HPER; k++) {
    const unsigned l = len[k];
    if (l) {
      const unsigned long long w = pos >> 5;
      const unsigned room = 32 - (unsigned)(pos & 31);
      if (room < l) {
        const unsigned rshift = l - room;
        atomicOr(&words[w], code[k] >> rshift);
        atomicOr(&words[w + 1], code[k] << (32 - rshift));
      } else {
        atomicOr(&words[w], code[k] << (room - l));
      }
      pos += l;
    }
    if ((missmask >> k) & 1u)
      misses[mpos++] = missv[k];
  }
}

struct HuffNode {
  int q;
  size_t cnt;
  int left, right;
};

// build_tree + build_codec (:70-115): returns false when a code would not fit
// the reference's 32-bit code word
bool build_cpu_huffman(const std::vector<size_t> &ft, std::vector<unsigned> &code, std::vector<unsigned> &len,
                       std::vector<HuffNode> &nodes, int &root) {
  struct ByCount { // LessThanByCnt (:59-63)
    const std::vector<HuffNode> *nodes;
    bool operator()(int a, int b) const { return (*nodes)[a].cnt > (*nodes)[b].cnt; }
  };
  nodes.clear();
  nodes.reserve(2 * 4096);
  std::priority_queue<int, std::vector<int>, ByCount> pq(ByCount{&nodes});
  for (int i = 0; i < NQL; i++)
    if (ft[i] != 0) {
      nodes.push_back(HuffNode{i, ft[i], -1, -1});
      pq.push((int)nodes.size() - 1);
    }
  code.assign(NQL, 0);
  len.assign(NQL, 0);
  root = -1;
  if (pq.empty())
    return true;
  while (pq.size() > 1) {
    const int a = pq.top();
    pq.pop();
    const int b = pq.top();
    pq.pop();
    nodes.push_back(HuffNode{-1, nodes[a].cnt + nodes[b].cnt, a, b});
    pq.push((int)nodes.size() - 1);
  }
  root = pq.top();
  // depth-first code assignment: left = 0, right = 1
  std::vector<std::pair<int, std::pair<unsigned, unsigned>>> stack;
  stack.push_back({root, {0u, 0u}});
  bool ok = true;
  while (!stack.empty()) {
    const auto cur = stack.back();
    stack.pop_back();
    const HuffNode &nd = nodes[cur.first];
    if (nd.left < 0 && nd.right < 0) {
      code[nd.q] = cur.second.first;
      len[nd.q] = cur.second.second;
      ok = ok && cur.second.second <= 32;
      continue;
    }
    stack.push_back({nd.right, {cur.second.first << 1 | 1u, cur.second.second + 1}});
    stack.push_back({nd.left, {cur.second.first << 1, cur.second.second + 1}});
  }
  return ok;
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
  if (a.total <= 0xffffffffull)
    MGB_LAUNCH(MGB_K_AXPY, st, (cpu_level_kernel<T, OP, uint32_t><<<(unsigned)blocks, 256, 0, st>>>(a)));
  else
    MGB_LAUNCH(MGB_K_AXPY, st, (cpu_level_kernel<T, OP, uint64_t><<<(unsigned)blocks, 256, 0, st>>>(a)));
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

// One pass instead of copy / D prolongation passes / subtract (MGB_CPU_UNFUSED=1: the
// reference's pass-by-pass sequence; both give the same bits)
inline bool fused_passes() {
  static const bool on = getenv("MGB_CPU_UNFUSED") == nullptr;
  return on;
}

// mgard::decompose on the nodal array `v` (decompose.tpp:129-174)
template <typename T> void decompose_nodal(mgb_cpu_plan *p, T *v, cudaStream_t st) {
  T *b0 = reinterpret_cast<T *>(p->d_b0), *b1 = reinterpret_cast<T *>(p->d_b1);
  for (int l = p->L; l > 0; l--) {
    LevelArgs<T> a = level_args<T>(p, l, -1, SEL_ALL);
    a.v = v;
    a.buf = b0;
    if (fused_passes()) {
      launch_level<T, OP_COEF_FUSED>(a, st);
    } else {
      launch_level<T, OP_COPY_OLD_ZERO_NEW>(a, st);
      for (int d = 0; d < CD; d++) {
        if ((p->flat >> d) & 1)
          continue;
        LevelArgs<T> pa = level_args<T>(p, l, d, SEL_NEW);
        pa.buf = b0;
        launch_level<T, OP_PROLONG>(pa, st);
      }
      launch_level<T, OP_SUB_NEW>(a, st);
    }
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
    if (fused_passes()) {
      LevelArgs<T> c = level_args<T>(p, l - 1, -1, SEL_ALL);
      c.v = v;
      c.buf = corr;
      launch_level<T, OP_RECOMP_OLD>(c, st);
      launch_level<T, OP_RECOMP_NEW>(a, st);
      continue;
    }
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


struct DeviceBuf {
  void *p = nullptr;
  ~DeviceBuf() { cudaFree(p); }
  int alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1) == cudaSuccess ? MGB_SUCCESS : MGB_CUDA_ERROR; }
  template <typename U> U *as() { return static_cast<U *>(p); }
};

// compress_memory_z (compressors.cpp:552-606): one deflate stream, level 9
int zlib_payload(const void *src, size_t src_bytes, std::vector<uint8_t> &out) {
  // the reference feeds the whole buffer through a 32-bit avail_in (:560)
  if (src_bytes > 0xffffffffull)
    return MGB_OUTPUT_TOO_LARGE;
  z_stream strm;
  memset(&strm, 0, sizeof(strm));
  if (deflateInit(&strm, Z_BEST_COMPRESSION) != Z_OK)
    return MGB_FAILURE;
  const size_t bound = deflateBound(&strm, (uLong)src_bytes);
  if (bound > 0xffffffffull) {
    deflateEnd(&strm);
    return MGB_OUTPUT_TOO_LARGE;
  }
  out.resize(bound);
  strm.next_in = const_cast<Bytef *>(static_cast<const Bytef *>(src));
  strm.avail_in = (uInt)src_bytes;
  strm.next_out = out.data();
  strm.avail_out = (uInt)bound;
  const int zr = deflate(&strm, Z_FINISH);
  out.resize(bound - strm.avail_out);
  deflateEnd(&strm);
  return zr == Z_STREAM_END ? MGB_SUCCESS : MGB_FAILURE;
}

// compress_memory_huffman, MGARD_ZSTD build (compressors.cpp:421-512): the
// frequency table and the bit packing run on the GPU over p->d_q
int huffman_zstd_payload(mgb_cpu_plan *p, cudaStream_t st, std::vector<uint8_t> &out) {
  const mgb_zstd_fns &z = mgb_zstd();
  if (!z.ok)
    return MGB_FAILURE; // libzstd.so.1 not found: the request is never silently downgraded
  const uint64_t n = p->N;
  if (n >= (1ull << 32))
    return MGB_OUTPUT_TOO_LARGE;
  const uint64_t nblk = (n + HCH - 1) / HCH;
  DeviceBuf hist, clen, cbits, bbits, bmiss, bmoff, totals, words, misses;
  if (hist.alloc(NQL * 4) || clen.alloc(NQL * 4) || cbits.alloc(NQL * 4) || bbits.alloc(nblk * 8) ||
      bmiss.alloc(nblk * 4) || bmoff.alloc(nblk * 8) || totals.alloc(16))
    return MGB_CUDA_ERROR;
  MGB_CUDA_CHECK(cudaMemsetAsync(hist.p, 0, NQL * 4, st));
  MGB_LAUNCH(MGB_K_CODEBOOK, st,
             (cpuhuff_hist_kernel<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 16), 256, 0, st>>>(
                 p->d_q, n, hist.as<unsigned>())));
  std::vector<unsigned> h32(NQL);
  MGB_CUDA_CHECK(cudaMemcpyAsync(h32.data(), hist.p, NQL * 4, cudaMemcpyDeviceToHost, st));
  MGB_CUDA_CHECK(cudaStreamSynchronize(st));
  std::vector<size_t> ft(h32.begin(), h32.end());
  std::vector<unsigned> code, len;
  std::vector<HuffNode> nodes;
  int root = -1;
  if (!build_cpu_huffman(ft, code, len, nodes, root))
    return MGB_FAILURE; // a code longer than the reference's 32-bit code word
  MGB_CUDA_CHECK(cudaMemcpyAsync(clen.p, len.data(), NQL * 4, cudaMemcpyHostToDevice, st));
  MGB_CUDA_CHECK(cudaMemcpyAsync(cbits.p, code.data(), NQL * 4, cudaMemcpyHostToDevice, st));
  MGB_LAUNCH(MGB_K_CHUNK_BITS, st,
             (cpuhuff_count_kernel<<<(unsigned)nblk, HTH, 0, st>>>(p->d_q, n, clen.as<unsigned>(),
                                                                  bbits.as<unsigned long long>(), bmiss.as<unsigned>())));
  MGB_LAUNCH(MGB_K_CHUNK_SCAN, st,
             (cpuhuff_scan_kernel<<<1, 1024, 0, st>>>(bbits.as<unsigned long long>(), bmiss.as<unsigned>(),
                                                      bmoff.as<unsigned long long>(), nblk,
                                                      totals.as<unsigned long long>())));
  unsigned long long tot[2] = {0, 0};
  MGB_CUDA_CHECK(cudaMemcpyAsync(tot, totals.p, 16, cudaMemcpyDeviceToHost, st));
  MGB_CUDA_CHECK(cudaStreamSynchronize(st));
  const size_t hit_bits = tot[0], nmiss = tot[1];
  const size_t hit_bytes = hit_bits / 8 + 4; // as the reference copies (:452-453)
  const size_t word_bytes = ((hit_bits + 31) / 32 + 2) * 4;
  if (words.alloc(word_bytes) || misses.alloc(nmiss * 4))
    return MGB_CUDA_ERROR;
  MGB_CUDA_CHECK(cudaMemsetAsync(words.p, 0, word_bytes, st));
  MGB_LAUNCH(MGB_K_ENCODE, st,
             (cpuhuff_pack_kernel<<<(unsigned)nblk, HTH, 0, st>>>(
                 p->d_q, n, clen.as<unsigned>(), cbits.as<unsigned>(), bbits.as<unsigned long long>(),
                 bmoff.as<unsigned long long>(), words.as<unsigned>(), misses.as<int>(), p->d_flag)));
  {
    const int frc = check_flag(p, st);
    if (frc)
      return frc;
  }
  // payload = frequency pairs | packed codes | misses (:447-463)
  size_t nonzero = 0;
  for (int i = 0; i < NQL; i++)
    nonzero += ft[i] > 0;
  const size_t tree_bytes = 2 * nonzero * sizeof(size_t), miss_bytes = nmiss * sizeof(int);
  std::vector<uint8_t> payload(tree_bytes + hit_bytes + miss_bytes);
  {
    size_t *cft = reinterpret_cast<size_t *>(payload.data());
    size_t off = 0;
    for (int i = 0; i < NQL; i++)
      if (ft[i] > 0) {
        cft[2 * off] = (size_t)i;
        cft[2 * off + 1] = ft[i];
        off++;
      }
  }
  MGB_CUDA_CHECK(cudaMemcpyAsync(payload.data() + tree_bytes, words.p, hit_bytes, cudaMemcpyDeviceToHost, st));
  if (miss_bytes)
    MGB_CUDA_CHECK(cudaMemcpyAsync(payload.data() + tree_bytes + hit_bytes, misses.p, miss_bytes,
                                   cudaMemcpyDeviceToHost, st));
  MGB_CUDA_CHECK(cudaStreamSynchronize(st));
  // compress_memory_zstd (:542-549): level 1; then the three sizes in front (:494-511)
  const size_t bound = z.bound(payload.size());
  out.resize(3 * sizeof(size_t) + bound);
  const size_t csize = z.compress(out.data() + 3 * sizeof(size_t), bound, payload.data(), payload.size(), 1);
  if (z.is_error(csize))
    return MGB_FAILURE;
  size_t head[3] = {tree_bytes, hit_bits, miss_bytes};
  memcpy(out.data(), head, sizeof(head));
  out.resize(3 * sizeof(size_t) + csize);
  return MGB_SUCCESS;
}

// decompress_memory_huffman + huffman_decoding (compressors.cpp:183-314): the
// stream carries no block index, so the bit-serial walk runs on the host
int huffman_zstd_decode(const uint8_t *src, size_t src_bytes, std::vector<long long> &q) {
  const mgb_zstd_fns &z = mgb_zstd();
  if (!z.ok)
    return MGB_FAILURE;
  if (src_bytes < 3 * sizeof(size_t))
    return MGB_BAD_STREAM;
  size_t head[3];
  memcpy(head, src, sizeof(head));
  const size_t tree_bytes = head[0], hit_bits = head[1], miss_bytes = head[2];
  if (tree_bytes % (2 * sizeof(size_t)) || miss_bytes % sizeof(int) || tree_bytes > (size_t)NQL * 16 ||
      hit_bits / 8 > q.size() * 4 + 8 || miss_bytes > q.size() * 4)
    return MGB_BAD_STREAM;
  const size_t hit_bytes = hit_bits / 8 + 4;
  std::vector<uint8_t> payload(tree_bytes + hit_bytes + miss_bytes);
  const size_t got = z.decompress(payload.data(), payload.size(), src + 3 * sizeof(size_t), src_bytes - 3 * sizeof(size_t));
  if (z.is_error(got) || got != payload.size())
    return MGB_BAD_STREAM;
  std::vector<size_t> ft(NQL, 0);
  {
    std::vector<size_t> cft(tree_bytes / sizeof(size_t));
    memcpy(cft.data(), payload.data(), tree_bytes);
    for (size_t j = 0; j + 1 < cft.size(); j += 2) {
      if (cft[j] >= (size_t)NQL)
        return MGB_BAD_STREAM;
      ft[cft[j]] = cft[j + 1];
    }
  }
  std::vector<unsigned> code, len;
  std::vector<HuffNode> nodes;
  int root = -1;
  build_cpu_huffman(ft, code, len, nodes, root);
  if (root < 0)
    return q.empty() ? MGB_SUCCESS : MGB_BAD_STREAM;
  std::vector<unsigned> words((hit_bytes + 3) / 4 + 3, 0);
  memcpy(words.data(), payload.data() + tree_bytes, hit_bytes);
  std::vector<int> miss(miss_bytes / sizeof(int));
  if (miss_bytes)
    memcpy(miss.data(), payload.data() + tree_bytes + hit_bytes, miss_bytes);
  // the reference walks the tree bit by bit (:211-229); same walk, with the first
  // LUT_BITS steps of every symbol taken from a table
  constexpr int LUT_BITS = 12;
  struct Step {
    int node;
    int used;
  };
  std::vector<Step> lut(1u << LUT_BITS);
  for (unsigned v = 0; v < (1u << LUT_BITS); v++) {
    int node = root, used = 0;
    while (nodes[node].left >= 0 && used < LUT_BITS) {
      node = (v >> (LUT_BITS - 1 - used)) & 1u ? nodes[node].right : nodes[node].left;
      used++;
    }
    lut[v] = Step{node, used};
  }
  size_t bit = 0, next_miss = 0;
  for (size_t i = 0; i < q.size(); i++) {
    const size_t w = bit >> 5;
    const unsigned long long window = ((unsigned long long)words[w] << 32) | words[w + 1];
    const unsigned v = (unsigned)(window >> (64 - LUT_BITS - (bit & 31))) & ((1u << LUT_BITS) - 1);
    int node = lut[v].node;
    bit += lut[v].used;
    while (nodes[node].left >= 0) {
      if (bit >= hit_bits)
        return MGB_BAD_STREAM;
      const unsigned flag = words[bit >> 5] & (0x80000000u >> (bit & 31));
      node = flag ? nodes[node].right : nodes[node].left;
      bit++;
    }
    if (bit > hit_bits)
      return MGB_BAD_STREAM;
    if (nodes[node].q != 0) {
      q[i] = (long long)nodes[node].q - NQL / 2;
    } else {
      if (next_miss >= miss.size())
        return MGB_BAD_STREAM;
      q[i] = (long long)miss[next_miss++] - NQL / 2;
    }
  }
  return bit == hit_bits && next_miss == miss.size() ? MGB_SUCCESS : MGB_BAD_STREAM;
}

} // namespace

namespace {
// One constituent operator on every line of level `level` along `dim`, on a NODAL array
// in place (the reference applies them to a shuffled array: shuffle, operator,
// unshuffle give the same nodal result - tests/src/test_TensorMassMatrix.cpp,
// test_TensorRestriction.cpp, test_TensorProlongation.cpp).
template <typename T> int apply_operator_t(mgb_cpu_plan *p, int op, int l, int d, T *v, cudaStream_t st) {
  const T *real = reinterpret_cast<const T *>(p->d_real);
  if (op == MGB_CPU_OP_MASS) {
    // out of place: the nodes outside the level keep their values
    T *tmp = reinterpret_cast<T *>(p->d_b0);
    MGB_CUDA_CHECK(cudaMemcpyAsync(tmp, v, p->N * sizeof(T), cudaMemcpyDeviceToDevice, st));
    LevelArgs<T> a = level_args<T>(p, l, d, SEL_ALL);
    a.src = v;
    a.dst = tmp;
    launch_level<T, OP_MASS>(a, st);
    MGB_CUDA_CHECK(cudaMemcpyAsync(v, tmp, p->N * sizeof(T), cudaMemcpyDeviceToDevice, st));
  } else if (op == MGB_CPU_OP_RESTRICTION) {
    LevelArgs<T> a = level_args<T>(p, l, d, SEL_OLD);
    a.buf = v;
    launch_level<T, OP_RESTRICT>(a, st);
  } else if (op == MGB_CPU_OP_PROLONGATION_ADDITION) {
    LevelArgs<T> a = level_args<T>(p, l, d, SEL_NEW);
    a.buf = v;
    launch_level<T, OP_PROLONG>(a, st);
  } else if (op == MGB_CPU_OP_MASS_INVERSE) {
    LevelArgs<T> a = level_args<T>(p, l, d, SEL_ONE);
    a.buf = v;
    const DimTables<double> &t = p->tab[l][d];
    const uint64_t blocks = (a.total + 127) / 128;
    MGB_LAUNCH(MGB_K_THOMAS_STRIDED, st,
               (cpu_thomas_kernel<T><<<(unsigned)blocks, 128, 0, st>>>(a, real + t.w, real + t.dv, real + t.cc)));
  } else {
    return MGB_BAD_ARGUMENT;
  }
  MGB_CUDA_CHECK(cudaGetLastError());
  return MGB_SUCCESS;
}

} // namespace

extern "C" {

static int cpu_plan_create_impl(int ndim, const uint64_t *shape, int dtype, const void *const *coords,
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
    if (p->N > (1ull << 40) / shape[d]) { // also keeps every product below 2^63
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

int mgb_cpu_apply_operator(mgb_cpu_plan *p, int op, int level, int dim, void *d_nodal, void *stream) {
  if (!p || !d_nodal || level < 0 || level > p->L || dim < 0 || dim >= p->ndim)
    return MGB_BAD_ARGUMENT;
  // restriction / prolongation relate level l to l - 1 (the reference constructors throw for l = 0)
  if ((op == MGB_CPU_OP_RESTRICTION || op == MGB_CPU_OP_PROLONGATION_ADDITION) && level == 0)
    return MGB_BAD_ARGUMENT;
  int rc = ensure_workspace(p, false);
  if (rc)
    return rc;
  const int d = dim + (CD - p->ndim); // user dimensions are right-aligned in the CD-dimensional plan
  if ((p->flat >> d) & 1)
    return MGB_SUCCESS; // a dimension of size 1: every operator is the identity
  cudaStream_t st = (cudaStream_t)stream;
  if (p->dtype == MGB_F32)
    return apply_operator_t<float>(p, op, level, d, (float *)d_nodal, st);
  return apply_operator_t<double>(p, op, level, d, (double *)d_nodal, st);
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
static int cpu_compress_impl(int ndim, int dtype, const uint64_t *shape, const void *const *coords, double s,
                             double tol, int compressor, const void *in, void **out, size_t *out_size) {
  if (!in || !out || !out_size || !(tol > 0) || (compressor != 1 && compressor != 2))
    return MGB_BAD_ARGUMENT;
  mgb_cpu_plan *p = nullptr;
  int rc = mgb_cpu_plan_create(ndim, shape, dtype, coords, &p);
  if (rc)
    return rc;
  struct Guard {
    mgb_cpu_plan *p;
    ~Guard() { delete p; }
  } guard{p};
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
  std::vector<uint8_t> payload;
  if (compressor == 2) {
    rc = huffman_zstd_payload(p, st, payload);
  } else {
    std::vector<long long> q(p->N);
    MGB_CUDA_CHECK(cudaMemcpy(q.data(), p->d_q, p->N * sizeof(long long), cudaMemcpyDeviceToHost));
    rc = zlib_payload(q.data(), q.size() * sizeof(long long), payload);
  }
  if (rc)
    return rc;

  mgb_header h;
  h.convention = 1;
  h.cpu_compressor = compressor;
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
  uint8_t *buf = (uint8_t *)malloc(head.size() + payload.size());
  if (!buf)
    return MGB_FAILURE;
  memcpy(buf, head.data(), head.size());
  memcpy(buf + head.size(), payload.data(), payload.size());
  *out = buf;
  *out_size = head.size() + payload.size();
  return MGB_SUCCESS;
}

// mgard::decompress(void const *, size_t) (compress.tpp:69-83 behind the header
// dispatch of include/compress.hpp:62-72).  *out is malloc'ed (host).
static int cpu_decompress_impl(const void *in, size_t in_size, void **out, int *ndim, uint64_t *shape,
                               int *dtype) {
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
  std::vector<long long> q(p->N);
  if (h.cpu_compressor == 2) {
    rc = huffman_zstd_decode((const uint8_t *)in + hb, in_size - hb, q);
    if (rc)
      return rc;
  } else { // decompress_memory_z (compressors.cpp:608-629)
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

// No exception crosses the ABI (a corrupt header may ask for an absurd allocation).
int mgb_cpu_plan_create(int ndim, const uint64_t *shape, int dtype, const void *const *coords,
                        mgb_cpu_plan **plan) {
  try {
    return cpu_plan_create_impl(ndim, shape, dtype, coords, plan);
  } catch (const std::exception &) {
    return MGB_FAILURE;
  }
}

int mgb_cpu_compress(int ndim, int dtype, const uint64_t *shape, const void *const *coords, double s, double tol,
                     int compressor, const void *in, void **out, size_t *out_size) {
  try {
    return cpu_compress_impl(ndim, dtype, shape, coords, s, tol, compressor, in, out, out_size);
  } catch (const std::exception &) {
    return MGB_FAILURE;
  }
}

int mgb_cpu_write_header(int ndim, int dtype, const uint64_t *shape, const void *const *coords, double s,
                         double tol, int compressor, uint8_t *out, uint64_t cap, uint64_t *size) {
  if (!shape || !out || !size || ndim < 1 || ndim > MGB_MAX_DIMS || (dtype != MGB_F32 && dtype != MGB_F64) ||
      (compressor != 1 && compressor != 2))
    return MGB_BAD_ARGUMENT;
  try {
    mgb_header h;
    h.convention = 1;
    h.cpu_compressor = compressor;
    h.ndim = ndim;
    h.dtype = dtype;
    h.ebtype = MGB_ABS;
    h.s = dtype == MGB_F32 ? (double)(float)s : s;
    h.tol = dtype == MGB_F32 ? (double)(float)tol : tol;
    for (int d = 0; d < ndim; d++) {
      if (shape[d] == 0 || shape[d] >= (1ull << 31))
        return MGB_BAD_ARGUMENT;
      h.shape[d] = shape[d];
    }
    if (coords) {
      h.coords.resize(ndim);
      for (int d = 0; d < ndim; d++) {
        h.coords[d].resize(shape[d]);
        for (uint64_t i = 0; i < shape[d]; i++)
          h.coords[d][i] = dtype == MGB_F32 ? (double)((const float *)coords[d])[i] : ((const double *)coords[d])[i];
      }
    }
    const std::vector<uint8_t> head = mgb_encode_stream_header(h);
    *size = head.size();
    if (head.size() > cap)
      return MGB_OUTPUT_TOO_LARGE;
    memcpy(out, head.data(), head.size());
    return MGB_SUCCESS;
  } catch (const std::exception &) {
    return MGB_FAILURE;
  }
}

int mgb_cpu_decompress(const void *in, size_t in_size, void **out, int *ndim, uint64_t *shape, int *dtype) {
  try {
    return cpu_decompress_impl(in, in_size, out, ndim, shape, dtype);
  } catch (const std::exception &) {
    return MGB_FAILURE;
  }
}

} // extern "C"
