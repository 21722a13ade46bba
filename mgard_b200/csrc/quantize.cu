// Norms, level-weighted quantizer fused with the Huffman histogram, and the
// inverse (sm_100a).  Compile with -fmad=false (bit-exact contract).
//
//   norm        reference norm_calculator            (CompressionLowLevel/NormCalculator.hpp:13-83)
//   quantizers  reference LinearQuantizer::CalcQuantizers (Quantization/LinearQuantization.hpp:495-545)
//   quantize    reference LevelwiseLinearQuantizerKernel<QUANTIZE>   (:148-248)
//               + HistogramKernel (Lossless/ParallelHuffman/Histogram.hpp:15-118)
//   dequantize  reference OutlierRestoreKernel + <DEQUANTIZE>       (:251-264,304-350)
//
// Differences in shape, not in results: symbols are 16 bit (dictionary <= 65536)
// instead of int64, the histogram is accumulated by the same pass that
// quantizes, and nothing round-trips through the host.
#include <cmath>
#include <cstring>
#include <limits>

#include "plan.h"

namespace {

typedef long long i64;

struct QParams {
  int D;
  int dict;
  int calc_level;
  int L;
  unsigned n[5];
  unsigned rows;
  const int *marks;
  unsigned long long marks_width;
  // per level: quantizer (reciprocal for quantize), volume factor
  double q[MGB_MAX_LEVELS];
  double vol[MGB_MAX_LEVELS];
};

template <typename T> struct Tables {
  T q[MGB_MAX_LEVELS];
  T vol[MGB_MAX_LEVELS];
};

template <typename T> __device__ __forceinline__ T absT(T x) { return fabs(x); }
template <> __device__ __forceinline__ float absT<float>(float x) { return fabsf(x); }

template <typename T>
__device__ __forceinline__ long long quantize_one(T t, T q, T vol, int dict) {
  // LinearQuantization.hpp:203-208
  T x = copysign((T)0.5 + absT(t * q * vol), t);
  long long qi = (long long)x;
  return qi + dict / 2;
}

// shared-memory histogram add with warp aggregation of equal bins
__device__ __forceinline__ void hist_add(unsigned *sh, unsigned bin, bool valid) {
  unsigned active = __ballot_sync(0xffffffffu, valid);
  if (!valid)
    return;
  unsigned peers = __match_any_sync(active, bin);
  int leader = __ffs(peers) - 1;
  if ((threadIdx.x & 31) == leader)
    atomicAdd(&sh[bin], (unsigned)__popc(peers));
}

// warp-aggregated append of outliers (one atomic per warp; ascending index
// order inside a warp).  Must be called by all 32 lanes.
__device__ __forceinline__ void
outlier_append(bool is_out, unsigned long long idx, long long qi,
               unsigned long long *__restrict__ ocount, uint64_t *__restrict__ oidx,
               i64 *__restrict__ oval, unsigned long long ocap) {
  unsigned m = __ballot_sync(0xffffffffu, is_out);
  if (m == 0)
    return;
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  unsigned long long base = 0;
  if (lane == leader)
    base = atomicAdd(ocount, (unsigned long long)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (is_out) {
    unsigned long long k = base + __popc(m & ((1u << lane) - 1));
    if (k < ocap) {
      oidx[k] = idx;
      oval[k] = qi;
    }
  }
}

// vector of 16 bytes of T
template <typename T> struct Vec16;
template <> struct Vec16<float> { typedef float4 type; static constexpr int n = 4; };
template <> struct Vec16<double> { typedef double2 type; static constexpr int n = 2; };

// s = inf: one quantizer for every node (LinearQuantization.hpp:170-176), so the
// array is a flat stream.  Each thread handles PER consecutive elements per
// iteration through 128-bit loads (two vectors in flight for fp32, four for
// fp64) and stores its symbols with one 128-bit store.  Histogram: shared memory,
// one atomic per run of equal symbols inside a thread, a single atomic per warp
// when all 32 lanes agree (smooth data); outliers take a ballot-guarded slow path.
template <typename T>
__global__ void __launch_bounds__(256, 4)
quantize_linear_kernel(const T *__restrict__ v, i64 N, T q, T vol, int dict,
                       uint16_t *__restrict__ sym, unsigned *__restrict__ ghist,
                       unsigned long long *__restrict__ ocount,
                       uint64_t *__restrict__ oidx, i64 *__restrict__ oval,
                       unsigned long long ocap, unsigned long long base,
                       const T *__restrict__ qdev) {
  // v / sym point at element `base` of the array (outlier indices are global)
  // qdev: the quantizer table lives in device memory (prepare_q_kernel)
  extern __shared__ unsigned sh[];
  if (qdev)
    q = qdev[0];
  typedef typename Vec16<T>::type V;
  constexpr int VN = Vec16<T>::n, PER = 8, NV = PER / VN;
  for (int i = threadIdx.x; i < dict; i += blockDim.x)
    sh[i] = 0;
  __syncthreads();
  // full groups of PER elements (none when an array is not 16-byte aligned)
  const bool aligned = (((uintptr_t)v | (uintptr_t)sym) & 15) == 0;
  const i64 ngroups = aligned ? N / PER : 0;
  const i64 gstride = (i64)gridDim.x * blockDim.x;
  const i64 ground = (ngroups + gstride - 1) / gstride * gstride;
  for (i64 gi = (i64)blockIdx.x * blockDim.x + threadIdx.x; gi < ground; gi += gstride) {
    const bool valid = gi < ngroups;
    unsigned s[PER];
    long long qi[PER];
    bool any_out = false;
    if (valid) {
      V raw[NV];
#pragma unroll
      for (int k = 0; k < NV; k++)
        raw[k] = __ldcs(reinterpret_cast<const V *>(v + gi * PER) + k);
      const T *x = reinterpret_cast<const T *>(raw);
#pragma unroll
      for (int k = 0; k < PER; k++) {
        qi[k] = quantize_one<T>(x[k], q, vol, dict);
        const bool in = qi[k] >= 0 && qi[k] < dict;
        s[k] = in ? (unsigned)qi[k] : 0u;
        any_out |= !in;
      }
      uint4 pk;
      pk.x = s[0] | (s[1] << 16);
      pk.y = s[2] | (s[3] << 16);
      pk.z = s[4] | (s[5] << 16);
      pk.w = s[6] | (s[7] << 16);
      reinterpret_cast<uint4 *>(sym)[gi] = pk;
    }
    if (__any_sync(0xffffffffu, any_out)) {
#pragma unroll
      for (int k = 0; k < PER; k++) {
        const bool o = valid && !(qi[k] >= 0 && qi[k] < dict);
        outlier_append(o, base + (unsigned long long)(gi * PER + k), valid ? qi[k] : 0, ocount, oidx,
                       oval, ocap);
      }
    }
    // histogram
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      unsigned run = 1;
      int allsame = 0;
#pragma unroll
      for (int k = 1; k < PER; k++)
        allsame += s[k] == s[0];
      int pred = 0;
      if (act == 0xffffffffu)
        __match_all_sync(0xffffffffu, allsame == PER - 1 ? s[0] : 0xffffffffu - (threadIdx.x & 31), &pred);
      if (pred) {
        if ((threadIdx.x & 31) == 0)
          atomicAdd(&sh[s[0]], 32u * PER);
      } else {
#pragma unroll
        for (int k = 1; k < PER; k++) {
          if (s[k] == s[k - 1]) {
            run++;
          } else {
            atomicAdd(&sh[s[k - 1]], run);
            run = 1;
          }
        }
        atomicAdd(&sh[s[PER - 1]], run);
      }
    }
  }
  // remainder (and everything, when the arrays are not 16-byte aligned): scalar
  for (i64 i0 = ngroups * PER + (i64)blockIdx.x * blockDim.x; i0 < N; i0 += gstride) {
    const i64 i = i0 + threadIdx.x;
    const bool valid = i < N;
    long long q1 = 0;
    unsigned s1 = 0;
    bool o = false;
    if (valid) {
      q1 = quantize_one<T>(v[i], q, vol, dict);
      if (q1 >= 0 && q1 < dict)
        s1 = (unsigned)q1;
      else
        o = true;
      sym[i] = (uint16_t)s1;
      atomicAdd(&sh[s1], 1u);
    }
    outlier_append(o, base + (unsigned long long)i, q1, ocount, oidx, oval, ocap);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < dict; i += blockDim.x) {
    unsigned c = sh[i];
    if (c)
      atomicAdd(&ghist[i], c);
  }
}

// s-norm variant: level = max_d level_marks[d][idx_d]; blocks own rows.
template <typename T>
__global__ void __launch_bounds__(256)
quantize_level_kernel(const QParams p, const Tables<T> tb, const T *__restrict__ v,
                      uint16_t *__restrict__ sym, unsigned *__restrict__ ghist,
                      unsigned long long *__restrict__ ocount,
                      uint64_t *__restrict__ oidx, i64 *__restrict__ oval,
                      unsigned long long ocap, const T *__restrict__ qdev) {
  extern __shared__ unsigned sh[];
  __shared__ T s_q[MGB_MAX_LEVELS];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = tid; i <= p.L; i += blockDim.x * blockDim.y)
    s_q[i] = qdev ? qdev[i] : tb.q[i];
  for (int i = tid; i < p.dict; i += blockDim.x * blockDim.y)
    sh[i] = 0;
  __syncthreads();
  const int D = p.D;
  const unsigned nf = p.n[D - 1];
  for (unsigned row0 = blockIdx.x * blockDim.y; row0 < p.rows;
       row0 += gridDim.x * blockDim.y) {
    unsigned row = row0 + threadIdx.y;
    bool rvalid = row < p.rows;
    int lvl = 0;
    unsigned rem = rvalid ? row : 0;
    for (int d = D - 2; d >= 0; d--) {
      unsigned i = rem % p.n[d];
      rem /= p.n[d];
      lvl = max(lvl, p.marks[(size_t)d * p.marks_width + i]);
    }
    i64 base = (i64)row * nf;
    unsigned nfr = (nf + blockDim.x - 1) / blockDim.x * blockDim.x;
    for (unsigned f = threadIdx.x; f < nfr; f += blockDim.x) {
      bool valid = rvalid && f < nf;
      unsigned s = 0;
      long long qi = 0;
      bool is_out = false;
      if (valid) {
        int l = max(lvl, p.marks[(size_t)(D - 1) * p.marks_width + f]);
        qi = quantize_one<T>(v[base + f], s_q[l], tb.vol[l], p.dict);
        if (qi >= 0 && qi < p.dict)
          s = (unsigned)qi;
        else
          is_out = true;
        sym[base + f] = (uint16_t)s;
      }
      outlier_append(is_out, (unsigned long long)(base + f), qi, ocount, oidx, oval, ocap);
      hist_add(sh, s, valid);
    }
  }
  __syncthreads();
  for (int i = tid; i < p.dict; i += blockDim.x * blockDim.y) {
    unsigned c = sh[i];
    if (c)
      atomicAdd(&ghist[i], c);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
dequantize_linear_kernel(const uint16_t *__restrict__ sym, i64 N, T qv, int dict,
                         T *__restrict__ v) {
  typedef typename Vec16<T>::type V;
  constexpr int VN = Vec16<T>::n, PER = 8, NV = PER / VN;
  const bool aligned = (((uintptr_t)v | (uintptr_t)sym) & 15) == 0;
  const i64 ngroups = aligned ? N / PER : 0;
  const i64 gstride = (i64)gridDim.x * blockDim.x;
  const int half = dict / 2;
  for (i64 gi = (i64)blockIdx.x * blockDim.x + threadIdx.x; gi < ngroups; gi += gstride) {
    const uint4 pk = __ldcs(reinterpret_cast<const uint4 *>(sym) + gi);
    const unsigned w[4] = {pk.x, pk.y, pk.z, pk.w};
    V outv[NV];
    T *x = reinterpret_cast<T *>(outv);
#pragma unroll
    for (int k = 0; k < PER; k++) {
      const long long qi = (long long)((w[k >> 1] >> (16 * (k & 1))) & 0xffffu) - half;
      x[k] = qv * (T)qi; // (quantizer * volume) * (T)quantized
    }
#pragma unroll
    for (int k = 0; k < NV; k++)
      reinterpret_cast<V *>(v + gi * PER)[k] = outv[k];
  }
  for (i64 i = ngroups * PER + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gstride) {
    long long qi = (long long)sym[i] - half;
    v[i] = qv * (T)qi;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
dequantize_level_kernel(const QParams p, const Tables<T> tb,
                        const uint16_t *__restrict__ sym, T *__restrict__ v) {
  const int D = p.D;
  const unsigned nf = p.n[D - 1];
  for (unsigned row0 = blockIdx.x * blockDim.y; row0 < p.rows;
       row0 += gridDim.x * blockDim.y) {
    unsigned row = row0 + threadIdx.y;
    if (row >= p.rows)
      continue;
    int lvl = 0;
    unsigned rem = row;
    for (int d = D - 2; d >= 0; d--) {
      unsigned i = rem % p.n[d];
      rem /= p.n[d];
      lvl = max(lvl, p.marks[(size_t)d * p.marks_width + i]);
    }
    i64 base = (i64)row * nf;
    for (unsigned f = threadIdx.x; f < nf; f += blockDim.x) {
      int l = max(lvl, p.marks[(size_t)(D - 1) * p.marks_width + f]);
      long long qi = (long long)sym[base + f] - p.dict / 2;
      v[base + f] = (tb.q[l] * tb.vol[l]) * (T)qi;
    }
  }
}

// outliers: v[idx] = (quantizer * volume) * (T)(outlier - dict/2)
template <typename T>
__global__ void outlier_restore_kernel(const QParams p, const Tables<T> tb,
                                       unsigned long long count,
                                       const uint64_t *__restrict__ oidx,
                                       const i64 *__restrict__ oval, T *__restrict__ v) {
  unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count)
    return;
  uint64_t idx = oidx[k];
  {
    // positions come from the stream: never write outside the array
    uint64_t total = 1;
    for (int d = 0; d < p.D; d++)
      total *= p.n[d];
    if (idx >= total)
      return;
  }
  int l = 0;
  if (p.calc_level) {
    uint64_t rem = idx;
    for (int d = p.D - 1; d >= 0; d--) {
      uint64_t i = rem % p.n[d];
      rem /= p.n[d];
      l = max(l, p.marks[(size_t)d * p.marks_width + i]);
    }
  }
  long long qi = oval[k] - p.dict / 2;
  v[idx] = (tb.q[l] * tb.vol[l]) * (T)qi;
}

// ------------------------------- norms -------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
norm_partial_kernel(const T *__restrict__ v, i64 N, double *__restrict__ part) {
  typedef typename Vec16<T>::type V;
  constexpr int VN = Vec16<T>::n;
  double mx = 0.0, ss = 0.0;
  // leading elements up to the first 16-byte boundary (sub-domain pointers are
  // only element aligned), then whole vectors, then the remainder
  i64 head = (i64)(((16 - ((uintptr_t)v & 15)) & 15) / sizeof(T));
  if (((uintptr_t)v % sizeof(T)) != 0 || head > N)
    head = N;
  if (blockIdx.x == 0 && threadIdx.x < head && head < 16) {
    double x = (double)v[threadIdx.x];
    mx = fmax(mx, fabs(x));
    ss += x * x;
  }
  if (head >= 16) { // misaligned element type: scalar everything
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < N; k += (i64)gridDim.x * blockDim.x) {
      double x = (double)v[k];
      mx = fmax(mx, fabs(x));
      ss += x * x;
    }
    N = 0;
    head = 0;
  }
  v += head;
  N -= head;
  const i64 nvec = N / VN;
  const i64 stride = (i64)gridDim.x * blockDim.x;
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  // two vectors in flight per thread
  for (; i + stride < nvec; i += 2 * stride) {
    V a = __ldcs(reinterpret_cast<const V *>(v) + i);
    V b = __ldcs(reinterpret_cast<const V *>(v) + i + stride);
    const T *xa = reinterpret_cast<const T *>(&a), *xb = reinterpret_cast<const T *>(&b);
#pragma unroll
    for (int k = 0; k < VN; k++) {
      double x = (double)xa[k];
      mx = fmax(mx, fabs(x));
      ss += x * x;
    }
#pragma unroll
    for (int k = 0; k < VN; k++) {
      double x = (double)xb[k];
      mx = fmax(mx, fabs(x));
      ss += x * x;
    }
  }
  for (; i < nvec; i += stride) {
    V a = reinterpret_cast<const V *>(v)[i];
    const T *xa = reinterpret_cast<const T *>(&a);
#pragma unroll
    for (int k = 0; k < VN; k++) {
      double x = (double)xa[k];
      mx = fmax(mx, fabs(x));
      ss += x * x;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < VN) {
    const i64 t = nvec * VN + threadIdx.x;
    if (t < N) {
      double x = (double)v[t];
      mx = fmax(mx, fabs(x));
      ss += x * x;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  __shared__ double smx[8], sss[8];
  int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    smx[w] = mx;
    sss[w] = ss;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; k++) {
      mx = fmax(mx, smx[k]);
      ss += sss[k];
    }
    part[2 * blockIdx.x] = mx;
    part[2 * blockIdx.x + 1] = ss;
  }
}

__global__ void norm_final_kernel(const double *__restrict__ part, int nblocks,
                                  double *__restrict__ out) {
  // single warp, fixed order => deterministic
  double mx = 0.0, ss = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 32) {
    mx = fmax(mx, part[2 * i]);
    ss += part[2 * i + 1];
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if (threadIdx.x == 0) {
    out[0] = mx;
    out[1] = ss;
  }
}

// LinearQuantizer::CalcQuantizers (LinearQuantization.hpp:495-545), evaluated
// with the reference's types: tol/s/norm are T, abs_tol is double.
// denominators of CalcQuantizers, evaluated with the reference's types (s is T, the
// products are double): quantizer[l] = abs_tol / denom[l]
template <typename T> void calc_denoms(const mgb_plan *p, double s_, double *denom) {
  const T s = (T)s_;
  const uint64_t l_target = (uint64_t)p->L;
  const size_t dof = p->N;
  for (int l = 0; l < p->L + 1; l++) {
    if (s == std::numeric_limits<T>::infinity()) {
      if (p->cfg.decomposition == 1) // SingleDim (LinearQuantization.hpp:516-520)
        denom[l] = ((l_target + 1) * (uint8_t)p->D * (1 + std::pow(3, 1)));
      else
        denom[l] = ((l_target + 1) * (1 + std::pow(3, p->D)));
    } else {
      denom[l] = (std::exp2(s * l) * std::sqrt(dof));
    }
  }
}

template <typename T>
void calc_quantizers(const mgb_plan *p, int ebtype, double tol_, double s_,
                     double norm_, bool reciprocal, T *out) {
  T tol = (T)tol_, norm = (T)norm_;
  double abs_tol = tol;
  if (ebtype == MGB_REL)
    abs_tol *= norm;
  abs_tol *= 2;
  double denom[MGB_MAX_LEVELS];
  calc_denoms<T>(p, s_, denom);
  for (int l = 0; l < p->L + 1; l++) {
    out[l] = (abs_tol) / denom[l];
    if (reciprocal)
      out[l] = 1.0f / out[l];
  }
}

// The same bookkeeping on the device, for bounds that depend on a norm which is only
// known in device memory (no host round trip between the norm and the quantizer):
//   norm      norm_calculator's last step (NormCalculator.hpp:44-71) or, decomposed,
//             calc_norm_decomposed (ErrorToleranceCalculator.hpp:91-132), in T, 0 -> epsilon
//   local tol calc_local_abs_tol (ErrorToleranceCalculator.hpp:134-155) when decomposed
//   table     CalcQuantizers (LinearQuantization.hpp:495-545), reciprocals
// IEEE division and square root are correctly rounded on both sides and nothing is
// contracted (-fmad=false), so the table equals the host's bit for bit.
struct QPrep {
  int nlevels;       // L + 1
  int rel;           // error_bound_type::REL
  int s_inf;
  int src;           // 0: fp32 max |x| bit pattern; 1: double {max |x|, sum x^2}
  unsigned long long n_total; // elements of the whole domain (s-norm)
  unsigned long long nsub;    // > 0: domain-decomposed into nsub sub-domains
  double tol;
  double denom[MGB_MAX_LEVELS];
};
template <typename T>
__global__ void prepare_q_kernel(const QPrep p, const unsigned *__restrict__ absmax_bits,
                                 const double *__restrict__ red, T *__restrict__ qtab,
                                 double *__restrict__ norm_out) {
  if (threadIdx.x || blockIdx.x)
    return;
  T norm = (T)1;
  if (p.rel) {
    if (p.src == 0) {
      norm = (T)__uint_as_float(*absmax_bits);
    } else if (p.s_inf) {
      norm = (T)red[0];
    } else if (sizeof(T) == 4) {
      norm = (T)sqrtf((float)red[1] / (float)p.n_total);
    } else {
      norm = (T)sqrt(red[1] / (double)p.n_total);
    }
    if (norm == (T)0)
      norm = std::numeric_limits<T>::epsilon();
  }
  *norm_out = (double)norm;
  T tol = (T)p.tol;
  int rel = p.rel;
  if (p.nsub) {
    // sub-domains are compressed with an absolute bound
    if (sizeof(T) == 4) {
      const float t = (float)p.tol, n = (float)norm;
      float lt;
      if (p.rel)
        lt = p.s_inf ? t * n : sqrtf((t * n) * (t * n) / (float)p.nsub);
      else
        lt = p.s_inf ? t : sqrtf((t * t) / (float)p.nsub);
      tol = (T)lt;
    } else {
      const double t = p.tol, n = (double)norm;
      double lt;
      if (p.rel)
        lt = p.s_inf ? t * n : sqrt((t * n) * (t * n) / (double)p.nsub);
      else
        lt = p.s_inf ? t : sqrt((t * t) / (double)p.nsub);
      tol = (T)lt;
    }
    rel = 0;
  }
  double abs_tol = (double)tol;
  if (rel)
    abs_tol *= (double)norm;
  abs_tol *= 2;
  for (int l = 0; l < p.nlevels; l++) {
    T q = (T)(abs_tol / p.denom[l]);
    qtab[l] = (T)(1.0f / q);
  }
}

// level volume factor: Hierarchy::calc_volume (Hierarchy.hpp:165-190) and the
// product / sqrt of LinearQuantization.hpp:190-198
template <typename T>
void calc_volumes(const mgb_plan *p, bool reciprocal, bool calc_vol, T *out) {
  for (int l = 0; l <= p->L; l++) {
    T volume = 1;
    if (calc_vol) {
      for (int d = p->D - 1; d >= 0; d--) {
        uint64_t dofd = p->lshape[l][d];
        T hv = 0.0;
        if (dofd > 1)
          hv = 1.0 / (T)(dofd - 1);
        if (reciprocal)
          hv = 1.0 / hv;
        volume *= hv;
      }
      if (sizeof(T) == sizeof(double))
        volume = std::sqrt(volume);
      else
        volume = sqrtf((float)volume);
    }
    out[l] = volume;
  }
}

template <typename T>
void make_params(const mgb_plan *p, int ebtype, double tol, double s, double norm,
                 bool dequant, QParams &qp, Tables<T> &tb) {
  memset(&qp, 0, sizeof(qp));
  qp.D = p->D;
  qp.dict = p->cfg.huff_dict_size;
  qp.L = p->L;
  bool calc_vol = !(std::isinf(s) && s > 0);
  qp.calc_level = calc_vol;
  unsigned rows = 1;
  for (int d = 0; d < p->D; d++) {
    qp.n[d] = (unsigned)p->shape[d];
    if (d < p->D - 1)
      rows *= (unsigned)p->shape[d];
  }
  qp.rows = rows;
  qp.marks = p->d_marks;
  qp.marks_width = p->marks_width;
  calc_quantizers<T>(p, ebtype, tol, s, norm, !dequant, tb.q);
  calc_volumes<T>(p, dequant, calc_vol, tb.vol);
}

template <typename T>
int quantize_t(mgb_plan *p, const T *d_coef, int ebtype, double tol, double s,
               double norm, uint16_t *d_sym, uint32_t *d_hist,
               unsigned long long *d_ocount, uint64_t *d_oidx, int64_t *d_oval,
               uint64_t ocap, cudaStream_t st, uint64_t first = 0, uint64_t count = ~0ull,
               bool zero = true, unsigned max_blocks = 148 * 4, const T *qdev = nullptr) {
  // [first, first + count): part of the array (s = inf only); zero: clear the
  // histogram and the outlier counter first
  QParams qp;
  Tables<T> tb;
  make_params<T>(p, ebtype, tol, s, norm, false, qp, tb);
  const int dict = qp.dict;
  if (count == ~0ull)
    count = p->N - first;
  if (zero) {
    MGB_CUDA_CHECK(cudaMemsetAsync(d_hist, 0, dict * sizeof(uint32_t), st));
    MGB_CUDA_CHECK(cudaMemsetAsync(d_ocount, 0, sizeof(unsigned long long), st));
  }
  size_t smem = dict * sizeof(unsigned);
  if (!qp.calc_level) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(quantize_linear_kernel<T>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (count == 0)
      return MGB_SUCCESS;
    unsigned blocks = (unsigned)std::min<i64>(((i64)count + 2047) / 2048, (i64)max_blocks);
    MGB_LAUNCH(MGB_K_QUANTIZE, st,
               (quantize_linear_kernel<T><<<blocks, 256, smem, st>>>(
                   d_coef + first, (i64)count, tb.q[0], tb.vol[0], dict, d_sym + first, d_hist,
                   d_ocount, d_oidx, (i64 *)d_oval, ocap, (unsigned long long)first, qdev)));
  } else {
    if (first != 0 || count != p->N)
      return MGB_BAD_ARGUMENT;
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(quantize_level_kernel<T>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int bx = 32;
    while (bx < (int)qp.n[p->D - 1] && bx < 256)
      bx <<= 1;
    dim3 block(bx, 256 / bx);
    unsigned blocks = std::min<unsigned>((qp.rows + block.y - 1) / block.y, 148 * 6);
    MGB_LAUNCH(MGB_K_QUANTIZE, st,
               (quantize_level_kernel<T><<<blocks, block, smem, st>>>(
                   qp, tb, d_coef, d_sym, d_hist, d_ocount, d_oidx, (i64 *)d_oval, ocap, qdev)));
  }
  MGB_CUDA_CHECK(cudaGetLastError());
  return MGB_SUCCESS;
}

template <typename T>
int dequantize_t(mgb_plan *p, const uint16_t *d_sym, uint64_t ocount,
                 const uint64_t *d_oidx, const int64_t *d_oval, int ebtype,
                 double tol, double s, double norm, T *d_coef, cudaStream_t st,
                 bool outliers_only = false) {
  QParams qp;
  Tables<T> tb;
  make_params<T>(p, ebtype, tol, s, norm, true, qp, tb);
  if (outliers_only) {
    // the decoder already wrote (quantizer * volume) * (T)quantized
  } else if (!qp.calc_level) {
    unsigned blocks = (unsigned)std::min<i64>((p->N + 2047) / 2048, 148 * 8);
    MGB_LAUNCH(MGB_K_DEQUANTIZE, st,
               (dequantize_linear_kernel<T><<<blocks, 256, 0, st>>>(
                   d_sym, (i64)p->N, tb.q[0] * tb.vol[0], qp.dict, d_coef)));
  } else {
    int bx = 32;
    while (bx < (int)qp.n[p->D - 1] && bx < 256)
      bx <<= 1;
    dim3 block(bx, 256 / bx);
    unsigned blocks = std::min<unsigned>((qp.rows + block.y - 1) / block.y, 148 * 16);
    MGB_LAUNCH(MGB_K_DEQUANTIZE, st,
               (dequantize_level_kernel<T><<<blocks, block, 0, st>>>(qp, tb, d_sym, d_coef)));
  }
  if (ocount) {
    MGB_LAUNCH(MGB_K_OUTLIER_RESTORE, st,
               (outlier_restore_kernel<T><<<(unsigned)((ocount + 255) / 256), 256, 0, st>>>(
                   qp, tb, ocount, d_oidx, (const i64 *)d_oval, d_coef)));
  }
  MGB_CUDA_CHECK(cudaGetLastError());
  return MGB_SUCCESS;
}

} // namespace

extern "C" int mgb_quantize(mgb_plan *plan, const void *d_coef, int ebtype,
                            double tol, double s, double norm, uint16_t *d_sym,
                            uint32_t *d_hist, unsigned long long *d_ocount,
                            uint64_t *d_oidx, int64_t *d_oval,
                            uint64_t outlier_cap, void *stream) {
  if (!plan || !d_coef || !d_sym || !d_hist || !d_ocount)
    return MGB_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  if (plan->dtype == MGB_F32)
    return quantize_t<float>(plan, (const float *)d_coef, ebtype, tol, s, norm,
                             d_sym, d_hist, d_ocount, d_oidx, d_oval,
                             outlier_cap, st);
  return quantize_t<double>(plan, (const double *)d_coef, ebtype, tol, s, norm,
                            d_sym, d_hist, d_ocount, d_oidx, d_oval,
                            outlier_cap, st);
}

extern "C" int mgb_dequantize(mgb_plan *plan, const uint16_t *d_sym,
                              uint64_t ocount, const uint64_t *d_oidx,
                              const int64_t *d_oval, int ebtype, double tol,
                              double s, double norm, void *d_coef, void *stream) {
  if (!plan || !d_coef || !d_sym)
    return MGB_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  if (plan->dtype == MGB_F32)
    return dequantize_t<float>(plan, d_sym, ocount, d_oidx, d_oval, ebtype, tol,
                               s, norm, (float *)d_coef, st);
  return dequantize_t<double>(plan, d_sym, ocount, d_oidx, d_oval, ebtype, tol,
                              s, norm, (double *)d_coef, st);
}

extern "C" int mgb_norm_partials(mgb_plan *plan, const void *d_in,
                                 double *absmax, double *sumsq) {
  if (!plan || !d_in)
    return MGB_BAD_ARGUMENT;
  const int nblocks = 148 * 8;
  if (!plan->d_norm_tmp)
    MGB_CUDA_CHECK(cudaMalloc(&plan->d_norm_tmp, (2 * nblocks + 2) * sizeof(double)));
  double *part = (double *)plan->d_norm_tmp;
  if (plan->dtype == MGB_F32)
    MGB_LAUNCH(MGB_K_NORM, 0,
               (norm_partial_kernel<float><<<nblocks, 256>>>((const float *)d_in, (i64)plan->N, part)));
  else
    MGB_LAUNCH(MGB_K_NORM, 0,
               (norm_partial_kernel<double><<<nblocks, 256>>>((const double *)d_in, (i64)plan->N, part)));
  MGB_LAUNCH(MGB_K_NORM, 0, (norm_final_kernel<<<1, 32>>>(part, nblocks, part + 2 * nblocks)));
  double h[2];
  MGB_CUDA_CHECK(cudaMemcpy(h, part + 2 * nblocks, sizeof(h), cudaMemcpyDeviceToHost));
  if (absmax)
    *absmax = h[0];
  if (sumsq)
    *sumsq = h[1];
  return MGB_SUCCESS;
}

extern "C" int mgb_norm(mgb_plan *plan, const void *d_in, double s, double *norm) {
  if (!norm)
    return MGB_BAD_ARGUMENT;
  double mx, ss;
  int rc = mgb_norm_partials(plan, d_in, &mx, &ss);
  if (rc)
    return rc;
  // NormCalculator.hpp:44-71 (normalize_coordinates = true), result in T
  double r;
  if (std::isinf(s) && s > 0) {
    r = mx;
  } else {
    if (plan->dtype == MGB_F32)
      r = std::sqrt((float)ss / plan->N);
    else
      r = std::sqrt(ss / plan->N);
  }
  if (plan->dtype == MGB_F32) {
    float f = (float)r;
    if (f == 0)
      f = std::numeric_limits<float>::epsilon();
    r = f;
  } else if (r == 0) {
    r = std::numeric_limits<double>::epsilon();
  }
  *norm = r;
  return MGB_SUCCESS;
}

// s = inf: the dequantizer is one scalar for every node, so the Huffman decoder
// can apply it while it flushes a chunk (no symbol array, no dequantize pass).
// Returns 1 and the factor (quantizer * volume, evaluated in T) if so.
int mgb_linear_dequant_scale(mgb_plan *plan, int ebtype, double tol, double s, double norm,
                             double *scale) {
  if (!(std::isinf(s) && s > 0))
    return 0;
  if (plan->dtype == MGB_F32) {
    QParams qp;
    Tables<float> tb;
    make_params<float>(plan, ebtype, tol, s, norm, true, qp, tb);
    *scale = (double)(tb.q[0] * tb.vol[0]);
  } else {
    QParams qp;
    Tables<double> tb;
    make_params<double>(plan, ebtype, tol, s, norm, true, qp, tb);
    *scale = tb.q[0] * tb.vol[0];
  }
  return 1;
}

int mgb_outlier_restore(mgb_plan *plan, uint64_t ocount, const uint64_t *d_oidx,
                        const int64_t *d_oval, int ebtype, double tol, double s, double norm,
                        void *d_coef, cudaStream_t st) {
  if (!ocount)
    return MGB_SUCCESS;
  if (plan->dtype == MGB_F32)
    return dequantize_t<float>(plan, nullptr, ocount, d_oidx, d_oval, ebtype, tol, s, norm,
                               (float *)d_coef, st, true);
  return dequantize_t<double>(plan, nullptr, ocount, d_oidx, d_oval, ebtype, tol, s, norm,
                              (double *)d_coef, st, true);
}

// Part of the array only (s = inf): lets the compressor quantize the level-l_target
// coefficients that are final early, next to the rest of the decomposition.
int mgb_quantize_range(mgb_plan *plan, const void *d_coef, int ebtype, double tol, double s,
                       double norm, uint16_t *d_sym, uint32_t *d_hist,
                       unsigned long long *d_ocount, uint64_t *d_oidx, int64_t *d_oval,
                       uint64_t outlier_cap, uint64_t first, uint64_t count, int zero,
                       unsigned max_blocks, cudaStream_t st, const void *d_qtab) {
  // d_qtab != nullptr: reciprocal quantizers in device memory (mgb_prepare_quantizers);
  // ebtype / tol / norm are not used then
  if (count == ~0ull)
    count = plan->N - first;
  if (plan->dtype == MGB_F32)
    return quantize_t<float>(plan, (const float *)d_coef, ebtype, tol, s, norm, d_sym, d_hist,
                             d_ocount, d_oidx, d_oval, outlier_cap, st, first, count, zero != 0,
                             max_blocks, (const float *)d_qtab);
  return quantize_t<double>(plan, (const double *)d_coef, ebtype, tol, s, norm, d_sym, d_hist,
                            d_ocount, d_oidx, d_oval, outlier_cap, st, first, count, zero != 0,
                            max_blocks, (const double *)d_qtab);
}

// Quantizer table of `plan` from a norm that lives in device memory (see
// prepare_q_kernel).  src 0: d_src = fp32 bit pattern of max |x| (coef3d by-product);
// src 1: d_src = double {max |x|, sum x^2} (norm kernels, possibly all-reduced).
// n_total: elements the s-norm is taken over; nsub > 0: domain-decomposed run.
int mgb_prepare_quantizers(mgb_plan *plan, int ebtype, double tol, double s, int src, const void *d_src,
                           uint64_t n_total, uint64_t nsub, void *d_qtab, double *d_norm_out,
                           cudaStream_t st) {
  QPrep q;
  memset(&q, 0, sizeof(q));
  q.nlevels = plan->L + 1;
  q.rel = ebtype == MGB_REL;
  q.s_inf = std::isinf(s) && s > 0;
  q.src = src;
  q.n_total = n_total;
  q.nsub = nsub;
  q.tol = tol;
  if (plan->dtype == MGB_F32) {
    calc_denoms<float>(plan, s, q.denom);
    MGB_LAUNCH(MGB_K_NORM, st,
               (prepare_q_kernel<float><<<1, 32, 0, st>>>(q, (const unsigned *)d_src, (const double *)d_src,
                                                          (float *)d_qtab, d_norm_out)));
  } else {
    calc_denoms<double>(plan, s, q.denom);
    MGB_LAUNCH(MGB_K_NORM, st,
               (prepare_q_kernel<double><<<1, 32, 0, st>>>(q, (const unsigned *)d_src, (const double *)d_src,
                                                           (double *)d_qtab, d_norm_out)));
  }
  MGB_CUDA_CHECK(cudaGetLastError());
  return MGB_SUCCESS;
}

// max |x| and sum x^2 of n elements into d_red[0..1] (doubles), asynchronously:
// two-stage deterministic reduction (norm_partial_kernel / norm_final_kernel)
int mgb_norm_raw(int dtype, const void *d_in, uint64_t n, double *d_part, double *d_red, cudaStream_t st) {
  // d_part: MGB_NORM_PART_DOUBLES doubles of scratch
  const int nblocks = 148 * 8;
  if (dtype == MGB_F32)
    MGB_LAUNCH(MGB_K_NORM, st,
               (norm_partial_kernel<float><<<nblocks, 256, 0, st>>>((const float *)d_in, (i64)n, d_part)));
  else
    MGB_LAUNCH(MGB_K_NORM, st,
               (norm_partial_kernel<double><<<nblocks, 256, 0, st>>>((const double *)d_in, (i64)n, d_part)));
  MGB_LAUNCH(MGB_K_NORM, st, (norm_final_kernel<<<1, 32, 0, st>>>(d_part, nblocks, d_red)));
  MGB_CUDA_CHECK(cudaGetLastError());
  return MGB_SUCCESS;
}
// {max, sum} pairs of several sub-domains -> one pair (fixed order: deterministic)
int mgb_norm_combine(const double *d_pairs, int count, double *d_red, cudaStream_t st) {
  MGB_LAUNCH(MGB_K_NORM, st, (norm_final_kernel<<<1, 32, 0, st>>>(d_pairs, count, d_red)));
  MGB_CUDA_CHECK(cudaGetLastError());
  return MGB_SUCCESS;
}
int mgb_norm_async(mgb_plan *plan, const void *d_in, uint64_t n, double *d_red, cudaStream_t st) {
  if (!plan->d_norm_tmp)
    MGB_CUDA_CHECK(cudaMalloc(&plan->d_norm_tmp, MGB_NORM_PART_DOUBLES * sizeof(double)));
  return mgb_norm_raw(plan->dtype, d_in, n, (double *)plan->d_norm_tmp, d_red, st);
}

// Outliers are appended with atomics, so their order depends on scheduling.
// Sorting them by index (what the reference's SERIAL adapter produces) makes the
// stream deterministic and byte-identical to the reference's.  Bitonic sort of the
// (index, value) pairs, in place: ST_BLOCKS blocks, each owning tiles of ST_TILE pairs.
// Compare-exchange distances below a tile run out of shared memory; the larger ones go
// through global memory with a grid barrier in between (the blocks are few enough to be
// co-resident; a block that has to wait for an SM only delays the others).  Lists longer
// than 65536 entries (or that do not fit their power-of-two padding) are left as they are.
namespace {
constexpr unsigned ST_TILE = 4096, ST_BLOCKS = 16, ST_MAX = ST_TILE * ST_BLOCKS;

// bar[0] arrivals, bar[1] generation (zero-initialised words of the plan)
__device__ __forceinline__ void sort_grid_barrier(unsigned *bar, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile unsigned *gen = &bar[1];
    const unsigned my = *gen;
    __threadfence();
    if (atomicAdd(&bar[0], 1u) == nblocks - 1) {
      bar[0] = 0;
      __threadfence();
      atomicAdd(&bar[1], 1u);
    } else {
      while (*gen == my)
        __nanosleep(64);
    }
    __threadfence();
  }
  __syncthreads();
}

// distances j = jmax .. 1 of stage k on the tile in shared memory
__device__ __forceinline__ void sort_tile_passes(uint64_t *sk, long long *sv, unsigned base, unsigned k,
                                                 unsigned jmax) {
  for (unsigned j = jmax; j > 0; j >>= 1) {
    for (unsigned t = threadIdx.x; t < ST_TILE / 2; t += blockDim.x) {
      // t-th pair (lo, lo ^ j) of the tile
      const unsigned lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;
      const uint64_t a = sk[lo], b = sk[hi];
      const bool up = ((base + lo) & k) == 0;
      if ((a > b) == up) {
        sk[lo] = b;
        sk[hi] = a;
        const long long va = sv[lo];
        sv[lo] = sv[hi];
        sv[hi] = va;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024)
sort_outliers_kernel(const unsigned long long *__restrict__ ocount, uint64_t *oidx,
                     long long *oval, unsigned long long cap, unsigned *bar) {
  extern __shared__ __align__(16) unsigned char sort_smem[];
  uint64_t *sk = reinterpret_cast<uint64_t *>(sort_smem);
  long long *sv = reinterpret_cast<long long *>(sort_smem) + ST_TILE;
  const unsigned long long n64 = *ocount;
  if (n64 < 2 || n64 > ST_MAX || n64 > cap)
    return;
  const unsigned n = (unsigned)n64;
  unsigned np = 1;
  while (np < n)
    np <<= 1;
  if (np > cap)
    return;
  // every decision below depends on n only, so all blocks take the same barriers
  const unsigned ntiles = (np + ST_TILE - 1) / ST_TILE; // power of two (or 1)
  const unsigned tile = blockIdx.x;
  const bool active = tile < ntiles;
  const unsigned base = tile * ST_TILE;
  const unsigned tlen = min(ST_TILE, np); // np < ST_TILE: one short tile
  // stages up to the tile size, entirely in shared memory
  if (active) {
    for (unsigned i = threadIdx.x; i < ST_TILE; i += blockDim.x) {
      const unsigned g = base + i;
      const bool real = i < tlen && g < n;
      sk[i] = real ? __ldcg(oidx + g) : ~0ull; // padding sorts to the end
      sv[i] = real ? __ldcg(oval + g) : 0;
    }
    __syncthreads();
    for (unsigned k = 2; k <= tlen; k <<= 1)
      sort_tile_passes(sk, sv, base, k, k >> 1);
    if (ntiles > 1)
      for (unsigned i = threadIdx.x; i < ST_TILE; i += blockDim.x) {
        __stcg(oidx + base + i, sk[i]);
        __stcg(oval + base + i, sv[i]);
      }
  }
  // larger stages: distances >= ST_TILE through global memory, the rest per tile
  for (unsigned k = 2 * ST_TILE; k <= np && ntiles > 1; k <<= 1) {
    for (unsigned j = k >> 1; j >= ST_TILE; j >>= 1) {
      __threadfence();
      sort_grid_barrier(bar, gridDim.x);
      // np / 2 pairs over all threads of the grid
      for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < np / 2; t += gridDim.x * blockDim.x) {
        const unsigned lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;
        // other blocks wrote these in the previous pass: read them from the L2
        const uint64_t a = __ldcg(oidx + lo), b = __ldcg(oidx + hi);
        const bool up = (lo & k) == 0;
        if ((a > b) == up) {
          __stcg(oidx + lo, b);
          __stcg(oidx + hi, a);
          const long long va = __ldcg(oval + lo), vb = __ldcg(oval + hi);
          __stcg(oval + lo, vb);
          __stcg(oval + hi, va);
        }
      }
    }
    __threadfence();
    sort_grid_barrier(bar, gridDim.x);
    if (active) {
      for (unsigned i = threadIdx.x; i < ST_TILE; i += blockDim.x) {
        sk[i] = __ldcg(oidx + base + i);
        sv[i] = __ldcg(oval + base + i);
      }
      __syncthreads();
      sort_tile_passes(sk, sv, base, k, ST_TILE >> 1);
      for (unsigned i = threadIdx.x; i < ST_TILE; i += blockDim.x) {
        __stcg(oidx + base + i, sk[i]);
        __stcg(oval + base + i, sv[i]);
      }
    }
  }
  if (ntiles == 1 && active)
    for (unsigned i = threadIdx.x; i < tlen; i += blockDim.x) {
      __stcg(oidx + i, sk[i]);
      __stcg(oval + i, sv[i]);
    }
}
} // namespace

// d_ocount: the plan's scalar block (d_scalars: [0] the count, [16..17] the barrier words)
int mgb_sort_outliers(const unsigned long long *d_ocount, uint64_t *d_oidx, int64_t *d_oval,
                      uint64_t cap, cudaStream_t st) {
  static bool configured[64] = {};
  if (mgb_first_use_on_device(configured))
    MGB_CUDA_CHECK(cudaFuncSetAttribute(sort_outliers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(ST_TILE * 16)));
  MGB_LAUNCH(MGB_K_OUTLIER_RESTORE, st,
             (sort_outliers_kernel<<<ST_BLOCKS, 1024, ST_TILE * 16, st>>>(
                 d_ocount, d_oidx, (long long *)d_oval, cap, (unsigned *)(d_ocount + 16))));
  MGB_CUDA_CHECK(cudaGetLastError());
  return MGB_SUCCESS;
}


// ---- Config::reorder = 1: level-linearised order of the quantised symbols ------
// calc_level_offset (LinearQuantization.hpp:46-146) behind the previous levels
// (:591-604): within a level, the nodes it introduces in the row-major order of
// that level's un-reordered mesh.
namespace {

struct LinearizeGeom {
  int D, L;
  uint64_t shape[MGB_MAX_DIMS];
  uint32_t ranges[MGB_MAX_LEVELS + 2][MGB_MAX_DIMS]; // level_ranges (Hierarchy.hpp:243-260)
  uint64_t base[MGB_MAX_LEVELS + 1];                  // elements of the levels below
  const int *marks;
  uint64_t marks_width;
};

__device__ __forceinline__ uint64_t level_linear_pos(const LinearizeGeom &g, uint64_t i) {
  uint32_t idx[MGB_MAX_DIMS];
  int mark[MGB_MAX_DIMS];
  int level = 0;
  for (int d = g.D - 1; d >= 0; d--) {
    idx[d] = (uint32_t)(i % g.shape[d]);
    i /= g.shape[d];
    mark[d] = g.marks[(uint64_t)d * g.marks_width + idx[d]];
    level = max(level, mark[d]);
  }
  uint64_t thread_off = 0, coarse_off = 0, stride = 1, cstride = 1;
  for (int d = g.D - 1; d >= 0; d--) {
    const uint32_t bit = mark[d] == level;
    const uint32_t n = g.ranges[level + 1][d];
    const uint32_t t = bit ? idx[d] - g.ranges[level][d] : idx[d];
    uint32_t gi;
    if (level == 0)
      gi = t;
    else if (n % 2 == 0 && t == n / 2)
      gi = n - 1;
    else
      gi = t * 2 + bit;
    thread_off += (uint64_t)gi * stride;
    stride *= n;
    if (gi % 2 != 0 && gi != n - 1)
      coarse_off = 0;
    if (gi)
      coarse_off += (uint64_t)((gi - 1) / 2 + 1) * cstride;
    cstride *= n / 2 + 1;
  }
  if (level == 0)
    coarse_off = 0;
  return g.base[level] + thread_off - coarse_off;
}

__global__ void __launch_bounds__(256) linearize_kernel(const __grid_constant__ LinearizeGeom g, uint64_t n,
                                                        const uint16_t *__restrict__ dense,
                                                        uint16_t *__restrict__ linear) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    linear[level_linear_pos(g, i)] = dense[i];
}

__global__ void __launch_bounds__(256) delinearize_kernel(const __grid_constant__ LinearizeGeom g, uint64_t n,
                                                          const uint16_t *__restrict__ linear,
                                                          uint16_t *__restrict__ dense,
                                                          uint32_t *__restrict__ inverse) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const uint64_t pos = level_linear_pos(g, i);
    dense[i] = linear[pos];
    inverse[pos] = (uint32_t)i;
  }
}

__global__ void __launch_bounds__(256) linearize_outliers_kernel(const __grid_constant__ LinearizeGeom g,
                                                                 const unsigned long long *d_count, uint64_t cap,
                                                                 uint64_t *oidx) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t count = min((uint64_t)*d_count, cap);
  if (k < count)
    oidx[k] = level_linear_pos(g, oidx[k]);
}

__global__ void __launch_bounds__(256) delinearize_outliers_kernel(uint64_t count, uint64_t n,
                                                                   const uint32_t *__restrict__ inverse,
                                                                   const uint64_t *__restrict__ oidx_linear,
                                                                   uint64_t *__restrict__ oidx_dense) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < count) {
    const uint64_t pos = oidx_linear[k];
    oidx_dense[k] = pos < n ? inverse[pos] : n; // out-of-range entries are dropped by the restore kernel
  }
}

LinearizeGeom make_geom(const mgb_plan *p) {
  LinearizeGeom g;
  memset(&g, 0, sizeof(g));
  g.D = p->D;
  g.L = p->L;
  for (int d = 0; d < p->D; d++)
    g.shape[d] = p->shape[d];
  for (int l = 0; l <= p->L; l++) {
    uint64_t below = 1;
    for (int d = 0; d < p->D; d++) {
      g.ranges[l + 1][d] = (uint32_t)p->lshape[l][d];
      below *= g.ranges[l][d];
    }
    g.base[l] = l == 0 ? 0 : below;
  }
  g.marks = p->d_marks;
  g.marks_width = p->marks_width;
  return g;
}

} // namespace

int mgb_linearize_symbols(mgb_plan *p, const uint16_t *d_dense, uint16_t *d_linear,
                          const unsigned long long *d_ocount, uint64_t *d_oidx, uint64_t ocap, cudaStream_t st) {
  if (p->w_elems * p->tsize < p->N * sizeof(uint16_t))
    return MGB_FAILURE;
  const LinearizeGeom g = make_geom(p);
  const unsigned blocks = (unsigned)((p->N + 255) / 256);
  MGB_LAUNCH(MGB_K_QUANTIZE, st, (linearize_kernel<<<blocks, 256, 0, st>>>(g, p->N, d_dense, d_linear)));
  const unsigned oblocks = (unsigned)std::min<uint64_t>((ocap + 255) / 256, 1u << 20);
  MGB_LAUNCH(MGB_K_QUANTIZE, st, (linearize_outliers_kernel<<<oblocks, 256, 0, st>>>(g, d_ocount, ocap, d_oidx)));
  MGB_CUDA_CHECK(cudaGetLastError());
  return MGB_SUCCESS;
}

int mgb_delinearize_symbols(mgb_plan *p, const uint16_t *d_linear, uint16_t *d_dense, uint32_t *d_inverse,
                            uint64_t ocount, const uint64_t *d_oidx_linear, uint64_t *d_oidx_dense,
                            cudaStream_t st) {
  if (p->w_elems * p->tsize < p->N * sizeof(uint16_t))
    return MGB_FAILURE;
  const LinearizeGeom g = make_geom(p);
  const unsigned blocks = (unsigned)((p->N + 255) / 256);
  MGB_LAUNCH(MGB_K_DEQUANTIZE, st,
             (delinearize_kernel<<<blocks, 256, 0, st>>>(g, p->N, d_linear, d_dense, d_inverse)));
  if (ocount)
    MGB_LAUNCH(MGB_K_DEQUANTIZE, st,
               (delinearize_outliers_kernel<<<(unsigned)((ocount + 255) / 256), 256, 0, st>>>(
                   ocount, p->N, d_inverse, d_oidx_linear, d_oidx_dense)));
  MGB_CUDA_CHECK(cudaGetLastError());
  return MGB_SUCCESS;
}
