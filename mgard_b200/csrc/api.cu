// C ABI: low-level (one sub-domain) and high-level (self-describing stream)
// compress / decompress.
//
//   low level   reference Compressor<D,T>::Compress / Decompress
//               (include/mgard-x/CompressionLowLevel/Compressor.hpp:193-272)
//   high level  reference general_compress / general_decompress
//               (include/mgard-x/CompressionHighLevel/CompressionHighLevel.hpp:49-314,379-594),
//               compress_pipeline_gpu / decompress_pipeline_gpu
//               (CompressionHighLevel/GPUPipelines.hpp:3-207,270-520),
//               DomainDecomposer MaxDim partition (DomainDecomposer/DomainDecomposer.hpp:124-169),
//               calc_local_abs_tol (CompressionHighLevel/ErrorToleranceCalculator.hpp:134-155)
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <limits>
#include <map>
#include <mutex>
#include <vector>

#include "format.h"
#include "plan.h"

// huffman.cu internals
int mgb_huff_workspace(mgb_plan *p);
int mgb_huffman_compress_async(mgb_plan *p, const uint16_t *d_sym, uint64_t n,
                               const uint32_t *d_hist,
                               const unsigned long long *d_ocount_ptr,
                               uint64_t ocount_fixed, const uint64_t *d_oidx,
                               const int64_t *d_oval, uint8_t *d_out, uint64_t cap,
                               cudaStream_t st);
int mgb_huffman_finish(mgb_plan *p, uint64_t *size, cudaStream_t st);
int mgb_huffman_decompress_impl(mgb_plan *p, const uint8_t *d_in, uint64_t size, uint16_t *d_sym,
                                uint64_t n, uint64_t *ocount, const uint64_t **d_oidx,
                                const int64_t **d_oval, void *stream, void *d_deq,
                                double deq_scale, int *fused);
// quantize.cu internals
int mgb_quantize_range(mgb_plan *plan, const void *d_coef, int ebtype, double tol, double s,
                       double norm, uint16_t *d_sym, uint32_t *d_hist,
                       unsigned long long *d_ocount, uint64_t *d_oidx, int64_t *d_oval,
                       uint64_t outlier_cap, uint64_t first, uint64_t count, int zero,
                       unsigned max_blocks, cudaStream_t st, const void *d_qtab);
int mgb_prepare_quantizers(mgb_plan *plan, int ebtype, double tol, double s, int src, const void *d_src,
                           uint64_t n_total, uint64_t nsub, void *d_qtab, double *d_norm_out,
                           cudaStream_t st);
int mgb_norm_async(mgb_plan *plan, const void *d_in, uint64_t n, double *d_red, cudaStream_t st);
int mgb_linearize_symbols(mgb_plan *p, const uint16_t *d_dense, uint16_t *d_linear, const unsigned long long *d_ocount,
                          uint64_t *d_oidx, uint64_t ocap, cudaStream_t st);
int mgb_delinearize_symbols(mgb_plan *p, const uint16_t *d_linear, uint16_t *d_dense, uint32_t *d_inverse,
                            uint64_t ocount, const uint64_t *d_oidx_linear, uint64_t *d_oidx_dense,
                            cudaStream_t st);
int mgb_sort_outliers(const unsigned long long *d_ocount, uint64_t *d_oidx, int64_t *d_oval,
                      uint64_t cap, cudaStream_t st);
int mgb_linear_dequant_scale(mgb_plan *plan, int ebtype, double tol, double s, double norm,
                             double *scale);
int mgb_outlier_restore(mgb_plan *plan, uint64_t ocount, const uint64_t *d_oidx,
                        const int64_t *d_oval, int ebtype, double tol, double s, double norm,
                        void *d_coef, cudaStream_t st);

namespace {

int ensure_lowlevel_workspace(mgb_plan *p) {
  int rc = mgb_plan_ensure_workspace(p);
  if (rc)
    return rc;
  rc = mgb_huff_workspace(p);
  if (rc)
    return rc;
  if (!p->d_coef)
    MGB_CUDA_CHECK(cudaMalloc(&p->d_coef, p->N * p->tsize));
  if (!p->d_sym)
    MGB_CUDA_CHECK(cudaMalloc(&p->d_sym, p->N * sizeof(uint16_t) + 64));
  if (!p->d_hist)
    MGB_CUDA_CHECK(cudaMalloc(&p->d_hist, p->cfg.huff_dict_size * sizeof(uint32_t)));
  if (!p->d_qtab)
    MGB_CUDA_CHECK(cudaMalloc(&p->d_qtab, MGB_MAX_LEVELS * sizeof(double)));
  if (!p->d_oidx) {
    // a power of two, so that the list can always be padded for the index sort
    p->outlier_cap = 4096;
    while (p->outlier_cap < p->N / 32)
      p->outlier_cap <<= 1;
    MGB_CUDA_CHECK(cudaMalloc(&p->d_oidx, p->outlier_cap * 8));
    MGB_CUDA_CHECK(cudaMalloc(&p->d_oval, p->outlier_cap * 8));
  }
  return MGB_SUCCESS;
}

bool is_inf(double s) { return std::isinf(s) && s > 0; }

} // namespace

// Everything of Compressor::Compress up to (not including) the final read-back of
// the block size: norm (relative bounds), decomposition, quantization + histogram,
// Huffman.  Nothing here waits for the device.
//   d_qtab_ext != nullptr  reciprocal quantizers already in device memory (global norm
//                          of a domain-decomposed run, mgb_prepare_quantizers)
//   otherwise REL          the norm is reduced on the device - by-product of the finest
//                          level's coefficient kernel (L-inf, fp32, tiled 3-D path) or
//                          the norm kernels - and turned into the table there
//   otherwise ABS          the table is computed on the host and passed by value
static int compress_lowlevel_async(mgb_plan *p, const void *d_in, int ebtype, double tol, double s,
                                   const void *d_qtab_ext, uint8_t *d_out, uint64_t cap,
                                   cudaStream_t st) {
  int rc = ensure_lowlevel_workspace(p);
  if (rc)
    return rc;
  double *d_normout = (double *)(p->d_scalars + 9), *d_red = (double *)(p->d_scalars + 10);
  const void *qtab = d_qtab_ext;
  p->fused_norm.armed = false;
  if (!qtab && ebtype == MGB_REL) {
    // Compressor.hpp:121-129: the norm is only computed for relative bounds
    const bool fuse_norm = is_inf(s) && p->dtype == MGB_F32 && p->D == 3 && !p->force_generic &&
                           p->shape[1] * p->shape[2] < (1ull << 31) && p->L >= 1 &&
                           p->cfg.decomposition == 0 && getenv("MGB_NO_FUSED_NORM") == nullptr;
    if (fuse_norm) {
      if (!p->d_absmax)
        MGB_CUDA_CHECK(cudaMalloc(&p->d_absmax, 8));
      MGB_CUDA_CHECK(cudaMemsetAsync(p->d_absmax, 0, 8, st));
      // decompose_t turns max |x| into the table right behind the coefficient kernel
      p->fused_norm.armed = true;
      p->fused_norm.tol = tol;
      p->fused_norm.s = s;
    } else {
      rc = mgb_norm_async(p, d_in, p->N, d_red, st);
      if (!rc)
        rc = mgb_prepare_quantizers(p, MGB_REL, tol, s, 1, d_red, p->N, 0, p->d_qtab, d_normout, st);
      if (rc)
        return rc;
    }
    qtab = p->d_qtab;
  }
  // s = inf, 3-D: the upper half of the coefficients is quantized while the coarse
  // levels are still being decomposed (refactor.cu: decompose_t)
  p->early_q.armed = is_inf(s) && p->D == 3 && !p->force_generic && p->L >= 3 &&
                     getenv("MGB_NO_EARLY_QUANTIZE") == nullptr && p->cfg.decomposition == 0;
  p->early_q.done = false;
  p->early_q.ebtype = ebtype;
  p->early_q.tol = tol;
  p->early_q.s = s;
  p->early_q.norm = 1.0;
  p->early_q.d_qtab = qtab;
  rc = mgb_decompose_impl(p, d_in, p->d_coef, st);
  p->early_q.armed = false;
  p->fused_norm.armed = false;
  if (rc)
    return rc;
  if (p->early_q.done) {
    MGB_CUDA_CHECK(cudaStreamWaitEvent(st, p->ev_qjoin, 0));
    rc = mgb_quantize_range(p, p->d_coef, ebtype, tol, s, 1.0, p->d_sym, p->d_hist, p->d_scalars,
                            p->d_oidx, p->d_oval, p->outlier_cap, 0, p->early_q.first, 0, 148 * 4, st,
                            qtab);
  } else {
    rc = mgb_quantize_range(p, p->d_coef, ebtype, tol, s, 1.0, p->d_sym, p->d_hist, p->d_scalars,
                            p->d_oidx, p->d_oval, p->outlier_cap, 0, ~0ull, 1, 148 * 4, st, qtab);
  }
  if (rc)
    return rc;
  const uint16_t *sym = p->d_sym;
  if (p->cfg.reorder) {
    // Config::reorder: symbols and outlier positions in level-linearised order
    // (LinearQuantization.hpp:46-146,232-248); the work buffer is free by now
    rc = mgb_linearize_symbols(p, p->d_sym, (uint16_t *)p->d_wA, p->d_scalars, p->d_oidx,
                               p->outlier_cap, st);
    if (rc)
      return rc;
    sym = (const uint16_t *)p->d_wA;
  }
  // index order: deterministic stream
  rc = mgb_sort_outliers(p->d_scalars, p->d_oidx, p->d_oval, p->outlier_cap, st);
  if (rc)
    return rc;
  // speculative: encode assuming the outlier buffer was large enough (checked by the caller
  // after mgb_huffman_finish)
  return mgb_huffman_compress_async(p, sym, p->N, p->d_hist, p->d_scalars, 0, p->d_oidx, p->d_oval,
                                    d_out, cap, st);
}

// After mgb_huffman_finish: the outlier list did not fit (LinearQuantization.hpp:661-675)
// -> grow the buffers; the caller runs the sub-domain again.
static int grow_outlier_buffers(mgb_plan *p, uint64_t oc) {
  cudaFree(p->d_oidx);
  cudaFree(p->d_oval);
  p->d_oidx = nullptr;
  p->d_oval = nullptr;
  while (p->outlier_cap < oc)
    p->outlier_cap <<= 1;
  MGB_CUDA_CHECK(cudaMalloc(&p->d_oidx, p->outlier_cap * 8));
  MGB_CUDA_CHECK(cudaMalloc(&p->d_oval, p->outlier_cap * 8));
  return MGB_SUCCESS;
}

// compress_lowlevel_async + size read-back (one synchronisation), repeated once
// with larger outlier buffers if the list overflowed.  *norm: in (ABS: unused), out
// (REL: the norm the device computed, or the value behind d_qtab_ext's table).
static int compress_lowlevel_sync(mgb_plan *p, const void *d_in, int ebtype, double tol, double s,
                                  const void *d_qtab_ext, double *norm, uint8_t *d_out, uint64_t cap,
                                  uint64_t *size, cudaStream_t st) {
  for (int attempt = 0; attempt < 2; attempt++) {
    int rc = compress_lowlevel_async(p, d_in, ebtype, tol, s, d_qtab_ext, d_out, cap, st);
    if (rc)
      return rc;
    rc = mgb_huffman_finish(p, size, st);
    const uint64_t oc = p->h_pinned[0];
    if (oc <= p->outlier_cap) {
      if (!d_qtab_ext && ebtype == MGB_REL && norm)
        memcpy(norm, &p->h_pinned[9], sizeof(double));
      return rc;
    }
    rc = grow_outlier_buffers(p, oc);
    if (rc)
      return rc;
  }
  return MGB_FAILURE;
}

static int compress_lowlevel_impl(mgb_plan *p, const void *d_in, int ebtype,
                                  double tol, double s, double *norm,
                                  uint8_t *d_out, uint64_t cap, uint64_t *size,
                                  void *stream) {
  if (!p || !d_in || !d_out || !size || !norm)
    return MGB_BAD_ARGUMENT;
  return compress_lowlevel_sync(p, d_in, ebtype, tol, s, nullptr, norm, d_out, cap, size,
                                (cudaStream_t)stream);
}

static int decompress_lowlevel_impl(mgb_plan *p, const uint8_t *d_in, uint64_t size,
                                       int ebtype, double tol, double s, double norm,
                                       void *d_out, void *stream) {
  if (!p || !d_in || !d_out)
    return MGB_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_lowlevel_workspace(p);
  if (rc)
    return rc;
  uint64_t oc = 0;
  const uint64_t *oidx = nullptr;
  const int64_t *oval = nullptr;
  // s = inf: one dequantization factor for every node, applied by the decoder
  // while it flushes its chunks (no symbol array, no dequantize pass)
  double scale = 0;
  const int linear = mgb_linear_dequant_scale(p, ebtype, tol, s, norm, &scale);
  int fused = 0;
  rc = mgb_huffman_decompress_impl(p, d_in, size, p->d_sym, p->N, &oc, &oidx, &oval, st,
                                   linear && !p->cfg.reorder ? p->d_coef : nullptr, scale, &fused);
  if (rc)
    return rc;
  if (p->cfg.reorder) {
    // level-linearised symbols back to the array order; the outlier positions are
    // translated through the inverse map, parked in the coefficient buffer that the
    // dequantizer overwrites next
    if (p->N >= (1ull << 32))
      return MGB_FAILURE;
    if (oc > p->outlier_cap) {
      cudaFree(p->d_oidx);
      cudaFree(p->d_oval);
      p->d_oidx = nullptr;
      p->d_oval = nullptr;
      while (p->outlier_cap < oc)
        p->outlier_cap <<= 1;
      MGB_CUDA_CHECK(cudaMalloc(&p->d_oidx, p->outlier_cap * 8));
      MGB_CUDA_CHECK(cudaMalloc(&p->d_oval, p->outlier_cap * 8));
    }
    rc = mgb_delinearize_symbols(p, p->d_sym, (uint16_t *)p->d_wA, (uint32_t *)p->d_coef, oc, oidx,
                                 p->d_oidx, st);
    if (rc)
      return rc;
    rc = mgb_dequantize(p, (const uint16_t *)p->d_wA, oc, p->d_oidx, oval, ebtype, tol, s, norm,
                        p->d_coef, st);
  } else if (fused)
    rc = mgb_outlier_restore(p, oc, oidx, oval, ebtype, tol, s, norm, p->d_coef, st);
  else
    rc = mgb_dequantize(p, p->d_sym, oc, oidx, oval, ebtype, tol, s, norm, p->d_coef, st);
  if (rc)
    return rc;
  rc = mgb_recompose_impl(p, p->d_coef, d_out, st);
  if (rc)
    return rc;
  MGB_CUDA_CHECK(cudaStreamSynchronize(st));
  return MGB_SUCCESS;
}

// ------------------------------ high level ---------------------------------
namespace {

struct CacheKey {
  int ndim, dtype, dict, chunk;
  uint64_t shape[MGB_MAX_DIMS];
  int dev;
  bool operator<(const CacheKey &o) const {
    return memcmp(this, &o, sizeof(CacheKey)) < 0;
  }
};

struct HighLevelCache {
  std::map<CacheKey, mgb_plan *> plans;
  unsigned char *d_stage = nullptr; // sub-domain input / output staging
  uint64_t stage_bytes = 0;
  unsigned char *d_payload = nullptr; // aligned compressed sub-domain
  uint64_t payload_bytes = 0;
  std::mutex mu;
};
HighLevelCache g_cache;

int ensure_bytes(unsigned char **ptr, uint64_t *have, uint64_t need) {
  if (*have >= need)
    return MGB_SUCCESS;
  cudaFree(*ptr);
  *ptr = nullptr;
  *have = 0;
  MGB_CUDA_CHECK(cudaMalloc(ptr, need));
  *have = need;
  return MGB_SUCCESS;
}


// ---- second-stage lossless: Zstandard on the host ------------------------------
// The reference's Huffman_Zstd stage (include/mgard-x/Lossless/Zstd.hpp:64-125)
// copies the Huffman block to the host, runs ZSTD_compress and stores
// `size_t input_count | zstd frame`.  Same here, through the system's libzstd
// (no header in this image: the four stable C entry points are resolved with
// dlopen).  Without the library the request fails, it is never ignored.
using ZstdApi = mgb_zstd_fns;
const ZstdApi &zstd_api() {
  static ZstdApi api = [] {
    ZstdApi a;
    void *h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!h)
      h = dlopen("libzstd.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      a.compress = (decltype(a.compress))dlsym(h, "ZSTD_compress");
      a.decompress = (decltype(a.decompress))dlsym(h, "ZSTD_decompress");
      a.bound = (decltype(a.bound))dlsym(h, "ZSTD_compressBound");
      a.is_error = (decltype(a.is_error))dlsym(h, "ZSTD_isError");
      a.ok = a.compress && a.decompress && a.bound && a.is_error;
    }
    return a;
  }();
  return api;
}

// uniform-grid plans are cached like the reference's CompressorCache
// (CompressionHighLevel.hpp:89-98); non-uniform ones are rebuilt
// (Hierarchy::can_reuse, Hierarchy.hpp:722-733).
int get_plan(int ndim, int dtype, const uint64_t *shape, const void *const *coords,
             const mgb_config *cfg, mgb_plan **plan, bool *owned) {
  if (coords) {
    *owned = true;
    return mgb_plan_create(ndim, shape, dtype, coords, cfg, plan);
  }
  CacheKey k;
  memset(&k, 0, sizeof(k));
  k.ndim = ndim;
  k.dtype = dtype;
  k.dict = cfg->huff_dict_size;
  k.chunk = cfg->huff_block_size;
  cudaGetDevice(&k.dev);
  for (int d = 0; d < ndim; d++)
    k.shape[d] = shape[d];
  auto it = g_cache.plans.find(k);
  if (it != g_cache.plans.end()) {
    *plan = it->second;
    (*plan)->cfg.reorder = cfg->reorder; // not part of the key: same tables either way
    (*plan)->cfg.decomposition = cfg->decomposition;
    *owned = false;
    return MGB_SUCCESS;
  }
  int rc = mgb_plan_create(ndim, shape, dtype, nullptr, cfg, plan);
  if (rc)
    return rc;
  g_cache.plans[k] = *plan;
  *owned = false;
  return MGB_SUCCESS;
}

bool is_device_pointer(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

struct Partition {
  bool decomposed = false;
  int dim = 0;
  uint64_t size = 0; // chunk size along dim
  uint64_t count = 1;
};

void subdomain_shape(const Partition &pt, int ndim, const uint64_t *shape, uint64_t id,
                     uint64_t *out) {
  for (int d = 0; d < ndim; d++)
    out[d] = shape[d];
  if (!pt.decomposed)
    return;
  // DomainDecomposer.hpp:131-144
  if (id < shape[pt.dim] / pt.size)
    out[pt.dim] = pt.size;
  else
    out[pt.dim] = shape[pt.dim] % pt.size;
}

// copy sub-domain `id` between the full array and a dense buffer
// (DomainDecomposer::copy_subdomain, DomainDecomposer.hpp:649-826)
int copy_subdomain(const Partition &pt, int ndim, const uint64_t *shape, size_t tsize,
                   uint64_t id, const void *full, void *dense, bool to_dense,
                   cudaStream_t st) {
  uint64_t sub[MGB_MAX_DIMS];
  subdomain_shape(pt, ndim, shape, id, sub);
  uint64_t inner = 1, outer = 1;
  const int dim = pt.decomposed ? pt.dim : 0;
  for (int d = dim + 1; d < ndim; d++)
    inner *= shape[d];
  for (int d = 0; d < dim; d++)
    outer *= shape[d];
  uint64_t start = pt.decomposed ? id * pt.size : 0;
  size_t width = sub[dim] * inner * tsize;
  size_t fpitch = shape[dim] * inner * tsize;
  const unsigned char *fp = (const unsigned char *)full + start * inner * tsize;
  if (to_dense)
    MGB_CUDA_CHECK(cudaMemcpy2DAsync(dense, width, fp, fpitch, width, outer,
                                     cudaMemcpyDefault, st));
  else
    MGB_CUDA_CHECK(cudaMemcpy2DAsync((void *)fp, fpitch, dense, width, width, outer,
                                     cudaMemcpyDefault, st));
  return MGB_SUCCESS;
}

// ErrorToleranceCalculator.hpp:134-155, evaluated in T
double local_abs_tol(int dtype, int ebtype, double norm, double tol, double s,
                     uint64_t nsub) {
  if (dtype == MGB_F32) {
    float n = (float)norm, t = (float)tol;
    if (ebtype == MGB_REL)
      return is_inf(s) ? t * n : std::sqrt((t * n) * (t * n) / nsub);
    return is_inf(s) ? t : std::sqrt((t * t) / nsub);
  }
  if (ebtype == MGB_REL)
    return is_inf(s) ? tol * norm : std::sqrt((tol * norm) * (tol * norm) / nsub);
  return is_inf(s) ? tol : std::sqrt((tol * tol) / nsub);
}

int make_partition(int ndim, const uint64_t *shape, size_t tsize, const mgb_config *cfg,
                   Partition &pt) {
  pt = Partition();
  uint64_t S = cfg->domain_decomposition_size;
  int dim = cfg->domain_decomposition_dim;
  if (dim < 0) {
    // MaxDim: the largest dimension (DomainDecomposer.hpp:199-207)
    uint64_t mx = 0;
    for (int d = 0; d < ndim; d++)
      if (shape[d] > mx) {
        mx = shape[d];
        dim = d;
      }
  }
  if (dim >= ndim)
    return MGB_BAD_ARGUMENT;
  if (S == 0) {
    // fit the working set into free device memory by halving the chunk
    // (DomainDecomposer.hpp:208-230)
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    uint64_t rest = 1;
    for (int d = 0; d < ndim; d++)
      if (d != dim)
        rest *= shape[d];
    uint64_t chunk = shape[dim];
    auto footprint = [&](uint64_t c) {
      return (double)c * rest * (tsize * 5.5 + 2.0) + (double)(64ull << 20);
    };
    while (footprint(chunk) > 0.85 * (double)free_b && chunk > 3)
      chunk = (chunk - 1) / 2 + 1;
    S = chunk;
  }
  if (S >= shape[dim]) {
    pt.decomposed = false;
    pt.dim = 0;
    pt.size = shape[0];
    pt.count = 1;
    return MGB_SUCCESS;
  }
  pt.decomposed = true;
  pt.dim = dim;
  pt.size = S;
  pt.count = (shape[dim] - 1) / S + 1;
  uint64_t left = shape[dim] % S;
  if (S < 3 || (left != 0 && left < 3))
    return MGB_BAD_ARGUMENT; // Hierarchy.hpp:748-756
  return MGB_SUCCESS;
}

void header_from(int ndim, int dtype, const uint64_t *shape, double tol, double s,
                 int ebtype, double norm, const void *const *coords,
                 const mgb_config *cfg, const Partition &pt, mgb_header &h) {
  h.ndim = ndim;
  h.dtype = dtype;
  for (int d = 0; d < ndim; d++)
    h.shape[d] = shape[d];
  h.ebtype = ebtype;
  h.tol = tol;
  h.s = s;
  h.norm = norm;
  h.decomposed = pt.decomposed;
  h.dd_dim = pt.dim;
  h.dd_size = pt.size;
  h.dict_size = cfg->huff_dict_size;
  h.block_size = cfg->huff_block_size;
  h.lossless = cfg->lossless;
  h.reorder = cfg->reorder ? 1 : 0;
  h.decomposition = cfg->decomposition == 1 ? 1 : 0;
  h.coords.clear();
  if (coords) {
    h.coords.resize(ndim);
    for (int d = 0; d < ndim; d++) {
      h.coords[d].resize(shape[d]);
      for (uint64_t i = 0; i < shape[d]; i++)
        h.coords[d][i] = dtype == MGB_F32 ? (double)((const float *)coords[d])[i]
                                          : ((const double *)coords[d])[i];
    }
  }
}

// compress sub-domains [first, first+count) into records `u64 size | payload`
// appended to `out` (host or device) starting at *offset.
int compress_records(int ndim, int dtype, const uint64_t *shape, const Partition &pt,
                     double local_tol, double s, int local_eb, double *norm,
                     const void *in_full_or_first, bool in_is_first_subdomain,
                     uint64_t first, uint64_t count, const void *const *coords,
                     const mgb_config *cfg, unsigned char *out, bool out_on_device,
                     uint64_t cap, uint64_t *offset, cudaStream_t st) {
  const size_t tsize = dtype == MGB_F32 ? 4 : 8;
  for (uint64_t id = first; id < first + count; id++) {
    uint64_t sub[MGB_MAX_DIMS];
    subdomain_shape(pt, ndim, shape, id, sub);
    uint64_t nsub = 1;
    for (int d = 0; d < ndim; d++)
      nsub *= sub[d];
    // coordinates of the sub-domain (DomainDecomposer.hpp:283-300)
    const void *subcoords[MGB_MAX_DIMS];
    if (coords) {
      for (int d = 0; d < ndim; d++)
        subcoords[d] = coords[d];
      if (pt.decomposed)
        subcoords[pt.dim] =
            (const unsigned char *)coords[pt.dim] + id * pt.size * tsize;
    }
    mgb_plan *plan = nullptr;
    bool owned = false;
    int rc = get_plan(ndim, dtype, sub, coords ? subcoords : nullptr, cfg, &plan, &owned);
    if (rc)
      return rc;
    const uint64_t raw_bytes = nsub * tsize;
    // dense device copy of the sub-domain
    const void *d_in;
    const bool in_dev = is_device_pointer(in_full_or_first);
    const bool contiguous = !pt.decomposed || pt.dim == 0;
    if (in_is_first_subdomain && !contiguous)
      return MGB_BAD_ARGUMENT;
    uint64_t plane = tsize; // bytes of one index along dim 0
    for (int d = 1; d < ndim; d++)
      plane *= shape[d];
    const unsigned char *sp = nullptr;
    if (contiguous) {
      uint64_t rel = in_is_first_subdomain ? id - first : id;
      sp = (const unsigned char *)in_full_or_first +
           (pt.decomposed ? rel * pt.size * plane : 0);
    }
    if (in_dev && contiguous) {
      d_in = sp;
    } else {
      rc = ensure_bytes(&g_cache.d_stage, &g_cache.stage_bytes, raw_bytes);
      if (rc)
        return rc;
      if (contiguous) {
        MGB_CUDA_CHECK(cudaMemcpyAsync(g_cache.d_stage, sp, raw_bytes,
                                       cudaMemcpyDefault, st));
      } else {
        rc = copy_subdomain(pt, ndim, shape, tsize, id, in_full_or_first,
                            g_cache.d_stage, true, st);
        if (rc)
          return rc;
      }
      d_in = g_cache.d_stage;
    }
    uint64_t pcap = raw_bytes + 2 * (1024 + 8ull * cfg->huff_dict_size) +
                    32 * ((nsub - 1) / cfg->huff_block_size + 1) + 4096;
    // device output whose payload position is 8-byte aligned: compress straight
    // into the record (no staging copy); otherwise through an aligned buffer
    unsigned char *direct = nullptr;
    if (cfg->lossless != 2 && out_on_device && *offset + 8 <= cap &&
        (((uintptr_t)(out + *offset + 8)) & 7) == 0)
      direct = out + *offset + 8;
    if (!direct) {
      rc = ensure_bytes(&g_cache.d_payload, &g_cache.payload_bytes, pcap);
      if (rc)
        return rc;
    }
    uint64_t psize = 0;
    rc = mgb_compress_lowlevel(plan, d_in, local_eb, local_tol, s, norm,
                               direct ? direct : g_cache.d_payload,
                               direct ? std::min<uint64_t>(pcap, cap - *offset - 8) : pcap, &psize,
                               st);
    if (owned)
      mgb_plan_destroy(plan);
    const void *payload = direct ? direct : g_cache.d_payload;
    std::vector<unsigned char> zbuf; // host: size_t count | zstd frame
    if (rc == MGB_SUCCESS && cfg->lossless == 2) {
      const ZstdApi &z = zstd_api();
      if (!z.ok)
        return MGB_FAILURE;
      std::vector<unsigned char> hpay(psize);
      MGB_CUDA_CHECK(cudaMemcpyAsync(hpay.data(), payload, psize, cudaMemcpyDeviceToHost, st));
      MGB_CUDA_CHECK(cudaStreamSynchronize(st));
      zbuf.resize(sizeof(size_t) + z.bound(psize));
      const size_t zs = z.compress(zbuf.data() + sizeof(size_t), zbuf.size() - sizeof(size_t),
                                   hpay.data(), psize, cfg->zstd_compress_level);
      if (z.is_error(zs))
        return MGB_FAILURE;
      const size_t count = psize;
      memcpy(zbuf.data(), &count, sizeof(size_t));
      zbuf.resize(sizeof(size_t) + zs);
      payload = zbuf.data(); // host memory from here on
      psize = zbuf.size();
    }
    if (rc == MGB_OUTPUT_TOO_LARGE || (rc == MGB_SUCCESS && psize >= raw_bytes)) {
      // GPUPipelines.hpp:139-155: store the sub-domain uncompressed
      payload = d_in;
      psize = raw_bytes;
      rc = MGB_SUCCESS;
    }
    if (rc)
      return rc;
    // GPUPipelines.hpp:157-193
    if (*offset + 8 + psize > cap)
      return MGB_OUTPUT_TOO_LARGE;
    uint64_t sz = psize;
    if (out_on_device) {
      MGB_CUDA_CHECK(cudaMemcpyAsync(out + *offset, &sz, 8, cudaMemcpyHostToDevice, st));
      MGB_CUDA_CHECK(cudaStreamSynchronize(st));
    } else {
      memcpy(out + *offset, &sz, 8);
    }
    if (payload != (const void *)(out + *offset + 8)) {
      MGB_CUDA_CHECK(cudaMemcpyAsync(out + *offset + 8, payload, psize, cudaMemcpyDefault, st));
      MGB_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    *offset += 8 + psize;
  }
  return MGB_SUCCESS;
}

int check_args(int ndim, int dtype, const uint64_t *shape) {
  if (!shape)
    return MGB_BAD_ARGUMENT;
  if (ndim < 1 || ndim > MGB_MAX_DIMS)
    return MGB_TOO_MANY_DIMS;
  if (dtype != MGB_F32 && dtype != MGB_F64)
    return MGB_BAD_DTYPE;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return MGB_BACKEND_NOT_AVAILABLE;
  }
  return MGB_SUCCESS;
}

} // namespace

static int compress_impl(int ndim, int dtype, const uint64_t *shape, double tol,
                            double s, int ebtype, const void *in, void **out,
                            size_t *out_size, const void *const *coords,
                            const mgb_config *cfg_in, int output_pre_allocated) {
  int rc = check_args(ndim, dtype, shape);
  if (rc)
    return rc;
  if (!in || !out || !out_size || (output_pre_allocated && !*out))
    return MGB_BAD_ARGUMENT;
  mgb_config cfg;
  if (cfg_in)
    cfg = *cfg_in;
  else
    mgb_config_default(&cfg);
  if (cfg.lossless != 0 && cfg.lossless != 2)
    return MGB_FAILURE; // Huffman_LZ4 (nvcomp) / CPU_Lossless are not built
  std::lock_guard<std::mutex> lock(g_cache.mu);
  const bool in_dev = is_device_pointer(in);
  if (in_dev) {
    cudaPointerAttributes a;
    cudaPointerGetAttributes(&a, in);
    cudaSetDevice(a.device);
  } else if (cfg.dev_id >= 0) {
    cudaSetDevice(cfg.dev_id);
  }
  const size_t tsize = dtype == MGB_F32 ? 4 : 8;
  uint64_t N = 0;
  if (!mgb_checked_elems(ndim, shape, tsize, &N))
    return MGB_BAD_ARGUMENT;
  Partition pt;
  rc = make_partition(ndim, shape, tsize, &cfg, pt);
  if (rc)
    return rc;
  cudaStream_t st = 0;
  double norm = 1;
  double ltol = tol;
  int leb = ebtype;
  if (pt.decomposed) {
    // CompressionHighLevel.hpp:128-139 + ErrorToleranceCalculator.hpp:70-155
    if (ebtype == MGB_REL) {
      double mx = 0, ss = 0;
      for (uint64_t id = 0; id < pt.count; id++) {
        uint64_t sub[MGB_MAX_DIMS];
        subdomain_shape(pt, ndim, shape, id, sub);
        mgb_plan *plan = nullptr;
        bool owned = false;
        mgb_config c2 = cfg;
        rc = get_plan(ndim, dtype, sub, nullptr, &c2, &plan, &owned);
        if (rc)
          return rc;
        uint64_t nsub = plan->N;
        rc = ensure_bytes(&g_cache.d_stage, &g_cache.stage_bytes, nsub * tsize);
        if (rc)
          return rc;
        rc = copy_subdomain(pt, ndim, shape, tsize, id, in, g_cache.d_stage, true, st);
        if (rc)
          return rc;
        MGB_CUDA_CHECK(cudaStreamSynchronize(st));
        double m1, s1;
        rc = mgb_norm_partials(plan, g_cache.d_stage, &m1, &s1);
        if (rc)
          return rc;
        mx = std::max(mx, m1);
        ss += s1;
      }
      if (is_inf(s))
        norm = mx;
      else
        norm = dtype == MGB_F32 ? (double)std::sqrt((float)ss / N) : std::sqrt(ss / N);
      if (dtype == MGB_F32)
        norm = (double)(float)norm;
    }
    ltol = local_abs_tol(dtype, ebtype, norm, tol, s, pt.count);
    leb = MGB_ABS;
  }
  mgb_header h;
  header_from(ndim, dtype, shape, tol, s, ebtype, norm, coords, &cfg, pt, h);
  std::vector<uint8_t> hdr = mgb_encode_stream_header(h);
  uint64_t cap;
  unsigned char *obuf;
  if (!output_pre_allocated) {
    // CompressionHighLevel.hpp:149-158 (OUTPUT_SAFTY_OVERHEAD = 1e6)
    cap = N * tsize + 1000000 + hdr.size() + 8 * pt.count;
    if (in_dev) {
      void *p = nullptr;
      MGB_CUDA_CHECK(cudaMalloc(&p, cap));
      obuf = (unsigned char *)p;
    } else {
      obuf = (unsigned char *)malloc(cap);
      if (!obuf)
        return MGB_FAILURE;
    }
  } else {
    cap = *out_size;
    obuf = (unsigned char *)*out;
  }
  const bool out_dev = is_device_pointer(obuf);
  uint64_t offset = hdr.size();
  if (offset > cap)
    rc = MGB_OUTPUT_TOO_LARGE;
  if (!rc)
    rc = compress_records(ndim, dtype, shape, pt, ltol, s, leb, &norm, in, false, 0,
                          pt.count, coords, &cfg, obuf, out_dev, cap, &offset, st);
  if (!rc) {
    // the norm of a non-decomposed REL run is known only now
    // (CompressionHighLevel.hpp:253-279 serialises the metadata again)
    header_from(ndim, dtype, shape, tol, s, ebtype, norm, coords, &cfg, pt, h);
    std::vector<uint8_t> hdr2 = mgb_encode_stream_header(h);
    if (hdr2.size() != hdr.size())
      rc = MGB_FAILURE;
    else if (out_dev)
      rc = cudaMemcpy(obuf, hdr2.data(), hdr2.size(), cudaMemcpyHostToDevice) == cudaSuccess
               ? MGB_SUCCESS
               : MGB_CUDA_ERROR;
    else
      memcpy(obuf, hdr2.data(), hdr2.size());
  }
  if (rc) {
    if (!output_pre_allocated) {
      if (in_dev)
        cudaFree(obuf);
      else
        free(obuf);
    }
    return rc;
  }
  *out = obuf;
  *out_size = offset;
  return MGB_SUCCESS;
}

static int peek_header_impl(const void *in, size_t in_size, int *ndim, uint64_t *shape,
                               int *dtype, int *ebtype, double *tol, double *s,
                               double *norm, uint64_t *header_bytes) {
  if (!in || in_size < 17)
    return MGB_BAD_ARGUMENT;
  std::vector<uint8_t> head;
  const uint8_t *hp = (const uint8_t *)in;
  if (is_device_pointer(in)) {
    uint8_t pre[17];
    if (cudaMemcpy(pre, in, 17, cudaMemcpyDeviceToHost) != cudaSuccess)
      return MGB_CUDA_ERROR;
    const uint64_t hs = mgb_preamble_header_size(pre, in_size);
    if (hs == UINT64_MAX)
      return MGB_BAD_STREAM;
    head.resize(17 + hs);
    if (cudaMemcpy(head.data(), in, 17 + hs, cudaMemcpyDeviceToHost) != cudaSuccess)
      return MGB_CUDA_ERROR;
    hp = head.data();
    in_size = head.size();
  }
  mgb_header h;
  uint64_t hb = 0;
  int rc = mgb_parse_stream_header(hp, in_size, h, hb);
  if (rc)
    return rc;
  if (ndim) *ndim = h.ndim;
  if (shape)
    for (int d = 0; d < h.ndim; d++)
      shape[d] = h.shape[d];
  if (dtype) *dtype = h.dtype;
  if (ebtype) *ebtype = h.ebtype;
  if (tol) *tol = h.tol;
  if (s) *s = h.s;
  if (norm) *norm = h.norm;
  if (header_bytes) *header_bytes = hb;
  return MGB_SUCCESS;
}

static int decompress_impl(const void *in, size_t in_size, void **out,
                              const mgb_config *cfg_in, int output_pre_allocated,
                              int *ndim_out, uint64_t *shape_out, int *dtype_out) {
  if (!in || !out || (output_pre_allocated && !*out))
    return MGB_BAD_ARGUMENT;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return MGB_BACKEND_NOT_AVAILABLE;
  }
  std::lock_guard<std::mutex> lock(g_cache.mu);
  const bool in_dev = is_device_pointer(in);
  if (in_dev) {
    cudaPointerAttributes a;
    cudaPointerGetAttributes(&a, in);
    cudaSetDevice(a.device);
  } else if (cfg_in && cfg_in->dev_id >= 0) {
    cudaSetDevice(cfg_in->dev_id);
  }
  // header
  std::vector<uint8_t> head;
  const uint8_t *hp = (const uint8_t *)in;
  size_t hsize = in_size;
  if (in_dev) {
    if (in_size < 17)
      return MGB_BAD_STREAM;
    uint8_t pre[17];
    MGB_CUDA_CHECK(cudaMemcpy(pre, in, 17, cudaMemcpyDeviceToHost));
    const uint64_t hs = mgb_preamble_header_size(pre, in_size);
    if (hs == UINT64_MAX)
      return MGB_BAD_STREAM;
    head.resize(17 + hs);
    MGB_CUDA_CHECK(cudaMemcpy(head.data(), in, 17 + hs, cudaMemcpyDeviceToHost));
    hp = head.data();
    hsize = head.size();
  }
  mgb_header h;
  uint64_t hb = 0;
  int rc = mgb_parse_stream_header(hp, hsize, h, hb);
  if (rc)
    return rc;
  if (h.convention != 0)
    return MGB_BAD_STREAM; // MGARD-CPU stream: mgb_cpu_decompress reads those
  mgb_config cfg;
  mgb_config_default(&cfg);
  if (cfg_in)
    cfg.dev_id = cfg_in->dev_id;
  // Metadata.cpp:129-136: the header overrides the configuration
  cfg.huff_dict_size = h.dict_size;
  cfg.huff_block_size = h.block_size;
  cfg.lossless = h.lossless;
  cfg.reorder = h.reorder;
  cfg.decomposition = h.decomposition;
  const int ndim = h.ndim, dtype = h.dtype;
  const size_t tsize = dtype == MGB_F32 ? 4 : 8;
  uint64_t N = 0;
  for (int d = 0; d < ndim; d++)
    if (h.shape[d] < 3)
      return MGB_BAD_STREAM;
  // a header that passes its CRC can still announce an absurd shape
  if (!mgb_checked_elems(ndim, h.shape, tsize, &N))
    return MGB_BAD_STREAM;
  Partition pt;
  pt.decomposed = h.decomposed;
  pt.dim = (int)h.dd_dim;
  pt.size = h.dd_size;
  pt.count = 1;
  if (pt.decomposed) {
    if (pt.dim >= ndim || pt.size < 3 || pt.size >= h.shape[pt.dim])
      return MGB_BAD_STREAM;
    pt.count = (h.shape[pt.dim] - 1) / pt.size + 1;
  }
  // CompressionHighLevel.hpp:456-464: coordinates go through float
  std::vector<std::vector<unsigned char>> cbytes;
  const void *cptr[MGB_MAX_DIMS];
  const bool nonuniform = !h.coords.empty();
  if (nonuniform) {
    cbytes.resize(ndim);
    for (int d = 0; d < ndim; d++) {
      cbytes[d].resize(h.shape[d] * tsize);
      for (uint64_t i = 0; i < h.shape[d]; i++) {
        float f = (float)h.coords[d][i];
        if (dtype == MGB_F32)
          ((float *)cbytes[d].data())[i] = f;
        else
          ((double *)cbytes[d].data())[i] = f;
      }
      cptr[d] = cbytes[d].data();
    }
  }
  double ltol = h.tol;
  int leb = h.ebtype;
  if (pt.decomposed) {
    ltol = local_abs_tol(dtype, h.ebtype, h.norm, h.tol, h.s, pt.count);
    leb = MGB_ABS;
  }
  unsigned char *obuf;
  if (!output_pre_allocated) {
    if (in_dev) {
      void *p = nullptr;
      MGB_CUDA_CHECK(cudaMalloc(&p, N * tsize));
      obuf = (unsigned char *)p;
    } else {
      obuf = (unsigned char *)malloc(N * tsize);
      if (!obuf)
        return MGB_FAILURE;
    }
  } else {
    obuf = (unsigned char *)*out;
  }
  const bool out_dev = is_device_pointer(obuf);
  cudaStream_t st = 0;
  uint64_t offset = hb;
  const unsigned char *ip = (const unsigned char *)in;
  for (uint64_t id = 0; id < pt.count && !rc; id++) {
    uint64_t sub[MGB_MAX_DIMS];
    subdomain_shape(pt, ndim, h.shape, id, sub);
    uint64_t nsub = 1;
    for (int d = 0; d < ndim; d++)
      nsub *= sub[d];
    const uint64_t raw_bytes = nsub * tsize;
    if (offset + 8 > in_size) {
      rc = MGB_BAD_STREAM;
      break;
    }
    uint64_t psize = 0;
    if (in_dev)
      cudaMemcpy(&psize, ip + offset, 8, cudaMemcpyDeviceToHost);
    else
      memcpy(&psize, ip + offset, 8);
    offset += 8;
    if (psize > in_size - offset) {
      rc = MGB_BAD_STREAM;
      break;
    }
    // dense device destination for this sub-domain
    const bool contiguous = !pt.decomposed || pt.dim == 0;
    unsigned char *d_dst;
    uint64_t plane = raw_bytes / sub[pt.decomposed ? pt.dim : 0];
    unsigned char *final_dst = obuf + (pt.decomposed && contiguous ? id * pt.size * plane : 0);
    if (out_dev && contiguous) {
      d_dst = final_dst;
    } else {
      rc = ensure_bytes(&g_cache.d_stage, &g_cache.stage_bytes, raw_bytes);
      if (rc)
        break;
      d_dst = g_cache.d_stage;
    }
    if (psize >= raw_bytes) {
      // raw sub-domain (GPUPipelines.hpp:417,458-466)
      if (cudaMemcpyAsync(d_dst, ip + offset, raw_bytes, cudaMemcpyDefault, st) != cudaSuccess)
        rc = MGB_CUDA_ERROR;
    } else {
      const void *subcoords[MGB_MAX_DIMS];
      if (nonuniform) {
        for (int d = 0; d < ndim; d++)
          subcoords[d] = cptr[d];
        if (pt.decomposed)
          subcoords[pt.dim] = (const unsigned char *)cptr[pt.dim] + id * pt.size * tsize;
      }
      mgb_plan *plan = nullptr;
      bool owned = false;
      rc = get_plan(ndim, dtype, sub, nonuniform ? subcoords : nullptr, &cfg, &plan, &owned);
      if (rc)
        break;
      uint64_t hsize = psize; // size of the Huffman block
      if (h.lossless == 2) {
        // Zstd.hpp:100-125: `size_t count | zstd frame` -> Huffman block, on the host
        const ZstdApi &z = zstd_api();
        std::vector<unsigned char> rec(psize), hpay;
        if (!z.ok || psize < sizeof(size_t))
          rc = z.ok ? MGB_BAD_STREAM : MGB_FAILURE;
        if (!rc && cudaMemcpy(rec.data(), ip + offset, psize, cudaMemcpyDefault) != cudaSuccess)
          rc = MGB_CUDA_ERROR;
        if (!rc) {
          size_t count = 0;
          memcpy(&count, rec.data(), sizeof(size_t));
          if (count > (size_t)raw_bytes * 2 + (1u << 24))
            rc = MGB_BAD_STREAM;
          if (!rc) {
            hpay.resize(count);
            const size_t got = z.decompress(hpay.data(), count, rec.data() + sizeof(size_t),
                                            psize - sizeof(size_t));
            if (z.is_error(got) || got != count)
              rc = MGB_BAD_STREAM;
          }
          hsize = count;
        }
        if (!rc)
          rc = ensure_bytes(&g_cache.d_payload, &g_cache.payload_bytes, hsize + 64);
        if (!rc && cudaMemcpy(g_cache.d_payload, hpay.data(), hsize, cudaMemcpyHostToDevice) !=
                       cudaSuccess)
          rc = MGB_CUDA_ERROR;
      } else {
        rc = ensure_bytes(&g_cache.d_payload, &g_cache.payload_bytes, psize + 64);
        if (!rc && cudaMemcpyAsync(g_cache.d_payload, ip + offset, psize, cudaMemcpyDefault,
                                   st) != cudaSuccess)
          rc = MGB_CUDA_ERROR;
      }
      if (!rc)
        rc = mgb_decompress_lowlevel(plan, g_cache.d_payload, hsize, leb, ltol, h.s,
                                     h.norm, d_dst, st);
      if (owned)
        mgb_plan_destroy(plan);
      if (rc)
        break;
    }
    if (d_dst != final_dst || !contiguous) {
      if (contiguous) {
        if (cudaMemcpyAsync(final_dst, d_dst, raw_bytes, cudaMemcpyDefault, st) != cudaSuccess)
          rc = MGB_CUDA_ERROR;
      } else {
        rc = copy_subdomain(pt, ndim, h.shape, tsize, id, obuf, d_dst, false, st);
      }
    }
    if (cudaStreamSynchronize(st) != cudaSuccess)
      rc = MGB_CUDA_ERROR;
    offset += psize;
  }
  if (rc) {
    if (!output_pre_allocated) {
      if (in_dev)
        cudaFree(obuf);
      else
        free(obuf);
    }
    return rc;
  }
  *out = obuf;
  if (ndim_out) *ndim_out = ndim;
  if (dtype_out) *dtype_out = dtype;
  if (shape_out)
    for (int d = 0; d < ndim; d++)
      shape_out[d] = h.shape[d];
  return MGB_SUCCESS;
}

extern "C" void mgb_release_cache(void) {
  std::lock_guard<std::mutex> lock(g_cache.mu);
  for (auto &kv : g_cache.plans)
    mgb_plan_destroy(kv.second);
  g_cache.plans.clear();
  cudaFree(g_cache.d_stage);
  cudaFree(g_cache.d_payload);
  g_cache.d_stage = g_cache.d_payload = nullptr;
  g_cache.stage_bytes = g_cache.payload_bytes = 0;
}

static int compress_subdomains_impl(int ndim, int dtype, const uint64_t *shape,
                                       double tol, double s, int ebtype, double norm,
                                       const void *d_in_first, uint64_t first,
                                       uint64_t count, const mgb_config *cfg_in,
                                       uint8_t *d_out, uint64_t cap, uint64_t *size) {
  int rc = check_args(ndim, dtype, shape);
  if (rc)
    return rc;
  if (!d_in_first || !d_out || !size || !cfg_in)
    return MGB_BAD_ARGUMENT;
  std::lock_guard<std::mutex> lock(g_cache.mu);
  const size_t tsize = dtype == MGB_F32 ? 4 : 8;
  uint64_t nall = 0;
  if (!mgb_checked_elems(ndim, shape, tsize, &nall))
    return MGB_BAD_ARGUMENT;
  Partition pt;
  rc = make_partition(ndim, shape, tsize, cfg_in, pt);
  if (rc)
    return rc;
  if (!pt.decomposed || pt.dim != 0 || first + count > pt.count)
    return MGB_BAD_ARGUMENT;
  double ltol = local_abs_tol(dtype, ebtype, norm, tol, s, pt.count);
  uint64_t offset = 0;
  double nrm = norm;
  rc = compress_records(ndim, dtype, shape, pt, ltol, s, MGB_ABS, &nrm, d_in_first, true,
                        first, count, nullptr, cfg_in, d_out, true, cap, &offset, 0);
  if (rc)
    return rc;
  *size = offset;
  return MGB_SUCCESS;
}

static int write_header_impl(int ndim, int dtype, const uint64_t *shape, double tol,
                                double s, int ebtype, double norm,
                                const void *const *coords, const mgb_config *cfg_in,
                                uint8_t *out, uint64_t cap, uint64_t *size) {
  if (!shape || !out || !size || !cfg_in)
    return MGB_BAD_ARGUMENT;
  const size_t tsize = dtype == MGB_F32 ? 4 : 8;
  Partition pt;
  int rc = make_partition(ndim, shape, tsize, cfg_in, pt);
  if (rc)
    return rc;
  mgb_header h;
  header_from(ndim, dtype, shape, tol, s, ebtype, norm, coords, cfg_in, pt, h);
  std::vector<uint8_t> hdr = mgb_encode_stream_header(h);
  *size = hdr.size();
  if (hdr.size() > cap)
    return MGB_OUTPUT_TOO_LARGE;
  memcpy(out, hdr.data(), hdr.size());
  return MGB_SUCCESS;
}

extern "C" int mgb_pin_memory(void *ptr, uint64_t num_bytes) {
  if (!ptr || !num_bytes)
    return MGB_BAD_ARGUMENT;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return MGB_BACKEND_NOT_AVAILABLE;
  }
  cudaError_t e = cudaHostRegister(ptr, num_bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) {
    cudaGetLastError();
    return MGB_SUCCESS;
  }
  return e == cudaSuccess ? MGB_SUCCESS : MGB_CUDA_ERROR;
}
extern "C" int mgb_check_memory_pinned(const void *ptr) {
  cudaPointerAttributes a;
  if (!ptr || cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return a.type == cudaMemoryTypeHost ? 1 : 0;
}
extern "C" int mgb_unpin_memory(void *ptr) {
  if (!ptr)
    return MGB_BAD_ARGUMENT;
  cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess)
    cudaGetLastError();
  return e == cudaSuccess ? MGB_SUCCESS : MGB_CUDA_ERROR;
}

// ---- exported entry points: no exception crosses the C ABI --------------------
extern "C" int mgb_compress_lowlevel(mgb_plan *p, const void *d_in, int ebtype, double tol, double s,
                                     double *norm, uint8_t *d_out, uint64_t cap, uint64_t *size,
                                     void *stream) {
  MGB_NOEXCEPT_CALL(compress_lowlevel_impl(p, d_in, ebtype, tol, s, norm, d_out, cap, size, stream));
}
extern "C" int mgb_decompress_lowlevel(mgb_plan *p, const uint8_t *d_in, uint64_t size, int ebtype,
                                       double tol, double s, double norm, void *d_out, void *stream) {
  MGB_NOEXCEPT_CALL(decompress_lowlevel_impl(p, d_in, size, ebtype, tol, s, norm, d_out, stream));
}
extern "C" int mgb_compress(int ndim, int dtype, const uint64_t *shape, double tol, double s, int ebtype,
                            const void *in, void **out, size_t *out_size, const void *const *coords,
                            const mgb_config *cfg_in, int output_pre_allocated) {
  MGB_NOEXCEPT_CALL(
      compress_impl(ndim, dtype, shape, tol, s, ebtype, in, out, out_size, coords, cfg_in, output_pre_allocated));
}
extern "C" int mgb_peek_header(const void *in, size_t in_size, int *ndim, uint64_t *shape, int *dtype,
                               int *ebtype, double *tol, double *s, double *norm, uint64_t *header_bytes) {
  MGB_NOEXCEPT_CALL(peek_header_impl(in, in_size, ndim, shape, dtype, ebtype, tol, s, norm, header_bytes));
}
extern "C" int mgb_decompress(const void *in, size_t in_size, void **out, const mgb_config *cfg_in,
                              int output_pre_allocated, int *ndim_out, uint64_t *shape_out, int *dtype_out) {
  MGB_NOEXCEPT_CALL(
      decompress_impl(in, in_size, out, cfg_in, output_pre_allocated, ndim_out, shape_out, dtype_out));
}
extern "C" int mgb_compress_subdomains(int ndim, int dtype, const uint64_t *shape, double tol, double s,
                                       int ebtype, double norm, const void *d_in_first, uint64_t first,
                                       uint64_t count, const mgb_config *cfg_in, uint8_t *d_out,
                                       uint64_t cap, uint64_t *size) {
  MGB_NOEXCEPT_CALL(compress_subdomains_impl(ndim, dtype, shape, tol, s, ebtype, norm, d_in_first, first,
                                             count, cfg_in, d_out, cap, size));
}
extern "C" int mgb_write_header(int ndim, int dtype, const uint64_t *shape, double tol, double s,
                                int ebtype, double norm, const void *const *coords,
                                const mgb_config *cfg_in, uint8_t *out, uint64_t cap, uint64_t *size) {
  MGB_NOEXCEPT_CALL(write_header_impl(ndim, dtype, shape, tol, s, ebtype, norm, coords, cfg_in, out, cap, size));
}

extern "C" uint64_t mgb_launch_count(void) { return g_mgb_launches; }
extern "C" const char *mgb_version(void) { return "mgard_b200 0.1 (sm_100a)"; }

// shared with cpu_convention.cu (CPU_HUFFMAN_ZSTD payload)
const mgb_zstd_fns &mgb_zstd() { return zstd_api(); }
