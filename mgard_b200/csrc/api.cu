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
int mgb_norm_raw(int dtype, const void *d_in, uint64_t n, double *d_part, double *d_red, cudaStream_t st);
int mgb_norm_combine(const double *d_pairs, int count, double *d_red, cudaStream_t st);
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
    // + 16: the load-vector kernel's bulk copies move whole 16-byte pieces, the last of which
    // may straddle the end of the array (masstrans3d.cuh)
    MGB_CUDA_CHECK(cudaMalloc(&p->d_coef, p->N * p->tsize + 16));
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
// front: norm, decomposition, quantization + histogram, outlier order (the block's
// destination is not needed yet); back: Huffman into d_out.
static int compress_lowlevel_front(mgb_plan *p, const void *d_in, int ebtype, double tol, double s,
                                   const void *d_qtab_ext, cudaStream_t st) {
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
  // Only below 2^28 nodes: the two launches together take longer than one over the whole
  // array (1.25 + 0.81 against 1.57 ms at 257 x 2049^2), which pays while the coarse levels are
  // a latency-bound chain worth hiding (513^3) and costs 0.47 ms per compression at 257 x 2049^2.
  p->early_q.armed = is_inf(s) && p->D == 3 && !p->force_generic && p->L >= 3 && p->N < (1ull << 28) &&
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
  if (p->cfg.reorder) {
    // Config::reorder: symbols and outlier positions in level-linearised order
    // (LinearQuantization.hpp:46-146,232-248); the work buffer is free by now
    rc = mgb_linearize_symbols(p, p->d_sym, (uint16_t *)p->d_wA, p->d_scalars, p->d_oidx,
                               p->outlier_cap, st);
    if (rc)
      return rc;
  }
  // index order: deterministic stream
  return mgb_sort_outliers(p->d_scalars, p->d_oidx, p->d_oval, p->outlier_cap, st);
}

// speculative: encodes assuming the outlier buffer was large enough (the flag read back
// with the block size tells)
static int compress_lowlevel_back(mgb_plan *p, uint8_t *d_out, uint64_t cap, cudaStream_t st) {
  const uint16_t *sym = p->cfg.reorder ? (const uint16_t *)p->d_wA : p->d_sym;
  return mgb_huffman_compress_async(p, sym, p->N, p->d_hist, p->d_scalars, 0, p->d_oidx, p->d_oval,
                                    d_out, cap, st);
}

static int compress_lowlevel_async(mgb_plan *p, const void *d_in, int ebtype, double tol, double s,
                                   const void *d_qtab_ext, uint8_t *d_out, uint64_t cap,
                                   cudaStream_t st) {
  int rc = compress_lowlevel_front(p, d_in, ebtype, tol, s, d_qtab_ext, st);
  if (rc)
    return rc;
  return compress_lowlevel_back(p, d_out, cap, st);
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

// Compressor::Decompress without the final synchronisation (the Huffman header is
// still read back to size the views, Huffman.hpp:264-320)
static int decompress_lowlevel_async(mgb_plan *p, const uint8_t *d_in, uint64_t size,
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
  return mgb_recompose_impl(p, p->d_coef, d_out, st);
}

static int decompress_lowlevel_impl(mgb_plan *p, const uint8_t *d_in, uint64_t size, int ebtype,
                                    double tol, double s, double norm, void *d_out, void *stream) {
  int rc = decompress_lowlevel_async(p, d_in, size, ebtype, tol, s, norm, d_out, stream);
  if (rc)
    return rc;
  MGB_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  return MGB_SUCCESS;
}

// ------------------------------ high level ---------------------------------
namespace {

struct CacheKey {
  int ndim, dtype, dict, chunk;
  uint64_t max_level;
  uint64_t shape[MGB_MAX_DIMS];
  int dev;
  bool operator<(const CacheKey &o) const {
    return memcmp(this, &o, sizeof(CacheKey)) < 0;
  }
};

struct HighLevelCache {
  std::map<CacheKey, mgb_plan *> plans;
  std::map<CacheKey, uint64_t> last_use; // call stamp of the last use (eviction order)
  uint64_t stamp = 0;
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
  k.max_level = cfg->max_larget_level;
  cudaGetDevice(&k.dev);
  for (int d = 0; d < ndim; d++)
    k.shape[d] = shape[d];
  g_cache.last_use[k] = g_cache.stamp;
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

// The reference's CompressorCache holds one hierarchy per <D, T> and rebuilds it when
// the shape changes; this cache keeps several shapes (the sub-domains of a decomposed
// domain alternate between two) but not without bound: at the START of a high-level call
// (nothing of this library is in flight then) the least recently used plans beyond
// MAX_PLANS are destroyed together with their workspaces.
void trim_plan_cache() {
  constexpr size_t MAX_PLANS = 12;
  g_cache.stamp++;
  while (g_cache.plans.size() > MAX_PLANS) {
    auto victim = g_cache.plans.begin();
    for (auto it = g_cache.plans.begin(); it != g_cache.plans.end(); ++it)
      if (g_cache.last_use[it->first] < g_cache.last_use[victim->first])
        victim = it;
    mgb_plan_destroy(victim->second);
    g_cache.last_use.erase(victim->first);
    g_cache.plans.erase(victim);
  }
}

bool is_device_pointer(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// DomainDecomposer (DomainDecomposer/DomainDecomposer.hpp:90-169,199-262): MaxDim and
// Variable cut one dimension (equal extents + a remainder / caller-given extents), Block
// cuts every dimension into edges of `size`; sub-domain ids run row-major over the cuts.
struct Partition {
  bool decomposed = false;
  int method = 1; // pb::DomainDecomposition::Method: 1 MAX_DIMENSION, 2 BLOCK, 3 VARIABLE
  int dim = 0;
  uint64_t size = 0; // extent of a full cut
  uint64_t count = 1;
  uint64_t cuts[MGB_MAX_DIMS] = {1, 1, 1, 1, 1}; // sub-domains per dimension
  std::vector<uint64_t> var_sizes, var_offsets;  // Variable
  // sub-domains are consecutive runs of planes of dim 0 (copied / used in place as one range)
  bool slabs() const { return !decomposed || (method != 2 && dim == 0); }
};

// extent and first index of sub-domain `id` per dimension
void subdomain_box(const Partition &pt, int ndim, const uint64_t *shape, uint64_t id,
                   uint64_t *ext, uint64_t *off) {
  for (int d = 0; d < ndim; d++) {
    ext[d] = shape[d];
    off[d] = 0;
  }
  if (!pt.decomposed)
    return;
  if (pt.method == 3) {
    ext[pt.dim] = pt.var_sizes[id];
    off[pt.dim] = pt.var_offsets[id];
  } else if (pt.method == 2) {
    // DomainDecomposer.hpp:104-111,146-157
    for (int d = ndim - 1; d >= 0; d--) {
      const uint64_t k = id % pt.cuts[d];
      id /= pt.cuts[d];
      ext[d] = k < shape[d] / pt.size ? pt.size : shape[d] % pt.size;
      off[d] = k * pt.size;
    }
  } else {
    // DomainDecomposer.hpp:131-144
    ext[pt.dim] = id < shape[pt.dim] / pt.size ? pt.size : shape[pt.dim] % pt.size;
    off[pt.dim] = id * pt.size;
  }
}
// slab partitions: first plane (index along dim 0) of sub-domain `id`
uint64_t subdomain_first_plane(const Partition &pt, uint64_t id) {
  if (!pt.decomposed)
    return 0;
  return pt.method == 3 ? pt.var_offsets[id] : id * pt.size;
}
// coordinates of sub-domain `id`: every cut dimension starts at its own offset
// (DomainDecomposer.hpp:258-300)
void subdomain_coords(const Partition &pt, int ndim, const uint64_t *shape, size_t tsize, uint64_t id,
                      const void *const *coords, const void **out) {
  uint64_t ext[MGB_MAX_DIMS], off[MGB_MAX_DIMS];
  subdomain_box(pt, ndim, shape, id, ext, off);
  for (int d = 0; d < ndim; d++)
    out[d] = (const unsigned char *)coords[d] + off[d] * tsize;
}
void subdomain_shape(const Partition &pt, int ndim, const uint64_t *shape, uint64_t id,
                     uint64_t *out) {
  uint64_t off[MGB_MAX_DIMS];
  subdomain_box(pt, ndim, shape, id, out, off);
}

// copy sub-domain `id` between the full array and a dense buffer
// (DomainDecomposer::copy_subdomain, DomainDecomposer.hpp:649-826)
int copy_subdomain(const Partition &pt, int ndim, const uint64_t *shape, size_t tsize,
                   uint64_t id, const void *full, void *dense, bool to_dense,
                   cudaStream_t st) {
  uint64_t sub[MGB_MAX_DIMS], off[MGB_MAX_DIMS];
  subdomain_box(pt, ndim, shape, id, sub, off);
  if (pt.decomposed && pt.method == 2) {
    // a D-dimensional box: one pitched copy per index of the dimensions above the last two
    uint64_t fstride[MGB_MAX_DIMS], dstride[MGB_MAX_DIMS];
    uint64_t a = 1, b = 1;
    for (int d = ndim - 1; d >= 0; d--) {
      fstride[d] = a;
      dstride[d] = b;
      a *= shape[d];
      b *= sub[d];
    }
    const int nlead = ndim >= 2 ? ndim - 2 : 0;
    uint64_t lead = 1;
    for (int d = 0; d < nlead; d++)
      lead *= sub[d];
    const size_t width = sub[ndim - 1] * tsize;
    const uint64_t rows = ndim >= 2 ? sub[ndim - 2] : 1;
    const size_t fpitch = shape[ndim - 1] * tsize;
    for (uint64_t k = 0; k < lead; k++) {
      uint64_t rem = k, fo = 0, dn = 0;
      for (int d = nlead - 1; d >= 0; d--) {
        const uint64_t i = rem % sub[d];
        rem /= sub[d];
        fo += (off[d] + i) * fstride[d];
        dn += i * dstride[d];
      }
      if (ndim >= 2)
        fo += off[ndim - 2] * fstride[ndim - 2];
      fo += off[ndim - 1];
      const unsigned char *fp = (const unsigned char *)full + fo * tsize;
      unsigned char *dp = (unsigned char *)dense + dn * tsize;
      if (to_dense)
        MGB_CUDA_CHECK(cudaMemcpy2DAsync(dp, width, fp, fpitch, width, rows, cudaMemcpyDefault, st));
      else
        MGB_CUDA_CHECK(cudaMemcpy2DAsync((void *)fp, fpitch, dp, width, width, rows, cudaMemcpyDefault, st));
    }
    return MGB_SUCCESS;
  }
  uint64_t inner = 1, outer = 1;
  const int dim = pt.decomposed ? pt.dim : 0;
  for (int d = dim + 1; d < ndim; d++)
    inner *= shape[d];
  for (int d = 0; d < dim; d++)
    outer *= shape[d];
  uint64_t start = pt.decomposed ? off[dim] : 0;
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
  size_t free_b = 0, total_b = 0;
  auto budget = [&]() {
    if (!total_b)
      cudaMemGetInfo(&free_b, &total_b);
    return std::min<double>(0.85 * (double)free_b, (double)cfg->max_memory_footprint);
  };
  // working set of one sub-domain of `elems` nodes (input copy, coefficients, symbols,
  // coarse boxes and correction workspaces, payload)
  auto footprint = [&](double elems) { return elems * (tsize * 5.5 + 2.0) + (double)(64ull << 20); };
  if (cfg->domain_decomposition == 2) {
    // Variable: caller-given extents along domain_decomposition_dim (DomainDecomposer.hpp:335-348)
    const int dim = cfg->domain_decomposition_dim < 0 ? 0 : cfg->domain_decomposition_dim;
    if (dim >= ndim || !cfg->domain_decomposition_sizes || !cfg->num_domain_decomposition_sizes)
      return MGB_BAD_ARGUMENT;
    uint64_t sum = 0;
    for (uint64_t k = 0; k < cfg->num_domain_decomposition_sizes; k++) {
      const uint64_t e = cfg->domain_decomposition_sizes[k];
      if (e < 3)
        return MGB_BAD_ARGUMENT; // Hierarchy.hpp:748-756
      pt.var_offsets.push_back(sum);
      pt.var_sizes.push_back(e);
      sum += e;
    }
    if (sum != shape[dim])
      return MGB_BAD_ARGUMENT;
    pt.decomposed = true;
    pt.method = 3;
    pt.dim = dim;
    pt.size = pt.var_sizes[0];
    pt.count = pt.var_sizes.size();
    pt.cuts[dim] = pt.count;
    return MGB_SUCCESS;
  }
  if (cfg->domain_decomposition == 1) {
    // Block: every dimension in edges of block_size, halved until one block fits
    // (DomainDecomposer.hpp:232-256,329-334); always decomposed, even into one block
    uint64_t S = cfg->block_size;
    if (S < 3)
      return MGB_BAD_ARGUMENT;
    while (footprint(std::pow((double)S, ndim)) > budget() && S > 3)
      S = (S - 1) / 2 + 1;
    pt.decomposed = true;
    pt.method = 2;
    pt.dim = 0;
    pt.size = S;
    pt.count = 1;
    for (int d = 0; d < ndim; d++) {
      pt.cuts[d] = (shape[d] - 1) / S + 1;
      pt.count *= pt.cuts[d];
      const uint64_t left = shape[d] % S;
      if (left != 0 && left < 3)
        return MGB_BAD_ARGUMENT; // Hierarchy.hpp:748-756
    }
    return MGB_SUCCESS;
  }
  if (cfg->domain_decomposition != 0)
    return MGB_BAD_ARGUMENT;
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
    // fit the working set into (free device memory, max_memory_footprint) by halving
    // the chunk (DomainDecomposer.hpp:208-230)
    uint64_t rest = 1;
    for (int d = 0; d < ndim; d++)
      if (d != dim)
        rest *= shape[d];
    uint64_t chunk = shape[dim];
    while (footprint((double)chunk * rest) > budget() && chunk > 3)
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
  pt.method = 1;
  pt.dim = dim;
  pt.size = S;
  pt.count = (shape[dim] - 1) / S + 1;
  pt.cuts[dim] = pt.count;
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
  h.dd_method = pt.method;
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

// ---- NCCL, resolved at run time ---------------------------------------------------
// The library does not link NCCL: one-process-per-GPU callers (torch.distributed,
// MPI + NCCL) already have libnccl.so.2 in the process; it is looked up with dlopen.
struct mgb_nccl_uid {
  char internal[128];
};
struct NcclApi {
  int (*GetUniqueId)(mgb_nccl_uid *) = nullptr;
  int (*CommInitRank)(void **, int, mgb_nccl_uid, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok = false;
};
const NcclApi &nccl_api() {
  static NcclApi api = [] {
    NcclApi a;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h)
      h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h)
      h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
      a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
      a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
      a.AllReduce = (decltype(a.AllReduce))dlsym(h, "ncclAllReduce");
      a.AllGather = (decltype(a.AllGather))dlsym(h, "ncclAllGather");
      a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
      a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.AllGather;
    }
    return a;
  }();
  return api;
}
enum { NCCL_UINT64 = 5, NCCL_FLOAT64 = 8, NCCL_SUM = 0 };
#define MGB_NCCL_CHECK(x)                                                                \
  do {                                                                                   \
    int e_ = (x);                                                                        \
    if (e_ != 0) {                                                                       \
      const NcclApi &a_ = nccl_api();                                                    \
      fprintf(stderr, "mgard_b200: NCCL error %s at %s:%d\n",                            \
              a_.GetErrorString ? a_.GetErrorString(e_) : "?", __FILE__, __LINE__);      \
      return MGB_FAILURE;                                                                \
    }                                                                                    \
  } while (0)

} // namespace

struct mgb_comm {
  void *nccl = nullptr; // ncclComm_t
  int rank = 0, nranks = 1;
  bool owned = false;
};

namespace {

// ---- per-device resources of the high-level calls -----------------------------------
// Three non-blocking streams as the reference's pipeline has three queues per device
// (GPUPipelines.hpp:88-207): host-to-device staging, compute, device-to-host drain;
// two staging slots each way.
struct DevRes {
  cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr};
  cudaEvent_t ev_entry = nullptr;
  unsigned char *stage[2] = {nullptr, nullptr};
  uint64_t stage_bytes[2] = {0, 0};
  unsigned char *payload[2] = {nullptr, nullptr};
  uint64_t payload_bytes[2] = {0, 0};
  unsigned char *d_full = nullptr; // whole local block staged once (relative bounds, host input)
  uint64_t full_bytes = 0;
  double *d_red = nullptr;   // [0..1] {max |x|, sum x^2} of the domain, [2] norm, [4..] reduction scratch
  double *d_pairs = nullptr; // {max, sum} per sub-domain
  uint64_t pairs_cap = 0;
  unsigned long long *d_sizes = nullptr; // size exchange: [0] mine, [8..8+nranks) all
  uint64_t sizes_cap = 0;
  unsigned long long *h_pin = nullptr; // pinned scratch
};
std::map<int, DevRes> g_devres;

int dev_res(DevRes **out) {
  int dev = 0;
  MGB_CUDA_CHECK(cudaGetDevice(&dev));
  DevRes &r = g_devres[dev];
  if (!r.s_comp) {
    MGB_CUDA_CHECK(cudaStreamCreateWithFlags(&r.s_in, cudaStreamNonBlocking));
    MGB_CUDA_CHECK(cudaStreamCreateWithFlags(&r.s_comp, cudaStreamNonBlocking));
    MGB_CUDA_CHECK(cudaStreamCreateWithFlags(&r.s_out, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
      MGB_CUDA_CHECK(cudaEventCreateWithFlags(&r.ev_in[k], cudaEventDisableTiming));
      MGB_CUDA_CHECK(cudaEventCreateWithFlags(&r.ev_out[k], cudaEventDisableTiming));
      MGB_CUDA_CHECK(cudaEventCreateWithFlags(&r.ev_comp[k], cudaEventDisableTiming));
    }
    MGB_CUDA_CHECK(cudaEventCreateWithFlags(&r.ev_entry, cudaEventDisableTiming));
    MGB_CUDA_CHECK(cudaMalloc(&r.d_red, (4 + MGB_NORM_PART_DOUBLES) * sizeof(double)));
    MGB_CUDA_CHECK(cudaMallocHost(&r.h_pin, 64 * sizeof(unsigned long long)));
  }
  *out = &r;
  return MGB_SUCCESS;
}

// Work issued before the call on the legacy default stream (and streams that
// synchronise with it) is ordered before ours: the caller's buffers are ready.
int join_caller(DevRes *r) {
  MGB_CUDA_CHECK(cudaEventRecord(r->ev_entry, cudaStreamLegacy));
  MGB_CUDA_CHECK(cudaStreamWaitEvent(r->s_in, r->ev_entry, 0));
  MGB_CUDA_CHECK(cudaStreamWaitEvent(r->s_comp, r->ev_entry, 0));
  MGB_CUDA_CHECK(cudaStreamWaitEvent(r->s_out, r->ev_entry, 0));
  return MGB_SUCCESS;
}

// owned (non-uniform) plans are destroyed on every path
struct PlanGuard {
  mgb_plan *p = nullptr;
  bool owned = false;
  ~PlanGuard() {
    if (owned && p)
      mgb_plan_destroy(p);
  }
};

uint64_t subdomain_elems(const Partition &pt, int ndim, const uint64_t *shape, uint64_t id) {
  uint64_t sub[MGB_MAX_DIMS];
  subdomain_shape(pt, ndim, shape, id, sub);
  uint64_t n = 1;
  for (int d = 0; d < ndim; d++)
    n *= sub[d];
  return n;
}

// One call of the high-level compressor on this process: sub-domains
// [first, first + count) of the partition `pt` of the domain `shape`.
struct CompressJob {
  int ndim = 0, dtype = 0;
  const uint64_t *shape = nullptr; // whole domain
  Partition pt;
  uint64_t first = 0, count = 1;
  // input: the whole array (local == false; any partition, single process) or this
  // process's sub-domains back to back (local == true; partition along dim 0)
  const void *in = nullptr;
  bool local = false;
  const void *const *coords = nullptr;
  const mgb_config *cfg = nullptr;
  double tol = 0, s = 0;
  int ebtype = MGB_ABS;
  bool norm_given = false; // norm holds the global norm (mgb_compress_subdomains)
  double norm = 1;         // out: norm of the original data (relative bounds)
  unsigned char *out = nullptr;
  bool out_dev = false;
  uint64_t cap = 0;
  uint64_t offset = 0; // in: where the first record goes; out: end of the last one
  mgb_comm *comm = nullptr;
};

// plan of sub-domain `id` (coordinates cut as DomainDecomposer.hpp:283-300 does)
int subdomain_plan(const CompressJob &j, uint64_t id, PlanGuard &g, uint64_t *sub) {
  subdomain_shape(j.pt, j.ndim, j.shape, id, sub);
  const size_t tsize = j.dtype == MGB_F32 ? 4 : 8;
  const void *subcoords[MGB_MAX_DIMS];
  if (j.coords)
    subdomain_coords(j.pt, j.ndim, j.shape, tsize, id, j.coords, subcoords);
  return get_plan(j.ndim, j.dtype, sub, j.coords ? subcoords : nullptr, j.cfg, &g.p, &g.owned);
}

// where sub-domain `id` starts in the input when it is contiguous there
const unsigned char *subdomain_ptr(const CompressJob &j, uint64_t id, uint64_t plane) {
  if (!j.pt.decomposed)
    return (const unsigned char *)j.in;
  const uint64_t p0 = subdomain_first_plane(j.pt, id) - (j.local ? subdomain_first_plane(j.pt, j.first) : 0);
  return (const unsigned char *)j.in + p0 * plane;
}

// dense copy of sub-domain `id` into device memory `dst` on stream st
int fetch_subdomain(const CompressJob &j, uint64_t id, uint64_t plane, bool contiguous, uint64_t raw_bytes,
                    unsigned char *dst, cudaStream_t st) {
  const size_t tsize = j.dtype == MGB_F32 ? 4 : 8;
  if (contiguous) {
    MGB_CUDA_CHECK(cudaMemcpyAsync(dst, subdomain_ptr(j, id, plane), raw_bytes, cudaMemcpyDefault, st));
    return MGB_SUCCESS;
  }
  return copy_subdomain(j.pt, j.ndim, j.shape, tsize, id, j.in, dst, true, st);
}

int compress_core(CompressJob &j) {
  DevRes *r = nullptr;
  int rc = dev_res(&r);
  if (rc)
    return rc;
  rc = join_caller(r);
  if (rc)
    return rc;
  const mgb_config *cfg = j.cfg;
  const size_t tsize = j.dtype == MGB_F32 ? 4 : 8;
  const bool in_dev = is_device_pointer(j.in);
  const bool contiguous = j.pt.slabs();
  if (j.local && !contiguous)
    return MGB_BAD_ARGUMENT;
  uint64_t plane = tsize; // bytes of one index along dim 0
  for (int d = 1; d < j.ndim; d++)
    plane *= j.shape[d];
  uint64_t n_total = 1;
  for (int d = 0; d < j.ndim; d++)
    n_total *= j.shape[d];
  const bool direct_in = in_dev && contiguous; // sub-domains are read where they lie
  cudaStream_t sc = r->s_comp;

  // ---- error control of a decomposed domain (CompressionHighLevel.hpp:128-139) -------
  // relative bounds need the norm of the WHOLE domain before any sub-domain is
  // quantized: {max |x|, sum x^2} per sub-domain (two-stage reduction each), summed
  // across processes (x + 0 is exact, so every process ends up with every pair), then
  // combined in sub-domain order - the result does not depend on the number of
  // processes.  Nothing of this visits the host; the quantizer tables are made on the
  // device from the reduced pair (quantize.cu: prepare_q_kernel).
  const bool device_norm = j.pt.decomposed && j.ebtype == MGB_REL && !j.norm_given;
  double ltol = j.tol;
  int leb = j.ebtype;
  if (j.pt.decomposed && !device_norm) {
    ltol = local_abs_tol(j.dtype, j.ebtype, j.norm, j.tol, j.s, j.pt.count);
    leb = MGB_ABS;
  }
  uint64_t local_bytes = 0, max_raw = 0;
  for (uint64_t id = j.first; id < j.first + j.count; id++) {
    const uint64_t b = subdomain_elems(j.pt, j.ndim, j.shape, id) * tsize;
    local_bytes += b;
    max_raw = std::max(max_raw, b);
  }
  bool staged_full = false; // the local block sits dense in r->d_full
  if (device_norm) {
    if (r->pairs_cap < j.pt.count) {
      cudaFree(r->d_pairs);
      r->d_pairs = nullptr;
      r->pairs_cap = 0;
      MGB_CUDA_CHECK(cudaMalloc(&r->d_pairs, 2 * j.pt.count * sizeof(double)));
      r->pairs_cap = j.pt.count;
    }
    MGB_CUDA_CHECK(cudaMemsetAsync(r->d_pairs, 0, 2 * j.pt.count * sizeof(double), sc));
    double *d_part = r->d_red + 4;
    if (direct_in) {
      for (uint64_t id = j.first; id < j.first + j.count; id++) {
        rc = mgb_norm_raw(j.dtype, subdomain_ptr(j, id, plane), subdomain_elems(j.pt, j.ndim, j.shape, id), d_part,
                          r->d_pairs + 2 * id, sc);
        if (rc)
          return rc;
      }
    } else {
      // host (or strided) input: one pass over the link if the block fits next to the
      // workspaces, else the norm pass re-reads it as the reference's
      // calc_norm_decomposed_w_prefetch does (ErrorToleranceCalculator.hpp:91-132)
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      staged_full = local_bytes <= r->full_bytes ||
                    (double)local_bytes + 8.0 * (double)max_raw + (double)(1ull << 30) < (double)(free_b + r->full_bytes);
      if (staged_full) {
        rc = ensure_bytes(&r->d_full, &r->full_bytes, local_bytes);
        if (rc)
          return rc;
      } else {
        for (int k = 0; k < 2; k++) {
          rc = ensure_bytes(&r->stage[k], &r->stage_bytes[k], max_raw);
          if (rc)
            return rc;
        }
      }
      uint64_t off = 0, k = 0;
      for (uint64_t id = j.first; id < j.first + j.count; id++, k++) {
        const uint64_t nsub = subdomain_elems(j.pt, j.ndim, j.shape, id);
        unsigned char *dst = staged_full ? r->d_full + off : r->stage[k & 1];
        if (!staged_full && k >= 2)
          MGB_CUDA_CHECK(cudaStreamWaitEvent(r->s_in, r->ev_comp[k & 1], 0));
        rc = fetch_subdomain(j, id, plane, contiguous, nsub * tsize, dst, r->s_in);
        if (rc)
          return rc;
        MGB_CUDA_CHECK(cudaEventRecord(r->ev_in[k & 1], r->s_in));
        MGB_CUDA_CHECK(cudaStreamWaitEvent(sc, r->ev_in[k & 1], 0));
        rc = mgb_norm_raw(j.dtype, dst, nsub, d_part, r->d_pairs + 2 * id, sc);
        if (rc)
          return rc;
        MGB_CUDA_CHECK(cudaEventRecord(r->ev_comp[k & 1], sc));
        off += nsub * tsize;
      }
      if (!staged_full) { // the slots are reused by the compression pass below
        MGB_CUDA_CHECK(cudaStreamWaitEvent(r->s_in, r->ev_comp[0], 0));
        MGB_CUDA_CHECK(cudaStreamWaitEvent(r->s_in, r->ev_comp[1], 0));
      }
    }
    if (j.comm && j.comm->nranks > 1) {
      const NcclApi &nc = nccl_api();
      if (!nc.ok || !j.comm->nccl)
        return MGB_FAILURE;
      MGB_NCCL_CHECK(nc.AllReduce(r->d_pairs, r->d_pairs, 2 * j.pt.count, NCCL_FLOAT64, NCCL_SUM, j.comm->nccl, sc));
    }
    rc = mgb_norm_combine(r->d_pairs, (int)j.pt.count, r->d_red, sc);
    if (rc)
      return rc;
  }

  // ---- records: H2D of k+1 | compute of k | D2H of k-1 ------------------------------
  const bool stage_in = !(direct_in || staged_full);
  if (stage_in)
    for (int k = 0; k < 2; k++) {
      rc = ensure_bytes(&r->stage[k], &r->stage_bytes[k], max_raw);
      if (rc)
        return rc;
    }
  auto issue_fetch = [&](uint64_t k) -> int {
    const uint64_t id = j.first + k;
    const uint64_t nsub = subdomain_elems(j.pt, j.ndim, j.shape, id);
    int e = fetch_subdomain(j, id, plane, contiguous, nsub * tsize, r->stage[k & 1], r->s_in);
    if (e)
      return e;
    MGB_CUDA_CHECK(cudaEventRecord(r->ev_in[k & 1], r->s_in));
    return MGB_SUCCESS;
  };
  if (stage_in) {
    rc = issue_fetch(0);
    if (rc)
      return rc;
  }
  uint64_t full_off = 0;
  for (uint64_t k = 0; k < j.count; k++) {
    const uint64_t id = j.first + k;
    uint64_t sub[MGB_MAX_DIMS];
    PlanGuard g;
    rc = subdomain_plan(j, id, g, sub);
    if (rc)
      return rc;
    mgb_plan *plan = g.p;
    const uint64_t nsub = plan->N, raw_bytes = nsub * tsize;
    const int slot = (int)(k & 1);
    // slot (k+1)&1 was last read by the compression of k-1, which the host has waited for
    if (stage_in && k + 1 < j.count) {
      rc = issue_fetch(k + 1);
      if (rc)
        return rc;
    }
    const void *d_in;
    if (direct_in)
      d_in = subdomain_ptr(j, id, plane);
    else if (staged_full)
      d_in = r->d_full + full_off;
    else {
      d_in = r->stage[slot];
      MGB_CUDA_CHECK(cudaStreamWaitEvent(sc, r->ev_in[slot], 0));
    }
    full_off += raw_bytes;
    const void *qtab = nullptr;
    if (device_norm) {
      rc = ensure_lowlevel_workspace(plan);
      if (rc)
        return rc;
      rc = mgb_prepare_quantizers(plan, MGB_REL, j.tol, j.s, 1, r->d_red, n_total, j.pt.count, plan->d_qtab,
                                  r->d_red + 2, sc);
      if (rc)
        return rc;
      qtab = plan->d_qtab;
    }
    uint64_t pcap = raw_bytes + 2 * (1024 + 8ull * cfg->huff_dict_size) +
                    32 * ((nsub - 1) / cfg->huff_block_size + 1) + 4096;
    // device output whose payload position is 8-byte aligned: compress straight into
    // the record; otherwise into a slot that the drain stream copies out
    unsigned char *direct = nullptr;
    if (cfg->lossless != 2 && j.out_dev && j.offset + 8 <= j.cap && (((uintptr_t)(j.out + j.offset + 8)) & 7) == 0)
      direct = j.out + j.offset + 8;
    if (!direct) {
      rc = ensure_bytes(&r->payload[slot], &r->payload_bytes[slot], pcap);
      if (rc)
        return rc;
      MGB_CUDA_CHECK(cudaStreamWaitEvent(sc, r->ev_out[slot], 0)); // drained (record k-2)
    }
    uint64_t psize = 0;
    double nrm = j.norm;
    rc = compress_lowlevel_sync(plan, d_in, leb, ltol, j.s, qtab, &nrm, direct ? direct : r->payload[slot],
                                direct ? std::min<uint64_t>(pcap, j.cap - j.offset - 8) : pcap, &psize, sc);
    if (!j.pt.decomposed && j.ebtype == MGB_REL)
      j.norm = nrm;
    const void *payload = direct ? direct : r->payload[slot];
    std::vector<unsigned char> zbuf; // host: size_t count | zstd frame
    if (rc == MGB_SUCCESS && cfg->lossless == 2) {
      const ZstdApi &z = zstd_api();
      if (!z.ok)
        return MGB_FAILURE;
      std::vector<unsigned char> hpay(psize);
      MGB_CUDA_CHECK(cudaMemcpyAsync(hpay.data(), payload, psize, cudaMemcpyDeviceToHost, sc));
      MGB_CUDA_CHECK(cudaStreamSynchronize(sc));
      zbuf.resize(sizeof(size_t) + z.bound(psize));
      const size_t zs = z.compress(zbuf.data() + sizeof(size_t), zbuf.size() - sizeof(size_t), hpay.data(), psize,
                                   cfg->zstd_compress_level);
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
    if (j.offset + 8 + psize > j.cap)
      return MGB_OUTPUT_TOO_LARGE;
    const bool host_payload = payload == (const void *)zbuf.data() && !zbuf.empty();
    if (j.out_dev) {
      // pageable source: staged by the runtime before the call returns
      const uint64_t sz = psize;
      MGB_CUDA_CHECK(cudaMemcpyAsync(j.out + j.offset, &sz, 8, cudaMemcpyHostToDevice, r->s_out));
    } else {
      memcpy(j.out + j.offset, &psize, 8);
    }
    if (payload != (const void *)(j.out + j.offset + 8)) {
      if (host_payload || payload == d_in) {
        // small or rare: host-resident zstd frame / raw fallback, waited for right away
        // (the source is a local buffer or a staging slot about to be reused)
        MGB_CUDA_CHECK(cudaMemcpyAsync(j.out + j.offset + 8, payload, psize, cudaMemcpyDefault, r->s_out));
        MGB_CUDA_CHECK(cudaStreamSynchronize(r->s_out));
      } else {
        // the compression of k is complete (its size was read back): drain it while
        // the next sub-domain is compressed
        MGB_CUDA_CHECK(cudaMemcpyAsync(j.out + j.offset + 8, payload, psize, cudaMemcpyDefault, r->s_out));
      }
    }
    MGB_CUDA_CHECK(cudaEventRecord(r->ev_out[slot], r->s_out));
    j.offset += 8 + psize;
  }
  if (device_norm) {
    MGB_CUDA_CHECK(cudaMemcpyAsync(&r->h_pin[40], r->d_red + 2, sizeof(double), cudaMemcpyDeviceToHost, sc));
    MGB_CUDA_CHECK(cudaStreamSynchronize(sc));
    memcpy(&j.norm, &r->h_pin[40], sizeof(double));
  }
  MGB_CUDA_CHECK(cudaStreamSynchronize(r->s_out));
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

// contiguous block of sub-domains per process (balanced, rank-major)
void owned_range(uint64_t nsub, int rank, int nranks, uint64_t *first, uint64_t *count) {
  const uint64_t base = nsub / nranks, rem = nsub % nranks;
  *first = rank * base + std::min<uint64_t>(rank, rem);
  *count = base + ((uint64_t)rank < rem ? 1 : 0);
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
  trim_plan_cache();
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
  CompressJob j;
  rc = make_partition(ndim, shape, tsize, &cfg, j.pt);
  if (rc)
    return rc;
  j.ndim = ndim;
  j.dtype = dtype;
  j.shape = shape;
  j.first = 0;
  j.count = j.pt.count;
  j.in = in;
  j.local = false;
  j.coords = coords;
  j.cfg = &cfg;
  j.tol = tol;
  j.s = s;
  j.ebtype = ebtype;
  mgb_header h;
  header_from(ndim, dtype, shape, tol, s, ebtype, 1.0, coords, &cfg, j.pt, h);
  std::vector<uint8_t> hdr = mgb_encode_stream_header(h);
  uint64_t cap;
  unsigned char *obuf;
  if (!output_pre_allocated) {
    // CompressionHighLevel.hpp:149-158 (OUTPUT_SAFTY_OVERHEAD = 1e6)
    cap = N * tsize + 1000000 + hdr.size() + 8 * j.pt.count;
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
  j.out = obuf;
  j.out_dev = is_device_pointer(obuf);
  j.cap = cap;
  j.offset = hdr.size();
  if (j.offset > cap)
    rc = MGB_OUTPUT_TOO_LARGE;
  if (!rc)
    rc = compress_core(j);
  if (!rc) {
    // the norm of a relative bound is known only now
    // (CompressionHighLevel.hpp:253-279 serialises the metadata again)
    header_from(ndim, dtype, shape, tol, s, ebtype, j.norm, coords, &cfg, j.pt, h);
    std::vector<uint8_t> hdr2 = mgb_encode_stream_header(h);
    if (hdr2.size() != hdr.size())
      rc = MGB_FAILURE;
    else if (j.out_dev)
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
  *out_size = j.offset;
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

namespace {

// One call of the high-level decompressor on this process: records of sub-domains
// [first, first + count), the first one starting at in + offset.
struct DecompressJob {
  const mgb_header *h = nullptr;
  Partition pt;
  uint64_t first = 0, count = 1;
  const unsigned char *in = nullptr;
  uint64_t in_size = 0, offset = 0;
  // output: the whole array (local == false) or this process's sub-domains back to
  // back (local == true; partition along dim 0)
  unsigned char *out = nullptr;
  bool local = false;
  const mgb_config *cfg = nullptr;
  const void *const *coords = nullptr; // whole-domain coordinates (T) or nullptr
};

int decompress_core(DecompressJob &j) {
  DevRes *r = nullptr;
  int rc = dev_res(&r);
  if (rc)
    return rc;
  rc = join_caller(r);
  if (rc)
    return rc;
  const mgb_header &h = *j.h;
  const int ndim = h.ndim, dtype = h.dtype;
  const size_t tsize = dtype == MGB_F32 ? 4 : 8;
  const bool in_dev = is_device_pointer(j.in), out_dev = is_device_pointer(j.out);
  const bool contiguous = j.pt.slabs();
  if (j.local && !contiguous)
    return MGB_BAD_ARGUMENT;
  double ltol = h.tol;
  int leb = h.ebtype;
  if (j.pt.decomposed) {
    ltol = local_abs_tol(dtype, h.ebtype, h.norm, h.tol, h.s, j.pt.count);
    leb = MGB_ABS;
  }
  uint64_t plane = tsize;
  for (int d = 1; d < ndim; d++)
    plane *= h.shape[d];
  cudaStream_t sc = r->s_comp;
  // record sizes: the u64 chain is walked up front (8 bytes per record)
  std::vector<uint64_t> rec_off(j.count), rec_size(j.count);
  uint64_t offset = j.offset;
  for (uint64_t k = 0; k < j.count; k++) {
    if (offset + 8 > j.in_size)
      return MGB_BAD_STREAM;
    uint64_t psize = 0;
    if (in_dev)
      MGB_CUDA_CHECK(cudaMemcpy(&psize, j.in + offset, 8, cudaMemcpyDeviceToHost));
    else
      memcpy(&psize, j.in + offset, 8);
    offset += 8;
    if (psize > j.in_size - offset)
      return MGB_BAD_STREAM;
    rec_off[k] = offset;
    rec_size[k] = psize;
    offset += psize;
  }
  // a record is decoded where it lies when it is in device memory and 8-byte aligned
  auto in_place = [&](uint64_t k) { return in_dev && ((uintptr_t)(j.in + rec_off[k]) & 7) == 0 && h.lossless != 2; };
  auto raw_of = [&](uint64_t k) { return subdomain_elems(j.pt, ndim, h.shape, j.first + k) * tsize; };
  auto issue_fetch = [&](uint64_t k) -> int {
    if (in_place(k) || h.lossless == 2 || rec_size[k] >= raw_of(k))
      return MGB_SUCCESS;
    const int slot = (int)(k & 1);
    int e = ensure_bytes(&r->payload[slot], &r->payload_bytes[slot], rec_size[k] + 64);
    if (e)
      return e;
    MGB_CUDA_CHECK(cudaStreamWaitEvent(r->s_in, r->ev_comp[slot], 0)); // decode of k-2 has read the slot
    MGB_CUDA_CHECK(cudaMemcpyAsync(r->payload[slot], j.in + rec_off[k], rec_size[k], cudaMemcpyDefault, r->s_in));
    MGB_CUDA_CHECK(cudaEventRecord(r->ev_in[slot], r->s_in));
    return MGB_SUCCESS;
  };
  if (j.count) {
    rc = issue_fetch(0);
    if (rc)
      return rc;
  }
  for (uint64_t k = 0; k < j.count; k++) {
    const uint64_t id = j.first + k;
    const int slot = (int)(k & 1);
    uint64_t sub[MGB_MAX_DIMS];
    subdomain_shape(j.pt, ndim, h.shape, id, sub);
    const uint64_t raw_bytes = raw_of(k), psize = rec_size[k];
    if (k + 1 < j.count) {
      rc = issue_fetch(k + 1);
      if (rc)
        return rc;
    }
    // dense device destination of this sub-domain
    const uint64_t p0 = subdomain_first_plane(j.pt, id) - (j.local ? subdomain_first_plane(j.pt, j.first) : 0);
    unsigned char *final_dst = j.out + (j.pt.decomposed && contiguous ? p0 * plane : 0);
    unsigned char *d_dst;
    const bool direct_out = out_dev && contiguous;
    if (direct_out) {
      d_dst = final_dst;
    } else {
      rc = ensure_bytes(&r->stage[slot], &r->stage_bytes[slot], raw_bytes);
      if (rc)
        return rc;
      d_dst = r->stage[slot];
      MGB_CUDA_CHECK(cudaStreamWaitEvent(sc, r->ev_out[slot], 0)); // drained (sub-domain k-2)
    }
    if (psize >= raw_bytes) {
      // raw sub-domain (GPUPipelines.hpp:417,458-466)
      MGB_CUDA_CHECK(cudaMemcpyAsync(d_dst, j.in + rec_off[k], raw_bytes, cudaMemcpyDefault, sc));
    } else {
      PlanGuard g;
      const void *subcoords[MGB_MAX_DIMS];
      if (j.coords)
        subdomain_coords(j.pt, ndim, h.shape, tsize, id, j.coords, subcoords);
      rc = get_plan(ndim, dtype, sub, j.coords ? subcoords : nullptr, j.cfg, &g.p, &g.owned);
      if (rc)
        return rc;
      const unsigned char *d_pay;
      uint64_t hsize = psize; // size of the Huffman block
      if (h.lossless == 2) {
        // Zstd.hpp:100-125: `size_t count | zstd frame` -> Huffman block, on the host
        const ZstdApi &z = zstd_api();
        if (!z.ok)
          return MGB_FAILURE;
        if (psize < sizeof(size_t))
          return MGB_BAD_STREAM;
        std::vector<unsigned char> rec(psize), hpay;
        MGB_CUDA_CHECK(cudaMemcpy(rec.data(), j.in + rec_off[k], psize, cudaMemcpyDefault));
        size_t count = 0;
        memcpy(&count, rec.data(), sizeof(size_t));
        if (count > (size_t)raw_bytes * 2 + (1u << 24))
          return MGB_BAD_STREAM;
        hpay.resize(count);
        const size_t got = z.decompress(hpay.data(), count, rec.data() + sizeof(size_t), psize - sizeof(size_t));
        if (z.is_error(got) || got != count)
          return MGB_BAD_STREAM;
        hsize = count;
        MGB_CUDA_CHECK(cudaStreamSynchronize(sc)); // the slot may still be read by the previous decode
        rc = ensure_bytes(&r->payload[slot], &r->payload_bytes[slot], hsize + 64);
        if (rc)
          return rc;
        MGB_CUDA_CHECK(cudaMemcpy(r->payload[slot], hpay.data(), hsize, cudaMemcpyHostToDevice));
        d_pay = r->payload[slot];
      } else if (in_place(k)) {
        d_pay = j.in + rec_off[k];
      } else {
        MGB_CUDA_CHECK(cudaStreamWaitEvent(sc, r->ev_in[slot], 0));
        d_pay = r->payload[slot];
      }
      rc = decompress_lowlevel_async(g.p, d_pay, hsize, leb, ltol, h.s, h.norm, d_dst, sc);
      if (rc)
        return rc;
      if (g.owned) // its workspaces go away with the guard
        MGB_CUDA_CHECK(cudaStreamSynchronize(sc));
    }
    MGB_CUDA_CHECK(cudaEventRecord(r->ev_comp[slot], sc));
    if (!direct_out) {
      MGB_CUDA_CHECK(cudaStreamWaitEvent(r->s_out, r->ev_comp[slot], 0));
      if (contiguous)
        MGB_CUDA_CHECK(cudaMemcpyAsync(final_dst, d_dst, raw_bytes, cudaMemcpyDefault, r->s_out));
      else {
        rc = copy_subdomain(j.pt, ndim, h.shape, tsize, id, j.out, d_dst, false, r->s_out);
        if (rc)
          return rc;
      }
      MGB_CUDA_CHECK(cudaEventRecord(r->ev_out[slot], r->s_out));
    }
  }
  MGB_CUDA_CHECK(cudaStreamSynchronize(sc));
  MGB_CUDA_CHECK(cudaStreamSynchronize(r->s_out));
  return MGB_SUCCESS;
}

// header of a stream in host or device memory
int read_header(const void *in, size_t in_size, mgb_header &h, uint64_t &hb) {
  std::vector<uint8_t> head;
  const uint8_t *hp = (const uint8_t *)in;
  size_t hsize = in_size;
  if (is_device_pointer(in)) {
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
  return mgb_parse_stream_header(hp, hsize, h, hb);
}

// what the header fixes for the decoder: configuration, partition, coordinates
struct DecodeSetup {
  mgb_config cfg;
  Partition pt;
  std::vector<std::vector<unsigned char>> cbytes;
  const void *cptr[MGB_MAX_DIMS];
  bool nonuniform = false;
  uint64_t N = 0;
};
int decode_setup(const mgb_header &h, const mgb_config *cfg_in, DecodeSetup &d) {
  if (h.convention != 0)
    return MGB_BAD_STREAM; // MGARD-CPU stream: mgb_cpu_decompress reads those
  mgb_config_default(&d.cfg);
  if (cfg_in)
    d.cfg.dev_id = cfg_in->dev_id;
  // Metadata.cpp:129-136: the header overrides the configuration
  d.cfg.huff_dict_size = h.dict_size;
  d.cfg.huff_block_size = h.block_size;
  d.cfg.lossless = h.lossless;
  d.cfg.reorder = h.reorder;
  d.cfg.decomposition = h.decomposition;
  const int ndim = h.ndim;
  const size_t tsize = h.dtype == MGB_F32 ? 4 : 8;
  for (int k = 0; k < ndim; k++)
    if (h.shape[k] < 3)
      return MGB_BAD_STREAM;
  // a header that passes its CRC can still announce an absurd shape
  if (!mgb_checked_elems(ndim, h.shape, tsize, &d.N))
    return MGB_BAD_STREAM;
  // not in the stream: the caller repeats them (as with the reference)
  if (cfg_in)
    d.cfg.max_larget_level = cfg_in->max_larget_level;
  d.pt.decomposed = h.decomposed;
  d.pt.method = h.dd_method;
  d.pt.dim = (int)h.dd_dim;
  d.pt.size = h.dd_size;
  d.pt.count = 1;
  if (d.pt.decomposed && d.pt.method == 3) {
    // Variable: the extents come from the caller's Config (CompressionHighLevel.hpp:485-493)
    if (!cfg_in || !cfg_in->domain_decomposition_sizes || !cfg_in->num_domain_decomposition_sizes ||
        d.pt.dim >= ndim)
      return MGB_BAD_ARGUMENT;
    uint64_t sum = 0;
    for (uint64_t k = 0; k < cfg_in->num_domain_decomposition_sizes; k++) {
      const uint64_t e = cfg_in->domain_decomposition_sizes[k];
      if (e < 3 || e > h.shape[d.pt.dim])
        return MGB_BAD_ARGUMENT;
      d.pt.var_offsets.push_back(sum);
      d.pt.var_sizes.push_back(e);
      sum += e;
    }
    if (sum != h.shape[d.pt.dim])
      return MGB_BAD_ARGUMENT;
    d.pt.count = d.pt.var_sizes.size();
    d.pt.cuts[d.pt.dim] = d.pt.count;
  } else if (d.pt.decomposed && d.pt.method == 2) {
    if (d.pt.size < 3)
      return MGB_BAD_STREAM;
    d.pt.dim = 0;
    for (int k = 0; k < ndim; k++) {
      d.pt.cuts[k] = (h.shape[k] - 1) / d.pt.size + 1;
      const uint64_t left = h.shape[k] % d.pt.size;
      if ((left != 0 && left < 3) || d.pt.count > (1ull << 40) / d.pt.cuts[k])
        return MGB_BAD_STREAM;
      d.pt.count *= d.pt.cuts[k];
    }
  } else if (d.pt.decomposed) {
    if (d.pt.dim >= ndim || d.pt.size < 3 || d.pt.size >= h.shape[d.pt.dim])
      return MGB_BAD_STREAM;
    d.pt.count = (h.shape[d.pt.dim] - 1) / d.pt.size + 1;
    d.pt.cuts[d.pt.dim] = d.pt.count;
  }
  // CompressionHighLevel.hpp:456-464: coordinates go through float
  d.nonuniform = !h.coords.empty();
  if (d.nonuniform) {
    d.cbytes.resize(ndim);
    for (int k = 0; k < ndim; k++) {
      d.cbytes[k].resize(h.shape[k] * tsize);
      for (uint64_t i = 0; i < h.shape[k]; i++) {
        float f = (float)h.coords[k][i];
        if (h.dtype == MGB_F32)
          ((float *)d.cbytes[k].data())[i] = f;
        else
          ((double *)d.cbytes[k].data())[i] = f;
      }
      d.cptr[k] = d.cbytes[k].data();
    }
  }
  return MGB_SUCCESS;
}

} // namespace

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
  trim_plan_cache();
  const bool in_dev = is_device_pointer(in);
  if (in_dev) {
    cudaPointerAttributes a;
    cudaPointerGetAttributes(&a, in);
    cudaSetDevice(a.device);
  } else if (cfg_in && cfg_in->dev_id >= 0) {
    cudaSetDevice(cfg_in->dev_id);
  }
  mgb_header h;
  uint64_t hb = 0;
  int rc = read_header(in, in_size, h, hb);
  if (rc)
    return rc;
  DecodeSetup ds;
  rc = decode_setup(h, cfg_in, ds);
  if (rc)
    return rc;
  const size_t tsize = h.dtype == MGB_F32 ? 4 : 8;
  unsigned char *obuf;
  if (!output_pre_allocated) {
    if (in_dev) {
      void *p = nullptr;
      MGB_CUDA_CHECK(cudaMalloc(&p, ds.N * tsize));
      obuf = (unsigned char *)p;
    } else {
      obuf = (unsigned char *)malloc(ds.N * tsize);
      if (!obuf)
        return MGB_FAILURE;
    }
  } else {
    obuf = (unsigned char *)*out;
  }
  DecompressJob j;
  j.h = &h;
  j.pt = ds.pt;
  j.first = 0;
  j.count = ds.pt.count;
  j.in = (const unsigned char *)in;
  j.in_size = in_size;
  j.offset = hb;
  j.out = obuf;
  j.local = false;
  j.cfg = &ds.cfg;
  j.coords = ds.nonuniform ? ds.cptr : nullptr;
  rc = decompress_core(j);
  if (rc) {
    cudaDeviceSynchronize();
    if (!output_pre_allocated) {
      if (in_dev)
        cudaFree(obuf);
      else
        free(obuf);
    }
    return rc;
  }
  *out = obuf;
  if (ndim_out) *ndim_out = h.ndim;
  if (dtype_out) *dtype_out = h.dtype;
  if (shape_out)
    for (int d = 0; d < h.ndim; d++)
      shape_out[d] = h.shape[d];
  return MGB_SUCCESS;
}

extern "C" void mgb_release_cache(void) {
  std::lock_guard<std::mutex> lock(g_cache.mu);
  cudaDeviceSynchronize();
  for (auto &kv : g_cache.plans)
    mgb_plan_destroy(kv.second);
  g_cache.plans.clear();
  g_cache.last_use.clear();
  for (auto &kv : g_devres) {
    DevRes &r = kv.second;
    for (int k = 0; k < 2; k++) {
      cudaFree(r.stage[k]);
      cudaFree(r.payload[k]);
      r.stage[k] = r.payload[k] = nullptr;
      r.stage_bytes[k] = r.payload_bytes[k] = 0;
    }
    cudaFree(r.d_full);
    r.d_full = nullptr;
    r.full_bytes = 0;
  }
}

// ---- one process per GPU ------------------------------------------------------------
// The MaxDim partition along dim 0 (DomainDecomposer.hpp:124-169) spread over the
// processes of a communicator: process r owns a contiguous block of sub-domains
// (owned_range), `local` holds them back to back.  Exchanges: one all-reduce of the
// {max |x|, sum x^2} pairs behind relative bounds, stream-ordered on the device (the
// norm never visits the host before the quantizer has used it), and one all-gather
// of the container sizes (GPUPipelines.hpp:189-193: every record is `u64 size |
// payload`, so a process needs the bytes written before it).  The records are the
// ones mgb_compress writes for the same domain, whatever the number of processes.
static int compress_sharded_impl(mgb_comm *comm, int ndim, int dtype, const uint64_t *shape, double tol, double s,
                                 int ebtype, const void *local, const mgb_config *cfg_in, void *out, uint64_t cap,
                                 uint64_t *local_size, uint64_t *offset, uint64_t *total_size, uint64_t *all_sizes,
                                 double *norm, uint8_t *header, uint64_t header_cap, uint64_t *header_size) {
  int rc = check_args(ndim, dtype, shape);
  if (rc)
    return rc;
  if (!local || !out || !local_size || !cfg_in)
    return MGB_BAD_ARGUMENT;
  std::lock_guard<std::mutex> lock(g_cache.mu);
  trim_plan_cache();
  mgb_config cfg = *cfg_in;
  if (cfg.lossless != 0 && cfg.lossless != 2)
    return MGB_FAILURE;
  if (is_device_pointer(local)) {
    cudaPointerAttributes a;
    cudaPointerGetAttributes(&a, local);
    cudaSetDevice(a.device);
  } else if (cfg.dev_id >= 0) {
    cudaSetDevice(cfg.dev_id);
  }
  const size_t tsize = dtype == MGB_F32 ? 4 : 8;
  uint64_t nall = 0;
  if (!mgb_checked_elems(ndim, shape, tsize, &nall))
    return MGB_BAD_ARGUMENT;
  CompressJob j;
  cfg.domain_decomposition_dim = 0;
  if (cfg.domain_decomposition == 1)
    return MGB_BAD_ARGUMENT; // Block sub-domains are not runs of planes
  rc = make_partition(ndim, shape, tsize, &cfg, j.pt);
  if (rc)
    return rc;
  const int rank = comm ? comm->rank : 0, nranks = comm ? comm->nranks : 1;
  if (!j.pt.decomposed && nranks > 1)
    return MGB_BAD_ARGUMENT;
  j.ndim = ndim;
  j.dtype = dtype;
  j.shape = shape;
  owned_range(j.pt.count, rank, nranks, &j.first, &j.count);
  j.in = local;
  j.local = true;
  j.cfg = &cfg;
  j.tol = tol;
  j.s = s;
  j.ebtype = ebtype;
  j.out = (unsigned char *)out;
  j.out_dev = is_device_pointer(out);
  j.cap = cap;
  j.offset = 0;
  j.comm = comm;
  rc = compress_core(j);
  if (rc)
    return rc;
  *local_size = j.offset;
  mgb_header h;
  header_from(ndim, dtype, shape, tol, s, ebtype, j.norm, nullptr, &cfg, j.pt, h);
  std::vector<uint8_t> hdr = mgb_encode_stream_header(h);
  if (header_size)
    *header_size = hdr.size();
  if (header) {
    if (hdr.size() > header_cap)
      return MGB_OUTPUT_TOO_LARGE;
    memcpy(header, hdr.data(), hdr.size());
  }
  if (norm)
    *norm = j.norm;
  // sizes of all containers -> my byte offset in the stream
  std::vector<uint64_t> sizes(nranks, 0);
  sizes[rank] = j.offset;
  if (nranks > 1) {
    DevRes *r = nullptr;
    rc = dev_res(&r);
    if (rc)
      return rc;
    const NcclApi &nc = nccl_api();
    if (!nc.ok || !comm->nccl)
      return MGB_FAILURE;
    if (r->sizes_cap < (uint64_t)nranks) {
      cudaFree(r->d_sizes);
      r->d_sizes = nullptr;
      MGB_CUDA_CHECK(cudaMalloc(&r->d_sizes, (8 + nranks) * sizeof(unsigned long long)));
      r->sizes_cap = nranks;
    }
    r->h_pin[0] = j.offset;
    MGB_CUDA_CHECK(cudaMemcpyAsync(r->d_sizes, r->h_pin, 8, cudaMemcpyHostToDevice, r->s_comp));
    MGB_NCCL_CHECK(nc.AllGather(r->d_sizes, r->d_sizes + 8, 1, NCCL_UINT64, comm->nccl, r->s_comp));
    if (nranks > 48)
      return MGB_FAILURE;
    MGB_CUDA_CHECK(cudaMemcpyAsync(r->h_pin + 8, r->d_sizes + 8, nranks * 8, cudaMemcpyDeviceToHost, r->s_comp));
    MGB_CUDA_CHECK(cudaStreamSynchronize(r->s_comp));
    for (int k = 0; k < nranks; k++)
      sizes[k] = r->h_pin[8 + k];
  }
  uint64_t off = hdr.size(), tot = hdr.size();
  for (int k = 0; k < nranks; k++) {
    if (k < rank)
      off += sizes[k];
    tot += sizes[k];
    if (all_sizes)
      all_sizes[k] = sizes[k];
  }
  if (offset)
    *offset = off;
  if (total_size)
    *total_size = tot;
  return MGB_SUCCESS;
}

// Inverse: `header` = the stream's preamble + metadata (host), `records` = this
// process's records (host or device), `local_out` = its sub-domains back to back.
// No exchange is needed: every process decodes what it owns.
static int decompress_sharded_impl(mgb_comm *comm, const uint8_t *header, uint64_t header_size, const void *records,
                                   uint64_t records_size, void *local_out, const mgb_config *cfg_in) {
  if (!header || !records || !local_out)
    return MGB_BAD_ARGUMENT;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return MGB_BACKEND_NOT_AVAILABLE;
  }
  std::lock_guard<std::mutex> lock(g_cache.mu);
  trim_plan_cache();
  if (is_device_pointer(local_out)) {
    cudaPointerAttributes a;
    cudaPointerGetAttributes(&a, local_out);
    cudaSetDevice(a.device);
  } else if (cfg_in && cfg_in->dev_id >= 0) {
    cudaSetDevice(cfg_in->dev_id);
  }
  mgb_header h;
  uint64_t hb = 0;
  int rc = mgb_parse_stream_header(header, header_size, h, hb);
  if (rc)
    return rc;
  DecodeSetup ds;
  rc = decode_setup(h, cfg_in, ds);
  if (rc)
    return rc;
  const int rank = comm ? comm->rank : 0, nranks = comm ? comm->nranks : 1;
  if (nranks > 1 && (!ds.pt.decomposed || !ds.pt.slabs()))
    return MGB_BAD_ARGUMENT;
  DecompressJob j;
  j.h = &h;
  j.pt = ds.pt;
  owned_range(ds.pt.count, rank, nranks, &j.first, &j.count);
  j.in = (const unsigned char *)records;
  j.in_size = records_size;
  j.offset = 0;
  j.out = (unsigned char *)local_out;
  j.local = true;
  j.cfg = &ds.cfg;
  j.coords = ds.nonuniform ? ds.cptr : nullptr;
  rc = decompress_core(j);
  if (rc)
    cudaDeviceSynchronize();
  return rc;
}

// offsets / sizes of the records of a stream in host memory (walks the u64 chain)
static int stream_records_impl(const void *stream, uint64_t size, uint64_t *offsets, uint64_t *sizes, uint64_t cap,
                               uint64_t *count) {
  if (!stream || !count)
    return MGB_BAD_ARGUMENT;
  mgb_header h;
  uint64_t hb = 0;
  int rc = read_header(stream, size, h, hb);
  if (rc)
    return rc;
  DecodeSetup ds;
  rc = decode_setup(h, nullptr, ds);
  if (rc)
    return rc;
  *count = ds.pt.count;
  const bool dev = is_device_pointer(stream);
  uint64_t off = hb;
  for (uint64_t k = 0; k < ds.pt.count; k++) {
    if (off + 8 > size)
      return MGB_BAD_STREAM;
    uint64_t ps = 0;
    if (dev)
      MGB_CUDA_CHECK(cudaMemcpy(&ps, (const unsigned char *)stream + off, 8, cudaMemcpyDeviceToHost));
    else
      memcpy(&ps, (const unsigned char *)stream + off, 8);
    if (ps > size - off - 8)
      return MGB_BAD_STREAM;
    if (k < cap) {
      if (offsets)
        offsets[k] = off;
      if (sizes)
        sizes[k] = 8 + ps;
    }
    off += 8 + ps;
  }
  return MGB_SUCCESS;
}

static int comm_unique_id_impl(uint8_t *id128) {
  const NcclApi &nc = nccl_api();
  if (!nc.ok || !id128)
    return nc.ok ? MGB_BAD_ARGUMENT : MGB_BACKEND_NOT_AVAILABLE;
  mgb_nccl_uid u;
  MGB_NCCL_CHECK(nc.GetUniqueId(&u));
  memcpy(id128, u.internal, 128);
  return MGB_SUCCESS;
}
static int comm_init_rank_impl(const uint8_t *id128, int nranks, int rank, mgb_comm **out) {
  if (!out || nranks < 1 || rank < 0 || rank >= nranks)
    return MGB_BAD_ARGUMENT;
  mgb_comm *c = new mgb_comm();
  c->rank = rank;
  c->nranks = nranks;
  if (nranks > 1) {
    const NcclApi &nc = nccl_api();
    if (!nc.ok || !id128) {
      delete c;
      return nc.ok ? MGB_BAD_ARGUMENT : MGB_BACKEND_NOT_AVAILABLE;
    }
    mgb_nccl_uid u;
    memcpy(u.internal, id128, 128);
    int e = nc.CommInitRank(&c->nccl, nranks, u, rank);
    if (e != 0) {
      fprintf(stderr, "mgard_b200: ncclCommInitRank failed: %s\n", nc.GetErrorString ? nc.GetErrorString(e) : "?");
      delete c;
      return MGB_FAILURE;
    }
    c->owned = true;
  }
  *out = c;
  return MGB_SUCCESS;
}

extern "C" int mgb_comm_unique_id(uint8_t *id128) { MGB_NOEXCEPT_CALL(comm_unique_id_impl(id128)); }
extern "C" int mgb_comm_init_rank(const uint8_t *id128, int nranks, int rank, mgb_comm **comm) {
  MGB_NOEXCEPT_CALL(comm_init_rank_impl(id128, nranks, rank, comm));
}
extern "C" int mgb_comm_from_nccl(void *nccl_comm, int nranks, int rank, mgb_comm **comm) {
  if (!comm || !nccl_comm || nranks < 1 || rank < 0 || rank >= nranks)
    return MGB_BAD_ARGUMENT;
  mgb_comm *c = new (std::nothrow) mgb_comm();
  if (!c)
    return MGB_FAILURE;
  c->nccl = nccl_comm;
  c->rank = rank;
  c->nranks = nranks;
  c->owned = false;
  *comm = c;
  return MGB_SUCCESS;
}
extern "C" void mgb_comm_destroy(mgb_comm *comm) {
  if (!comm)
    return;
  if (comm->owned && comm->nccl && nccl_api().ok)
    nccl_api().CommDestroy(comm->nccl);
  delete comm;
}
extern "C" int mgb_comm_rank(const mgb_comm *comm) { return comm ? comm->rank : 0; }
extern "C" int mgb_comm_size(const mgb_comm *comm) { return comm ? comm->nranks : 1; }
extern "C" int mgb_owned_subdomains(const mgb_comm *comm, uint64_t num_subdomains, uint64_t *first, uint64_t *count) {
  if (!first || !count)
    return MGB_BAD_ARGUMENT;
  owned_range(num_subdomains, comm ? comm->rank : 0, comm ? comm->nranks : 1, first, count);
  return MGB_SUCCESS;
}
extern "C" int mgb_compress_sharded(mgb_comm *comm, int ndim, int dtype, const uint64_t *shape, double tol, double s,
                                    int ebtype, const void *local, const mgb_config *cfg, void *out, uint64_t cap,
                                    uint64_t *local_size, uint64_t *offset, uint64_t *total_size,
                                    uint64_t *all_sizes, double *norm, uint8_t *header, uint64_t header_cap,
                                    uint64_t *header_size) {
  MGB_NOEXCEPT_CALL(compress_sharded_impl(comm, ndim, dtype, shape, tol, s, ebtype, local, cfg, out, cap, local_size,
                                          offset, total_size, all_sizes, norm, header, header_cap, header_size));
}
extern "C" int mgb_decompress_sharded(mgb_comm *comm, const uint8_t *header, uint64_t header_size,
                                      const void *records, uint64_t records_size, void *local_out,
                                      const mgb_config *cfg) {
  MGB_NOEXCEPT_CALL(decompress_sharded_impl(comm, header, header_size, records, records_size, local_out, cfg));
}
extern "C" int mgb_stream_records(const void *stream, uint64_t size, uint64_t *offsets, uint64_t *sizes,
                                  uint64_t cap, uint64_t *count) {
  MGB_NOEXCEPT_CALL(stream_records_impl(stream, size, offsets, sizes, cap, count));
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
  trim_plan_cache();
  const size_t tsize = dtype == MGB_F32 ? 4 : 8;
  uint64_t nall = 0;
  if (!mgb_checked_elems(ndim, shape, tsize, &nall))
    return MGB_BAD_ARGUMENT;
  CompressJob j;
  rc = make_partition(ndim, shape, tsize, cfg_in, j.pt);
  if (rc)
    return rc;
  if (!j.pt.decomposed || !j.pt.slabs() || first + count > j.pt.count)
    return MGB_BAD_ARGUMENT;
  j.ndim = ndim;
  j.dtype = dtype;
  j.shape = shape;
  j.first = first;
  j.count = count;
  j.in = d_in_first;
  j.local = true;
  j.cfg = cfg_in;
  j.tol = tol;
  j.s = s;
  j.ebtype = ebtype;
  j.norm_given = true;
  j.norm = norm;
  j.out = d_out;
  j.out_dev = true;
  j.cap = cap;
  j.offset = 0;
  rc = compress_core(j);
  if (rc)
    return rc;
  *size = j.offset;
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
