// Internal: plan = hierarchy tables (host + device) and device workspaces.
// Host bookkeeping mirrors mgard_x::Hierarchy<D,T>
// (reference include/mgard-x/Hierarchy/Hierarchy.hpp:23-418), evaluated in the
// working precision T with the reference's operation order so that every table
// entry is bit-identical.
#pragma once
#include <cstdint>
#include <cstdio>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/mgard_b200.h"

#define MGB_MAX_LEVELS 48

#define MGB_CUDA_CHECK(x)                                                      \
  do {                                                                         \
    cudaError_t e_ = (x);                                                      \
    if (e_ != cudaSuccess) {                                                   \
      fprintf(stderr, "mgard_b200: CUDA error %s at %s:%d\n",                  \
              cudaGetErrorString(e_), __FILE__, __LINE__);                     \
      return MGB_CUDA_ERROR;                                                   \
    }                                                                          \
  } while (0)

struct mgb_dim_tables {
  // offsets (in elements of T) into the flat table buffer
  uint64_t dist, ratio, am, bm, fw; // fw[i] = am[i] / bm[i] (Thomas forward)
  uint64_t mt; // [9][nc] mass_trans coefficients for level l -> l-1 (l >= 1)
  uint64_t n;
};

struct mgb_plan {
  int D = 0;
  int dtype = MGB_F32;
  size_t tsize = 4;
  uint64_t shape[MGB_MAX_DIMS] = {1, 1, 1, 1, 1};
  int L = 0; // l_target
  uint64_t lshape[MGB_MAX_LEVELS][MGB_MAX_DIMS];
  uint64_t N = 0;
  bool uniform = true;
  mgb_config cfg;
  bool force_generic = false; // tests: use the dimension-generic kernels for D == 3
  std::vector<std::vector<double>> coords; // as doubles, for the header
  // hierarchy tables
  mgb_dim_tables tab[MGB_MAX_LEVELS][MGB_MAX_DIMS];
  std::vector<unsigned char> h_tables; // T elements
  unsigned char *d_tables = nullptr;
  // level marks for the s-norm quantizer: int32[D][max shape]
  int *d_marks = nullptr;
  uint64_t marks_width = 0;

  // ---- device workspaces (allocated on demand) ----
  unsigned char *d_coef = nullptr; // N * T: decomposed coefficients
  unsigned char *d_cbuf = nullptr; // dense coarse boxes, levels L-1..0
  uint64_t cbuf_off[MGB_MAX_LEVELS]; // element offsets per level
  uint64_t cbuf_elems = 0;
  unsigned char *d_wA = nullptr, *d_wB = nullptr;
  unsigned char *d_sd = nullptr; // N * T scratch of the SingleDim decomposition (on demand)
  uint64_t w_elems = 0;
  uint16_t *d_sym = nullptr;
  uint32_t *d_hist = nullptr;
  unsigned long long *d_codebook = nullptr;
  unsigned long long *d_decodebook = nullptr;
  unsigned long long *d_chunk_bits = nullptr;  // nchunk
  unsigned long long *d_chunk_woff = nullptr;  // nchunk + 1
  unsigned *d_chunk_sub = nullptr;             // 8 per chunk: bits of every eighth (huffman_serial.cuh)
  unsigned long long *d_scalars = nullptr; // [0] outlier count, [1] total words, ...
  uint64_t *d_oidx = nullptr;
  int64_t *d_oval = nullptr;
  uint64_t outlier_cap = 0;
  unsigned char *d_norm_tmp = nullptr; // reduction partials (doubles)
  unsigned char *d_cbwork = nullptr;   // codebook kernel scratch
  unsigned *d_declut = nullptr;        // chunk-serial decoder tables (huffman_serial.cuh)
  unsigned *d_dec_sub = nullptr;       // decoder sub-sequence bookkeeping
  uint64_t dec_sub_cap = 0;
  unsigned long long *h_pinned = nullptr; // pinned host scalars
  // side stream + events: work that does not depend on the level recursion runs
  // next to the latency-bound small levels (refactor.cu: recompose_t)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // compression, s = inf: the upper half of the coefficient array (slowest index >=
  // coarse size) holds finest-level coefficients only.  It is quantized on a second
  // side stream in the shadow of the latency-bound coarse levels; armed by
  // mgb_compress_lowlevel, consumed by decompose_t
  cudaStream_t side_q = nullptr;
  cudaEvent_t ev_qfork = nullptr, ev_qjoin = nullptr;
  struct {
    bool armed = false, done = false;
    int ebtype = 0;
    double tol = 0, s = 0, norm = 0;
    const void *d_qtab = nullptr; // quantizer table in device memory (else tol / norm above)
    uint64_t first = 0; // elements [first, N) were quantized early
  } early_q;

  // relative L-infinity bound, fp32, tiled 3-D path: max |x| comes out of the finest
  // level's coefficient kernel instead of a separate pass (armed by
  // mgb_compress_lowlevel, produced and collected in refactor.cu)
  unsigned *d_absmax = nullptr;
  struct {
    bool armed = false;
    double tol = 0, s = 0;
  } fused_norm;
  // reciprocal quantizers per level, computed on the device from a norm that never
  // visits the host (quantize.cu: prepare_q_kernel)
  void *d_qtab = nullptr;

  const unsigned char *dtab(uint64_t off) const { return d_tables + off * tsize; }
};

// Number of elements of a user / stream supplied shape with overflow checks: every
// extent below 2^31 (the kernels keep per-dimension indices in 32 bits) and the byte
// size below 2^62.  false: reject the shape.
inline bool mgb_checked_elems(int ndim, const uint64_t *shape, size_t tsize, uint64_t *N) {
  unsigned __int128 n = 1;
  for (int d = 0; d < ndim; d++) {
    if (shape[d] == 0 || shape[d] >= (1ull << 31))
      return false;
    n *= shape[d];
    if (n * tsize >= ((unsigned __int128)1 << 62))
      return false;
  }
  *N = (uint64_t)n;
  return true;
}

// No exception crosses the C ABI (std::bad_alloc from a host table, ...).
#define MGB_NOEXCEPT_CALL(expr)                                                \
  do {                                                                         \
    try {                                                                      \
      return (expr);                                                           \
    } catch (...) {                                                            \
      return MGB_FAILURE;                                                      \
    }                                                                          \
  } while (0)

// Function attributes (dynamic shared memory limits) are per device: true the
// first time the calling site runs on the current device.
inline bool mgb_first_use_on_device(bool (&seen)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (seen[dev])
    return false;
  seen[dev] = true;
  return true;
}

// api.cu: the system's libzstd resolved with dlopen (no header in this image)
struct mgb_zstd_fns {
  size_t (*compress)(void *, size_t, const void *, size_t, int) = nullptr;
  size_t (*decompress)(void *, size_t, const void *, size_t) = nullptr;
  size_t (*bound)(size_t) = nullptr;
  unsigned (*is_error)(size_t) = nullptr;
  bool ok = false;
};
const mgb_zstd_fns &mgb_zstd();

// quantize.cu: scratch of the two-stage norm reduction (2 doubles per block + result)
#define MGB_NORM_PART_DOUBLES (2 * 148 * 8 + 2)

// plan.cu
int mgb_plan_ensure_workspace(mgb_plan *p);
uint64_t mgb_level_elems(const mgb_plan *p, int l);

// refactor.cu
int mgb_decompose_impl(mgb_plan *p, const void *d_in, void *d_out,
                       cudaStream_t st);
int mgb_recompose_impl(mgb_plan *p, const void *d_in, void *d_out,
                       cudaStream_t st);

// launch counter + optional per-kernel timing (CUDA events on the launching
// stream; enabled by mgb_profile_enable, read back by mgb_profile_report)
extern unsigned long long g_mgb_launches;
enum mgb_kernel_id {
  MGB_K_COEF = 0, MGB_K_RESTORE, MGB_K_MASSTRANS, MGB_K_THOMAS_CONTIG,
  MGB_K_THOMAS_STRIDED, MGB_K_AXPY, MGB_K_BOXCOPY, MGB_K_QUANTIZE,
  MGB_K_DEQUANTIZE, MGB_K_OUTLIER_RESTORE, MGB_K_NORM, MGB_K_CODEBOOK,
  MGB_K_CHUNK_BITS, MGB_K_CHUNK_SCAN, MGB_K_ENCODE, MGB_K_SERIALIZE,
  MGB_K_DECODE, MGB_K_PARSE, MGB_K_COUNT
};
void mgb_prof_begin(int id, cudaStream_t st);
void mgb_prof_end(int id, cudaStream_t st);
extern int g_mgb_profile;
#define MGB_LAUNCH(id, st, ...)                                                \
  do {                                                                         \
    if (g_mgb_profile)                                                         \
      mgb_prof_begin(id, st);                                                  \
    __VA_ARGS__;                                                               \
    ++g_mgb_launches;                                                          \
    if (g_mgb_profile)                                                         \
      mgb_prof_end(id, st);                                                    \
  } while (0)
