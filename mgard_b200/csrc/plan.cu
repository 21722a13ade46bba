// Plan construction: host-side hierarchy bookkeeping (C++), device upload.
// Restates mgard_x::Hierarchy<D,T>::init and helpers
// (reference include/mgard-x/Hierarchy/Hierarchy.hpp:23-162,193-418,686-706)
// with the reference's operation order in precision T, so that dist / ratio /
// am / bm are bit-identical to the reference's tables.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <utility>

#include "plan.h"

unsigned long long g_mgb_launches = 0;
int g_mgb_profile = 0;

namespace {
struct ProfSlot {
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
};
ProfSlot g_prof[MGB_K_COUNT];
const char *g_prof_names[MGB_K_COUNT] = {
    "coef", "restore", "mass_trans", "thomas_contig", "thomas_strided", "axpy",
    "box_copy", "quantize_hist", "dequantize", "outlier_restore", "norm", "codebook",
    "chunk_bits", "chunk_scan", "encode", "serialize", "decode", "parse"};
} // namespace

void mgb_prof_begin(int id, cudaStream_t st) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a, st);
  g_prof[id].ev.push_back({a, b});
}
void mgb_prof_end(int id, cudaStream_t st) {
  cudaEventRecord(g_prof[id].ev.back().second, st);
}

extern "C" void mgb_profile_enable(int on) {
  g_mgb_profile = on;
  if (on)
    for (int k = 0; k < MGB_K_COUNT; k++) {
      for (auto &e : g_prof[k].ev) {
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
      }
      g_prof[k].ev.clear();
    }
}

// per kernel family: launches, total ms, max single-launch ms
extern "C" int mgb_profile_report(int id, const char **name, unsigned long long *launches,
                                  double *total_ms, double *max_ms) {
  if (id < 0 || id >= MGB_K_COUNT)
    return MGB_BAD_ARGUMENT;
  cudaDeviceSynchronize();
  double tot = 0, mx = 0;
  for (auto &e : g_prof[id].ev) {
    float ms = 0;
    cudaEventElapsedTime(&ms, e.first, e.second);
    tot += ms;
    mx = std::max(mx, (double)ms);
  }
  if (name) *name = g_prof_names[id];
  if (launches) *launches = g_prof[id].ev.size();
  if (total_ms) *total_ms = tot;
  if (max_ms) *max_ms = mx;
  return MGB_SUCCESS;
}

namespace {

template <typename T> struct DimTab {
  std::vector<T> dist, ratio, am, bm, fw, mt;
};

// Per coarse index i the nine constants mass_trans evaluates from h1..h4
// (reference Correction/LPKFunctor.h:47-66, non-FMA branch), with
// h1..h4 = dist[2i-2 .. 2i+1] and zeros outside [0, n)
// (Correction/LinearProcessingKernel3D.hpp:262-300).
template <typename T>
std::vector<T> calc_mass_trans(const std::vector<T> &dist, size_t nc) {
  size_t n = dist.size();
  auto h = [&](long long k) -> T {
    return (k < 0 || k >= (long long)n) ? (T)0 : dist[k];
  };
  std::vector<T> mt(9 * nc);
  for (size_t i = 0; i < nc; i++) {
    long long b = 2 * (long long)i;
    T h1 = h(b - 2), h2 = h(b - 1), h3 = h(b), h4 = h(b + 1);
    T r1, r4;
    if (h1 + h2 != 0)
      r1 = h1 / (h1 + h2);
    else
      r1 = 0.0;
    if (h3 + h4 != 0)
      r4 = h4 / (h3 + h4);
    else
      r4 = 0.0;
    mt[0 * nc + i] = h1 / 6;
    mt[1 * nc + i] = (h1 + h2) / 3;
    mt[2 * nc + i] = h2 / 6;
    mt[3 * nc + i] = (h2 + h3) / 3;
    mt[4 * nc + i] = h3 / 6;
    mt[5 * nc + i] = (h3 + h4) / 3;
    mt[6 * nc + i] = h4 / 6;
    mt[7 * nc + i] = r1;
    mt[8 * nc + i] = r4;
  }
  return mt;
}

// Hierarchy.hpp:23-51
template <typename T> std::vector<T> coord_to_dist(const std::vector<T> &c) {
  size_t n = c.size();
  std::vector<T> d(n, (T)0);
  for (size_t i = 0; i + 1 < n; i++)
    d[i] = c[i + 1] - c[i];
  if (n != 2 && n % 2 == 0) {
    T last = d[n - 2];
    d[n - 2] = last / 2.0;
    d[n - 1] = last / 2.0;
  }
  return d;
}

// Hierarchy.hpp:53-80
template <typename T> std::vector<T> dist_to_ratio(const std::vector<T> &d) {
  size_t n = d.size();
  std::vector<T> r(n, (T)0);
  for (size_t i = 0; i + 2 < n; i++)
    r[i] = d[i] / (d[i + 1] + d[i]);
  if (n % 2 == 0)
    r[n - 2] = d[n - 2] / (d[n - 1] + d[n - 2]);
  return r;
}

// Hierarchy.hpp:82-109
template <typename T> std::vector<T> reduce_dist(const std::vector<T> &d) {
  size_t n = d.size();
  size_t n2 = n / 2 + 1;
  std::vector<T> d2(n2, (T)0);
  for (size_t i = 0; i + 1 < n2; i++)
    d2[i] = d[i * 2] + d[i * 2 + 1];
  if (n2 != 2 && n2 % 2 == 0) {
    T last = d2[n2 - 2];
    d2[n2 - 2] = last / 2.0;
    d2[n2 - 1] = last / 2.0;
  }
  return d2;
}

// Hierarchy.hpp:112-162 (non-FMA branch)
template <typename T>
void calc_am_bm(const std::vector<T> &dist, std::vector<T> &am,
                std::vector<T> &bm, std::vector<T> &fw) {
  size_t n = dist.size();
  std::vector<T> ham(n + 1, (T)0), hbm(n + 1, (T)0);
  hbm[0] = 2 * dist[0] / 6;
  ham[0] = 0.0;
  for (size_t i = 1; i + 1 < n; i++) {
    T a_j = dist[i - 1] / 6;
    T w = a_j / hbm[i - 1];
    hbm[i] = 2 * (dist[i - 1] + dist[i]) / 6 - w * a_j;
    ham[i] = a_j;
  }
  T a_j = dist[n - 2] / 6;
  T w = a_j / hbm[n - 2];
  hbm[n - 1] = 2 * dist[n - 2] / 6 - w * a_j;
  ham[n - 1] = a_j;
  am.assign(n + 1, (T)0);
  bm.assign(n + 1, (T)0);
  for (size_t i = 0; i < n; i++) {
    am[i] = ham[i];
    bm[i + 1] = hbm[i];
  }
  bm[0] = 1;
  am[n] = 0;
  // forward-elimination multiplier used by tridiag_forward2
  // (Correction/IPKFunctor.h:29: curr - prev * (am / bm))
  fw.assign(n + 1, (T)0);
  for (size_t i = 0; i <= n; i++)
    fw[i] = am[i] / bm[i];
}

template <typename T>
int build_tables(mgb_plan *p, const void *const *coords_in) {
  const int D = p->D, L = p->L;
  std::vector<std::vector<DimTab<T>>> tabs(L + 1, std::vector<DimTab<T>>(D));
  p->coords.assign(D, std::vector<double>());
  for (int d = 0; d < D; d++) {
    size_t n = p->shape[d];
    std::vector<T> c(n);
    if (coords_in) {
      const T *src = (const T *)coords_in[d];
      for (size_t i = 0; i < n; i++)
        c[i] = src[i];
      for (size_t i = 0; i + 1 < n; i++)
        if (!(c[i + 1] > c[i]))
          return MGB_BAD_ARGUMENT;
    } else {
      // create_uniform_coords with normalize_coordinates (Hierarchy.hpp:686-706)
      for (size_t i = 0; i < n; i++)
        c[i] = (T)i / (p->shape[d] - 1);
    }
    p->coords[d].assign(c.begin(), c.end());
    tabs[L][d].dist = coord_to_dist(c);
    tabs[L][d].ratio = dist_to_ratio(tabs[L][d].dist);
  }
  for (int l = L - 1; l >= 0; l--)
    for (int d = 0; d < D; d++) {
      tabs[l][d].dist = reduce_dist(tabs[l + 1][d].dist);
      tabs[l][d].ratio = dist_to_ratio(tabs[l][d].dist);
    }
  uint64_t off = 0;
  for (int l = 0; l <= L; l++)
    for (int d = 0; d < D; d++) {
      DimTab<T> &t = tabs[l][d];
      calc_am_bm(t.dist, t.am, t.bm, t.fw);
      if (l >= 1)
        t.mt = calc_mass_trans(t.dist, (size_t)p->lshape[l - 1][d]);
      mgb_dim_tables &m = p->tab[l][d];
      m.n = t.dist.size();
      // pad dist/ratio with zeros on both sides so kernels may index
      // [-2, n+2) without bounds checks (LinearProcessingKernel3D.hpp loads
      // zeros outside [0, n))
      off += 4;
      m.dist = off;
      off += m.n + 4;
      off += 4;
      m.ratio = off;
      off += m.n + 4;
      m.am = off;
      off += m.n + 1;
      m.bm = off;
      off += m.n + 1;
      m.fw = off;
      off += m.n + 1;
      m.mt = off;
      off += t.mt.size();
    }
  p->h_tables.assign(off * sizeof(T), 0);
  T *buf = (T *)p->h_tables.data();
  for (int l = 0; l <= L; l++)
    for (int d = 0; d < D; d++) {
      DimTab<T> &t = tabs[l][d];
      mgb_dim_tables &m = p->tab[l][d];
      memcpy(buf + m.dist, t.dist.data(), m.n * sizeof(T));
      memcpy(buf + m.ratio, t.ratio.data(), m.n * sizeof(T));
      memcpy(buf + m.am, t.am.data(), (m.n + 1) * sizeof(T));
      memcpy(buf + m.bm, t.bm.data(), (m.n + 1) * sizeof(T));
      memcpy(buf + m.fw, t.fw.data(), (m.n + 1) * sizeof(T));
      if (!t.mt.empty())
        memcpy(buf + m.mt, t.mt.data(), t.mt.size() * sizeof(T));
    }
  return MGB_SUCCESS;
}

} // namespace

uint64_t mgb_level_elems(const mgb_plan *p, int l) {
  uint64_t n = 1;
  for (int d = 0; d < p->D; d++)
    n *= p->lshape[l][d];
  return n;
}

extern "C" void mgb_config_default(mgb_config *cfg) {
  // src/mgard-x/Config/Config.cpp:14-43
  cfg->dev_id = 0;
  cfg->huff_dict_size = 8192;
  cfg->huff_block_size = 1024 * 20;
  cfg->domain_decomposition_dim = -1;
  cfg->domain_decomposition_size = 0;
  cfg->normalize_coordinates = 1;
  cfg->lossless = 0;
  cfg->zstd_compress_level = 3;
  cfg->reorder = 0;
  cfg->decomposition = 0;
  cfg->domain_decomposition = 0;
  cfg->max_larget_level = ~0ull;
  cfg->block_size = 256;
  cfg->domain_decomposition_sizes = nullptr;
  cfg->num_domain_decomposition_sizes = 0;
  cfg->max_memory_footprint = ~0ull;
}

static int plan_create_impl(int ndim, const uint64_t *shape, int dtype,
                            const void *const *coords,
                            const mgb_config *cfg, mgb_plan **out) {
  if (!out || !shape)
    return MGB_BAD_ARGUMENT;
  *out = nullptr;
  if (ndim < 1 || ndim > MGB_MAX_DIMS)
    return MGB_TOO_MANY_DIMS;
  if (dtype != MGB_F32 && dtype != MGB_F64)
    return MGB_BAD_DTYPE;
  // Hierarchy.hpp:748-756: every dimension must be >= 3
  for (int d = 0; d < ndim; d++)
    if (shape[d] < 3)
      return MGB_BAD_ARGUMENT;
  uint64_t nelems = 0;
  if (!mgb_checked_elems(ndim, shape, dtype == MGB_F32 ? 4 : 8, &nelems))
    return MGB_BAD_ARGUMENT;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return MGB_BACKEND_NOT_AVAILABLE;

  mgb_plan *p = new mgb_plan();
  p->D = ndim;
  p->dtype = dtype;
  p->tsize = dtype == MGB_F32 ? 4 : 8;
  if (cfg)
    p->cfg = *cfg;
  else
    mgb_config_default(&p->cfg);
  if (p->cfg.huff_dict_size < 2 || p->cfg.huff_dict_size > 65536 ||
      p->cfg.huff_block_size < 1) {
    delete p;
    return MGB_BAD_ARGUMENT;
  }
  p->uniform = coords == nullptr;
  p->N = 1;
  for (int d = 0; d < ndim; d++) {
    p->shape[d] = shape[d];
    p->N *= shape[d];
  }
  // Hierarchy.hpp:199-230: n -> n/2+1 until 2; l_target = min over dims
  std::vector<std::vector<uint64_t>> per_dim(ndim);
  size_t nlevel = (size_t)-1;
  for (int d = 0; d < ndim; d++) {
    uint64_t n = shape[d];
    while (n > 2) {
      per_dim[d].push_back(n);
      n = n / 2 + 1;
    }
    per_dim[d].push_back(2);
    nlevel = std::min(nlevel, per_dim[d].size());
  }
  p->L = (int)nlevel - 1;
  // Config::max_larget_level (Hierarchy.hpp:216-217): fewer levels, a larger coarsest mesh
  if ((uint64_t)p->L > p->cfg.max_larget_level)
    p->L = (int)p->cfg.max_larget_level;
  if (p->L >= MGB_MAX_LEVELS) {
    delete p;
    return MGB_BAD_ARGUMENT;
  }
  for (int l = 0; l <= p->L; l++)
    for (int d = 0; d < MGB_MAX_DIMS; d++)
      p->lshape[l][d] = d < ndim ? per_dim[d][p->L - l] : 1;

  int rc = dtype == MGB_F32 ? build_tables<float>(p, coords)
                            : build_tables<double>(p, coords);
  if (rc != MGB_SUCCESS) {
    delete p;
    return rc;
  }
  if (cudaMalloc(&p->d_tables, p->h_tables.size()) != cudaSuccess ||
      cudaMemcpy(p->d_tables, p->h_tables.data(), p->h_tables.size(),
                 cudaMemcpyHostToDevice) != cudaSuccess) {
    mgb_plan_destroy(p);
    return MGB_CUDA_ERROR;
  }
  // level_marks (Hierarchy.hpp:262-282)
  p->marks_width = 0;
  for (int d = 0; d < ndim; d++)
    p->marks_width = std::max<uint64_t>(p->marks_width, shape[d]);
  std::vector<int> marks(ndim * p->marks_width, 0);
  for (int d = 0; d < ndim; d++) {
    uint64_t i = 0;
    for (int l = 0; l <= p->L; l++)
      for (; i < p->lshape[l][d]; i++)
        marks[d * p->marks_width + i] = l;
  }
  if (cudaMalloc(&p->d_marks, marks.size() * sizeof(int)) != cudaSuccess ||
      cudaMemcpy(p->d_marks, marks.data(), marks.size() * sizeof(int),
                 cudaMemcpyHostToDevice) != cudaSuccess) {
    mgb_plan_destroy(p);
    return MGB_CUDA_ERROR;
  }
  *out = p;
  return MGB_SUCCESS;
}

extern "C" int mgb_plan_create(int ndim, const uint64_t *shape, int dtype,
                               const void *const *coords,
                               const mgb_config *cfg, mgb_plan **out) {
  MGB_NOEXCEPT_CALL(plan_create_impl(ndim, shape, dtype, coords, cfg, out));
}

extern "C" void mgb_plan_destroy(mgb_plan *p) {
  if (!p)
    return;
  cudaFree(p->d_tables);
  cudaFree(p->d_marks);
  cudaFree(p->d_coef);
  cudaFree(p->d_cbuf);
  cudaFree(p->d_wA);
  cudaFree(p->d_sd);
  cudaFree(p->d_absmax);
  cudaFree(p->d_qtab);
  cudaFree(p->d_wB);
  if (p->side)
    cudaStreamDestroy(p->side);
  if (p->side_q)
    cudaStreamDestroy(p->side_q);
  if (p->ev_qfork)
    cudaEventDestroy(p->ev_qfork);
  if (p->ev_qjoin)
    cudaEventDestroy(p->ev_qjoin);
  if (p->ev_fork)
    cudaEventDestroy(p->ev_fork);
  if (p->ev_join)
    cudaEventDestroy(p->ev_join);
  cudaFree(p->d_sym);
  cudaFree(p->d_hist);
  cudaFree(p->d_codebook);
  cudaFree(p->d_decodebook);
  cudaFree(p->d_chunk_bits);
  cudaFree(p->d_chunk_woff);
  cudaFree(p->d_chunk_sub);
  cudaFree(p->d_scalars);
  cudaFree(p->d_oidx);
  cudaFree(p->d_oval);
  cudaFree(p->d_norm_tmp);
  cudaFree(p->d_cbwork);
  cudaFree(p->d_dec_sub);
  cudaFree(p->d_declut);
  if (p->h_pinned)
    cudaFreeHost(p->h_pinned);
  delete p;
}

extern "C" int mgb_plan_l_target(const mgb_plan *p) { return p ? p->L : -1; }
extern "C" void mgb_plan_set_generic(mgb_plan *p, int on) {
  if (p)
    p->force_generic = on != 0;
}
extern "C" uint64_t mgb_plan_num_elems(const mgb_plan *p) {
  return p ? p->N : 0;
}
extern "C" uint64_t mgb_plan_level_shape(const mgb_plan *p, int level,
                                         int dim) {
  if (!p || level < 0 || level > p->L)
    return 0;
  if (dim < 0 || dim >= p->D)
    return 1; // Hierarchy.hpp:570-576
  return p->lshape[level][dim];
}

extern "C" uint64_t mgb_plan_table(const mgb_plan *p, int which, int level,
                                   int dim, void *out, uint64_t count) {
  if (!p || level < 0 || level > p->L || dim < 0 || dim >= p->D)
    return 0;
  const mgb_dim_tables &m = p->tab[level][dim];
  uint64_t off, len;
  switch (which) {
  case 0: off = m.dist; len = m.n; break;
  case 1: off = m.ratio; len = m.n; break;
  case 2: off = m.am; len = m.n + 1; break;
  case 3: off = m.bm; len = m.n + 1; break;
  default: return 0;
  }
  if (out)
    memcpy(out, p->h_tables.data() + off * p->tsize,
           std::min(count, len) * p->tsize);
  return len;
}

int mgb_plan_ensure_workspace(mgb_plan *p) {
  if (p->d_cbuf)
    return MGB_SUCCESS;
  // dense coarse boxes for levels L-1 .. 0
  uint64_t off = 0;
  for (int l = p->L - 1; l >= 0; l--) {
    p->cbuf_off[l] = off;
    off += (mgb_level_elems(p, l) + 63) / 64 * 64;
  }
  p->cbuf_elems = off;
  // ping-pong buffers for the mass-trans passes: the first pass output has
  // shape (n_0.. n_{D-2}, nc_{D-1}) of the finest level
  uint64_t w = 1;
  for (int d = 0; d < p->D; d++)
    w *= (d == p->D - 1 && p->L > 0) ? p->lshape[p->L - 1][d] : p->shape[d];
  p->w_elems = w;
  MGB_CUDA_CHECK(cudaMalloc(&p->d_cbuf, std::max<uint64_t>(off, 64) * p->tsize));
  MGB_CUDA_CHECK(cudaMalloc(&p->d_wA, w * p->tsize));
  MGB_CUDA_CHECK(cudaMalloc(&p->d_wB, w * p->tsize));
  return MGB_SUCCESS;
}
