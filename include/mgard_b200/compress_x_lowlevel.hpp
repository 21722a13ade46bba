// C++ mirror of the reference's MGARD-X LOW-LEVEL API (doc/MGARD-X.md:205-262;
// include/compress_x_lowlevel.hpp, mgard-x/Hierarchy/Hierarchy.h:18-64,
// mgard-x/CompressionLowLevel/Compressor.h:29-90, mgard-x/RuntimeX/DataStructures/Array.h:15-60)
// as thin inline wrappers over the C ABI (include/mgard_b200.h): same class names,
// template parameters, member names and argument order, so
//
//   mgard_x::Hierarchy<3, float, mgard_x::CUDA> hierarchy(shape, config);
//   mgard_x::Compressor<3, float, mgard_x::CUDA> compressor(hierarchy, config);
//   mgard_x::Array<3, float, mgard_x::CUDA> in_array(shape);   in_array.load(u);
//   mgard_x::Array<1, unsigned char, mgard_x::CUDA> compressed;
//   compressor.Compress(in_array, mgard_x::error_bound_type::REL, tol, s, norm, compressed, 0);
//   mgard_x::DeviceRuntime<mgard_x::CUDA>::SyncQueue(0);
//
// compiles unchanged against this header (tests/cxx/lowlevel_roundtrip.cpp).  Only the
// CUDA device type exists here.  One difference, to the caller's advantage: Compress does
// not alter in_array (the reference destroys its input).
#ifndef MGARD_B200_COMPRESS_X_LOWLEVEL_HPP
#define MGARD_B200_COMPRESS_X_LOWLEVEL_HPP

#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "compress_x.hpp"

namespace mgard_x {

struct CUDA {};
#define MGARDX_SYNCHRONIZED_QUEUE 0

// queues are CUDA streams; queue 0 is the default stream (DeviceRuntime, RuntimeX.h)
template <typename DeviceType> struct DeviceRuntime {
  static void SyncQueue(int) { cudaDeviceSynchronize(); }
  static void SyncDevice() { cudaDeviceSynchronize(); }
  static void SelectDevice(int dev_id) { cudaSetDevice(dev_id); }
};

// Array.h:15-60: a managed dense device array (never pitched here: ld == fastest extent)
template <DIM D, typename T, typename DeviceType> class Array {
public:
  Array() {}
  explicit Array(std::vector<SIZE> shape, bool = true, bool = false, int = 0) { resize(shape); }
  Array(const Array &o) { copy_from(o); }
  Array &operator=(const Array &o) {
    if (this != &o)
      copy_from(o);
    return *this;
  }
  Array(Array &&o) noexcept { swap(o); }
  Array &operator=(Array &&o) noexcept {
    swap(o);
    return *this;
  }
  ~Array() { release(); }
  void resize(std::vector<SIZE> shape, int = 0) {
    SIZE n = 1;
    for (SIZE e : shape)
      n *= e;
    if (n > cap_) {
      release_device();
      if (cudaMalloc(&dv_, n * sizeof(T)) != cudaSuccess)
        throw std::runtime_error("mgard_x::Array: device allocation failed");
      cap_ = n;
    }
    shape_ = shape;
  }
  // data may be a host or a device pointer; ld: leading dimension of the source (0: dense)
  void load(const T *data, SIZE ld = 0, int = 0) {
    const SIZE nf = shape_.empty() ? 0 : shape_.back();
    SIZE rows = 1;
    for (size_t d = 0; d + 1 < shape_.size(); d++)
      rows *= shape_[d];
    if (ld == 0 || ld == nf)
      cudaMemcpy(dv_, data, rows * nf * sizeof(T), cudaMemcpyDefault);
    else
      cudaMemcpy2D(dv_, nf * sizeof(T), data, ld * sizeof(T), nf * sizeof(T), rows, cudaMemcpyDefault);
  }
  T *hostCopy(bool keep = false, int = 0) {
    free(hv_);
    hv_ = (T *)malloc(elems() * sizeof(T));
    cudaMemcpy(hv_, dv_, elems() * sizeof(T), cudaMemcpyDeviceToHost);
    keep_ = keep;
    return hv_;
  }
  T *data(SIZE &ld) {
    ld = shape_.empty() ? 0 : shape_.back();
    return dv_;
  }
  T *data() { return dv_; }
  SIZE &shape(DIM d) { return shape_[d]; }
  std::vector<SIZE> &shape() { return shape_; }
  SIZE ld(DIM d) { return shape_[d]; }
  bool isPitched() { return false; }
  SIZE elems() const {
    SIZE n = shape_.empty() ? 0 : 1;
    for (SIZE e : shape_)
      n *= e;
    return n;
  }

private:
  void release_device() {
    if (dv_)
      cudaFree(dv_);
    dv_ = nullptr;
    cap_ = 0;
  }
  void release() {
    release_device();
    if (hv_ && !keep_)
      free(hv_);
    hv_ = nullptr;
  }
  void copy_from(const Array &o) {
    resize(o.shape_);
    if (o.dv_)
      cudaMemcpy(dv_, o.dv_, elems() * sizeof(T), cudaMemcpyDeviceToDevice);
  }
  void swap(Array &o) {
    std::swap(dv_, o.dv_);
    std::swap(hv_, o.hv_);
    std::swap(cap_, o.cap_);
    std::swap(keep_, o.keep_);
    std::swap(shape_, o.shape_);
  }
  std::vector<SIZE> shape_;
  T *dv_ = nullptr, *hv_ = nullptr;
  SIZE cap_ = 0;
  bool keep_ = false;
};

namespace detail {
template <typename T> struct dtype_of;
template <> struct dtype_of<float> { static constexpr int value = MGB_F32; };
template <> struct dtype_of<double> { static constexpr int value = MGB_F64; };
} // namespace detail

// Hierarchy.h:18-64.  The plan behind it also carries the Compressor's workspaces
// (allocated on first use), so a Compressor shares its Hierarchy's plan.
template <DIM D, typename T, typename DeviceType> class Hierarchy {
public:
  Hierarchy() {}
  Hierarchy(std::vector<SIZE> shape, Config config) { init(shape, std::vector<T *>(), config); }
  Hierarchy(std::vector<SIZE> shape, std::vector<T *> coords, Config config) { init(shape, coords, config); }
  Hierarchy(const Hierarchy &o) : shape_(o.shape_), coords_(o.coords_), config_(o.config_) {
    if (o.plan_)
      build();
  }
  Hierarchy &operator=(const Hierarchy &o) {
    if (this != &o) {
      destroy();
      shape_ = o.shape_;
      coords_ = o.coords_;
      config_ = o.config_;
      if (o.plan_)
        build();
    }
    return *this;
  }
  ~Hierarchy() { destroy(); }
  SIZE total_num_elems() { return mgb_plan_num_elems(plan_); }
  SIZE l_target() { return (SIZE)mgb_plan_l_target(plan_); }
  SIZE level_shape(SIZE level, DIM dim) { return mgb_plan_level_shape(plan_, (int)level, (int)dim); }
  std::vector<SIZE> level_shape(SIZE level) {
    std::vector<SIZE> s(D);
    for (DIM d = 0; d < D; d++)
      s[d] = level_shape(level, d);
    return s;
  }
  // Hierarchy::can_reuse (Hierarchy.hpp:722-733)
  bool can_reuse(std::vector<SIZE> shape) { return coords_.empty() && shape == shape_; }
  mgb_plan *plan() const { return plan_; }
  bool uniform() const { return coords_.empty(); }

private:
  void init(std::vector<SIZE> shape, std::vector<T *> coords, Config config) {
    if (shape.size() != D || (!coords.empty() && coords.size() != D))
      throw std::invalid_argument("mgard_x::Hierarchy: shape / coordinates do not match D");
    shape_ = shape;
    coords_.clear();
    for (size_t d = 0; d < coords.size(); d++)
      coords_.emplace_back(coords[d], coords[d] + shape[d]);
    config_ = config;
    build();
  }
  void build() {
    mgb_config c = detail::to_c(config_);
    std::vector<const void *> cp;
    for (auto &v : coords_)
      cp.push_back(v.data());
    const int rc = mgb_plan_create((int)D, shape_.data(), detail::dtype_of<T>::value, cp.empty() ? nullptr : cp.data(),
                                   &c, &plan_);
    if (rc != MGB_SUCCESS)
      throw std::runtime_error("mgard_x::Hierarchy: mgb_plan_create failed");
  }
  void destroy() {
    if (plan_)
      mgb_plan_destroy(plan_);
    plan_ = nullptr;
  }
  std::vector<SIZE> shape_;
  std::vector<std::vector<T>> coords_;
  Config config_;
  mgb_plan *plan_ = nullptr;
};

// Compressor.h:29-90.  queue_idx selects the CUDA stream; only queue 0 (the default
// stream) exists here, and the calls return when the block is complete (the reference is
// asynchronous until DeviceRuntime::SyncQueue - calling it afterwards is harmless).
template <DIM D, typename T, typename DeviceType> class Compressor {
public:
  Compressor() : initialized(false), hierarchy(nullptr) {}
  Compressor(Hierarchy<D, T, DeviceType> &hierarchy, Config config)
      : initialized(true), hierarchy(&hierarchy), config(config) {}
  void Adapt(Hierarchy<D, T, DeviceType> &h, Config c, int) {
    hierarchy = &h;
    config = c;
    initialized = true;
  }
  static size_t EstimateMemoryFootprint(std::vector<SIZE> shape, Config) {
    size_t n = 1;
    for (SIZE e : shape)
      n *= e;
    return (size_t)(n * (sizeof(T) * 5.5 + 2.0)) + (64u << 20);
  }
  void Compress(Array<D, T, DeviceType> &original_data, enum error_bound_type ebtype, T tol, T s, T &norm,
                Array<1, Byte, DeviceType> &compressed_data, int queue_idx) {
    (void)queue_idx;
    const SIZE n = hierarchy->total_num_elems();
    // room for a block that does not compress (the low-level call has no raw fallback;
    // the high-level one stores such a sub-domain uncompressed, GPUPipelines.hpp:139-155)
    const uint64_t cap = 2 * n * sizeof(T) + 2 * (1024 + 8ull * config.huff_dict_size) +
                         32 * ((n - 1) / config.huff_block_size + 1) + (1u << 20);
    compressed_data.resize({(SIZE)cap});
    double nrm = (double)norm;
    uint64_t size = 0;
    const int rc = mgb_compress_lowlevel(hierarchy->plan(), original_data.data(), (int)ebtype, (double)tol,
                                         (double)s, &nrm, compressed_data.data(), cap, &size, nullptr);
    if (rc != MGB_SUCCESS)
      throw std::runtime_error("mgard_x::Compressor::Compress failed with status " + std::to_string(rc));
    norm = (T)nrm;
    compressed_data.shape(0) = size;
  }
  void Decompress(Array<1, Byte, DeviceType> &compressed_data, enum error_bound_type ebtype, T tol, T s, T &norm,
                  Array<D, T, DeviceType> &decompressed_data, int queue_idx) {
    (void)queue_idx;
    std::vector<SIZE> shape(D);
    for (DIM d = 0; d < D; d++)
      shape[d] = hierarchy->level_shape(hierarchy->l_target(), d);
    decompressed_data.resize(shape);
    const int rc = mgb_decompress_lowlevel(hierarchy->plan(), compressed_data.data(), compressed_data.shape(0),
                                           (int)ebtype, (double)tol, (double)s, (double)norm,
                                           decompressed_data.data(), nullptr);
    if (rc != MGB_SUCCESS)
      throw std::runtime_error("mgard_x::Compressor::Decompress failed with status " + std::to_string(rc));
  }
  // the stages of Compress / Decompress, device resident (Compressor.h:41-71)
  void Decompose(Array<D, T, DeviceType> &original_data, int) {
    Array<D, T, DeviceType> out(original_data.shape());
    if (mgb_decompose(hierarchy->plan(), original_data.data(), out.data(), nullptr) != MGB_SUCCESS)
      throw std::runtime_error("mgard_x::Compressor::Decompose failed");
    original_data = std::move(out);
  }
  void Recompose(Array<D, T, DeviceType> &decompressed_data, int) {
    Array<D, T, DeviceType> out(decompressed_data.shape());
    if (mgb_recompose(hierarchy->plan(), decompressed_data.data(), out.data(), nullptr) != MGB_SUCCESS)
      throw std::runtime_error("mgard_x::Compressor::Recompose failed");
    decompressed_data = std::move(out);
  }
  void CalculateNorm(Array<D, T, DeviceType> &original_data, enum error_bound_type ebtype, T s, T &norm, int) {
    if (ebtype != error_bound_type::REL)
      return;
    double r = 0;
    if (mgb_norm(hierarchy->plan(), original_data.data(), (double)s, &r) != MGB_SUCCESS)
      throw std::runtime_error("mgard_x::Compressor::CalculateNorm failed");
    norm = (T)r;
  }

  bool initialized;
  Hierarchy<D, T, DeviceType> *hierarchy;
  Config config;
};

} // namespace mgard_x

#endif
