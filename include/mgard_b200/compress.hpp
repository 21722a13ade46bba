// C++ mirror of the reference's MGARD-CPU API, computed on the GPU, as thin inline
// wrappers over the C ABI (include/mgard_b200.h, mgb_cpu_*).
//
// Same names, argument order and meaning as the reference (namespace mgard):
//
//   TensorMeshHierarchy<N, Real>     include/TensorMeshHierarchy.hpp:30-200
//   CompressedDataset<N, Real>       include/CompressedDataset.hpp:18-70
//   DecompressedDataset<N, Real>     include/CompressedDataset.hpp:73-110
//   MemoryBuffer<T>                  include/utilities.hpp:421-451
//   compress / decompress            include/compress.hpp:33-72
//
// A translation unit that includes this header instead of <compress.hpp> and
// links libmgard_b200.so compiles unchanged for this subset.  Errors surface as
// the reference's exception types.  `pb::Header` is not mirrored (protobuf is not
// a dependency here): CompressedDataset keeps the serialised stream instead, and
// `write` emits exactly the bytes the reference's `write` emits.
#ifndef MGARD_B200_COMPRESS_HPP
#define MGARD_B200_COMPRESS_HPP

#include <array>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <vector>

#include "../mgard_b200.h"

namespace mgard {

template <typename T> struct MemoryBuffer {
  MemoryBuffer(T *ptr, const std::size_t size) : data(ptr), size(size) {}
  std::unique_ptr<T[]> data;
  std::size_t size;
};

namespace detail {
inline void throw_status(const int status, const char *where) {
  switch (status) {
  case MGB_SUCCESS: return;
  case MGB_BAD_ARGUMENT: throw std::invalid_argument(where);
  case MGB_FAILURE: throw std::domain_error(where); // e.g. "number too large to be quantized"
  default: throw std::runtime_error(where);
  }
}
template <typename Real> constexpr int dtype_code() {
  static_assert(sizeof(Real) == 4 || sizeof(Real) == 8, "float or double");
  return sizeof(Real) == 4 ? MGB_F32 : MGB_F64;
}
} // namespace detail

template <std::size_t N, typename Real> class TensorMeshHierarchy {
public:
  // uniform nodes on [0, 1] in every dimension
  explicit TensorMeshHierarchy(const std::array<std::size_t, N> &shape) : uniform(true) { init(shape, nullptr); }
  TensorMeshHierarchy(const std::array<std::size_t, N> &shape, const std::array<std::vector<Real>, N> &coordinates)
      : coordinates(coordinates), uniform(false) {
    for (std::size_t i = 0; i < N; ++i)
      if (coordinates.at(i).size() != shape.at(i))
        throw std::invalid_argument("incorrect number of node coordinates given");
    init(shape, &this->coordinates);
  }
  TensorMeshHierarchy(const TensorMeshHierarchy &other)
      : coordinates(other.coordinates), uniform(other.uniform) {
    init(other.shapes.back(), uniform ? nullptr : &coordinates);
  }
  TensorMeshHierarchy &operator=(const TensorMeshHierarchy &) = delete;
  ~TensorMeshHierarchy() { mgb_cpu_plan_destroy(plan_); }

  std::size_t ndof() const { return ndof(L); }
  std::size_t ndof(const std::size_t l) const { return mgb_cpu_plan_ndof(plan_, (int)l); }

  std::vector<std::array<std::size_t, N>> shapes; // shapes[l], l = 0..L
  std::array<std::vector<Real>, N> coordinates;   // empty vectors when uniform
  std::size_t L = 0;
  bool uniform;
  mgb_cpu_plan *plan() const { return plan_; }

private:
  void init(const std::array<std::size_t, N> &shape, const std::array<std::vector<Real>, N> *coords) {
    static_assert(N >= 1 && N <= MGB_MAX_DIMS, "1 to 5 dimensions");
    uint64_t shp[N];
    const void *cptr[N];
    for (std::size_t i = 0; i < N; ++i) {
      shp[i] = shape[i];
      cptr[i] = coords ? (*coords)[i].data() : nullptr;
    }
    const int rc = mgb_cpu_plan_create((int)N, shp, detail::dtype_code<Real>(), coords ? cptr : nullptr, &plan_);
    if (rc == MGB_BAD_ARGUMENT)
      throw std::domain_error("dataset must have size larger than 1 in some dimension and increasing coordinates");
    detail::throw_status(rc, "TensorMeshHierarchy");
    L = (std::size_t)mgb_cpu_plan_levels(plan_);
    shapes.resize(L + 1);
    for (std::size_t l = 0; l <= L; ++l)
      for (std::size_t i = 0; i < N; ++i)
        shapes[l][i] = mgb_cpu_plan_level_shape(plan_, (int)l, (int)i);
  }
  mgb_cpu_plan *plan_ = nullptr;
};

template <std::size_t N, typename Real> class CompressedDataset {
public:
  // `stream` (malloc'ed, owned) = preamble + header + payload
  CompressedDataset(const TensorMeshHierarchy<N, Real> &hierarchy, const Real s, const Real tolerance,
                    unsigned char *stream, const std::size_t stream_size)
      : hierarchy(hierarchy), s(s), tolerance(tolerance), stream_(stream, &std::free), stream_size_(stream_size) {
    uint64_t hb = 0;
    detail::throw_status(mgb_peek_header(stream, stream_size, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                         nullptr, &hb),
                         "CompressedDataset");
    header_bytes_ = hb;
  }
  const TensorMeshHierarchy<N, Real> hierarchy;
  const Real s;
  const Real tolerance;
  // the payload, as in the reference (the header travels separately there)
  void const *data() const { return stream_.get() + header_bytes_; }
  std::size_t size() const { return stream_size_ - header_bytes_; }
  // preamble + header + payload: the reference's self-describing format
  void write(std::ostream &ostream) const {
    ostream.write(reinterpret_cast<const char *>(stream_.get()), (std::streamsize)stream_size_);
  }
  void const *stream() const { return stream_.get(); }
  std::size_t stream_size() const { return stream_size_; }

private:
  std::unique_ptr<unsigned char, void (*)(void *)> stream_;
  std::size_t stream_size_, header_bytes_ = 0;
};

template <std::size_t N, typename Real> class DecompressedDataset {
public:
  DecompressedDataset(const CompressedDataset<N, Real> &compressed, Real const *const data)
      : hierarchy(compressed.hierarchy), s(compressed.s), tolerance(compressed.tolerance), data_(data) {}
  const TensorMeshHierarchy<N, Real> hierarchy;
  const Real s;
  const Real tolerance;
  Real const *data() const { return data_.get(); }

private:
  std::unique_ptr<const Real[]> data_;
};

// include/compress.hpp:33-36.  `v` may be a host or a device pointer.  As in the
// reference, defining MGARD_ZSTD selects the Huffman + zstd payload
// (src/format.cpp:124-131); otherwise the zlib payload is written.
template <std::size_t N, typename Real>
CompressedDataset<N, Real> compress(const TensorMeshHierarchy<N, Real> &hierarchy, Real *const v, const Real s,
                                    const Real tolerance) {
  uint64_t shp[N];
  const void *cptr[N];
  for (std::size_t i = 0; i < N; ++i) {
    shp[i] = hierarchy.shapes.back()[i];
    cptr[i] = hierarchy.uniform ? nullptr : hierarchy.coordinates[i].data();
  }
  void *out = nullptr;
  std::size_t out_size = 0;
  detail::throw_status(mgb_cpu_compress((int)N, detail::dtype_code<Real>(), shp, hierarchy.uniform ? nullptr : cptr,
                                        (double)s, (double)tolerance,
#ifdef MGARD_ZSTD
                                        2,
#else
                                        1,
#endif
                                        v, &out, &out_size),
                       "mgard::compress");
  return CompressedDataset<N, Real>(hierarchy, s, tolerance, static_cast<unsigned char *>(out), out_size);
}

// include/compress.hpp:62-72 (self-describing stream in, raw array bytes out)
inline MemoryBuffer<const unsigned char> decompress(void const *const data, const std::size_t size) {
  void *out = nullptr;
  int ndim = 0, dtype = 0;
  uint64_t shape[MGB_MAX_DIMS];
  detail::throw_status(mgb_cpu_decompress(data, size, &out, &ndim, shape, &dtype), "mgard::decompress");
  std::size_t bytes = dtype == MGB_F32 ? 4 : 8;
  for (int d = 0; d < ndim; ++d)
    bytes *= shape[d];
  unsigned char *copy = new unsigned char[bytes];
  std::memcpy(copy, out, bytes);
  std::free(out);
  return MemoryBuffer<const unsigned char>(copy, bytes);
}

// include/compress.hpp:56-58
template <std::size_t N, typename Real>
DecompressedDataset<N, Real> decompress(const CompressedDataset<N, Real> &compressed) {
  MemoryBuffer<const unsigned char> raw = decompress(compressed.stream(), compressed.stream_size());
  Real *v = new Real[compressed.hierarchy.ndof()];
  std::memcpy(v, raw.data.get(), raw.size);
  return DecompressedDataset<N, Real>(compressed, v);
}

} // namespace mgard
#endif
