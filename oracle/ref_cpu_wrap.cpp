// TEST INFRASTRUCTURE -- not part of the product.
// C entry points around the UNMODIFIED reference MGARD-CPU templates
// (/root/reference/include/{TensorMeshHierarchy,shuffle,decompose,
// TensorMultilevelCoefficientQuantizer}.hpp and src/compressors.cpp), compiled
// in place by oracle/Makefile into oracle/_ref/libmgard_cpu_ref.so.  The stage
// order is the one of mgard::compress / mgard::decompress
// (reference include/compress.tpp:35-83).
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "TensorMeshHierarchy.hpp"
#include "TensorMultilevelCoefficientQuantizer.hpp"
#include "compressors.hpp"
#include "decompose.hpp"
#include "format.hpp"
#include "shuffle.hpp"

namespace mgard {
// normally in src/format.cpp:48-54 (not linked: it needs libprotobuf)
template <> pb::Dataset::Type type_to_dataset_type<float>() {
  return pb::Dataset::FLOAT;
}
template <> pb::Dataset::Type type_to_dataset_type<double>() {
  return pb::Dataset::DOUBLE;
}
// normally in src/format.cpp (validates, then returns the field); named by
// src/compressors.cpp:674, whose header-driven entry point is not exercised here
pb::Encoding::Compressor read_encoding_compressor(const pb::Header &header) {
  return header.encoding().compressor();
}
} // namespace mgard

extern "C" {
enum {
  REFCPU_INFO = 0,
  REFCPU_DECOMPOSE,  // nodal values -> shuffled multilevel coefficients
  REFCPU_RECOMPOSE,  // shuffled multilevel coefficients -> nodal values
  REFCPU_QUANTIZE,   // shuffled coefficients -> int64
  REFCPU_DEQUANTIZE, // int64 -> shuffled coefficients
  REFCPU_SHUFFLE,
  REFCPU_UNSHUFFLE
};
struct refcpu_args {
  int32_t op, ndim, dtype, pad;
  const uint64_t *shape;
  const void *coords[4]; // all null: uniform hierarchy constructor
  double s, tol;
  const void *in;
  void *out;
  uint64_t L;        // out
  uint64_t ndof[64]; // out: ndof(l), l = 0..L
};
}

namespace {

template <std::size_t N, typename Real> int run(refcpu_args *a) {
  std::array<std::size_t, N> shape;
  for (std::size_t d = 0; d < N; d++)
    shape[d] = a->shape[d];
  mgard::TensorMeshHierarchy<N, Real> *hp;
  if (a->coords[0]) {
    std::array<std::vector<Real>, N> coords;
    for (std::size_t d = 0; d < N; d++) {
      const Real *c = static_cast<const Real *>(a->coords[d]);
      coords[d].assign(c, c + shape[d]);
    }
    hp = new mgard::TensorMeshHierarchy<N, Real>(shape, coords);
  } else {
    hp = new mgard::TensorMeshHierarchy<N, Real>(shape);
  }
  const mgard::TensorMeshHierarchy<N, Real> &h = *hp;
  const std::size_t ndof = h.ndof();
  a->L = h.L;
  for (std::size_t l = 0; l <= h.L && l < 64; l++)
    a->ndof[l] = h.ndof(l);
  mgard::pb::Header header;
  header.mutable_function_decomposition()->set_transform(
      mgard::pb::FunctionDecomposition::MULTILEVEL_COEFFICIENTS);
  const Real s = a->s, tol = a->tol;
  switch (a->op) {
  case REFCPU_INFO:
    break;
  case REFCPU_SHUFFLE:
    mgard::shuffle(h, static_cast<const Real *>(a->in), static_cast<Real *>(a->out));
    break;
  case REFCPU_UNSHUFFLE:
    mgard::unshuffle(h, static_cast<const Real *>(a->in), static_cast<Real *>(a->out));
    break;
  case REFCPU_DECOMPOSE: {
    Real *u = static_cast<Real *>(a->out);
    mgard::shuffle(h, static_cast<const Real *>(a->in), u);
    mgard::decompose(h, header, u);
    break;
  }
  case REFCPU_RECOMPOSE: {
    std::vector<Real> u(static_cast<const Real *>(a->in),
                        static_cast<const Real *>(a->in) + ndof);
    mgard::recompose(h, header, u.data());
    mgard::unshuffle(h, u.data(), static_cast<Real *>(a->out));
    break;
  }
  case REFCPU_QUANTIZE: {
    const mgard::TensorMultilevelCoefficientQuantizer<N, Real, std::int64_t> Q(h, s, tol);
    std::int64_t *q = static_cast<std::int64_t *>(a->out);
    for (const std::int64_t x : Q(static_cast<const Real *>(a->in)))
      *q++ = x;
    break;
  }
  case REFCPU_DEQUANTIZE: {
    const mgard::TensorMultilevelCoefficientDequantizer<N, std::int64_t, Real> D(h, s, tol);
    const std::int64_t *q = static_cast<const std::int64_t *>(a->in);
    Real *u = static_cast<Real *>(a->out);
    for (const Real x : D(q, q + ndof))
      *u++ = x;
    break;
  }
  default:
    delete hp;
    return -2;
  }
  delete hp;
  return 0;
}

template <typename Real> int by_dim(refcpu_args *a) {
  switch (a->ndim) {
  case 1: return run<1, Real>(a);
  case 2: return run<2, Real>(a);
  case 3: return run<3, Real>(a);
  case 4: return run<4, Real>(a);
  default: return -1;
  }
}

} // namespace

extern "C" int refcpu_run(refcpu_args *a) {
  try {
    return a->dtype == 0 ? by_dim<float>(a) : by_dim<double>(a);
  } catch (const std::exception &e) {
    fprintf(stderr, "refcpu_run: %s\n", e.what());
    return -3;
  }
}

// reference src/compressors.cpp:552-606 / :608-629
extern "C" int64_t refcpu_zlib_compress(const void *src, uint64_t n, void *dst,
                                        uint64_t cap) {
  const mgard::MemoryBuffer<unsigned char> out =
      mgard::compress_memory_z(const_cast<void *>(src), n);
  if (out.size > cap)
    return -(int64_t)out.size;
  memcpy(dst, out.data.get(), out.size);
  return (int64_t)out.size;
}
extern "C" void refcpu_zlib_decompress(const void *src, uint64_t n, void *dst,
                                       uint64_t dst_bytes) {
  mgard::decompress_memory_z(const_cast<void *>(src), n,
                             static_cast<unsigned char *>(dst), dst_bytes);
}

// Preamble of an MGARD-CPU stream, byte order included, through the reference's own
// templates: SIGNATURE (include/format.hpp:28) and serialize<> / deserialize<>
// (include/format.tpp:11-41), combined as write_metadata / read_metadata do
// (src/format.cpp:202-233; format.cpp itself needs libprotobuf and is not linked).
#include <zlib.h>
extern "C" void refcpu_preamble(const unsigned char *header, uint64_t n, unsigned char *out17) {
  memcpy(out17, mgard::SIGNATURE.data(), mgard::SIGNATURE.size());
  const auto sz = mgard::serialize<std::uint_least64_t, mgard::HEADER_SIZE_SIZE>(n);
  uLong crc = crc32_z(0, Z_NULL, 0);
  crc = crc32_z(crc, header, n); // compute_crc32, src/format.cpp:179-187
  const auto cb = mgard::serialize<std::uint_least32_t, mgard::HEADER_CRC32_SIZE>((std::uint_least32_t)crc);
  memcpy(out17 + 5, sz.data(), sz.size());
  memcpy(out17 + 13, cb.data(), cb.size());
}
extern "C" int refcpu_read_preamble(const unsigned char *in17, uint64_t *size, uint32_t *crc) {
  if (memcmp(in17, mgard::SIGNATURE.data(), mgard::SIGNATURE.size()) != 0)
    return -1;
  std::array<unsigned char, mgard::HEADER_SIZE_SIZE> sb;
  std::array<unsigned char, mgard::HEADER_CRC32_SIZE> cb;
  memcpy(sb.data(), in17 + 5, sb.size());
  memcpy(cb.data(), in17 + 13, cb.size());
  *size = mgard::deserialize<std::uint_least64_t, mgard::HEADER_SIZE_SIZE>(sb);
  *crc = mgard::deserialize<std::uint_least32_t, mgard::HEADER_CRC32_SIZE>(cb);
  return 0;
}
