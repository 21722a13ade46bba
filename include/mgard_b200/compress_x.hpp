// C++ mirror of the reference's MGARD-X high-level API for the hot path, as
// thin inline wrappers over the C ABI (include/mgard_b200.h).
//
// Same names, argument order and meaning as the reference's
// include/compress_x.hpp:31-178 (namespace mgard_x): a translation unit that
// includes this header instead of <compress_x.hpp> and links
// libmgard_b200.so compiles unchanged for the subset below.
//
//   compress / decompress overloads      compress_x.hpp:54-146
//   release_cache                        compress_x.hpp:159
//   Config (fields the hot path reads)   mgard-x/Config/Config.h:10-42, Config.cpp:14-43
//   enums                                mgard-x/Utilities/Types.h:18-66
//
// Unsupported Config choices (Hybrid decomposition, SingleDim beyond 3-D, LZ4 second
// stage, Block / Variable domain decomposition, ZFP) return
// compress_status_type::Failure instead of silently doing something else.
#ifndef MGARD_B200_COMPRESS_X_HPP
#define MGARD_B200_COMPRESS_X_HPP

#include <cstddef>
#include <cstdint>
#include <limits>
#include <vector>

#include "../mgard_b200.h"

namespace mgard_x {

using SIZE = uint64_t;
using DIM = uint8_t;
using Byte = unsigned char;

enum class decomposition_type : uint8_t { MultiDim, SingleDim, Hybrid };
enum class error_bound_type : uint8_t { REL, ABS };
enum class lossless_type : uint8_t { Huffman, Huffman_LZ4, Huffman_Zstd, CPU_Lossless };
enum class data_type : uint8_t { Float, Double };
enum class domain_decomposition_type : uint8_t { MaxDim, Block, Variable };
enum class compressor_type : uint8_t { MGARD, ZFP };
enum class device_type : uint8_t { AUTO, SERIAL, OPENMP, CUDA, HIP, SYCL, NONE };
enum class compress_status_type : uint8_t {
  Success,
  Failure,
  OutputTooLargeFailure,
  NotSupportHigherNumberOfDimensionsFailure,
  NotSupportDataTypeFailure,
  BackendNotAvailableFailure
};

struct Config {
  device_type dev_type = device_type::AUTO;
  int dev_id = 0;
  compressor_type compressor = compressor_type::MGARD;
  domain_decomposition_type domain_decomposition = domain_decomposition_type::MaxDim;
  decomposition_type decomposition = decomposition_type::MultiDim;
  double estimate_outlier_ratio = 1.0;
  SIZE huff_dict_size = 8192;
  SIZE huff_block_size = 1024 * 20;
  bool normalize_coordinates = true;
  lossless_type lossless = lossless_type::Huffman;
  int zstd_compress_level = 3;
  int reorder = 0;
  SIZE domain_decomposition_dim = 0;
  // mgard_b200 extension: planes per MaxDim sub-domain (0: decide from free
  // device memory as DomainDecomposer.hpp:199-230 does)
  SIZE domain_decomposition_size = 0;
  bool auto_cache_release = false;
};

namespace detail {
inline bool supported(const Config &c) {
  return c.compressor == compressor_type::MGARD &&
         (c.decomposition == decomposition_type::MultiDim || c.decomposition == decomposition_type::SingleDim) &&
         (c.lossless == lossless_type::Huffman || c.lossless == lossless_type::Huffman_Zstd) &&
         (c.reorder == 0 || c.reorder == 1) &&
         c.domain_decomposition == domain_decomposition_type::MaxDim &&
         c.normalize_coordinates &&
         (c.dev_type == device_type::AUTO || c.dev_type == device_type::CUDA);
}
inline mgb_config to_c(const Config &c) {
  mgb_config m;
  mgb_config_default(&m);
  m.dev_id = c.dev_id;
  m.huff_dict_size = (int32_t)c.huff_dict_size;
  m.huff_block_size = (int32_t)c.huff_block_size;
  m.domain_decomposition_dim = c.domain_decomposition_size ? (int32_t)c.domain_decomposition_dim : -1;
  m.domain_decomposition_size = c.domain_decomposition_size;
  m.lossless = (int32_t)c.lossless;
  m.zstd_compress_level = c.zstd_compress_level;
  m.reorder = c.reorder;
  m.decomposition = c.decomposition == decomposition_type::SingleDim ? 1 : 0;
  return m;
}
inline compress_status_type status(int rc) {
  return rc <= 5 ? (compress_status_type)rc : compress_status_type::Failure;
}
} // namespace detail

inline compress_status_type
compress(DIM D, data_type dtype, std::vector<SIZE> shape, double tol, double s,
         error_bound_type mode, const void *original_data, void *&compressed_data,
         size_t &compressed_size, std::vector<const Byte *> coords, Config config,
         bool output_pre_allocated) {
  if (!detail::supported(config) || shape.size() != D ||
      (!coords.empty() && coords.size() != D))
    return compress_status_type::Failure;
  mgb_config c = detail::to_c(config);
  std::vector<const void *> cp(coords.begin(), coords.end());
  int rc = mgb_compress((int)D, (int)dtype, shape.data(), tol, s, (int)mode, original_data,
                        &compressed_data, &compressed_size, cp.empty() ? nullptr : cp.data(),
                        &c, output_pre_allocated ? 1 : 0);
  if (config.auto_cache_release)
    mgb_release_cache();
  return detail::status(rc);
}
inline compress_status_type
compress(DIM D, data_type dtype, std::vector<SIZE> shape, double tol, double s,
         error_bound_type mode, const void *original_data, void *&compressed_data,
         size_t &compressed_size, Config config, bool output_pre_allocated) {
  return compress(D, dtype, shape, tol, s, mode, original_data, compressed_data,
                  compressed_size, std::vector<const Byte *>(), config, output_pre_allocated);
}
inline compress_status_type
compress(DIM D, data_type dtype, std::vector<SIZE> shape, double tol, double s,
         error_bound_type mode, const void *original_data, void *&compressed_data,
         size_t &compressed_size, bool output_pre_allocated) {
  return compress(D, dtype, shape, tol, s, mode, original_data, compressed_data,
                  compressed_size, Config(), output_pre_allocated);
}
inline compress_status_type
compress(DIM D, data_type dtype, std::vector<SIZE> shape, double tol, double s,
         error_bound_type mode, const void *original_data, void *&compressed_data,
         size_t &compressed_size, std::vector<const Byte *> coords,
         bool output_pre_allocated) {
  return compress(D, dtype, shape, tol, s, mode, original_data, compressed_data,
                  compressed_size, coords, Config(), output_pre_allocated);
}

inline compress_status_type
decompress(const void *compressed_data, size_t compressed_size, void *&decompressed_data,
           std::vector<SIZE> &shape, data_type &dtype, Config config,
           bool output_pre_allocated) {
  mgb_config c = detail::to_c(config);
  int nd = 0, dt = 0;
  uint64_t shp[MGB_MAX_DIMS];
  int rc = mgb_decompress(compressed_data, compressed_size, &decompressed_data, &c,
                          output_pre_allocated ? 1 : 0, &nd, shp, &dt);
  if (rc == MGB_SUCCESS) {
    shape.assign(shp, shp + nd);
    dtype = (data_type)dt;
  }
  if (config.auto_cache_release)
    mgb_release_cache();
  return detail::status(rc);
}
inline compress_status_type
decompress(const void *compressed_data, size_t compressed_size, void *&decompressed_data,
           Config config, bool output_pre_allocated) {
  std::vector<SIZE> shape;
  data_type dtype;
  return decompress(compressed_data, compressed_size, decompressed_data, shape, dtype, config,
                    output_pre_allocated);
}
inline compress_status_type
decompress(const void *compressed_data, size_t compressed_size, void *&decompressed_data,
           bool output_pre_allocated) {
  return decompress(compressed_data, compressed_size, decompressed_data, Config(),
                    output_pre_allocated);
}
inline compress_status_type
decompress(const void *compressed_data, size_t compressed_size, void *&decompressed_data,
           std::vector<SIZE> &shape, data_type &dtype, bool output_pre_allocated) {
  return decompress(compressed_data, compressed_size, decompressed_data, shape, dtype, Config(),
                    output_pre_allocated);
}

inline compress_status_type release_cache(Config) {
  mgb_release_cache();
  return compress_status_type::Success;
}

// compress_x.hpp:162-178
inline void pin_memory(void *ptr, SIZE num_bytes, Config) { mgb_pin_memory(ptr, num_bytes); }
inline bool check_memory_pinned(void *ptr, Config) { return mgb_check_memory_pinned(ptr) != 0; }
inline void unpin_memory(void *ptr, Config) { mgb_unpin_memory(ptr); }

} // namespace mgard_x

#endif
