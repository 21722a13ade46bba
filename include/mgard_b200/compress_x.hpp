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
// stage, ZFP, non-CUDA devices) return compress_status_type::Failure instead of
// silently doing something else.
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

enum class cpu_parallelization_mode : uint8_t { INTRA_BLOCK, INTER_BLOCK };

// Every field of the reference's struct, same names, types and defaults
// (mgard-x/Config/Config.h:10-42, Config.cpp:14-43), so code that fills a Config for
// the reference compiles against this header.  What the hot path reads travels in
// mgb_config; fields that select code this engine does not have are checked by
// detail::supported() - the call fails, they are never silently ignored; fields that
// only tune the reference's runtime (logging, prefetch, CPU threading, dry runs) are
// accepted and have no effect.
struct Config {
  device_type dev_type = device_type::AUTO;
  int dev_id = 0;
  compressor_type compressor = compressor_type::MGARD;
  domain_decomposition_type domain_decomposition = domain_decomposition_type::MaxDim;
  decomposition_type decomposition = decomposition_type::MultiDim;
  double estimate_outlier_ratio = 1.0;
  SIZE huff_dict_size = 8192;
  SIZE huff_block_size = 1024 * 20;
  SIZE lz4_block_size = 1 << 15;
  int zstd_compress_level = 3;
  bool normalize_coordinates = true;
  lossless_type lossless = lossless_type::Huffman;
  int reorder = 0;
  int log_level = 0;
  bool prefetch = false;
  bool auto_pin_host_buffers = true;
  SIZE max_larget_level = std::numeric_limits<SIZE>::max();
  SIZE max_memory_footprint = std::numeric_limits<SIZE>::max();
  SIZE total_num_bitplanes = 32;
  SIZE block_size = 256;
  SIZE domain_decomposition_dim = 0;
  std::vector<SIZE> domain_decomposition_sizes;
  bool mdr_adaptive_resolution = false;
  bool adjust_shape = false;
  bool compress_with_dryrun = false;
  int num_local_refactoring_level = 1;
  bool auto_cache_release = false;
  cpu_parallelization_mode cpu_mode = cpu_parallelization_mode::INTER_BLOCK;
  // mgard_b200 extension: planes per MaxDim sub-domain along domain_decomposition_dim
  // (0: the largest dimension, halved until the working set fits, as
  // DomainDecomposer.hpp:199-230 does)
  SIZE domain_decomposition_size = 0;
  void apply() {}
};

namespace detail {
inline bool supported(const Config &c) {
  return c.compressor == compressor_type::MGARD &&
         (c.decomposition == decomposition_type::MultiDim || c.decomposition == decomposition_type::SingleDim) &&
         (c.lossless == lossless_type::Huffman || c.lossless == lossless_type::Huffman_Zstd) &&
         (c.reorder == 0 || c.reorder == 1) &&
         c.normalize_coordinates && !c.compress_with_dryrun &&
         (c.dev_type == device_type::AUTO || c.dev_type == device_type::CUDA);
}
inline mgb_config to_c(const Config &c) {
  mgb_config m;
  mgb_config_default(&m);
  m.dev_id = c.dev_id;
  m.huff_dict_size = (int32_t)c.huff_dict_size;
  m.huff_block_size = (int32_t)c.huff_block_size;
  m.domain_decomposition = (int32_t)c.domain_decomposition;
  if (c.domain_decomposition == domain_decomposition_type::Variable)
    m.domain_decomposition_dim = (int32_t)c.domain_decomposition_dim;
  else
    m.domain_decomposition_dim = c.domain_decomposition_size ? (int32_t)c.domain_decomposition_dim : -1;
  m.domain_decomposition_size = c.domain_decomposition_size;
  m.domain_decomposition_sizes = c.domain_decomposition_sizes.empty() ? nullptr : c.domain_decomposition_sizes.data();
  m.num_domain_decomposition_sizes = c.domain_decomposition_sizes.size();
  m.block_size = c.block_size;
  m.max_larget_level = c.max_larget_level;
  m.max_memory_footprint = c.max_memory_footprint;
  m.lossless = (int32_t)c.lossless;
  m.zstd_compress_level = c.zstd_compress_level;
  m.reorder = c.reorder;
  m.decomposition = c.decomposition == decomposition_type::SingleDim ? 1 : 0;
  return m;
}
// Config::adjust_shape (CompressionHighLevel/ShapeAdjustment.hpp:43-84): the prime
// factors of the largest extent, largest first, go to whichever dimension is currently
// the smallest; the data are reinterpreted with the new shape.
inline void adjust_shape(std::vector<SIZE> &shape, const Config &c) {
  SIZE steps = 1;
  const bool variable = c.domain_decomposition == domain_decomposition_type::Variable;
  if (variable && !c.domain_decomposition_sizes.empty()) {
    steps = shape[0] / c.domain_decomposition_sizes[0];
    shape[0] = c.domain_decomposition_sizes[0];
  }
  size_t big = 0;
  for (size_t d = 1; d < shape.size(); d++)
    if (shape[d] > shape[big])
      big = d;
  std::vector<SIZE> factors;
  SIZE n = shape[big];
  for (SIZE z = 2; z * z <= n;) {
    if (n % z == 0) {
      factors.push_back(z);
      n /= z;
    } else {
      z++;
    }
  }
  if (n > 1)
    factors.push_back(n);
  shape[big] = 1;
  for (size_t k = factors.size(); k-- > 0;) {
    size_t small = 0;
    for (size_t d = 1; d < shape.size(); d++)
      if (shape[d] < shape[small])
        small = d;
    shape[small] *= factors[k];
  }
  if (variable)
    shape[0] *= steps;
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
  if (config.adjust_shape) {
    if (!coords.empty())
      return compress_status_type::Failure;
    detail::adjust_shape(shape, config);
  }
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
