// Internal: parsed / to-be-written stream header (host side).
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/mgard_b200.h"

struct mgb_header {
  int ndim = 0;
  int dtype = MGB_F32;
  uint64_t shape[MGB_MAX_DIMS] = {1, 1, 1, 1, 1};
  int ebtype = MGB_ABS;
  double tol = 0, s = 0, norm = 0;
  bool decomposed = false;
  uint64_t dd_dim = 0, dd_size = 0;
  int dd_method = 1; // pb::DomainDecomposition::Method when decomposed: 1 MAX_DIMENSION, 2 BLOCK, 3 VARIABLE
  int dict_size = 8192, block_size = 20480;
  int lossless = 0; // 0: X_HUFFMAN, 2: X_HUFFMAN_ZSTD (mgard_x::lossless_type values)
  // 0: MGARD-X stream (Metadata.cpp); 1: MGARD-CPU stream (src/format.cpp:110-140:
  // POWER_OF_TWO_PLUS_ONE hierarchy, SHUFFLE preprocessor, CPU_HUFFMAN_ZLIB payload)
  int decomposition = 0; // 0 MULTIDIMENSION_WITH_GHOST_NODES, 1 ONE_DIM_AT_A_TIME_WITH_GHOST_NODES
  int reorder = 0; // Encoding.preprocessor = SHUFFLE (Metadata.cpp:408-412)
  int convention = 0;
  int cpu_compressor = 1; // pb::Encoding::Compressor of an MGARD-CPU stream: 1 zlib, 2 Huffman + zstd
  std::vector<std::vector<double>> coords; // empty: uniform grid
};

std::vector<uint8_t> mgb_encode_stream_header(const mgb_header &h);
int mgb_parse_stream_header(const uint8_t *data, size_t size, mgb_header &h,
                            uint64_t &total_bytes);
// header size announced by the 17-byte preamble in whichever byte order fits the
// stream (UINT64_MAX: not an MGARD preamble / nothing fits)
uint64_t mgb_preamble_header_size(const uint8_t *pre17, size_t stream_size);
