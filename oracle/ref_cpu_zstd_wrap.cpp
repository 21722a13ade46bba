// TEST INFRASTRUCTURE -- not part of the product.
// C entry points around the UNMODIFIED reference src/compressors.cpp compiled
// with -DMGARD_ZSTD (the reference's default when libzstd is found,
// CMakeLists.txt:114-127): compress_memory_huffman / decompress_memory_huffman
// (src/compressors.cpp:316-512), i.e. the CPU_HUFFMAN_ZSTD payload.  Built by
// oracle/Makefile into oracle/_ref/libmgard_cpu_ref_zstd.so (a separate object:
// the zlib and zstd builds define the same symbols).
#include <cstdint>
#include <cstring>
#include <vector>

#include "compressors.hpp"

namespace mgard {
pb::Encoding::Compressor read_encoding_compressor(const pb::Header &header) {
  return header.encoding().compressor();
}
// src/format.cpp:56-77 (only INT64_T is used by the Huffman path)
MemoryBuffer<unsigned char> quantization_buffer(const pb::Header &, const std::size_t ndof) {
  return MemoryBuffer<unsigned char>(ndof * sizeof(std::int64_t));
}
} // namespace mgard

// `q` is modified by the reference (build_ft shifts it in place): pass a copy.
extern "C" int64_t refcpu_huffman_zstd_compress(const int64_t *q, uint64_t n, void *dst, uint64_t cap) {
  std::vector<long int> tmp(q, q + n);
  const mgard::MemoryBuffer<unsigned char> out = mgard::compress_memory_huffman(tmp.data(), n);
  if (out.size > cap)
    return -(int64_t)out.size;
  memcpy(dst, out.data.get(), out.size);
  return (int64_t)out.size;
}
extern "C" void refcpu_huffman_zstd_decompress(const void *src, uint64_t n, int64_t *dst, uint64_t dst_bytes) {
  std::vector<unsigned char> tmp((const unsigned char *)src, (const unsigned char *)src + n);
  mgard::decompress_memory_huffman(tmp.data(), n, reinterpret_cast<long int *>(dst), dst_bytes);
}
