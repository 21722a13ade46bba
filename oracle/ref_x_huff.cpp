/* TEST INFRASTRUCTURE — not part of the product path.
 * C-ABI access to the reference's Huffman stage alone (SERIAL adapter):
 *   include/mgard-x/Lossless/ParallelHuffman/Huffman.hpp:61-362
 *   include/mgard-x/Lossless/ParallelHuffman/GetCodebook.hpp:23-146
 * plus the (D,T) dispatcher for ref_x_wrap.cpp. */
#include "compress_x_lowlevel.hpp"
#include "ref_x_wrap.h"
#include <cstring>

using namespace mgard_x;
using Dev = SERIAL;
using HuffT = Huffman<QUANTIZED_UNSIGNED_INT, QUANTIZED_INT, HUFFMAN_CODE, Dev>;

#define DECL(D, S) extern "C" int refx_run_##D##S(refx_args *a);
DECL(1, f) DECL(2, f) DECL(3, f) DECL(4, f) DECL(5, f)
DECL(1, d) DECL(2, d) DECL(3, d) DECL(4, d) DECL(5, d)

extern "C" int refx_run(refx_args *a) {
#define CASE(D)                                                                \
  case D:                                                                      \
    return a->dtype == 0 ? refx_run_##D##f(a) : refx_run_##D##d(a);
  switch (a->ndim) {
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5)
  }
  return 1;
}

extern "C" int refx_huffman_compress(const uint64_t *symbols, uint64_t n,
                                     int dict_size, int chunk_size,
                                     uint8_t *out, uint64_t cap,
                                     uint64_t *out_size) {
  HuffT huff(n, dict_size, chunk_size, 1.0);
  Array<1, QUANTIZED_UNSIGNED_INT, Dev> primary({(SIZE)n});
  primary.load((QUANTIZED_UNSIGNED_INT *)symbols);
  Array<1, Byte, Dev> compressed;
  huff.outlier_count = 0;
  huff.CompressPrimary(primary, compressed, 0);
  huff.Serialize(compressed, 0);
  DeviceRuntime<Dev>::SyncQueue(0);
  *out_size = compressed.shape(0);
  if (compressed.shape(0) > cap)
    return 2;
  memcpy(out, compressed.hostCopy(), compressed.shape(0));
  return 0;
}

extern "C" int refx_huffman_decompress(const uint8_t *in, uint64_t in_size,
                                       uint64_t *symbols, uint64_t n) {
  HuffT huff(n, 8192, 20480, 1.0);
  Array<1, Byte, Dev> compressed({(SIZE)in_size});
  compressed.load((Byte *)in);
  Array<1, QUANTIZED_UNSIGNED_INT, Dev> primary({(SIZE)n});
  huff.Deserialize(compressed, 0);
  huff.DecompressPrimary(compressed, primary, 0);
  DeviceRuntime<Dev>::SyncQueue(0);
  memcpy(symbols, primary.hostCopy(), n * sizeof(uint64_t));
  return 0;
}

extern "C" int refx_codebook(const uint32_t *freq, int dict_size,
                             uint64_t *codebook, uint8_t *decodebook,
                             uint32_t *cl_out, int *nz_out) {
  HuffT huff(1024, dict_size, 1024, 1.0);
  huff.workspace.reset(0);
  MemoryManager<Dev>::Copy1D(huff.workspace.freq_subarray.data(),
                             (unsigned int *)freq, dict_size, 0);
  GetCodebook(dict_size, huff.workspace.freq_subarray,
              huff.workspace.codebook_subarray,
              huff.workspace.decodebook_subarray, huff.workspace, 0);
  DeviceRuntime<Dev>::SyncQueue(0);
  memcpy(codebook, huff.workspace.codebook_subarray.data(),
         dict_size * sizeof(uint64_t));
  memcpy(decodebook, huff.workspace.decodebook_subarray.data(),
         8 * 128 + 8 * (size_t)dict_size);
  int nz = 0;
  for (int i = 0; i < dict_size; i++)
    nz += freq[i] != 0;
  *nz_out = nz;
  memcpy(cl_out, huff.workspace.CL_subarray.data(), nz * sizeof(uint32_t));
  return 0;
}
