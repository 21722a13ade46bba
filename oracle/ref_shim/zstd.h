/* Minimal declarations of the stable zstd C ABI (the image ships
 * libzstd.so.1 without its header). Test infrastructure only. */
#ifndef ORACLE_ZSTD_SHIM_H
#define ORACLE_ZSTD_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
size_t ZSTD_compress(void *dst, size_t dstCapacity, const void *src,
                     size_t srcSize, int compressionLevel);
size_t ZSTD_decompress(void *dst, size_t dstCapacity, const void *src,
                       size_t compressedSize);
size_t ZSTD_compressBound(size_t srcSize);
unsigned ZSTD_isError(size_t code);
const char *ZSTD_getErrorName(size_t code);
unsigned long long ZSTD_getFrameContentSize(const void *src, size_t srcSize);
#ifdef __cplusplus
}
#endif
#endif
