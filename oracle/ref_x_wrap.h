/* TEST INFRASTRUCTURE. Argument block for oracle/_ref/libmgardx_ref.so
 * (see ref_x_wrap.cpp). Plain C so that ctypes can mirror it. */
#ifndef REF_X_WRAP_H
#define REF_X_WRAP_H
#include <stdint.h>

enum {
  REFX_OP_TABLES = 0,
  REFX_OP_DECOMPOSE = 1,
  REFX_OP_RECOMPOSE = 2,
  REFX_OP_COMPRESS = 3,
  REFX_OP_DECOMPRESS = 4
};

typedef struct refx_args {
  int32_t op;
  int32_t ndim;
  int32_t dtype; /* 0 = float, 1 = double */
  int32_t ebtype; /* 0 = REL, 1 = ABS (mgard_x::error_bound_type) */
  int32_t s_is_inf;
  int32_t dict_size;
  int32_t chunk_size;
  int32_t l_target; /* out */
  const uint64_t *shape;
  const void *coords[5]; /* T arrays, or coords[0] == NULL for uniform */
  void *data;            /* in/out field, dense row major */
  double tol, s, norm;   /* norm: out on compress, in on decompress */
  void *tables_out;
  uint64_t tables_count;  /* out */
  void *decomposed_out;   /* optional T[N] */
  int64_t *quantized_out; /* optional int64[N] (dict-shifted, outliers = 0) */
  uint64_t outlier_count; /* out */
  uint8_t *payload;
  uint64_t payload_cap;
  uint64_t payload_size; /* out on compress, in on decompress */
  int32_t lossless;      /* mgard_x::lossless_type: 0 Huffman, 2 Huffman_Zstd */
  int32_t zstd_level;    /* 0: reference default (3) */
  int32_t reorder;       /* Config::reorder: 1 = level-linearised quantised order */
  int32_t decomposition; /* decomposition_type: 0 MultiDim, 1 SingleDim */
  int32_t max_level;     /* Config::max_larget_level; <= 0: no limit */
} refx_args;

#ifdef __cplusplus
extern "C" {
#endif
int refx_run(refx_args *a);
/* Huffman stage alone: Q = uint64 symbols in [0, dict_size). */
int refx_huffman_compress(const uint64_t *symbols, uint64_t n, int dict_size,
                          int chunk_size, uint8_t *out, uint64_t cap,
                          uint64_t *out_size);
int refx_huffman_decompress(const uint8_t *in, uint64_t in_size,
                            uint64_t *symbols, uint64_t n);
/* Codebook alone: freq[dict] -> codebook[dict] (len<<56|code), decodebook
 * bytes (first[64] entry[64] keys[dict]), CL (nz_dict entries, ascending
 * length order as left by GenerateCW) */
int refx_codebook(const uint32_t *freq, int dict_size, uint64_t *codebook,
                  uint8_t *decodebook, uint32_t *cl_out, int *nz_out);
#ifdef __cplusplus
}
#endif
#endif
