/*
 * mgard_b200 — C ABI of the B200-native MGARD hot path: the MGARD-X convention
 * (mgb_*) and the MGARD-CPU convention (mgb_cpu_*, further down).
 *
 * Plain C: pointers, sizes and status codes only.  This is the drop-in
 * boundary: every entry point names the reference interface it replaces
 * (paths under /root/reference).  The C++ mirror of the reference API
 * (mgard_x::compress / mgard_x::decompress, include/compress_x.hpp:31-178) is
 * include/mgard_b200/compress_x.hpp and is a thin inline wrapper over these.
 *
 * All `d_*` pointers are CUDA device pointers on the current device; `stream`
 * is a cudaStream_t passed as void* (NULL = default stream).  Functions are
 * asynchronous on `stream` unless stated; they return 0 (MGB_SUCCESS) or a
 * status mirroring mgard_x::compress_status_type
 * (include/mgard-x/Utilities/Types.h:56-63).  No exceptions cross this ABI,
 * and there is no CPU fallback: without a CUDA device every compute call
 * returns MGB_BACKEND_NOT_AVAILABLE.
 */
#ifndef MGARD_B200_H
#define MGARD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* mgard_x::compress_status_type (Types.h:56-63), same numeric values */
enum {
  MGB_SUCCESS = 0,
  MGB_FAILURE = 1,
  MGB_OUTPUT_TOO_LARGE = 2,
  MGB_TOO_MANY_DIMS = 3,
  MGB_BAD_DTYPE = 4,
  MGB_BACKEND_NOT_AVAILABLE = 5,
  /* extensions (argument / stream-format errors) */
  MGB_BAD_ARGUMENT = 16,
  MGB_BAD_STREAM = 17,
  MGB_CUDA_ERROR = 18
};

/* mgard_x::data_type (Types.h:38) */
enum { MGB_F32 = 0, MGB_F64 = 1 };
/* mgard_x::error_bound_type (Types.h:30) */
enum { MGB_REL = 0, MGB_ABS = 1 };

#define MGB_MAX_DIMS 5

/* The fields of mgard_x::Config (include/mgard-x/Config/Config.h:10-42) that change what
 * the hot path computes; defaults as src/mgard-x/Config/Config.cpp:14-43.  The remaining
 * fields of the reference's struct (logging, prefetch, CPU threading, MDR / ZFP / LZ4
 * knobs ...) are carried by the C++ mirror include/mgard_b200/compress_x.hpp. */
typedef struct mgb_config {
  int32_t dev_id;          /* Config::dev_id */
  int32_t huff_dict_size;  /* 8192 */
  int32_t huff_block_size; /* 20480 */
  int32_t domain_decomposition_dim;   /* -1: largest dim (MaxDim) */
  uint64_t domain_decomposition_size; /* 0: do not decompose */
  int32_t normalize_coordinates;      /* 1 (only value supported) */
  int32_t lossless;            /* mgard_x::lossless_type: 0 Huffman (default), 2 Huffman_Zstd */
  int32_t zstd_compress_level; /* Config::zstd_compress_level, 3 */
  int32_t reorder;             /* Config::reorder: 1 = quantised symbols in level-linearised order (LevelLinearizer) */
  int32_t decomposition;       /* mgard_x::decomposition_type: 0 MultiDim (default), 1 SingleDim (D <= 3) */
  int32_t domain_decomposition; /* mgard_x::domain_decomposition_type: 0 MaxDim (default), 1 Block, 2 Variable */
  uint64_t max_larget_level;   /* Config::max_larget_level: l_target = min(levels - 1, this); UINT64_MAX = no limit
                                  (Hierarchy.hpp:195-217).  Not stored in the stream: the decompressing call must
                                  pass the same value, as with the reference */
  uint64_t block_size;         /* Config::block_size (256): edge of the Block sub-domains */
  const uint64_t *domain_decomposition_sizes; /* Config::domain_decomposition_sizes (Variable): extents along */
  uint64_t num_domain_decomposition_sizes;    /* domain_decomposition_dim; not stored in the stream either */
  uint64_t max_memory_footprint; /* Config::max_memory_footprint: bytes the working set of one sub-domain may
                                    take when the MaxDim / Block size is chosen automatically; UINT64_MAX: what
                                    the device has free */
} mgb_config;

void mgb_config_default(mgb_config *cfg);

/* ---- plan = mgard_x::Hierarchy<D,T> + Compressor<D,T> workspaces ----------
 * replaces Hierarchy ctor (include/mgard-x/Hierarchy/Hierarchy.hpp:193-418,
 * 737-800) and Compressor ctor (CompressionLowLevel/Compressor.hpp:31-57). */
typedef struct mgb_plan mgb_plan;

/* coords: NULL for a uniform grid, else ndim host arrays (dtype T) of length
 * shape[d] (slowest dim first), strictly increasing. */
int mgb_plan_create(int ndim, const uint64_t *shape, int dtype,
                    const void *const *coords, const mgb_config *cfg,
                    mgb_plan **plan);
void mgb_plan_destroy(mgb_plan *plan);
int mgb_plan_l_target(const mgb_plan *plan);
/* Testing aid: route D == 3 through the dimension-generic kernels instead of
 * the fused 3-D level kernel (both are bit-identical by contract). */
void mgb_plan_set_generic(mgb_plan *plan, int on);
uint64_t mgb_plan_num_elems(const mgb_plan *plan);
/* level_shape(l, d), Hierarchy.hpp:560-577 */
uint64_t mgb_plan_level_shape(const mgb_plan *plan, int level, int dim);
/* Host copy of one hierarchy table; which: 0 dist, 1 ratio, 2 am, 3 bm
 * (Hierarchy.hpp:23-162).  Writes `count` values of dtype T to out; returns
 * the table length (n for dist/ratio, n+1 for am/bm). */
uint64_t mgb_plan_table(const mgb_plan *plan, int which, int level, int dim,
                        void *out, uint64_t count);

/* ---- stage entry points (device resident, used by the parity tests) ------ */

/* data_refactoring::multi_dimension::decompose / recompose
 * (DataRefactoring/MultiDimension/DataRefactoring.hpp:25-177,180-317).
 * d_in: dense row-major field (not modified); d_out: coefficients in the
 * reference's in-place "coarse nodes first along every dimension" layout. */
int mgb_decompose(mgb_plan *plan, const void *d_in, void *d_out, void *stream);
int mgb_recompose(mgb_plan *plan, const void *d_in, void *d_out, void *stream);

/* norm_calculator (CompressionLowLevel/NormCalculator.hpp:13-83): max|u| for
 * s = +inf, else sqrt(sum u^2 / N).  Synchronises the stream. */
int mgb_norm(mgb_plan *plan, const void *d_in, double s, double *norm);
/* Partial reductions for the sharded path (max|u| and sum u^2 as doubles). */
int mgb_norm_partials(mgb_plan *plan, const void *d_in, double *absmax,
                      double *sumsq);

/* LevelwiseLinearQuantizerKernel<QUANTIZE> fused with the Huffman histogram
 * (Quantization/LinearQuantization.hpp:148-266,495-545 +
 * Lossless/ParallelHuffman/Histogram.hpp).  d_sym: uint16 symbols (dictionary
 * shifted, outliers = 0); d_hist: uint32[dict] (zeroed by the call);
 * outliers appended to (d_oidx, d_oval); *d_ocount device counter. */
int mgb_quantize(mgb_plan *plan, const void *d_coef, int ebtype, double tol,
                 double s, double norm, uint16_t *d_sym, uint32_t *d_hist,
                 unsigned long long *d_ocount, uint64_t *d_oidx,
                 int64_t *d_oval, uint64_t outlier_cap, void *stream);
/* OutlierRestore + LevelwiseLinearQuantizerKernel<DEQUANTIZE>
 * (LinearQuantization.hpp:251-264,304-350). */
int mgb_dequantize(mgb_plan *plan, const uint16_t *d_sym, uint64_t ocount,
                   const uint64_t *d_oidx, const int64_t *d_oval, int ebtype,
                   double tol, double s, double norm, void *d_coef,
                   void *stream);

/* GetCodebook (Lossless/ParallelHuffman/GetCodebook.hpp:23-146): histogram ->
 * codebook[dict] (len<<56|code) and decodebook (first[64] entry[64]
 * keys[dict], all uint64).  Device resident. */
int mgb_codebook(mgb_plan *plan, const uint32_t *d_hist, uint64_t *d_codebook,
                 uint64_t *d_decodebook, void *stream);

/* Huffman::CompressPrimary + Serialize (ParallelHuffman/Huffman.hpp:61-262):
 * writes the serialised Huffman block to d_out (capacity cap bytes) and its
 * size to *size (host).  Synchronises the stream. */
int mgb_huffman_compress(mgb_plan *plan, const uint16_t *d_sym, uint64_t n,
                         const uint32_t *d_hist, uint64_t ocount,
                         const uint64_t *d_oidx, const int64_t *d_oval,
                         uint8_t *d_out, uint64_t cap, uint64_t *size,
                         void *stream);
/* Huffman::Deserialize + DecompressPrimary (Huffman.hpp:264-362).  Outlier
 * arrays are returned as pointers into d_in. */
int mgb_huffman_decompress(mgb_plan *plan, const uint8_t *d_in, uint64_t size,
                           uint16_t *d_sym, uint64_t n, uint64_t *ocount,
                           const uint64_t **d_oidx, const int64_t **d_oval,
                           void *stream);

/* ---- low level: Compressor<D,T>::Compress / Decompress --------------------
 * (CompressionLowLevel/Compressor.hpp:193-272) on one sub-domain.  `norm` is
 * in/out exactly as the reference's `T &norm`.  Synchronous. */
int mgb_compress_lowlevel(mgb_plan *plan, const void *d_in, int ebtype,
                          double tol, double s, double *norm, uint8_t *d_out,
                          uint64_t cap, uint64_t *size, void *stream);
int mgb_decompress_lowlevel(mgb_plan *plan, const uint8_t *d_in, uint64_t size,
                            int ebtype, double tol, double s, double norm,
                            void *d_out, void *stream);

/* ---- high level: mgard_x::compress / mgard_x::decompress ------------------
 * (include/compress_x.hpp:54-86,120-146; CompressionHighLevel.hpp:49-314,
 * 379-594).  `in` and `*out` may be host or device pointers; the output is
 * allocated in the memory space of the input (malloc / cudaMalloc) unless
 * output_pre_allocated, in which case *out_size carries the capacity in.
 * Writes/reads the reference's self-describing stream (Metadata.cpp). */
int mgb_compress(int ndim, int dtype, const uint64_t *shape, double tol,
                 double s, int ebtype, const void *in, void **out,
                 size_t *out_size, const void *const *coords,
                 const mgb_config *cfg, int output_pre_allocated);
int mgb_decompress(const void *in, size_t in_size, void **out,
                   const mgb_config *cfg, int output_pre_allocated,
                   int *ndim, uint64_t *shape, int *dtype);
/* Parse only the header (Metadata.cpp:475-739 infer_* helpers). */
int mgb_peek_header(const void *in, size_t in_size, int *ndim, uint64_t *shape,
                    int *dtype, int *ebtype, double *tol, double *s,
                    double *norm, uint64_t *header_bytes);
/* mgard_x::release_cache (compress_x.hpp:159) */
void mgb_release_cache(void);

/* sharded variant for one-process-per-GPU runs (no reference counterpart; the
 * reference processes sub-domains serially, GPUPipelines.hpp:88-207).  Each
 * rank compresses sub-domains [first, first+count) of the MaxDim partition
 * with the caller-supplied global norm; the container of `u64 size|payload`
 * records is written to d_out. */
int mgb_compress_subdomains(int ndim, int dtype, const uint64_t *shape,
                            double tol, double s, int ebtype, double norm,
                            const void *d_in_first, uint64_t first,
                            uint64_t count, const mgb_config *cfg,
                            uint8_t *d_out, uint64_t cap, uint64_t *size);
/* ---- one process per GPU: slab-sharded compress / decompress -----------------
 * (no reference counterpart - the reference walks the sub-domains
 * of DomainDecomposer.hpp:124-169 serially on one device, GPUPipelines.hpp:88-207, and
 * its published multi-GPU runs use one MPI rank per GPU outside the library).
 * The domain `shape` is MaxDim-decomposed along dim 0 in sub-domains of
 * cfg->domain_decomposition_size planes; process `rank` of `nranks` owns the
 * contiguous block of sub-domains mgb_owned_subdomains reports and passes them back to
 * back in `local` (host or device).  Collectives (NCCL, resolved with dlopen - the
 * library does not link it): one all-reduce of the per-sub-domain {max|u|, sum u^2}
 * pairs behind relative bounds, stream-ordered on the device, and one all-gather of
 * the container sizes.  The records written are those mgb_compress writes for the same
 * domain on one GPU, for any number of processes. */
typedef struct mgb_comm mgb_comm;
/* ncclGetUniqueId -> 128 bytes (rank 0; ship them to the other ranks), then
 * ncclCommInitRank on every rank.  nranks == 1 needs neither NCCL nor an id. */
int mgb_comm_unique_id(uint8_t *id128);
int mgb_comm_init_rank(const uint8_t *id128, int nranks, int rank, mgb_comm **comm);
/* wrap a communicator the caller already has (ncclComm_t passed as void*) */
int mgb_comm_from_nccl(void *nccl_comm, int nranks, int rank, mgb_comm **comm);
void mgb_comm_destroy(mgb_comm *comm);
int mgb_comm_rank(const mgb_comm *comm);
int mgb_comm_size(const mgb_comm *comm);
/* sub-domains [first, first + count) of `num_subdomains` owned by this process */
int mgb_owned_subdomains(const mgb_comm *comm, uint64_t num_subdomains,
                         uint64_t *first, uint64_t *count);
/* comm == NULL: single process.  out (host or device, capacity cap) receives this
 * process's `u64 size | payload` records; *local_size their bytes, *offset where they
 * start in the assembled stream (header + records of lower ranks), *total_size the
 * stream's size, all_sizes[nranks] every container's size, *norm the norm of the whole
 * domain (relative bounds), header[0..*header_size) the preamble + metadata (every
 * rank gets the same bytes; rank 0 writes them).  Optional outputs may be NULL.
 * Synchronous at return. */
int mgb_compress_sharded(mgb_comm *comm, int ndim, int dtype, const uint64_t *shape,
                         double tol, double s, int ebtype, const void *local,
                         const mgb_config *cfg, void *out, uint64_t cap,
                         uint64_t *local_size, uint64_t *offset, uint64_t *total_size,
                         uint64_t *all_sizes, double *norm, uint8_t *header,
                         uint64_t header_cap, uint64_t *header_size);
/* header: the stream's preamble + metadata (host); records: this process's records
 * (host or device); local_out: its sub-domains back to back (host or device).  No
 * exchange: every process decodes what it owns. */
int mgb_decompress_sharded(mgb_comm *comm, const uint8_t *header, uint64_t header_size,
                           const void *records, uint64_t records_size, void *local_out,
                           const mgb_config *cfg);
/* byte offset and size (8 + payload) of each record of an assembled stream: walks the
 * u64 chain (CompressionHighLevel.hpp:485-520); *count = number of sub-domains */
int mgb_stream_records(const void *stream, uint64_t size, uint64_t *offsets,
                       uint64_t *sizes, uint64_t cap, uint64_t *count);

/* Serialised metadata for the whole (decomposed) domain, host buffer. */
int mgb_write_header(int ndim, int dtype, const uint64_t *shape, double tol,
                     double s, int ebtype, double norm,
                     const void *const *coords, const mgb_config *cfg,
                     uint8_t *out, uint64_t cap, uint64_t *size);

/* Page-lock / query / release a host buffer so that the copies of the high-level
 * calls run at full PCIe speed: mgard_x::pin_memory / check_memory_pinned /
 * unpin_memory (reference include/compress_x.hpp:162-178,
 * CompressionHighLevel/DynamicAPI.cpp:606-740; cudaHostRegister underneath, as in
 * RuntimeX/DeviceAdapters/DeviceAdapterCuda.h MemoryManager::HostRegister). */
int mgb_pin_memory(void *ptr, uint64_t num_bytes);
int mgb_check_memory_pinned(const void *ptr);
int mgb_unpin_memory(void *ptr);

/* ---- MGARD-CPU convention -------------------------------------------------
 * mgard::compress / mgard::decompress (reference include/compress.hpp:33-72,
 * include/compress.tpp:35-83) computed on the GPU, bit-identical to the CPU
 * reference: same multilevel coefficients, same int64 quantisation, same zlib
 * payload and same protobuf header as a reference build without MGARD_ZSTD.
 * Any shape (dyadic or not, sizes of 1 allowed), uniform or explicit coordinates. */
typedef struct mgb_cpu_plan mgb_cpu_plan;

/* mgard::TensorMeshHierarchy<N, Real> (include/TensorMeshHierarchy.tpp:40-166).
 * coords: NULL for the uniform constructor, else ndim host arrays of dtype T. */
int mgb_cpu_plan_create(int ndim, const uint64_t *shape, int dtype,
                        const void *const *coords, mgb_cpu_plan **plan);
void mgb_cpu_plan_destroy(mgb_cpu_plan *plan);
/* TensorMeshHierarchy::L, ::ndof(l), ::shapes[l][dim] */
int mgb_cpu_plan_levels(const mgb_cpu_plan *plan);
uint64_t mgb_cpu_plan_ndof(const mgb_cpu_plan *plan, int level);
uint64_t mgb_cpu_plan_level_shape(const mgb_cpu_plan *plan, int level, int dim);

/* stage entry points, device resident (parity tests).  "Shuffled" is the level
 * order of mgard::shuffle (include/shuffle.tpp:8-37). */
int mgb_cpu_shuffle(mgb_cpu_plan *plan, const void *d_nodal, void *d_shuffled,
                    void *stream);
int mgb_cpu_unshuffle(mgb_cpu_plan *plan, const void *d_shuffled, void *d_nodal,
                      void *stream);
/* shuffle + mgard::decompose (include/decompose.tpp:129-174): nodal values ->
 * shuffled multilevel coefficients; and its inverse, recompose + unshuffle. */
int mgb_cpu_decompose(mgb_cpu_plan *plan, const void *d_nodal, void *d_coef,
                      void *stream);
int mgb_cpu_recompose(mgb_cpu_plan *plan, const void *d_coef, void *d_nodal,
                      void *stream);
/* One constituent operator of the multilevel transform on every line of level `level`
 * along user dimension `dim`, in place on a NODAL device array: ConstituentMassMatrix
 * (include/TensorMassMatrix.tpp:15-90), ConstituentMassMatrixInverse (:123-290),
 * ConstituentRestriction (include/TensorRestriction.tpp:24-71),
 * ConstituentProlongationAddition (include/TensorProlongation.tpp:22-69).  The reference
 * applies them to shuffled arrays; shuffle -> operator -> unshuffle is this call. */
enum {
  MGB_CPU_OP_MASS = 0,
  MGB_CPU_OP_MASS_INVERSE = 1,
  MGB_CPU_OP_RESTRICTION = 2,
  MGB_CPU_OP_PROLONGATION_ADDITION = 3
};
int mgb_cpu_apply_operator(mgb_cpu_plan *plan, int op, int level, int dim,
                           void *d_nodal, void *stream);
/* TensorMultilevelCoefficientQuantizer / Dequantizer on shuffled coefficients
 * (include/TensorMultilevelCoefficientQuantizer.tpp:13-77,242-259).  quantize
 * synchronises and returns MGB_FAILURE where the reference throws
 * std::domain_error ("number too large to be quantized"). */
int mgb_cpu_quantize(mgb_cpu_plan *plan, const void *d_coef, double s, double tol,
                     int64_t *d_q, void *stream);
int mgb_cpu_dequantize(mgb_cpu_plan *plan, const int64_t *d_q, double s,
                       double tol, void *d_coef, void *stream);

/* mgard::compress followed by CompressedDataset::write
 * (include/CompressedDataset.tpp:26-29): `in` is a host or device array; *out is
 * a malloc'ed host buffer holding preamble + header + payload (caller frees).
 * s = +inf selects the L-infinity norm; tol is absolute.  `compressor` is the
 * pb::Encoding::Compressor the reference picks at build time
 * (src/format.cpp:124-131): 1 = CPU_HUFFMAN_ZLIB (deflate level 9 of the int64
 * quanta, src/compressors.cpp:552-606), 2 = CPU_HUFFMAN_ZSTD (the reference's
 * default when libzstd is present: CPU Huffman coder + zstd level 1,
 * src/compressors.cpp:316-549; histogram and bit packing run on the GPU). */
int mgb_cpu_compress(int ndim, int dtype, const uint64_t *shape,
                     const void *const *coords, double s, double tol,
                     int compressor, const void *in, void **out,
                     size_t *out_size);
/* Preamble + header of an MGARD-CPU stream alone (host only, no device needed):
 * write_metadata (src/format.cpp:219-233) -- "MGARD", header size (u64) and
 * CRC32 (u32) BIG-endian (include/format.tpp:27-41), then the protobuf header of
 * populate_defaults / TensorMeshHierarchy::populate (src/format.cpp:102-140,
 * include/TensorMeshHierarchy.tpp:293-348). */
int mgb_cpu_write_header(int ndim, int dtype, const uint64_t *shape,
                         const void *const *coords, double s, double tol,
                         int compressor, uint8_t *out, uint64_t cap,
                         uint64_t *size);
/* mgard::decompress(void const *, std::size_t) (include/compress.hpp:62-72):
 * *out is a malloc'ed host array of the decoded dtype and shape. */
int mgb_cpu_decompress(const void *in, size_t in_size, void **out, int *ndim,
                       uint64_t *shape, int *dtype);

/* Tuning knobs (tests and A/B measurements; results never depend on them).
 * MGB_TUNE_SERIAL_MIN_CHUNKS: Huffman blocks with at least this many chunks are
 * encoded / decoded by the thread-per-chunk kernels, smaller ones by the
 * block-per-chunk kernels (0: always thread-per-chunk, negative: never).
 * MGB_TUNE_RING_DECODER: 1 (default): the thread-per-chunk decoder reads its bit
 * stream through a shared-memory ring; 0: the register-queue formulation.
 * MGB_TUNE_SUB_ENCODER: 1 (default): eight threads encode a chunk (when its size is a
 * multiple of 64 symbols); 0: one thread per chunk. */
enum { MGB_TUNE_SERIAL_MIN_CHUNKS = 0, MGB_TUNE_RING_DECODER = 1, MGB_TUNE_SUB_ENCODER = 2 };
int mgb_tune(int key, long long value);

/* kernel launch counter (bench.py's gpu_launches) */
uint64_t mgb_launch_count(void);
/* Per-kernel-family timing with CUDA events on the launching stream
 * (bench.py's roofline leg).  enable(1) clears and starts, enable(0) stops;
 * report(id) returns MGB_BAD_ARGUMENT past the last family. */
void mgb_profile_enable(int on);
int mgb_profile_report(int id, const char **name, unsigned long long *launches,
                       double *total_ms, double *max_ms);
const char *mgb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MGARD_B200_H */
