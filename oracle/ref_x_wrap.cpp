/* TEST INFRASTRUCTURE — not part of the product path.
 *
 * C-ABI wrapper around the UNMODIFIED reference MGARD-X low-level API
 * (SERIAL device adapter), compiled in place from /root/reference by
 * oracle/Makefile into oracle/_ref/libmgardx_ref.so.  It exists so that the
 * numpy restatement in oracle/mgardx_oracle.py and the CUDA engine can be
 * checked against what the reference itself computes, stage by stage.
 *
 * One translation unit per (D, T): compile with -DREFX_D=<1..5>
 * -DREFX_T=<float|double> -DREFX_NAME=refx_run_<D><f|d>.
 *
 * Reference entry points used (all under /root/reference/include/mgard-x):
 *   Hierarchy/Hierarchy.hpp:193-418          Hierarchy::init (level tables)
 *   CompressionLowLevel/Compressor.hpp:121-272  Compressor stages
 *   DataRefactoring/DataRefactor.hpp:73-140  Decompose / Recompose
 */
#include "compress_x_lowlevel.hpp"

#include "mgard-x/DataRefactoring/MultiDimension/Coefficient/CalcCoefficients3D.hpp"
#include "mgard-x/DataRefactoring/MultiDimension/Coefficient/CalcCoefficientsND.hpp"
#include "mgard-x/DataRefactoring/MultiDimension/Coefficient/CoefficientsRestore3D.hpp"
#include "mgard-x/DataRefactoring/MultiDimension/Coefficient/CoefficientsRestoreND.hpp"
#include "mgard-x/DataRefactoring/MultiDimension/CopyND/AddND.hpp"
#include "mgard-x/DataRefactoring/MultiDimension/CopyND/CopyND.hpp"
#include "mgard-x/DataRefactoring/MultiDimension/CopyND/SubtractND.hpp"
#include "mgard-x/DataRefactoring/MultiDimension/Correction/CalcCorrection3D.hpp"
#include "mgard-x/DataRefactoring/MultiDimension/Correction/CalcCorrectionND.hpp"
#include "mgard-x/DataRefactoring/MultiDimension/DataRefactoring.hpp"
#include "mgard-x/DataRefactoring/SingleDimension/Coefficient/CalcCoefficients.hpp"
#include "mgard-x/DataRefactoring/SingleDimension/Coefficient/CoefficientsRestore.hpp"
#include "mgard-x/DataRefactoring/SingleDimension/Correction/CalcCorrection.hpp"
#include "mgard-x/DataRefactoring/SingleDimension/DataRefactoring.hpp"

#include "ref_x_wrap.h"

#include <cstring>
#include <limits>
#include <vector>

using namespace mgard_x;
using Dev = SERIAL;
typedef REFX_T T;
constexpr DIM D = REFX_D;

static Config make_config(const refx_args *a) {
  Config cfg;
  cfg.dev_type = device_type::SERIAL;
  cfg.lossless = a->lossless == 2 ? lossless_type::Huffman_Zstd : lossless_type::Huffman;
  if (a->zstd_level)
    cfg.zstd_compress_level = a->zstd_level;
  cfg.reorder = a->reorder;
  cfg.decomposition = a->decomposition == 1 ? decomposition_type::SingleDim : decomposition_type::MultiDim;
  cfg.huff_dict_size = a->dict_size;
  cfg.huff_block_size = a->chunk_size;
  cfg.normalize_coordinates = true;
  cfg.log_level = log::ERR;
  if (a->max_level > 0)
    cfg.max_larget_level = a->max_level;
  return cfg;
}

extern "C" int REFX_NAME(refx_args *a) {
  Config cfg = make_config(a);
  std::vector<SIZE> shape(a->shape, a->shape + D);
  size_t n = 1;
  for (DIM d = 0; d < D; d++)
    n *= shape[d];

  Hierarchy<D, T, Dev> *hp;
  if (a->coords[0] != nullptr) {
    std::vector<T *> coords(D);
    for (DIM d = 0; d < D; d++)
      coords[d] = (T *)a->coords[d];
    hp = new Hierarchy<D, T, Dev>(shape, coords, cfg);
  } else {
    hp = new Hierarchy<D, T, Dev>(shape, cfg);
  }
  Hierarchy<D, T, Dev> &h = *hp;
  a->l_target = h.l_target();
  T s = a->s_is_inf ? std::numeric_limits<T>::infinity() : (T)a->s;
  error_bound_type eb =
      a->ebtype == 0 ? error_bound_type::REL : error_bound_type::ABS;
  int rc = 0;

  if (a->op == REFX_OP_TABLES) {
    /* for l, d: dist[n] ratio[n] am[n+1] bm[n+1]; then volumes (L+1,D,maxn) */
    T *out = (T *)a->tables_out;
    size_t off = 0;
    for (SIZE l = 0; l <= h.l_target(); l++)
      for (DIM d = 0; d < D; d++) {
        SIZE m = h.level_shape(l, d);
        memcpy(out + off, h.dist(l, d).hostCopy(), m * sizeof(T));
        off += m;
        memcpy(out + off, h.ratio(l, d).hostCopy(), m * sizeof(T));
        off += m;
        memcpy(out + off, h.am(l, d).hostCopy(), (m + 1) * sizeof(T));
        off += m + 1;
        memcpy(out + off, h.bm(l, d).hostCopy(), (m + 1) * sizeof(T));
        off += m + 1;
      }
    a->tables_count = off;
  } else if (a->op == REFX_OP_DECOMPOSE || a->op == REFX_OP_RECOMPOSE) {
    Compressor<D, T, Dev> c(h, cfg);
    Array<D, T, Dev> arr(shape);
    arr.load((T *)a->data);
    if (a->op == REFX_OP_DECOMPOSE)
      c.Decompose(arr, 0);
    else
      c.Recompose(arr, 0);
    DeviceRuntime<Dev>::SyncQueue(0);
    memcpy(a->data, arr.hostCopy(), n * sizeof(T));
  } else if (a->op == REFX_OP_COMPRESS) {
    Compressor<D, T, Dev> c(h, cfg);
    Array<D, T, Dev> arr(shape);
    arr.load((T *)a->data);
    Array<1, Byte, Dev> out;
    T norm = (T)a->norm;
    T tol = (T)a->tol;
    c.CalculateNorm(arr, eb, s, norm, 0);
    c.Decompose(arr, 0);
    DeviceRuntime<Dev>::SyncQueue(0);
    if (a->decomposed_out)
      memcpy(a->decomposed_out, arr.hostCopy(), n * sizeof(T));
    c.Quantize(arr, eb, tol, s, norm, 0);
    DeviceRuntime<Dev>::SyncQueue(0);
    if (a->quantized_out)
      memcpy(a->quantized_out, c.quantized_array.hostCopy(),
             n * sizeof(int64_t));
    a->outlier_count = c.lossless_compressor.huffman.outlier_count;
    c.LosslessCompress(out, 0);
    c.Serialize(out, 0);
    DeviceRuntime<Dev>::SyncQueue(0);
    a->norm = (double)norm;
    a->payload_size = out.shape(0);
    if (out.shape(0) <= a->payload_cap)
      memcpy(a->payload, out.hostCopy(), out.shape(0));
    else
      rc = 2;
  } else if (a->op == REFX_OP_DECOMPRESS) {
    Compressor<D, T, Dev> c(h, cfg);
    Array<1, Byte, Dev> in({(SIZE)a->payload_size});
    in.load(a->payload);
    Array<D, T, Dev> arr(shape);
    T norm = (T)a->norm;
    c.Decompress(in, eb, (T)a->tol, s, norm, arr, 0);
    DeviceRuntime<Dev>::SyncQueue(0);
    memcpy(a->data, arr.hostCopy(), n * sizeof(T));
  } else {
    rc = 1;
  }
  delete hp;
  return rc;
}
