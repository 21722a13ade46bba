"""Python mirror of the reference's MGARD-X interface for the hot path.

Names and argument meaning follow the reference:
  mgard_x::compress / decompress       include/compress_x.hpp:54-146
  mgard_x::Config                      include/mgard-x/Config/Config.h:10-42
  error_bound_type / data_type / compress_status_type
                                       include/mgard-x/Utilities/Types.h:30-63
  Hierarchy / Compressor (low level)   doc/MGARD-X.md:205-262 -> class Plan

Arrays may be numpy arrays (host) or torch CUDA tensors (device); the output
lives in the memory space of the input, as in the reference
(CompressionHighLevel.hpp:150-158).
"""
import ctypes as C
import enum
import math

import numpy as np

from . import _lib
from ._lib import MgardError, MgbConfig, check


class error_bound_type(enum.IntEnum):
    REL = 0
    ABS = 1


class data_type(enum.IntEnum):
    Float = 0
    Double = 1


class compress_status_type(enum.IntEnum):
    Success = 0
    Failure = 1
    OutputTooLargeFailure = 2
    NotSupportHigherNumberOfDimensionsFailure = 3
    NotSupportDataTypeFailure = 4
    BackendNotAvailableFailure = 5


class decomposition_type(enum.IntEnum):
    MultiDim = 0
    SingleDim = 1  # D <= 3
    Hybrid = 2     # not built


class lossless_type(enum.IntEnum):
    Huffman = 0
    Huffman_LZ4 = 1   # not built (nvcomp)
    Huffman_Zstd = 2
    CPU_Lossless = 3  # not built


class domain_decomposition_type(enum.IntEnum):
    MaxDim = 0
    Block = 1
    Variable = 2


class Config:
    """mgard_x::Config with its defaults (include/mgard-x/Config/Config.h:10-42,
    src/mgard-x/Config/Config.cpp:14-43).  Fields that select code this engine does
    not have (ZFP, Hybrid decomposition, LZ4, CPU threading, MDR knobs) are carried for
    source compatibility and checked by `_c()`: unsupported choices raise, they are
    never ignored."""

    UNLIMITED = (1 << 64) - 1

    def __init__(self):
        self.dev_type = "AUTO"
        self.dev_id = 0
        self.compressor = "MGARD"
        self.domain_decomposition = domain_decomposition_type.MaxDim
        self.decomposition = decomposition_type.MultiDim
        self.estimate_outlier_ratio = 1.0
        self.huff_dict_size = 8192
        self.huff_block_size = 1024 * 20
        self.lz4_block_size = 1 << 15
        self.zstd_compress_level = 3
        self.normalize_coordinates = True
        self.lossless = lossless_type.Huffman
        self.reorder = 0
        self.log_level = 0
        self.prefetch = False
        self.auto_pin_host_buffers = True
        self.max_larget_level = Config.UNLIMITED
        self.max_memory_footprint = Config.UNLIMITED
        self.total_num_bitplanes = 32
        self.block_size = 256
        self.domain_decomposition_dim = -1   # -1: the largest dimension (MaxDim)
        self.domain_decomposition_sizes = []
        self.mdr_adaptive_resolution = False
        self.adjust_shape = False
        self.compress_with_dryrun = False
        self.num_local_refactoring_level = 1
        self.auto_cache_release = False
        self.cpu_mode = "INTER_BLOCK"
        # mgard_b200 extension: planes per MaxDim sub-domain (0: from free device memory,
        # DomainDecomposer.hpp:199-230)
        self.domain_decomposition_size = 0

    def _c(self):
        c = MgbConfig()
        _lib.lib().mgb_config_default(C.byref(c))
        c.dev_id = self.dev_id
        c.huff_dict_size = self.huff_dict_size
        c.huff_block_size = self.huff_block_size
        c.domain_decomposition_dim = self.domain_decomposition_dim
        c.domain_decomposition_size = self.domain_decomposition_size
        c.normalize_coordinates = 1 if self.normalize_coordinates else 0
        c.lossless = int(self.lossless)
        c.zstd_compress_level = int(self.zstd_compress_level)
        c.reorder = int(self.reorder)
        c.decomposition = int(self.decomposition)
        if self.compressor != "MGARD" or self.dev_type not in ("AUTO", "CUDA") or not self.normalize_coordinates:
            raise MgardError(_lib.FAILURE, "Config: compressor / device / coordinate choice not supported")
        c.domain_decomposition = int(self.domain_decomposition)
        c.max_larget_level = int(self.max_larget_level)
        c.block_size = int(self.block_size)
        c.max_memory_footprint = int(self.max_memory_footprint)
        if self.domain_decomposition_sizes:
            self._sizes = (C.c_uint64 * len(self.domain_decomposition_sizes))(*self.domain_decomposition_sizes)
            c.domain_decomposition_sizes = self._sizes
            c.num_domain_decomposition_sizes = len(self.domain_decomposition_sizes)
        return c


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _dtype_code(dt):
    dt = np.dtype(dt)
    if dt == np.float32:
        return data_type.Float
    if dt == np.float64:
        return data_type.Double
    raise MgardError(_lib.BAD_DTYPE, "dtype")


def _np_dtype_of(x):
    if _is_torch(x):
        import torch
        return {torch.float32: np.float32, torch.float64: np.float64}.get(x.dtype)
    return x.dtype


def _coords_arg(coords, dtype, keep):
    if coords is None:
        return None
    arr = (C.c_void_p * len(coords))()
    for d, c in enumerate(coords):
        cc = np.ascontiguousarray(c, dtype=dtype)
        keep.append(cc)
        arr[d] = cc.ctypes.data
    return arr


def _shape_arg(shape):
    return (C.c_uint64 * len(shape))(*[int(s) for s in shape])


def pin_memory(array):
    """mgard_x::pin_memory (compress_x.hpp:162-167): page-lock a numpy array in place."""
    check(_lib.lib().mgb_pin_memory(array.ctypes.data, array.nbytes), "pin_memory")


def check_memory_pinned(array):
    return bool(_lib.lib().mgb_check_memory_pinned(array.ctypes.data))


def unpin_memory(array):
    check(_lib.lib().mgb_unpin_memory(array.ctypes.data), "unpin_memory")


TUNE_SERIAL_MIN_CHUNKS = 0
TUNE_RING_DECODER = 1
TUNE_SUB_ENCODER = 2


def tune(key, value):
    """mgb_tune: A/B knobs (results never depend on them)."""
    check(_lib.lib().mgb_tune(int(key), int(value)), "tune")


def launch_count():
    """Number of kernels this library has launched so far (bench.py gpu_launches)."""
    return int(_lib.lib().mgb_launch_count())


def release_cache():
    """mgard_x::release_cache (compress_x.hpp:159)."""
    _lib.lib().mgb_release_cache()


def adjust_shape(shape, config=None):
    """Config::adjust_shape (CompressionHighLevel/ShapeAdjustment.hpp:43-84): the prime
    factors of the largest extent, largest first, are handed to whichever dimension is
    currently the smallest; the data are reinterpreted with that shape (same bytes)."""
    shape = [int(n) for n in shape]
    steps = 1
    variable = config is not None and int(config.domain_decomposition) == 2
    if variable:
        steps = shape[0] // int(config.domain_decomposition_sizes[0])
        shape[0] = int(config.domain_decomposition_sizes[0])
    big = max(range(len(shape)), key=lambda d: (shape[d], -d))
    n, factors, z = shape[big], [], 2
    while z * z <= n:
        if n % z == 0:
            factors.append(z)
            n //= z
        else:
            z += 1
    if n > 1:
        factors.append(n)
    shape[big] = 1
    for f in reversed(factors):
        small = min(range(len(shape)), key=lambda d: (shape[d], d))
        shape[small] *= f
    if variable:
        shape[0] *= steps
    return tuple(shape)


def compress(data, tol, s, mode, coords=None, config=None, out=None):
    """mgard_x::compress(D, dtype, shape, tol, s, mode, original_data,
    compressed_data, compressed_size[, coords][, config], output_pre_allocated).

    data: numpy array or torch CUDA tensor (float32 / float64, 1-5 dims).
    out:  optional pre-allocated uint8 buffer (same memory space rules as the
          reference); its size is the capacity.
    Returns the compressed stream (numpy uint8 array or torch uint8 tensor).
    """
    L = _lib.lib()
    cfg = (config or Config())._c()
    keep = []
    tdev = _is_torch(data)
    if tdev:
        import torch
        if not data.is_cuda:
            data = data.numpy()
            tdev = False
    if tdev:
        import torch
        data = data.contiguous()
        npdt = _np_dtype_of(data)
        if npdt is None:
            raise MgardError(_lib.BAD_DTYPE, "compress")
        shape = tuple(data.shape)
        in_ptr = data.data_ptr()
        cfg.dev_id = data.device.index or 0
    else:
        data = np.ascontiguousarray(data)
        npdt = data.dtype
        shape = data.shape
        in_ptr = data.ctypes.data
    if config is not None and config.adjust_shape:
        if coords is not None:
            raise MgardError(_lib.BAD_ARGUMENT, "adjust_shape with explicit coordinates")
        shape = adjust_shape(shape, config)
    dt = _dtype_code(npdt)
    carr = _coords_arg(coords, npdt, keep)
    outp = C.c_void_p(0)
    size = C.c_size_t(0)
    pre = 0
    if out is not None:
        pre = 1
        if _is_torch(out):
            outp = C.c_void_p(out.data_ptr())
            size = C.c_size_t(out.numel())
        else:
            outp = C.c_void_p(out.ctypes.data)
            size = C.c_size_t(out.size)
    rc = L.mgb_compress(len(shape), int(dt), _shape_arg(shape), float(tol), float(s),
                        int(mode), in_ptr, C.byref(outp), C.byref(size), carr,
                        C.byref(cfg), pre)
    check(rc, "mgard_x::compress")
    n = size.value
    if out is not None:
        return out[:n]
    if tdev:
        import torch
        # adopt the cudaMalloc'ed buffer: copy into a torch tensor, then free
        res = torch.empty(n, dtype=torch.uint8, device=data.device)
        _cuda_memcpy(res.data_ptr(), outp.value, n)
        _cuda_free(outp.value)
        return res
    buf = (C.c_uint8 * n).from_address(outp.value)
    res = np.frombuffer(buf, dtype=np.uint8).copy()
    _libc_free(outp.value)
    return res


def decompress(stream, config=None, out=None):
    """mgard_x::decompress(compressed_data, compressed_size, decompressed_data,
    shape, dtype[, config], output_pre_allocated) — returns the array."""
    L = _lib.lib()
    cfg = (config or Config())._c()
    tdev = _is_torch(stream) and stream.is_cuda
    if _is_torch(stream) and not tdev:
        stream = stream.numpy()
    if tdev:
        stream = stream.contiguous()
        in_ptr, nbytes = stream.data_ptr(), stream.numel()
        cfg.dev_id = stream.device.index or 0
    else:
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        in_ptr, nbytes = stream.ctypes.data, stream.size
    info = peek_header(stream)
    npdt = np.float32 if info["dtype"] == data_type.Float else np.float64
    shape = info["shape"]
    outp = C.c_void_p(0)
    pre = 0
    if out is None:
        if tdev:
            import torch
            out = torch.empty(shape, dtype=torch.float32 if npdt == np.float32 else torch.float64,
                              device=stream.device)
        else:
            out = np.empty(shape, dtype=npdt)
    pre = 1
    outp = C.c_void_p(out.data_ptr() if _is_torch(out) else out.ctypes.data)
    nd = C.c_int(0)
    dtc = C.c_int(0)
    shp = (C.c_uint64 * 5)()
    rc = L.mgb_decompress(in_ptr, nbytes, C.byref(outp), C.byref(cfg), pre, C.byref(nd),
                          shp, C.byref(dtc))
    check(rc, "mgard_x::decompress")
    return out


def peek_header(stream):
    """Shape / dtype / error control stored in a stream (Metadata.cpp:475-739)."""
    L = _lib.lib()
    if _is_torch(stream):
        ptr, n = stream.data_ptr(), stream.numel()
    else:
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        ptr, n = stream.ctypes.data, stream.size
    nd, dt, eb = C.c_int(0), C.c_int(0), C.c_int(0)
    shp = (C.c_uint64 * 5)()
    tol, s, norm = C.c_double(0), C.c_double(0), C.c_double(0)
    hb = C.c_uint64(0)
    rc = L.mgb_peek_header(ptr, n, C.byref(nd), shp, C.byref(dt), C.byref(eb),
                           C.byref(tol), C.byref(s), C.byref(norm), C.byref(hb))
    check(rc, "peek_header")
    return dict(shape=tuple(int(shp[i]) for i in range(nd.value)),
                dtype=data_type(dt.value), mode=error_bound_type(eb.value),
                tol=tol.value, s=s.value, norm=norm.value, header_bytes=hb.value)


_cudart = None


def _rt():
    global _cudart
    if _cudart is None:
        import torch  # noqa: F401  (loads libcudart into the process)
        for name in ("libcudart.so.12", "libcudart.so"):
            try:
                _cudart = C.CDLL(name)
                break
            except OSError:
                continue
        if _cudart is None:
            raise RuntimeError("libcudart not found")
    return _cudart


def _cuda_memcpy(dst, src, n):
    rc = _rt().cudaMemcpy(C.c_void_p(dst), C.c_void_p(src), C.c_size_t(n), 4)
    if rc != 0:
        raise RuntimeError(f"cudaMemcpy failed: {rc}")


def _cuda_free(p):
    _rt().cudaFree(C.c_void_p(p))


def _libc_free(p):
    C.CDLL(None).free(C.c_void_p(p))


class Plan:
    """Hierarchy<D,T> + Compressor<D,T> of the reference's low-level API
    (doc/MGARD-X.md:205-262), device resident.  All arrays are torch CUDA
    tensors; used by the parity tests and the device-resident bench."""

    def __init__(self, shape, dtype, coords=None, config=None):
        import torch
        self.torch = torch
        L = _lib.lib()
        self.shape = tuple(int(s) for s in shape)
        self.np_dtype = np.dtype(dtype)
        self.t_dtype = torch.float32 if self.np_dtype == np.float32 else torch.float64
        self.config = config or Config()
        cfg = self.config._c()
        keep = []
        carr = _coords_arg(coords, self.np_dtype, keep)
        h = C.c_void_p(0)
        rc = L.mgb_plan_create(len(self.shape), _shape_arg(self.shape),
                               int(_dtype_code(self.np_dtype)), carr, C.byref(cfg),
                               C.byref(h))
        check(rc, "Hierarchy")
        self._h = h
        self.n = int(L.mgb_plan_num_elems(h))
        self.l_target = int(L.mgb_plan_l_target(h))
        self.dict_size = self.config.huff_dict_size

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().mgb_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def set_generic(self, on=True):
        """Use the dimension-generic kernels for D == 3 (testing aid)."""
        _lib.lib().mgb_plan_set_generic(self._h, 1 if on else 0)

    def level_shape(self, l):
        L = _lib.lib()
        return tuple(int(L.mgb_plan_level_shape(self._h, l, d)) for d in range(len(self.shape)))

    def table(self, which, level, dim):
        L = _lib.lib()
        code = {"dist": 0, "ratio": 1, "am": 2, "bm": 3}[which]
        n = int(L.mgb_plan_table(self._h, code, level, dim, None, 0))
        out = np.zeros(n, dtype=self.np_dtype)
        L.mgb_plan_table(self._h, code, level, dim, out.ctypes.data, n)
        return out

    def _stream(self):
        return self.torch.cuda.current_stream().cuda_stream

    def decompose(self, u):
        out = self.torch.empty_like(u)
        check(_lib.lib().mgb_decompose(self._h, u.data_ptr(), out.data_ptr(), self._stream()),
              "decompose")
        return out

    def recompose(self, v):
        out = self.torch.empty_like(v)
        check(_lib.lib().mgb_recompose(self._h, v.data_ptr(), out.data_ptr(), self._stream()),
              "recompose")
        return out

    def norm(self, u, s):
        r = C.c_double(0)
        self.torch.cuda.current_stream().synchronize()
        check(_lib.lib().mgb_norm(self._h, u.data_ptr(), float(s), C.byref(r)), "norm")
        return r.value

    def quantize(self, coef, mode, tol, s, norm, outlier_cap=None):
        t = self.torch
        cap = outlier_cap or max(self.n // 8, 4096)
        sym = t.empty(self.n, dtype=t.int16, device=coef.device)
        hist = t.empty(self.dict_size, dtype=t.int32, device=coef.device)
        ocount = t.zeros(1, dtype=t.int64, device=coef.device)
        oidx = t.empty(cap, dtype=t.int64, device=coef.device)
        oval = t.empty(cap, dtype=t.int64, device=coef.device)
        check(_lib.lib().mgb_quantize(self._h, coef.data_ptr(), int(mode), float(tol), float(s),
                                      float(norm), sym.data_ptr(), hist.data_ptr(),
                                      ocount.data_ptr(), oidx.data_ptr(), oval.data_ptr(),
                                      cap, self._stream()), "quantize")
        k = int(ocount.item())
        if k > cap:
            return self.quantize(coef, mode, tol, s, norm, outlier_cap=k)
        return sym, hist, oidx[:k], oval[:k]

    def dequantize(self, sym, oidx, oval, mode, tol, s, norm):
        t = self.torch
        out = t.empty(self.shape, dtype=self.t_dtype, device=sym.device)
        k = int(oidx.numel())
        check(_lib.lib().mgb_dequantize(self._h, sym.data_ptr(), k,
                                        oidx.data_ptr() if k else None,
                                        oval.data_ptr() if k else None, int(mode),
                                        float(tol), float(s), float(norm), out.data_ptr(),
                                        self._stream()), "dequantize")
        return out

    def codebook(self, hist):
        t = self.torch
        cb = t.empty(self.dict_size, dtype=t.int64, device=hist.device)
        db = t.empty(128 + self.dict_size, dtype=t.int64, device=hist.device)
        check(_lib.lib().mgb_codebook(self._h, hist.data_ptr(), cb.data_ptr(), db.data_ptr(),
                                      self._stream()), "codebook")
        return cb, db

    def huffman_compress(self, sym, hist, oidx, oval, cap=None):
        t = self.torch
        n = int(sym.numel())
        cap = cap or (n * 8 + 8 * (128 + self.dict_size) + 32 * (n // self.config.huff_block_size + 2) + 4096 + 16 * int(oidx.numel()))
        out = t.empty(cap, dtype=t.uint8, device=sym.device)
        size = C.c_uint64(0)
        k = int(oidx.numel())
        check(_lib.lib().mgb_huffman_compress(self._h, sym.data_ptr(), n, hist.data_ptr(), k,
                                              oidx.data_ptr() if k else None,
                                              oval.data_ptr() if k else None,
                                              out.data_ptr(), cap, C.byref(size),
                                              self._stream()), "huffman_compress")
        return out[:size.value]

    def huffman_decompress(self, payload, n):
        t = self.torch
        sym = t.empty(n, dtype=t.int16, device=payload.device)
        oc = C.c_uint64(0)
        pi, pv = C.c_void_p(0), C.c_void_p(0)
        check(_lib.lib().mgb_huffman_decompress(self._h, payload.data_ptr(), payload.numel(),
                                                sym.data_ptr(), n, C.byref(oc), C.byref(pi),
                                                C.byref(pv), self._stream()),
              "huffman_decompress")
        k = oc.value
        base = payload.data_ptr()
        if k:
            o0 = pi.value - base
            raw = payload[o0:o0 + 16 * k].clone()
            oidx = raw[:8 * k].view(t.int64)
            oval = raw[8 * k:].view(t.int64)
        else:
            oidx = t.empty(0, dtype=t.int64, device=payload.device)
            oval = t.empty(0, dtype=t.int64, device=payload.device)
        return sym, oidx, oval

    def compress(self, u, mode, tol, s, norm=1.0, cap=None, out=None):
        """Compressor::Compress: returns (payload tensor, norm)."""
        t = self.torch
        if out is None:
            cap = cap or (self.n * self.np_dtype.itemsize + 8 * (128 + self.dict_size) + (1 << 20))
            out = t.empty(cap, dtype=t.uint8, device=u.device)
        cap = out.numel()
        size = C.c_uint64(0)
        nrm = C.c_double(norm)
        check(_lib.lib().mgb_compress_lowlevel(self._h, u.data_ptr(), int(mode), float(tol),
                                               float(s), C.byref(nrm), out.data_ptr(), cap,
                                               C.byref(size), self._stream()),
              "Compressor::Compress")
        return out[:size.value], nrm.value

    def decompress(self, payload, mode, tol, s, norm, out=None):
        t = self.torch
        if out is None:
            out = t.empty(self.shape, dtype=self.t_dtype, device=payload.device)
        check(_lib.lib().mgb_decompress_lowlevel(self._h, payload.data_ptr(), payload.numel(),
                                                 int(mode), float(tol), float(s), float(norm),
                                                 out.data_ptr(), self._stream()),
              "Compressor::Decompress")
        return out
