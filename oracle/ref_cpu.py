"""TEST INFRASTRUCTURE — ctypes binding of oracle/_ref/libmgard_cpu_ref.so.

The shared object is the UNMODIFIED reference MGARD-CPU (TensorMeshHierarchy,
shuffle, decompose/recompose, TensorMultilevelCoefficientQuantizer and the zlib
leg of src/compressors.cpp) built by oracle/Makefile from /root/reference; see
ref_cpu_wrap.cpp.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libmgard_cpu_ref.so")

INFO, DECOMPOSE, RECOMPOSE, QUANTIZE, DEQUANTIZE, SHUFFLE, UNSHUFFLE = range(7)


class RefCpuArgs(C.Structure):
    _fields_ = [
        ("op", C.c_int32), ("ndim", C.c_int32), ("dtype", C.c_int32), ("pad", C.c_int32),
        ("shape", C.POINTER(C.c_uint64)),
        ("coords", C.c_void_p * 4),
        ("s", C.c_double), ("tol", C.c_double),
        ("inp", C.c_void_p), ("out", C.c_void_p),
        ("L", C.c_uint64), ("ndof", C.c_uint64 * 64),
    ]


_lib = None


def available():
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_LIB_PATH)
        _lib.refcpu_run.argtypes = [C.POINTER(RefCpuArgs)]
        _lib.refcpu_run.restype = C.c_int
        _lib.refcpu_zlib_compress.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        _lib.refcpu_zlib_compress.restype = C.c_int64
        _lib.refcpu_zlib_decompress.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        _lib.refcpu_zlib_decompress.restype = None
    return _lib


def _run(op, shape, dtype, coords, inp, out, s=0.0, tol=0.0):
    a = RefCpuArgs()
    keep = []
    a.op, a.ndim = op, len(shape)
    a.dtype = 0 if np.dtype(dtype) == np.float32 else 1
    shp = (C.c_uint64 * len(shape))(*shape)
    a.shape = shp
    if coords is not None:
        for d, c in enumerate(coords):
            c = np.ascontiguousarray(c, dtype=dtype)
            keep.append(c)
            a.coords[d] = c.ctypes.data
    a.s, a.tol = float(s), float(tol)
    if inp is not None:
        a.inp = inp.ctypes.data
    if out is not None:
        a.out = out.ctypes.data
    rc = lib().refcpu_run(C.byref(a))
    if rc != 0:
        raise RuntimeError(f"refcpu_run failed: {rc}")
    return a


def info(shape, dtype=np.float64, coords=None):
    a = _run(INFO, shape, dtype, coords, None, None)
    return int(a.L), [int(a.ndof[l]) for l in range(int(a.L) + 1)]


def _map(op, u, coords, out_dtype=None, in_dtype=None, shape=None, real=None, **kw):
    shape = u.shape if shape is None else shape
    real = u.dtype if real is None else real
    u = np.ascontiguousarray(u)
    out = np.empty(int(np.prod(shape)), dtype=out_dtype or real)
    _run(op, shape, real, coords, u, out, **kw)
    return out


def decompose(u, coords=None):
    """nodal array -> shuffled multilevel coefficients (shuffle + decompose)."""
    return _map(DECOMPOSE, u, coords)


def recompose(c, shape, coords=None):
    return _map(RECOMPOSE, c, coords, shape=shape).reshape(shape)


def shuffle(u, coords=None):
    return _map(SHUFFLE, u, coords)


def unshuffle(c, shape, coords=None):
    return _map(UNSHUFFLE, c, coords, shape=shape).reshape(shape)


def quantize(c, shape, s, tol, coords=None):
    return _map(QUANTIZE, c, coords, out_dtype=np.int64, shape=shape, s=s, tol=tol)


def dequantize(q, shape, dtype, s, tol, coords=None):
    return _map(DEQUANTIZE, q, coords, out_dtype=dtype, shape=shape, real=dtype, s=s, tol=tol)


def preamble(header):
    """17 preamble bytes for `header` through the reference's SIGNATURE and
    serialize<> templates (include/format.hpp:28, include/format.tpp:27-41)."""
    out = (C.c_ubyte * 17)()
    hb = bytes(header)
    lib().refcpu_preamble(hb, C.c_uint64(len(hb)), out)
    return bytes(out)


def read_preamble(pre17):
    size, crc = C.c_uint64(0), C.c_uint32(0)
    rc = lib().refcpu_read_preamble(bytes(pre17[:17]), C.byref(size), C.byref(crc))
    if rc:
        raise ValueError("bad magic number")
    return size.value, crc.value


def zlib_compress(buf):
    buf = np.ascontiguousarray(buf).view(np.uint8).ravel()
    cap = buf.size + buf.size // 2 + 4096
    out = np.empty(cap, dtype=np.uint8)
    n = lib().refcpu_zlib_compress(buf.ctypes.data, buf.size, out.ctypes.data, cap)
    assert n >= 0
    return out[:n].copy()


def zlib_decompress(buf, nbytes):
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    out = np.empty(nbytes, dtype=np.uint8)
    lib().refcpu_zlib_decompress(buf.ctypes.data, buf.size, out.ctypes.data, nbytes)
    return out


# ---- CPU_HUFFMAN_ZSTD payload: src/compressors.cpp built with -DMGARD_ZSTD ----
_ZLIB_PATH = os.path.join(_HERE, "_ref", "libmgard_cpu_ref_zstd.so")
_zlib_lib = None


def zstd_available():
    return os.path.exists(_ZLIB_PATH)


def _zstd_lib():
    global _zlib_lib
    if _zlib_lib is None:
        _zlib_lib = C.CDLL(_ZLIB_PATH)
        _zlib_lib.refcpu_huffman_zstd_compress.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        _zlib_lib.refcpu_huffman_zstd_compress.restype = C.c_int64
        _zlib_lib.refcpu_huffman_zstd_decompress.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        _zlib_lib.refcpu_huffman_zstd_decompress.restype = None
    return _zlib_lib


def huffman_zstd_compress(q):
    """compress_memory_huffman (reference src/compressors.cpp:421-512)."""
    q = np.ascontiguousarray(q, dtype=np.int64)
    cap = q.nbytes + q.nbytes // 2 + (1 << 22)
    out = np.empty(cap, dtype=np.uint8)
    n = _zstd_lib().refcpu_huffman_zstd_compress(q.ctypes.data, q.size, out.ctypes.data, cap)
    assert n >= 0
    return out[:n].copy()


def huffman_zstd_decompress(buf, count):
    """decompress_memory_huffman (reference src/compressors.cpp:273-314)."""
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    out = np.empty(count, dtype=np.int64)
    _zstd_lib().refcpu_huffman_zstd_decompress(buf.ctypes.data, buf.size, out.ctypes.data, out.nbytes)
    return out
