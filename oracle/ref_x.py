"""TEST INFRASTRUCTURE — ctypes binding of oracle/_ref/libmgardx_ref.so.

The shared object is the UNMODIFIED reference MGARD-X (SERIAL adapter) built by
oracle/Makefile from /root/reference; see ref_x_wrap.cpp / ref_x_wrap.h.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libmgardx_ref.so")

REL, ABS = 0, 1
OP_TABLES, OP_DECOMPOSE, OP_RECOMPOSE, OP_COMPRESS, OP_DECOMPRESS = range(5)


class RefxArgs(C.Structure):
    _fields_ = [
        ("op", C.c_int32), ("ndim", C.c_int32), ("dtype", C.c_int32),
        ("ebtype", C.c_int32), ("s_is_inf", C.c_int32),
        ("dict_size", C.c_int32), ("chunk_size", C.c_int32),
        ("l_target", C.c_int32),
        ("shape", C.POINTER(C.c_uint64)),
        ("coords", C.c_void_p * 5),
        ("data", C.c_void_p),
        ("tol", C.c_double), ("s", C.c_double), ("norm", C.c_double),
        ("tables_out", C.c_void_p), ("tables_count", C.c_uint64),
        ("decomposed_out", C.c_void_p),
        ("quantized_out", C.c_void_p),
        ("outlier_count", C.c_uint64),
        ("payload", C.c_void_p), ("payload_cap", C.c_uint64),
        ("payload_size", C.c_uint64),
        ("lossless", C.c_int32), ("zstd_level", C.c_int32),
        ("reorder", C.c_int32), ("decomposition", C.c_int32),
        ("max_level", C.c_int32),
    ]


_lib = None


def available():
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_LIB_PATH)
        _lib.refx_run.argtypes = [C.POINTER(RefxArgs)]
        _lib.refx_run.restype = C.c_int
    return _lib


def _base_args(shape, dtype, coords, dict_size, chunk_size, keep):
    a = RefxArgs()
    a.ndim = len(shape)
    a.dtype = 0 if np.dtype(dtype) == np.float32 else 1
    shp = (C.c_uint64 * len(shape))(*shape)
    keep.append(shp)
    a.shape = shp
    a.dict_size = dict_size
    a.chunk_size = chunk_size
    if coords is not None:
        for d, c in enumerate(coords):
            cc = np.ascontiguousarray(c, dtype=dtype)
            keep.append(cc)
            a.coords[d] = cc.ctypes.data
    return a


def _set_s(a, s):
    if np.isinf(s):
        a.s_is_inf, a.s = 1, 0.0
    else:
        a.s_is_inf, a.s = 0, float(s)


def level_shapes(shape):
    """Hierarchy.hpp:199-230."""
    per_dim = []
    for n in shape:
        sizes = []
        while n > 2:
            sizes.append(n)
            n = n // 2 + 1
        sizes.append(2)
        per_dim.append(sizes)
    L = min(len(s) for s in per_dim) - 1
    return [[per_dim[d][L - l] for d in range(len(shape))] for l in range(L + 1)]


def tables(shape, dtype, coords=None):
    """Per (level, dim): dict(dist, ratio, am, bm) as the reference builds them."""
    keep = []
    a = _base_args(shape, dtype, coords, 8192, 20480, keep)
    a.op = OP_TABLES
    ls = level_shapes(shape)
    total = sum(4 * n + 2 for lv in ls for n in lv)
    out = np.zeros(total, dtype=dtype)
    a.tables_out = out.ctypes.data
    rc = lib().refx_run(C.byref(a))
    assert rc == 0 and a.tables_count == total
    res, off = [], 0
    for lv in ls:
        row = []
        for n in lv:
            t = {}
            for name, m in (("dist", n), ("ratio", n), ("am", n + 1), ("bm", n + 1)):
                t[name] = out[off:off + m].copy()
                off += m
            row.append(t)
        res.append(row)
    return res


def decompose(u, coords=None, decomposition=0):
    keep = []
    v = np.array(u, copy=True, order="C")
    a = _base_args(v.shape, v.dtype, coords, 8192, 20480, keep)
    a.op = OP_DECOMPOSE
    a.decomposition = decomposition
    a.data = v.ctypes.data
    assert lib().refx_run(C.byref(a)) == 0
    return v


def recompose(v, coords=None, decomposition=0):
    keep = []
    u = np.array(v, copy=True, order="C")
    a = _base_args(u.shape, u.dtype, coords, 8192, 20480, keep)
    a.op = OP_RECOMPOSE
    a.decomposition = decomposition
    a.data = u.ctypes.data
    assert lib().refx_run(C.byref(a)) == 0
    return u


def compress(u, ebtype, tol, s, coords=None, dict_size=8192, chunk_size=20480, lossless=0, reorder=0, decomposition=0,
             max_level=0):
    """Low-level Compressor::Compress staged; returns dict with payload bytes,
    norm, decomposed coefficients, quantized (dict-shifted) int64, outlier count."""
    keep = []
    v = np.array(u, copy=True, order="C")
    a = _base_args(v.shape, v.dtype, coords, dict_size, chunk_size, keep)
    a.op = OP_COMPRESS
    a.max_level = max_level
    a.lossless = lossless
    a.reorder = reorder
    a.decomposition = decomposition
    a.ebtype = ebtype
    a.tol = tol
    _set_s(a, s)
    a.data = v.ctypes.data
    dec = np.zeros_like(v)
    q = np.zeros(v.shape, dtype=np.int64)
    cap = v.nbytes * 3 + (1 << 20)
    payload = np.zeros(cap, dtype=np.uint8)
    a.decomposed_out = dec.ctypes.data
    a.quantized_out = q.ctypes.data
    a.payload = payload.ctypes.data
    a.payload_cap = cap
    rc = lib().refx_run(C.byref(a))
    assert rc == 0, rc
    return dict(payload=payload[:a.payload_size].copy(), norm=a.norm,
                decomposed=dec, quantized=q, outlier_count=a.outlier_count,
                l_target=a.l_target)


def decompress(payload, shape, dtype, ebtype, tol, s, norm, coords=None,
               dict_size=8192, chunk_size=20480, lossless=0, reorder=0, decomposition=0, max_level=0):
    keep = []
    out = np.zeros(shape, dtype=dtype)
    a = _base_args(shape, dtype, coords, dict_size, chunk_size, keep)
    a.op = OP_DECOMPRESS
    a.max_level = max_level
    a.lossless = lossless
    a.reorder = reorder
    a.decomposition = decomposition
    a.ebtype = ebtype
    a.tol = tol
    a.norm = norm
    _set_s(a, s)
    a.data = out.ctypes.data
    p = np.ascontiguousarray(payload, dtype=np.uint8)
    a.payload = p.ctypes.data
    a.payload_size = p.size
    assert lib().refx_run(C.byref(a)) == 0
    return out


def huffman_compress(symbols, dict_size=8192, chunk_size=20480):
    sym = np.ascontiguousarray(symbols, dtype=np.uint64)
    cap = sym.nbytes * 2 + (1 << 20)
    out = np.zeros(cap, dtype=np.uint8)
    size = C.c_uint64(0)
    f = lib().refx_huffman_compress
    f.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p,
                  C.c_uint64, C.POINTER(C.c_uint64)]
    rc = f(sym.ctypes.data, sym.size, dict_size, chunk_size, out.ctypes.data,
           cap, C.byref(size))
    assert rc == 0
    return out[:size.value].copy()


def huffman_decompress(payload, n):
    p = np.ascontiguousarray(payload, dtype=np.uint8)
    out = np.zeros(n, dtype=np.uint64)
    f = lib().refx_huffman_decompress
    f.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    assert f(p.ctypes.data, p.size, out.ctypes.data, n) == 0
    return out


def codebook(freq):
    fr = np.ascontiguousarray(freq, dtype=np.uint32)
    d = fr.size
    cb = np.zeros(d, dtype=np.uint64)
    db = np.zeros(8 * 128 + 8 * d, dtype=np.uint8)
    cl = np.zeros(d, dtype=np.uint32)
    nz = C.c_int(0)
    f = lib().refx_codebook
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                  C.POINTER(C.c_int)]
    assert f(fr.ctypes.data, d, cb.ctypes.data, db.ctypes.data, cl.ctypes.data,
             C.byref(nz)) == 0
    return dict(codebook=cb, first=db[:512].view(np.uint64).copy(),
                entry=db[512:1024].view(np.uint64).copy(),
                keys=db[1024:].view(np.uint64).copy(), cl=cl[:nz.value].copy())
