"""Python mirror of the reference's MGARD-CPU interface, computed on the GPU.

Names and argument meaning follow the reference:
  mgard::TensorMeshHierarchy<N, Real>   include/TensorMeshHierarchy.hpp:30-200
  mgard::compress(hierarchy, v, s, tolerance)      include/compress.hpp:33-47
  mgard::decompress(data, size)                    include/compress.hpp:62-72
  shuffle / unshuffle                              include/shuffle.hpp
  decompose / recompose                            include/decompose.hpp
  TensorMultilevelCoefficientQuantizer/Dequantizer include/TensorMultilevelCoefficientQuantizer.hpp

`compress` returns the bytes `CompressedDataset::write` would emit (preamble,
protobuf header, zlib payload); they are bit-identical to the CPU reference's.
Stage functions take and return torch CUDA tensors.  There is no CPU fallback.
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import check

_LIBC = C.CDLL(None)
_LIBC.free.argtypes = [C.c_void_p]


def _dtype_code(dt):
    dt = np.dtype(dt)
    if dt == np.float32:
        return 0
    if dt == np.float64:
        return 1
    raise TypeError("MGARD supports float32 and float64")


def _coord_args(coords, dt, keep):
    if coords is None:
        return None
    arr = (C.c_void_p * len(coords))()
    for d, c in enumerate(coords):
        c = np.ascontiguousarray(c, dtype=dt)
        keep.append(c)
        arr[d] = c.ctypes.data
    return arr


class TensorMeshHierarchy:
    """mgard::TensorMeshHierarchy<N, Real>: `shape` slowest dimension first;
    `coordinates` None (uniform on [0, 1]) or one increasing array per dimension."""

    def __init__(self, shape, coordinates=None, dtype=np.float64):
        self.shape = tuple(int(n) for n in shape)
        self.dtype = np.dtype(dtype)
        self.coordinates = None if coordinates is None else [np.asarray(c, dtype=dtype) for c in coordinates]
        if self.coordinates is not None:
            for n, c in zip(self.shape, self.coordinates):
                if c.shape != (n,):
                    raise ValueError("incorrect number of node coordinates given")
        keep = []
        shp = (C.c_uint64 * len(self.shape))(*self.shape)
        h = C.c_void_p()
        check(_lib.lib().mgb_cpu_plan_create(len(self.shape), shp, _dtype_code(dtype),
                                             _coord_args(self.coordinates, dtype, keep), C.byref(h)),
              "TensorMeshHierarchy")
        self._h = h
        self.L = _lib.lib().mgb_cpu_plan_levels(h)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            _lib.lib().mgb_cpu_plan_destroy(h)
            self._h = None

    def ndof(self, l=None):
        return int(_lib.lib().mgb_cpu_plan_ndof(self._h, self.L if l is None else l))

    def level_shape(self, l):
        return tuple(int(_lib.lib().mgb_cpu_plan_level_shape(self._h, l, d)) for d in range(len(self.shape)))

    # ---- stages (torch CUDA tensors) ----
    def _stage(self, fn, x, out_dtype=None, extra=()):
        import torch
        if not x.is_cuda:
            raise ValueError("stage functions take CUDA tensors")
        x = x.contiguous()
        out = torch.empty(self.ndof(), dtype=out_dtype or x.dtype, device=x.device)
        st = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            check(fn(self._h, x.data_ptr(), *extra, out.data_ptr(), st), fn.__name__)
        return out

    def shuffle(self, v):
        return self._stage(_lib.lib().mgb_cpu_shuffle, v)

    def unshuffle(self, u):
        return self._stage(_lib.lib().mgb_cpu_unshuffle, u).reshape(self.shape)

    MASS, MASS_INVERSE, RESTRICTION, PROLONGATION_ADDITION = range(4)

    def apply_operator(self, op, level, dimension, v):
        """Constituent{MassMatrix, MassMatrixInverse, Restriction, ProlongationAddition}
        (hierarchy, level, dimension) applied to every line of a nodal array (a copy is
        returned): what shuffle -> operator on each line -> unshuffle gives in the
        reference (tests/src/test_Tensor{MassMatrix,Restriction,Prolongation}.cpp)."""
        import torch
        out = v.contiguous().clone()
        check(_lib.lib().mgb_cpu_apply_operator(self._h, int(op), int(level), int(dimension), out.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream), "apply_operator")
        return out

    def decompose(self, v):
        """shuffle + decompose: nodal values -> shuffled multilevel coefficients."""
        return self._stage(_lib.lib().mgb_cpu_decompose, v)

    def recompose(self, u):
        return self._stage(_lib.lib().mgb_cpu_recompose, u).reshape(self.shape)

    def quantize(self, u, s, tolerance):
        import torch
        return self._stage(_lib.lib().mgb_cpu_quantize, u, torch.int64, (float(s), float(tolerance)))

    def dequantize(self, q, s, tolerance):
        import torch
        dt = torch.float32 if self.dtype == np.float32 else torch.float64
        return self._stage(_lib.lib().mgb_cpu_dequantize, q, dt, (float(s), float(tolerance)))


CPU_HUFFMAN_ZLIB, CPU_HUFFMAN_ZSTD = 1, 2  # pb::Encoding::Compressor (src/mgard.proto)


def compress(hierarchy, v, s, tolerance, compressor=CPU_HUFFMAN_ZLIB):
    """mgard::compress + CompressedDataset::write -> bytes.  `v`: numpy array
    (host) or torch CUDA tensor of hierarchy.shape; s = math.inf for L-infinity.
    `compressor`: the lossless stage the reference fixes at build time --
    CPU_HUFFMAN_ZLIB (build without zstd) or CPU_HUFFMAN_ZSTD (default build)."""
    is_torch = type(v).__module__.startswith("torch")
    if is_torch:
        v = v.contiguous()
        if tuple(v.shape) != hierarchy.shape:
            raise ValueError("array shape does not match the hierarchy")
        ptr = v.data_ptr()
        if _dtype_code(str(v.dtype).replace("torch.", "")) != _dtype_code(hierarchy.dtype):
            raise TypeError("array dtype does not match the hierarchy")
    else:
        v = np.ascontiguousarray(v, dtype=hierarchy.dtype)
        if v.shape != hierarchy.shape:
            raise ValueError("array shape does not match the hierarchy")
        ptr = v.ctypes.data
    keep = []
    shp = (C.c_uint64 * len(hierarchy.shape))(*hierarchy.shape)
    out = C.c_void_p()
    size = C.c_size_t()
    check(_lib.lib().mgb_cpu_compress(len(hierarchy.shape), _dtype_code(hierarchy.dtype), shp,
                                      _coord_args(hierarchy.coordinates, hierarchy.dtype, keep),
                                      float(s), float(tolerance), int(compressor), ptr, C.byref(out), C.byref(size)),
          "mgard::compress")
    try:
        return C.string_at(out.value, size.value)
    finally:
        _LIBC.free(out)


def decompress(data):
    """mgard::decompress(data, size) -> numpy array of the stored shape and type."""
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    out = C.c_void_p()
    ndim = C.c_int()
    dtype = C.c_int()
    shape = (C.c_uint64 * 5)()
    check(_lib.lib().mgb_cpu_decompress(buf.ctypes.data, buf.size, C.byref(out), C.byref(ndim), shape,
                                        C.byref(dtype)), "mgard::decompress")
    shp = tuple(int(shape[d]) for d in range(ndim.value))
    dt = np.float32 if dtype.value == 0 else np.float64
    n = int(np.prod(shp))
    try:
        return np.frombuffer(C.string_at(out.value, n * np.dtype(dt).itemsize), dtype=dt).reshape(shp).copy()
    finally:
        _LIBC.free(out)
