"""ctypes loader for mgard_b200/libmgard_b200.so (the C ABI of include/mgard_b200.h).

The product path fails loudly when the CUDA library is missing: there is no
CPU fallback anywhere in this package.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmgard_b200.so")

# status codes, include/mgard_b200.h (mirror mgard_x::compress_status_type)
SUCCESS, FAILURE, OUTPUT_TOO_LARGE, TOO_MANY_DIMS, BAD_DTYPE, BACKEND_NOT_AVAILABLE = range(6)
BAD_ARGUMENT, BAD_STREAM, CUDA_ERROR = 16, 17, 18
STATUS_NAMES = {
    0: "Success", 1: "Failure", 2: "OutputTooLargeFailure",
    3: "NotSupportHigherNumberOfDimensionsFailure", 4: "NotSupportDataTypeFailure",
    5: "BackendNotAvailableFailure", 16: "BadArgument", 17: "BadStream", 18: "CudaError",
}


class MgbConfig(C.Structure):
    """mgb_config: the fields of mgard_x::Config (include/mgard-x/Config/Config.h:10-42) that change
    what the hot path computes."""
    _fields_ = [
        ("dev_id", C.c_int32), ("huff_dict_size", C.c_int32),
        ("huff_block_size", C.c_int32), ("domain_decomposition_dim", C.c_int32),
        ("domain_decomposition_size", C.c_uint64),
        ("normalize_coordinates", C.c_int32), ("lossless", C.c_int32),
        ("zstd_compress_level", C.c_int32), ("reorder", C.c_int32),
        ("decomposition", C.c_int32), ("domain_decomposition", C.c_int32),
        ("max_larget_level", C.c_uint64), ("block_size", C.c_uint64),
        ("domain_decomposition_sizes", C.POINTER(C.c_uint64)),
        ("num_domain_decomposition_sizes", C.c_uint64),
        ("max_memory_footprint", C.c_uint64),
    ]


_lib = None

_vp, _u64, _i32, _dbl = C.c_void_p, C.c_uint64, C.c_int, C.c_double
_pu64 = C.POINTER(C.c_uint64)

# name -> (restype, argtypes); every symbol declared in include/mgard_b200.h
SIGNATURES = {
    "mgb_config_default": (None, [C.POINTER(MgbConfig)]),
    "mgb_plan_create": (_i32, [_i32, _pu64, _i32, C.POINTER(_vp), C.POINTER(MgbConfig), C.POINTER(_vp)]),
    "mgb_plan_destroy": (None, [_vp]),
    "mgb_plan_l_target": (_i32, [_vp]),
    "mgb_plan_set_generic": (None, [_vp, _i32]),
    "mgb_plan_num_elems": (_u64, [_vp]),
    "mgb_plan_level_shape": (_u64, [_vp, _i32, _i32]),
    "mgb_plan_table": (_u64, [_vp, _i32, _i32, _i32, _vp, _u64]),
    "mgb_decompose": (_i32, [_vp, _vp, _vp, _vp]),
    "mgb_recompose": (_i32, [_vp, _vp, _vp, _vp]),
    "mgb_norm": (_i32, [_vp, _vp, _dbl, C.POINTER(_dbl)]),
    "mgb_norm_partials": (_i32, [_vp, _vp, C.POINTER(_dbl), C.POINTER(_dbl)]),
    "mgb_quantize": (_i32, [_vp, _vp, _i32, _dbl, _dbl, _dbl, _vp, _vp, _vp, _vp, _vp, _u64, _vp]),
    "mgb_dequantize": (_i32, [_vp, _vp, _u64, _vp, _vp, _i32, _dbl, _dbl, _dbl, _vp, _vp]),
    "mgb_codebook": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "mgb_huffman_compress": (_i32, [_vp, _vp, _u64, _vp, _u64, _vp, _vp, _vp, _u64, _pu64, _vp]),
    "mgb_huffman_decompress": (_i32, [_vp, _vp, _u64, _vp, _u64, _pu64, C.POINTER(_vp), C.POINTER(_vp), _vp]),
    "mgb_compress_lowlevel": (_i32, [_vp, _vp, _i32, _dbl, _dbl, C.POINTER(_dbl), _vp, _u64, _pu64, _vp]),
    "mgb_decompress_lowlevel": (_i32, [_vp, _vp, _u64, _i32, _dbl, _dbl, _dbl, _vp, _vp]),
    "mgb_compress": (_i32, [_i32, _i32, _pu64, _dbl, _dbl, _i32, _vp, C.POINTER(_vp), C.POINTER(C.c_size_t), C.POINTER(_vp), C.POINTER(MgbConfig), _i32]),
    "mgb_decompress": (_i32, [_vp, C.c_size_t, C.POINTER(_vp), C.POINTER(MgbConfig), _i32, C.POINTER(_i32), _pu64, C.POINTER(_i32)]),
    "mgb_peek_header": (_i32, [_vp, C.c_size_t, C.POINTER(_i32), _pu64, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_dbl), C.POINTER(_dbl), C.POINTER(_dbl), _pu64]),
    "mgb_release_cache": (None, []),
    "mgb_compress_subdomains": (_i32, [_i32, _i32, _pu64, _dbl, _dbl, _i32, _dbl, _vp, _u64, _u64, C.POINTER(MgbConfig), _vp, _u64, _pu64]),
    "mgb_comm_unique_id": (_i32, [_vp]),
    "mgb_comm_init_rank": (_i32, [_vp, _i32, _i32, C.POINTER(_vp)]),
    "mgb_comm_from_nccl": (_i32, [_vp, _i32, _i32, C.POINTER(_vp)]),
    "mgb_comm_destroy": (None, [_vp]),
    "mgb_comm_rank": (_i32, [_vp]),
    "mgb_comm_size": (_i32, [_vp]),
    "mgb_owned_subdomains": (_i32, [_vp, _u64, _pu64, _pu64]),
    "mgb_compress_sharded": (_i32, [_vp, _i32, _i32, _pu64, _dbl, _dbl, _i32, _vp, C.POINTER(MgbConfig), _vp, _u64,
                                    _pu64, _pu64, _pu64, _pu64, C.POINTER(_dbl), _vp, _u64, _pu64]),
    "mgb_decompress_sharded": (_i32, [_vp, _vp, _u64, _vp, _u64, _vp, C.POINTER(MgbConfig)]),
    "mgb_stream_records": (_i32, [_vp, _u64, _pu64, _pu64, _u64, _pu64]),
    "mgb_write_header": (_i32, [_i32, _i32, _pu64, _dbl, _dbl, _i32, _dbl, C.POINTER(_vp), C.POINTER(MgbConfig), _vp, _u64, _pu64]),
    "mgb_pin_memory": (_i32, [_vp, _u64]),
    "mgb_check_memory_pinned": (_i32, [_vp]),
    "mgb_unpin_memory": (_i32, [_vp]),
    "mgb_cpu_plan_create": (_i32, [_i32, _pu64, _i32, C.POINTER(_vp), C.POINTER(_vp)]),
    "mgb_cpu_plan_destroy": (None, [_vp]),
    "mgb_cpu_plan_levels": (_i32, [_vp]),
    "mgb_cpu_plan_ndof": (_u64, [_vp, _i32]),
    "mgb_cpu_plan_level_shape": (_u64, [_vp, _i32, _i32]),
    "mgb_cpu_shuffle": (_i32, [_vp, _vp, _vp, _vp]),
    "mgb_cpu_unshuffle": (_i32, [_vp, _vp, _vp, _vp]),
    "mgb_cpu_decompose": (_i32, [_vp, _vp, _vp, _vp]),
    "mgb_cpu_recompose": (_i32, [_vp, _vp, _vp, _vp]),
    "mgb_cpu_apply_operator": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "mgb_cpu_quantize": (_i32, [_vp, _vp, _dbl, _dbl, _vp, _vp]),
    "mgb_cpu_dequantize": (_i32, [_vp, _vp, _dbl, _dbl, _vp, _vp]),
    "mgb_cpu_compress": (_i32, [_i32, _i32, _pu64, C.POINTER(_vp), _dbl, _dbl, _i32, _vp, C.POINTER(_vp), C.POINTER(C.c_size_t)]),
    "mgb_cpu_write_header": (_i32, [_i32, _i32, _pu64, C.POINTER(_vp), _dbl, _dbl, _i32, _vp, _u64, _pu64]),
    "mgb_cpu_decompress": (_i32, [_vp, C.c_size_t, C.POINTER(_vp), C.POINTER(_i32), _pu64, C.POINTER(_i32)]),
    "mgb_tune": (_i32, [_i32, C.c_longlong]),
    "mgb_launch_count": (_u64, []),
    "mgb_profile_enable": (None, [_i32]),
    "mgb_profile_report": (_i32, [_i32, C.POINTER(C.c_char_p), C.POINTER(C.c_ulonglong), C.POINTER(_dbl), C.POINTER(_dbl)]),
    "mgb_version": (C.c_char_p, []),
}


def lib():
    """Load the CUDA shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import "
                "__graft_entry__ as g; g.build()'` (mgard_b200 has no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


class MgardError(RuntimeError):
    def __init__(self, status, where):
        self.status = status
        super().__init__(f"{where}: {STATUS_NAMES.get(status, status)} ({status})")


def check(status, where):
    if status != SUCCESS:
        raise MgardError(status, where)
