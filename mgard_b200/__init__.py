"""mgard_b200 — B200-native MGARD-X hot path (CUDA, sm_100a) behind the
reference's own API.

Host-side mirror of the reference interface (include/compress_x.hpp:31-178,
include/mgard-x/Config/Config.h, include/mgard-x/Utilities/Types.h) over the C
ABI in include/mgard_b200.h.  Torch is used for device memory only.
"""
from .api import (Config, Plan, compress, decompress, peek_header, release_cache,
                  error_bound_type, data_type, compress_status_type, MgardError,
                  launch_count, tune, TUNE_SERIAL_MIN_CHUNKS, TUNE_RING_DECODER, TUNE_SUB_ENCODER, lossless_type, decomposition_type, domain_decomposition_type, adjust_shape, pin_memory, check_memory_pinned,
                  unpin_memory)

__all__ = ["Config", "Plan", "compress", "decompress", "peek_header",
           "release_cache", "error_bound_type", "data_type",
           "compress_status_type", "MgardError", "launch_count", "tune", "TUNE_SERIAL_MIN_CHUNKS", "TUNE_RING_DECODER", "TUNE_SUB_ENCODER", "lossless_type", "decomposition_type", "domain_decomposition_type", "adjust_shape", "pin_memory", "check_memory_pinned",
           "unpin_memory"]
