"""One-process-per-GPU slab-decomposed compression (torch.distributed).

The reference has no distributed code: its "multi-GPU" path is the MaxDim
domain decomposition processed serially on one device
(include/mgard-x/CompressionHighLevel/GPUPipelines.hpp:88-207), and the
published scaling runs used one MPI rank per GPU outside the library
(doc/MGARD-X.md:287).  Here the SAME MaxDim partition
(include/mgard-x/DomainDecomposer/DomainDecomposer.hpp:124-169) is spread over
ranks: rank r owns `count` consecutive sub-domains starting at `first`.  The
only exchanges are

  1. one all-reduce of a double — max|u| (s = inf) or sum u^2 (s-norm) — behind
     relative error bounds (ErrorToleranceCalculator.hpp:70-155), and
  2. one all-gather of the per-rank container sizes, from which every rank
     knows its byte offset in the stream (GPUPipelines.hpp:189-193).

The stream that results is byte-identical for any number of ranks (up to the
order of the outlier list, which is unordered in the reference as well) and is
decodable by mgard_x::decompress.

`local_compress` is injectable so that the host-side logic can be exercised on
CPU with the gloo backend (tests/test_sharded_gloo.py).

On CUDA tensors the work is done by the C entry points mgb_compress_sharded /
mgb_decompress_sharded (include/mgard_b200.h) over the library's own NCCL
communicator (class Comm): the norm all-reduce runs stream-ordered on the device
and the quantizer reads it from device memory - no host round trip, one
synchronisation per record.  Sub-domain ownership: mgb_owned_subdomains ==
owned_range below.
"""
import ctypes as C
import math

import numpy as np


def partition(n0, size):
    """Sub-domain extents along dim 0 (DomainDecomposer.hpp:131-144)."""
    count = (n0 - 1) // size + 1
    ext = []
    for i in range(count):
        ext.append(size if i < n0 // size else n0 % size)
    return ext


def owned_range(num_subdomains, rank, world):
    """Contiguous block of sub-domains per rank (balanced, rank-major)."""
    base, rem = divmod(num_subdomains, world)
    first = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return first, count


def global_norm(absmax, sumsq, total_elems, s, dtype, dist=None, group=None, device=None):
    """calc_norm_decomposed (ErrorToleranceCalculator.hpp:91-132) across ranks."""
    import torch
    if dist is not None and dist.is_initialized() and dist.get_world_size(group) > 1:
        if math.isinf(s):
            t = torch.tensor([absmax], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            absmax = float(t.item())
        else:
            t = torch.tensor([sumsq], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            sumsq = float(t.item())
    if math.isinf(s):
        norm = absmax
    elif np.dtype(dtype) == np.float32:
        norm = float(np.sqrt(np.float32(sumsq) / np.float32(total_elems)))
    else:
        norm = math.sqrt(sumsq / total_elems)
    if np.dtype(dtype) == np.float32:
        norm = float(np.float32(norm))
        if norm == 0:
            norm = float(np.finfo(np.float32).eps)
    elif norm == 0:
        norm = float(np.finfo(np.float64).eps)
    return norm


def exchange_sizes(local_size, dist=None, group=None, device=None):
    """All-gather of the per-rank container sizes -> (sizes, my offset)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [int(local_size)], 0
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = torch.tensor([int(local_size)], dtype=torch.int64, device=device)
    allsz = torch.zeros(world, dtype=torch.int64, device=device)
    if hasattr(dist, "all_gather_into_tensor") and device is not None:
        dist.all_gather_into_tensor(allsz, mine, group=group)
        sizes = [int(x) for x in allsz.tolist()]  # one device-to-host read
    else:  # gloo (CPU tests)
        parts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        sizes = [int(x.item()) for x in parts]
    return sizes, sum(sizes[:rank])


def compress_sharded(local, global_shape, tol, s, mode, decomposition_size, config=None,
                     dist=None, group=None, local_compress=None, local_partials=None,
                     write_header=None):
    """Compress this rank's slabs of a field of `global_shape` that is MaxDim
    decomposed along dim 0 with `decomposition_size` planes per sub-domain.

    local: this rank's planes (torch CUDA tensor or numpy array in tests),
           shape (owned planes, *global_shape[1:]).
    Returns dict(records=<bytes-like of this rank's `u64 size|payload` records>,
                 sizes=[per-rank sizes], offset=<my offset after the header>,
                 header=<header bytes>, norm=<global norm>).
    """
    world = dist.get_world_size(group) if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    ext = partition(int(global_shape[0]), int(decomposition_size))
    first, count = owned_range(len(ext), rank, world)
    total = int(np.prod(global_shape))
    is_torch = type(local).__module__.startswith("torch")
    device = local.device if is_torch and local.is_cuda else None
    np_dtype = np.float32 if ("float32" in str(local.dtype)) else np.float64
    if local_compress is None and device is not None:
        return compress_sharded_native(local, global_shape, tol, s, mode, decomposition_size, config=config,
                                       comm=default_comm(dist, group))
    if local_compress is None:
        local_compress, local_partials, write_header = _cuda_backend(config)
    norm = 1.0
    if int(mode) == 0:  # REL
        absmax, sumsq = local_partials(local)
        norm = global_norm(absmax, sumsq, total, s, np_dtype, dist, group, device)
    records = local_compress(local, global_shape, tol, s, mode, norm, first, count,
                             decomposition_size)
    nbytes = records.numel() if is_torch else len(records)
    sizes, offset = exchange_sizes(nbytes, dist, group, device)
    header = write_header(global_shape, np_dtype, tol, s, mode, norm, decomposition_size)
    return dict(records=records, sizes=sizes, offset=offset, header=header, norm=norm,
                first=first, count=count)


_PARTIAL_PLANS = {}


def _cuda_backend(config):
    """Local work through the C ABI (mgb_norm_partials / mgb_compress_subdomains /
    mgb_write_header)."""
    import torch
    from . import _lib
    from .api import Config, Plan, _shape_arg, _dtype_code
    L = _lib.lib()
    cfg_py = config or Config()

    def cfg_for(size):
        c = cfg_py._c()
        c.domain_decomposition_dim = 0
        c.domain_decomposition_size = int(size)
        return c

    def partials(local):
        # the plan only carries the reduction scratch here; keep it across calls
        key = (tuple(local.shape), str(local.dtype), local.device.index)
        p = _PARTIAL_PLANS.get(key)
        if p is None:
            p = Plan(tuple(local.shape), np.float32 if local.dtype == torch.float32 else np.float64,
                     config=cfg_py)
            _PARTIAL_PLANS[key] = p
        mx, ss = C.c_double(0), C.c_double(0)
        torch.cuda.current_stream().synchronize()
        _lib.check(L.mgb_norm_partials(p._h, local.data_ptr(), C.byref(mx), C.byref(ss)),
                   "norm_partials")
        return mx.value, ss.value

    def compress(local, gshape, tol, s, mode, norm, first, count, size):
        npdt = np.float32 if local.dtype == torch.float32 else np.float64
        cap = local.numel() * local.element_size() + count * (8 * (128 + cfg_py.huff_dict_size) + (1 << 20))
        out = torch.empty(cap, dtype=torch.uint8, device=local.device)
        sz = C.c_uint64(0)
        c = cfg_for(size)
        torch.cuda.current_stream().synchronize()
        _lib.check(L.mgb_compress_subdomains(len(gshape), int(_dtype_code(npdt)),
                                             _shape_arg(gshape), float(tol), float(s),
                                             int(mode), float(norm), local.data_ptr(),
                                             first, count, C.byref(c), out.data_ptr(), cap,
                                             C.byref(sz)), "compress_subdomains")
        return out[:sz.value]

    def header(gshape, npdt, tol, s, mode, norm, size):
        buf = np.zeros(1 << 16, dtype=np.uint8)
        sz = C.c_uint64(0)
        c = cfg_for(size)
        _lib.check(L.mgb_write_header(len(gshape), int(_dtype_code(npdt)), _shape_arg(gshape),
                                      float(tol), float(s), int(mode), float(norm), None,
                                      C.byref(c), buf.ctypes.data, buf.size, C.byref(sz)),
                   "write_header")
        return buf[:sz.value].tobytes()

    return compress, partials, header


# ---------------------------------------------------------------------------------
# native path: mgb_compress_sharded / mgb_decompress_sharded over the library's NCCL
# communicator
# ---------------------------------------------------------------------------------
class Comm:
    """mgb_comm: rank / size + an NCCL communicator created by the library
    (ncclGetUniqueId on rank 0, shipped through torch.distributed, ncclCommInitRank)."""

    def __init__(self, dist=None, group=None):
        import torch
        from . import _lib
        L = _lib.lib()
        self.rank, self.size = 0, 1
        if dist is not None and dist.is_initialized():
            self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)
        uid = (C.c_uint8 * 128)()
        if self.size > 1:
            if self.rank == 0:
                _lib.check(L.mgb_comm_unique_id(uid), "ncclGetUniqueId")
            box = [bytes(uid)]
            dist.broadcast_object_list(box, src=0, group=group)
            uid = (C.c_uint8 * 128).from_buffer_copy(box[0])
        h = C.c_void_p(0)
        _lib.check(L.mgb_comm_init_rank(uid, self.size, self.rank, C.byref(h)), "ncclCommInitRank")
        self._h = h

    def close(self):
        from . import _lib
        if getattr(self, "_h", None):
            _lib.lib().mgb_comm_destroy(self._h)
            self._h = None

    def owned(self, num_subdomains):
        from . import _lib
        f, c = C.c_uint64(0), C.c_uint64(0)
        _lib.check(_lib.lib().mgb_owned_subdomains(self._h, num_subdomains, C.byref(f), C.byref(c)), "owned")
        return f.value, c.value


_DEFAULT_COMM = {}


def default_comm(dist=None, group=None):
    key = id(group)
    c = _DEFAULT_COMM.get(key)
    if c is None:
        c = _DEFAULT_COMM[key] = Comm(dist, group)
    return c


def compress_sharded_native(local, global_shape, tol, s, mode, decomposition_size, config=None, comm=None,
                            out=None):
    """mgb_compress_sharded.  local / out: torch CUDA tensors or numpy (host) arrays."""
    from . import _lib
    from .api import Config, _shape_arg, _dtype_code, _is_torch
    L = _lib.lib()
    cfg = (config or Config())._c()
    cfg.domain_decomposition_dim = 0
    cfg.domain_decomposition_size = int(decomposition_size)
    tdev = _is_torch(local)
    if tdev:
        import torch
        local = local.contiguous()
        npdt = np.float32 if local.dtype == torch.float32 else np.float64
        in_ptr, nbytes = local.data_ptr(), local.numel() * local.element_size()
        if local.is_cuda:
            cfg.dev_id = local.device.index or 0
    else:
        local = np.ascontiguousarray(local)
        npdt = local.dtype
        in_ptr, nbytes = local.ctypes.data, local.nbytes
    ext = partition(int(global_shape[0]), int(decomposition_size))
    nranks = comm.size if comm else 1
    if out is None:
        cap = nbytes + len(ext) * (8 * (128 + cfg.huff_dict_size) + (1 << 20))
        if tdev:
            import torch
            out = torch.empty(cap, dtype=torch.uint8, device=local.device)
        else:
            out = np.empty(cap, dtype=np.uint8)
    if _is_torch(out):
        out_ptr, cap = out.data_ptr(), out.numel()
    else:
        out_ptr, cap = out.ctypes.data, out.size
    lsz, off, tot, nrm, hsz = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_double(0), C.c_uint64(0)
    sizes = (C.c_uint64 * nranks)()
    hdr = (C.c_uint8 * (1 << 16))()
    _lib.check(L.mgb_compress_sharded(comm._h if comm else None, len(global_shape), int(_dtype_code(npdt)),
                                      _shape_arg(global_shape), float(tol), float(s), int(mode), in_ptr,
                                      C.byref(cfg), out_ptr, cap, C.byref(lsz), C.byref(off), C.byref(tot), sizes,
                                      C.byref(nrm), hdr, len(hdr), C.byref(hsz)), "compress_sharded")
    first, count = owned_range(len(ext), comm.rank if comm else 0, nranks)
    return dict(records=out[:lsz.value], sizes=[int(x) for x in sizes], offset=off.value - hsz.value,
                stream_offset=off.value, total=tot.value, header=bytes(hdr[:hsz.value]), norm=nrm.value,
                first=first, count=count)


def decompress_sharded_native(header, records, out, config=None, comm=None):
    """mgb_decompress_sharded: `records` (this rank's, host or device) -> `out` (this
    rank's sub-domains back to back, host or device)."""
    from . import _lib
    from .api import Config, _is_torch
    L = _lib.lib()
    cfg = (config or Config())._c()
    if _is_torch(records):
        rp, rn = records.data_ptr(), records.numel()
    else:
        records = np.ascontiguousarray(records, dtype=np.uint8)
        rp, rn = records.ctypes.data, records.size
    if _is_torch(out):
        op = out.data_ptr()
        if out.is_cuda:
            cfg.dev_id = out.device.index or 0
    else:
        op = out.ctypes.data
    hb = (C.c_uint8 * len(header)).from_buffer_copy(header)
    _lib.check(L.mgb_decompress_sharded(comm._h if comm else None, hb, len(header), rp, rn, op, C.byref(cfg)),
               "decompress_sharded")
    return out
