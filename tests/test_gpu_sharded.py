"""mgb_compress_sharded / mgb_decompress_sharded (one process per GPU, MaxDim slabs along
dim 0, DomainDecomposer.hpp:124-169) through the C ABI.

* one process: the header + records equal mgb_compress's stream byte for byte, for
  relative L-inf and s-norm bounds (global norm reduced on the device) and for device
  and host buffers (the three-stream pipeline of GPUPipelines.hpp:88-207);
* the records decode with mgb_decompress_sharded and with mgb_decompress;
* the s-norm stream equals the oracle's records with the oracle's global norm and local
  tolerance (ErrorToleranceCalculator.hpp:70-155);
* two processes on two GPUs (runs where two devices are visible): the assembled stream
  equals the one-process stream byte for byte and each rank decodes its own part."""
import os
import struct
import sys

import numpy as np
import pytest

import mgardx_oracle as mo

pytestmark = pytest.mark.gpu

SHAPE = (70, 33, 40)
SIZE = 24  # 24 + 24 + 22


def field(shape=SHAPE, seed=2):
    rng = np.random.default_rng(seed)
    g = np.meshgrid(*[np.linspace(0, 1, n) for n in shape], indexing="ij")
    return (np.sin(5 * g[0]) * np.cos(3 * g[1]) + g[2] ** 2 + 0.02 * rng.standard_normal(shape)).astype(np.float32)


@pytest.fixture(scope="module")
def env():
    import torch
    import mgard_b200 as mg
    from mgard_b200 import sharded
    assert torch.cuda.is_available()
    return torch, mg, sharded, torch.device("cuda:0")


@pytest.mark.parametrize("mode,tol,s", [(mo.REL, 1e-3, np.inf), (mo.REL, 1e-2, 0.0), (mo.ABS, 1e-3, np.inf),
                                        (mo.ABS, 1e-2, 0.0)])
def test_one_process_equals_high_level_stream(env, mode, tol, s):
    torch, mg, sharded, d = env
    u = field()
    cfg = mg.Config()
    cfg.domain_decomposition_dim = 0
    cfg.domain_decomposition_size = SIZE
    stream = mg.compress(u, tol, s, mode, config=cfg)            # host in, host out
    dstream = mg.compress(torch.from_numpy(u).to(d), tol, s, mode, config=cfg)  # device
    assert dstream.cpu().numpy().tobytes() == stream.tobytes()
    for local in (torch.from_numpy(u).to(d), u):
        r = sharded.compress_sharded_native(local, SHAPE, tol, s, mode, SIZE)
        rec = r["records"].cpu().numpy() if hasattr(r["records"], "cpu") else r["records"]
        assert r["header"] + rec.tobytes() == stream.tobytes()
        assert r["total"] == stream.size and r["stream_offset"] == len(r["header"])
        info = mg.peek_header(stream)
        if mode == mo.REL:  # the norm of the original data is only stored for relative bounds
            assert r["norm"] == info["norm"]
    # decode: sharded entry point (device and host output) and the plain one
    want = mg.decompress(stream)
    out = torch.empty(SHAPE, dtype=torch.float32, device=d)
    sharded.decompress_sharded_native(r["header"], torch.from_numpy(rec).to(d), out)
    assert np.array_equal(out.cpu().numpy(), want)
    hout = np.empty(SHAPE, dtype=np.float32)
    sharded.decompress_sharded_native(r["header"], rec, hout)
    assert np.array_equal(hout, want)
    if np.isinf(s):
        assert np.abs(want - u).max() <= tol * (np.abs(u).max() if mode == mo.REL else 1)
    # walking the chain
    import ctypes as C
    from mgard_b200 import _lib
    offs, sizes, cnt = (C.c_uint64 * 8)(), (C.c_uint64 * 8)(), C.c_uint64(0)
    _lib.check(_lib.lib().mgb_stream_records(stream.ctypes.data, stream.size, offs, sizes, 8, C.byref(cnt)), "records")
    assert cnt.value == 3 and offs[0] == info["header_bytes"] and offs[2] + sizes[2] == stream.size


def test_snorm_records_match_oracle(env):
    """REL, s = 0, three sub-domains: global norm sqrt(sum u^2 / N) in T, local tolerance
    sqrt((tol*norm)^2 / 3), every record byte-identical to the oracle's."""
    torch, mg, sharded, d = env
    u = field(seed=5)
    tol = 1e-2
    r = sharded.compress_sharded_native(torch.from_numpy(u).to(d), SHAPE, tol, 0.0, mo.REL, SIZE)
    norm = np.float32(r["norm"])
    want_norm = np.sqrt(np.float32((u.astype(np.float64) ** 2).sum()) / np.float32(u.size))
    assert abs(float(norm) - float(want_norm)) <= 2e-7 * float(want_norm)
    ltol = float(np.sqrt((np.float32(tol) * norm) * (np.float32(tol) * norm) / np.float32(3)))
    raw = r["records"].cpu().numpy().tobytes()
    off, lo = 0, 0
    for e in (24, 24, 22):
        size = struct.unpack_from("<Q", raw, off)[0]
        sub = np.ascontiguousarray(u[lo:lo + e])
        ref = mo.compress_lowlevel(mo.Hierarchy(sub.shape, np.float32), sub, mo.ABS, ltol, 0.0)
        if size == sub.nbytes:
            assert len(ref["payload"]) >= sub.nbytes
        else:
            assert raw[off + 8:off + 8 + size] == ref["payload"]
        off += 8 + size
        lo += e
    assert off == len(raw)


def test_zero_field_relative_bound_decomposed(env):
    """norm 0 -> epsilon (NormCalculator.hpp:50-52) in the decomposed path too."""
    torch, mg, sharded, d = env
    z = np.zeros(SHAPE, dtype=np.float32)
    cfg = mg.Config()
    cfg.domain_decomposition_dim, cfg.domain_decomposition_size = 0, SIZE
    st = mg.compress(z, 1e-3, np.inf, mo.REL, config=cfg)
    assert mg.peek_header(st)["norm"] == float(np.finfo(np.float32).eps)
    assert np.array_equal(mg.decompress(st), z)
    st = mg.compress(z, 1e-3, 0.0, mo.REL, config=cfg)
    assert np.array_equal(mg.decompress(st), z)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import mgard_b200 as mg
    from mgard_b200 import sharded
    torch.cuda.set_device(rank)
    d = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=d)
    u = field()
    ext = sharded.partition(SHAPE[0], SIZE)
    comm = sharded.Comm(dist)
    first, count = comm.owned(len(ext))
    lo, hi = sum(ext[:first]), sum(ext[:first + count])
    res = {}
    for s in (np.inf, 0.0):
        r = sharded.compress_sharded_native(torch.from_numpy(u[lo:hi].copy()).to(d), SHAPE, 1e-3, s, mo.REL, SIZE,
                                            comm=comm)
        out = torch.empty((hi - lo,) + SHAPE[1:], dtype=torch.float32, device=d)
        sharded.decompress_sharded_native(r["header"], r["records"], out, comm=comm)
        res[str(s)] = (r["stream_offset"], r["records"].cpu().numpy().tobytes(), r["header"], r["total"], r["norm"],
                       out.cpu().numpy(), lo, hi)
    q.put((rank, res))
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


def test_two_processes_two_gpus_equal_one_process(env):
    torch, mg, sharded, d = env
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs (run once with gpurun --gpus 2; see profiles/)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    u = field()
    cfg = mg.Config()
    cfg.domain_decomposition_dim, cfg.domain_decomposition_size = 0, SIZE
    for s in (np.inf, 0.0):
        stream = mg.compress(u, 1e-3, s, mo.REL, config=cfg).tobytes()
        want = mg.decompress(np.frombuffer(stream, dtype=np.uint8))
        body = bytearray(len(stream))
        for rank in (0, 1):
            off, rec, hdr, total, norm, back, lo, hi = got[rank][str(s)]
            assert total == len(stream)
            body[:len(hdr)] = hdr
            body[off:off + len(rec)] = rec
            assert np.array_equal(back, want[lo:hi])
        assert bytes(body) == stream
