"""CPU, world_size 2, gloo: the host-side logic of the slab-sharded path
(partition, norm all-reduce, size all-gather / offsets, header) with the oracle
plugged in as the local compressor.  The assembled stream must equal the
single-process MaxDim-decomposed stream."""
import os
import struct
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mgardx_oracle as mo
from mgard_b200 import sharded

SHAPE = (23, 9, 10)
SIZE = 6  # 6 + 6 + 6 + 5
TOL = 1e-2


def make_field():
    rng = np.random.default_rng(3)
    g = np.meshgrid(*[np.linspace(0, 1, n) for n in SHAPE], indexing="ij")
    return (np.sin(4 * g[0]) + g[1] * g[2] + 0.01 * rng.standard_normal(SHAPE)).astype(np.float32)


def oracle_backend(s):
    def partials(local):
        a = np.asarray(local, dtype=np.float64)
        return float(np.abs(a).max()), float((a * a).sum())

    def compress(local, gshape, tol, s_, mode, norm, first, count, size):
        ext = sharded.partition(gshape[0], size)
        ltol = float(np.float32(tol) * np.float32(norm)) if np.isinf(s_) else float(
            np.sqrt((np.float32(tol) * np.float32(norm)) ** 2 / np.float32(len(ext))))
        out = b""
        lo = 0
        for i in range(first, first + count):
            sub = np.ascontiguousarray(local[lo:lo + ext[i]])
            h = mo.Hierarchy(sub.shape, sub.dtype)
            pay = mo.compress_lowlevel(h, sub, mo.ABS, ltol, s_)["payload"]
            if len(pay) >= sub.nbytes:
                pay = sub.tobytes()
            out += struct.pack("<Q", len(pay)) + pay
            lo += ext[i]
        return out

    def header(gshape, npdt, tol, s_, mode, norm, size):
        return mo.encode_preamble(mo.encode_header(gshape, npdt, mode, tol, s_, norm, None, True, 0, size))

    return compress, partials, header


def single_process_stream(u, s):
    comp, part, hdr = oracle_backend(s)
    mx, ss = part(u)
    norm = sharded.global_norm(mx, ss, u.size, s, np.float32)
    ext = sharded.partition(SHAPE[0], SIZE)
    return hdr(SHAPE, np.float32, TOL, s, mo.REL, norm, SIZE) + comp(u, SHAPE, TOL, s, mo.REL, norm, 0, len(ext), SIZE)


def worker(rank, world, port, s, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    u = make_field()
    ext = sharded.partition(SHAPE[0], SIZE)
    first, count = sharded.owned_range(len(ext), rank, world)
    lo = sum(ext[:first])
    hi = lo + sum(ext[first:first + count])
    comp, part, hdr = oracle_backend(s)
    r = sharded.compress_sharded(u[lo:hi], SHAPE, TOL, s, mo.REL, SIZE, dist=dist,
                                 local_compress=comp, local_partials=part, write_header=hdr)
    gathered = [None] * world
    dist.all_gather_object(gathered, (r["offset"], bytes(r["records"]), r["sizes"], r["norm"]))
    if rank == 0:
        body = bytearray(sum(r["sizes"]))
        for off, rec, sizes, norm in gathered:
            body[off:off + len(rec)] = rec
            assert sizes == r["sizes"] and norm == r["norm"]
        q.put(r["header"] + bytes(body))
    dist.destroy_process_group()


@pytest.mark.parametrize("s", [float("inf"), 0.0])
def test_two_ranks_equal_single_process(s):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (0 if np.isinf(s) else 1)
    procs = [ctx.Process(target=worker, args=(r, 2, port, s, q)) for r in range(2)]
    for p in procs:
        p.start()
    stream = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert stream == single_process_stream(make_field(), s)


def test_partition_rules():
    # DomainDecomposer.hpp:131-144 and SURVEY §8e (2049 = 7*257 + 250)
    assert sharded.partition(2049, 257) == [257] * 7 + [250]
    assert sharded.partition(70, 24) == [24, 24, 22]
    for world in (1, 2, 4, 8):
        got = [sharded.owned_range(8, r, world) for r in range(world)]
        assert sum(c for _, c in got) == 8 and got[0][0] == 0
        assert all(got[i][0] + got[i][1] == got[i + 1][0] for i in range(world - 1))
