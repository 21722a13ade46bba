"""Developer A/B tool (run under gpurun): per-kernel-family CUDA-event times of one
low-level compress + decompress for a shape, under one or more settings of
mgb_tune's serial_min_chunks.  Usage:
  python scripts/kernel_times.py 513,513,513 [257,2049,2049 ...] [--serial 0,-1] [--tol 1e-3]"""
import argparse, ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mgard_b200 as mg
from mgard_b200 import _lib
import bench

ap = argparse.ArgumentParser()
ap.add_argument("shapes", nargs="+")
ap.add_argument("--serial", default="0,-1")
ap.add_argument("--tol", type=float, default=1e-3)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--plane0", type=int, default=0, help="first plane of the slab in the domain")
ap.add_argument("--full-n0", type=int, default=0, help="planes of the domain the shape is a slab of (bench field of C5: 2049)")
a = ap.parse_args()
L = _lib.lib()
dev = torch.device("cuda:0")


def families(reps):
    fam, k = {}, 0
    while True:
        name, n_l, tot, mx = C.c_char_p(), C.c_ulonglong(0), C.c_double(0), C.c_double(0)
        if L.mgb_profile_report(k, C.byref(name), C.byref(n_l), C.byref(tot), C.byref(mx)) != 0:
            break
        if n_l.value:
            fam[name.value.decode()] = dict(n=n_l.value / reps, ms=round(tot.value / reps, 4), max_ms=round(mx.value, 4))
        k += 1
    return fam


for sh in a.shapes:
    shape = tuple(int(x) for x in sh.split(","))
    u = bench.field_torch(shape, dev, plane0=a.plane0, full_n0=a.full_n0 or None)
    p = mg.Plan(shape, np.float32)
    for serial in [int(x) for x in a.serial.split(",")]:
        mg.tune(mg.TUNE_SERIAL_MIN_CHUNKS, serial)
        for _ in range(2):
            payload, norm = p.compress(u, mg.error_bound_type.REL, a.tol, float("inf"))
            back = p.decompress(payload, mg.error_bound_type.REL, a.tol, float("inf"), norm)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tc = td = 0.0
        for _ in range(a.reps):
            ev[0].record()
            payload, norm = p.compress(u, mg.error_bound_type.REL, a.tol, float("inf"))
            ev[1].record()
            back = p.decompress(payload, mg.error_bound_type.REL, a.tol, float("inf"), norm)
            ev[2].record()
            torch.cuda.synchronize()
            tc += ev[0].elapsed_time(ev[1]) / a.reps
            td += ev[1].elapsed_time(ev[2]) / a.reps
        L.mgb_profile_enable(1)
        for _ in range(a.reps):
            payload, norm = p.compress(u, mg.error_bound_type.REL, a.tol, float("inf"))
            back = p.decompress(payload, mg.error_bound_type.REL, a.tol, float("inf"), norm)
        torch.cuda.synchronize()
        L.mgb_profile_enable(0)
        err = float((back - u).abs().max())
        print(json.dumps(dict(shape=shape, serial_min_chunks=serial, compress_ms=round(tc, 3), decompress_ms=round(td, 3),
                              GBs=round(2 * u.numel() * 4 / (tc + td) / 1e6, 1), ratio=round(u.numel() * 4 / payload.numel(), 3),
                              err_ok=err <= a.tol * norm, fam=families(a.reps))), flush=True)
    del u, p
    torch.cuda.empty_cache()
