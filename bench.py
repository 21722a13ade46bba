#!/usr/bin/env python
"""bench.py — MGARD-X hot path on B200: compress + decompress throughput.

Workload (BASELINE.json configs[1], SURVEY.md §8d "C2"): 3-D fp32 513^3
synthetic field, relative L-inf bound 1e-3 (s = inf), Huffman lossless, dict
8192, block 20480.  One "step" = one compress + one decompress of the field.

  value  (device resident)  original bytes moved through the codec per second:
         2 * N * 4 B / (t_compress + t_decompress); inputs already in HBM.
  e2e    same metric through the public host API mgard_b200.compress /
         decompress with pinned HOST buffers (H2D of the field, D2H of the
         stream, and back) inside the timed region.
  N > 1  weak scaling: the domain is (N*513) x 513 x 513, MaxDim-decomposed
         along dim 0 with 513 planes per sub-domain, one sub-domain per rank,
         global norm all-reduce + size all-gather (mgard_b200/sharded.py).

`--impl reference` times the UNMODIFIED reference (MGARD-X SERIAL adapter built
from /root/reference as oracle/_ref) on the host, on a bounded 129^3 / 257^3
sample of the same field.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "compress/decompress GB/s at 1/2/4/8 B200 vs HBM roofline; ratio at bound"
SHAPE = (513, 513, 513)
TOL, S = 1e-3, float("inf")
SEED = 2049
_REAL_STDOUT = None  # saved stdout fd when N > 1 (see main)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def field_numpy(shape, lo=0, seed=SEED, full_shape=None):
    """SURVEY §8d C2 field on `shape` (row-major index offset `lo` elements)."""
    import numpy as np
    full_shape = full_shape or shape
    x = [np.arange(n, dtype=np.float64) / (n - 1) for n in shape]
    g = np.meshgrid(*x, indexing="ij")
    u = (np.sin(6 * np.pi * g[0]) * np.cos(4 * np.pi * g[1]) * np.sin(2 * np.pi * g[2])
         + 0.3 * np.sin(40 * np.pi * g[0] * g[1]))
    i = np.arange(u.size, dtype=np.uint64) + np.uint64(lo)
    with np.errstate(over="ignore"):
        z = i + np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    xi = (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -52 - 1.0
    return (u + 1e-3 * xi.reshape(shape)).astype(np.float32)


def field_torch(shape, device, seed=SEED, plane0=0):
    """Same field generated on the device (planes plane0.. of a taller domain
    share the noise stream by linear index)."""
    import torch
    n0, n1, n2 = shape
    x0 = (torch.arange(n0, device=device, dtype=torch.float64) / (n0 - 1)).view(-1, 1, 1)
    x1 = (torch.arange(n1, device=device, dtype=torch.float64) / (n1 - 1)).view(1, -1, 1)
    x2 = (torch.arange(n2, device=device, dtype=torch.float64) / (n2 - 1)).view(1, 1, -1)
    out = torch.empty(shape, dtype=torch.float32, device=device)
    M = (1 << 64) - 1

    def srl(z, k):  # logical shift right on int64
        return (z >> k) & ((1 << (64 - k)) - 1)

    def c(v):  # python int -> wrapped int64
        v &= M
        return v - (1 << 64) if v >= (1 << 63) else v

    step = 64
    for a in range(0, n0, step):
        b = min(n0, a + step)
        u = (torch.sin(6 * math.pi * x0[a:b]) * torch.cos(4 * math.pi * x1) * torch.sin(2 * math.pi * x2)
             + 0.3 * torch.sin(40 * math.pi * x0[a:b] * x1))
        i = (torch.arange((b - a) * n1 * n2, device=device, dtype=torch.int64)
             + (plane0 + a) * n1 * n2)
        z = i + c(seed + 0x9E3779B97F4A7C15)
        z = (z ^ srl(z, 30)) * c(0xBF58476D1CE4E5B9)
        z = (z ^ srl(z, 27)) * c(0x94D049BB133111EB)
        z = z ^ srl(z, 31)
        xi = srl(z, 11).to(torch.float64) * 2.0 ** -52 - 1.0
        out[a:b] = (u + 1e-3 * xi.view(b - a, n1, n2)).to(torch.float32)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.samples:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[2 + k].lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def reference_arm(args):
    """Times the reference's own CPU implementation (oracle/_ref, SERIAL adapter,
    1 core) on a bounded sample of the workload."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import ref_x
    n = args.ref_size
    shape = (n, n, n)
    u = field_numpy(shape)
    times_c, times_d = [], []
    cr = None
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = ref_x.compress(u, ref_x.REL, TOL, S)
        t1 = time.perf_counter()
        back = ref_x.decompress(r["payload"], shape, u.dtype, ref_x.REL, TOL, S, r["norm"])
        t2 = time.perf_counter()
        if it >= args.warmup:
            times_c.append(t1 - t0)
            times_d.append(t2 - t1)
        cr = u.nbytes / r["payload"].size
    tc, td = sum(times_c) / len(times_c), sum(times_d) / len(times_d)
    value = 2 * u.nbytes / (tc + td) / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": (tc + td) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "3D fp32 513x513x513 synthetic field, relative L-inf 1e-3, Huffman lossless",
                   "sample": f"{n}^3 sample of the same field"},
        "compress_gbs": u.nbytes / tc / 1e9, "decompress_gbs": u.nbytes / td / 1e9,
        "ratio": cr,
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": 1, "kind": "reference",
                         "sample": f"MGARD-X SERIAL Compressor::Compress+Decompress on a {n}^3 fp32 sample, REL 1e-3 s=inf"},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def cpu_baseline(n=129):
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import ref_x
        if not ref_x.available():
            raise RuntimeError("oracle/_ref not built")
        u = field_numpy((n, n, n))
        t0 = time.perf_counter()
        r = ref_x.compress(u, ref_x.REL, TOL, S)
        t1 = time.perf_counter()
        ref_x.decompress(r["payload"], u.shape, u.dtype, ref_x.REL, TOL, S, r["norm"])
        t2 = time.perf_counter()
        return {"value": 2 * u.nbytes / (t2 - t0) / 1e9, "unit": "GB/s", "cores": 1,
                "kind": "reference",
                "sample": f"MGARD-X SERIAL (oracle/_ref) compress+decompress of a {n}^3 fp32 sample of the workload field, {t2 - t0:.1f} s",
                "compress_gbs": u.nbytes / (t1 - t0) / 1e9, "decompress_gbs": u.nbytes / (t2 - t1) / 1e9,
                "ratio": u.nbytes / r["payload"].size}
    except Exception as e:  # the oracle always exists; report why it could not run
        return {"value": None, "unit": "GB/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-size", type=int, default=129)
    ap.add_argument("--cpu-size", type=int, default=193)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import mgard_b200 as mg
    from mgard_b200 import _lib, sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (mgard_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # keep stdout to the one JSON line: whatever NCCL / torch print while the
        # communicator comes up (e.g. the version banner at NCCL_DEBUG >= VERSION) is
        # sent to stderr; the JSON line is written to the saved stdout at the end
        sys.stdout.flush()
        global _REAL_STDOUT
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    K = args.steps
    L = _lib.lib()

    N = int(np.prod(SHAPE))
    nbytes = N * 4
    gshape = (SHAPE[0] * world,) + SHAPE[1:]
    u = field_torch(SHAPE, dev, plane0=rank * SHAPE[0])
    cfg = mg.Config()
    cfg.dev_id = local_rank
    plan = mg.Plan(SHAPE, np.float32, config=cfg)
    cap = nbytes + 8 * (128 + cfg.huff_dict_size) + (1 << 20)
    out = torch.empty(cap, dtype=torch.uint8, device=dev)
    back = torch.empty(SHAPE, dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_compress():
        if world == 1:
            payload, norm = plan.compress(u, mg.error_bound_type.REL, TOL, S, out=out)
            return payload, norm, 0.0
        r = sharded.compress_sharded(u, gshape, TOL, S, mg.error_bound_type.REL, SHAPE[0],
                                     config=cfg, dist=dist)
        return r["records"], r["norm"], r

    def one_decompress(payload, norm):
        if world == 1:
            return plan.decompress(payload, mg.error_bound_type.REL, TOL, S, norm, out=back)
        # sharded: the rank's own record `u64 size | payload`, ABS with tol*norm
        return plan.decompress(payload[8:], mg.error_bound_type.ABS,
                               float(np.float32(TOL) * np.float32(norm)), S, norm, out=back)

    # ---- warm-up (also builds workspaces) ----
    for _ in range(W):
        payload, norm, _r = one_compress()
        one_decompress(payload, norm)
    barrier()
    launches0 = mg.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tc = td = 0.0
    barrier()
    for _ in range(K):
        ev[0].record()
        payload, norm, _r = one_compress()
        ev[1].record()
        one_decompress(payload, norm)
        ev[2].record()
        torch.cuda.synchronize()
        tc += ev[0].elapsed_time(ev[1])
        td += ev[1].elapsed_time(ev[2])
    barrier()
    clocks = sampler.stop()
    launches = mg.launch_count() - launches0
    tc /= K
    td /= K
    if world > 1:
        t = torch.tensor([tc, td], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tc, td = float(t[0]), float(t[1])
    stream_bytes = int(payload.numel())
    err = float((back - u).abs().max())
    bound = TOL * float(u.abs().max()) if world == 1 else TOL * norm
    if world > 1:
        t = torch.tensor([stream_bytes], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total_stream = float(t[0])
        e = torch.tensor([err], dtype=torch.float64, device=dev)
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        err = float(e[0])
    else:
        total_stream = stream_bytes
    value = 2 * nbytes * world / ((tc + td) * 1e-3) / 1e9

    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": tc + td, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "3D fp32 513x513x513 synthetic field per GPU, relative L-inf 1e-3 (s=inf), Huffman lossless, dict 8192, block 20480",
                   "step": "compress + decompress of the field (device resident)",
                   "l2": "input (540 MB) and coefficient/symbol arrays exceed the 126 MB L2; no explicit flush",
                   "multi_gpu": "MaxDim slabs of 513 planes along dim 0, one per rank; norm all-reduce + size all-gather" if world > 1 else "single GPU"},
        "compress_gbs": nbytes * world / (tc * 1e-3) / 1e9,
        "decompress_gbs": nbytes * world / (td * 1e-3) / 1e9,
        "compress_ms": tc, "decompress_ms": td,
        "ratio": nbytes * world / total_stream,
        "max_abs_error": err, "error_bound": bound, "bound_ok": bool(err <= bound),
        "gpu_launches": int(launches), "clocks": clocks,
    }

    if rank == 0:
        # ---- roofline of the dominant kernel (separate profiled pass) ----
        peak, peak_src = measured_peaks()
        L.mgb_profile_enable(1)
        reps = 3
        for _ in range(reps):
            payload1, norm1 = plan.compress(u, mg.error_bound_type.REL, TOL, S, out=out)
            plan.decompress(payload1, mg.error_bound_type.REL, TOL, S, norm1, out=back)
        torch.cuda.synchronize()
        L.mgb_profile_enable(0)
        import ctypes as C
        fam = []
        k = 0
        while True:
            name, n_l, tot, mx = C.c_char_p(), C.c_ulonglong(0), C.c_double(0), C.c_double(0)
            if L.mgb_profile_report(k, C.byref(name), C.byref(n_l), C.byref(tot), C.byref(mx)) != 0:
                break
            if n_l.value:
                fam.append({"kernel": name.value.decode(), "launches_per_step": n_l.value / reps,
                            "ms_per_step": tot.value / reps, "max_launch_ms": mx.value})
            k += 1
        fam.sort(key=lambda f: -f["ms_per_step"])
        line["kernel_breakdown"] = fam
        # algorithmic bytes of the largest launch of each family (finest level), fp32
        # (DESIGN.md section 4):
        nl = N
        cs = (SHAPE[0] // 2 + 1) * (SHAPE[1] // 2 + 1) * (SHAPE[2] // 2 + 1)
        alg = {
            "coef": 2 * nl * 4,                    # read the level box, write coefficients + coarse
            "restore": 2 * nl * 4 + cs * 4,        # read coefficients + coarse, write the level box
            "mass_trans": (nl + cs) * 4,           # fused f/c/r pass: read n, write n/8
            "quantize_hist": nl * 4 + nl * 2,      # read T, write u16 symbols
            "encode": nl * 2 + total_stream / world,
            "chunk_bits": nl * 2,
            "decode": total_stream / world + nl * 4,   # s=inf: dequantized while flushing
            "thomas_contig": 2 * cs * 4, "thomas_strided": 2 * cs * 4,
            "norm": nl * 4,
        }
        # DRAM bytes per launch measured by ncu --set full on the same workload
        # (profiles/r1_ncu_traffic.json, produced by scripts/make_profiles.py)
        ncu = {}
        try:
            ncu = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")))
        except Exception:
            pass

        def traffic_of(k):
            t = ncu.get(k)
            return (t["dram_read_bytes"] + t["dram_write_bytes"]) if t else None

        # the quantizer runs as two launches when the upper half of the coefficients is
        # quantized early (api.cu): its largest launch covers that share of the array
        qshare = 1.0
        for f in fam:
            if f["kernel"] == "quantize_hist" and f["launches_per_step"] > 1.5:
                first = -(-(SHAPE[0] // 2 + 1) * SHAPE[1] * SHAPE[2] // 8) * 8
                qshare = max(first, N - first) / N
        alg["quantize_hist"] *= qshare
        if "quantize_hist" in ncu and qshare < 1.0:
            ncu["quantize_hist"] = dict(ncu["quantize_hist"])
            for key in ("dram_read_bytes", "dram_write_bytes"):
                ncu["quantize_hist"][key] *= qshare

        per_kernel = []
        for f in fam:
            a = alg.get(f["kernel"])
            if a:
                ach = a / (f["max_launch_ms"] * 1e-3) / 1e9
                per_kernel.append({"kernel": f["kernel"], "achieved": ach, "frac": ach / peak,
                                   "algorithmic_bytes": a, "traffic": traffic_of(f["kernel"]),
                                   "launch_ms": f["max_launch_ms"]})
        line["roofline_kernels"] = per_kernel
        if fam:
            top = fam[0]
            a = alg.get(top["kernel"])
            if a:
                achieved = a / (top["max_launch_ms"] * 1e-3) / 1e9
                line["roofline"] = {"bound": "hbm", "kernel": top["kernel"], "achieved": achieved,
                                    "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                    "traffic": traffic_of(top["kernel"]), "peak_source": peak_src,
                                    "note": "largest (finest-level) launch of the dominant family: algorithmic bytes / "
                                            "CUDA-event duration; traffic = dram read+write bytes of that launch from "
                                            "ncu --set full (profiles/r1_ncu_traffic.json)"}
        # time-weighted DRAM efficiency over the finest-level launches:
        # sum of ncu DRAM bytes / sum of live launch durations / peak
        tb = sum(k["traffic"] for k in per_kernel if k["traffic"])
        tt = sum(k["launch_ms"] for k in per_kernel if k["traffic"]) * 1e-3
        if tt > 0:
            line["roofline"]["dram_efficiency"] = tb / tt / 1e9 / peak
        # whole-codec view on the B_alg basis of SURVEY §8d
        line["roofline_codec"] = {
            "compress_frac": (nbytes + total_stream / world) / (tc * 1e-3) / 1e9 / peak,
            "decompress_frac": (nbytes + total_stream / world) / (td * 1e-3) / 1e9 / peak,
            "basis": "B_alg = N*4 + stream bytes per direction"}

    # ---- e2e through the public host API with pinned host buffers ----
    if not args.no_e2e and world == 1:
        hin = torch.empty(SHAPE, dtype=torch.float32, pin_memory=True)
        hin.copy_(u)
        hout = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
        hback = torch.empty(SHAPE, dtype=torch.float32, pin_memory=True)
        hin_np, hout_np, hback_np = hin.numpy(), hout.numpy(), hback.numpy()
        ek = max(3, min(K, 5))
        for _ in range(2):
            s_ = mg.compress(hin_np, TOL, S, mg.error_bound_type.REL, config=cfg, out=hout_np)
            mg.decompress(s_, config=cfg, out=hback_np)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(ek):
            s_ = mg.compress(hin_np, TOL, S, mg.error_bound_type.REL, config=cfg, out=hout_np)
            mg.decompress(s_, config=cfg, out=hback_np)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        et = (t1 - t0) / ek
        e2e_err = float(np.abs(hback_np - hin_np).max())
        line["e2e"] = {"value": 2 * nbytes / et / 1e9, "unit": "GB/s",
                       "h2d_bytes_per_step": int(nbytes + s_.size),
                       "d2h_bytes_per_step": int(s_.size + nbytes),
                       "ms_per_step": et * 1e3, "max_abs_error": e2e_err,
                       "api": "mgard_b200.compress / decompress (mgard_x::compress mirror), pinned host buffers"}
    elif not args.no_e2e and world > 1:
        # sharded API with HOST buffers: every rank copies its slab in from pinned
        # memory, compresses it (norm all-reduce + size all-gather inside), copies its
        # records out; then the way back.  Copies are inside the timed region.
        hin = torch.empty(SHAPE, dtype=torch.float32, pin_memory=True)
        hin.copy_(u)
        hrec = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
        hback = torch.empty(SHAPE, dtype=torch.float32, pin_memory=True)
        du = torch.empty_like(u)
        drec = torch.empty(cap, dtype=torch.uint8, device=dev)
        ek = max(3, min(K, 5))
        rec_bytes = 0

        def e2e_step():
            du.copy_(hin, non_blocking=True)
            r = sharded.compress_sharded(du, gshape, TOL, S, mg.error_bound_type.REL, SHAPE[0],
                                         config=cfg, dist=dist)
            n = int(r["records"].numel())
            hrec[:n].copy_(r["records"], non_blocking=True)
            torch.cuda.synchronize()
            drec[:n].copy_(hrec[:n], non_blocking=True)
            b = plan.decompress(drec[8:n], mg.error_bound_type.ABS,
                                float(np.float32(TOL) * np.float32(r["norm"])), S, r["norm"], out=back)
            hback.copy_(b, non_blocking=True)
            torch.cuda.synchronize()
            return n

        for _ in range(2):
            rec_bytes = e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ek):
            rec_bytes = e2e_step()
        barrier()
        et = (time.perf_counter() - t0) / ek
        tt = torch.tensor([et, float(rec_bytes)], dtype=torch.float64, device=dev)
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        e2e_err = float((hback - hin).abs().max())
        line["e2e"] = {"value": 2 * nbytes * world / float(tmax[0]) / 1e9, "unit": "GB/s",
                       "h2d_bytes_per_step": int(nbytes * world + float(tt[1])),
                       "d2h_bytes_per_step": int(nbytes * world + float(tt[1])),
                       "ms_per_step": float(tmax[0]) * 1e3, "max_abs_error": e2e_err,
                       "api": "mgard_b200.sharded.compress_sharded / Plan.decompress per rank, pinned host "
                              "buffers, H2D + D2H inside the timed region; max over ranks"}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.cpu_size)
    if rank == 0:
        if _REAL_STDOUT is not None:
            sys.stdout.flush()
            os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())
        else:
            print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
