#!/usr/bin/env python
"""bench.py — MGARD-X hot path on B200: compress + decompress throughput.

Workload at every N (BASELINE.json configs[4], SURVEY.md 8d/8e "C5"): 3-D fp32
2049^3 synthetic field (34.4 GB), relative L-inf bound 1e-3 (s = inf), Huffman
lossless, dict 8192, block 20480, MaxDim-decomposed along dim 0 in 8 sub-domains of
257 planes (7 x 257 + 250, DomainDecomposer.hpp:124-169).  GPU g of N owns the
sub-domains [g*8/N, (g+1)*8/N): STRONG scaling, the stream does not depend on N.
One "step" = one compress + one decompress of the whole domain through
mgb_compress_sharded / mgb_decompress_sharded (global norm all-reduce + size
all-gather over NCCL inside the timed region).

  value  (device resident)  original bytes through the codec per second:
         2 * 34.4 GB / (t_compress + t_decompress), max over ranks; inputs in HBM.
  e2e    the same calls with pinned HOST buffers (H2D of the field, D2H of the
         records, and back, inside the timed region; three-stream pipeline).
  c2     (N = 1 only) BASELINE configs[1]: 513^3 fp32, same bound, one sub-domain,
         through Compressor::Compress / Decompress (mgb_*_lowlevel), with the
         per-kernel roofline table.

`--impl reference` times the UNMODIFIED reference (MGARD-X SERIAL adapter built from
/root/reference as oracle/_ref) on the host: one 257^3 block of the same field per
host core, all cores at once (what the reference's CPU pipeline does with one
sub-domain per thread, CPUPipelines.hpp:88-135).
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
GSHAPE = (2049, 2049, 2049)
DD_SIZE = 257
SHAPE = (513, 513, 513)  # C2
TOL, S = 1e-3, float("inf")
SEED = 2049
_REAL_STDOUT = None  # saved stdout fd when N > 1 (see main)
WORKLOAD = ("3D fp32 2049x2049x2049 synthetic field (34.4 GB), relative L-inf 1e-3 (s=inf), Huffman lossless, "
            "dict 8192, block 20480, MaxDim sub-domains of 257 planes (7x257+250) spread over the GPUs")


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


def field_torch(shape, device, seed=SEED, plane0=0, full_n0=None):
    """Same field generated on the device: planes [plane0, plane0 + shape[0]) of a
    domain with full_n0 planes (coordinates and noise index of the full domain)."""
    import torch
    n0, n1, n2 = shape
    full_n0 = full_n0 or n0
    x0 = ((torch.arange(n0, device=device, dtype=torch.float64) + plane0) / (full_n0 - 1)).view(-1, 1, 1)
    x1 = (torch.arange(n1, device=device, dtype=torch.float64) / (n1 - 1)).view(1, -1, 1)
    x2 = (torch.arange(n2, device=device, dtype=torch.float64) / (n2 - 1)).view(1, 1, -1)
    out = torch.empty(shape, dtype=torch.float32, device=device)
    M = (1 << 64) - 1

    def srl(z, k):  # logical shift right on int64
        return (z >> k) & ((1 << (64 - k)) - 1)

    def c(v):  # python int -> wrapped int64
        v &= M
        return v - (1 << 64) if v >= (1 << 63) else v

    step = max(1, min(64, (1 << 26) // (n1 * n2)))
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


def host_link(torch, dist, world, dev, h_src, h_dst):
    """Pinned-memory copy rates with every rank copying at the same time: H2D, D2H and both
    directions together (up to 1 GiB each, max time over ranks)."""
    n = min(h_src.numel(), h_dst.numel(), 1 << 28)
    src, dst = h_src.view(-1)[:n], h_dst.view(-1)[:n]
    d_a = torch.empty(n, dtype=torch.float32, device=dev)
    d_b = torch.zeros(n, dtype=torch.float32, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {}
    for kind in ("h2d", "d2h", "both"):
        def once():
            if kind in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_a.copy_(src, non_blocking=True)
            if kind in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    dst.copy_(d_b, non_blocking=True)
        once()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            once()
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / 3], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        out[kind] = round(n * 4 * (2 if kind == "both" else 1) * world / float(dt) / 1e9, 1)
    out["unit"] = "GB/s, all ranks together"
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


def _ref_block(arg):
    """One worker of the reference arm: MGARD-X SERIAL compress + decompress of one
    n^3 block of the workload field (planes from `plane0` of the 2049^3 domain)."""
    n, plane0, reps = arg
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import baseline_fields as bf
    import ref_x
    u = bf.c2_like(GSHAPE, plane0, n, crop=(n, n))
    tc = td = 0.0
    size = 0
    err = 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        r = ref_x.compress(u, ref_x.REL, TOL, S)
        t1 = time.perf_counter()
        back = ref_x.decompress(r["payload"], u.shape, u.dtype, ref_x.REL, TOL, S, r["norm"])
        t2 = time.perf_counter()
        tc += t1 - t0
        td += t2 - t1
        size = int(r["payload"].size)
        err = float(np.abs(back - u).max() / np.abs(u).max())
    return u.nbytes, tc / reps, td / reps, size, err


def reference_cpu(n, cores, steps=1):
    """All host cores at once, one n^3 block per core (the reference's own CPU
    pipeline runs one sub-domain per thread, CPUPipelines.hpp:88-135).  Returns the
    aggregate rate of (compress + decompress)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_ref_block, [(n, (37 * k) % (GSHAPE[0] - n), steps) for k in range(cores)])
        wall = time.perf_counter() - t0
    nbytes = sum(r[0] for r in res)
    # blocks run concurrently: the step takes as long as the slowest worker
    tc, td = max(r[1] for r in res), max(r[2] for r in res)
    return {"bytes": nbytes, "tc": tc, "td": td, "wall": wall, "ratio": nbytes / sum(r[3] for r in res),
            "rel_err": max(r[4] for r in res)}


def reference_arm(args):
    """Times the reference's own CPU implementation (oracle/_ref, MGARD-X SERIAL
    adapter) on a bounded sample of the workload: one 257^3 block per host core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    n = args.ref_size
    r = reference_cpu(n, cores, steps=max(1, min(args.steps, 3)))
    value = 2 * r["bytes"] / (r["tc"] + r["td"]) / 1e9
    sample = (f"{cores} blocks of {n}^3 fp32 of the 2049^3 workload field, one per host core, MGARD-X SERIAL "
              f"Compressor::Compress+Decompress, REL 1e-3 s=inf")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": (r["tc"] + r["td"]) * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "compress_gbs": r["bytes"] / r["tc"] / 1e9, "decompress_gbs": r["bytes"] / r["td"] / 1e9,
        "ratio": r["ratio"],
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def cpu_baseline(n=257):
    """Reference arms on the box's host cores: (i) MGARD-X SERIAL, one n^3 block of
    the workload field per core; (ii) MGARD-CPU mgard::compress / decompress stages
    (OpenMP, all cores) on BASELINE config 1."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    cores = os.cpu_count() or 1
    try:
        import ref_x
        if not ref_x.available():
            raise RuntimeError("oracle/_ref not built")
        r = reference_cpu(n, cores)
        out = {"value": 2 * r["bytes"] / (r["tc"] + r["td"]) / 1e9, "unit": "GB/s", "cores": cores,
               "kind": "reference",
               "sample": f"MGARD-X SERIAL (oracle/_ref) compress+decompress, {cores} blocks of {n}^3 fp32 of the "
                         f"workload field, one per host core at once, {r['wall']:.1f} s",
               "compress_gbs": r["bytes"] / r["tc"] / 1e9, "decompress_gbs": r["bytes"] / r["td"] / 1e9,
               "ratio": r["ratio"]}
    except Exception as e:  # the oracle always exists; report why it could not run
        out = {"value": None, "unit": "GB/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import baseline_fields as bf
        import ref_cpu
        u = bf.c1()
        t0 = time.perf_counter()
        c = ref_cpu.decompose(u)
        q = ref_cpu.quantize(c, u.shape, float("inf"), 1e-4)
        blob = ref_cpu.huffman_zstd_compress(q) if ref_cpu.zstd_available() else ref_cpu.zlib_compress(q)
        t1 = time.perf_counter()
        q2 = (ref_cpu.huffman_zstd_decompress(blob, q.size) if ref_cpu.zstd_available()
              else ref_cpu.zlib_decompress(blob, q.nbytes))
        back = ref_cpu.recompose(ref_cpu.dequantize(np.asarray(q2).reshape(-1), u.shape, u.dtype, float("inf"), 1e-4),
                                 u.shape)
        t2 = time.perf_counter()
        out["mgard_cpu"] = {"value": 2 * u.nbytes / (t2 - t0) / 1e9, "unit": "GB/s",
                            "cores": int(os.environ.get("OMP_NUM_THREADS", cores)), "kind": "reference",
                            "sample": "MGARD-CPU mgard::compress + decompress stages (reference templates, OpenMP) on "
                                      "BASELINE config 1, 129^3 fp64 ABS 1e-4 s=inf, Huffman+zstd payload",
                            "compress_gbs": u.nbytes / (t1 - t0) / 1e9, "decompress_gbs": u.nbytes / (t2 - t1) / 1e9,
                            "ratio": u.nbytes / len(blob),
                            "max_abs_error": float(np.abs(np.asarray(back).reshape(u.shape) - u).max())}
    except Exception as e:
        out["mgard_cpu"] = {"value": None, "sample": f"unavailable: {e}"}
    return out


def kernel_table(L, reps):
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
    return fam


def algorithmic_bytes(shape, stream_bytes):
    """Per kernel family: algorithmic bytes of its largest (finest-level) launch on one
    sub-domain of `shape`, fp32."""
    nl = shape[0] * shape[1] * shape[2]
    cs = (shape[0] // 2 + 1) * (shape[1] // 2 + 1) * (shape[2] // 2 + 1)
    return {
        "coef": 2 * nl * 4,                    # read the level box, write coefficients + coarse
        "restore": 2 * nl * 4 + cs * 4,        # read coefficients + coarse, write the level box
        "mass_trans": (nl + cs) * 4,           # fused f/c/r pass: read n, write n/8
        "quantize_hist": nl * 4 + nl * 2,      # read T, write u16 symbols
        "encode": nl * 2 + stream_bytes,
        "chunk_bits": nl * 2,
        "decode": stream_bytes + nl * 4,       # s=inf: dequantized while flushing
        "thomas_contig": 2 * cs * 4, "thomas_strided": 2 * cs * 4,
        "norm": nl * 4,
    }


def roofline_tables(fam, shape, stream_bytes, peak, peak_src, ncu_file, nsub=1):
    """roofline (dominant kernel) + per-kernel list from a kernel_table."""
    alg = algorithmic_bytes(shape, stream_bytes)
    ncu = {}
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", ncu_file)))
    except Exception:
        pass

    def traffic_of(k):
        t = ncu.get(k)
        return (t["dram_read_bytes"] + t["dram_write_bytes"]) if t else None

    # the quantizer runs as two launches when the upper half of the coefficients is
    # quantized early (api.cu): its largest launch covers that share of the array
    N = shape[0] * shape[1] * shape[2]
    qshare = 1.0
    for f in fam:
        if f["kernel"] == "quantize_hist" and f["launches_per_step"] / nsub > 1.5:  # nsub sub-domains per step
            first = -(-(shape[0] // 2 + 1) * shape[1] * shape[2] // 8) * 8
            qshare = max(first, N - first) / N
    alg["quantize_hist"] *= qshare
    per_kernel = []
    for f in fam:
        a = alg.get(f["kernel"])
        if a and f["max_launch_ms"] * 1e-3 * peak * 1e9 * 1.2 < a:
            continue  # not that kernel's full pass (e.g. the norm as a by-product: only its last step is a launch)
        if a:
            ach = a / (f["max_launch_ms"] * 1e-3) / 1e9
            per_kernel.append({"kernel": f["kernel"], "achieved": ach, "frac": ach / peak,
                               "algorithmic_bytes": a, "traffic": traffic_of(f["kernel"]),
                               "launch_ms": f["max_launch_ms"]})
    roof = None
    if fam:
        top = fam[0]
        a = alg.get(top["kernel"])
        if a:
            achieved = a / (top["max_launch_ms"] * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": top["kernel"], "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic_of(top["kernel"]), "peak_source": peak_src,
                    "note": "largest (finest-level) launch of the dominant family on one sub-domain: algorithmic "
                            "bytes / CUDA-event duration measured in this run; traffic = dram read+write bytes of "
                            f"that launch from ncu --set full (profiles/{ncu_file}), null if not captured"}
            tb = sum(k["traffic"] for k in per_kernel if k["traffic"])
            tt = sum(k["launch_ms"] for k in per_kernel if k["traffic"]) * 1e-3
            if tt > 0:
                roof["dram_efficiency"] = tb / tt / 1e9 / peak
    return roof, per_kernel


def bench_c2(mg, L, dev, cfg, W, K, peak, peak_src):
    """BASELINE configs[1] on one GPU: 513^3 fp32, one sub-domain, device resident."""
    import numpy as np
    import torch
    N = int(np.prod(SHAPE))
    nbytes = N * 4
    u = field_torch(SHAPE, dev)
    plan = mg.Plan(SHAPE, np.float32, config=cfg)
    out = torch.empty(nbytes + 8 * (128 + cfg.huff_dict_size) + (1 << 20), dtype=torch.uint8, device=dev)
    back = torch.empty(SHAPE, dtype=torch.float32, device=dev)
    for _ in range(W):
        payload, norm = plan.compress(u, mg.error_bound_type.REL, TOL, S, out=out)
        plan.decompress(payload, mg.error_bound_type.REL, TOL, S, norm, out=back)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tc = td = 0.0
    for _ in range(K):
        ev[0].record()
        payload, norm = plan.compress(u, mg.error_bound_type.REL, TOL, S, out=out)
        ev[1].record()
        plan.decompress(payload, mg.error_bound_type.REL, TOL, S, norm, out=back)
        ev[2].record()
        torch.cuda.synchronize()
        tc += ev[0].elapsed_time(ev[1])
        td += ev[1].elapsed_time(ev[2])
    tc /= K
    td /= K
    err = float((back - u).abs().max())
    bound = TOL * float(u.abs().max())
    stream = int(payload.numel())
    L.mgb_profile_enable(1)
    reps = 3
    for _ in range(reps):
        payload1, norm1 = plan.compress(u, mg.error_bound_type.REL, TOL, S, out=out)
        plan.decompress(payload1, mg.error_bound_type.REL, TOL, S, norm1, out=back)
    torch.cuda.synchronize()
    L.mgb_profile_enable(0)
    fam = kernel_table(L, reps)
    roof, per_kernel = roofline_tables(fam, SHAPE, stream, peak, peak_src, "r2_ncu_c2_traffic.json")
    res = {"workload": "3D fp32 513x513x513 synthetic field, relative L-inf 1e-3 (s=inf), Huffman lossless, one "
                       "sub-domain, Compressor::Compress / Decompress device resident",
           "value": 2 * nbytes / ((tc + td) * 1e-3) / 1e9, "unit": "GB/s", "compress_ms": tc, "decompress_ms": td,
           "compress_gbs": nbytes / (tc * 1e-3) / 1e9, "decompress_gbs": nbytes / (td * 1e-3) / 1e9,
           "ratio": nbytes / stream, "max_abs_error": err, "error_bound": bound, "bound_ok": bool(err <= bound),
           "steps": K, "warmup": W, "roofline": roof, "roofline_kernels": per_kernel, "kernel_breakdown": fam,
           "roofline_codec": {"compress_frac": (nbytes + stream) / (tc * 1e-3) / 1e9 / peak,
                              "decompress_frac": (nbytes + stream) / (td * 1e-3) / 1e9 / peak,
                              "basis": "B_alg = N*4 + stream bytes per direction"}}
    del plan, u, out, back
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-size", type=int, default=257)
    ap.add_argument("--cpu-size", type=int, default=257)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-c2", action="store_true")
    ap.add_argument("--domain", type=int, default=GSHAPE[0], help="edge of the cubic domain (default 2049)")
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
    peak, peak_src = measured_peaks()

    gshape = (args.domain,) * 3
    ext = sharded.partition(gshape[0], DD_SIZE)
    comm = sharded.Comm(dist if world > 1 else None)
    first, count = comm.owned(len(ext))
    plane0, planes = sum(ext[:first]), sum(ext[first:first + count])
    lshape = (planes,) + gshape[1:]
    n_total = int(np.prod(gshape))
    total_bytes = n_total * 4
    local_bytes = int(np.prod(lshape)) * 4
    B = {"u": field_torch(lshape, dev, plane0=plane0, full_n0=gshape[0])}  # device buffers (freed before e2e)
    cfg = mg.Config()
    cfg.dev_id = local_rank
    cap = int(local_bytes * 0.4) + count * (8 * (128 + cfg.huff_dict_size) + (1 << 20))
    B["out"] = torch.empty(cap, dtype=torch.uint8, device=dev)
    B["back"] = torch.empty(lshape, dtype=torch.float32, device=dev)
    REL = mg.error_bound_type.REL

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_compress(src=None, dst=None):
        return sharded.compress_sharded_native(B["u"] if src is None else src, gshape, TOL, S, REL, DD_SIZE,
                                               config=cfg, comm=comm, out=B["out"] if dst is None else dst)

    def one_decompress(r, dst=None):
        return sharded.decompress_sharded_native(r["header"], r["records"], B["back"] if dst is None else dst,
                                                 config=cfg, comm=comm)

    # ---- warm-up (also builds workspaces) ----
    for _ in range(W):
        r = one_compress()
        one_decompress(r)
    barrier()
    launches0 = mg.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # The K steps are timed as a whole on every rank (CUDA events, barrier + synchronize on
    # both sides) and the MAX over ranks is the job's time.  The two phases are timed per
    # step as well, but only to split that time: ranks meet once per step (in the norm
    # all-reduce of the compression), so a rank whose sub-domain decodes faster spends the
    # difference WAITING inside its next compression - per-phase maxima over ranks would
    # count that wait twice.  compress_ms / decompress_ms are therefore means over ranks
    # (their sum is the step time of every rank), scaled to the job's time.
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    tc = td = 0.0
    barrier()
    ev[3].record()
    for _ in range(K):
        ev[0].record()
        r = one_compress()
        ev[1].record()
        one_decompress(r)
        ev[2].record()
        torch.cuda.synchronize()
        tc += ev[0].elapsed_time(ev[1])
        td += ev[1].elapsed_time(ev[2])
    ev[4].record()
    barrier()
    t_all = ev[3].elapsed_time(ev[4]) / K
    clocks = sampler.stop()
    launches = mg.launch_count() - launches0
    tc /= K
    td /= K
    err = 0.0
    for a in range(0, planes, 64):
        err = max(err, float((B["back"][a:a + 64] - B["u"][a:a + 64]).abs().max()))
    norm = r["norm"]
    phase_max = [tc, td]
    if world > 1:
        t = torch.tensor([t_all, tc, td, err], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_all, err = float(t[0]), float(t[3])
        phase_max = [float(t[1]), float(t[2])]
        t = torch.tensor([tc, td], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        tc, td = float(t[0]) / world, float(t[1]) / world
    # split the job's step time in the proportion of the mean phase times
    tc, td = t_all * tc / (tc + td), t_all * td / (tc + td)
    total_stream = int(r["total"])
    bound = TOL * norm
    value = 2 * total_bytes / ((tc + td) * 1e-3) / 1e9

    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": tc + td, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD if args.domain == GSHAPE[0] else f"{args.domain}^3 variant of: " + WORKLOAD,
                   "step": "compress + decompress of the whole domain (device resident), mgb_compress_sharded / "
                           "mgb_decompress_sharded",
                   "l2": "every sub-domain (4.3 GB) and its coefficient / symbol arrays exceed the 126 MB L2; no "
                         "explicit flush",
                   "multi_gpu": f"GPU g owns sub-domains [g*8/N, (g+1)*8/N) ({count} here); NCCL all-reduce of the "
                                "per-sub-domain {max|u|, sum u^2} pairs (device-ordered) + all-gather of container "
                                "sizes" if world > 1 else "single GPU: all 8 sub-domains"},
        "compress_gbs": total_bytes / (tc * 1e-3) / 1e9,
        "decompress_gbs": total_bytes / (td * 1e-3) / 1e9,
        "compress_ms": tc, "decompress_ms": td,
        "phase_ms_max_over_ranks": {"compress": phase_max[0], "decompress": phase_max[1],
                                    "note": "per-phase maxima include the wait for the slowest rank's previous "
                                            "phase; they do not add up to ms_per_step"},
        "ratio": total_bytes / total_stream, "stream_bytes": total_stream,
        "max_abs_error": err, "error_bound": bound, "bound_ok": bool(err <= bound),
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline_codec": {
            "compress_frac": (total_bytes + total_stream) / (tc * 1e-3) / 1e9 / (peak * world),
            "decompress_frac": (total_bytes + total_stream) / (td * 1e-3) / 1e9 / (peak * world),
            "basis": "B_alg = N*4 + stream bytes per direction, against n_gpus x the measured copy bandwidth"},
    }

    # ---- roofline of the dominant kernel: per-kernel CUDA events, one more pass ----
    L.mgb_profile_enable(1)
    r1 = one_compress()
    one_decompress(r1)
    torch.cuda.synchronize()
    L.mgb_profile_enable(0)
    if rank == 0:
        line["kernel_breakdown"] = kernel_table(L, 1)
    # ... and once more with the early quantization switched off (api.cu: the upper half of
    # the coefficients is otherwise quantized on a side stream WHILE the coarse levels are
    # decomposed, which stretches those small launches): every kernel alone on the GPU, so
    # the longest launch of a family is its finest-level launch
    os.environ["MGB_NO_EARLY_QUANTIZE"] = "1"
    L.mgb_profile_enable(1)
    r1 = one_compress()
    one_decompress(r1)
    torch.cuda.synchronize()
    L.mgb_profile_enable(0)
    del os.environ["MGB_NO_EARLY_QUANTIZE"]
    if rank == 0:
        fam = kernel_table(L, 1)
        sub_stream = int(r1["records"].numel()) / max(count, 1)
        roof, per_kernel = roofline_tables(fam, (ext[first],) + gshape[1:], sub_stream, peak, peak_src,
                                           "r2_ncu_c5_traffic.json", nsub=max(count, 1))
        if roof:
            line["roofline"] = roof
        line["roofline_kernels"] = per_kernel

    # ---- e2e: the same calls with pinned HOST buffers ----
    if not args.no_e2e:
        hin = torch.empty(lshape, dtype=torch.float32, pin_memory=True)
        hin.copy_(B["u"])
        hrec = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
        hback = torch.empty(lshape, dtype=torch.float32, pin_memory=True)
        hin_np, hrec_np, hback_np = hin.numpy(), hrec.numpy(), hback.numpy()
        del r, r1
        B.clear()
        torch.cuda.empty_cache()
        ek = 2 if total_bytes > (8 << 30) else max(3, min(K, 5))

        def e2e_step():
            rr = one_compress(hin_np, hrec_np)
            one_decompress(rr, hback_np)
            return rr

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ek):
            rr = e2e_step()
        barrier()
        et = (time.perf_counter() - t0) / ek
        rec_bytes = float(len(rr["records"]))
        e2e_err = 0.0  # every 16th plane and the last one (the host pass over 2 x 34 GB is the slow part)
        for a in list(range(0, planes, 16)) + [planes - 1]:
            e2e_err = max(e2e_err, float(np.abs(hback_np[a] - hin_np[a]).max()))
        if world > 1:
            tt = torch.tensor([et, e2e_err], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            et, e2e_err = float(tt[0]), float(tt[1])
            tb = torch.tensor([rec_bytes], dtype=torch.float64, device=dev)
            dist.all_reduce(tb, op=dist.ReduceOp.SUM)
            rec_bytes = float(tb[0])
        line["e2e"] = {"value": 2 * total_bytes / et / 1e9, "unit": "GB/s",
                       "h2d_bytes_per_step": int(total_bytes + rec_bytes),
                       "d2h_bytes_per_step": int(rec_bytes + total_bytes),
                       "ms_per_step": et * 1e3, "max_abs_error": e2e_err, "steps": ek,
                       "api": "mgb_compress_sharded / mgb_decompress_sharded with pinned host buffers: H2D of "
                              "sub-domain k+1, compute of k and D2H of k-1 overlap on three streams; max over ranks"}
        # what the host side of the box gives ALL ranks at once (pinned copies of the same
        # buffers, no codec): the ceiling of e2e
        line["e2e"]["host_link"] = host_link(torch, dist if world > 1 else None, world, dev, hin, hback)
        del hin, hrec, hback

    if rank == 0 and world == 1 and not args.no_c2:
        B.clear()
        torch.cuda.empty_cache()
        mg.release_cache()
        line["c2"] = bench_c2(mg, L, dev, cfg, W, K, peak, peak_src)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.cpu_size)
    if rank == 0:
        if _REAL_STDOUT is not None:
            sys.stdout.flush()
            os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())
        else:
            print(json.dumps(line))
    comm.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
