"""MGARD-CPU convention on the GPU next to the reference CPU build (SURVEY 8d:
C1 129^3 fp64 and C3 1000^2 fp32 non-uniform): parity, ratio, error, timings.
Writes one JSON object to stdout.  GPU box only; uses oracle/_ref when present.
"""
import json
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mgard_b200.cpu as mc  # noqa: E402
import ref_cpu  # noqa: E402


def c1():
    n = 129
    x = np.arange(n) / (n - 1)
    x0, x1, x2 = np.meshgrid(x, x, x, indexing="ij")
    u = np.sin(2 * np.pi * x0) * np.cos(3 * np.pi * x1) + 0.5 * np.sin(5 * np.pi * x2) + 0.25 * x0 * x1
    return "C1 129^3 fp64 ABS 1e-4 s=inf", u, None, math.inf, 1e-4


def c3():
    n = 1000
    coords = []
    for k in (7, 11):
        h = 1 + 0.5 * np.sin(2 * np.pi * k * np.arange(n - 1) / 999)
        xx = np.concatenate([[0.0], np.cumsum(h)])
        coords.append((xx / xx[-1]).astype(np.float32))
    x0, x1 = np.meshgrid(coords[0].astype(np.float64), coords[1].astype(np.float64), indexing="ij")
    u = (np.exp(-8 * ((x0 - .5) ** 2 + (x1 - .4) ** 2)) + 0.1 * np.sin(30 * x0)).astype(np.float32)
    return "C3 1000^2 fp32 non-uniform ABS 1e-2 s=0 (CPU convention)", u, coords, 0.0, 1e-2


def gpu_ms(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def wall_s(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t)
    return best


def main():
    out = []
    for name, u, coords, s, tol in (c1(), c3()):
        H = mc.TensorMeshHierarchy(u.shape, coords, u.dtype)
        du = torch.from_numpy(u).cuda()
        r = {"case": name, "L": H.L, "bytes": int(u.nbytes)}
        coef = H.decompose(du)
        q = H.quantize(coef, s, tol)
        r["gpu_decompose_ms"] = gpu_ms(lambda: H.decompose(du))
        r["gpu_quantize_ms"] = gpu_ms(lambda: H.quantize(coef, s, tol))
        r["gpu_recompose_ms"] = gpu_ms(lambda: H.recompose(coef))
        r["gpu_decompose_GBps"] = u.nbytes / r["gpu_decompose_ms"] / 1e6
        blob = mc.compress(H, u, s, tol)
        blob2 = mc.compress(H, u, s, tol, mc.CPU_HUFFMAN_ZSTD)
        r["huffman_zstd_ratio"] = u.nbytes / len(blob2)
        r["huffman_zstd_compress_total_s"] = wall_s(lambda: mc.compress(H, u, s, tol, mc.CPU_HUFFMAN_ZSTD))
        r["huffman_zstd_decompress_total_s"] = wall_s(lambda: mc.decompress(blob2))
        r["compress_total_s"] = wall_s(lambda: mc.compress(H, u, s, tol))
        r["decompress_total_s"] = wall_s(lambda: mc.decompress(blob))
        r["ratio"] = u.nbytes / len(blob)
        back = mc.decompress(blob)
        r["linf_error"] = float(np.abs(back.astype(np.float64) - u).max())
        r["rms_error"] = float(math.sqrt(np.mean((back.astype(np.float64) - u) ** 2)))
        if ref_cpu.available():
            c_ref = ref_cpu.decompose(u, coords)
            q_ref = ref_cpu.quantize(c_ref, u.shape, s, tol, coords)
            r["coefficients_bit_identical"] = bool(np.array_equal(coef.cpu().numpy().view(np.uint8), c_ref.view(np.uint8)))
            r["quanta_identical"] = bool(np.array_equal(q.cpu().numpy(), q_ref))
            pay = ref_cpu.zlib_compress(q_ref).tobytes()
            r["payload_identical"] = blob.endswith(pay)
            r["ref_cpu_decompose_s"] = wall_s(lambda: ref_cpu.decompose(u, coords), 2)
            r["ref_cpu_quantize_s"] = wall_s(lambda: ref_cpu.quantize(c_ref, u.shape, s, tol, coords), 2)
            r["ref_cpu_zlib_s"] = wall_s(lambda: ref_cpu.zlib_compress(q_ref), 1)
            r["ref_cpu_recompose_s"] = wall_s(lambda: ref_cpu.recompose(c_ref, u.shape, coords), 2)
            r["ref_cpu_threads"] = os.cpu_count()  # built with -fopenmp (line loops)
            r["decompose_speedup_vs_ref_cpu"] = r["ref_cpu_decompose_s"] * 1e3 / r["gpu_decompose_ms"]
        out.append(r)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
