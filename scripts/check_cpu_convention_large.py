"""MGARD-CPU convention at the bench size (513^3 fp32): stage timings on the GPU next
to the reference CPU build's decomposition (OpenMP, all host threads)."""
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


def gpu_ms(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 513
    x = torch.linspace(0, 1, n, device="cuda", dtype=torch.float64)
    x0, x1, x2 = torch.meshgrid(x, x, x, indexing="ij")
    du = (torch.sin(6 * math.pi * x0) * torch.cos(4 * math.pi * x1) * torch.sin(2 * math.pi * x2)
          + 0.3 * torch.sin(40 * math.pi * x0 * x1)).float().contiguous()
    del x0, x1, x2
    H = mc.TensorMeshHierarchy((n, n, n), None, np.float32)
    r = {"case": f"{n}^3 fp32 uniform, s=inf, tol 1e-3 (MGARD-CPU convention)", "L": H.L, "bytes": du.numel() * 4}
    coef = H.decompose(du)
    r["gpu_decompose_ms"] = gpu_ms(lambda: H.decompose(du))
    r["gpu_recompose_ms"] = gpu_ms(lambda: H.recompose(coef))
    r["gpu_quantize_ms"] = gpu_ms(lambda: H.quantize(coef, math.inf, 1e-3))
    r["gpu_decompose_GBps"] = r["bytes"] / r["gpu_decompose_ms"] / 1e6
    t = time.perf_counter()
    blob = mc.compress(H, du, math.inf, 1e-3, mc.CPU_HUFFMAN_ZSTD)
    r["compress_huffman_zstd_device_in_s"] = time.perf_counter() - t
    r["ratio"] = r["bytes"] / len(blob)
    t = time.perf_counter()
    back = mc.decompress(blob)
    r["decompress_s"] = time.perf_counter() - t
    r["linf_error"] = float(np.abs(back - du.cpu().numpy()).max())
    if ref_cpu.available() and n <= 513:
        u = du.cpu().numpy()
        t = time.perf_counter()
        c_ref = ref_cpu.decompose(u)
        r["ref_cpu_decompose_s"] = time.perf_counter() - t
        r["ref_cpu_threads"] = os.cpu_count()
        r["coefficients_bit_identical"] = bool(np.array_equal(coef.cpu().numpy().view(np.uint8), c_ref.view(np.uint8)))
    print(json.dumps(r, indent=1))


if __name__ == "__main__":
    main()
