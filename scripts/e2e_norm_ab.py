"""A/B of the host-buffer path with and without the fused norm (same process)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench, mgard_b200 as mg
dev = torch.device("cuda:0")
n = 513
u = bench.field_torch((n, n, n), dev).cpu().numpy()
mg.pin_memory(u)
for rep in range(3):
    for flag in ("0", "1"):
        if flag == "1":
            os.environ["MGB_NO_FUSED_NORM"] = "1"
        else:
            os.environ.pop("MGB_NO_FUSED_NORM", None)
        ts = []
        for i in range(6):
            torch.cuda.synchronize()
            t = time.perf_counter()
            st = mg.compress(u, 1e-3, float("inf"), mg.error_bound_type.REL)
            ts.append(time.perf_counter() - t)
        print("no_fused_norm", flag, "compress host->host ms:", " ".join(f"{x*1e3:.1f}" for x in ts), flush=True)
