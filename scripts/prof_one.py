"""Runs one compress + decompress of the bench workload (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench, mgard_b200 as mg
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 513
shape = (n, n, n)
u = bench.field_torch(shape, dev)
p = mg.Plan(shape, np.float32)
for _ in range(2):
    payload, norm = p.compress(u, mg.error_bound_type.REL, 1e-3, float("inf"))
    back = p.decompress(payload, mg.error_bound_type.REL, 1e-3, float("inf"), norm)
torch.cuda.synchronize()
print("CR", u.numel() * 4 / payload.numel(), "err", float((back - u).abs().max()))
