"""One compress + decompress of a single sub-domain through Compressor::Compress / Decompress
(for ncu captures).  Usage: prof_slab.py n0 n1 n2 [iterations]   (default: the C5 slab)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench, mgard_b200 as mg
dev = torch.device("cuda:0")
shape = tuple(int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (257, 2049, 2049)
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 1
u = bench.field_torch(shape, dev, full_n0=2049 if shape[1] == 2049 else None)
p = mg.Plan(shape, np.float32)
tol = float(np.float32(1e-3) * np.float32(1.3009516)) if shape[1] == 2049 else 1e-3
mode = mg.error_bound_type.ABS if shape[1] == 2049 else mg.error_bound_type.REL
for _ in range(iters):
    payload, norm = p.compress(u, mode, tol, float("inf"))
    back = p.decompress(payload, mode, tol, float("inf"), norm)
torch.cuda.synchronize()
print("launches", mg.launch_count(), "CR", u.numel() * 4 / payload.numel(), "err", float((back - u).abs().max()))
