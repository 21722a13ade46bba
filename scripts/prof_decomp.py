"""Runs decompose + recompose of one sub-domain twice (for ncu captures of the level kernels).
Usage: prof_decomp.py n0 n1 n2"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench, mgard_b200 as mg
dev = torch.device("cuda:0")
shape = tuple(int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (513, 513, 513)
u = bench.field_torch(shape, dev)
p = mg.Plan(shape, np.float32)
for _ in range(2):
    c = p.decompose(u)
    b = p.recompose(c)
torch.cuda.synchronize()
print("max err", float((b - u).abs().max()))
