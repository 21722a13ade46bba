"""Per-kernel CUDA-event breakdown of one compress + decompress for a BASELINE config
(C1, C3, C4, C5 of scripts/check_configs.py).  Usage: breakdown.py C4"""
import os, sys, math, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mgard_b200 as mg
from mgard_b200 import _lib
import bench
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "C4"
INF = float("inf")
coords = None
if which == "C4":
    shape = (8, 16395, 39, 39)
    i0 = torch.arange(shape[0], device=dev, dtype=torch.float64).view(-1, 1, 1, 1)
    x1 = torch.linspace(0, 1, shape[1], device=dev, dtype=torch.float64).view(1, -1, 1, 1)
    x2 = torch.linspace(0, 1, shape[2], device=dev, dtype=torch.float64).view(1, 1, -1, 1)
    x3 = torch.linspace(0, 1, shape[3], device=dev, dtype=torch.float64).view(1, 1, 1, -1)
    g = sum((1.0 / k) * torch.sin(2 * math.pi * (2 * k + 1) * x1) for k in range(1, 6))
    u = ((1 + 0.1 * i0) * g * torch.exp(-((x2 - .5) ** 2 + (x3 - .5) ** 2) / 0.08)).contiguous()
    mode, tol, s = mg.error_bound_type.REL, 1e-3, 0.0
elif which == "C5":
    u = bench.field_torch((257, 2049, 2049), dev); mode, tol, s = mg.error_bound_type.ABS, 1.3e-3, INF
elif which == "C1":
    n = 129
    x = [torch.linspace(0, 1, n, dtype=torch.float64, device=dev) for _ in range(3)]
    X0, X1, X2 = torch.meshgrid(*x, indexing="ij")
    u = (torch.sin(2 * math.pi * X0) * torch.cos(3 * math.pi * X1) + 0.5 * torch.sin(5 * math.pi * X2) + 0.25 * X0 * X1).contiguous()
    mode, tol, s = mg.error_bound_type.ABS, 1e-4, INF
else:
    raise SystemExit("unknown config")
npdt = np.float32 if u.dtype == torch.float32 else np.float64
plan = mg.Plan(tuple(u.shape), npdt, coords=coords)
for _ in range(2):
    payload, norm = plan.compress(u, mode, tol, s)
    back = plan.decompress(payload, mode, tol, s, norm)
torch.cuda.synchronize()
L = _lib.lib()
L.mgb_profile_enable(1)
payload, norm = plan.compress(u, mode, tol, s)
back = plan.decompress(payload, mode, tol, s, norm)
torch.cuda.synchronize()
L.mgb_profile_enable(0)
k = 0
rows = []
while True:
    name, n_l, tot, mx = C.c_char_p(), C.c_ulonglong(0), C.c_double(0), C.c_double(0)
    if L.mgb_profile_report(k, C.byref(name), C.byref(n_l), C.byref(tot), C.byref(mx)) != 0:
        break
    if n_l.value:
        rows.append((tot.value, name.value.decode(), n_l.value, mx.value))
    k += 1
for tot, name, n, mx in sorted(rows, reverse=True):
    print(f"{name:18s} launches {n:4d} total {tot:8.3f} ms  max {mx:8.3f} ms")
print("sum", sum(r[0] for r in rows), "ms; bytes", u.numel() * u.element_size() / 1e6, "MB; ratio", u.numel() * u.element_size() / payload.numel())
