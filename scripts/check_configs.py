"""BASELINE.json configs other than the bench workload (C2): run once on the GPU,
verify the requested bound on reconstruction, report ratio and device time.
  C1 129^3 fp64 ABS 1e-4 s=inf         C3 1000^2 fp32 non-uniform s=0 ABS/REL 1e-2
  C4 8x16395x39x39 fp64 REL 1e-3 s=0   C5 one 257x2049x2049 fp32 slab, ABS (local tol)"""
import os, sys, math, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mgard_b200 as mg
import bench
dev = torch.device("cuda:0")
INF = float("inf")
out = []

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1) / reps

def run(name, u, mode, tol, s, coords=None):
    shape = tuple(u.shape)
    npdt = np.float32 if u.dtype == torch.float32 else np.float64
    plan = mg.Plan(shape, npdt, coords=coords)
    (payload, norm), tc = timed(lambda: plan.compress(u, mode, tol, s))
    back, td = timed(lambda: plan.decompress(payload, mode, tol, s, norm))
    n = u.numel()
    diff = (back.double() - u.double())
    if math.isinf(s):
        err = float(diff.abs().max()); ref = float(u.abs().max()) if mode == mg.error_bound_type.REL else 1.0
    else:  # X-convention L2: sqrt(sum(e^2)/N) (ErrorCalculator.h:36-53)
        err = float(torch.sqrt((diff * diff).sum() / n)); ref = float(torch.sqrt((u.double() ** 2).sum() / n)) if mode == mg.error_bound_type.REL else 1.0
    bound = tol * ref
    nbytes = n * u.element_size()
    rec = {"config": name, "shape": shape, "dtype": str(u.dtype), "levels": plan.l_target, "ratio": nbytes / payload.numel(),
           "err": err, "bound": bound, "bound_ok": err <= bound, "compress_ms": tc, "decompress_ms": td,
           "compress_gbs": nbytes / tc / 1e6, "decompress_gbs": nbytes / td / 1e6}
    print(json.dumps(rec)); out.append(rec)
    del plan, payload, back
    torch.cuda.empty_cache()

which = sys.argv[1:] or ["C1", "C3", "C4", "C5"]
if "C1" in which:
    n = 129
    x = [torch.linspace(0, 1, n, dtype=torch.float64, device=dev) for _ in range(3)]
    X0, X1, X2 = torch.meshgrid(*x, indexing="ij")
    u = torch.sin(2 * math.pi * X0) * torch.cos(3 * math.pi * X1) + 0.5 * torch.sin(5 * math.pi * X2) + 0.25 * X0 * X1
    run("C1 129^3 fp64 ABS 1e-4 s=inf (X convention)", u.contiguous(), mg.error_bound_type.ABS, 1e-4, INF)
if "C3" in which:
    n = 1000
    cs = []
    for k in (7, 11):
        i = np.arange(n - 1)
        hsp = 1 + 0.5 * np.sin(2 * np.pi * k * i / 999)
        xx = np.concatenate([[0.0], np.cumsum(hsp)]); xx /= xx[-1]
        cs.append(xx.astype(np.float32))
    X0, X1 = np.meshgrid(cs[0].astype(np.float64), cs[1].astype(np.float64), indexing="ij")
    u = (np.exp(-8 * ((X0 - .5) ** 2 + (X1 - .4) ** 2)) + 0.1 * np.sin(30 * X0)).astype(np.float32)
    ut = torch.from_numpy(u).to(dev)
    run("C3 1000^2 fp32 non-uniform s=0 ABS 1e-2", ut, mg.error_bound_type.ABS, 1e-2, 0.0, coords=cs)
    run("C3 1000^2 fp32 non-uniform s=0 REL 1e-2", ut, mg.error_bound_type.REL, 1e-2, 0.0, coords=cs)
if "C4" in which:
    shape = (8, 16395, 39, 39)
    i0 = torch.arange(shape[0], device=dev, dtype=torch.float64).view(-1, 1, 1, 1)
    x1 = torch.linspace(0, 1, shape[1], device=dev, dtype=torch.float64).view(1, -1, 1, 1)
    x2 = torch.linspace(0, 1, shape[2], device=dev, dtype=torch.float64).view(1, 1, -1, 1)
    x3 = torch.linspace(0, 1, shape[3], device=dev, dtype=torch.float64).view(1, 1, 1, -1)
    g = sum((1.0 / k) * torch.sin(2 * math.pi * (2 * k + 1) * x1) for k in range(1, 6))
    u = (1 + 0.1 * i0) * g * torch.exp(-((x2 - .5) ** 2 + (x3 - .5) ** 2) / 0.08)
    u = u.contiguous()
    run("C4 8x16395x39x39 fp64 REL 1e-3 s=0", u, mg.error_bound_type.REL, 1e-3, 0.0)
    del u
if "C5" in which:
    shape = (257, 2049, 2049)
    u = bench.field_torch(shape, dev)
    # one MaxDim slab of the 2049^3 domain, compressed in ABS mode with the local tolerance
    run("C5 slab 257x2049x2049 fp32 ABS 1.3e-3 s=inf", u, mg.error_bound_type.ABS, 1.3e-3, INF)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "check_configs.json"), "w"), indent=1)
