"""BASELINE config C5 on ONE GPU: 2049^3 fp32 (34.4 GB), relative L-inf 1e-3, MaxDim
sub-domains of 257 planes (7 x 257 + 250), through the high-level API on device buffers.
Verifies the bound on reconstruction and reports ratio and throughput."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mgard_b200 as mg
import bench
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2049
shape = (n, n, n)
u = bench.field_torch(shape, dev)
cfg = mg.Config()
cfg.domain_decomposition_dim = 0
cfg.domain_decomposition_size = 257
nbytes = u.numel() * 4
out = torch.empty(nbytes // 2 + (64 << 20), dtype=torch.uint8, device=dev)
back = torch.empty_like(u)
res = {}
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    stream = mg.compress(u, 1e-3, float("inf"), mg.error_bound_type.REL, config=cfg, out=out)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    mg.decompress(stream, out=back)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    res = {"shape": shape, "compress_s": t1 - t0, "decompress_s": t2 - t1,
           "compress_gbs": nbytes / (t1 - t0) / 1e9, "decompress_gbs": nbytes / (t2 - t1) / 1e9,
           "ratio": nbytes / stream.numel()}
err = 0.0
amax = 0.0
for a in range(0, n, 128):
    err = max(err, float((back[a:a + 128] - u[a:a + 128]).abs().max()))
    amax = max(amax, float(u[a:a + 128].abs().max()))
res.update({"max_abs_error": err, "bound": 1e-3 * amax, "bound_ok": err <= 1e-3 * amax,
            "header": {k: v for k, v in mg.peek_header(stream[:4096].cpu().numpy()).items() if k != "coords"}})
print(json.dumps(res, default=str))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "check_c5.json"), "w"), indent=1, default=str)
