import sys, time, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch, numpy as np
import mgard_b200 as mg, bench
dev = torch.device("cuda:0")
n = 513**3
h = torch.empty(n, dtype=torch.float32, pin_memory=True); h.normal_()
d = torch.empty(n, dtype=torch.float32, device=dev)
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(name, "540 MB pinned:", round(dt * 1e3, 2), "ms", round(n * 4 / dt / 1e9, 1), "GB/s")
u = bench.field_torch((513, 513, 513), dev)
hin = torch.empty((513, 513, 513), dtype=torch.float32, pin_memory=True); hin.copy_(u)
hout = torch.empty(n * 4 + (1 << 21), dtype=torch.uint8, pin_memory=True)
hback = torch.empty((513, 513, 513), dtype=torch.float32, pin_memory=True)
a, o, b = hin.numpy(), hout.numpy(), hback.numpy()
for _ in range(2):
    s = mg.compress(a, 1e-3, float("inf"), mg.error_bound_type.REL, out=o); mg.decompress(s, out=b)
torch.cuda.synchronize()
t0 = time.perf_counter(); s = mg.compress(a, 1e-3, float("inf"), mg.error_bound_type.REL, out=o); torch.cuda.synchronize(); t1 = time.perf_counter()
mg.decompress(s, out=b); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host API compress", round((t1 - t0) * 1e3, 2), "ms; decompress", round((t2 - t1) * 1e3, 2), "ms; stream", s.size / 1e6, "MB")
