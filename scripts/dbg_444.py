import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import mgard_b200 as mg
import mgardx_oracle as mo
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
for shape, dt in [((4, 4, 4), np.float64), ((5, 6, 9), np.float32), ((6, 5, 5), np.float32), ((5, 6, 5), np.float32), ((5, 5, 6), np.float32), ((8, 8, 8), np.float32)]:
    u = rng.standard_normal(shape).astype(dt)
    h = mo.Hierarchy(shape, dt)
    p = mg.Plan(shape, dt)
    oc = mo.decompose(h, u)
    g = p.decompose(torch.from_numpy(u).to(dev)).cpu().numpy()
    bad = np.argwhere(g != oc)
    print(shape, "levels", h.l_target, [h.level_shape[l] for l in range(h.l_target + 1)], "mismatch", len(bad))
    for b in bad[:12]:
        print("   ", tuple(b), g[tuple(b)], oc[tuple(b)])
