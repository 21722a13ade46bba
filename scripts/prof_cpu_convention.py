"""One decomposition of a 513^3 fp32 field through the MGARD-CPU convention (for ncu launch lists)."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mgard_b200.cpu as mc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 513
x = torch.linspace(0, 1, n, device="cuda")
u = (torch.sin(6 * x)[:, None, None] * torch.cos(4 * x)[None, :, None] + x[None, None, :] ** 2).float().contiguous()
H = mc.TensorMeshHierarchy((n, n, n), None, np.float32)
c = H.decompose(u)
torch.cuda.synchronize()
print("ok", float(c.abs().max()))
