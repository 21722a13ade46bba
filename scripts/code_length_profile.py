"""Code-length distribution of one sub-domain's Huffman code (developer tool, run under
gpurun): fraction of the symbols whose codeword is longer than L bits, overall and for
the worst chunk.  Usage: code_length_profile.py [n0 n1 n2]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench, mgard_b200 as mg
dev = torch.device("cuda:0")
shape = tuple(int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (257, 2049, 2049)
plane0 = int(sys.argv[4]) if len(sys.argv) > 4 else 0
u = bench.field_torch(shape, dev, plane0=plane0, full_n0=2049 if shape[1] == 2049 else None)
p = mg.Plan(shape, np.float32)
norm = float(u.abs().max())
coef = p.decompose(u)
sym, hist, oi, ov = p.quantize(coef, mg.error_bound_type.REL, 1e-3, float("inf"), norm)
cb, db = p.codebook(hist)
lens = (cb.cpu().numpy().view(np.uint64) >> np.uint64(56)).astype(np.int64)
h = hist.cpu().numpy().astype(np.int64)
tot = h.sum()
print("symbols", tot, "used", int((h > 0).sum()), "avg bits", float((h * lens).sum()) / tot, "max len", int(lens[h > 0].max()))
first = np.full(64, -1, dtype=np.int64)
cw = cb.cpu().numpy().view(np.uint64)
code = (cw & np.uint64((1 << 56) - 1)).astype(np.uint64)
used = h > 0
# Kraft mass of the codewords longer than L bits, in units of 2^-L: prefixes of L bits that hold longer codes
for L in (12, 16, 20, 24):
    m = sum(2.0 ** (L - int(l)) for l in lens[used & (lens > L)])
    print(f"  {L}-bit prefixes holding longer codewords: {m:.1f}")
for L in range(8, 31):
    print(f"  longer than {L:2d} bits: {h[lens > L].sum() / tot:.5f}")
# per chunk: share of long codewords (> 12 bits) in the worst chunks
l_dev = torch.from_numpy(lens.astype(np.int16)).to(dev)
sl = l_dev[sym.to(torch.int64) & 0xffff]
chunk = 20480
nfull = sl.numel() // chunk
per = (sl[: nfull * chunk].view(nfull, chunk) > 12).float().mean(dim=1)
bits = sl[: nfull * chunk].view(nfull, chunk).to(torch.float32).mean(dim=1)
q = torch.tensor([0.5, 0.9, 0.99, 0.999, 1.0], device=dev)
print("share of > 12-bit codewords per chunk, quantiles 50/90/99/99.9/100:", torch.quantile(per, q).cpu().numpy())
print("bits per symbol per chunk, quantiles:", torch.quantile(bits, q).cpu().numpy())
