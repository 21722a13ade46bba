"""Developer check (run under gpurun): CUDA stages vs the numpy oracle and vs
the reference build, printing per-stage agreement."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import mgard_b200 as mg
import mgardx_oracle as mo
import ref_x

rng = np.random.default_rng(0)
dev = torch.device("cuda:0")

def field(shape, dtype):
    g = np.meshgrid(*[np.linspace(0, 1, n) for n in shape], indexing="ij")
    u = sum(np.sin((3 + 2 * i) * x + i) for i, x in enumerate(g)) + 0.05 * rng.standard_normal(shape)
    return u.astype(dtype)

def T(a): return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

cases = [((17,), np.float32), ((100,), np.float64), ((9, 9), np.float32), ((10, 7), np.float64), ((64, 33), np.float32),
         ((5, 6, 9), np.float32), ((33, 20, 17), np.float64), ((12, 13, 14), np.float32), ((65, 65, 65), np.float32),
         ((5, 6, 7, 9), np.float32), ((4, 17, 5, 6), np.float64), ((5, 5, 6, 7, 5), np.float32), ((129, 129, 129), np.float32)]
if len(sys.argv) > 1: cases = cases[:int(sys.argv[1])]
allok = True
for shape, dt in cases:
    u = field(shape, dt)
    h = mo.Hierarchy(shape, dt)
    p = mg.Plan(shape, dt)
    tab_ok = all(np.array_equal(p.table(k, l, d), getattr(h, k)[l][d]) for l in range(h.l_target + 1) for d in range(h.D) for k in ("dist", "ratio", "am", "bm"))
    du = T(u)
    dc = p.decompose(du); torch.cuda.synchronize()
    oc = mo.decompose(h, u)
    dec_ok = np.array_equal(dc.cpu().numpy(), oc)
    dmax = np.abs(dc.cpu().numpy() - oc).max()
    rc = p.recompose(T(oc)).cpu().numpy()
    orc = mo.recompose(h, oc)
    rec_ok = np.array_equal(rc, orc)
    res = [tab_ok, dec_ok, rec_ok]
    for (eb, tol, s) in [(mo.REL, 1e-3, np.inf), (mo.ABS, 1e-2, 0.0)]:
        norm = mo.calc_norm(u, s)
        gn = p.norm(du, s)
        q, oi, ov = mo.quantize(h, oc, eb, tol, s, norm)
        sym, hist, goi, gov = p.quantize(T(oc), eb, tol, s, float(norm))
        sym_np = sym.cpu().numpy().astype(np.uint16).astype(np.int64).reshape(shape)
        q_ok = np.array_equal(sym_np, q)
        o1 = np.argsort(goi.cpu().numpy()); 
        o_ok = np.array_equal(goi.cpu().numpy()[o1].astype(np.uint64), oi) and np.array_equal(gov.cpu().numpy()[o1], ov)
        hist_ok = np.array_equal(hist.cpu().numpy().astype(np.int64), np.bincount(q.ravel(), minlength=8192))
        cb = mo.get_codebook(np.bincount(q.ravel(), minlength=8192))
        gcb, gdb = p.codebook(hist)
        gcb = gcb.cpu().numpy().view(np.uint64); gdb = gdb.cpu().numpy().view(np.uint64)
        cb_ok = np.array_equal(gcb, cb["codebook"]) and np.array_equal(gdb[:64], cb["first"]) and np.array_equal(gdb[64:128], cb["entry"]) and np.array_equal(gdb[128:], cb["keys"])
        # payload with outliers sorted
        order = np.argsort(goi.cpu().numpy())
        goi_s = goi[torch.from_numpy(order).to(dev)] if len(order) else goi
        gov_s = gov[torch.from_numpy(order).to(dev)] if len(order) else gov
        pay = p.huffman_compress(sym, hist, goi_s, gov_s).cpu().numpy()
        opay = np.frombuffer(mo.huffman_compress(q, 8192, 20480, oi, ov), dtype=np.uint8)
        pay_ok = pay.size == opay.size and np.array_equal(pay, opay)
        sym2, oi2, ov2 = p.huffman_decompress(T(opay), int(np.prod(shape)))
        dcd_ok = np.array_equal(sym2.cpu().numpy(), sym.cpu().numpy())
        dq = p.dequantize(sym, goi, gov, eb, tol, s, float(norm)).cpu().numpy()
        odq = mo.dequantize(h, q, oi, ov, eb, tol, s, norm)
        dq_ok = np.array_equal(dq, odq)
        # end to end low level
        payload, gnorm = p.compress(du, eb, tol, s)
        back = p.decompress(payload, eb, tol, s, gnorm).cpu().numpy()
        err = np.abs(back - u).max()
        res += [abs(gn - float(norm)) <= 1e-6 * abs(float(norm)), q_ok, o_ok, hist_ok, cb_ok, pay_ok, dcd_ok, dq_ok]
        print("   ", "eb", eb, "s", s, "norm", gn, float(norm), "q", q_ok, "outl", o_ok, len(oi), "hist", hist_ok, "cb", cb_ok, "payload", pay_ok, pay.size, opay.size,
              "decode", dcd_ok, "dequant", dq_ok, "e2e err", err, "CR", u.nbytes / payload.numel())
    print(shape, dt.__name__, "tables", tab_ok, "decompose", dec_ok, dmax, "recompose", rec_ok, "ALL", all(res))
    allok &= all(res)
# high level round trip + reference decode of our low-level payload
u = field((65, 65, 65), np.float32)
stream = mg.compress(u, 1e-3, np.inf, mg.error_bound_type.REL)
back = mg.decompress(stream)
print("high-level host roundtrip err", np.abs(back - u).max(), "bound", 1e-3 * np.abs(u).max(), "CR", u.nbytes / stream.size)
ostream = mo.compress(u, mo.REL, 1e-3, np.inf)
print("stream == oracle stream:", stream.tobytes() == ostream, stream.size, len(ostream))
p = mg.Plan(u.shape, u.dtype)
payload, gnorm = p.compress(T(u), mo.REL, 1e-3, np.inf)
rb = ref_x.decompress(payload.cpu().numpy(), u.shape, u.dtype, ref_x.REL, 1e-3, np.inf, gnorm)
print("reference decodes our payload: err", np.abs(rb - u).max())
rr = ref_x.compress(u, ref_x.REL, 1e-3, np.inf)
ours = p.decompress(T(rr["payload"]), mo.REL, 1e-3, np.inf, rr["norm"]).cpu().numpy()
print("we decode reference payload: err", np.abs(ours - u).max(), "identical to reference's own decode:",
      np.array_equal(ours, ref_x.decompress(rr["payload"], u.shape, u.dtype, ref_x.REL, 1e-3, np.inf, rr["norm"])))
du = T(u)
ds = mg.compress(du, 1e-3, np.inf, mg.error_bound_type.REL)
print("device stream equal host stream:", np.array_equal(ds.cpu().numpy(), stream))
db = mg.decompress(ds)
print("device roundtrip err", (db - du).abs().max().item())
print("ALLOK", allok)
