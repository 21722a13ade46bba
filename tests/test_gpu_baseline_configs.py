"""BASELINE.json configurations at FULL size, CUDA path against the reference build.

The reference (oracle/_ref/libmgardx_ref.so = unmodified MGARD-X, SERIAL adapter) was run
on these inputs by tests/golden/make_baseline_digests.py; tests/golden/baseline_digests.json
holds SHA-256 digests of its decomposed coefficients, quantized symbols, Huffman block
(field by field, outlier list as a set) and reconstruction.  Here the same inputs are
regenerated (tests/baseline_fields.py; the input digest must reproduce), pushed through
the C ABI on the GPU and compared digest by digest: bit-exact coefficients (stricter than
north_star's 1e-5 / 1e-12), identical symbols, byte-identical Huffman block, identical
reconstruction, identical compression ratio; and the requested bound is verified on the
reconstruction.  Small configurations additionally run the reference live.

Relative bounds with a finite s: the reference sums the squares in T in whatever order
its backend reduces (sequentially in the SERIAL build, NormCalculator.hpp:44-68 ->
DeviceAdapterSerial.h:1370-1385); this path reduces in double.  With the reference's norm
handed in the results are byte-identical; with this path's own norm the differing quanta
are COUNTED and reported (north_star: ties counted), and the ratio must agree within 1 %.

  C4 (1.6 GB) and the C5 slab (4.3 GB) take a few minutes of host time for input
  generation and hashing: they run when MGB_SLOW=1 (results of this round's run are in
  profiles/r2_parity_configs.json)."""
import json
import os

import numpy as np
import pytest

import baseline_fields as bf
import mgardx_oracle as mo
import ref_x

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
DIGESTS = json.load(open(os.path.join(HERE, "golden", "baseline_digests.json")))
SLOW = os.environ.get("MGB_SLOW", "0") == "1"
REPORT = os.path.join(os.path.dirname(HERE), "gpurun_out", "parity_configs.jsonl")

GEN = {
    "C1": lambda: (bf.c1(), None),
    "C2": lambda: (bf.c2(), None),
    "C3abs": bf.c3,
    "C3rel": bf.c3,
    "C4crop": lambda: (bf.c4(2051), None),
    "C4": lambda: (bf.c4(), None),
    "C5slab": lambda: (bf.c5_slab(0), None),
}


@pytest.fixture(scope="module")
def env():
    import torch
    import mgard_b200 as mg
    assert torch.cuda.is_available()
    return torch, mg, torch.device("cuda:0")


def sha_t(t):
    return bf.sha(t.cpu().numpy())


def report(rec):
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        with open(REPORT, "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass


def reference_record(name, u, coords, want):
    """The committed digests when the input reproduces bit for bit on this machine,
    otherwise the reference run live (small configurations only)."""
    if bf.sha(u) == want["input"]:
        return want, "committed digests"
    if u.nbytes > (1 << 30) or not ref_x.available():
        pytest.skip("input does not reproduce on this host (libm / SIMD differences) and the "
                    "configuration is too large to run the reference inside the suite")
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_baseline_digests as mk
    mk.CONFIGS[name] = (lambda: (u, coords),) + mk.CONFIGS[name][1:]
    return mk.run(name), "reference run live"


def run_config(env, name):
    torch, mg, d = env
    want = DIGESTS[name]
    u, coords = GEN[name]()
    assert list(u.shape) == want["shape"] and u.dtype.name == want["dtype"]
    want, source = reference_record(name, u, coords, want)
    eb, tol = want["ebtype"], want["tol"]
    s = bf.INF if want["s"] == "inf" else float(want["s"])
    ref_norm = float(want["norm"])
    p = mg.Plan(u.shape, u.dtype, coords=coords)
    assert p.l_target == want["l_target"]
    du = torch.from_numpy(u).to(d)
    rec = {"config": name, "shape": want["shape"], "dtype": want["dtype"], "reference": source}

    # --- stage by stage --------------------------------------------------------------
    coef = p.decompose(du)
    assert sha_t(coef) == want["decomposed"], "decomposed coefficients differ from the reference"
    rec["coefficients"] = "bit-exact"
    if eb == mo.REL:
        own_norm = p.norm(du, s)
        if np.isinf(s):
            assert own_norm == ref_norm
        else:
            # this path reduces in double: within rounding of the exact value.  The SERIAL
            # reference accumulates sequentially in T (DeviceAdapterSerial.h:1370-1385): in
            # fp32 over 10^6 squares that is good to ~4 digits only, and reported here
            exact = float(np.sqrt((u.astype(np.float64) ** 2).sum() / u.size))
            assert abs(own_norm - exact) <= (2e-7 if u.dtype == np.float32 else 1e-14) * exact
            assert abs(own_norm - ref_norm) <= (2e-3 if u.dtype == np.float32 else 1e-10) * ref_norm
            rec["reference_norm_relative_deviation_from_exact"] = abs(ref_norm - exact) / exact
        rec["norm"] = own_norm
        rec["reference_norm"] = ref_norm
    sym, hist, oi, ov = p.quantize(coef, eb, tol, s, ref_norm if eb == mo.REL else 1.0)
    assert sha_t(sym) == want["symbols"], "quantized symbols differ from the reference"
    assert int(oi.numel()) == want["outlier_count"]
    rec["symbols"] = "identical"
    rec["outliers"] = int(oi.numel())
    pay = p.huffman_compress(sym, hist, oi, ov)
    got = bf.payload_digests(mo.huffman_parse(pay.cpu().numpy().tobytes()))
    assert got == want["payload"], {k: (got[k], want["payload"][k]) for k in got if got[k] != want["payload"][k]}
    assert pay.numel() == want["payload_bytes"]
    rec["huffman_block"] = "byte-identical (outlier list as a set)"
    rec["ratio"] = u.nbytes / pay.numel()
    rec["reference_ratio"] = want["ratio"]
    del coef, hist

    # --- Compressor::Compress / Decompress (the fused path the bench times) -----------
    payload, norm = p.compress(du, eb, tol, s)
    exact_norm = eb != mo.REL or np.isinf(s)
    if exact_norm:
        got = bf.payload_digests(mo.huffman_parse(payload.cpu().numpy().tobytes()))
        assert got == want["payload"], "Compressor::Compress block differs from the reference"
        back = p.decompress(payload, eb, tol, s, norm)
        assert sha_t(back) == want["decompressed"], "reconstruction differs from the reference"
        rec["reconstruction"] = "bit-exact"
    else:
        # own norm (double reduction): count the quanta that differ from the reference's
        sym2, _, oi2, _ = p.quantize(p.decompose(du), eb, tol, s, norm)
        differ = int((sym2 != sym).sum())
        maxstep = int((sym2.to(torch.int32) - sym.to(torch.int32)).abs().max()) if differ else 0
        rec["differing_quanta_with_own_norm"] = differ
        rec["differing_quanta_fraction"] = differ / u.size
        rec["differing_quanta_max_step"] = maxstep
        # a relative change r of the norm moves a quantum q by at most |q| r (|q| <= dict/2)
        rel = abs(norm - ref_norm) / ref_norm
        assert maxstep <= 1 + int(4096 * rel + 1)
        assert abs(payload.numel() - want["payload_bytes"]) <= 0.01 * want["payload_bytes"]
        rec["ratio_with_own_norm"] = u.nbytes / payload.numel()
        # the reference's own block decodes to the reference's reconstruction
        back = p.decompress(pay, eb, tol, s, ref_norm)
        assert sha_t(back) == want["decompressed"], "reconstruction differs from the reference"
        rec["reconstruction"] = "bit-exact (reference norm)"
        back = p.decompress(payload, eb, tol, s, norm)
    # --- the requested bound, on the reconstruction -------------------------------------
    diff = back.double() - du.double()
    if np.isinf(s):
        err = float(diff.abs().max())
        bound = tol * (float(du.abs().max()) if eb == mo.REL else 1.0)
    else:  # X convention: sqrt(sum e^2 / N) (ErrorCalculator.h:36-53)
        err = float(torch.sqrt((diff * diff).sum() / u.size))
        bound = tol * (float(torch.sqrt((du.double() ** 2).sum() / u.size)) if eb == mo.REL else 1.0)
    assert err <= bound
    rec["error"], rec["bound"] = err, bound
    report(rec)
    return rec


@pytest.mark.parametrize("name", ["C1", "C2", "C3abs", "C3rel", "C4crop"])
def test_baseline_config_matches_reference(env, name):
    run_config(env, name)


@pytest.mark.skipif(not SLOW, reason="MGB_SLOW=1 runs the 1.6 GB / 4.3 GB configurations")
@pytest.mark.parametrize("name", ["C4", "C5slab"])
def test_large_baseline_config_matches_reference(env, name):
    run_config(env, name)


@pytest.mark.skipif(not ref_x.available(), reason="oracle/_ref not in the snapshot")
def test_c3_x_convention_live_reference(env):
    """C3 (1000^2 fp32, non-uniform coordinates, s = 0) against the reference run in this
    process: every stage as arrays, not digests, and both directions of cross-decoding."""
    torch, mg, d = env
    u, cs = bf.c3()
    p = mg.Plan(u.shape, np.float32, coords=cs)
    du = torch.from_numpy(u).to(d)
    for eb, tol in ((mo.ABS, 1e-2), (mo.REL, 1e-2)):
        r = ref_x.compress(u, eb, tol, 0.0, cs)
        assert np.array_equal(p.decompose(du).cpu().numpy(), r["decomposed"])
        nrm = r["norm"] if eb == mo.REL else 1.0
        sym, hist, oi, ov = p.quantize(p.decompose(du), eb, tol, 0.0, nrm)
        assert np.array_equal(sym.cpu().numpy().astype(np.uint16).astype(np.int64).reshape(u.shape),
                              r["quantized"])
        ours = p.decompress(torch.from_numpy(r["payload"]).to(d), eb, tol, 0.0, nrm).cpu().numpy()
        theirs = ref_x.decompress(r["payload"], u.shape, u.dtype, eb, tol, 0.0, nrm, cs)
        assert np.array_equal(ours, theirs)
        payload, n2 = p.compress(du, eb, tol, 0.0)
        theirs2 = ref_x.decompress(payload.cpu().numpy(), u.shape, u.dtype, eb, tol, 0.0, n2, cs)
        assert np.array_equal(p.decompress(payload, eb, tol, 0.0, n2).cpu().numpy(), theirs2)
        assert abs(payload.numel() - r["payload"].size) <= 0.01 * r["payload"].size
