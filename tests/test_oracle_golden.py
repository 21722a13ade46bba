"""CPU: the numpy oracle against the committed golden fixtures (generated from
the unmodified reference by tests/golden/make_golden.py) and against the
reference's own known-answer vectors for this path."""
import glob
import os

import numpy as np
import pytest

import mgardx_oracle as mo

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "d*_*.npz")))


def load(path):
    z = np.load(path)
    shape = z["u"].shape
    coords = [z[f"coords{d}"] for d in range(len(shape))] if "coords0" in z else None
    return z, shape, coords


def sorted_outliers(p):
    o = np.argsort(p["oidx"])
    return p["oidx"][o], p["oval"][o]


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_fixture(path):
    z, shape, coords = load(path)
    u = z["u"]
    h = mo.Hierarchy(shape, u.dtype, coords)
    assert h.l_target == int(z["l_target"])
    for d in range(len(shape)):
        for k in ("dist", "ratio", "am", "bm"):
            assert np.array_equal(getattr(h, k)[h.l_target][d], z[f"tab_{k}_L_{d}"])
            assert np.array_equal(getattr(h, k)[0][d], z[f"tab_{k}_0_{d}"])
    # bit-exact decomposition / recomposition
    dec = mo.decompose(h, u)
    assert np.array_equal(dec, z["decomposed"])
    assert np.array_equal(mo.recompose(h, z["decomposed"]), z["recomposed"])
    # bit-exact quantized indices (given the reference's norm)
    eb, tol, s, norm = int(z["ebtype"]), float(z["tol"]), float(z["s"]), u.dtype.type(z["norm"])
    q, oi, ov = mo.quantize(h, dec, eb, tol, s, norm)
    assert np.array_equal(q, z["quantized"])
    # Huffman block: identical bytes except the (unordered) outlier list and the
    # reference's out-of-bounds read in GenerateCL (see oracle docstring)
    ref = mo.huffman_parse(z["payload"].tobytes())
    ok = False
    for oob in (0, 0xFFFFFFFF):
        mine = mo.huffman_parse(mo.huffman_compress(q, 8192, 20480, oi, ov, oob_value=oob))
        same = all(np.array_equal(mine[k], ref[k]) for k in
                   ("bits", "word_offset", "first", "entry", "keys", "ddata"))
        ok = ok or same
    assert ok
    a, b = sorted_outliers(ref)
    assert np.array_equal(a, oi) and np.array_equal(b, ov)
    # decoding the reference payload reproduces the reference's reconstruction
    back = mo.decompress_lowlevel(h, z["payload"].tobytes(), eb, tol, s, norm)
    assert np.array_equal(back, z["decompressed"])


def test_codebook_fixtures():
    z = np.load(os.path.join(HERE, "golden", "codebooks.npz"))
    i = 0
    matched0 = 0
    while f"freq{i}" in z:
        res = []
        for oob in (0, 0xFFFFFFFF):
            cb = mo.get_codebook(z[f"freq{i}"], oob)
            res.append(all(np.array_equal(cb[k], z[f"{k}{i}"]) for k in ("codebook", "first", "entry", "keys")))
        assert any(res), i
        matched0 += res[0]
        i += 1
    assert i >= 4 and matched0 >= i - 1


def test_header_fixtures():
    """encode_header == proto3 canonical bytes produced by Python protobuf from the
    reference's src/mgard.proto with the fields MetadataBase::Serialize sets."""
    z = np.load(os.path.join(HERE, "golden", "headers.npz"))
    i = 0
    while f"hdr{i}" in z:
        shape = tuple(int(x) for x in z[f"shape{i}"])
        dt, eb, tol, s, norm, dec, dd, ds = z[f"meta{i}"]
        coords = [z[f"coords{i}_{d}"] for d in range(len(shape))] if f"coords{i}_0" in z else None
        b = mo.encode_header(shape, np.float32 if dt == 0 else np.float64, int(eb), tol, s, norm,
                             coords, bool(dec), int(dd), int(ds))
        assert b == z[f"hdr{i}"].tobytes()
        i += 1
    assert i == 5


# ---- the reference's own known-answer vectors (MGARD-CPU tests) ----------------
# tests/src/test_decompose.cpp:277-301 (1-D, 33 nodes, float) and :303-337 (2-D 5x5,
# double): multilevel coefficients in nodal order.  On dyadic uniform grids the
# MGARD-X transform is the same operator, stored coarse-first; compare after
# undoing the layout.

U33 = [10, 3, -8, -6, 3, 0, -5, 0, 0, -2, -8, -5, -10, -7, 8, -2, 3, -1, 0, 9, -4, -6,
       -8, -5, -10, 1, 3, 7, -8, 1, 10, -2, 8]
E33 = {
    1: [11.0, 2.0, -7.0],
    2: [4.4375, 2.0, -14.5, -3.5, -6.687500000000001],
    3: [0.4374999999999991, 2.0, -15.678571428571429, -3.5, -4.625000000000002, 1.0,
        -5.321428571428571, 2.5, -2.4375000000000004],
    4: [-0.95703125, 2.0, -15.652199926362297, -3.5, -4.122767857142856, 1.0,
        -4.978599042709867, 2.5, -4.765625000000001, 2.0, -1.173186671575852, 4.0,
        -10.689732142857139, -6.0, 9.303985640648008, -7.5, -3.1054687499999987],
    5: [-2.640624999999999, 2.0, -15.652212055333662, -3.5, -4.04917617820324, 1.0,
        -4.978756719337627, 2.5, -6.024553571428571, 2.0, -1.1753820153931231, 4.0,
        -9.73304031664212, -6.0, 9.273408503833881, -7.5, -3.0234374999999996, -2.5,
        3.878101069067445, 11.0, -1.7446382547864507, 0.0, -3.6049935368896406, 4.0,
        -8.00669642857143, 4.5, 13.02698941447754, 9.5, 2.7768547496318097, 0.0,
        8.232845339575196, -11.0, 0.14062500000000222],
}
U55 = [7, 4, 5, -10, -6, 6, -8, -5, 6, -2, 2, -2, 9, 2, -10, 3, 8, -8, -3, 7, -8, -9, -6, -1, -4]
E55 = [3.8007812499999973, -2.0, -2.9062499999998854, -9.5, -2.910156250000001, 1.5, -13.75,
       -12.0, 6.5, 6.0, -1.593749999999881, -7.5, 2.8750000000004396, 2.5, -1.0312499999998854,
       6.0, 8.75, -9.5, -0.25, 14.0, -2.5039062500000013, -2.0, -10.218749999999885, 4.0,
       -0.6992187500000024]


def to_nodal(h, v):
    """Undo the coarse-first layout level by level."""
    out = np.array(v, copy=True)
    for l in range(1, h.l_target + 1):
        n, nc = h.level_shape[l], h.level_shape[l - 1]
        box = tuple(slice(0, k) for k in n)
        X = mo._from_octants(out[box], n, nc)
        assert list(X.shape) == list(n)  # dyadic: no ghost slots
        out[box] = X
    return out


@pytest.mark.parametrize("level", [1, 2, 3, 4, 5])
def test_reference_kat_1d_33(level):
    # the reference test feeds the first 2^L + 1 entries (test_decompose.cpp:37-60)
    u = np.array(U33, dtype=np.float32)[: 2 ** level + 1]
    h = mo.Hierarchy(u.shape, np.float32)
    got = to_nodal(h, mo.decompose(h, u))
    np.testing.assert_allclose(got, np.array(E33[level]), rtol=1e-5, atol=1e-5)


def test_reference_kat_2d_5x5():
    u = np.array(U55, dtype=np.float64).reshape(5, 5)
    h = mo.Hierarchy(u.shape, np.float64)
    got = to_nodal(h, mo.decompose(h, u))
    np.testing.assert_allclose(got.ravel(), np.array(E55), rtol=1e-12, atol=1e-10)


def test_piecewise_linear_has_zero_coefficients():
    """tests/src/test_decompose.cpp:459-492 property on non-dyadic, non-uniform grids."""
    rng = np.random.default_rng(5)
    for shape in [(129,), (21, 20), (14, 10, 17)]:
        coords = [np.cumsum(rng.uniform(1, 2, n)) for n in shape]
        coords = [(c - c[0]) / (c[-1] - c[0]) for c in coords]
        h = mo.Hierarchy(shape, np.float64, coords)
        g = np.meshgrid(*coords, indexing="ij")
        u = sum((i + 1.5) * x for i, x in enumerate(g)) + 0.25
        v = mo.decompose(h, u)
        lv = mo.node_levels(h)
        assert np.abs(v[lv > 0]).max() < 1e-12


def test_round_trip_and_error_bound():
    rng = np.random.default_rng(11)
    for shape, dt in [((65,), np.float32), ((17, 19), np.float64), ((10, 5, 12), np.float32)]:
        g = np.meshgrid(*[np.linspace(0, 1, n) for n in shape], indexing="ij")
        u = (sum(np.cos(4 * x) for x in g) + 0.01 * rng.standard_normal(shape)).astype(dt)
        h = mo.Hierarchy(shape, dt)
        assert np.abs(mo.recompose(h, mo.decompose(h, u)) - u).max() < (1e-5 if dt == np.float32 else 1e-13)
        for tol in (1e-1, 1e-3):
            r = mo.compress_lowlevel(h, u, mo.REL, tol, np.inf)
            back = mo.decompress_lowlevel(h, r["payload"], mo.REL, tol, np.inf, r["norm"])
            assert np.abs(back - u).max() <= tol * np.abs(u).max()


def test_reorder_and_single_dimension_fixtures():
    """tests/golden/x_modes.npz (make_golden_modes.py): Config::reorder = 1 and
    decomposition_type::SingleDim outputs of the unmodified reference build."""
    z = np.load(os.path.join(HERE, "golden", "x_modes.npz"))
    for i in range(int(z["count"])):
        shape = tuple(int(x) for x in z[f"shape{i}"])
        dt = np.float64 if int(z[f"dtype{i}"]) == 1 else np.float32
        coords = None
        if int(z[f"explicit{i}"]):
            flat, coords, off = z[f"coords{i}"], [], 0
            for m in shape:
                coords.append(flat[off:off + m].astype(dt))
                off += m
        tol, s = float(z[f"tol{i}"]), float(z[f"s{i}"])
        reorder, sd = int(z[f"reorder{i}"]), bool(int(z[f"single{i}"]))
        h = mo.Hierarchy(shape, dt, coords)
        u = z[f"u{i}"]
        dec = mo.decompose_single(h, u) if sd else mo.decompose(h, u)
        assert np.array_equal(dec, z[f"decomposed{i}"])
        m = mo.compress_lowlevel(h, u, mo.REL, tol, s, dt(z[f"norm{i}"]), reorder=reorder, single_dim=sd)
        assert np.array_equal(np.asarray(m["quantized"]).ravel(), z[f"quantized{i}"])
        back = mo.decompress_lowlevel(h, m["payload"], mo.REL, tol, s, dt(z[f"norm{i}"]), reorder=reorder, single_dim=sd)
        assert np.array_equal(back, z[f"decompressed{i}"])
