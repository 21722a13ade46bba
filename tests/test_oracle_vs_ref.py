"""CPU: the numpy oracle against the reference itself (oracle/_ref, the
unmodified MGARD-X SERIAL build) on fresh seeded inputs.  Skipped where the
reference build is absent."""
import numpy as np
import pytest

import mgardx_oracle as mo
import ref_x

pytestmark = pytest.mark.skipif(not ref_x.available(), reason="oracle/_ref not built")


def field(shape, dtype, seed):
    rng = np.random.default_rng(seed)
    g = np.meshgrid(*[np.linspace(0, 1, n) for n in shape], indexing="ij")
    u = sum(np.sin((3 + 2 * i) * x + i) for i, x in enumerate(g)) + 0.05 * rng.standard_normal(shape)
    return u.astype(dtype)


SHAPES = [(5,), (6,), (100,), (9, 9), (10, 7), (64, 33), (5, 6, 9), (33, 20, 17), (4, 4, 4),
          (3, 3, 3), (5, 6, 7, 9), (5, 5, 6, 7, 5)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_tables_and_transform_bit_exact(shape, dtype):
    u = field(shape, dtype, len(shape))
    h = mo.Hierarchy(shape, dtype)
    tb = ref_x.tables(shape, dtype)
    for l in range(h.l_target + 1):
        for d in range(h.D):
            for k in ("dist", "ratio", "am", "bm"):
                assert np.array_equal(tb[l][d][k], getattr(h, k)[l][d])
    a = ref_x.decompose(u)
    assert np.array_equal(a, mo.decompose(h, u))
    assert np.array_equal(ref_x.recompose(a), mo.recompose(h, a))


@pytest.mark.parametrize("case", [
    ((100,), np.float32, mo.REL, 1e-3, np.inf), ((33, 20), np.float32, mo.REL, 1e-2, 0.0),
    ((64, 65), np.float64, mo.REL, 1e-3, np.inf), ((17, 19, 21), np.float32, mo.REL, 1e-3, np.inf),
    ((33, 33, 33), np.float32, mo.ABS, 1e-2, 0.5), ((40, 24, 30), np.float64, mo.REL, 1e-4, -1.0),
    ((6, 9, 5, 7), np.float32, mo.ABS, 1e-3, np.inf)])
def test_quantized_and_payload(case):
    shape, dt, eb, tol, s = case
    u = field(shape, dt, 3)
    r = ref_x.compress(u, eb, tol, s)
    h = mo.Hierarchy(shape, dt)
    m = mo.compress_lowlevel(h, u, eb, tol, s, dt(r["norm"]) if eb == mo.REL else None)
    assert np.array_equal(m["quantized"], r["quantized"])
    ref = mo.huffman_parse(r["payload"].tobytes())
    ok = False
    for oob in (0, 0xFFFFFFFF):
        mine = mo.huffman_parse(mo.huffman_compress(m["quantized"], 8192, 20480, m["oidx"], m["oval"], oob))
        ok = ok or all(np.array_equal(mine[k], ref[k]) for k in
                       ("bits", "word_offset", "first", "entry", "keys", "ddata"))
    assert ok
    o = np.argsort(ref["oidx"])
    assert np.array_equal(ref["oidx"][o], m["oidx"]) and np.array_equal(ref["oval"][o], m["oval"])
    if u.size <= 30000:
        back = mo.decompress_lowlevel(h, r["payload"].tobytes(), eb, tol, s, dt(r["norm"]))
        assert np.array_equal(back, ref_x.decompress(r["payload"], shape, dt, eb, tol, s, r["norm"]))


def test_nonuniform_c3_shape():
    """BASELINE config 3 geometry: 2-D non-dyadic, non-uniform, s = 0."""
    n = 200
    coords = []
    for k in (7, 11):
        hh = 1 + 0.5 * np.sin(2 * np.pi * k * np.arange(n - 1) / (n - 1))
        x = np.concatenate([[0], np.cumsum(hh)])
        coords.append((x / x[-1]).astype(np.float32))
    g = np.meshgrid(*coords, indexing="ij")
    u = (np.exp(-8 * ((g[0] - .5) ** 2 + (g[1] - .4) ** 2)) + 0.1 * np.sin(30 * g[0])).astype(np.float32)
    h = mo.Hierarchy(u.shape, np.float32, coords)
    a = ref_x.decompose(u, coords)
    assert np.array_equal(a, mo.decompose(h, u))
    r = ref_x.compress(u, mo.ABS, 1e-2, 0.0, coords)
    q, _, _ = mo.quantize(h, a, mo.ABS, 1e-2, 0.0, np.float32(1))
    assert np.array_equal(q, r["quantized"])


def test_huffman_standalone_against_reference():
    rng = np.random.default_rng(0)
    sym = np.clip(np.round(rng.normal(4096, 30, 70000)), 0, 8191).astype(np.int64)
    ref = ref_x.huffman_compress(sym.astype(np.uint64))
    mine = [mo.huffman_compress(sym, oob_value=o) for o in (0, 0xFFFFFFFF)]
    assert ref.tobytes() in mine
    assert np.array_equal(mo.huffman_decode(mo.huffman_parse(ref.tobytes())).astype(np.int64), sym)
    assert np.array_equal(ref_x.huffman_decompress(np.frombuffer(mine[0], dtype=np.uint8), sym.size).astype(np.int64), sym)


@pytest.mark.parametrize("case", [
    ((17,), np.float32, np.inf, 1e-3), ((10, 7), np.float64, np.inf, 1e-3), ((17, 19, 21), np.float32, np.inf, 1e-4),
    ((12, 13, 14), np.float64, 0.0, 1e-3), ((5, 6, 9), np.float32, np.inf, 1e-6), ((4, 17, 5, 6), np.float64, np.inf, 1e-3),
    ((33, 20), np.float32, 1.0, 1e-2), ((5, 5, 6, 7, 5), np.float32, np.inf, 1e-3)])
def test_level_linearised_order(case):
    """Config::reorder = 1 (LevelLinearizer order, LinearQuantization.hpp:46-146): the
    reference's linearised quantised array and Huffman block; outliers as a set (the
    reference appends them in thread-block order)."""
    shape, dt, s, tol = case
    u = np.random.default_rng(0).standard_normal(shape).astype(dt)
    h = mo.Hierarchy(shape, dt)
    lin = mo.level_linear_index(h)
    assert np.array_equal(np.sort(lin.ravel()), np.arange(lin.size))
    r = ref_x.compress(u, ref_x.REL, tol, s, reorder=1)
    m = mo.compress_lowlevel(h, u, mo.REL, tol, s, dt(r["norm"]), reorder=1)
    assert np.array_equal(r["quantized"].ravel(), m["quantized"].ravel())
    ref = mo.huffman_parse(r["payload"].tobytes())
    ok = False
    for oob in (0, 0xFFFFFFFF):
        mine = mo.huffman_parse(mo.huffman_compress(m["quantized"], 8192, 20480, m["oidx"], m["oval"], oob))
        ok = ok or all(np.array_equal(mine[k], ref[k]) for k in
                       ("bits", "word_offset", "first", "entry", "keys", "ddata"))
    assert ok
    o = np.argsort(ref["oidx"])
    assert np.array_equal(ref["oidx"][o], m["oidx"]) and np.array_equal(ref["oval"][o], m["oval"])
    back = mo.decompress_lowlevel(h, r["payload"].tobytes(), mo.REL, tol, s, dt(r["norm"]), reorder=1)
    assert np.array_equal(back, ref_x.decompress(r["payload"], shape, dt, ref_x.REL, tol, s, r["norm"], reorder=1))


def _rand_coords(rng, shape, dt):
    out = []
    for n in shape:
        x = np.concatenate([[0.0], np.cumsum(rng.uniform(1, 2, n - 1))])
        out.append((x / x[-1]).astype(dt))
    return out


@pytest.mark.parametrize("shape,dt,nonuniform", [
    ((17,), np.float32, False), ((6,), np.float64, False), ((10, 7), np.float64, False), ((9, 9), np.float32, True),
    ((17, 19, 21), np.float32, False), ((12, 13, 14), np.float64, True), ((5, 6, 9), np.float32, False),
    ((64, 33), np.float32, False)])
def test_single_dimension_decomposition(shape, dt, nonuniform):
    """decomposition_type::SingleDim (DataRefactoring/SingleDimension/*.hpp), D <= 3:
    coefficients, recomposition, quantised values and Huffman block vs the reference."""
    rng = np.random.default_rng(1)
    coords = _rand_coords(rng, shape, dt) if nonuniform else None
    u = field(shape, dt, 4)
    h = mo.Hierarchy(shape, dt, coords)
    a = ref_x.decompose(u, coords, decomposition=1)
    assert np.array_equal(a, mo.decompose_single(h, u))
    assert np.array_equal(ref_x.recompose(a, coords, decomposition=1), mo.recompose_single(h, a))
    for s, tol in ((np.inf, 1e-3), (0.0, 1e-2)):
        r = ref_x.compress(u, ref_x.REL, tol, s, coords, decomposition=1)
        m = mo.compress_lowlevel(h, u, mo.REL, tol, s, dt(r["norm"]), single_dim=True)
        assert np.array_equal(m["quantized"], r["quantized"])
        ref = mo.huffman_parse(r["payload"].tobytes())
        ok = False
        for oob in (0, 0xFFFFFFFF):
            mine = mo.huffman_parse(mo.huffman_compress(m["quantized"], 8192, 20480, m["oidx"], m["oval"], oob))
            ok = ok or all(np.array_equal(mine[k], ref[k]) for k in
                           ("bits", "word_offset", "first", "entry", "keys", "ddata"))
        assert ok
        back = mo.decompress_lowlevel(h, r["payload"].tobytes(), mo.REL, tol, s, dt(r["norm"]), single_dim=True)
        assert np.array_equal(back, ref_x.decompress(r["payload"], shape, dt, ref_x.REL, tol, s, r["norm"], coords,
                                                     decomposition=1))
