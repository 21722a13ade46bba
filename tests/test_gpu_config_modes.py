"""mgard_x::Config choices beyond the defaults (include/mgard-x/Config/Config.h:10-42):

* domain_decomposition = Block / Variable (DomainDecomposer.hpp:90-169,232-256,335-348):
  every record equals the oracle's Compressor::Compress of the same box with the local
  tolerance of ErrorToleranceCalculator.hpp:134-155, the header equals the oracle's proto3
  bytes, the stream decodes and meets the bound; device and host buffers agree;
* max_larget_level (Hierarchy.hpp:195-217): stages and payload against the reference build
  with the same limit;
* adjust_shape (CompressionHighLevel/ShapeAdjustment.hpp:43-84): the adjusted shape, and a
  stream that decodes to the same bytes."""
import struct

import numpy as np
import pytest

import mgardx_oracle as mo
import ref_x

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    import mgard_b200 as mg
    assert torch.cuda.is_available()
    return torch, mg, torch.device("cuda:0")


def field(shape, dtype=np.float32, seed=0):
    rng = np.random.default_rng(seed)
    g = np.meshgrid(*[np.linspace(0, 1, n) for n in shape], indexing="ij")
    u = sum(np.sin((3 + 2 * i) * x + i) for i, x in enumerate(g)) + 0.03 * rng.standard_normal(shape)
    return u.astype(dtype)


def local_tol(tol, norm, s, nsub, rel):
    t, n = np.float32(tol), np.float32(norm)
    if rel:
        return float(t * n) if np.isinf(s) else float(np.sqrt((t * n) * (t * n) / np.float32(nsub)))
    return float(t) if np.isinf(s) else float(np.sqrt((t * t) / np.float32(nsub)))


def check_records(stream, hb, u, boxes, ltol, s):
    raw = stream.tobytes()
    off = hb
    for sl in boxes:
        size = struct.unpack_from("<Q", raw, off)[0]
        sub = np.ascontiguousarray(u[sl])
        ref = mo.compress_lowlevel(mo.Hierarchy(sub.shape, u.dtype), sub, mo.ABS, ltol, s)
        if size == sub.nbytes:
            assert len(ref["payload"]) >= sub.nbytes and raw[off + 8:off + 8 + size] == sub.tobytes()
        else:
            assert raw[off + 8:off + 8 + size] == ref["payload"], sl
        off += 8 + size
    assert off == len(raw)


@pytest.mark.parametrize("mode,tol,s", [(mo.REL, 1e-3, np.inf), (mo.ABS, 1e-2, 0.0), (mo.REL, 1e-2, 0.0)])
def test_block_decomposition(env, mode, tol, s):
    torch, mg, d = env
    shape, B = (40, 37, 51), 16  # remainders 8, 5, 3
    u = field(shape, seed=1)
    cfg = mg.Config()
    cfg.domain_decomposition = mg.domain_decomposition_type.Block
    cfg.block_size = B
    stream = mg.compress(u, tol, s, mode, config=cfg)
    info = mg.peek_header(stream)
    cuts = [range(0, n, B) for n in shape]
    boxes = [tuple(slice(o, min(o + B, n)) for o, n in zip((a, b, c), shape))
             for a in cuts[0] for b in cuts[1] for c in cuts[2]]
    assert len(boxes) == 3 * 3 * 4
    norm = info["norm"] if mode == mo.REL else 1.0
    if mode == mo.REL and np.isinf(s):
        assert norm == float(np.abs(u).max())
    check_records(stream, info["header_bytes"], u, boxes, local_tol(tol, norm, s, len(boxes), mode == mo.REL), s)
    hdr = mo.encode_preamble(mo.encode_header(shape, np.float32, mode, tol, s, np.float32(norm), None, True, 0, B,
                                              dd_method=2))
    assert stream[:info["header_bytes"]].tobytes() == hdr
    back = mg.decompress(stream)
    if np.isinf(s):
        assert np.abs(back - u).max() <= tol * (np.abs(u).max() if mode == mo.REL else 1)
    else:
        bound = tol * (np.sqrt((u.astype(np.float64) ** 2).mean()) if mode == mo.REL else 1)
        assert np.sqrt(((back.astype(np.float64) - u) ** 2).mean()) <= bound
    ds = mg.compress(torch.from_numpy(u).to(d), tol, s, mode, config=cfg)
    assert ds.cpu().numpy().tobytes() == stream.tobytes()
    assert np.array_equal(mg.decompress(ds).cpu().numpy(), back)


def test_block_decomposition_2d_and_4d(env):
    torch, mg, d = env
    for shape, B in (((70, 45), 32), ((9, 20, 11, 12), 8)):  # remainders 6, 13 / 0 (one cut), 4, 3, 4
        u = field(shape, np.float64, seed=len(shape))
        cfg = mg.Config()
        cfg.domain_decomposition = mg.domain_decomposition_type.Block
        cfg.block_size = B
        if any(n % B in (1, 2) for n in shape):
            with pytest.raises(mg.MgardError):
                mg.compress(u, 1e-3, np.inf, mo.REL, config=cfg)
            continue
        stream = mg.compress(u, 1e-3, np.inf, mo.REL, config=cfg)
        back = mg.decompress(stream)
        assert back.dtype == np.float64 and np.abs(back - u).max() <= 1e-3 * np.abs(u).max()


def test_variable_decomposition(env):
    torch, mg, d = env
    shape, sizes = (40, 33, 20), [20, 13, 7]
    u = field(shape, seed=3)
    cfg = mg.Config()
    cfg.domain_decomposition = mg.domain_decomposition_type.Variable
    cfg.domain_decomposition_dim = 0
    cfg.domain_decomposition_sizes = sizes
    tol, s = 1e-3, np.inf
    stream = mg.compress(u, tol, s, mo.REL, config=cfg)
    info = mg.peek_header(stream)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    boxes = [(slice(int(a), int(b)),) for a, b in zip(offs[:-1], offs[1:])]
    check_records(stream, info["header_bytes"], u, boxes, local_tol(tol, info["norm"], s, 3, True), s)
    # the extents are not in the stream: the decompressing call repeats them
    back = mg.decompress(stream, config=cfg)
    assert np.abs(back - u).max() <= tol * np.abs(u).max()
    with pytest.raises(mg.MgardError):
        mg.decompress(stream)
    # a cut along another dimension
    cfg.domain_decomposition_dim = 2
    cfg.domain_decomposition_sizes = [9, 11]
    st2 = mg.compress(u, tol, s, mo.REL, config=cfg)
    assert np.abs(mg.decompress(st2, config=cfg) - u).max() <= tol * np.abs(u).max()


@pytest.mark.skipif(not ref_x.available(), reason="oracle/_ref not in the snapshot")
@pytest.mark.parametrize("shape,dtype,level", [((33, 40, 65), np.float32, 2), ((100, 90), np.float64, 3),
                                               ((65, 65, 65), np.float32, 1)])
def test_max_larget_level_against_reference(env, shape, dtype, level):
    torch, mg, d = env
    u = field(shape, dtype, seed=7)
    cfg = mg.Config()
    cfg.max_larget_level = level
    p = mg.Plan(shape, dtype, config=cfg)
    assert p.l_target == level
    du = torch.from_numpy(u).to(d)
    for eb, tol, s in ((mo.REL, 1e-3, np.inf), (mo.ABS, 1e-2, 0.0)):
        r = ref_x.compress(u, eb, tol, s, max_level=level)
        assert r["l_target"] == level
        assert np.array_equal(p.decompose(du).cpu().numpy(), r["decomposed"])
        nrm = r["norm"] if eb == mo.REL else 1.0
        sym, hist, oi, ov = p.quantize(p.decompose(du), eb, tol, s, nrm)
        q = r["quantized"]
        assert np.array_equal(sym.cpu().numpy().astype(np.uint16).astype(np.int64).reshape(shape), q)
        payload, norm = p.compress(du, eb, tol, s)
        if eb == mo.ABS or np.isinf(s):
            # the reference's code lengths depend on a word read past its frequency array
            # (GenerateCL.hpp:252-257): the engine writes the
            # oracle's variant 0, the reference one of the two
            a = mo.huffman_parse(payload.cpu().numpy().tobytes())
            oidx, oval = np.asarray(a["oidx"]), np.asarray(a["oval"])
            variants = [mo.huffman_compress(q, 8192, 20480, oidx, oval, oob_value=o) for o in (0, 0xFFFFFFFF)]
            assert payload.cpu().numpy().tobytes() == variants[0]
            b = mo.huffman_parse(r["payload"])
            assert any(np.array_equal(b["ddata"], mo.huffman_parse(v)["ddata"]) for v in variants)
            assert np.array_equal(np.sort(oidx), np.sort(np.asarray(b["oidx"])))
        ours = p.decompress(torch.from_numpy(r["payload"]).to(d), eb, tol, s, r["norm"] if eb == mo.REL else 1.0)
        theirs = ref_x.decompress(r["payload"], shape, dtype, eb, tol, s, r["norm"], max_level=level)
        assert np.array_equal(ours.cpu().numpy(), theirs)
    # through the high-level API: the limit is not stored in the stream
    st = mg.compress(u, 1e-3, np.inf, mo.REL, config=cfg)
    assert np.abs(mg.decompress(st, config=cfg) - u).max() <= 1e-3 * np.abs(u).max()


def test_adjust_shape(env):
    torch, mg, d = env
    # ShapeAdjustment.hpp:43-84: 360 = 2*2*2*3*3*5 -> largest factors to the smallest dims
    assert mg.adjust_shape((360, 4, 5)) == (15, 24, 20)
    shape = (360, 4, 5)
    new = mg.adjust_shape(shape)
    assert np.prod(new) == np.prod(shape) and max(new) < 360
    u = field(shape, seed=2)
    cfg = mg.Config()
    cfg.adjust_shape = True
    st = mg.compress(u, 1e-3, np.inf, mo.REL, config=cfg)
    info = mg.peek_header(st)
    assert info["shape"] == new
    back = mg.decompress(st)
    assert back.shape == new
    assert np.abs(back.reshape(shape) - u).max() <= 1e-3 * np.abs(u).max()
    same = mg.compress(u.reshape(new), 1e-3, np.inf, mo.REL)
    assert same.tobytes() == st.tobytes()
