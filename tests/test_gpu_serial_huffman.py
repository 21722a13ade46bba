"""Thread-per-chunk Huffman kernels (mgard_b200/csrc/huffman_serial.cuh), forced on
through mgb_tune so that small inputs reach them too: payload bytes and decoded
symbols identical to the oracle's restatement of the reference coder
(Lossless/ParallelHuffman/{Deflate,Decode}.hpp), to the reference build where it
travelled with the snapshot, and to the block-per-chunk kernels on a large input."""
import numpy as np
import pytest

import mgardx_oracle as mo
import ref_x

pytestmark = pytest.mark.gpu


@pytest.fixture()
def env():
    import torch
    import mgard_b200 as mg
    assert torch.cuda.is_available()
    mg.tune(mg.TUNE_SERIAL_MIN_CHUNKS, 0)
    yield torch, mg, torch.device("cuda:0")
    mg.tune(mg.TUNE_SERIAL_MIN_CHUNKS, 16384)


def field(shape, dtype, seed=0, noise=0.05):
    rng = np.random.default_rng(seed)
    g = np.meshgrid(*[np.linspace(0, 1, n) for n in shape], indexing="ij")
    u = sum(np.sin((3 + 2 * i) * x + i) for i, x in enumerate(g)) + noise * rng.standard_normal(shape)
    return u.astype(dtype)


def dev(torch, a, d):
    return torch.from_numpy(np.ascontiguousarray(a).copy()).to(d)


def parsed_equal(a_bytes, b_bytes):
    a, b = mo.huffman_parse(a_bytes), mo.huffman_parse(b_bytes)
    for k in ("n", "dict_size", "chunk_size"):
        assert a[k] == b[k]
    for k in ("bits", "word_offset", "first", "entry", "keys", "ddata"):
        assert np.array_equal(a[k], b[k]), k
    oa, ob = np.argsort(a["oidx"]), np.argsort(b["oidx"])
    assert np.array_equal(np.asarray(a["oidx"])[oa], np.asarray(b["oidx"])[ob])
    assert np.array_equal(np.asarray(a["oval"])[oa], np.asarray(b["oval"])[ob])


@pytest.mark.parametrize("shape,dt,dict_size,chunk", [
    ((65, 65, 65), np.float32, 8192, 20480),
    ((70, 33, 41), np.float32, 8192, 1000),    # chunk starts not 32-byte aligned
    ((70, 33, 41), np.float64, 4096, 4096),
    ((300, 250), np.float32, 8192, 77),        # odd chunk size: scalar paths
    ((9, 20, 11, 12), np.float64, 32768, 2048),  # codebook too large for shared memory
    ((5000,), np.float32, 64, 16),             # many outliers, tiny chunks
])
def test_serial_kernels_match_oracle(env, shape, dt, dict_size, chunk):
    torch, mg, d = env
    u = field(shape, dt, 3)
    cfg = mg.Config()
    cfg.huff_dict_size, cfg.huff_block_size = dict_size, chunk
    p = mg.Plan(shape, dt, config=cfg)
    h = mo.Hierarchy(shape, dt)
    for eb, tol, s in [(mo.REL, 1e-3, np.inf), (mo.ABS, 1e-2, 0.0)]:
        ref = mo.compress_lowlevel(h, u, eb, tol, s, dict_size=dict_size, chunk_size=chunk)
        payload, norm = p.compress(dev(torch, u, d), eb, tol, s)
        parsed_equal(payload.cpu().numpy().tobytes(), ref["payload"])
        # decoder, symbols out (stage API) ...
        sym, oi, ov = p.huffman_decompress(dev(torch, np.frombuffer(ref["payload"], dtype=np.uint8), d), u.size)
        q = np.asarray(ref["quantized"]).ravel().copy()
        q[np.asarray(ref["oidx"], dtype=np.int64)] = 0
        assert np.array_equal(sym.cpu().numpy().astype(np.uint16).astype(np.int64), q)
        # ... and dequantized on the fly (s = inf) / through the symbol array (s-norm)
        back = p.decompress(payload, eb, tol, s, norm).cpu().numpy()
        nrm = u.dtype.type(norm) if eb == mo.REL else np.float32(1)
        want = mo.recompose(h, mo.dequantize(h, ref["quantized"], ref["oidx"], ref["oval"], eb, tol, s, nrm,
                                             dict_size=dict_size))
        assert np.array_equal(back, want)


@pytest.mark.skipif(not ref_x.available(), reason="oracle/_ref not in the snapshot")
def test_serial_kernels_cross_decode_with_reference_build(env):
    torch, mg, d = env
    for shape, dt, eb, tol, s in [((65, 65, 65), np.float32, mo.REL, 1e-3, np.inf),
                                  ((40, 33, 50), np.float64, mo.REL, 1e-4, 0.0)]:
        u = field(shape, dt, 5)
        p = mg.Plan(shape, dt)
        payload, norm = p.compress(dev(torch, u, d), eb, tol, s)
        ours = p.decompress(payload, eb, tol, s, norm).cpu().numpy()
        theirs = ref_x.decompress(payload.cpu().numpy(), shape, dt, eb, tol, s, norm)
        assert np.array_equal(ours, theirs)
        r = ref_x.compress(u, eb, tol, s)
        assert np.array_equal(p.decompress(dev(torch, r["payload"], d), eb, tol, s, r["norm"]).cpu().numpy(),
                              ref_x.decompress(r["payload"], shape, dt, eb, tol, s, r["norm"]))


def test_serial_and_block_kernels_agree_at_scale(env):
    """257 x 513 x 513 fp32 (3303 chunks): both kernel families write the same bytes
    and decode them to the same values, skewed and flat histograms alike."""
    torch, mg, d = env
    import bench
    shape = (257, 513, 513)
    u = bench.field_torch(shape, d)
    p = mg.Plan(shape, np.float32)
    for tol in (1e-3, 3e-2):  # ~12 and ~7 bits per symbol
        outs = []
        for serial in (0, -1):
            mg.tune(mg.TUNE_SERIAL_MIN_CHUNKS, serial)
            payload, norm = p.compress(u, mo.REL, tol, np.inf)
            back = p.decompress(payload, mo.REL, tol, np.inf, norm)
            assert float((back - u).abs().max()) <= tol * norm
            sym, oi, ov = p.huffman_decompress(payload, u.numel())
            outs.append((payload.clone(), back.clone(), sym.clone()))
        # more than 65536 outliers keep their append order:
        # the blocks are compared field by field with the outlier list as a set
        parsed_equal(outs[0][0].cpu().numpy().tobytes(), outs[1][0].cpu().numpy().tobytes())
        assert torch.equal(outs[0][1], outs[1][1])
        assert torch.equal(outs[0][2], outs[1][2])
        # cross: bytes written by one family decoded by the other
        mg.tune(mg.TUNE_SERIAL_MIN_CHUNKS, 0)
        assert torch.equal(p.decompress(outs[1][0], mo.REL, tol, np.inf, norm), outs[1][1])


def test_serial_decoder_survives_corrupted_chunk_fields(env):
    """bits / word offsets come from the stream: garbage there must not fault."""
    torch, mg, d = env
    shape = (33, 40, 65)
    u = field(shape, np.float32, 1)
    p = mg.Plan(shape, np.float32)
    payload, norm = p.compress(dev(torch, u, d), mo.REL, 1e-3, np.inf)
    raw = payload.cpu().numpy().copy()
    nchunk = (u.size - 1) // 20480 + 1
    rng = np.random.default_rng(0)
    for trial in range(8):
        bad = raw.copy()
        meta = bad[24:24 + 16 * nchunk].view(np.uint64)
        k = int(rng.integers(0, 2 * nchunk))
        meta[k] = rng.integers(0, 2 ** 63, dtype=np.uint64) if trial % 2 else np.uint64(0)
        try:
            p.decompress(dev(torch, bad, d), mo.REL, 1e-3, np.inf, norm)
        except mg.MgardError:
            pass
        torch.cuda.synchronize()
    # the device is still healthy
    back = p.decompress(payload, mo.REL, 1e-3, np.inf, norm).cpu().numpy()
    assert np.abs(back - u).max() <= 1e-3 * np.abs(u).max()


def test_ring_decoder_long_codewords(env):
    """Fibonacci symbol counts give a code whose lengths run from 1 to beyond 32 bits: every
    level of the ring decoder (first table, second-level tables of both depths, canonical
    walk on the window, codewords of more than 32 bits from global memory) decodes what the
    encoder wrote, and the first formulation of the decoder agrees."""
    torch, mg, d = env
    counts = [1, 1]
    while len(counts) < 36:
        counts.append(counts[-1] + counts[-2])
    rng = np.random.default_rng(11)
    sym = np.repeat(np.arange(36, dtype=np.int64), counts)
    rng.shuffle(sym)
    cfg = mg.Config()
    cfg.huff_dict_size, cfg.huff_block_size = 64, 20480
    p = mg.Plan((sym.size,), np.float32, config=cfg)
    ds = torch.from_numpy(sym.astype(np.uint16).view(np.int16)).to(d)
    hist = torch.from_numpy(np.bincount(sym, minlength=64).astype(np.uint32).view(np.int32)).to(d)
    e = torch.empty(0, dtype=torch.int64, device=d)
    pay = p.huffman_compress(ds, hist, e, e)
    cb, _ = p.codebook(hist)
    lens = (cb.cpu().numpy().view(np.uint64) >> np.uint64(56)).astype(np.int64)
    assert lens[:36].max() > 32 and lens[:36].min() <= 2
    outs = []
    for ring in (1, 0):
        mg.tune(mg.TUNE_RING_DECODER, ring)
        back, _, _ = p.huffman_decompress(pay, sym.size)
        outs.append(back.cpu().numpy().view(np.uint16).copy())
    mg.tune(mg.TUNE_RING_DECODER, 1)
    assert np.array_equal(outs[0], sym.astype(np.uint16))
    assert np.array_equal(outs[1], outs[0])


def test_ring_decoder_more_prefixes_than_tables(env):
    """A nearly flat histogram over 8192 symbols: thousands of 12-bit prefixes continue in a
    second-level table, far more than shared memory holds - the prefixes left without one
    are decoded by the canonical walk; values equal the first formulation's."""
    torch, mg, d = env
    rng = np.random.default_rng(5)
    n = 20480 * 600
    sym = rng.integers(0, 8192, n).astype(np.int64)
    sym[rng.random(n) < 0.3] = 4096  # one short codeword among 13 / 14-bit ones
    p = mg.Plan((n,), np.float32)
    ds = torch.from_numpy(sym.astype(np.uint16).view(np.int16)).to(d)
    hist = torch.from_numpy(np.bincount(sym, minlength=8192).astype(np.uint32).view(np.int32)).to(d)
    e = torch.empty(0, dtype=torch.int64, device=d)
    pay = p.huffman_compress(ds, hist, e, e)
    back, _, _ = p.huffman_decompress(pay, n)
    assert np.array_equal(back.cpu().numpy().view(np.uint16), sym.astype(np.uint16))
