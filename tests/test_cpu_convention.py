"""MGARD-CPU convention (mgard::compress / mgard::decompress).

CPU tests pin the numpy restatement (oracle/mgardcpu_oracle.py) to the reference's
own known-answer vectors and, when it has been built, to the compiled reference
(oracle/_ref/libmgard_cpu_ref.so).  GPU tests compare the CUDA path, through the
C ABI, with both -- bit for bit: coefficients, int64 quanta, zlib payload, header.
"""
import math
import os
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)

import mgardcpu_oracle as mo  # noqa: E402
import ref_cpu  # noqa: E402

needs_ref = pytest.mark.skipif(not ref_cpu.available(), reason="oracle/_ref/libmgard_cpu_ref.so not built")


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def random_coords(rng, shape, dt):
    """Spacings uniform in [1, 2), as the reference's tests use
    (tests/src/test_compress.cpp:32-34), scaled to [0, 1]."""
    out = []
    for n in shape:
        if n == 1:
            out.append(np.zeros(1, dtype=dt))
            continue
        x = np.concatenate([[0.0], np.cumsum(rng.uniform(1, 2, n - 1))])
        out.append((x / x[-1]).astype(dt))
    return out


SHAPES = [
    ((17,), False), ((33, 17), False), ((9, 9, 9), False), ((20,), True), ((10, 7), True),
    ((33, 20, 17), True), ((5, 6, 7, 9), True), ((12, 1, 9), False), ((129, 65), False),
    ((100, 65), True), ((3, 3), False), ((2, 5), False), ((1, 31, 1, 18), True),
]


# ----------------------------------------------------------------------------
# CPU: oracle vs the reference's known answers and the compiled reference
# ----------------------------------------------------------------------------

def test_oracle_shuffle_known_answers():
    # reference tests/src/test_shuffle.cpp:33-50
    for shape, dt, expected in [
        ((9,), np.float32, [0, 8, 4, 2, 6, 1, 3, 5, 7]),
        ((6, 4), np.float32, [0, 3, 8, 11, 20, 23, 1, 4, 5, 7, 9, 12, 13, 15, 21, 2, 6, 10, 14, 16, 17, 18, 19, 22]),
        ((3, 2, 2), np.float64, list(range(12))),
    ]:
        h = mo.Hierarchy(shape, dt)
        u = np.arange(h.ndof(), dtype=dt).reshape(shape)
        assert mo.shuffle(h, u).tolist() == expected
        assert np.array_equal(mo.unshuffle(h, mo.shuffle(h, u)), u)


def test_oracle_decompose_known_answers():
    # reference tests/src/test_decompose.cpp:277-337 (tolerance 1e-4 as there)
    u1 = [10, 3, -8, -6, 3, 0, -5, 0, 0, -2, -8, -5, -10, -7, 8, -2, 3, -1, 0, 9, -4, -6, -8, -5, -10, 1, 3, 7, -8,
          1, 10, -2, 8]
    exp1 = [
        [10.0, 3.0],
        [11.0, 2.0, -7.0],
        [4.4375, 2.0, -14.5, -3.5, -6.687500000000001],
        [0.4374999999999991, 2.0, -15.678571428571429, -3.5, -4.625000000000002, 1.0, -5.321428571428571, 2.5,
         -2.4375000000000004],
        [-0.95703125, 2.0, -15.652199926362297, -3.5, -4.122767857142856, 1.0, -4.978599042709867, 2.5,
         -4.765625000000001, 2.0, -1.173186671575852, 4.0, -10.689732142857139, -6.0, 9.303985640648008, -7.5,
         -3.1054687499999987],
    ]
    for L, expected in enumerate(exp1):
        n = (1 << L) + 1
        h = mo.Hierarchy((n,), np.float32)
        got = mo.decompose_nodal(h, np.array(u1[:n], dtype=np.float32))
        assert np.allclose(got, expected, rtol=1e-4, atol=1e-6)
    u2 = [7, 4, 5, -10, -6, 6, -8, -5, 6, -2, 2, -2, 9, 2, -10, 3, 8, -8, -3, 7, -8, -9, -6, -1, -4]
    exp2 = [
        [7.0, 4.0, 5.0, -10.0],
        [0.9999999999999973, -2.0, 3.9999999999999982, -9.5, -8.5, 0.5, -15.000000000000004, -4.0,
         3.9999999999999973],
        [3.8007812499999973, -2.0, -2.9062499999998854, -9.5, -2.910156250000001, 1.5, -13.75, -12.0, 6.5, 6.0,
         -1.593749999999881, -7.5, 2.8750000000004396, 2.5, -1.0312499999998854, 6.0, 8.75, -9.5, -0.25, 14.0,
         -2.5039062500000013, -2.0, -10.218749999999885, 4.0, -0.6992187500000024],
    ]
    for L, expected in enumerate(exp2):
        n = (1 << L) + 1
        h = mo.Hierarchy((n, n), np.float64)
        got = mo.decompose_nodal(h, np.array(u2[:n * n], dtype=np.float64).reshape(n, n))
        assert np.allclose(got.ravel(), expected, rtol=1e-4, atol=1e-9)
        back = mo.recompose_nodal(h, got)
        assert np.allclose(back.ravel(), u2[:n * n], rtol=1e-12, atol=1e-12)


def test_oracle_hierarchy_levels():
    # level shapes of reference tests/src/test_TensorMeshHierarchy.cpp:19-52 style cases
    h = mo.Hierarchy((129, 129, 129), np.float64)
    assert h.L == 7 and h.ndof(0) == 8
    h = mo.Hierarchy((1000, 1000), np.float32)
    assert h.L == 10 and h.shapes[9] == (513, 513)
    h = mo.Hierarchy((5, 3), np.float32)
    assert h.L == 1 and h.shapes[0] == (3, 2)
    with pytest.raises(ValueError):
        mo.Hierarchy((1, 1), np.float32)


def test_oracle_header_matches_protobuf_golden():
    # tests/golden/cpu_headers.npz: bytes produced by python protobuf from the
    # reference's src/mgard.proto (tests/golden/make_cpu_header_golden.py)
    g = np.load(os.path.join(ROOT, "tests", "golden", "cpu_headers.npz"), allow_pickle=False)
    n = int(g["count"])
    assert n >= 4
    for i in range(n):
        shape = tuple(int(x) for x in g[f"shape{i}"])
        dt = np.float64 if int(g[f"dtype{i}"]) == 1 else np.float32
        coords = None
        if int(g[f"explicit{i}"]):
            flat = g[f"coords{i}"]
            coords, off = [], 0
            for m in shape:
                coords.append(flat[off:off + m].astype(dt))
                off += m
        h = mo.Hierarchy(shape, dt, coords)
        got = mo.header_bytes(h, float(g[f"s{i}"]), float(g[f"tol{i}"]))
        assert got == g[f"bytes{i}"].tobytes()


def test_preamble_is_big_endian_known_answers():
    """MGARD-CPU framing: header size and CRC32 are stored big-endian
    (include/format.tpp:11-41).  Known answers of the reference's own test
    (tests/src/test_format.cpp:23-50), checked on the oracle's packing and, when
    built, on the reference's serialize<> / deserialize<> templates themselves."""
    import struct
    assert struct.pack(">Q", 144965140814303507) == bytes([0x02, 0x03, 0x05, 0x07, 0x0b, 0x0d, 0x11, 0x13])
    assert struct.pack(">I", 2017) == bytes([0x00, 0x00, 0x07, 0xe1])
    assert struct.pack(">I", 13117532) == bytes([0x00, 0xc8, 0x28, 0x5c])
    assert struct.unpack(">Q", bytes([0xa1, 0xb2, 0xc3, 0xd4, 0xe5, 0xf6, 0x07, 0x18]))[0] == 11651590505119483672
    hdr = bytes(range(247))
    pre = mo.preamble(hdr)
    assert pre[:5] == b"MGARD" and pre[5:13] == bytes([0, 0, 0, 0, 0, 0, 0, 0xf7])
    assert pre[13:17] == struct.pack(">I", zlib.crc32(hdr))
    assert mo.read_preamble(pre + hdr) == (247, zlib.crc32(hdr))
    if ref_cpu.available():
        rng = np.random.default_rng(3)
        for n in (1, 20, 247, 300, 70000):
            body = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
            assert ref_cpu.preamble(body) == mo.preamble(body)
            assert ref_cpu.read_preamble(mo.preamble(body)) == (n, zlib.crc32(body))
        assert ref_cpu.read_preamble(b"MGARD" + bytes([0xa1, 0xb2, 0xc3, 0xd4, 0xe5, 0xf6, 0x07, 0x18,
                                                     0x00, 0xac, 0x00, 0x00])) == (11651590505119483672, 11272192)


def test_library_writes_and_reads_the_cpu_framing():
    """mgb_cpu_write_header (host only): bytes equal to the oracle's stream head --
    big-endian preamble included -- for uniform and explicit grids; the parser takes
    big-endian framing for MGARD-CPU headers and little-endian for MGARD-X headers
    only (each reference reader accepts its own byte order, src/format.cpp:160-175 vs
    src/mgard-x/Metadata/Metadata.cpp:475-500)."""
    import ctypes as C
    import struct
    from mgard_b200 import _lib
    import mgard_b200 as mg
    L = _lib.lib()
    rng = np.random.default_rng(11)
    for shape, explicit, dt, s, tol, comp in (((17, 9), False, np.float32, math.inf, 1e-3, 1),
                                               ((33, 20, 17), True, np.float64, 0.0, 1e-2, 2),
                                               ((300,), True, np.float32, 1.0, 0.5, 2)):
        coords = random_coords(rng, shape, dt) if explicit else None
        h = mo.Hierarchy(shape, dt, coords)
        want = mo.stream(h, s, tol, b"", comp)
        carr = None
        if coords is not None:
            carr = (C.c_void_p * len(shape))(*[c.ctypes.data for c in coords])
        out = np.zeros(1 << 16, dtype=np.uint8)
        sz = C.c_uint64(0)
        rc = L.mgb_cpu_write_header(len(shape), 0 if dt == np.float32 else 1, (C.c_uint64 * len(shape))(*shape),
                                    carr, s, tol, comp, out.ctypes.data, out.size, C.byref(sz))
        assert rc == 0
        got = out[:sz.value].tobytes()
        assert got == want
        hs = len(got) - 17
        assert got[5:13] == struct.pack(">Q", hs) and got[13:17] == struct.pack(">I", zlib.crc32(got[17:]))
        if ref_cpu.available():
            assert got[:17] == ref_cpu.preamble(got[17:])
        info = mg.peek_header(np.frombuffer(got + b"\0" * 16, dtype=np.uint8))
        assert list(info["shape"]) == list(shape) and info["header_bytes"] == len(got)
        # the same header framed little-endian is not an MGARD-CPU stream
        le = b"MGARD" + struct.pack("<Q", hs) + struct.pack("<I", zlib.crc32(got[17:])) + got[17:]
        with pytest.raises(mg.MgardError):
            mg.peek_header(np.frombuffer(le + b"\0" * 16, dtype=np.uint8))
    # and an MGARD-X header framed big-endian is rejected as well
    z = np.load(os.path.join(ROOT, "tests", "golden", "headers.npz"))
    hdr = z["hdr0"].tobytes()
    good = b"MGARD" + struct.pack("<Q", len(hdr)) + struct.pack("<I", zlib.crc32(hdr)) + hdr
    assert mg.peek_header(np.frombuffer(good + b"\0" * 16, dtype=np.uint8))["header_bytes"] == len(good)
    be = b"MGARD" + struct.pack(">Q", len(hdr)) + struct.pack(">I", zlib.crc32(hdr)) + hdr
    with pytest.raises(mg.MgardError):
        mg.peek_header(np.frombuffer(be + b"\0" * 16, dtype=np.uint8))


@needs_ref
@pytest.mark.parametrize("shape,explicit", SHAPES)
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_oracle_matches_reference_build(shape, explicit, dt):
    rng = np.random.default_rng(hash((shape, explicit)) & 0xffff)
    coords = random_coords(rng, shape, dt) if explicit else None
    h = mo.Hierarchy(shape, dt, coords)
    L, ndof = ref_cpu.info(shape, dt, coords)
    assert L == h.L and ndof == [h.ndof(l) for l in range(h.L + 1)]
    u = rng.standard_normal(shape).astype(dt)
    assert bits_equal(ref_cpu.shuffle(u, coords), mo.shuffle(h, u))
    c_ref = ref_cpu.decompose(u, coords)
    assert bits_equal(c_ref, mo.decompose(h, u))
    assert bits_equal(ref_cpu.recompose(c_ref, shape, coords), mo.recompose(h, c_ref))
    for s in (math.inf, 0.0, 1.0, -0.5):
        q_ref = ref_cpu.quantize(c_ref, shape, s, 1e-3, coords)
        assert np.array_equal(q_ref, mo.quantize(h, s, 1e-3, c_ref))
        assert bits_equal(ref_cpu.dequantize(q_ref, shape, dt, s, 1e-3, coords), mo.dequantize(h, s, 1e-3, q_ref))
    q = ref_cpu.quantize(c_ref, shape, math.inf, 1e-3, coords)
    assert ref_cpu.zlib_compress(q).tobytes() == mo.zlib_payload(q)


needs_ref_zstd = pytest.mark.skipif(not ref_cpu.zstd_available(),
                                    reason="oracle/_ref/libmgard_cpu_ref_zstd.so not built")


def _quanta_cases():
    rng = np.random.default_rng(5)
    cases = []
    for n, scale in [(200000, 30), (5000, 3), (17, 1), (1, 1), (70000, 2000), (3000, 0), (40000, 1e5)]:
        q = np.round(rng.standard_normal(n) * scale).astype(np.int64)
        if n >= 70000:  # misses, and the edges of the in-range window (0, 131072) after the shift
            q[::5000] = 10 ** 7
            q[7], q[9], q[10], q[11], q[12] = -10 ** 6, 65535, 65536, -65536, -65535
        cases.append(q)
    return cases


@needs_ref_zstd
def test_oracle_huffman_zstd_matches_reference_build():
    """CPU_HUFFMAN_ZSTD payload: tree ties (std::priority_queue order), MSB-first
    32-bit packing, miss list, zstd frame -- byte for byte."""
    for q in _quanta_cases():
        ref = ref_cpu.huffman_zstd_compress(q).tobytes()
        assert mo.huffman_zstd_payload(q) == ref
        assert np.array_equal(ref_cpu.huffman_zstd_decompress(np.frombuffer(ref, np.uint8), q.size), q)


def test_oracle_huffman_codes_are_prefix_free_and_optimal_length():
    rng = np.random.default_rng(3)
    ft = np.zeros(mo.NQL, dtype=np.int64)
    ft[rng.integers(1, mo.NQL, 300)] = rng.integers(1, 1000, 300)
    codes = mo.huffman_codes(ft)
    words = sorted(format(c, "b").zfill(l) for c, l in codes.values())
    assert all(not b.startswith(a) for a, b in zip(words, words[1:]))
    assert abs(sum(2.0 ** -l for _, l in codes.values()) - 1.0) < 1e-12  # Kraft equality


def test_cpu_convention_symbols_and_no_gpu_behaviour():
    import torch
    from mgard_b200 import _lib
    lib = _lib.lib()
    for name in ("mgb_cpu_plan_create", "mgb_cpu_compress", "mgb_cpu_decompress", "mgb_cpu_decompose",
                 "mgb_cpu_recompose", "mgb_cpu_quantize", "mgb_cpu_dequantize", "mgb_cpu_shuffle",
                 "mgb_cpu_unshuffle"):
        assert hasattr(lib, name)
    if not torch.cuda.is_available():
        import mgard_b200.cpu as mc
        with pytest.raises(_lib.MgardError) as e:
            mc.TensorMeshHierarchy((9, 9))
        assert e.value.status == _lib.BACKEND_NOT_AVAILABLE  # no CPU fallback
        h = mo.Hierarchy((9, 9), np.float64)
        blob = mo.compress(h, np.zeros((9, 9)), math.inf, 1e-3)
        with pytest.raises(_lib.MgardError):
            mc.decompress(blob)


# ----------------------------------------------------------------------------
# GPU: CUDA path through the C ABI vs the oracle / the compiled reference
# ----------------------------------------------------------------------------

def _ref_or_oracle_decompose(h, u, coords):
    return ref_cpu.decompose(u, coords) if ref_cpu.available() else mo.decompose(h, u)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,explicit", SHAPES)
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_gpu_stages_bit_exact(shape, explicit, dt):
    import torch
    import mgard_b200.cpu as mc
    rng = np.random.default_rng(hash((shape, explicit, 7)) & 0xffff)
    coords = random_coords(rng, shape, dt) if explicit else None
    h = mo.Hierarchy(shape, dt, coords)
    H = mc.TensorMeshHierarchy(shape, coords, dt)
    assert H.L == h.L
    assert [H.ndof(l) for l in range(h.L + 1)] == [h.ndof(l) for l in range(h.L + 1)]
    assert [H.level_shape(l) for l in range(h.L + 1)] == [tuple(s) for s in h.shapes]
    u = rng.standard_normal(shape).astype(dt)
    du = torch.from_numpy(u).cuda()
    assert bits_equal(H.shuffle(du).cpu().numpy(), mo.shuffle(h, u))
    assert bits_equal(H.unshuffle(H.shuffle(du)).cpu().numpy(), u)
    c_ref = _ref_or_oracle_decompose(h, u, coords)
    c_gpu = H.decompose(du)
    assert bits_equal(c_gpu.cpu().numpy(), c_ref)
    assert bits_equal(c_ref, mo.decompose(h, u))
    r_ref = ref_cpu.recompose(c_ref, shape, coords) if ref_cpu.available() else mo.recompose(h, c_ref)
    assert bits_equal(H.recompose(c_gpu).cpu().numpy(), r_ref)
    for s in (math.inf, 0.0, 1.0, -0.5):
        q_ref = mo.quantize(h, s, 1e-3, c_ref)
        if ref_cpu.available():
            assert np.array_equal(q_ref, ref_cpu.quantize(c_ref, shape, s, 1e-3, coords))
        q_gpu = H.quantize(c_gpu, s, 1e-3)
        assert np.array_equal(q_gpu.cpu().numpy(), q_ref)
        assert bits_equal(H.dequantize(q_gpu, s, 1e-3).cpu().numpy(), mo.dequantize(h, s, 1e-3, q_ref))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,explicit,dt,s,tol", [
    ((33, 20, 17), True, np.float64, math.inf, 1e-3),
    ((100, 65), True, np.float32, 0.0, 1e-2),
    ((65, 65, 65), False, np.float32, math.inf, 1e-3),
    ((300,), False, np.float64, 1.0, 1e-2),
    ((5, 6, 7, 9), True, np.float64, -0.5, 1e-2),
])
def test_gpu_compress_stream_identical_and_round_trip(shape, explicit, dt, s, tol):
    import torch
    import mgard_b200.cpu as mc
    rng = np.random.default_rng(11)
    coords = random_coords(rng, shape, dt) if explicit else None
    grids = np.meshgrid(*[np.linspace(0, 1, n) for n in shape], indexing="ij")
    u = sum(np.sin((3 + k) * g) for k, g in enumerate(grids)).astype(dt) + 0.01 * rng.standard_normal(shape).astype(dt)
    h = mo.Hierarchy(shape, dt, coords)
    H = mc.TensorMeshHierarchy(shape, coords, dt)
    blob = mc.compress(H, u, s, tol)
    expect = mo.compress(h, u, s, tol)
    assert blob == expect  # preamble + protobuf header + zlib payload, byte for byte
    blob_dev = mc.compress(H, torch.from_numpy(u).cuda(), s, tol)
    assert blob_dev == blob
    if ref_cpu.available():
        q = ref_cpu.quantize(ref_cpu.decompose(u, coords), shape, s, tol, coords)
        hdr = len(blob) - len(ref_cpu.zlib_compress(q))
        assert blob[hdr:] == ref_cpu.zlib_compress(q).tobytes()
        back_ref = ref_cpu.recompose(ref_cpu.dequantize(q, shape, dt, s, tol, coords), shape, coords)
    else:
        q = mo.quantize(h, s, tol, mo.decompose(h, u))
        back_ref = mo.recompose(h, mo.dequantize(h, s, tol, q))
    back = mc.decompress(blob)
    assert back.dtype == np.dtype(dt) and back.shape == tuple(shape)
    assert bits_equal(back, back_ref)
    if math.isinf(s):
        assert np.abs(back.astype(np.float64) - u.astype(np.float64)).max() <= tol


@pytest.mark.gpu
def test_gpu_c1_129cubed_fp64():
    """BASELINE config C1: 129^3 fp64, ABS 1e-4, s = inf, CPU convention."""
    import mgard_b200.cpu as mc
    n = 129
    x = np.arange(n) / (n - 1)
    x0, x1, x2 = np.meshgrid(x, x, x, indexing="ij")
    u = np.sin(2 * np.pi * x0) * np.cos(3 * np.pi * x1) + 0.5 * np.sin(5 * np.pi * x2) + 0.25 * x0 * x1
    H = mc.TensorMeshHierarchy((n, n, n), None, np.float64)
    assert H.L == 7
    blob = mc.compress(H, u, math.inf, 1e-4)
    if ref_cpu.available():
        q = ref_cpu.quantize(ref_cpu.decompose(u), u.shape, math.inf, 1e-4)
        payload = ref_cpu.zlib_compress(q).tobytes()
        h = mo.Hierarchy(u.shape, np.float64)
        assert blob == mo.stream(h, math.inf, 1e-4, payload)
    back = mc.decompress(blob)
    assert np.abs(back - u).max() <= 1e-4
    assert u.nbytes / len(blob) > 5


@pytest.mark.gpu
def test_gpu_c3_nonuniform_1000sq_fp32():
    """BASELINE config C3 under the CPU convention: 1000^2 fp32, non-uniform
    coordinates, s = 0; L = 10 with level 9 = 513^2."""
    import mgard_b200.cpu as mc
    n = 1000
    coords = []
    for k in (7, 11):
        hsp = 1 + 0.5 * np.sin(2 * np.pi * k * np.arange(n - 1) / 999)
        xx = np.concatenate([[0.0], np.cumsum(hsp)])
        coords.append((xx / xx[-1]).astype(np.float32))
    x0, x1 = np.meshgrid(coords[0].astype(np.float64), coords[1].astype(np.float64), indexing="ij")
    u = (np.exp(-8 * ((x0 - .5) ** 2 + (x1 - .4) ** 2)) + 0.1 * np.sin(30 * x0)).astype(np.float32)
    H = mc.TensorMeshHierarchy((n, n), coords, np.float32)
    assert H.L == 10 and H.level_shape(9) == (513, 513)
    blob = mc.compress(H, u, 0.0, 1e-2)
    if ref_cpu.available():
        c = ref_cpu.decompose(u, coords)
        q = ref_cpu.quantize(c, u.shape, 0.0, 1e-2, coords)
        h = mo.Hierarchy(u.shape, np.float32, coords)
        assert blob == mo.stream(h, 0.0, 1e-2, ref_cpu.zlib_compress(q).tobytes())
        back_ref = ref_cpu.recompose(ref_cpu.dequantize(q, u.shape, np.float32, 0.0, 1e-2, coords), u.shape, coords)
        assert bits_equal(mc.decompress(blob), back_ref)
    back = mc.decompress(blob)
    # L2 error well inside the tolerance (the s = 0 bound is on the mass-weighted L2 norm)
    assert math.sqrt(np.mean((back.astype(np.float64) - u) ** 2)) <= 1e-2


@pytest.mark.gpu
def test_gpu_cpu_convention_errors():
    import mgard_b200.cpu as mc
    from mgard_b200 import _lib
    import mgard_b200 as mg
    with pytest.raises(_lib.MgardError):
        mc.TensorMeshHierarchy((1, 1))
    with pytest.raises(_lib.MgardError):
        mc.TensorMeshHierarchy((4,), [np.array([0.0, 0.5, 0.25, 1.0])])
    H = mc.TensorMeshHierarchy((17, 17), None, np.float64)
    u = np.full((17, 17), 1e300)
    with pytest.raises(_lib.MgardError):  # "number too large to be quantized"
        mc.compress(H, u, math.inf, 1e-300)
    blob = mc.compress(H, np.ones((17, 17)), math.inf, 1e-3)
    bad = bytearray(blob)
    bad[20] ^= 0xff
    with pytest.raises(_lib.MgardError):
        mc.decompress(bytes(bad))
    with pytest.raises(_lib.MgardError):
        mc.decompress(blob[:-5])
    # an MGARD-X stream is not a CPU stream and vice versa
    xs = mg.compress(np.ones((17, 17), dtype=np.float32), 1e-3, math.inf, mg.error_bound_type.ABS)
    with pytest.raises(_lib.MgardError):
        mc.decompress(bytes(xs))
    with pytest.raises(_lib.MgardError):
        mg.decompress(np.frombuffer(blob, dtype=np.uint8))


@pytest.mark.gpu
def test_gpu_cxx_cpu_api_mirror(tmp_path):
    """include/mgard_b200/compress.hpp (mgard::compress / decompress mirror)."""
    import subprocess
    exe = tmp_path / "cpu_api_roundtrip"
    subprocess.check_call(["g++", "-std=c++17", f"-I{ROOT}/include", f"{ROOT}/tests/cxx/cpu_api_roundtrip.cpp",
                           "-o", str(exe), f"-L{ROOT}/mgard_b200", "-lmgard_b200",
                           f"-Wl,-rpath,{ROOT}/mgard_b200", "-L/usr/local/cuda/lib64", "-lcudart"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "self-describing decompress identical" in out.stdout and "degenerate shape rejected" in out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("shape,explicit,dt,s,tol", [
    ((33, 20, 17), True, np.float64, math.inf, 1e-3),
    ((100, 65), True, np.float32, 0.0, 1e-2),
    ((65, 65, 65), False, np.float32, math.inf, 1e-6),   # wide quanta: many misses
    ((300,), False, np.float64, 1.0, 1e-2),
    ((40, 40), False, np.float64, math.inf, 10.0),       # every quantum is zero: one-leaf tree
])
def test_gpu_huffman_zstd_stream_identical_and_round_trip(shape, explicit, dt, s, tol):
    """CPU_HUFFMAN_ZSTD (the reference's default lossless stage): histogram and bit
    packing on the GPU, tree on the host -- stream equal to the oracle's and, where
    built, to the reference's compress_memory_huffman."""
    import mgard_b200.cpu as mc
    rng = np.random.default_rng(13)
    coords = random_coords(rng, shape, dt) if explicit else None
    grids = np.meshgrid(*[np.linspace(0, 1, n) for n in shape], indexing="ij")
    u = sum(np.sin((3 + k) * g) for k, g in enumerate(grids)).astype(dt) + 0.01 * rng.standard_normal(shape).astype(dt)
    h = mo.Hierarchy(shape, dt, coords)
    H = mc.TensorMeshHierarchy(shape, coords, dt)
    blob = mc.compress(H, u, s, tol, mc.CPU_HUFFMAN_ZSTD)
    q = mo.quantize(h, s, tol, mo.decompose(h, u))
    assert blob == mo.stream(h, s, tol, mo.huffman_zstd_payload(q), 2)
    if ref_cpu.zstd_available():
        assert blob.endswith(ref_cpu.huffman_zstd_compress(q).tobytes())
    back = mc.decompress(blob)
    assert bits_equal(back, mo.recompose(h, mo.dequantize(h, s, tol, q)))
    assert bits_equal(back, mc.decompress(mc.compress(H, u, s, tol, mc.CPU_HUFFMAN_ZLIB)))
    if math.isinf(s):
        assert np.abs(back.astype(np.float64) - u.astype(np.float64)).max() <= tol


@pytest.mark.gpu
def test_gpu_reads_reference_written_huffman_zstd_stream():
    """A stream whose payload the reference itself wrote decodes to the
    reference's reconstruction."""
    if not (ref_cpu.available() and ref_cpu.zstd_available()):
        pytest.skip("reference builds not present")
    import mgard_b200.cpu as mc
    shape, dt, s, tol = (33, 20, 17), np.float64, math.inf, 1e-3
    rng = np.random.default_rng(17)
    u = rng.standard_normal(shape)
    q = ref_cpu.quantize(ref_cpu.decompose(u), shape, s, tol)
    h = mo.Hierarchy(shape, dt)
    blob = mo.stream(h, s, tol, ref_cpu.huffman_zstd_compress(q).tobytes(), 2)
    expect = ref_cpu.recompose(ref_cpu.dequantize(q, shape, dt, s, tol), shape)
    assert bits_equal(mc.decompress(blob), expect)


def _cli():
    exe = os.path.join(ROOT, "mgard_b200", "mgard-b200")
    if not os.path.exists(exe):
        pytest.skip("CLI not built")
    return exe


def test_mgard_cli_usage_and_missing_backend(tmp_path):
    """mgard-b200 (the reference `mgard` executable's sub-commands, src/cli/executable.cpp):
    usage text; without a CUDA device a request fails loudly (status 5), no CPU fallback."""
    import subprocess
    import torch
    exe = _cli()
    r = subprocess.run([exe, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "compress" in r.stdout and "decompress" in r.stdout
    r = subprocess.run([exe, "compress", "--input", "nowhere"], capture_output=True, text=True)
    assert r.returncode == 1 and "Required argument" in r.stderr
    if torch.cuda.is_available():
        return
    src = tmp_path / "u.bin"
    np.zeros(81).tofile(src)
    r = subprocess.run([exe, "compress", "--datatype", "double", "--shape", "9x9", "--smoothness", "inf",
                        "--tolerance", "1e-3", "--input", str(src), "--output", str(tmp_path / "u.mgard")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "status 5" in r.stderr


@pytest.mark.gpu
def test_gpu_mgard_cli_round_trip_and_stream_identity(tmp_path):
    import subprocess
    import mgard_b200.cpu as mc
    exe = _cli()
    shape = (33, 20, 17)
    rng = np.random.default_rng(23)
    u = np.cumsum(rng.standard_normal(shape), axis=0)
    src, mid, dst = tmp_path / "u.bin", tmp_path / "u.mgard", tmp_path / "u.out"
    u.tofile(src)
    for lossless, kind in (("zlib", mc.CPU_HUFFMAN_ZLIB), ("zstd", mc.CPU_HUFFMAN_ZSTD)):
        r = subprocess.run([exe, "compress", "--datatype", "double", "--shape", "33x20x17", "--smoothness", "inf",
                            "--tolerance", "1e-2", "--input", str(src), "--output", str(mid), "--lossless", lossless],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        blob = mid.read_bytes()
        h = mo.Hierarchy(shape, np.float64)
        assert blob == mo.compress(h, u, math.inf, 1e-2, kind)  # the file the reference would write
        r = subprocess.run([exe, "decompress", "--input", str(mid), "--output", str(dst)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        back = np.fromfile(dst).reshape(shape)
        assert np.abs(back - u).max() <= 1e-2
        assert bits_equal(back, mc.decompress(blob))
    r = subprocess.run([exe, "compress", "--datatype", "double", "--shape", "33x20", "--smoothness", "inf",
                        "--tolerance", "1e-2", "--input", str(src), "--output", str(mid)], capture_output=True, text=True)
    assert r.returncode == 1 and "expected" in r.stderr


def test_cxx_cpu_api_mirror_compiles(tmp_path):
    """include/mgard_b200/compress.hpp compiles and links against the C ABI (without a
    GPU the program stops at the hierarchy constructor: no CPU fallback)."""
    import subprocess
    import torch
    exe = tmp_path / "cpu_api_roundtrip"
    for define in ([], ["-DMGARD_ZSTD"]):
        subprocess.check_call(["g++", "-std=c++17", *define, f"-I{ROOT}/include", f"{ROOT}/tests/cxx/cpu_api_roundtrip.cpp",
                               "-o", str(exe), f"-L{ROOT}/mgard_b200", "-lmgard_b200",
                               f"-Wl,-rpath,{ROOT}/mgard_b200", "-L/usr/local/cuda/lib64", "-lcudart"])
    if not torch.cuda.is_available():
        out = subprocess.run([str(exe)], capture_output=True, text=True)
        assert out.returncode != 0


def _golden_cases():
    g = np.load(os.path.join(ROOT, "tests", "golden", "cpu_convention.npz"), allow_pickle=False)
    for i in range(int(g["count"])):
        shape = tuple(int(x) for x in g[f"shape{i}"])
        dt = np.float64 if int(g[f"dtype{i}"]) == 1 else np.float32
        coords = None
        if int(g[f"explicit{i}"]):
            flat, coords, off = g[f"coords{i}"], [], 0
            for m in shape:
                coords.append(flat[off:off + m].astype(dt))
                off += m
        yield (shape, dt, coords, float(g[f"s{i}"]), float(g[f"tol{i}"]), g[f"u{i}"], g[f"coef{i}"],
               g[f"quanta{i}"], g[f"recomposed{i}"], g[f"zlib{i}"].tobytes(), g[f"huffzstd{i}"].tobytes())


def test_oracle_against_committed_reference_vectors():
    """tests/golden/cpu_convention.npz: outputs of the unmodified reference MGARD-CPU build
    (tests/golden/make_cpu_convention_golden.py) - pins the restatement where oracle/_ref
    is not available."""
    n = 0
    for shape, dt, coords, s, tol, u, coef, quanta, recomposed, zl, hz in _golden_cases():
        h = mo.Hierarchy(shape, dt, coords)
        assert bits_equal(mo.decompose(h, u), coef)
        assert np.array_equal(mo.quantize(h, s, tol, coef), quanta)
        assert bits_equal(mo.recompose(h, mo.dequantize(h, s, tol, quanta)), recomposed)
        assert mo.zlib_payload(quanta) == zl
        assert mo.huffman_payload(quanta)[3] is not None
        # the zstd frame depends on the libzstd version: compare what precedes it
        tree_bytes, hit_bits, miss_bytes, _ = mo.huffman_payload(quanta)
        assert np.frombuffer(hz[:24], dtype="<u8").tolist() == [tree_bytes, hit_bits, miss_bytes]
        n += 1
    assert n >= 8


@pytest.mark.gpu
def test_gpu_against_committed_reference_vectors():
    import torch
    import mgard_b200.cpu as mc
    for shape, dt, coords, s, tol, u, coef, quanta, recomposed, zl, hz in _golden_cases():
        H = mc.TensorMeshHierarchy(shape, coords, dt)
        du = torch.from_numpy(u).cuda()
        c = H.decompose(du)
        assert bits_equal(c.cpu().numpy(), coef)
        assert np.array_equal(H.quantize(c, s, tol).cpu().numpy(), quanta)
        blob = mc.compress(H, u, s, tol)
        assert blob.endswith(zl)
        assert bits_equal(mc.decompress(blob), recomposed)
        blob2 = mc.compress(H, u, s, tol, mc.CPU_HUFFMAN_ZSTD)
        assert bits_equal(mc.decompress(blob2), recomposed)
