"""CPU: the C-ABI library loads, exports every symbol include/mgard_b200.h
declares, fails loudly without a GPU, and its host-side format code (header
writer / parser) matches the protobuf-generated fixtures."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from mgard_b200 import _lib
import mgard_b200 as mg

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mgard_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mgb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), n
        assert n in _lib.SIGNATURES, f"{n} missing from the ctypes signature table"
    assert b"sm_100a" in L.mgb_version()


def test_config_defaults_match_reference():
    c = _lib.MgbConfig()
    _lib.lib().mgb_config_default(C.byref(c))
    # src/mgard-x/Config/Config.cpp:14-43
    assert (c.huff_dict_size, c.huff_block_size, c.normalize_coordinates) == (8192, 20480, 1)


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu():
    u = np.zeros((5, 5, 5), dtype=np.float32)
    with pytest.raises(mg.MgardError) as e:
        mg.compress(u, 1e-3, float("inf"), mg.error_bound_type.REL)
    assert e.value.status == _lib.BACKEND_NOT_AVAILABLE
    h = C.c_void_p(0)
    rc = _lib.lib().mgb_plan_create(3, (C.c_uint64 * 3)(5, 5, 5), 0, None, None, C.byref(h))
    assert rc == _lib.BACKEND_NOT_AVAILABLE


def test_argument_errors_mirror_reference_status_codes():
    L = _lib.lib()
    h = C.c_void_p(0)
    # NotSupportHigherNumberOfDimensionsFailure / NotSupportDataTypeFailure (Types.h:56-63)
    assert L.mgb_plan_create(6, (C.c_uint64 * 6)(5, 5, 5, 5, 5, 5), 0, None, None, C.byref(h)) == 3
    assert L.mgb_plan_create(3, (C.c_uint64 * 3)(5, 5, 5), 7, None, None, C.byref(h)) == 4
    # Hierarchy.hpp:748-756: every dimension must be >= 3
    assert L.mgb_plan_create(2, (C.c_uint64 * 2)(2, 9), 0, None, None, C.byref(h)) == _lib.BAD_ARGUMENT


def test_header_writer_and_parser_against_protobuf_fixtures():
    L = _lib.lib()
    z = np.load(os.path.join(HERE, "golden", "headers.npz"))
    import zlib, struct
    i = 0
    while f"hdr{i}" in z:
        shape = [int(x) for x in z[f"shape{i}"]]
        dt, eb, tol, s, norm, dec, dd, ds = z[f"meta{i}"]
        npdt = np.float32 if dt == 0 else np.float64
        cfg = _lib.MgbConfig()
        L.mgb_config_default(C.byref(cfg))
        if dec:
            cfg.domain_decomposition_dim, cfg.domain_decomposition_size = int(dd), int(ds)
        else:
            cfg.domain_decomposition_size = 1 << 40
        keep, carr = [], None
        if f"coords{i}_0" in z:
            carr = (C.c_void_p * len(shape))()
            for d in range(len(shape)):
                cc = np.ascontiguousarray(z[f"coords{i}_{d}"], dtype=npdt)
                keep.append(cc)
                carr[d] = cc.ctypes.data
        out = np.zeros(1 << 16, dtype=np.uint8)
        sz = C.c_uint64(0)
        rc = L.mgb_write_header(len(shape), int(dt), (C.c_uint64 * len(shape))(*shape), tol, s,
                                int(eb), norm, carr, C.byref(cfg), out.ctypes.data, out.size,
                                C.byref(sz))
        assert rc == 0
        hdr = z[f"hdr{i}"].tobytes()
        want = b"MGARD" + struct.pack("<Q", len(hdr)) + struct.pack("<I", zlib.crc32(hdr)) + hdr
        got = out[:sz.value].tobytes()
        assert got == want, i
        # parser round trip
        info = mg.peek_header(np.frombuffer(got + b"\0" * 16, dtype=np.uint8))
        assert list(info["shape"]) == shape and int(info["dtype"]) == int(dt)
        assert int(info["mode"]) == int(eb) and info["tol"] == tol
        assert info["s"] == s or (np.isinf(s) and np.isinf(info["s"]))
        assert info["header_bytes"] == len(want)
        if eb == 0:
            assert info["norm"] == norm
        i += 1
    assert i == 5


def test_parser_rejects_corruption():
    L = _lib.lib()
    z = np.load(os.path.join(HERE, "golden", "headers.npz"))
    import zlib, struct
    hdr = z["hdr0"].tobytes()
    good = b"MGARD" + struct.pack("<Q", len(hdr)) + struct.pack("<I", zlib.crc32(hdr)) + hdr
    bad_magic = b"MGARX" + good[5:]
    bad_crc = good[:20] + bytes([good[20] ^ 1]) + good[21:]
    for b in (bad_magic, bad_crc, good[:30]):
        with pytest.raises(mg.MgardError) as e:
            mg.peek_header(np.frombuffer(b, dtype=np.uint8))
        assert e.value.status in (_lib.BAD_STREAM, _lib.BAD_ARGUMENT)


def test_cli_builds_and_reports_missing_backend(tmp_path):
    """mgard-x-b200 (the reference CLI's options on this engine): usage text without
    arguments; without a CUDA device a compression request fails loudly with
    BackendNotAvailableFailure (5), never a CPU fallback."""
    import subprocess
    exe = os.path.join(ROOT, "mgard_b200", "mgard-x-b200")
    if not os.path.exists(exe):
        pytest.skip("CLI not built")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "--compress" in r.stdout and "--decompress" in r.stdout
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    src = tmp_path / "u.bin"
    np.sin(np.arange(9 * 9 * 9) * 0.1).astype(np.float32).tofile(src)
    r = subprocess.run([exe, "-z", "-i", str(src), "-o", str(tmp_path / "u.mgard"), "-dt", "s", "-dim", "3",
                        "9", "9", "9", "-em", "rel", "-e", "1e-3", "-s", "inf"], capture_output=True, text=True)
    assert r.returncode == 1 and "status 5" in r.stderr


def test_header_records_the_second_stage_lossless():
    """Encoding.compressor (src/mgard.proto:139-145): X_HUFFMAN = 3 by default,
    X_HUFFMAN_ZSTD = 5 for lossless_type::Huffman_Zstd; nothing else changes."""
    L = _lib.lib()
    outs = []
    for lossless in (0, 2):
        cfg = _lib.MgbConfig()
        L.mgb_config_default(C.byref(cfg))
        cfg.domain_decomposition_size = 1 << 40
        cfg.lossless = lossless
        out = np.zeros(1 << 12, dtype=np.uint8)
        sz = C.c_uint64(0)
        shape = (C.c_uint64 * 3)(33, 34, 35)
        assert L.mgb_write_header(3, 0, shape, 1e-3, float("inf"), 0, 1.0, None, C.byref(cfg),
                                  out.ctypes.data, out.size, C.byref(sz)) == 0
        outs.append(out[:sz.value].copy())
    a, b = outs
    assert a.size == b.size
    diff = np.nonzero(a[17:] != b[17:])[0]  # after magic | size | crc
    assert diff.size == 1 and a[17 + diff[0]] == 3 and b[17 + diff[0]] == 5
    assert mg.peek_header(a)["header_bytes"] == mg.peek_header(b)["header_bytes"] == a.size


def test_header_records_reorder_and_single_dimension():
    """Config::reorder -> Encoding.preprocessor = SHUFFLE, decomposition_type::SingleDim ->
    FunctionDecomposition.hierarchy = ONE_DIM_AT_A_TIME_WITH_GHOST_NODES
    (Metadata.cpp:360-370,408-412): header bytes equal to the oracle's proto3 encoding."""
    import mgardx_oracle as mo
    L = _lib.lib()
    shape = (33, 34, 35)
    for reorder, decomposition in ((0, 0), (1, 0), (0, 1), (1, 1)):
        cfg = _lib.MgbConfig()
        L.mgb_config_default(C.byref(cfg))
        cfg.domain_decomposition_size = 1 << 40
        cfg.reorder = reorder
        cfg.decomposition = decomposition
        out = np.zeros(1 << 12, dtype=np.uint8)
        sz = C.c_uint64(0)
        cshape = (C.c_uint64 * 3)(*shape)
        assert L.mgb_write_header(3, 0, cshape, 1e-3, float("inf"), 0, 2.5, None, C.byref(cfg),
                                  out.ctypes.data, out.size, C.byref(sz)) == 0
        got = out[:sz.value].tobytes()
        hdr = mo.encode_header(shape, np.float32, mo.REL, 1e-3, float("inf"), np.float32(2.5), reorder=reorder)
        if decomposition:  # field 8 (FunctionDecomposition), field 2 (hierarchy): 1 -> 2
            assert hdr.count(b"\x42\x02\x10\x01") == 1
            hdr = hdr.replace(b"\x42\x02\x10\x01", b"\x42\x02\x10\x02")
        assert got == mo.encode_preamble(hdr)
        assert mg.peek_header(np.frombuffer(got, dtype=np.uint8))["header_bytes"] == len(got)


def test_tuning_knobs():
    """mgb_tune (include/mgard_b200.h): the three knobs are accepted without a device and
    restore to their defaults; an unknown key is an argument error."""
    import mgard_b200 as mg
    for key, default in ((mg.TUNE_SERIAL_MIN_CHUNKS, 16384), (mg.TUNE_RING_DECODER, 1), (mg.TUNE_SUB_ENCODER, 1)):
        mg.tune(key, 0)
        mg.tune(key, default)
    from mgard_b200 import _lib
    assert _lib.lib().mgb_tune(99, 0) != 0
