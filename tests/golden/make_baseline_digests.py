"""TEST INFRASTRUCTURE.  Runs the UNMODIFIED reference (oracle/_ref/libmgardx_ref.so,
MGARD-X SERIAL adapter compiled in place from /root/reference by oracle/Makefile) on the
BASELINE.json configurations at FULL size and records SHA-256 digests of everything it
produces: decomposed coefficients, quantized symbols, the Huffman block field by field
(outlier list as a set), the reconstruction, plus norm, sizes and timings.

    python tests/golden/make_baseline_digests.py C2 C3 C4 C4crop C5slab

writes / updates tests/golden/baseline_digests.json.  The GPU tests
(tests/test_gpu_baseline_configs.py) regenerate the same inputs (tests/baseline_fields.py,
digest checked) and compare the CUDA path's digests with these; where the input digest
does not reproduce on the test machine they run the reference live instead."""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import baseline_fields as bf  # noqa: E402
import mgardx_oracle as mo  # noqa: E402
import ref_x  # noqa: E402

OUT = os.path.join(HERE, "baseline_digests.json")

CONFIGS = {
    # name: (generator, ebtype, tol, s)
    "C1": (lambda: (bf.c1(), None), ref_x.ABS, 1e-4, bf.INF),
    "C2": (lambda: (bf.c2(), None), ref_x.REL, 1e-3, bf.INF),
    "C3abs": (bf.c3, ref_x.ABS, 1e-2, 0.0),
    "C3rel": (bf.c3, ref_x.REL, 1e-2, 0.0),
    "C4": (lambda: (bf.c4(), None), ref_x.REL, 1e-3, 0.0),
    "C4crop": (lambda: (bf.c4(2051), None), ref_x.REL, 1e-3, 0.0),
    # one MaxDim sub-domain of C5, compressed the way the high-level API does it: ABS with
    # the local tolerance tol * global norm (CompressionHighLevel.hpp:128-139); the global
    # max |u| of the C2 formula on 2049^3 is not needed exactly for a parity record, 1.3 is used
    "C5slab": (lambda: (bf.c5_slab(0), None), ref_x.ABS, float(np.float32(1e-3) * np.float32(1.3)), bf.INF),
}


def oracle_block_digests(q, cb, chunk_size, ref_parsed, group=256):
    """Digests of the Huffman block the oracle's encoder (mgardx_oracle, restating
    Deflate.hpp:21-77 / Huffman.hpp:130-262) writes for symbols `q` with codebook `cb`,
    computed group of chunks by group of chunks so that 10^9 symbols fit in memory.
    Fields that do not depend on the codebook are taken from the reference's block."""
    import hashlib
    sym = q.ravel()
    n = sym.size
    nchunk = (n - 1) // chunk_size + 1
    h_bits, h_words = hashlib.sha256(), hashlib.sha256()
    nwords = np.zeros(nchunk, dtype=np.uint64)
    for c0 in range(0, nchunk, group):
        c1 = min(nchunk, c0 + group)
        bits, words = mo.huffman_encode_chunks(sym[c0 * chunk_size:min(n, c1 * chunk_size)],
                                               cb["codebook"], chunk_size)
        h_bits.update(bits.astype("<u8").tobytes())
        for k, w in enumerate(words):
            nwords[c0 + k] = w.size
            h_words.update(w.astype("<u8").tobytes())
    entry = np.zeros(nchunk, dtype=np.uint64)
    entry[1:] = np.cumsum(nwords)[:-1]
    d = dict(ref_parsed)
    d.update(bits=h_bits.hexdigest(), word_offset=bf.sha(entry), ddata=h_words.hexdigest(),
             first=bf.sha(cb["first"]), entry=bf.sha(cb["entry"]), keys=bf.sha(cb["keys"]))
    total_words = int(nwords.sum())
    # size: the reference's minus its words plus these
    d["size"] = int(ref_parsed["size"]) + 8 * (total_words - int(ref_parsed["nwords"]))
    d.pop("nwords", None)
    return d


def run(name):
    gen, eb, tol, s = CONFIGS[name]
    u, coords = gen()
    rec = {"shape": list(u.shape), "dtype": u.dtype.name, "ebtype": int(eb), "tol": tol,
           "s": "inf" if np.isinf(s) else s, "input": bf.sha(u)}
    t0 = time.perf_counter()
    r = ref_x.compress(u, eb, tol, s, coords)
    t1 = time.perf_counter()
    rec["reference_compress_s"] = t1 - t0
    rec["norm"] = float(r["norm"])
    rec["norm_hex"] = float(r["norm"]).hex()
    rec["l_target"] = int(r["l_target"])
    rec["decomposed"] = bf.sha(r["decomposed"])
    q = r.pop("quantized")
    r.pop("decomposed")
    assert q.min() >= 0 and q.max() < 65536
    rec["symbols"] = bf.sha(q.astype(np.uint16))
    rec["outlier_count"] = int(r["outlier_count"])
    parsed = mo.huffman_parse(r["payload"])
    ref_dig = bf.payload_digests(parsed)
    rec["reference_payload"] = ref_dig
    rec["reference_payload_bytes"] = int(r["payload"].size)
    # The reference's code lengths depend on a word it reads one past the end of its
    # frequency array (GenerateCL.hpp:252-257): its block is
    # one of the oracle's variants.  The engine defines the word as 0; "payload" holds
    # the digests of THAT variant - the reference's own block when it took it, else the
    # oracle encoder's output for the reference's symbols.
    freq = np.bincount(q.ravel(), minlength=int(parsed["dict_size"])).astype(np.uint32)
    variants = {str(o): mo.get_codebook(freq, o) for o in (0, 0xFFFFFFFF)}

    def same(cb):
        return (np.array_equal(cb["first"], parsed["first"]) and np.array_equal(cb["entry"], parsed["entry"])
                and np.array_equal(cb["keys"], parsed["keys"]))

    took = [k for k, cb in variants.items() if same(cb)]
    rec["reference_oob_variant"] = took
    assert took, "the reference's codebook is neither variant of the oracle's"
    if "0" in took:
        rec["payload"] = ref_dig
        rec["payload_bytes"] = int(r["payload"].size)
    else:
        base = dict(ref_dig)
        base["nwords"] = int(np.asarray(parsed["ddata"]).size)
        rec["payload"] = oracle_block_digests(q, variants["0"], int(parsed["chunk_size"]), base)
        rec["payload_bytes"] = rec["payload"]["size"]
        if q.size <= (1 << 28):
            # the oracle's encoder with the reference's variant reproduces the reference's block
            chk = oracle_block_digests(q, variants[took[0]], int(parsed["chunk_size"]), base)
            assert chk == ref_dig, "oracle encoder does not reproduce the reference's block"
            rec["oracle_encoder_pinned_at_this_size"] = True
    del q
    rec["ratio"] = u.nbytes / rec["payload_bytes"]
    rec["reference_ratio"] = u.nbytes / r["payload"].size
    t1 = time.perf_counter()
    back = ref_x.decompress(r["payload"], u.shape, u.dtype, eb, tol, s, r["norm"], coords)
    rec["reference_decompress_s"] = time.perf_counter() - t1
    rec["decompressed"] = bf.sha(back)
    if np.isinf(s):
        rec["max_abs_error"] = float(np.abs(back.astype(np.float64) - u).max())
    else:
        d = back.astype(np.float64) - u
        rec["rms_error"] = float(np.sqrt((d * d).sum() / u.size))
    return rec


if __name__ == "__main__":
    names = sys.argv[1:] or list(CONFIGS)
    for nm in names:
        rec = run(nm)
        print(nm, json.dumps(rec), flush=True)
        allrec = json.load(open(OUT)) if os.path.exists(OUT) else {}
        allrec[nm] = rec
        json.dump(allrec, open(OUT, "w"), indent=1, sort_keys=True)
