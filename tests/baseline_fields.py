"""Synthetic inputs of the five BASELINE.json configurations, exactly as SURVEY.md
section 8(d) defines them (closed forms + splitmix64 noise by row-major linear index),
generated with numpy in plane blocks so that the 2049^3 slabs fit in host memory.
Shared by the GPU parity tests, tests/golden/make_baseline_digests.py and the scripts."""
import hashlib

import numpy as np

INF = float("inf")


def splitmix_noise(lo, count, seed):
    """xi(i) = (splitmix64(seed + i) >> 11) * 2^-52 - 1 in [-1, 1), i = lo .. lo+count-1."""
    i = np.arange(count, dtype=np.uint64) + np.uint64(lo)
    with np.errstate(over="ignore"):
        z = i + np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -52 - 1.0


def c1():
    """129^3 fp64, ABS 1e-4, s = inf."""
    n = 129
    x = np.arange(n, dtype=np.float64) / (n - 1)
    X0, X1, X2 = np.meshgrid(x, x, x, indexing="ij")
    return np.sin(2 * np.pi * X0) * np.cos(3 * np.pi * X1) + 0.5 * np.sin(5 * np.pi * X2) + 0.25 * X0 * X1


def c2_like(full_shape, plane0=0, planes=None, seed=2049, block=16, crop=None):
    """C2 / C5 field on planes [plane0, plane0 + planes) of a domain `full_shape`
    (coordinates and the noise index are those of the full domain); crop = (m1, m2)
    keeps the first m1 x m2 nodes of every plane.  Plane blocks are independent; they
    are filled by a few threads (numpy releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    import os
    n0, n1, n2 = full_shape
    m1, m2 = crop or (n1, n2)
    planes = n0 - plane0 if planes is None else planes
    out = np.empty((planes, m1, m2), dtype=np.float32)
    x1 = (np.arange(m1, dtype=np.float64) / (n1 - 1)).reshape(1, -1, 1)
    x2 = (np.arange(m2, dtype=np.float64) / (n2 - 1)).reshape(1, 1, -1)
    c1s2 = np.cos(4 * np.pi * x1) * np.sin(2 * np.pi * x2)

    def fill(a):
        b = min(planes, a + block)
        x0 = ((np.arange(a, b, dtype=np.float64) + plane0) / (n0 - 1)).reshape(-1, 1, 1)
        u = np.sin(6 * np.pi * x0) * c1s2 + 0.3 * np.sin(40 * np.pi * x0 * x1)
        if crop is None:
            xi = splitmix_noise((plane0 + a) * n1 * n2, (b - a) * n1 * n2, seed).reshape(b - a, n1, n2)
        else:
            xi = np.empty((b - a, m1, m2), dtype=np.float64)
            for p in range(a, b):
                for j in range(m1):
                    xi[p - a, j] = splitmix_noise(((plane0 + p) * n1 + j) * n2, m2, seed)
        out[a:b] = (u + 1e-3 * xi).astype(np.float32)

    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        list(ex.map(fill, range(0, planes, block)))
    return out


def c2():
    """513^3 fp32, REL 1e-3, s = inf."""
    return c2_like((513, 513, 513))


def c3():
    """1000^2 fp32 on non-uniform coordinates; returns (u, coords)."""
    n = 1000
    cs = []
    for k in (7, 11):
        i = np.arange(n - 1)
        h = 1 + 0.5 * np.sin(2 * np.pi * k * i / 999)
        x = np.concatenate([[0.0], np.cumsum(h)])
        cs.append((x / x[-1]).astype(np.float32))
    X0, X1 = np.meshgrid(cs[0].astype(np.float64), cs[1].astype(np.float64), indexing="ij")
    u = (np.exp(-8 * ((X0 - .5) ** 2 + (X1 - .4) ** 2)) + 0.1 * np.sin(30 * X0)).astype(np.float32)
    return u, cs


def c4(n1=16395, seed=16395):
    """8 x n1 x 39 x 39 fp64 (n1 = 16395: the full XGC-shaped config; smaller: a crop
    along dim 1 with the full domain's coordinates and noise indices)."""
    full = (8, 16395, 39, 39)
    i0 = np.arange(8, dtype=np.float64).reshape(-1, 1, 1, 1)
    x1 = (np.arange(n1, dtype=np.float64) / (full[1] - 1)).reshape(1, -1, 1, 1)
    x2 = (np.arange(39, dtype=np.float64) / 38).reshape(1, 1, -1, 1)
    x3 = (np.arange(39, dtype=np.float64) / 38).reshape(1, 1, 1, -1)
    g = sum((1.0 / k) * np.sin(2 * np.pi * (2 * k + 1) * x1) for k in range(1, 6))
    e = np.exp(-((x2 - .5) ** 2 + (x3 - .5) ** 2) / 0.08)
    out = np.empty((8, n1, 39, 39), dtype=np.float64)
    inner = 39 * 39
    for a in range(8):
        lo = a * full[1] * inner  # rows j < n1 of plane a are contiguous in the full index
        xi = splitmix_noise(lo, n1 * inner, seed).reshape(n1, 39, 39)
        out[a] = ((1 + 0.1 * i0[a]) * g[0] * e[0]) + 1e-4 * xi
    return out


def c5_slab(k=0):
    """Sub-domain k of the 2049^3 fp32 domain MaxDim-decomposed in 257-plane slabs."""
    planes = 257 if k < 7 else 250
    return c2_like((2049, 2049, 2049), 257 * k, planes)


def sha(a):
    a = np.ascontiguousarray(a)
    h = hashlib.sha256()
    mv = memoryview(a).cast("B")
    step = 1 << 28
    for o in range(0, len(mv), step):
        h.update(mv[o:o + step])
    return h.hexdigest()


def payload_digests(parsed):
    """Digests of a parsed Huffman block (oracle huffman_parse); outliers as a set."""
    d = {k: int(parsed[k]) for k in ("n", "dict_size", "chunk_size", "size")}
    for k in ("bits", "word_offset", "first", "entry", "keys", "ddata"):
        d[k] = sha(np.asarray(parsed[k]))
    oi = np.asarray(parsed["oidx"]).astype(np.uint64)
    o = np.argsort(oi, kind="stable")
    d["outliers"] = int(oi.size)
    d["oidx"] = sha(oi[o])
    d["oval"] = sha(np.asarray(parsed["oval"]).astype(np.int64)[o])
    return d
