"""Generates tests/golden/cpu_convention.npz from the UNMODIFIED reference MGARD-CPU build
(oracle/_ref/libmgard_cpu_ref.so and ..._zstd.so; run `make -C oracle` in the build container
first).  Inputs are seeded; outputs are the reference's shuffled multilevel coefficients, its
int64 quanta and its two lossless payloads."""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import ref_cpu  # noqa: E402

CASES = [
    ((17,), np.float32, False, math.inf, 1e-3),
    ((20,), np.float64, True, 0.0, 1e-2),
    ((10, 7), np.float32, True, 1.0, 1e-2),
    ((33, 17), np.float64, False, math.inf, 1e-4),
    ((9, 12, 13), np.float32, True, math.inf, 1e-3),
    ((17, 9, 11), np.float64, False, -0.5, 1e-2),
    ((5, 6, 7, 9), np.float64, True, math.inf, 1e-3),
    ((12, 1, 9), np.float32, False, 0.0, 1e-2),
]


def main():
    rng = np.random.default_rng(2024)
    out = {"count": np.int64(len(CASES))}
    for i, (shape, dt, explicit, s, tol) in enumerate(CASES):
        coords = None
        if explicit:
            coords = []
            for n in shape:
                x = np.concatenate([[0.0], np.cumsum(rng.uniform(1, 2, n - 1))]) if n > 1 else np.zeros(1)
                coords.append((x / max(x[-1], 1)).astype(dt))
        u = np.cumsum(rng.standard_normal(shape), axis=0).astype(dt)
        c = ref_cpu.decompose(u, coords)
        q = ref_cpu.quantize(c, shape, s, tol, coords)
        out[f"shape{i}"] = np.array(shape, dtype=np.int64)
        out[f"dtype{i}"] = np.int64(1 if dt is np.float64 else 0)
        out[f"coords{i}"] = np.concatenate(coords).astype(np.float64) if explicit else np.zeros(0)
        out[f"explicit{i}"] = np.int64(1 if explicit else 0)
        out[f"s{i}"] = np.float64(s)
        out[f"tol{i}"] = np.float64(tol)
        out[f"u{i}"] = u
        out[f"coef{i}"] = c
        out[f"quanta{i}"] = q
        out[f"recomposed{i}"] = ref_cpu.recompose(ref_cpu.dequantize(q, shape, dt, s, tol, coords), shape, coords)
        out[f"zlib{i}"] = ref_cpu.zlib_compress(q)
        out[f"huffzstd{i}"] = ref_cpu.huffman_zstd_compress(q)
    np.savez_compressed(os.path.join(HERE, "cpu_convention.npz"), **out)


if __name__ == "__main__":
    main()
