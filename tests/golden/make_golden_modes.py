"""Generates tests/golden/x_modes.npz from the UNMODIFIED reference MGARD-X build
(oracle/_ref/libmgardx_ref.so): Config::reorder = 1 and decomposition_type::SingleDim.
Run in the build container only; the fixture is committed."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import ref_x  # noqa: E402
from make_golden import field, nonuniform  # noqa: E402

CASES = [
    # shape, dtype, nonuniform?, tol, s, reorder, single_dim
    ((17,), np.float32, False, 1e-3, np.inf, 1, 0),
    ((10, 7), np.float64, True, 1e-3, 0.0, 1, 0),
    ((17, 19, 21), np.float32, False, 1e-4, np.inf, 1, 0),
    ((6,), np.float64, False, 1e-3, np.inf, 0, 1),
    ((9, 12), np.float32, True, 1e-2, np.inf, 0, 1),
    ((12, 13, 14), np.float64, False, 1e-3, 0.5, 0, 1),
    ((5, 6, 9), np.float32, False, 1e-3, np.inf, 1, 1),
]


def main():
    out = {"count": np.int64(len(CASES))}
    for i, (shape, dt, nonuni, tol, s, reorder, sd) in enumerate(CASES):
        u = field(shape, dt, 40 + i)
        coords = [nonuniform(n, 3 + 2 * k, dt) for k, n in enumerate(shape)] if nonuni else None
        r = ref_x.compress(u, ref_x.REL, tol, s, coords, reorder=reorder, decomposition=sd)
        back = ref_x.decompress(r["payload"], shape, dt, ref_x.REL, tol, s, r["norm"], coords,
                                reorder=reorder, decomposition=sd)
        out[f"shape{i}"] = np.array(shape, dtype=np.int64)
        out[f"dtype{i}"] = np.int64(1 if dt is np.float64 else 0)
        out[f"explicit{i}"] = np.int64(1 if nonuni else 0)
        out[f"coords{i}"] = np.concatenate(coords).astype(np.float64) if nonuni else np.zeros(0)
        out[f"tol{i}"], out[f"s{i}"] = np.float64(tol), np.float64(s)
        out[f"reorder{i}"], out[f"single{i}"] = np.int64(reorder), np.int64(sd)
        out[f"u{i}"], out[f"norm{i}"] = u, np.float64(r["norm"])
        out[f"decomposed{i}"] = r["decomposed"]
        out[f"quantized{i}"] = r["quantized"].ravel()
        out[f"decompressed{i}"] = back
    np.savez_compressed(os.path.join(HERE, "x_modes.npz"), **out)


if __name__ == "__main__":
    main()
