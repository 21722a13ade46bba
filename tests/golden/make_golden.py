"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref built
from /root/reference by oracle/Makefile) and, for the header bytes, from Python
protobuf driven by the reference's own src/mgard.proto.  Run in the build
container only (needs /root/reference); the fixtures are committed."""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_x  # noqa: E402


def field(shape, dtype, seed):
    rng = np.random.default_rng(seed)
    g = np.meshgrid(*[np.linspace(0, 1, n) for n in shape], indexing="ij")
    u = sum(np.sin((3 + 2 * i) * x + i) for i, x in enumerate(g)) + 0.05 * rng.standard_normal(shape)
    return u.astype(dtype)


def nonuniform(n, k, dtype):
    h = 1 + 0.5 * np.sin(2 * np.pi * k * np.arange(n - 1) / (n - 1))
    x = np.concatenate([[0], np.cumsum(h)])
    return (x / x[-1]).astype(dtype)


CASES = [
    # name, shape, dtype, nonuniform?, ebtype, tol, s
    ("d1_f32_17", (17,), np.float32, False, ref_x.REL, 1e-3, np.inf),
    ("d1_f64_100", (100,), np.float64, True, ref_x.ABS, 1e-4, 0.0),
    ("d2_f32_10x7", (10, 7), np.float32, False, ref_x.REL, 1e-2, 0.0),
    ("d2_f64_33x20", (33, 20), np.float64, True, ref_x.REL, 1e-3, np.inf),
    ("d3_f32_5x6x9", (5, 6, 9), np.float32, False, ref_x.REL, 1e-3, np.inf),
    ("d3_f32_17x19x21", (17, 19, 21), np.float32, True, ref_x.ABS, 1e-2, 0.5),
    ("d3_f64_12x13x14", (12, 13, 14), np.float64, False, ref_x.REL, 1e-4, -1.0),
    ("d4_f64_4x17x5x6", (4, 17, 5, 6), np.float64, False, ref_x.REL, 1e-3, 0.0),
    ("d5_f32_5x5x6x7x5", (5, 5, 6, 7, 5), np.float32, False, ref_x.REL, 1e-3, np.inf),
]


def main():
    for seed, (name, shape, dt, nonuni, eb, tol, s) in enumerate(CASES):
        u = field(shape, dt, seed)
        coords = [nonuniform(n, 3 + 2 * i, dt) for i, n in enumerate(shape)] if nonuni else None
        r = ref_x.compress(u, eb, tol, s, coords)
        back = ref_x.decompress(r["payload"], shape, dt, eb, tol, s, r["norm"], coords)
        tb = ref_x.tables(shape, dt, coords)
        L = len(tb) - 1
        out = dict(u=u, ebtype=eb, tol=tol, s=s, norm=r["norm"], decomposed=r["decomposed"],
                   quantized=r["quantized"], payload=r["payload"], decompressed=back,
                   recomposed=ref_x.recompose(r["decomposed"], coords), l_target=L)
        if coords is not None:
            for d, c in enumerate(coords):
                out[f"coords{d}"] = c
        for d in range(len(shape)):
            for k in ("dist", "ratio", "am", "bm"):
                out[f"tab_{k}_L_{d}"] = tb[L][d][k]
                out[f"tab_{k}_0_{d}"] = tb[0][d][k]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "payload", r["payload"].size, "outliers", r["outlier_count"])
    # codebooks
    rng = np.random.default_rng(99)
    cbs = {}
    for i, d in enumerate([16, 64, 1024, 8192, 8192]):
        if i % 2 == 0:
            fr = rng.integers(0, 50, d)
        else:
            fr = np.zeros(d, dtype=np.int64)
            k = min(d, 150)
            fr[d // 2 - k // 2: d // 2 - k // 2 + k] = (1e6 * np.exp(-0.5 * ((np.arange(k) - k / 2) / (k / 8)) ** 2)).astype(np.int64) + (rng.random(k) < 0.5)
        c = ref_x.codebook(fr)
        cbs[f"freq{i}"] = fr.astype(np.uint32)
        for k2 in ("codebook", "first", "entry", "keys"):
            cbs[f"{k2}{i}"] = c[k2]
    np.savez_compressed(os.path.join(HERE, "codebooks.npz"), **cbs)
    # header bytes from Python protobuf + the reference's mgard.proto
    import torch
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    protoc = os.path.join(os.path.dirname(torch.__file__), "bin", "protoc")
    d = tempfile.mkdtemp()
    subprocess.check_call([protoc, "--proto_path=/root/reference/src",
                           f"--descriptor_set_out={d}/mgard.desc", "/root/reference/src/mgard.proto"])
    fds = descriptor_pb2.FileDescriptorSet()
    fds.ParseFromString(open(f"{d}/mgard.desc", "rb").read())
    pool = descriptor_pool.DescriptorPool()
    for f in fds.file:
        pool.Add(f)
    Header = message_factory.GetMessageClass(pool.FindMessageTypeByName("mgard.pb.Header"))

    def header(shape, dtype, eb, tol, s, norm, coords, dec, dd, ds):
        # fields exactly as MetadataBase::Serialize sets them (Metadata.cpp:249-439)
        h = Header()
        h.mgard_version.major_, h.mgard_version.minor_, h.mgard_version.patch_ = 1, 0, 0
        h.file_format_version.SetInParent()
        h.domain.topology = 0
        h.domain.cartesian_grid_topology.dimension = len(shape)
        h.domain.cartesian_grid_topology.shape.extend(shape)
        h.domain.geometry = 0 if coords is None else 1
        if coords is not None:
            for c in coords:
                h.domain.explicit_cube_geometry.coordinates.extend([float(x) for x in c])
        h.dataset.type = 0 if dtype == np.float32 else 1
        h.dataset.dimension = 1
        if eb == ref_x.ABS:
            h.error_control.mode = 0
        else:
            h.error_control.mode = 1
            h.error_control.norm_of_original_data = norm
        h.error_control.norm = 0 if np.isinf(s) else 1
        h.error_control.s = s
        h.error_control.tolerance = tol
        h.domain_decomposition.method = 1 if dec else 0
        h.domain_decomposition.decomposition_dimension = dd if dec else 0
        h.domain_decomposition.decomposition_size = ds if dec else shape[0]
        h.function_decomposition.transform = 0
        h.function_decomposition.hierarchy = 1
        h.function_decomposition.L_target = 0
        h.quantization.method, h.quantization.bin_widths, h.quantization.type = 1, 0, 3
        h.quantization.big_endian = False
        h.bitplane_encoding.method = 0
        h.encoding.preprocessor, h.encoding.compressor = 0, 3
        h.encoding.huffman_dictionary_size, h.encoding.huffman_block_size = 8192, 20480
        h.device.backend = 3
        return h.SerializeToString()

    hdrs = {}
    hc = [((513, 513, 513), np.float32, ref_x.REL, 1e-3, np.inf, 1.2345, None, False, 0, 0),
          ((40, 30), np.float32, ref_x.ABS, 1e-2, 0.0, 1.0,
           [nonuniform(40, 3, np.float32), nonuniform(30, 5, np.float32)], False, 0, 0),
          ((2049, 2049, 2049), np.float32, ref_x.REL, 1e-3, np.inf, 2.5, None, True, 0, 257),
          ((8, 16395, 39, 39), np.float64, ref_x.REL, 1e-3, 0.0, 0.77, None, False, 0, 0),
          ((5,), np.float64, ref_x.ABS, 0.5, -1.5, 1.0, None, False, 0, 0)]
    for i, (shape, dt, eb, tol, s, norm, coords, dec, dd, ds) in enumerate(hc):
        hdrs[f"hdr{i}"] = np.frombuffer(header(shape, dt, eb, tol, s, norm, coords, dec, dd, ds), dtype=np.uint8)
        hdrs[f"shape{i}"] = np.array(shape)
        hdrs[f"meta{i}"] = np.array([0 if dt == np.float32 else 1, eb, tol, s, norm, int(dec), dd, ds], dtype=np.float64)
        if coords is not None:
            for d2, c in enumerate(coords):
                hdrs[f"coords{i}_{d2}"] = c
    np.savez_compressed(os.path.join(HERE, "headers.npz"), **hdrs)


if __name__ == "__main__":
    main()
