"""Generates tests/golden/cpu_headers.npz (run in the build container only: it
reads the reference's src/mgard.proto).

The bytes are what the reference's `header.SerializeToArray` produces for the
header `mgard::compress` fills in (reference include/compress.tpp:41-55,
src/format.cpp:102-140, include/TensorMeshHierarchy.tpp:293-348): python protobuf
is driven field by field in the same way, and proto3 serialisation is canonical.
"""
import os
import subprocess
import tempfile

import numpy as np

PROTO_DIR = "/root/reference/src"
PROTOC = "/opt/prime-rl/.venv/lib/python3.12/site-packages/torch/bin/protoc"


def header_class():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    with tempfile.TemporaryDirectory() as tmp:
        desc = os.path.join(tmp, "mgard.desc")
        subprocess.check_call([PROTOC, f"--proto_path={PROTO_DIR}", f"--descriptor_set_out={desc}",
                               os.path.join(PROTO_DIR, "mgard.proto")])
        fds = descriptor_pb2.FileDescriptorSet()
        fds.ParseFromString(open(desc, "rb").read())
    pool = descriptor_pool.DescriptorPool()
    for f in fds.file:
        pool.Add(f)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("mgard.pb.Header"))


def reference_header(Header, shape, is_double, coords, s, tol):
    m = Header()
    m.mgard_version.major_, m.mgard_version.minor_, m.mgard_version.patch_ = 1, 6, 0
    m.file_format_version.major_, m.file_format_version.minor_, m.file_format_version.patch_ = 1, 0, 0
    m.function_decomposition.transform = 0
    m.quantization.method, m.quantization.bin_widths = 1, 0
    m.quantization.type, m.quantization.big_endian = 3, False
    m.encoding.preprocessor, m.encoding.compressor = 1, 1
    m.device.backend = 0
    m.domain.topology = 0
    m.domain.cartesian_grid_topology.dimension = len(shape)
    m.domain.cartesian_grid_topology.shape.extend(shape)
    if coords is None:
        m.domain.geometry = 0
    else:
        m.domain.geometry = 1
        m.domain.explicit_cube_geometry.coordinates.extend([float(x) for c in coords for x in c])
    m.dataset.type, m.dataset.dimension = (1 if is_double else 0), 1
    m.function_decomposition.hierarchy = 0
    m.error_control.mode = 0
    if np.isinf(s):
        m.error_control.norm = 0
    else:
        m.error_control.norm = 1
        m.error_control.s = s
    m.error_control.tolerance = tol
    return m.SerializeToString()


def main():
    Header = header_class()
    cases = [
        ((129, 129, 129), np.float64, False, np.inf, 1e-4),
        ((10, 7), np.float32, True, 0.0, 1e-2),
        ((33, 20, 17), np.float64, True, 1.5, 0.1),
        ((300,), np.float32, False, -1.0, 3.0),
        ((1000, 1000), np.float32, True, 0.0, 1e-2),
        ((4, 1, 9, 5), np.float64, False, 0.25, 1e-6),
    ]
    out = {"count": np.int64(len(cases))}
    for i, (shape, dt, explicit, s, tol) in enumerate(cases):
        coords = None
        if explicit:
            coords = [(np.linspace(0, 1, n) ** 1.5).astype(dt) for n in shape]
        # the reference widens its Real-typed s and tolerance to double
        s64, tol64 = float(dt(s)), float(dt(tol))
        blob = reference_header(Header, shape, dt is np.float64, coords, s64, tol64)
        out[f"shape{i}"] = np.array(shape, dtype=np.int64)
        out[f"dtype{i}"] = np.int64(1 if dt is np.float64 else 0)
        out[f"explicit{i}"] = np.int64(1 if explicit else 0)
        out[f"coords{i}"] = np.concatenate(coords).astype(np.float64) if explicit else np.zeros(0)
        out[f"s{i}"] = np.float64(s)
        out[f"tol{i}"] = np.float64(tol)
        out[f"bytes{i}"] = np.frombuffer(blob, dtype=np.uint8)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpu_headers.npz"), **out)


if __name__ == "__main__":
    main()
