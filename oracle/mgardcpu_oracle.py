"""TEST INFRASTRUCTURE — numpy restatement of the reference's MGARD-CPU path
(`mgard::compress` / `mgard::decompress`, reference include/compress.tpp:35-83).

Not part of the product: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  Every function
cites the reference lines it restates; the arithmetic keeps the reference's
operation order so that results are bit-identical to the compiled reference
(oracle/_ref/libmgard_cpu_ref.so, pinned in tests/test_cpu_convention.py
together with the reference's own known-answer vectors).

Layout: the reference works on a "shuffled" array (nodes ordered by the level
that introduced them); its operators address nodes through the hierarchy, so the
values do not depend on the layout.  This restatement keeps the array nodal
(row-major) and shuffles at the end.
"""
import ctypes
import math
import zlib

import numpy as np

_libm = ctypes.CDLL("libm.so.6")
_libm.exp2f.argtypes = [ctypes.c_float]
_libm.exp2f.restype = ctypes.c_float
_libm.exp2.argtypes = [ctypes.c_double]
_libm.exp2.restype = ctypes.c_double


class Hierarchy:
    """TensorMeshHierarchy<N, Real> (reference include/TensorMeshHierarchy.tpp:40-139).

    shapes[l][d]; indices[d][l] = floor(j (N_d - 1) / (n_l - 1)); dates of birth;
    uniform coordinates j * (1 / (n - 1)) evaluated in Real (:145-157)."""

    def __init__(self, shape, dtype, coords=None):
        self.real = np.dtype(dtype).type
        self.shape = tuple(int(n) for n in shape)
        self.N = len(self.shape)
        self.uniform = coords is None
        if coords is None:
            coords = []
            for n in self.shape:
                h = self.real(1) / self.real(n - 1) if n > 1 else self.real(0)
                coords.append(np.arange(n).astype(self.real) * h)
        self.coords = [np.asarray(c, dtype=self.real) for c in coords]
        if any(n == 0 for n in self.shape):
            raise ValueError("dataset must have size larger than 0 in every dimension")
        if all(n == 1 for n in self.shape):
            raise ValueError("dataset must have size larger than 1 in some dimension")
        nlev = [(n - 1).bit_length() - 1 for n in self.shape if n > 1]  # log2(n - 1)
        L_dyadic = min(nlev)
        rounded = [1 if n == 1 else (1 << ((n - 1).bit_length() - 1)) + 1 for n in self.shape]
        nondyadic = any(r != n for r, n in zip(rounded, self.shape))
        self.L = L_dyadic + 1 if nondyadic else L_dyadic
        cur = [((r - 1) >> L_dyadic) + 1 for r in rounded]
        self.shapes = []
        for _ in range(self.L):
            self.shapes.append(tuple(cur))
            cur = [((n - 1) << 1) + 1 for n in cur]
        self.shapes.append(self.shape)
        self.indices = []
        for d, n_top in enumerate(self.shape):
            per_level = []
            for l in range(self.L + 1):
                n = self.shapes[l][d]
                if n_top == 1:
                    per_level.append(np.zeros(1, dtype=np.int64))
                else:
                    per_level.append((np.arange(n, dtype=np.int64) * (n_top - 1)) // (n - 1))
            self.indices.append(per_level)
        self.dob = []
        for d, n_top in enumerate(self.shape):
            dob = np.zeros(n_top, dtype=np.int64)
            for l in range(self.L, -1, -1):
                dob[self.indices[d][l]] = l
            self.dob.append(dob)

    def ndof(self, l=None):
        return int(np.prod(self.shapes[self.L if l is None else l]))

    def box(self, l):
        """Index tuple selecting the level-l mesh inside the nodal array."""
        return np.ix_(*[self.indices[d][l] for d in range(self.N)])

    def old_masks(self, l):
        """Per dimension: which level-l nodes already belong to level l - 1."""
        return [self.dob[d][self.indices[d][l]] < l if self.shape[d] > 1
                else np.ones(1, dtype=bool) for d in range(self.N)]

    def new_mask(self, l):
        """Boolean level-l box: nodes introduced by level l (date_of_birth == l,
        TensorMeshHierarchy.tpp `date_of_birth` = max over dimensions)."""
        if l == 0:
            return np.ones(self.shapes[0], dtype=bool)
        old = self.old_masks(l)
        allold = np.ones(self.shapes[l], dtype=bool)
        for d in range(self.N):
            sh = [1] * self.N
            sh[d] = -1
            allold = allold & old[d].reshape(sh)
        return ~allold


def _ax(a, d, N):
    sh = [1] * N
    sh[d] = -1
    return np.asarray(a).reshape(sh)


def _take(b, idx, d):
    return np.take(b, idx, axis=d)


def _neighbours(h, l, d):
    """For the level-l nodes of dimension d that are new: their positions in the
    level-l index list and those of the enclosing level-(l-1) nodes (adjacent in
    the level-l list: at most one new node sits between two old ones)."""
    old = h.old_masks(l)[d]
    new = np.nonzero(~old)[0]
    return new, new - 1, new + 1


def prolongation_addition(h, l, b):
    """TensorProlongationAddition on the dense level-l box `b`
    (reference include/TensorProlongation.tpp:22-69), one dimension after another
    (include/TensorLinearOperator.tpp:71-109)."""
    for d in range(h.N):
        if h.shape[d] == 1:
            continue
        x = h.coords[d][h.indices[d][l]]
        new, lo, hi = _neighbours(h, l, d)
        if new.size == 0:
            continue
        xl, xm, xr = (_ax(x[i], d, h.N) for i in (lo, new, hi))
        wr = h.real(1) / (xr - xl)
        vl, vr = _take(b, lo, d), _take(b, hi, d)
        upd = _take(b, new, d) + (vl * (xr - xm) + vr * (xm - xl)) * wr
        sl = [slice(None)] * h.N
        sl[d] = new
        b[tuple(sl)] = upd
    return b


def mass_matrix(h, l, b):
    """TensorMassMatrix (reference include/TensorMassMatrix.tpp:15-90)."""
    for d in range(h.N):
        if h.shape[d] == 1:
            continue
        x = h.coords[d][h.indices[d][l]]
        n = x.size
        hs = _ax(x[1:] - x[:-1], d, h.N)  # h_right of node j = hs[j]
        out = np.empty_like(b)
        sl = lambda a, s: a[tuple([slice(None)] * d + [s] + [slice(None)] * (h.N - d - 1))]
        first = sl(hs, slice(0, 1))
        sl(out, slice(0, 1))[...] = first / 3 * sl(b, slice(0, 1)) + first / 6 * sl(b, slice(1, 2))
        if n > 2:
            hl, hr = sl(hs, slice(0, n - 2)), sl(hs, slice(1, n - 1))
            sl(out, slice(1, n - 1))[...] = (hl / 6 * sl(b, slice(0, n - 2))
                                             + (hl + hr) / 3 * sl(b, slice(1, n - 1))
                                             + hr / 6 * sl(b, slice(2, n)))
        last = sl(hs, slice(n - 2, n - 1))
        sl(out, slice(n - 1, n))[...] = (last / 6 * sl(b, slice(n - 2, n - 1))
                                         + last / 3 * sl(b, slice(n - 1, n)))
        b = out
    return b


def restriction(h, l, b):
    """TensorRestriction (reference include/TensorRestriction.tpp:24-71).  The
    result is only consumed on the level-(l-1) nodes, so each pass keeps the old
    nodes of its dimension.  A coarse node first receives the contribution of the
    interval on its left, then the one on its right (loop order of :52-69)."""
    for d in range(h.N):
        if h.shape[d] == 1:
            continue
        x = h.coords[d][h.indices[d][l]]
        old = h.old_masks(l)[d]
        oldpos = np.nonzero(old)[0]
        out = _take(b, oldpos, d).copy()
        new, lo, hi = _neighbours(h, l, d)
        if new.size:
            xl, xm, xr = (_ax(x[i], d, h.N) for i in (lo, new, hi))
            wr = h.real(1) / (xr - xl)
            vm = _take(b, new, d)
            to_left = vm * (xr - xm) * wr    # added to the old node at `lo`
            to_right = vm * (xm - xl) * wr   # added to the old node at `hi`
            rank = np.cumsum(old) - 1        # level-l position -> coarse position
            sl = [slice(None)] * h.N
            # interval on the left of a coarse node: it is the `hi` neighbour
            sl[d] = rank[hi]
            out[tuple(sl)] = _take(out, rank[hi], d) + to_right
            sl[d] = rank[lo]
            out[tuple(sl)] = _take(out, rank[lo], d) + to_left
        b = out
    return b


def mass_inverse_tables(h, l, d):
    """divisors of ConstituentMassMatrixInverse (TensorMassMatrix.tpp:123-176)."""
    x = h.coords[d][h.indices[d][l]]
    n = x.size
    hs = x[1:] - x[:-1]
    div = np.empty(n, dtype=h.real)
    div[0] = 2 * hs[0] / 6
    for j in range(1, n - 1):
        a = hs[j - 1] / 6
        w = a / div[j - 1]
        div[j] = 2 * (hs[j - 1] + hs[j]) / 6 - w * a
    a = hs[n - 2] / 6
    w = a / div[n - 2]
    div[n - 1] = 2 * hs[n - 2] / 6 - w * a
    return hs, div


def mass_inverse(h, l, b):
    """TensorMassMatrixInverse on the dense level-l box (TensorMassMatrix.tpp:178-290)."""
    for d in range(h.N):
        if h.shape[d] == 1:
            continue
        hs, div = mass_inverse_tables(h, l, d)
        n = div.size
        b = np.moveaxis(b, d, 0).copy()
        prev = b[0].copy()
        for j in range(1, n - 1):
            w = (hs[j - 1] / 6) / div[j - 1]
            b[j] = b[j] - w * prev
            prev = b[j]
        w = (hs[n - 2] / 6) / div[n - 2]
        b[n - 1] = b[n - 1] - w * prev
        b[n - 1] = b[n - 1] / div[n - 1]
        nxt = b[n - 1]
        for j in range(n - 2, -1, -1):
            c = hs[j] / 6
            b[j] = b[j] - c * nxt
            b[j] = b[j] / div[j]
            nxt = b[j]
        b = np.moveaxis(b, 0, d)
    return np.ascontiguousarray(b)


def _old_box_in_level(h, l):
    return np.ix_(*[np.nonzero(m)[0] for m in h.old_masks(l)])


def decompose_nodal(h, u):
    """mgard::decompose (reference include/decompose.tpp:129-174) on a nodal array."""
    v = np.array(u, dtype=h.real, copy=True).reshape(h.shape)
    for l in range(h.L, 0, -1):
        V = v[h.box(l)]
        oldbox = _old_box_in_level(h, l)
        new = h.new_mask(l)
        buf = np.zeros_like(V)
        buf[oldbox] = V[oldbox]                       # copy_on_old_zero_on_new
        buf = prolongation_addition(h, l, buf)
        V = np.where(new, V - buf, V)                 # zero_on_old_subtract_and_copy_back_on_new
        buf = np.where(new, V, h.real(0))
        buf = mass_matrix(h, l, buf)
        buf = restriction(h, l, buf)
        buf = mass_inverse(h, l - 1, buf)
        V[oldbox] = V[oldbox] + h.real(1) * buf       # add_on_old_add_on_new (axpy)
        v[h.box(l)] = V
    return v


def recompose_nodal(h, c):
    """mgard::recompose (reference include/decompose.tpp:177-219) on a nodal array."""
    v = np.array(c, dtype=h.real, copy=True).reshape(h.shape)
    for l in range(1, h.L + 1):
        V = v[h.box(l)]
        oldbox = _old_box_in_level(h, l)
        new = h.new_mask(l)
        buf = np.where(new, V, h.real(0))             # zero_on_old_copy_on_new
        buf = mass_matrix(h, l, buf)
        buf = restriction(h, l, buf)
        buf = mass_inverse(h, l - 1, buf)
        full = np.zeros_like(V)                       # subtract_on_old_zero_on_new
        full[oldbox] = buf + h.real(-1) * V[oldbox]
        full = prolongation_addition(h, l, full)
        out = np.where(new, V + h.real(-1) * full, V)  # copy_negation_on_old_subtract_on_new
        out[oldbox] = -full[oldbox]
        v[h.box(l)] = out
    return v


def shuffle(h, v):
    """mgard::shuffle (reference include/shuffle.tpp:8-21): level by level, the
    nodes a level introduces in row-major order of that level's mesh
    (ShuffledTensorNodeRange, TensorMeshHierarchyIteration.tpp:204-227)."""
    v = np.asarray(v).reshape(h.shape)
    return np.concatenate([v[h.box(l)][h.new_mask(l)] for l in range(h.L + 1)])


def unshuffle(h, s):
    s = np.asarray(s)
    v = np.empty(h.shape, dtype=s.dtype)
    pos = 0
    for l in range(h.L + 1):
        m = h.new_mask(l)
        k = int(m.sum())
        V = v[h.box(l)]
        V[m] = s[pos:pos + k]
        v[h.box(l)] = V
        pos += k
    return v


def decompose(h, u):
    """shuffle + decompose, as mgard::compress does (compress.tpp:39-44)."""
    return shuffle(h, decompose_nodal(h, u))


def recompose(h, c):
    return recompose_nodal(h, unshuffle(h, c))


def supremum_quantum(h, tol):
    """reference include/TensorMultilevelCoefficientQuantizer.tpp:13-27."""
    d = sum(1 for n in h.shape if n > 1)
    return h.real(float(h.real(2) * h.real(tol)) / ((h.L + 1) * (1 + math.pow(3, d))))


def quanta(h, s, tol):
    """Per-node quantum in shuffled order (ibid. :38-77): s = inf -> one value;
    otherwise 2 tol / (2^(s l) sqrt(ndof * volume)), the volume being taken in
    the mesh that introduced the node."""
    real = h.real
    if math.isinf(s):
        return np.full(h.ndof(), supremum_quantum(h, tol), dtype=real)
    out = []
    ndof = real(h.ndof())
    for l in range(h.L + 1):
        vf = np.ones(h.shapes[l], dtype=real)
        for d in range(h.N):
            if h.shape[d] == 1:
                continue
            x = h.coords[d][h.indices[d][l]]
            succ = np.concatenate([x[1:], x[-1:]])
            pred = np.concatenate([x[:1], x[:-1]])
            vf = vf * _ax((succ - pred) / 2, d, h.N)
        e2 = real(_libm.exp2f(float(real(s) * real(l)))) if real is np.float32 \
            else real(_libm.exp2(float(s) * l))
        q = (real(2) * real(tol)) / (e2 * np.sqrt(ndof * vf))
        out.append(q[h.new_mask(l)])
    return np.concatenate(out).astype(real)


def quantize(h, s, tol, coeffs):
    """LinearQuantizer (reference include/LinearQuantizer.tpp:8-26) with the
    multilevel quantum: copysign(0.5 + |x / quantum|, x) truncated to int64."""
    c = np.asarray(coeffs, dtype=h.real)
    q = quanta(h, s, tol)
    mag = 0.5 + np.abs(c / q).astype(np.float64)
    return np.copysign(mag, c.astype(np.float64)).astype(np.int64)


def dequantize(h, s, tol, ints):
    """LinearDequantizer (ibid. :41-53): quantum * n evaluated in Real."""
    return quanta(h, s, tol) * np.asarray(ints).astype(h.real)


def zlib_payload(ints):
    """compress_memory_z (reference src/compressors.cpp:552-606): one deflate
    stream at Z_BEST_COMPRESSION over the raw int64 array."""
    return zlib.compress(np.ascontiguousarray(ints, dtype=np.int64).tobytes(), 9)


def _varint(x):
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        if x:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _f_varint(field, x):
    return b"" if x == 0 else _varint(field << 3) + _varint(x)


def _f_double(field, x):
    import struct
    bits = struct.pack("<d", x)
    return b"" if bits == b"\0" * 8 else _varint((field << 3) | 1) + bits


def _f_msg(field, body):
    return _varint((field << 3) | 2) + _varint(len(body)) + body


def header_bytes(h, s, tol, compressor=1):
    """The protobuf header mgard::compress builds (populate_defaults,
    reference src/format.cpp:102-140; TensorMeshHierarchy::populate,
    include/TensorMeshHierarchy.tpp:293-348; error control, compress.tpp:45-55)
    in proto3 canonical form (src/mgard.proto); compressor 1 = CPU_HUFFMAN_ZLIB
    (build without MGARD_ZSTD), 2 = CPU_HUFFMAN_ZSTD (default build)."""
    import struct
    topo = _f_varint(1, h.N) + _f_msg(2, b"".join(_varint(n) for n in h.shape))
    dom = _f_msg(2, topo)
    if not h.uniform:
        packed = b"".join(struct.pack("<d", float(x)) for c in h.coords for x in c)
        dom += _f_varint(3, 1) + _f_msg(4, _f_msg(2, packed))
    err = b""
    if not math.isinf(s):
        err += _f_varint(2, 1) + _f_double(3, float(h.real(s)))
    err += _f_double(5, float(h.real(tol)))
    hdr = _f_msg(2, _f_varint(1, 1) + _f_varint(2, 6) + _f_varint(3, 0))
    hdr += _f_msg(3, _f_varint(1, 1))
    hdr += _f_msg(4, dom)
    hdr += _f_msg(5, _f_varint(1, 1 if h.real is np.float64 else 0) + _f_varint(2, 1))
    hdr += _f_msg(6, err)
    hdr += _f_msg(8, b"")
    hdr += _f_msg(9, _f_varint(1, 1) + _f_varint(3, 3))
    hdr += _f_msg(11, _f_varint(1, 1) + _f_varint(2, compressor))
    hdr += _f_msg(12, b"")
    return hdr


def stream(h, s, tol, payload, compressor=1):
    """CompressedDataset::write (reference include/CompressedDataset.tpp:26-29)
    with write_metadata (src/format.cpp:219-233): magic, u64 header size, u32
    CRC32 of the header -- both BIG-endian (serialize<>, include/format.tpp:27-41;
    known answers in tests/src/test_format.cpp:23-50) --, header, payload."""
    hdr = header_bytes(h, s, tol, compressor)
    return preamble(hdr) + hdr + bytes(payload)


def preamble(hdr):
    import struct
    return b"MGARD" + struct.pack(">Q", len(hdr)) + struct.pack(">I", zlib.crc32(hdr))


def read_preamble(blob):
    """read_metadata up to the header bytes (src/format.cpp:150-208): returns
    (header_size, crc32); raises on a bad magic number or CRC."""
    import struct
    if blob[:5] != b"MGARD":
        raise ValueError("bad magic number")
    size, crc = struct.unpack(">QI", blob[5:17])
    if size > len(blob) - 17 or zlib.crc32(blob[17:17 + size]) != crc:
        raise ValueError("header CRC32 mismatch")
    return size, crc


def compress(h, u, s, tol, compressor=1):
    """mgard::compress (reference include/compress.tpp:35-67) + write."""
    q = quantize(h, s, tol, decompose(h, u))
    return stream(h, s, tol, zlib_payload(q) if compressor == 1 else huffman_zstd_payload(q), compressor)


# ---- CPU_HUFFMAN_ZSTD payload (reference src/compressors.cpp, MGARD_ZSTD build) ----
NQL = 32768 * 4


def _push_heap(heap, hole, top, value, cnt):
    # libstdc++ std::__push_heap with LessThanByCnt (compressors.cpp:59-63): the
    # reference's std::priority_queue decides how equal counts are ordered, so the
    # container's algorithm is restated rather than replaced by heapq
    parent = (hole - 1) // 2
    while hole > top and cnt[heap[parent]] > cnt[value]:
        heap[hole] = heap[parent]
        hole = parent
        parent = (hole - 1) // 2
    heap[hole] = value


def _pq_push(heap, value, cnt):
    heap.append(value)
    _push_heap(heap, len(heap) - 1, 0, value, cnt)


def _pq_pop(heap, cnt):
    # std::pop_heap (std::__pop_heap + std::__adjust_heap) then pop_back
    top = heap[0]
    value = heap[-1]
    length = len(heap) - 1
    if length > 0:
        hole, child = 0, 0
        while child < (length - 1) // 2:
            child = 2 * (child + 1)
            if cnt[heap[child]] > cnt[heap[child - 1]]:
                child -= 1
            heap[hole] = heap[child]
            hole = child
        if (length & 1) == 0 and child == (length - 2) // 2:
            child = 2 * (child + 1)
            heap[hole] = heap[child - 1]
            hole = child - 1
        _push_heap(heap, hole, 0, value, cnt)
    heap.pop()
    return top


def huffman_codes(ft):
    """build_tree + build_codec (compressors.cpp:70-115): {symbol: (code, len)}."""
    cnt, sym, left, right, heap = [], [], [], [], []
    for i in np.nonzero(ft)[0]:
        cnt.append(int(ft[i])); sym.append(int(i)); left.append(-1); right.append(-1)
        _pq_push(heap, len(cnt) - 1, cnt)
    if not heap:
        return {}
    while len(heap) > 1:
        a = _pq_pop(heap, cnt)
        b = _pq_pop(heap, cnt)
        cnt.append(cnt[a] + cnt[b]); sym.append(-1); left.append(a); right.append(b)
        _pq_push(heap, len(cnt) - 1, cnt)
    codes, stack = {}, [(heap[0], 0, 0)]
    while stack:
        node, code, length = stack.pop()
        if left[node] < 0:
            codes[sym[node]] = (code, length)
        else:
            stack.append((left[node], code << 1, length + 1))
            stack.append((right[node], (code << 1) | 1, length + 1))
    return codes


def huffman_payload(ints):
    """huffman_encoding + the concatenation of compress_memory_huffman
    (compressors.cpp:316-463): returns (tree_bytes, hit_bits, miss_bytes, payload)
    before the zstd stage."""
    import struct
    q = np.asarray(ints, dtype=np.int64) + NQL // 2
    qi = q.astype(np.int32)  # `int q = quantized_data[i]` (:343)
    inrange64 = (q > 0) & (q < NQL)
    ft = np.bincount(np.where(inrange64, q, 0), minlength=NQL)
    codes = huffman_codes(ft)
    hit = (qi > 0) & (qi < NQL)
    symbols = np.where(hit, qi, 0)
    length = np.zeros(NQL, dtype=np.int64)
    code = np.zeros(NQL, dtype=np.uint64)
    for sym_, (c, l) in codes.items():
        code[sym_], length[sym_] = c, l
    lens = length[symbols]
    cds = code[symbols]
    total = int(lens.sum())
    bits = np.zeros(total, dtype=np.uint8)
    starts = np.cumsum(lens) - lens
    for b in range(int(lens.max()) if lens.size else 0):
        sel = lens > b
        bits[starts[sel] + b] = ((cds[sel] >> (lens[sel] - 1 - b).astype(np.uint64)) & np.uint64(1)).astype(np.uint8)
    nwords = (total // 8 + 4 + 3) // 4 + 1
    padded = np.zeros(nwords * 32, dtype=np.uint8)
    padded[:total] = bits
    words = np.packbits(padded).view(">u4").astype("<u4")  # MSB-first 32-bit words
    hit_bytes = words.tobytes()[: total // 8 + 4]
    tree = b"".join(struct.pack("<QQ", int(i), int(ft[i])) for i in np.nonzero(ft)[0])
    miss = qi[~hit].astype("<i4").tobytes()
    return len(tree), total, len(miss), tree + hit_bytes + miss


def huffman_zstd_payload(ints):
    """compress_memory_huffman (compressors.cpp:421-512): three sizes, then the
    zstd level-1 frame (compress_memory_zstd, :542-549) from the system libzstd."""
    import struct
    tree_bytes, hit_bits, miss_bytes, payload = huffman_payload(ints)
    z = ctypes.CDLL("libzstd.so.1")
    z.ZSTD_compressBound.restype = ctypes.c_size_t
    z.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
    z.ZSTD_compress.restype = ctypes.c_size_t
    z.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int]
    cap = z.ZSTD_compressBound(len(payload))
    dst = ctypes.create_string_buffer(cap)
    n = z.ZSTD_compress(dst, cap, payload, len(payload), 1)
    return struct.pack("<QQQ", tree_bytes, hit_bits, miss_bytes) + dst.raw[:n]
