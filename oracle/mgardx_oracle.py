"""TEST INFRASTRUCTURE — CPU restatement (numpy) of the MGARD-X hot path.

This file restates, operation for operation and in the working precision T,
what the reference's MGARD-X kernels compute, so that results are BIT-EXACT
with the reference's SERIAL adapter built as oracle/_ref (parity pinned by
tests/test_oracle_vs_ref.py against the reference itself, and by the committed
fixtures under tests/golden/ generated from it).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it; the product path (mgard_b200/) never does.

Reference provenance (paths under /root/reference/include/mgard-x unless noted):
  Hierarchy/Hierarchy.hpp:23-190,193-418      level shapes, dist, ratio, am/bm, volumes
  DataRefactoring/MultiDimension/DataRefactoring.hpp:25-317   level loop
  .../Coefficient/GridProcessingKernel3D.hpp:21-1229,1231-2400 + GPKFunctor.h:13-26 (lerp)
  .../Correction/LinearProcessingKernel3D.hpp + LPKFunctor.h:14-66 (mass_trans)
  .../Correction/IterativeProcessingKernel3D.hpp + IPKFunctor.h:14-51 (Thomas)
  Quantization/LinearQuantization.hpp:148-266,495-545
  Lossless/ParallelHuffman/{GetCodebook,GenerateCL,GenerateCW,Deflate,Decode,Huffman}.hpp
  src/mgard-x/Metadata/Metadata.cpp:249-462, src/mgard.proto
"""
import struct
import zlib

import numpy as np

REL, ABS = 0, 1
UINT_MAX = 0xFFFFFFFF

# --------------------------------------------------------------------------
# Hierarchy (Hierarchy.hpp:193-418)
# --------------------------------------------------------------------------


def level_shapes(shape):
    """n -> n/2+1 until 2; l_target = min over dims (Hierarchy.hpp:199-230)."""
    per_dim = []
    for n in shape:
        sizes = []
        while n > 2:
            sizes.append(int(n))
            n = n // 2 + 1
        sizes.append(2)
        per_dim.append(sizes)
    L = min(len(s) for s in per_dim) - 1
    return [[per_dim[d][L - l] for d in range(len(shape))] for l in range(L + 1)]


def _coord_to_dist(coord, T):
    """Hierarchy.hpp:23-51."""
    n = coord.size
    dist = np.zeros(n, dtype=T)
    dist[: n - 1] = coord[1:] - coord[:-1]
    if n != 2 and n % 2 == 0:
        last = dist[n - 2]
        dist[n - 2] = T(np.float64(last) / 2.0)
        dist[n - 1] = T(np.float64(last) / 2.0)
    return dist


def _dist_to_ratio(dist, T):
    """Hierarchy.hpp:53-80."""
    n = dist.size
    ratio = np.zeros(n, dtype=T)
    if n > 2:
        ratio[: n - 2] = dist[: n - 2] / (dist[1: n - 1] + dist[: n - 2])
    if n % 2 == 0:
        ratio[n - 2] = dist[n - 2] / (dist[n - 1] + dist[n - 2])
    return ratio


def _reduce_dist(dist, T):
    """Hierarchy.hpp:82-109."""
    n = dist.size
    n2 = n // 2 + 1
    d2 = np.zeros(n2, dtype=T)
    for i in range(n2 - 1):
        d2[i] = dist[2 * i] + dist[2 * i + 1]
    if n2 != 2 and n2 % 2 == 0:
        last = d2[n2 - 2]
        d2[n2 - 2] = T(np.float64(last) / 2.0)
        d2[n2 - 1] = T(np.float64(last) / 2.0)
    return d2


def _calc_am_bm(dist, T):
    """Hierarchy.hpp:112-162 (non-FMA): am[i]=h_{i-1}/6, bm[i+1]=divisor_i,
    bm[0]=1, am[n]=0."""
    n = dist.size
    ham = np.zeros(n + 1, dtype=T)
    hbm = np.zeros(n + 1, dtype=T)
    hbm[0] = T(2) * dist[0] / T(6)
    for i in range(1, n - 1):
        a_j = dist[i - 1] / T(6)
        w = a_j / hbm[i - 1]
        hbm[i] = T(2) * (dist[i - 1] + dist[i]) / T(6) - w * a_j
        ham[i] = a_j
    a_j = dist[n - 2] / T(6)
    w = a_j / hbm[n - 2]
    hbm[n - 1] = T(2) * dist[n - 2] / T(6) - w * a_j
    ham[n - 1] = a_j
    am = np.zeros(n + 1, dtype=T)
    bm = np.zeros(n + 1, dtype=T)
    am[:n] = ham[:n]
    bm[1: n + 1] = hbm[:n]
    bm[0] = T(1)
    am[n] = T(0)
    return am, bm


class Hierarchy:
    def __init__(self, shape, dtype, coords=None):
        self.T = T = np.dtype(dtype).type
        self.shape = [int(s) for s in shape]
        self.D = D = len(shape)
        assert all(s >= 3 for s in self.shape)
        self.level_shape = level_shapes(self.shape)
        self.l_target = L = len(self.level_shape) - 1
        if coords is None:
            # create_uniform_coords, normalize_coordinates=true (Hierarchy.hpp:686-706)
            coords = [np.arange(n, dtype=T) / T(n - 1) for n in self.shape]
            self.uniform = True
        else:
            coords = [np.asarray(c, dtype=T) for c in coords]
            self.uniform = False
        self.coords = coords
        with np.errstate(all="ignore"):
            self.dist = [[None] * D for _ in range(L + 1)]
            self.ratio = [[None] * D for _ in range(L + 1)]
            for d in range(D):
                self.dist[L][d] = _coord_to_dist(coords[d], T)
                self.ratio[L][d] = _dist_to_ratio(self.dist[L][d], T)
            for l in range(L - 1, -1, -1):
                for d in range(D):
                    self.dist[l][d] = _reduce_dist(self.dist[l + 1][d], T)
                    self.ratio[l][d] = _dist_to_ratio(self.dist[l][d], T)
            self.am = [[None] * D for _ in range(L + 1)]
            self.bm = [[None] * D for _ in range(L + 1)]
            for l in range(L + 1):
                for d in range(D):
                    self.am[l][d], self.bm[l][d] = _calc_am_bm(self.dist[l][d], T)
        # level_marks (Hierarchy.hpp:262-282): smallest l with i < level_shape[l][d]
        self.level_marks = []
        for d in range(D):
            m = np.zeros(self.shape[d], dtype=np.int32)
            i = 0
            for l in range(L + 1):
                m[i: self.level_shape[l][d]] = l
                i = self.level_shape[l][d]
            self.level_marks.append(m)
        # calc_volume (Hierarchy.hpp:165-190): 1/(dof-1) as T; reciprocal 1/that
        self.volume = [[T(1.0 / float(T(n - 1))) for n in self.level_shape[l]]
                       for l in range(L + 1)]
        self.volume_recip = [[T(1.0 / float(v)) for v in row] for row in self.volume]

    @property
    def total(self):
        return int(np.prod(self.shape))


# --------------------------------------------------------------------------
# Level kernels
# --------------------------------------------------------------------------


def _lerp(v0, v1, t):
    """GPKFunctor.h:13-26 non-FMA branch: r = v0 + v0*t*-1; r = r + t*v1."""
    T = v0.dtype.type
    r = v0 + v0 * t * T(-1)
    return r + t * v1


def _bshape(vec, axis, ndim):
    s = [1] * ndim
    s[axis] = vec.size
    return vec.reshape(s)


def _pad_even(box):
    """Replicate the last node along every even-sized dim (ghost node,
    GridProcessingKernel3D.hpp:171-300)."""
    for ax, n in enumerate(box.shape):
        if n > 1 and n % 2 == 0:
            last = np.take(box, [n - 1], axis=ax)
            box = np.concatenate([box, last], axis=ax)
    return box


def _sl(ndim, axis, s):
    idx = [slice(None)] * ndim
    idx[axis] = s
    return tuple(idx)


def _interpolate(P, ratios):
    """Multilinear interpolant of the even-index (coarse) nodes of padded P,
    built dimension by dimension, fastest dim first (f, then c, then r...)."""
    A = P.copy()
    nd = P.ndim
    for ax in range(nd - 1, -1, -1):
        n = P.shape[ax]
        if n < 3:
            continue
        left = A[_sl(nd, ax, slice(0, n - 2, 2))]
        right = A[_sl(nd, ax, slice(2, n, 2))]
        t = _bshape(ratios[ax][0: n - 2: 2], ax, nd)
        A[_sl(nd, ax, slice(1, n - 1, 2))] = _lerp(left, right, t)
    return A


def _to_octants(X, n, nc):
    """Nodal (padded) -> coarse-first layout along every dim: even padded
    indices go to [0:nc), odd indices 1,3,.. go to [nc:n)."""
    nd = X.ndim
    for ax in range(nd):
        if n[ax] == 1:
            continue
        ev = X[_sl(nd, ax, slice(0, None, 2))]
        od = X[_sl(nd, ax, slice(1, 2 * (n[ax] - nc[ax]), 2))]
        X = np.concatenate([ev, od], axis=ax)
    return X


def _from_octants(V, n, nc):
    """Inverse of _to_octants: returns padded nodal array (odd sizes); the
    dropped odd slot of an even dim is zero-filled."""
    nd = V.ndim
    for ax in range(nd):
        if n[ax] == 1:
            continue
        npad = 2 * nc[ax] - 1
        shp = list(V.shape)
        shp[ax] = npad
        X = np.zeros(shp, dtype=V.dtype)
        X[_sl(nd, ax, slice(0, None, 2))] = V[_sl(nd, ax, slice(0, nc[ax]))]
        ncoef = n[ax] - nc[ax]
        X[_sl(nd, ax, slice(1, 2 * ncoef, 2))] = V[_sl(nd, ax, slice(nc[ax], n[ax]))]
        V = X
    return V


def _mass_trans_axis(W, ax, n, nc, dist):
    """Lpk{1,2,3}Reo3D + mass_trans (LPKFunctor.h:47-66, non-FMA): fused
    mass-matrix x restriction along `ax` on the coarse-first layout."""
    T = W.dtype.type
    nd = W.ndim
    npad = 2 * nc - 1
    # padded nodal line: coarse at even slots, coefficients at odd (missing -> 0)
    shp = list(W.shape)
    shp[ax] = npad + 4  # two zero slots each side
    X = np.zeros(shp, dtype=T)
    X[_sl(nd, ax, slice(2, 2 + npad, 2))] = W[_sl(nd, ax, slice(0, nc))]
    ncoef = n - nc
    X[_sl(nd, ax, slice(3, 3 + 2 * ncoef, 2))] = W[_sl(nd, ax, slice(nc, n))]
    h = np.zeros(npad + 5, dtype=T)  # h[k+2] = dist[k], zero outside [0,n)
    h[2: 2 + n] = dist[:n]
    i = np.arange(nc)
    h1 = h[2 * i]
    h2 = h[2 * i + 1]
    h3 = h[2 * i + 2]
    h4 = h[2 * i + 3]
    with np.errstate(all="ignore"):
        r1 = np.where(h1 + h2 != 0, h1 / (h1 + h2), T(0)).astype(T)
        r4 = np.where(h3 + h4 != 0, h4 / (h3 + h4), T(0)).astype(T)
    six, three = T(6), T(3)
    c16, c13, c26 = h1 / six, (h1 + h2) / three, h2 / six
    c23, c36 = (h2 + h3) / three, h3 / six
    c34, c46 = (h3 + h4) / three, h4 / six
    B = lambda v: _bshape(v, ax, nd)
    a = X[_sl(nd, ax, slice(0, 2 * nc, 2))]
    b = X[_sl(nd, ax, slice(1, 2 * nc + 1, 2))]
    c = X[_sl(nd, ax, slice(2, 2 * nc + 2, 2))]
    d = X[_sl(nd, ax, slice(3, 2 * nc + 3, 2))]
    e = X[_sl(nd, ax, slice(4, 2 * nc + 4, 2))]
    tb = a * B(c16) + b * B(c13) + c * B(c26)
    tc = b * B(c26) + c * B(c23) + d * B(c36)
    td = c * B(c36) + d * B(c34) + e * B(c46)
    return tc + (tb * B(r1) + td * B(r4))


def _thomas_axis(X, ax, am, bm):
    """Ipk{1,2,3}Reo3D + tridiag_forward2/backward2 (IPKFunctor.h:14-51)."""
    nd = X.ndim
    n = X.shape[ax]
    X = X.copy()
    prev = np.zeros_like(X[_sl(nd, ax, 0)])
    for i in range(n):
        cur = X[_sl(nd, ax, i)] - prev * (am[i] / bm[i])
        X[_sl(nd, ax, i)] = cur
        prev = cur
    prev = np.zeros_like(prev)
    for k in range(n - 1, -1, -1):
        cur = (X[_sl(nd, ax, k)] - am[k + 1] * prev) / bm[k + 1]
        X[_sl(nd, ax, k)] = cur
        prev = cur
    return X


def _correction(h, V, l):
    """CalcCorrection3D.hpp:30-196: coefficient function (coarse block zeroed)
    -> mass*restriction f,c,r -> Thomas f,c,r with level l-1 tables."""
    n = h.level_shape[l]
    nc = h.level_shape[l - 1]
    W = V.copy()
    W[tuple(slice(0, c) for c in nc)] = 0
    nd = W.ndim
    for ax in range(nd - 1, -1, -1):
        W = _mass_trans_axis(W, ax, n[ax], nc[ax], h.dist[l][ax])
    for ax in range(nd - 1, -1, -1):
        W = _thomas_axis(W, ax, h.am[l - 1][ax], h.bm[l - 1][ax])
    return W


def decompose(h, u):
    """multi_dimension::decompose (DataRefactoring.hpp:25-177), D <= 3 path
    generalised dimension-by-dimension."""
    v = np.array(u, dtype=h.T, copy=True, order="C")
    assert list(v.shape) == h.shape
    with np.errstate(all="ignore"):
        for l in range(h.l_target, 0, -1):
            n = h.level_shape[l]
            nc = h.level_shape[l - 1]
            box = tuple(slice(0, k) for k in n)
            cbox = tuple(slice(0, k) for k in nc)
            P = _pad_even(v[box])
            A = _interpolate(P, h.ratio[l])
            C = P - A
            # coarse nodes keep their value
            ev = tuple(slice(0, None, 2) for _ in n)
            C[ev] = P[ev]
            V = _to_octants(C, n, nc)
            corr = _correction(h, V, l)
            V[cbox] = V[cbox] + corr
            v[box] = V
    return v


def recompose(h, v):
    """multi_dimension::recompose (DataRefactoring.hpp:180-317)."""
    v = np.array(v, dtype=h.T, copy=True, order="C")
    with np.errstate(all="ignore"):
        for l in range(1, h.l_target + 1):
            n = h.level_shape[l]
            nc = h.level_shape[l - 1]
            box = tuple(slice(0, k) for k in n)
            cbox = tuple(slice(0, k) for k in nc)
            V = v[box].copy()
            corr = _correction(h, V, l)
            V[cbox] = V[cbox] - corr
            X = _from_octants(V, n, nc)  # padded nodal: coarse at even, coef at odd
            ev = tuple(slice(0, None, 2) for _ in n)
            Pc = np.zeros_like(X)
            Pc[ev] = X[ev]
            A = _interpolate(Pc, h.ratio[l])
            R = X + A
            R[ev] = X[ev]
            # drop ghost slots: for even n the ghost (padded index n) is node n-1
            for ax in range(R.ndim):
                if n[ax] > 1 and n[ax] % 2 == 0:
                    nd = R.ndim
                    keep = R[_sl(nd, ax, slice(0, n[ax]))].copy()
                    keep[_sl(nd, ax, n[ax] - 1)] = R[_sl(nd, ax, n[ax])]
                    R = keep
            v[box] = R
    return v


# --------------------------------------------------------------------------
# Norm, quantizer (NormCalculator.hpp:13-83, LinearQuantization.hpp)
# --------------------------------------------------------------------------



# --------------------------------------------------------------------------
# decomposition_type::SingleDim (DataRefactoring/SingleDimension/*.hpp)
# --------------------------------------------------------------------------


def _mass_trans_13(a, b, c, d, e, h1, h2, h3, h4):
    """mass_trans with explicit spacings (LPKFunctor.h:14-66, non-FMA branch):
    r1, r4 are recomputed from the spacings."""
    T = b.dtype.type
    with np.errstate(all="ignore"):
        r1 = np.where(h1 + h2 != 0, h1 / (h1 + h2), T(0)).astype(T)
        r4 = np.where(h3 + h4 != 0, h4 / (h3 + h4), T(0)).astype(T)
    tb = a * (h1 / 6) + b * ((h1 + h2) / 3) + c * (h2 / 6)
    tc = b * (h2 / 6) + c * ((h2 + h3) / 3) + d * (h3 / 6)
    td = c * (h3 / 6) + d * ((h3 + h4) / 3) + e * (h4 / 6)
    return tc + (tb * r1 + td * r4)


def _single_dim_correction(h, coeff, nc, l, ax):
    """CalcCorrection (SingleDimension/Correction/CalcCorrection.hpp:26-96): the
    load vector of the coefficients along one dimension (MassTransKernel.hpp:36-101,
    including its treatment of the last coarse nodes) and the Thomas solve with the
    level l-1 tables."""
    T = h.T
    nd = coeff.ndim
    ncoef = coeff.shape[ax]
    n = ncoef + nc
    dist = h.dist[l][ax]
    j = np.arange(nc)
    zero = np.zeros((), dtype=T)
    has_b = (j > 0) & (j < ncoef)
    has_d = j < ncoef
    b = np.where(_bshape(has_b, ax, nd), np.take(coeff, np.clip(j - 1, 0, ncoef - 1), axis=ax), zero)
    d = np.where(_bshape(has_d, ax, nd), np.take(coeff, np.clip(j, 0, ncoef - 1), axis=ax), zero)
    left = (j > 0) & (2 * j < n - 1)
    right = 2 * j < n - 1
    pick = lambda k, m: np.where(m, dist[np.clip(k, 0, len(dist) - 1)], T(0)).astype(T)
    h1, h2 = pick(2 * j - 2, left), pick(2 * j - 1, left)
    h3, h4 = pick(2 * j, right), pick(2 * j + 1, right)
    z = np.zeros_like(b)
    w = _mass_trans_13(z, b, z, d, z, *(_bshape(x, ax, nd) for x in (h1, h2, h3, h4))).astype(T)
    return _thomas_axis(w, ax, h.am[l - 1][ax], h.bm[l - 1][ax])


def decompose_single(h, u):
    """single_dimension::decompose (SingleDimension/DataRefactoring.hpp:25-108)."""
    T = h.T
    v = np.array(u, dtype=T, copy=True)
    D = h.D
    for l in range(h.l_target, 0, -1):
        for ax in range(D - 1, -1, -1):
            fine = [h.level_shape[l - 1][d] if d > ax else h.level_shape[l][d] for d in range(D)]
            box = tuple(slice(0, m) for m in fine)
            w = v[box].copy()
            n, nc = fine[ax], h.level_shape[l - 1][ax]
            ncoef = n - nc
            i = np.arange(ncoef)
            ratio = _bshape(h.ratio[l][ax][2 * i], ax, D)
            left, mid, right = (np.take(w, 2 * i + k, axis=ax) for k in (0, 1, 2))
            coeff = mid - _lerp(left, right, ratio)
            # CoefficientKernel.hpp:88-104: coarse = even nodes (+ the last node of an even size)
            cidx = list(2 * i) + [2 * ncoef] + ([2 * ncoef + 1] if n % 2 == 0 else [])
            coarse = np.take(w, cidx, axis=ax)
            coarse = coarse + _single_dim_correction(h, coeff, nc, l, ax)
            v[box] = np.concatenate([coarse, coeff], axis=ax)
    return v


def recompose_single(h, c):
    """single_dimension::recompose (SingleDimension/DataRefactoring.hpp:110-194)."""
    T = h.T
    v = np.array(c, dtype=T, copy=True)
    D = h.D
    for l in range(h.l_target):
        for ax in range(D):
            fine = [h.level_shape[l][d] if d > ax else h.level_shape[l + 1][d] for d in range(D)]
            box = tuple(slice(0, m) for m in fine)
            cur = v[box]
            n, nc = fine[ax], h.level_shape[l][ax]
            ncoef = n - nc
            coarse = np.take(cur, np.arange(nc), axis=ax)
            coeff = np.take(cur, np.arange(nc, n), axis=ax)
            coarse = coarse - _single_dim_correction(h, coeff, nc, l + 1, ax)
            i = np.arange(ncoef)
            ratio = _bshape(h.ratio[l + 1][ax][2 * i], ax, D)
            left = np.take(coarse, i, axis=ax)
            right = np.take(coarse, i + 1, axis=ax)
            mid = coeff + _lerp(left, right, ratio)
            out = np.empty_like(cur)
            sl = lambda idx: tuple(idx if d == ax else slice(None) for d in range(D))
            out[sl(2 * i)] = left
            out[sl(2 * i + 1)] = mid
            out[sl([2 * ncoef])] = np.take(coarse, [ncoef], axis=ax)
            if n % 2 == 0:
                out[sl([2 * ncoef + 1])] = np.take(coarse, [ncoef + 1], axis=ax)
            v[box] = out
    return v


def calc_norm(u, s):
    T = u.dtype.type
    if np.isinf(s):
        norm = T(np.max(np.abs(u)))
    else:
        # sequential sum in T is what SERIAL does; tests only use this loosely
        norm = T(np.sqrt(T(np.sum(u.astype(np.float64) ** 2)) / T(u.size)))
    if norm == 0:
        norm = np.finfo(T).eps
    return T(norm)


def calc_quantizers(h, ebtype, tol, s, norm, reciprocal, single_dim=False):
    """LinearQuantizer::CalcQuantizers (LinearQuantization.hpp:495-545)."""
    T = h.T
    abs_tol = float(T(tol))
    if ebtype == REL:
        abs_tol *= float(T(norm))
    abs_tol *= 2
    L = h.l_target
    q = np.zeros(L + 1, dtype=T)
    for l in range(L + 1):
        if np.isinf(s) and single_dim:
            q[l] = T(abs_tol / ((L + 1) * h.D * (1 + 3.0 ** 1)))
        elif np.isinf(s):
            q[l] = T(abs_tol / ((L + 1) * (1 + 3.0 ** h.D)))
        else:
            # std::exp2(s * l) is evaluated in T (float overload for fp32)
            e2 = float(np.exp2(T(T(s) * T(l))))
            q[l] = T(abs_tol / (e2 * np.sqrt(float(h.total))))
        if reciprocal:
            q[l] = T(np.float32(1.0) / q[l]) if T is np.float32 else T(1.0 / q[l])
    return q


def node_levels(h):
    lv = np.zeros(h.shape, dtype=np.int32)
    for d in range(h.D):
        lv = np.maximum(lv, _bshape(h.level_marks[d], d, h.D))
    return lv


def level_linear_index(h):
    """Config::reorder = 1: position of every element of the decomposed array in
    the level-linearised quantised array (`calc_level_offset`,
    LinearQuantization.hpp:46-146, placed behind the previous levels,
    :591-604).  Within a level the nodes it introduces keep the row-major order
    of that level's un-reordered mesh."""
    D = h.D
    lv = node_levels(h)
    idx = np.meshgrid(*[np.arange(n, dtype=np.int64) for n in h.shape], indexing="ij")
    ranges = np.zeros((h.l_target + 2, D), dtype=np.int64)
    for l in range(h.l_target + 1):
        ranges[l + 1] = h.level_shape[l]
    out = np.zeros(h.shape, dtype=np.int64)
    for l in range(h.l_target + 1):
        m = lv == l
        if not m.any():
            continue
        g = []
        for d in range(D):
            i = idx[d][m]
            bit = (np.asarray(h.level_marks[d])[i] == l).astype(np.int64)
            t = np.where(bit == 1, i - ranges[l][d], i)
            n = ranges[l + 1][d]
            if l == 0:
                gd = t
            else:
                gd = np.where((n % 2 == 0) & (t == n // 2), n - 1, t * 2 + bit)
            g.append(gd)
        cto = np.zeros_like(g[0])
        stride = 1
        for d in range(D - 1, -1, -1):
            cto = cto + g[d] * stride
            stride *= int(ranges[l + 1][d])
        clo = np.zeros_like(g[0])
        stride = 1
        for d in range(D - 1, -1, -1):
            n = int(ranges[l + 1][d])
            clo = np.where((g[d] % 2 != 0) & (g[d] != n - 1), 0, clo)
            clo = clo + np.where(g[d] != 0, ((g[d] - 1) // 2 + 1) * stride, 0)
            stride *= n // 2 + 1
        if l == 0:
            clo = np.zeros_like(clo)
        out[m] = int(np.prod(ranges[l])) + cto - clo
    return out


def linearize(h, q, oidx, oval):
    """Quantised symbols and outlier indices in level-linearised order."""
    lin = level_linear_index(h).ravel()
    ql = np.empty(lin.size, dtype=np.asarray(q).dtype)
    ql[lin] = np.asarray(q).ravel()
    if len(oidx):
        # canonical list order: ascending stored (linearised) position.  The
        # reference appends in the order its thread blocks run (tile by tile in
        # the SERIAL adapter, atomic arrival on GPUs); decoding is order-blind.
        new = lin[np.asarray(oidx, dtype=np.int64)]
        order = np.argsort(new, kind="stable")
        return ql, new[order].astype(np.uint64), np.asarray(oval)[order]
    return ql, np.asarray(oidx), np.asarray(oval)


def delinearize(h, ql):
    return np.asarray(ql).ravel()[level_linear_index(h).ravel()].reshape(h.shape)


def quantize(h, v, ebtype, tol, s, norm, dict_size=8192, single_dim=False):
    """LevelwiseLinearQuantizerKernel<QUANTIZE> (LinearQuantization.hpp:148-248).
    Returns (symbols int64 with outliers zeroed, outlier_idx, outlier_val)."""
    T = h.T
    quantizers = calc_quantizers(h, ebtype, tol, s, norm, True, single_dim)
    with np.errstate(all="ignore"):
        if np.isinf(s):
            x = v * quantizers[0] * T(1)
        else:
            lv = node_levels(h)
            vol = np.zeros(h.l_target + 1, dtype=T)
            for l in range(h.l_target + 1):
                p = T(1)
                for d in range(h.D - 1, -1, -1):
                    p = p * h.volume[l][d]
                vol[l] = np.sqrt(p)
            x = v * quantizers[lv] * vol[lv]
        y = np.copysign(T(0.5) + np.abs(x), v)
    q = y.astype(np.int64)  # C++ truncation toward zero
    q = q + dict_size // 2
    out = (q < 0) | (q >= dict_size)
    oidx = np.flatnonzero(out.ravel()).astype(np.uint64)
    oval = q.ravel()[oidx.astype(np.int64)].astype(np.int64)
    q = np.where(out, 0, q)
    return q, oidx, oval


def dequantize(h, q, oidx, oval, ebtype, tol, s, norm, dict_size=8192, single_dim=False):
    """OutlierRestore + LevelwiseLinearQuantizerKernel<DEQUANTIZE>
    (LinearQuantization.hpp:251-264,304-350)."""
    T = h.T
    quantizers = calc_quantizers(h, ebtype, tol, s, norm, False, single_dim)
    q = np.array(q, dtype=np.int64, copy=True).reshape(-1)
    if len(oidx):
        q[np.asarray(oidx, dtype=np.int64)] = oval
    q = q.reshape(h.shape) - dict_size // 2
    if np.isinf(s):
        return ((quantizers[0] * T(1)) * q.astype(T)).astype(T)
    lv = node_levels(h)
    vol = np.zeros(h.l_target + 1, dtype=T)
    for l in range(h.l_target + 1):
        p = T(1)
        for d in range(h.D - 1, -1, -1):
            p = p * h.volume_recip[l][d]
        vol[l] = np.sqrt(p)
    return ((quantizers[lv] * vol[lv]) * q.astype(T)).astype(T)


# --------------------------------------------------------------------------
# Huffman (Lossless/ParallelHuffman)
# --------------------------------------------------------------------------


def generate_cl(freq_sorted, oob_value=0):
    """GenerateCL.hpp:29-520 restated sequentially.  `freq_sorted`: non-zero
    frequencies in ascending order.  Returns code lengths in the same order.

    `oob_value` stands for the word the reference reads one past the end of the
    frequency array (GenerateCL.hpp:252-257 indexes histogram[lNodesCur +
    curLeavesNum] without a bound check); with glibc/SERIAL this word is 0."""
    n = len(freq_sorted)
    lfreq = [int(x) for x in freq_sorted]
    CL = [0] * n
    lleader = [-1] * n
    ifreq = [0] * n
    ileader = [-1] * n
    front = rear = lcur = isize = 0
    MOD = lambda a: a % n

    def hist(i):
        return lfreq[i] if i < n else oob_value

    while lcur < n or isize > 1:
        # Operation2: combine the two least frequent nodes
        mid = [(UINT_MAX, 0)] * 4
        if lcur < n:
            mid[0] = (lfreq[lcur], 1)
        if lcur < n - 1:
            mid[1] = (lfreq[lcur + 1], 1)
        if isize >= 1:
            mid[2] = (ifreq[front], 0)
        if isize >= 2:
            mid[3] = (ifreq[MOD(front + 1)], 0)
        for (i, j) in ((1, 3), (0, 2), (0, 1), (2, 3), (1, 2)):
            if mid[i][0] > mid[j][0]:
                mid[i], mid[j] = mid[j], mid[i]
        minfreq = mid[0][0]
        if mid[1][0] < UINT_MAX:
            minfreq += mid[1][0]
        ifreq[rear] = minfreq & UINT_MAX
        ileader[rear] = -1
        for k in (0, 1):
            if mid[k][0] < UINT_MAX:
                if mid[k][1]:
                    lleader[lcur] = rear
                    CL[lcur] += 1
                    lcur += 1
                else:
                    ileader[front] = rear
                    front = MOD(front + 1)
        isize = MOD(rear - front)
        # Operation3/4: leaves with freq <= minFreq
        cur_leaves = 0
        while lcur + cur_leaves < n and lfreq[lcur + cur_leaves] <= minfreq:
            cur_leaves += 1
        copy = [(lfreq[lcur + k], lcur + k, 1) for k in range(cur_leaves)]
        # Operation5
        merge_rear, merge_front = rear, front
        if (cur_leaves + isize) % 2 == 0:
            front = rear
        elif isize != 0 and (cur_leaves == 0 or
                             hist(lcur + cur_leaves) <= ifreq[MOD(rear - 1)]):
            merge_rear = MOD(merge_rear - 1)
            front = MOD(rear - 1)
        else:
            front = rear
            cur_leaves -= 1
        lcur += cur_leaves
        rear = MOD(rear + 1)
        blen = MOD(merge_rear - merge_front)
        temp_len = cur_leaves + blen
        if temp_len > 0:
            # Operations 6-11: merge (leaf first on ties)
            temp = []
            a, b = 0, 0
            while a < cur_leaves and b < blen:
                bi = MOD(merge_front + b)
                if copy[a][0] <= ifreq[bi]:
                    temp.append(copy[a])
                    a += 1
                else:
                    temp.append((ifreq[bi], bi, 0))
                    b += 1
            while a < cur_leaves:
                temp.append(copy[a])
                a += 1
            while b < blen:
                bi = MOD(merge_front + b)
                temp.append((ifreq[bi], bi, 0))
                b += 1
            # Operation12: meld pairs
            for i in range(temp_len // 2):
                ind = MOD(rear + i)
                ifreq[ind] = (temp[2 * i][0] + temp[2 * i + 1][0]) & UINT_MAX
                ileader[ind] = -1
                for t in (temp[2 * i], temp[2 * i + 1]):
                    if t[2]:
                        lleader[t[1]] = ind
                        CL[t[1]] += 1
                    else:
                        ileader[t[1]] = ind
            rear = MOD(rear + temp_len // 2)
        # Operation14: update leaders
        for i in range(n):
            if lleader[i] != -1 and ileader[lleader[i]] != -1:
                lleader[i] = ileader[lleader[i]]
                CL[i] += 1
        isize = MOD(rear - front)
    return CL


def generate_cw(CL_sorted_asc_freq):
    """GenerateCW.hpp:38-218.  Input: code lengths in ascending-frequency order
    (as GenerateCL leaves them).  Returns (codewords in the same order with
    len<<56, first[64], entry[64])."""
    n = len(CL_sorted_asc_freq)
    CL = list(reversed(CL_sorted_asc_freq))  # ascending length
    M64 = (1 << 64) - 1
    CW = [0] * n
    first = [M64] * 64
    entry = [M64] * 64  # workspace.reset memsets the decodebook to 0xff
    ccl = CL[0]
    cdpi = 0
    entry[ccl] = 0
    first[ccl] = (0 ^ ((1 << CL[0]) - 1)) & M64
    entry[ccl + 1] = 1
    for i in range(ccl):
        first[i] = M64
        entry[i] = 0
    while cdpi < n - 1:
        newcdpi = n - 1
        for i in range(n - 1):
            if CL[i + 1] > ccl:
                newcdpi = i
                break
        update_end = 64 if newcdpi >= n - 1 else CL[newcdpi + 1]
        cur_entry = entry[ccl]
        num_ccl = newcdpi - cdpi + 1
        CW[newcdpi] = 0 if cdpi == 0 else CW[cdpi]
        base = CW[newcdpi]
        for i in range(cdpi, newcdpi):
            CW[i] = base + (newcdpi - i)
        for i in range(ccl + 1, update_end):
            entry[i] = cur_entry + num_ccl
        if update_end < 64:
            entry[update_end] = cur_entry + num_ccl
        first[ccl] = (CW[cdpi] ^ ((1 << CL[cdpi]) - 1)) & M64
        for i in range(ccl + 1, update_end):
            first[i] = M64
        if newcdpi < n - 1:
            diff = CL[newcdpi + 1] - CL[newcdpi]
            CW[newcdpi + 1] = ((CW[cdpi] + 1) << diff) & M64
            ccl = CL[newcdpi + 1]
            newcdpi += 1
        cdpi = newcdpi
    out = [((CW[i] | ((CL[i] & 0xFF) << 56)) ^ ((1 << CL[i]) - 1)) & M64
           for i in range(n)]
    out.reverse()
    return out, first, entry


def get_codebook(freq, oob_value=0):
    """GetCodebook.hpp:23-146.  freq: uint32[dict].  Returns dict(codebook
    uint64[dict] indexed by symbol, first, entry, keys)."""
    freq = np.asarray(freq, dtype=np.uint32)
    dict_size = freq.size
    order = np.argsort(freq, kind="stable")  # ascending freq, stable on symbol
    sfreq = freq[order]
    first_nz = int(np.searchsorted(sfreq, 1, side="left"))
    nz = dict_size - first_nz
    CL = generate_cl(sfreq[first_nz:], oob_value)
    cw_nz, first, entry = generate_cw(CL)
    cb_sorted = np.zeros(dict_size, dtype=np.uint64)
    cb_sorted[first_nz:] = np.array(cw_nz, dtype=np.uint64)
    # ReverseArray(codebook), ReverseArray(qcode), ReorderByIndex
    cb_rev = cb_sorted[::-1]
    keys = order[::-1].astype(np.uint64)
    codebook = np.zeros(dict_size, dtype=np.uint64)
    codebook[keys.astype(np.int64)] = cb_rev
    return dict(codebook=codebook, first=np.array(first, dtype=np.uint64),
                entry=np.array(entry, dtype=np.uint64), keys=keys,
                cl=np.array(list(reversed(CL)), dtype=np.uint32), nz=nz)


def huffman_encode_chunks(symbols, codebook, chunk_size):
    """EncodeFixedLen + Deflate (Deflate.hpp:21-77): per chunk MSB-first bit
    packing into uint64 words; returns (bits per chunk, list of word arrays)."""
    sym = np.asarray(symbols).astype(np.int64).ravel()
    n = sym.size
    cw = codebook[sym]
    lens = (cw >> np.uint64(56)).astype(np.int64)
    codes = cw & np.uint64((1 << 56) - 1)
    nchunk = (n - 1) // chunk_size + 1
    bits = np.zeros(nchunk, dtype=np.uint64)
    words = []
    for c in range(nchunk):
        lo, hi = c * chunk_size, min(n, (c + 1) * chunk_size)
        l = lens[lo:hi]
        cd = codes[lo:hi]
        end = np.cumsum(l)
        start = end - l
        total = int(end[-1]) if l.size else 0
        bits[c] = total
        nw = (total - 1) // 64 + 1 if total > 0 else ((0 - 1) & ((1 << 64) - 1)) // 64 + 1
        w = np.zeros(nw + 1, dtype=np.uint64)
        wi = start // 64
        off = start % 64
        room = 64 - off
        fits = l <= room
        # part in first word
        sh_l = np.where(fits, room - l, 0).astype(np.uint64)
        sh_r = np.where(fits, 0, l - room).astype(np.uint64)
        part1 = np.where(fits, cd << sh_l, cd >> sh_r)
        np.bitwise_or.at(w, wi, part1)
        sp = ~fits
        if sp.any():
            rem = (l - room)[sp]
            part2 = cd[sp] << (64 - rem).astype(np.uint64)
            np.bitwise_or.at(w, wi[sp] + 1, part2)
        words.append(w[:nw])
    return bits, words


def huffman_serialize(n, dict_size, chunk_size, cb, bits, words, oidx, oval):
    """Huffman::Serialize (Huffman.hpp:130-262)."""
    nchunk = len(words)
    nwords = np.array([w.size for w in words], dtype=np.uint64)
    entry = np.zeros(nchunk, dtype=np.uint64)
    entry[1:] = np.cumsum(nwords)[:-1]
    out = bytearray()
    out += struct.pack("<Q", n)
    out += struct.pack("<i", dict_size)
    out += struct.pack("<i", chunk_size)
    out += struct.pack("<Q", 2 * nchunk)
    out += bits.astype("<u8").tobytes()
    out += entry.astype("<u8").tobytes()
    out += struct.pack("<Q", 8 * 128 + 8 * dict_size)
    out += cb["first"].astype("<u8").tobytes()
    out += cb["entry"].astype("<u8").tobytes()
    out += cb["keys"].astype("<u8").tobytes()
    out += struct.pack("<Q", int(nwords.sum()))
    for w in words:
        out += w.astype("<u8").tobytes()
    out += struct.pack("<Q", len(oidx))
    out += np.asarray(oidx, dtype="<u8").tobytes()
    out += np.asarray(oval, dtype="<i8").tobytes()
    return bytes(out)


def huffman_compress(symbols, dict_size=8192, chunk_size=20480, oidx=(), oval=(),
                     oob_value=0):
    sym = np.asarray(symbols).astype(np.int64).ravel()
    freq = np.bincount(sym, minlength=dict_size).astype(np.uint32)
    cb = get_codebook(freq, oob_value)
    bits, words = huffman_encode_chunks(sym, cb["codebook"], chunk_size)
    return huffman_serialize(sym.size, dict_size, chunk_size, cb, bits, words,
                             oidx, oval)


def huffman_parse(payload):
    """Huffman::Deserialize (Huffman.hpp:264-320)."""
    b = memoryview(bytes(payload))
    off = 0
    n, = struct.unpack_from("<Q", b, off); off += 8
    dict_size, chunk_size = struct.unpack_from("<ii", b, off); off += 8
    meta, = struct.unpack_from("<Q", b, off); off += 8
    nchunk = meta // 2
    bits = np.frombuffer(b, "<u8", nchunk, off); off += 8 * nchunk
    entry = np.frombuffer(b, "<u8", nchunk, off); off += 8 * nchunk
    dbs, = struct.unpack_from("<Q", b, off); off += 8
    first = np.frombuffer(b, "<u8", 64, off)
    ent = np.frombuffer(b, "<u8", 64, off + 512)
    keys = np.frombuffer(b, "<u8", (dbs - 1024) // 8, off + 1024)
    off += dbs
    nwords, = struct.unpack_from("<Q", b, off); off += 8
    ddata = np.frombuffer(b, "<u8", nwords, off); off += 8 * nwords
    nout, = struct.unpack_from("<Q", b, off); off += 8
    oidx = np.frombuffer(b, "<u8", nout, off); off += 8 * nout
    oval = np.frombuffer(b, "<i8", nout, off); off += 8 * nout
    return dict(n=n, dict_size=dict_size, chunk_size=chunk_size, bits=bits,
                word_offset=entry, first=first, entry=ent, keys=keys,
                ddata=ddata, oidx=oidx, oval=oval, size=off)


def huffman_decode(p):
    """DecodeKernel (Decode.hpp:66-116), bit-serial; small inputs only."""
    n, chunk = int(p["n"]), int(p["chunk_size"])
    first = [int(x) for x in p["first"]]
    entry = [int(x) for x in p["entry"]]
    keys = p["keys"]
    out = np.zeros(n, dtype=np.uint64)
    dd = p["ddata"]
    for c in range(len(p["bits"])):
        total = int(p["bits"][c])
        base = int(p["word_offset"][c])
        nw = (total - 1) // 64 + 1
        by = dd[base: base + nw].astype(">u8").tobytes()
        bitarr = np.unpackbits(np.frombuffer(by, dtype=np.uint8))
        i, o = 0, c * chunk
        while i < total:
            v, l = int(bitarr[i]), 1
            while v < first[l]:
                i += 1
                v = (v << 1) | int(bitarr[i])
                l += 1
            out[o] = keys[entry[l] + v - first[l]]
            o += 1
            i += 1
    return out


# --------------------------------------------------------------------------
# Low-level Compressor (CompressionLowLevel/Compressor.hpp:193-272)
# --------------------------------------------------------------------------


def compress_lowlevel(h, u, ebtype, tol, s, norm=None, dict_size=8192,
                      chunk_size=20480, oob_value=0, reorder=0, single_dim=False):
    T = h.T
    if ebtype == REL and norm is None:
        norm = calc_norm(np.asarray(u, dtype=T), s)
    if norm is None:
        norm = T(1)
    v = decompose_single(h, u) if single_dim else decompose(h, u)
    q, oidx, oval = quantize(h, v, ebtype, tol, s, norm, dict_size, single_dim)
    if reorder:
        q, oidx, oval = linearize(h, q, oidx, oval)
    payload = huffman_compress(q, dict_size, chunk_size, oidx, oval, oob_value)
    return dict(payload=payload, norm=norm, decomposed=v, quantized=q,
                oidx=oidx, oval=oval)


def decompress_lowlevel(h, payload, ebtype, tol, s, norm, reorder=0, single_dim=False):
    p = huffman_parse(payload)
    sym = huffman_decode(p).astype(np.int64)
    if reorder:  # outliers are indexed in the linearised order: restore them first
        if len(p["oidx"]):
            sym[np.asarray(p["oidx"], dtype=np.int64)] = p["oval"]
        sym = delinearize(h, sym)
        v = dequantize(h, sym, (), (), ebtype, tol, s, norm, int(p["dict_size"]), single_dim)
        return recompose_single(h, v) if single_dim else recompose(h, v)
    v = dequantize(h, sym, p["oidx"], p["oval"], ebtype, tol, s, norm,
                   int(p["dict_size"]), single_dim)
    if single_dim:
        return recompose_single(h, v)
    return recompose(h, v)


# --------------------------------------------------------------------------
# Stream framing: preamble + protobuf header (Metadata.cpp:249-462, mgard.proto)
# --------------------------------------------------------------------------

MGARD_FILE_VERSION = (1, 0, 0)  # CMakeLists.txt:17-19


def _varint(x):
    out = bytearray()
    x &= (1 << 64) - 1
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
    bits = struct.pack("<d", x)
    return b"" if bits == b"\0" * 8 else _varint((field << 3) | 1) + bits


def _f_msg(field, body):
    return _varint((field << 3) | 2) + _varint(len(body)) + body


def encode_header(shape, dtype, ebtype, tol, s, norm, coords=None,
                  decomposed=False, dd_dim=0, dd_size=0, dict_size=8192,
                  chunk_size=20480, backend=3, lossless=3, reorder=0, dd_method=1):
    """proto3 canonical bytes of mgard.pb.Header as MetadataBase::Serialize
    fills it (including the version-field quirk, Metadata.cpp:267-271).
    backend: Device.Backend (1 X_SERIAL, 3 X_CUDA); lossless: Encoding.Compressor."""
    ver = (_f_varint(1, MGARD_FILE_VERSION[0]) + _f_varint(2, MGARD_FILE_VERSION[1])
           + _f_varint(3, MGARD_FILE_VERSION[2]))
    topo = _f_varint(1, len(shape)) + _f_msg(
        2, b"".join(_varint(int(x)) for x in shape))
    dom = _f_msg(2, topo)
    if coords is not None:
        dom += _f_varint(3, 1)
        flat = np.concatenate([np.asarray(c, dtype=np.float64) for c in coords])
        dom += _f_msg(4, _f_msg(2, flat.astype("<f8").tobytes()))
    dataset = _f_varint(1, 0 if np.dtype(dtype) == np.float32 else 1) + _f_varint(2, 1)
    err = b""
    if ebtype == REL:
        err += _f_varint(1, 1)
    if not np.isinf(s):
        err += _f_varint(2, 1)
    err += _f_double(3, float(s))
    if ebtype == REL:
        err += _f_double(4, float(norm))
    err += _f_double(5, float(tol))
    if not decomposed:
        # DomainDecomposer.hpp:333-337: dim 0 / size shape[0] when not decomposed
        dd_dim, dd_size = 0, int(shape[0])
    # DomainDecomposition.Method: 1 MAX_DIMENSION, 2 BLOCK, 3 VARIABLE (Metadata.cpp:335-354)
    dd = (_f_varint(1, dd_method if decomposed else 0) + _f_varint(2, dd_dim)
          + _f_varint(3, dd_size))
    fd = _f_varint(2, 1)
    quant = _f_varint(1, 1) + _f_varint(3, 3)
    enc = (_f_varint(1, 1 if reorder else 0) + _f_varint(2, lossless)
           + _f_varint(3, dict_size) + _f_varint(4, chunk_size))
    dev = _f_varint(1, backend)
    return (_f_msg(2, ver) + _f_msg(3, b"") + _f_msg(4, dom) + _f_msg(5, dataset)
            + _f_msg(6, err) + _f_msg(7, dd) + _f_msg(8, fd) + _f_msg(9, quant)
            + _f_msg(10, b"") + _f_msg(11, enc) + _f_msg(12, dev))


def encode_preamble(header):
    """Metadata.cpp:441-459: 'MGARD' | u64 LE size | u32 LE crc32 | header."""
    return (b"MGARD" + struct.pack("<Q", len(header))
            + struct.pack("<I", zlib.crc32(header) & 0xFFFFFFFF) + header)


def compress(u, ebtype, tol, s, coords=None, dict_size=8192, chunk_size=20480,
             backend=3, oob_value=0):
    """mgard_x::compress, single sub-domain (CompressionHighLevel.hpp:49-314 +
    GPUPipelines.hpp:3-207): metadata | u64 size | payload (raw if CR < 1)."""
    u = np.ascontiguousarray(u)
    h = Hierarchy(u.shape, u.dtype, coords)
    r = compress_lowlevel(h, u, ebtype, tol, s, None, dict_size, chunk_size,
                          oob_value)
    payload = r["payload"]
    if u.nbytes / len(payload) < 1.0:
        payload = u.tobytes()
    norm = r["norm"] if ebtype == REL else u.dtype.type(1)
    hdr = encode_header(u.shape, u.dtype, ebtype, tol, s, norm, coords,
                        dict_size=dict_size, chunk_size=chunk_size, backend=backend)
    return encode_preamble(hdr) + struct.pack("<Q", len(payload)) + payload
