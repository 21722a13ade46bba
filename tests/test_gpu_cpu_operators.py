"""The reference's own known answers for the constituent operators of the MGARD-CPU
transform, reproduced on the GPU through mgb_cpu_apply_operator:

  ConstituentMassMatrix             tests/src/test_TensorMassMatrix.cpp:21-207
  ConstituentMassMatrixInverse      tests/src/test_TensorMassMatrix.cpp:304-387
  ConstituentRestriction            tests/src/test_TensorRestriction.cpp:18-143
  ConstituentProlongationAddition   tests/src/test_TensorProlongation.cpp:16-106

The reference applies an operator to ONE line of a shuffled array; the entry point
applies it to every line of the level along the dimension, so for the 2-D cases only the
nodes of the lines the reference touched are compared.  Input vectors and expected values
are the reference tests' (data, not code); comparison as there (Catch::Approx: relative
1.2e-5 of the value)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    import mgard_b200.cpu as mc
    assert torch.cuda.is_available()
    return torch, mc, torch.device("cuda:0")


def approx(a, b, margin=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= np.maximum(margin, 1.2e-5 * np.abs(b)) + 1e-30 * (margin == 0))


def apply(torch, d, H, op, l, dim, u):
    t = torch.from_numpy(np.ascontiguousarray(u, dtype=H.dtype)).to(d)
    return H.apply_operator(op, l, dim, t).cpu().numpy()


def test_mass_matrix_1d(env):
    torch, mc, d = env
    H = mc.TensorMeshHierarchy((9,), None, np.float32)
    u = np.array([-2, 1, 1, 8, 3, -2, -7, -4, 0], dtype=np.float32)
    got = apply(torch, d, H, H.MASS, 3, 0, u)
    assert approx(got, np.array([-3, 3, 13, 36, 18, -12, -34, -23, -4]) / 48)
    want = np.array([-1, 1, 1, 8, 10, -2, -7, -4, 3], dtype=np.float64)
    want[[0, 4, 8]] /= 12
    assert approx(apply(torch, d, H, H.MASS, 1, 0, u), want)
    # nondyadic
    H = mc.TensorMeshHierarchy((7,), None, np.float32)
    u = np.array([-1, 8, -9, -9, -1, -10, 6], dtype=np.float32)
    wants = [[2. / 3, 8, -9, -9, -1, -10, 11. / 6],
             [-11. / 12, 8, -9, -31. / 12, -1, -10, 0.25],
             [1. / 6, 29. / 36, -9, -39. / 36, -1. / 12, -10, 11. / 18],
             [1. / 6, 11. / 18, -37. / 36, -23. / 18, -23. / 36, -35. / 36, 1. / 18]]
    for l, want in enumerate(wants):
        assert approx(apply(torch, d, H, H.MASS, l, 0, u), want), l


def test_mass_matrix_2d_custom_spacing(env):
    torch, mc, d = env
    H = mc.TensorMeshHierarchy((5, 5), [[0, 0.1, 0.5, 0.75, 1], [0, 0.65, 0.70, 0.75, 1]], np.float64)
    u = np.array([8, -3, 1, 5, 10, 9, -10, -3, 8, 10, -3, 6, -7, -3, 3, 3, -9, 0, -1, 8, -6, 7, 1, -2, 10],
                 dtype=np.float64).reshape(5, 5)
    # l = 2 along dimension 0, lines at columns 0 and 3
    got = apply(torch, d, H, H.MASS, 2, 0, u)
    col0 = [2.5 / 6, 2.6 / 6 + 6.0 / 6, 1.2 / 6 + -0.75 / 6, 0.75 / 6 + 0.0 / 6, -2.25 / 6]
    col3 = [1.8 / 6, 2.1 / 6 + 5.2 / 6, 0.8 / 6 + -1.75 / 6, -1.25 / 6 + -1.0 / 6, -1.25 / 6]
    assert approx(got[:, 0], col0, 1e-15) and approx(got[:, 3], col3, 1e-15)
    # l = 2 along dimension 1, rows 1 and 2
    got = apply(torch, d, H, H.MASS, 2, 1, u)
    row1 = [5.2 / 6, -7.15 / 6 + -1.15 / 6, -0.8 / 6 + 0.1 / 6, 0.65 / 6 + 6.5 / 6, 7.0 / 6]
    row2 = [0.0 / 6, 5.85 / 6 + 0.25 / 6, -0.4 / 6 + -0.85 / 6, -0.65 / 6 + -0.75 / 6, 0.75 / 6]
    assert approx(got[1], row1, 1e-15) and approx(got[2], row2, 1e-15)
    # l = 1 along dimension 1, row 4 (level-1 nodes: columns 0, 2, 4); the others keep their values
    got = apply(torch, d, H, H.MASS, 1, 1, u)
    assert approx(got[4], [-7.7 / 6, 7, -2.8 / 6 + 3.6 / 6, -2, 6.3 / 6], 1e-15)
    assert np.array_equal(got[:, [1, 3]], u[:, [1, 3]])


def test_mass_matrix_inverse(env):
    torch, mc, d = env
    H = mc.TensorMeshHierarchy((9,), None, np.float32)
    u = np.array([-1, -1, 8, 2, 4, 3, -1, -4, 3], dtype=np.float32)
    v = apply(torch, d, H, H.MASS_INVERSE, 3, 0, apply(torch, d, H, H.MASS, 3, 0, u))
    assert approx(v, u)
    v = apply(torch, d, H, H.MASS, 1, 0, apply(torch, d, H, H.MASS_INVERSE, 1, 0, u))
    assert approx(v, u)
    # exhaustive_constituent_inverse_test on the reference's 9 x 17 custom-spacing mesh:
    # M then M^-1 along each dimension at each level reproduces the input
    xs = [0.469, 1.207, 1.918, 2.265, 2.499, 2.525, 2.879, 3.109, 3.713]
    ys = [0.137, 0.907, 1.363, 1.856, 2.188, 3.008, 3.643, 4.580, 5.320, 5.464, 6.223, 6.856, 7.083, 7.459, 7.748,
          8.641, 8.740]
    H = mc.TensorMeshHierarchy((9, 17), [xs, ys], np.float64)
    rng = np.random.default_rng(731617)
    u = rng.integers(-10, 11, (9, 17)).astype(np.float64)
    for l in range(H.L + 1):
        for dim in (0, 1):
            v = apply(torch, d, H, H.MASS_INVERSE, l, dim, apply(torch, d, H, H.MASS, l, dim, u))
            assert np.abs(v - u).max() <= 1e-9 * np.abs(u).max(), (l, dim)
    # 3-D nondyadic, fp32
    H = mc.TensorMeshHierarchy((9, 8, 7), None, np.float32)
    u = rng.uniform(-5, -3, (9, 8, 7)).astype(np.float32)
    for l in range(H.L + 1):
        for dim in (0, 1, 2):
            v = apply(torch, d, H, H.MASS_INVERSE, l, dim, apply(torch, d, H, H.MASS, l, dim, u))
            assert np.abs(v - u).max() <= 2e-5 * np.abs(u).max(), (l, dim)


def test_restriction(env):
    torch, mc, d = env
    H = mc.TensorMeshHierarchy((9,), None, np.float32)
    u = np.array([9, 2, 4, -4, 7, 5, -2, 5, 6], dtype=np.float32)
    wants = {3: [10, 2, 3, -4, 7.5, 5, 3, 5, 8.5], 2: [11, 2, 4, -4, 8, 5, -2, 5, 5], 1: [12.5, 2, 4, -4, 7, 5, -2, 5, 9.5]}
    for l, want in wants.items():
        assert np.array_equal(apply(torch, d, H, H.RESTRICTION, l, 0, u), np.array(want, dtype=np.float32)), l
    # custom spacing, nondyadic
    H = mc.TensorMeshHierarchy((4,), [[0.0, 0.1, 0.9, 1.0]], np.float64)
    u = np.array([5, 2, 2, 4], dtype=np.float64)
    assert approx(apply(torch, d, H, H.RESTRICTION, 2, 0, u), [5, 20. / 9, 2, 52. / 9])
    assert approx(apply(torch, d, H, H.RESTRICTION, 1, 0, u), [6.8, 2, 2, 4.2])
    # 2-D custom spacing: all lines at once = the reference's per-line results put together
    H = mc.TensorMeshHierarchy((3, 3), [[0, 0.75, 1], [0, 0.25, 1]], np.float64)
    u = np.array([-9, -5, 1, 9, 3, 2, 9, 5, 6], dtype=np.float64).reshape(3, 3)
    got = apply(torch, d, H, H.RESTRICTION, 1, 0, u)
    assert approx(got[:, 0], [-6.75, 9, 15.75]) and approx(got[:, 1], [-4.25, 3, 7.25])
    got = apply(torch, d, H, H.RESTRICTION, 1, 1, u)
    assert approx(got[1], [11.25, 3, 2.75]) and approx(got[2], [12.75, 5, 7.25])
    # the reference's constructors throw for level 0
    with pytest.raises(Exception):
        apply(torch, d, H, H.RESTRICTION, 0, 0, u)


def test_prolongation_addition(env):
    torch, mc, d = env
    H = mc.TensorMeshHierarchy((9,), None, np.float32)
    u = np.array([-10, 10, -2, 3, 9, -8, 9, -8, 4], dtype=np.float32)
    wants = {3: [-10, 4, -2, 6.5, 9, 1, 9, -1.5, 4], 2: [-10, 10, -2.5, 3, 9, -8, 15.5, -8, 4],
             1: [-10, 10, -2, 3, 6, -8, 9, -8, 4]}
    for l, want in wants.items():
        assert np.array_equal(apply(torch, d, H, H.PROLONGATION_ADDITION, l, 0, u), np.array(want, dtype=np.float32)), l
    H = mc.TensorMeshHierarchy((3, 3), [[0, 0.9, 1], [1, 1.4, 2]], np.float64)
    u = np.array([4, -1, 4, 7, 3, -3, -4, 1, -7], dtype=np.float64).reshape(3, 3)
    got = apply(torch, d, H, H.PROLONGATION_ADDITION, 1, 0, u)
    assert approx(got[:, 0], [4, 3.8, -4]) and approx(got[:, 1], [-1, 3.8, 1])
    got = apply(torch, d, H, H.PROLONGATION_ADDITION, 1, 1, u)
    assert approx(got[1], [7, 6.0, -3]) and approx(got[2], [-4, -4.2, -7])
    with pytest.raises(Exception):
        apply(torch, d, H, H.PROLONGATION_ADDITION, 0, 0, u)
