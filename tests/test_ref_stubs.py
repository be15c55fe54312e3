"""The third-party STAND-INS under oracle/ref_stubs_full/ held against independent implementations.

The oracle is pinned against the reference's own code compiled over minimal stand-ins of Eigen and Sophus
(oracle/build_ref_pba.py).  That pin is only as good as the stand-ins: here their LDL^T solve, pseudo-inverse, SE3
exponential / adjoint / inverse / product and a handful of the block / array / colwise semantics the reference relies on
are compared with NumPy and SciPy.  (Needs the library: built from /root/reference here, shipped prebuilt to the GPU box.)
"""
import ctypes as C

import numpy as np
import pytest

from oracle import ref_pba

pytestmark = pytest.mark.skipif(not ref_pba.available(), reason="neither /root/reference nor oracle/_ref is present")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def test_ldlt_solve_matches_numpy():
    lib = ref_pba.load()
    rng = np.random.default_rng(0)
    for n in (3, 8, 24, 64):
        A = rng.normal(size=(n + 3, n)) * np.logspace(0, 4, n)[None, :]
        H = np.ascontiguousarray(A.T @ A + np.eye(n) * 1e-3)
        b = rng.normal(size=n)
        x = np.zeros(n)
        lib.refstub_ldlt_solve(n, _ptr(H), _ptr(b), _ptr(x))
        ref = np.linalg.solve(H, b)
        assert np.abs(x - ref).max() <= 1e-8 * np.abs(ref).max() + 1e-12, n
    # indefinite but non-singular (LDL^T, not Cholesky): pivoting must cope
    H = np.array([[0.0, 2.0, 1.0], [2.0, 1.0, 0.5], [1.0, 0.5, -3.0]])
    b = np.array([1.0, -2.0, 0.5])
    x = np.zeros(3)
    lib.refstub_ldlt_solve(3, _ptr(H), _ptr(b), _ptr(x))
    assert np.abs(H @ x - b).max() <= 1e-12


def test_pseudo_inverse_matches_numpy():
    lib = ref_pba.load()
    rng = np.random.default_rng(1)
    for m, n, rank in ((8, 8, 8), (8, 8, 5), (16, 8, 8), (6, 10, 6), (8, 8, 1)):
        A = np.ascontiguousarray(rng.normal(size=(m, rank)) @ rng.normal(size=(rank, n)))
        P = np.zeros((n, m))
        lib.refstub_pseudo_inverse(m, n, _ptr(A), _ptr(P))
        ref = np.linalg.pinv(A, rcond=1e-12)
        assert np.abs(P - ref).max() <= 1e-9 * max(np.abs(ref).max(), 1.0), (m, n, rank)


def test_se3_matches_the_matrix_exponential():
    scipy_linalg = pytest.importorskip("scipy.linalg")
    lib = ref_pba.load()
    rng = np.random.default_rng(2)

    def hat6(xi):  # Sophus tangent order: translation first, then rotation
        v, w = xi[:3], xi[3:]
        M = np.zeros((4, 4))
        M[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
        M[:3, 3] = v
        return M

    for scale in (1e-12, 1e-6, 1e-2, 1.0, 3.0):
        a, b = rng.normal(size=6) * scale, rng.normal(size=6) * scale
        T, Adj, Ti, Pr = np.zeros((3, 4)), np.zeros((6, 6)), np.zeros((3, 4)), np.zeros((3, 4))
        lib.refstub_se3(_ptr(a), _ptr(b), _ptr(T), _ptr(Adj), _ptr(Ti), _ptr(Pr))
        Ta, Tb = scipy_linalg.expm(hat6(a)), scipy_linalg.expm(hat6(b))
        assert np.abs(T - Ta[:3]).max() <= 1e-12 * max(1.0, np.abs(Ta).max())
        assert np.abs(Ti - np.linalg.inv(Ta)[:3]).max() <= 1e-12 * max(1.0, np.abs(Ta).max())
        assert np.abs(Pr - (Ta @ Tb)[:3]).max() <= 1e-11 * max(1.0, np.abs(Ta @ Tb).max())
        # Adj is defined by  T exp(xi^) T^-1 = exp((Adj xi)^)
        R, t = Ta[:3, :3], Ta[:3, 3]
        th = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
        want = np.block([[R, th @ R], [np.zeros((3, 3)), R]])
        assert np.abs(Adj - want).max() <= 1e-12 * max(1.0, np.abs(want).max())
        xi = rng.normal(size=6) * 1e-3
        lhs = Ta @ scipy_linalg.expm(hat6(xi)) @ np.linalg.inv(Ta)
        rhs = scipy_linalg.expm(hat6(Adj @ xi))
        assert np.abs(lhs - rhs).max() <= 1e-10 * max(1.0, np.abs(lhs).max())


def test_eager_eigen_semantics():
    lib = ref_pba.load()
    M = np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 7.0], [2.0, 0.5, 4.0]])
    out = np.zeros(32)
    lib.refstub_eigen_semantics(_ptr(np.ascontiguousarray(M)), _ptr(out))
    assert np.allclose(out[0:6].reshape(2, 3), M[1:3] * 2 - M[0:2])
    assert np.allclose(out[6:12].reshape(2, 3), M[:2] / M[2:3])                       # colwise().hnormalized()
    assert np.allclose(out[12:21].reshape(3, 3), M + (M[:, 2] + 0.5 * M[:, 2])[:, None])  # colwise() += vector
    assert out[21] == 1.0
    assert np.isclose(out[22], np.trace(M.T @ M))
    assert out[23] == M[2, 0]                                                         # selfadjointView<Lower>
    assert out[24] == M[1, 2]                                                         # row -> column assignment
    assert np.isclose(out[25], (M[:, 0] * M[:, 1]).sum())
    assert np.isclose(out[26], 1.0 / np.sqrt(M[1, 1]))
