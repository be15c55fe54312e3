"""CPU tests of the C++ host side above the C ABI (no GPU needed): NormalLinearSystem::solve / reduce_system and
the pseudo-inverse against the NumPy oracle (which restates normal_linear_system.cpp:10-59 and
eigen_photometric_bundle_adjustment.cpp:31-45)."""
import numpy as np
import pytest

from dsopp_b200 import host
from oracle import pba_oracle as O


def spd(n, seed, scale=1e3):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(3 * n, n))
    return A.T @ A * scale, rng.normal(size=n) * scale


@pytest.mark.parametrize("n", [16, 24, 64, 128])
def test_normal_solve(n):
    H, b = spd(n, n)
    H[:8, :8] += np.eye(8) * 1e16  # fixed-frame prior
    H[14, 14] += 1e12  # affine prior
    x = host.normal_solve(H, b)
    ref = O.normal_solve(H, b)
    assert np.allclose(x, ref, rtol=1e-7, atol=1e-12 * np.abs(ref).max())
    assert np.abs(H @ x - b)[8:].max() <= 1e-6 * np.abs(b).max()


@pytest.mark.parametrize("n,elim", [(24, range(0, 8)), (64, range(8, 16)), (32, list(range(0, 8)) + list(range(24, 32)))])
def test_reduce_system(n, elim):
    H, b = spd(n, 100 + n)
    Hr, br = host.reduce_system(H, b, list(elim))
    Hr_ref, br_ref = O.reduce_system(H, b, list(elim))
    assert Hr.shape == Hr_ref.shape
    assert np.allclose(Hr, Hr_ref, rtol=1e-8, atol=1e-9 * np.abs(Hr_ref).max())
    assert np.allclose(br, br_ref, rtol=1e-8, atol=1e-9 * np.abs(br_ref).max())


def test_reduce_system_with_rank_deficient_block():
    H, b = spd(24, 7)
    H[3, :] = 0
    H[:, 3] = 0  # an unobservable direction inside the eliminated block
    Hr, br = host.reduce_system(H, b, list(range(8)))
    Hr_ref, br_ref = O.reduce_system(H, b, list(range(8)))
    assert np.allclose(Hr, Hr_ref, rtol=1e-7, atol=1e-8 * np.abs(Hr_ref).max())
    assert np.allclose(br, br_ref, rtol=1e-7, atol=1e-8 * np.abs(br_ref).max())


@pytest.mark.parametrize("n_null", [0, 1])
def test_pseudo_inverse(n_null):
    H, _ = spd(40, 3)
    P = host.sym_pinv(H, n_null)
    ref = O.pseudo_inverse(H, n_null)
    assert np.allclose(P, ref, rtol=1e-7, atol=1e-9 * np.abs(ref).max())


def test_radix_select_is_nth_element():
    """The selection rule of csrc/energy_quantile.cu restated in NumPy (order-preserving integer image of the floats,
    four most-significant-byte-first passes narrowing (prefix, k)) returns the element std::nth_element returns at
    k = size_t(n * 0.75)  (photometric_bundle_adjustment.cpp:358-361)."""
    rng = np.random.default_rng(0)

    def select(x, frac=0.75):
        u = x.astype(np.float32).view(np.uint32)
        key = np.where(u & 0x80000000, ~u, u | 0x80000000).astype(np.uint32)
        n = len(key)
        k = int(float(n) * frac)
        prefix, mask = np.uint32(0), np.uint32(0)
        for shift in (24, 16, 8, 0):
            sel = key[(key & mask) == prefix]
            hist = np.bincount((sel >> np.uint32(shift)) & np.uint32(255), minlength=256)
            below, b = 0, 0
            while b < 255 and below + hist[b] <= k:
                below += hist[b]
                b += 1
            k -= below
            prefix |= np.uint32(b << shift)
            mask |= np.uint32(255 << shift)
        back = np.uint32(prefix & 0x7FFFFFFF) if prefix & 0x80000000 else np.uint32(~prefix)
        return np.array([back], dtype=np.uint32).view(np.float32)[0]

    for n in (1, 2, 3, 7, 1000, 4097):
        for x in (rng.exponential(50.0, n), rng.normal(0.0, 3.0, n), np.full(n, 2.5), np.round(rng.uniform(0, 4, n))):
            x = x.astype(np.float32)
            k = int(float(n) * 0.75)
            assert select(x) == np.partition(x, k)[k]
