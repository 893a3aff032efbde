"""CPU tests of the __host__ __device__ numerics of the CUDA path (compiled for the host in
libekb200_hostcheck.so) against LAPACK: the secular-equation solver vs dlaed4."""
import ctypes
import os

import numpy as np
import pytest

from oracle import lapack_twin as lt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HC = os.path.join(ROOT, "eigenkernel_b200", "libekb200_hostcheck.so")
dp = ctypes.POINTER(ctypes.c_double)
ip = ctypes.POINTER(ctypes.c_int)


@pytest.fixture(scope="module")
def hc():
    if not os.path.exists(HC):
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(HC)


def ours(hc, d, z, rho):
    k = len(d)
    lam, tau = np.zeros(k), np.zeros(k)
    orig, it = np.zeros(k, dtype=np.int32), np.zeros(k, dtype=np.int32)
    hc.ekb200_host_secular(k, d.ctypes.data_as(dp), z.ctypes.data_as(dp), ctypes.c_double(rho),
                           lam.ctypes.data_as(dp), orig.ctypes.data_as(ip), tau.ctypes.data_as(dp), it.ctypes.data_as(ip))
    return lam, orig, tau, it


def dlaed4(d, z, rho):
    k = len(d)
    lam, delta = np.zeros(k), np.zeros(k)
    f = lt._LIB.scipy_dlaed4_
    for i in range(1, k + 1):
        dl, info = ctypes.c_double(), ctypes.c_int()
        f(ctypes.byref(ctypes.c_int(k)), ctypes.byref(ctypes.c_int(i)), d.ctypes.data_as(dp), z.ctypes.data_as(dp),
          delta.ctypes.data_as(dp), ctypes.byref(ctypes.c_double(rho)), ctypes.byref(dl), ctypes.byref(info))
        assert info.value == 0
        lam[i - 1] = dl.value
    return lam


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_secular_roots_match_dlaed4(hc, kind):
    rng = np.random.default_rng(kind)
    for trial in range(60):
        k = int(rng.integers(2, 200))
        if kind == 0:
            d = np.sort(rng.standard_normal(k))
        elif kind == 1:  # near-degenerate pole pairs, tiny weights
            d = np.sort(rng.standard_normal(k))
            d[1::2] = d[0::2][: len(d[1::2])] + 1e-9 * rng.random(len(d[1::2]))
            d = np.sort(d)
        elif kind == 2:  # gaps over 12 orders of magnitude
            d = np.cumsum(10.0 ** rng.uniform(-12, 0, k))
        else:
            d = np.sort(rng.standard_normal(k)) * 1e-3
        z = rng.standard_normal(k)
        if kind == 1:
            z *= 10.0 ** rng.uniform(-7, 0, k)
        z /= np.linalg.norm(z)
        rho = float(10.0 ** rng.uniform(-6, 1))
        lam, orig, tau, it = ours(hc, d, z, rho)
        ref = dlaed4(d, z, rho)
        scale = max(np.abs(d).max(), rho)
        assert np.max(np.abs(lam - ref)) <= 1e-14 * scale
        assert it.max() < 80
        # interlacing and origin = nearer pole: strict in the (origin, tau) representation the kernels use
        # (d_i - lambda_j is formed as (d_i - d_origin) - tau); the rounded lam = d_origin + tau may tie with
        # a pole when the weight is tiny (such poles are deflated before the secular solve in stedc.cu)
        j = np.arange(k - 1)
        assert np.all(((orig[:-1] == j) & (tau[:-1] > 0)) | ((orig[:-1] == j + 1) & (tau[:-1] < 0)))
        assert orig[-1] == k - 1 and tau[-1] > 0
        assert np.all(lam[:-1] >= d[:-1]) and np.all(lam[:-1] <= d[1:]) and lam[-1] >= d[-1]
        assert np.all(np.abs(tau[:-1]) <= 0.5 * np.diff(d) * (1 + 1e-12))


def test_secular_single_pole(hc):
    lam, orig, tau, it = ours(hc, np.array([0.5]), np.array([1.0]), 2.0)
    assert lam[0] == 2.5 and orig[0] == 0
