"""CPU tests of the bisection + inverse-iteration numerics of the CUDA path (eigenkernel_b200/csrc/tridiag.cuh, compiled
for the host in libekb200_hostcheck.so: the SAME functions the kernels of stebz.cu execute, including the per-cluster
orchestration) against LAPACK's dstebz / dstein and dsteqr via SciPy.

Bars: eigenvalues within a few ulp of |T|; eigenvectors with residual |T x - lambda x| <= 1e-13 |T| sqrt(n)-ish,
orthogonality |X^T X - I|_F <= 1e-12 n (BASELINE.json's tolerance)."""
import ctypes
import os

import numpy as np
import pytest
import scipy.linalg as sla

from eigenkernel_b200 import app_io
from oracle import lapack_twin as lt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HC = os.path.join(ROOT, "eigenkernel_b200", "libekb200_hostcheck.so")
dp = ctypes.POINTER(ctypes.c_double)
ll = ctypes.c_longlong


@pytest.fixture(scope="module")
def hc():
    if not os.path.exists(HC):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(HC)
    lib.ekb200_host_stebz.argtypes = [ll, dp, dp, dp, ctypes.POINTER(ctypes.c_int)]
    lib.ekb200_host_stebz_k.argtypes = [ll, dp, dp, ctypes.c_int, dp, ctypes.POINTER(ctypes.c_int)]
    lib.ekb200_host_stein_lanes.argtypes = [ll, dp, dp, ll, dp, ctypes.c_int, dp, ll]
    lib.ekb200_host_stein_slab.argtypes = [ll, dp, dp, ll, dp, ll, ll, dp, ll]
    lib.ekb200_host_stein.argtypes = [ll, dp, dp, ll, dp, dp, ll, ctypes.POINTER(ll), ctypes.POINTER(ll)]
    return lib


def stebz(hc, d, e):
    n = len(d)
    w = np.zeros(n)
    it = ctypes.c_int()
    ee = np.ascontiguousarray(e if n > 1 else np.zeros(1))
    hc.ekb200_host_stebz(n, d.ctypes.data_as(dp), ee.ctypes.data_as(dp), w.ctypes.data_as(dp), ctypes.byref(it))
    return w, it.value


def stein(hc, d, e, w, nev):
    n = len(d)
    Z = np.zeros((n, nev), order="F")
    nc, mc = ll(), ll()
    ee = np.ascontiguousarray(e if n > 1 else np.zeros(1))
    fail = hc.ekb200_host_stein(n, d.ctypes.data_as(dp), ee.ctypes.data_as(dp), nev, w.ctypes.data_as(dp),
                                Z.ctypes.data_as(dp), n, ctypes.byref(nc), ctypes.byref(mc))
    return Z, fail, nc.value, mc.value


def tmul(d, e, X):
    Y = d[:, None] * X
    Y[:-1] += e[:, None] * X[1:]
    Y[1:] += e[:, None] * X[:-1]
    return Y


def check(d, e, w, Z, tol_orth=None):
    n, k = Z.shape
    tn = max(np.abs(d).max() + 2 * (np.abs(e).max() if len(e) else 0.0), 1e-300)
    R = tmul(d, e, Z) - Z * w[None, :k]
    assert np.linalg.norm(R, axis=0).max() <= 2e-13 * tn * np.sqrt(n), np.linalg.norm(R, axis=0).max() / tn
    G = Z.T @ Z - np.eye(k)
    assert np.linalg.norm(G, "fro") <= (tol_orth or 1e-12 * n), np.linalg.norm(G, "fro")


def cases():
    rng = np.random.default_rng(11)
    out = {}
    n = 600
    out["random"] = (rng.standard_normal(n), rng.standard_normal(n - 1))
    out["toeplitz121"] = (2.0 * np.ones(1500), -np.ones(1499))
    m = 10
    out["wilkinson21"] = (np.abs(np.arange(-m, m + 1)).astype(float), np.ones(2 * m))
    # two identical blocks glued by a zero: exactly double eigenvalues
    db, eb = rng.standard_normal(150), rng.standard_normal(149)
    out["glued_exact_double"] = (np.concatenate([db, db]), np.concatenate([eb, [0.0], eb]))
    # ... and by a tiny coupling: pairs split by ~1e-12
    out["glued_tiny_coupling"] = (np.concatenate([db, db]), np.concatenate([eb, [1e-11], eb]))
    # graded over 12 orders of magnitude
    g = 10.0 ** np.linspace(0, -12, 200)
    out["graded"] = (g, 0.3 * np.sqrt(g[:-1] * g[1:]))
    out["scaled_small"] = (1e-150 * rng.standard_normal(100), 1e-150 * rng.standard_normal(99))
    out["scaled_big"] = (1e150 * rng.standard_normal(100), 1e150 * rng.standard_normal(99))
    out["n1"] = (np.array([3.5]), np.zeros(0))
    out["n2"] = (np.array([1.0, -1.0]), np.array([0.5]))
    out["diagonal"] = (np.array([3.0, 1.0, 2.0, 1.0, 5.0]), np.zeros(4))
    return out


@pytest.mark.parametrize("name", list(cases()))
def test_bisection_matches_lapack_eigenvalues(hc, name):
    d, e = cases()[name]
    w, it = stebz(hc, d, e)
    ref = sla.eigvalsh_tridiagonal(d, e, lapack_driver="stebz") if len(d) > 1 else d.copy()
    tn = np.abs(d).max() + 2 * (np.abs(e).max() if len(e) else 0.0)
    assert np.all(np.diff(w) >= 0)
    assert np.max(np.abs(w - ref)) <= 8 * np.finfo(float).eps * tn * max(1.0, np.log2(len(d) + 1))
    assert it <= 128


@pytest.mark.parametrize("sections", [3, 7])
@pytest.mark.parametrize("name", list(cases()))
def test_multisection_agrees_with_plain_bisection(hc, name, sections):
    """The CUDA kernel cuts the bracket by 4 or 8 per sweep (3 / 7 interleaved Sturm chains); the counts are the same
    function, so the final brackets overlap to dstebz's tolerance and far fewer sweeps are needed."""
    d, e = cases()[name]
    n = len(d)
    w1, it1 = stebz(hc, d, e)
    wk = np.zeros(n)
    it = ctypes.c_int()
    ee = np.ascontiguousarray(e if n > 1 else np.zeros(1))
    hc.ekb200_host_stebz_k(n, d.ctypes.data_as(dp), ee.ctypes.data_as(dp), sections, wk.ctypes.data_as(dp), ctypes.byref(it))
    assert np.all(np.diff(wk) >= 0)
    assert np.all(np.abs(wk - w1) <= 4 * np.finfo(float).eps * np.maximum(np.abs(w1), np.abs(wk)) + 1e-300 +
                  4 * np.finfo(float).tiny * max(1.0, float(np.max(e ** 2)) if len(e) else 1.0))
    if n > 2 and it1 > 20:
        assert it.value <= it1 // (2 if sections == 3 else 3) + 6


@pytest.mark.parametrize("name", list(cases()))
def test_inverse_iteration_all_vectors(hc, name):
    d, e = cases()[name]
    w, _ = stebz(hc, d, e)
    Z, fail, nc, mc = stein(hc, d, e, w, len(d))
    assert fail == 0
    check(d, e, w, Z)
    if name == "glued_exact_double":
        assert mc >= 2          # the double eigenvalues were clustered and reorthogonalised
    if name == "random":
        assert mc <= 32         # ... while a generic spectrum stays in tiny clusters


def test_selected_vectors_match_dstein_up_to_sign(hc):
    rng = np.random.default_rng(5)
    n, k = 900, 120
    d, e = rng.standard_normal(n), rng.standard_normal(n - 1)
    w, _ = stebz(hc, d, e)
    Z, fail, nc, mc = stein(hc, d, e, w, k)
    assert fail == 0 and Z.shape == (n, k)
    check(d, e, w, Z)
    wr, Zr = sla.eigh_tridiagonal(d, e, select="i", select_range=(0, k - 1), lapack_driver="stebz")
    assert np.max(np.abs(w[:k] - wr)) <= 1e-14 * np.abs(w).max()
    gaps = np.minimum(np.diff(w[:k + 1])[1:], np.diff(w[:k + 1])[:-1])
    for j in range(1, k - 1):
        if gaps[j - 1] > 1e-3:   # well separated: the eigenvector is determined to ~eps/gap
            s = np.sign(Z[:, j] @ Zr[:, j])
            assert np.linalg.norm(Z[:, j] - s * Zr[:, j]) <= 1e-11


def test_tridiagonal_of_the_shipped_vcnt400_matrix(hc, golden_dir):
    """configs[1]'s matrix has (near-)degenerate eigenvalues: the clusters must come out orthogonal, and the selected
    lowest pairs must reproduce the shipped answer file after the back-transformation."""
    mA = app_io.read_matrix_file(os.path.join(golden_dir, "ELSES_MATRIX_VCNT400std_A.mtx"))
    A = app_io.sparse_to_dense(mA)
    H, Q = sla.hessenberg(A, calc_q=True)
    d, e = np.ascontiguousarray(np.diag(H)), np.ascontiguousarray(np.diag(H, -1))
    w, _ = stebz(hc, d, e)
    E = app_io.read_indexed_values(os.path.join(golden_dir, "ELSES_MATRIX_VCNT400std_E.txt"))
    assert np.max(np.abs(w - E)) <= 6e-13
    for k in (40, 400):
        Z, fail, nc, mc = stein(hc, d, e, w, k)
        assert fail == 0
        check(d, e, w, Z)
        X = Q @ Z
        r = lt.residual_metrics(A, w[:k], X)
        o = lt.orthogonality_metrics(X)
        assert r["res_max_over_A"] <= 1e-12 * 400 and o["orth_fro"] <= 1e-12 * 400


def test_dense_spectrum_keeps_clusters_small_and_orthogonality(hc):
    """Semicircle-like spectrum (the synthetic benchmark matrices): dstein's 1e-3 |T| rule would chain everything into
    one serial cluster; the adaptive threshold must not, and orthogonality must hold without it."""
    n, k = 3000, 300
    A, _ = lt.synthetic_pair(n, 20240603)
    d, e, *_ = lt.sytrd_lower(A) if hasattr(lt, "sytrd_lower") else (None, None)
    if d is None:
        H = sla.hessenberg(A)
        d, e = np.ascontiguousarray(np.diag(H)), np.ascontiguousarray(np.diag(H, -1))
    w, _ = stebz(hc, d, e)
    Z, fail, nc, mc = stein(hc, d, e, w, k)
    assert fail == 0 and mc <= 16 and nc >= k // 8
    check(d, e, w, Z)


def test_slabs_of_a_sharded_solve_reproduce_the_single_rank_vectors(hc):
    """Rank-per-GPU runs: each rank processes every cluster that touches its column slab as a whole, so the slabs put
    together are bit-identical to the single-rank result even when a cluster of (near-)degenerate eigenvalues straddles
    a slab border."""
    rng = np.random.default_rng(8)
    db, eb = rng.standard_normal(130), rng.standard_normal(129)
    d = np.concatenate([db, db, db])                      # exactly triple eigenvalues: clusters of 3 everywhere
    e = np.concatenate([eb, [0.0], eb, [0.0], eb])
    n, nev = len(d), 300
    w, _ = stebz(hc, d, e)
    Z1, fail, nc, mc = stein(hc, d, e, w, nev)
    assert fail == 0 and mc >= 3
    ee = np.ascontiguousarray(e)
    for bounds in ([0, 128, 256, 300], [0, 100, 200, 300], [0, 1, 299, 300], [0, 300]):
        Zs = np.zeros((n, nev), order="F")
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            Zr = np.full((n, nev), np.nan, order="F")
            f = hc.ekb200_host_stein_slab(n, d.ctypes.data_as(dp), ee.ctypes.data_as(dp), nev, w.ctypes.data_as(dp), lo, hi,
                                          Zr.ctypes.data_as(dp), n)
            assert f == 0
            assert np.all(np.isfinite(Zr[:, lo:hi]))
            Zs[:, lo:hi] = Zr[:, lo:hi]
        assert np.array_equal(Zs, Z1)
    check(d, e, w, Z1)


@pytest.mark.parametrize("name", ["random", "glued_exact_double", "glued_tiny_coupling", "wilkinson21", "n2", "diagonal"])
def test_warp_style_execution_of_the_cluster_procedure(hc, name):
    """stein_cluster executed the way the CUDA kernel executes it: 32 (here also 8) lanes in lock step, lane 0 on the
    serial recurrences, all lanes on the strided vector work, butterfly reductions -- same quality as the 1-lane run."""
    d, e = cases()[name]
    n = len(d)
    nev = min(n, 120)
    w, _ = stebz(hc, d, e)
    ee = np.ascontiguousarray(e if n > 1 else np.zeros(1))
    for lanes in (32, 8):
        Z = np.zeros((n, nev), order="F")
        fail = hc.ekb200_host_stein_lanes(n, d.ctypes.data_as(dp), ee.ctypes.data_as(dp), nev, w.ctypes.data_as(dp), lanes,
                                          Z.ctypes.data_as(dp), n)
        assert fail == 0
        check(d, e, w, Z)
