"""Two-stage tridiagonalization + D&C + back-transformation parity (GPU, through the C-ABI)."""
import numpy as np
import pytest

from oracle import lapack_twin as lt

pytestmark = pytest.mark.gpu


def band_to_dense(AB, n, b):
    M = np.zeros((n, n))
    for d in range(b + 1):
        idx = np.arange(n - d)
        M[idx + d, idx] = AB[d, : n - d]
        M[idx, idx + d] = AB[d, : n - d]
    return M


def q1_from_panels(Aout, T1, n, b):
    """Q1 = prod_p (I - V_p T_p V_p^T), V_p explicit in A(j+b:, j:j+b)."""
    Q = np.eye(n)
    p, j = 0, 0
    while n - j - b >= 2:
        V = np.zeros((n, b))
        V[j + b:, :] = Aout[j + b:, j:j + b]
        T = T1[p]
        Q = Q @ (np.eye(n) - V @ T @ V.T)
        p, j = p + 1, j + b
    return Q


@pytest.mark.parametrize("n,band", [(67, 64), (130, 64), (300, 64), (300, 32), (777, 64), (1024, 64)])
def test_sy2sb_band_is_orthogonally_similar(ctx, n, band):
    ctx.set_option("band", band)
    b = band
    A, _ = lt.synthetic_pair(n, 500 + n)
    dA = ctx.from_numpy(A)
    ldab = 2 * b
    dAB = ctx.matrix(ldab, n)
    npan = ctx.lib.ekb200_sy2sb_num_panels(ctx.h, n)
    dT = ctx.matrix(b * b, max(npan, 1))
    assert ctx.call("ekb200_sy2sb", n, dA.ptr, dA.ld, dAB.ptr, dAB.ld, dT.ptr) == 0
    AB = dAB.download()
    assert np.all(AB[b + 1:, :] == 0.0)
    Bd = band_to_dense(AB, n, b)
    w_ref = np.linalg.eigvalsh(A)
    w = np.linalg.eigvalsh(Bd)
    anorm = np.abs(w_ref).max()
    assert np.max(np.abs(w - w_ref)) <= 1e-13 * n * anorm
    if n <= 400:
        Aout = dA.download()
        T1 = dT.download().T.reshape(max(npan, 1), b, b).transpose(0, 2, 1)
        Q = q1_from_panels(Aout, T1, n, b)
        assert np.max(np.abs(Q.T @ Q - np.eye(n))) <= 1e-13 * n
        assert np.max(np.abs(Q.T @ A @ Q - Bd)) <= 1e-13 * n * anorm
    ctx.set_option("band", 64)
    for d in (dA, dAB, dT):
        d.free()


def q2_from_reflectors(V2, TAU2, n, b):
    """Q2 = prod_{s ascending} prod_{t ascending} H(s,t)."""
    Q = np.eye(n)
    for s in range(n - 2):
        t = 0
        while True:
            r0 = s + 1 + t * b
            nr = min(b, n - r0)
            if nr < 2:
                break
            v = V2[r0:r0 + nr, s]
            tau = TAU2[t, s]
            Q[:, r0:r0 + nr] -= tau * np.outer(Q[:, r0:r0 + nr] @ v, v)
            t += 1
    return Q


def _run_sb2st(ctx, Bd, n, b):
    AB = np.zeros((2 * b, n), order="F")
    for d in range(min(b, n - 1) + 1):
        AB[d, : n - d] = np.diagonal(Bd, -d)
    dAB = ctx.from_numpy(AB)
    dV2 = ctx.matrix(n, n)
    ntm = ctx.lib.ekb200_sb2st_max_tasks(ctx.h, n)
    dTAU = ctx.matrix(ntm, n)
    dd, de = ctx.matrix(n, 1), ctx.matrix(n, 1)
    assert ctx.call("ekb200_sb2st", n, dAB.ptr, dAB.ld, dV2.ptr, dV2.ld, dTAU.ptr, dTAU.ld, dd.ptr, de.ptr) == 0
    d = dd.download()[:, 0]
    e = de.download()[: n - 1, 0]
    out = d, e, dV2.download(), dTAU.download()
    for x in (dAB, dV2, dTAU, dd, de):
        x.free()
    return out


@pytest.mark.parametrize("n,band", [(3, 64), (40, 64), (66, 64), (130, 64), (200, 32), (333, 64), (1000, 64)])
def test_sb2st_tridiagonal_is_orthogonally_similar(ctx, n, band):
    ctx.set_option("band", band)
    b = band
    rng = np.random.default_rng(n)
    M = rng.standard_normal((n, n))
    M = M + M.T
    Bd = np.triu(np.tril(M, b), -b)
    d, e, V2, TAU2 = _run_sb2st(ctx, Bd, n, b)
    ctx.set_option("band", 64)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    w_ref = np.linalg.eigvalsh(Bd)
    w = np.linalg.eigvalsh(T)
    anorm = np.abs(w_ref).max()
    assert np.max(np.abs(w - w_ref)) <= 1e-13 * n * anorm
    if n <= 340:
        Q = q2_from_reflectors(V2, TAU2, n, b)
        assert np.max(np.abs(Q.T @ Q - np.eye(n))) <= 1e-13 * n
        assert np.max(np.abs(Q.T @ Bd @ Q - T)) <= 1e-13 * n * anorm


def _tridiag_cases():
    rng = np.random.default_rng(7)
    cases = []
    for n in (1, 2, 3, 31, 32, 33, 64, 100, 257, 1000, 2500):
        cases.append((f"rand{n}", rng.standard_normal(n), rng.standard_normal(max(n - 1, 0))))
    n = 501
    cases.append(("toeplitz", np.zeros(n), np.ones(n - 1)))                       # heavy deflation
    m = 200
    cases.append(("wilkinson", np.abs(np.arange(-m, m + 1)).astype(float), np.ones(2 * m)))  # close pairs
    e = rng.standard_normal(399)
    e[[50, 51, 199, 300]] = 0.0
    cases.append(("split", rng.standard_normal(400), e))                           # exact zeros on cuts or not
    cases.append(("graded", 10.0 ** (-np.arange(300) / 20.0), 10.0 ** (-np.arange(299) / 20.0 - 1)))
    cases.append(("negrho", rng.standard_normal(640), -np.abs(rng.standard_normal(639))))
    cases.append(("identity", np.ones(300), np.zeros(299)))
    cases.append(("glued", np.tile(np.r_[np.arange(10.0), np.arange(10.0)[::-1]], 16), np.r_[np.ones(319)] * 1e-8 + 1.0))
    return cases


@pytest.mark.parametrize("name,d,e", _tridiag_cases(), ids=[c[0] for c in _tridiag_cases()])
def test_stedc_matches_oracle(ctx, name, d, e):
    import ctypes
    n = d.shape[0]
    w_ref, Z_ref, info = lt.stedc_I(d, e)
    assert info == 0
    dd, de = ctx.from_numpy(d), ctx.from_numpy(np.r_[e, 0.0])
    dw, dZ = ctx.matrix(n, 1), ctx.matrix(n, n)
    fl = ctypes.c_double()
    assert ctx.call("ekb200_stedc", n, dd.ptr, de.ptr, dw.ptr, dZ.ptr, dZ.ld, ctypes.byref(fl)) == 0
    w = dw.download()[:, 0]
    Z = dZ.download()
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    tn = max(np.abs(w_ref).max(), 1e-300)
    assert np.all(np.diff(w) >= 0)
    assert np.max(np.abs(w - w_ref)) <= 1e-14 * max(n, 10) * tn
    assert np.max(np.abs(T @ Z - Z * w[None, :])) <= 1e-14 * max(n, 10) * tn
    assert np.max(np.abs(Z.T @ Z - np.eye(n))) <= 1e-14 * max(n, 10)
    for x in (dd, de, dw, dZ):
        x.free()
