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
