"""Back-transformation parity (GPU, through the C-ABI): the two halves of pdormtr('L','L','N')
(reference src/solver_scalapack_all.f90:115-116) against explicitly accumulated Q2 / Q1."""
import numpy as np
import pytest

from test_gpu_twostage import _run_sb2st, q1_from_panels, q2_from_reflectors

pytestmark = pytest.mark.gpu


# kc = 0: the library picks the slab width (narrow right-hand sides take the 8-columns-per-warp kernel);
# kc = 64 / 112 force the 16-columns-per-warp kernels the full-width solve uses, 1008 / 1012 the 8-per-warp ones
# with 8 / 12 warps
@pytest.mark.parametrize("kc", [0, 64, 112, 1008, 1012])
@pytest.mark.parametrize("n,band,k", [(3, 64, 3), (40, 64, 40), (66, 64, 17), (130, 64, 130), (200, 32, 64),
                                      (333, 64, 333), (500, 64, 65), (700, 32, 700)])
def test_apply_q2_matches_explicit_product(ctx, n, band, k, kc):
    ctx.set_option("band", band)
    ctx.set_option("q2_kc", kc)
    try:
        b = band
        rng = np.random.default_rng(n + k)
        M = rng.standard_normal((n, n))
        M = M + M.T
        Bd = np.triu(np.tril(M, b), -b)
        d, e, V2, TAU2 = _run_sb2st(ctx, Bd, n, b)
        Q = q2_from_reflectors(V2, TAU2, n, b)
        Z = rng.standard_normal((n, k))
        dV2, dTAU, dZ = ctx.from_numpy(V2), ctx.from_numpy(TAU2), ctx.from_numpy(Z)
        assert ctx.call("ekb200_apply_q2", n, k, dV2.ptr, dV2.ld, dTAU.ptr, dTAU.ld, dZ.ptr, dZ.ld) == 0
        got = dZ.download()
        ref = Q @ Z
        assert np.max(np.abs(got - ref)) <= 1e-13 * n * np.abs(ref).max()
        for x in (dV2, dTAU, dZ):
            x.free()
    finally:
        ctx.set_option("band", 64)
        ctx.set_option("q2_kc", 0)


@pytest.mark.parametrize("n,band,k", [(67, 64, 67), (130, 64, 5), (300, 64, 300), (300, 32, 77), (390, 64, 390),
                                      (777, 64, 100), (1100, 64, 1100)])
def test_apply_q1_matches_explicit_product(ctx, n, band, k):
    from oracle import lapack_twin as lt
    ctx.set_option("band", band)
    try:
        b = band
        A, _ = lt.synthetic_pair(n, 900 + n)
        dA = ctx.from_numpy(A)
        dAB = ctx.matrix(2 * b, n)
        npan = ctx.lib.ekb200_sy2sb_num_panels(ctx.h, n)
        dT = ctx.matrix(b * b, max(npan, 1))
        assert ctx.call("ekb200_sy2sb", n, dA.ptr, dA.ld, dAB.ptr, dAB.ld, dT.ptr) == 0
        Aout = dA.download()
        T1 = dT.download().T.reshape(max(npan, 1), b, b).transpose(0, 2, 1)
        Q = q1_from_panels(Aout, T1, n, b)
        Z = np.random.default_rng(n).standard_normal((n, k))
        dZ = ctx.from_numpy(Z)
        assert ctx.call("ekb200_apply_q1", n, k, dA.ptr, dA.ld, dT.ptr, dZ.ptr, dZ.ld) == 0
        got = dZ.download()
        ref = Q @ Z
        assert np.max(np.abs(got - ref)) <= 1e-13 * n * np.abs(ref).max()
        for x in (dA, dAB, dT, dZ):
            x.free()
    finally:
        ctx.set_option("band", 64)
