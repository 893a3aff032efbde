"""Device-side result checks (csrc/verify.cu) against the oracle's restatement of verifier.f90 and get_ipratios,
and against the shipped IPR answer file.  GPU only; calls go through the C-ABI via the reference-named mirror."""
import os

import numpy as np
import pytest

from eigenkernel_b200 import app_io, verifier
from eigenkernel_b200.solver import Argument, Eigenpairs, EigenpairsBlacs, Process
from oracle import lapack_twin as lt

pytestmark = pytest.mark.gpu


def _coo(M):
    n = M.shape[0]
    i, j = np.tril_indices(n)
    ij = np.stack([i + 1, j + 1], axis=1).astype(np.int32)
    return app_io.SparseMat(size=n, num_non_zeros=len(i), value=np.ascontiguousarray(M[i, j]), suffix=ij)


def _pairs(w, X):
    n, k = X.shape
    ep = Eigenpairs(type_number=2)
    ep.blacs = EigenpairsBlacs(values=np.array(w), desc=np.array([1, 0, n, k, 64, 64, 0, 0, n], dtype=np.int32),
                               Vectors=np.asfortranarray(X))
    return ep


@pytest.mark.parametrize("n,generalized,ncheck", [(30, True, 30), (257, True, 100), (500, False, 500), (1000, True, 1000)])
def test_residual_orthogonality_ipr_match_oracle(ctx, n, generalized, ncheck):
    A, B = lt.synthetic_pair(n, 100 + n)
    if generalized:
        w, X, _ = lt.general_scalapack_twin(A, B)
    else:
        w, X = lt.scalapack_twin(A)
    rng = np.random.default_rng(n)
    X = np.asfortranarray(X + 1e-9 * rng.standard_normal(X.shape))  # make the metrics non-trivial
    Bm = B if generalized else None
    ep = _pairs(w, X)
    arg = Argument(solver_type="general_b200" if generalized else "b200", is_generalized_problem=generalized,
                   n_vec=n, n_check_vec=ncheck)
    mA, mB = _coo(A), (_coo(B) if generalized else None)
    a_norm, ave, mx = verifier.eval_residual_norm_blacs(arg, mA, ep, mB, ctx=ctx)
    ref = lt.residual_metrics(A, w[:ncheck], X[:, :ncheck], Bm)
    assert abs(a_norm - ref["A_norm"]) <= 1e-13 * ref["A_norm"]
    assert abs(mx - ref["res_max_over_A"]) <= 1e-6 * ref["res_max_over_A"] + 1e-18
    assert abs(ave - ref["res_avg_over_A"]) <= 1e-6 * ref["res_avg_over_A"] + 1e-18
    i1, i2 = 1 + n // 7, n - n // 5
    o = verifier.eval_orthogonality_blacs(i1, i2, ep, mB, ctx=ctx)
    oref = lt.orthogonality_metrics(X[:, i1 - 1:i2], Bm)["verifier_orthogonality"]
    assert abs(o - oref) <= 1e-6 * oref + 1e-18
    ipr = verifier.get_ipratios(Process(), X, ep.blacs.desc, mB, ctx=ctx)
    assert np.max(np.abs(ipr - lt.ipratios(X, Bm)) / lt.ipratios(X, Bm)) <= 1e-12


def test_bnz30_ipratios_from_oracle_vectors_match_shipped_file(ctx, golden_dir):
    A = lt.read_mtx_dense(os.path.join(golden_dir, "ELSES_MATRIX_BNZ30_A.mtx"))
    B = lt.read_mtx_dense(os.path.join(golden_dir, "ELSES_MATRIX_BNZ30_B.mtx"))
    w, X, _ = lt.general_scalapack_twin(A, B)
    ipr_ref = app_io.read_indexed_values(os.path.join(golden_dir, "ELSES_MATRIX_BNZ30_ipr.txt"))
    ipr = verifier.get_ipratios(Process(), X, _pairs(w, X).blacs.desc, _coo(B), ctx=ctx)
    assert np.max(np.abs(ipr - ipr_ref) / ipr_ref) <= 1e-7


def test_argument_errors_mirror_reference(ctx):
    A, B = lt.synthetic_pair(40, 1)
    w, X, _ = lt.general_scalapack_twin(A, B)
    ep = _pairs(w, X)
    arg = Argument(solver_type="general_b200", is_generalized_problem=True, n_vec=40, n_check_vec=40)
    with pytest.raises(app_io.TerminateError):
        verifier.eval_residual_norm_blacs(arg, _coo(A), ep, None, ctx=ctx)
    ep.blacs.desc[5] = 32
    with pytest.raises(app_io.TerminateError):
        verifier.eval_orthogonality_blacs(1, 10, ep, _coo(B), ctx=ctx)


@pytest.mark.parametrize("n,generalized", [(257, True), (500, False)])
def test_b_orthonormality_dev_matches_oracle(ctx, n, generalized):
    """|| X^T B X - I ||_F (BASELINE.json's B-orthogonality metric; bench.py's acceptance block) next to the
    reference's scaled-Gram number, both from one device pass, against the oracle's restatement."""
    import ctypes

    A, B = lt.synthetic_pair(n, 300 + n)
    X = lt.general_scalapack_twin(A, B)[1] if generalized else lt.scalapack_twin(A)[1]
    X = np.asfortranarray(X * (1.0 + 1e-9 * np.arange(n))[None, :] + 1e-10 * np.random.default_rng(n).standard_normal(X.shape))
    dX, dB = ctx.from_numpy(X), ctx.from_numpy(B)
    o, g = ctypes.c_double(), ctypes.c_double()
    i1, i2 = 3, n - 5
    ctx.call("ekb200_eval_b_orthonormality_dev", n, i1, i2, dX.ptr, dX.ld, dB.ptr if generalized else None, dB.ld,
             ctypes.byref(o), ctypes.byref(g))
    ref = lt.orthogonality_metrics(X[:, i1 - 1:i2], B if generalized else None)
    assert abs(o.value - ref["verifier_orthogonality"]) <= 1e-6 * ref["verifier_orthogonality"]
    assert abs(g.value - ref["orth_fro"]) <= 1e-6 * ref["orth_fro"]
    dX.free()
    dB.free()
