"""GPU tests of the bisection + inverse-iteration path of the selecting solvers (csrc/stebz.cu; the pdstebz + pdstein
half of pdsyevx, reference src/solver_scalapack_select.f90:52-60): the stage-level entry point against the host run of
the same numerics (libekb200_hostcheck.so) and LAPACK, and the whole `-n` solves with option "select_method" = 2
against the dsyevx twin and the shipped VCNT400 answer file.  Tolerances as in test_gpu_solve.py."""
import ctypes
import os

import numpy as np
import pytest
import scipy.linalg as sla

from eigenkernel_b200 import app_io
from oracle import lapack_twin as lt
from test_gpu_solve import _arg, _dense_to_coo, _info, check_pairs      # tests/ is on sys.path (rootdir conftest)
from test_host_tridiag import HC, cases, check, dp, ll

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hc():
    lib = ctypes.CDLL(HC)
    lib.ekb200_host_stebz.argtypes = [ll, dp, dp, dp, ctypes.POINTER(ctypes.c_int)]
    return lib


def _dev_stebz_stein(ctx, d, e, nev):
    n = len(d)
    dd = ctx.from_numpy(d)
    de = ctx.from_numpy(e if n > 1 else np.zeros(1))
    dw = ctx.matrix(n, 1)
    dZ = ctx.matrix(n, max(nev, 1))
    info = ctx.call("ekb200_stebz_stein", n, nev, dd.ptr, de.ptr, dw.ptr, dZ.ptr, dZ.ld)
    w = dw.download()[:, 0]
    Z = dZ.download()[:, :nev]
    for m in (dd, de, dw, dZ):
        m.free()
    return info, w, Z


@pytest.mark.parametrize("name", list(cases()))
def test_stage_matches_host_run_of_the_same_numerics(ctx, hc, name):
    d, e = cases()[name]
    n = len(d)
    info, w, Z = _dev_stebz_stein(ctx, d, e, n)
    assert info == 0
    w_host = np.zeros(n)
    it = ctypes.c_int()
    ee = np.ascontiguousarray(e if n > 1 else np.zeros(1))
    hc.ekb200_host_stebz(n, d.ctypes.data_as(dp), ee.ctypes.data_as(dp), w_host.ctypes.data_as(dp), ctypes.byref(it))
    tn = np.abs(d).max() + 2 * (np.abs(e).max() if len(e) else 0.0)
    # same recurrence, IEEE division on both sides: the Sturm counts agree, so the bisection results do
    assert np.max(np.abs(w - w_host)) <= 4 * np.finfo(float).eps * tn
    assert np.all(np.diff(w) >= 0)
    check(d, e, w, Z)


def test_stage_selected_subset_larger_problem(ctx):
    rng = np.random.default_rng(21)
    n, k = 6000, 500
    d, e = rng.standard_normal(n), rng.standard_normal(n - 1)
    info, w, Z = _dev_stebz_stein(ctx, d, e, k)
    assert info == 0 and Z.shape == (n, k)
    ref = sla.eigvalsh_tridiagonal(d, e, lapack_driver="stebz")
    assert np.max(np.abs(w - ref)) <= 1e-13 * np.abs(ref).max()
    check(d, e, w, Z)


@pytest.mark.parametrize("n,k", [(300, 1), (300, 30), (1000, 100), (777, 776), (2500, 250)])
def test_select_solvers_by_bisection_and_inverse_iteration(ctx, n, k):
    from eigenkernel_b200.solver import eigen_solver, validate_argument
    A, B = lt.synthetic_pair(n, 123 + n)
    mA, mB = _dense_to_coo(A), _dense_to_coo(B)
    ctx.set_option("select_method", 2)
    try:
        w_ref, _ = lt.general_scalapack_select_twin(A, B, k)
        arg = _arg("general_b200_select", _info(n, mA.num_non_zeros), _info(n, mB.num_non_zeros), n_vec=k)
        validate_argument(arg)
        ep, _ = eigen_solver(arg, mA, mB, ctx=ctx)
        assert ep.blacs.Vectors.shape == (n, k)
        check_pairs(A, B, ep.blacs.values, ep.blacs.Vectors, w_ref)
        assert "eigen_solver_b200:stebz_stein" in [e[0] for e in ctx.events()]
        w_ref, _ = lt.scalapack_select_twin(A, k)
        arg = _arg("b200_select", _info(n, mA.num_non_zeros), n_vec=k)
        ep, _ = eigen_solver(arg, mA, ctx=ctx)
        check_pairs(A, None, ep.blacs.values, ep.blacs.Vectors, w_ref)
    finally:
        ctx.set_option("select_method", 0)


def test_select_by_bisection_on_the_shipped_vcnt400_matrix(ctx, golden_dir):
    from eigenkernel_b200.solver import eigen_solver
    fa = os.path.join(golden_dir, "ELSES_MATRIX_VCNT400std_A.mtx")
    ia = app_io.read_matrix_info(fa)
    mA = app_io.read_matrix_file(fa, ia)
    E = app_io.read_indexed_values(os.path.join(golden_dir, "ELSES_MATRIX_VCNT400std_E.txt"))
    A = app_io.sparse_to_dense(mA)
    ctx.set_option("select_method", 2)
    try:
        for k in (40, 399):
            ep, _ = eigen_solver(_arg("b200_select", ia, n_vec=k), mA, ctx=ctx)
            assert np.max(np.abs(ep.blacs.values - E)) <= 6e-13     # all n eigenvalues come back, like pdsyevx's `values`
            r = lt.residual_metrics(A, ep.blacs.values[:k], ep.blacs.Vectors)
            o = lt.orthogonality_metrics(ep.blacs.Vectors)
            assert r["res_max_over_A"] <= 1e-12 * 400 and o["orth_fro"] <= 1e-12 * 400
    finally:
        ctx.set_option("select_method", 0)
