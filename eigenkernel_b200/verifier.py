"""Host-side mirror of the reference's result checks, computed on the B200 (csrc/verify.cu):

  eval_residual_norm_blacs   src/verifier.f90:75-204   (option -c)
  eval_orthogonality_blacs   src/verifier.f90:233-330  (option -t)
  get_ipratios               src/distribute_matrix.f90:18-78 (ipratios.dat)

Same names, argument meaning and error behaviour as the reference routines; the eigenpairs are the type-2
container `eigen_solver` returns (solver.py).  All arithmetic happens in libekb200.so; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes

import numpy as np

from .app_io import SparseMat, TerminateError
from .device import Context


def _coo(m: SparseMat | None):
    if m is None:
        return 0, None, None, None, None
    ij = np.ascontiguousarray(m.suffix, dtype=np.int32)
    v = np.ascontiguousarray(m.value, dtype=np.float64)
    return m.num_non_zeros, ij.ctypes.data, v.ctypes.data, ij, v


def _vectors(eigenpairs):
    if eigenpairs.type_number != 2:
        raise TerminateError("verifier: eigenpairs of type 2 (BLACS) expected", 1)
    X = np.asfortranarray(eigenpairs.blacs.Vectors, dtype=np.float64)
    n, nvec = int(eigenpairs.blacs.desc[2]), int(eigenpairs.blacs.desc[3])
    return X, n, nvec


def eval_residual_norm_blacs(arg, matrix_A: SparseMat, eigenpairs, matrix_B: SparseMat | None = None, *,
                             ctx: Context | None = None, device: int = 0):
    """verifier.f90:75-204 -> (A_norm, res_norm_ave, res_norm_max) over the first arg.n_check_vec vectors."""
    if arg.is_generalized_problem and matrix_B is None:
        raise TerminateError("eval_residual_norm_blacs: matrix_B is not provided", 1)
    X, n, nvec = _vectors(eigenpairs)
    w = np.ascontiguousarray(eigenpairs.blacs.values, dtype=np.float64)
    nnzA, pijA, pvA, _ka, _kb = _coo(matrix_A)
    nnzB, pijB, pvB, _kc, _kd = _coo(matrix_B if arg.is_generalized_problem else None)
    a, ave, mx = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    own = ctx is None
    if own:
        ctx = Context(device)
    try:
        ctx.call("ekb200_eval_residual_norm", n, nvec, int(arg.n_check_vec), nnzA, pijA, pvA, nnzB, pijB, pvB,
                 w.ctypes.data, X.ctypes.data, max(X.shape[0], 1), ctypes.byref(a), ctypes.byref(ave), ctypes.byref(mx))
    finally:
        if own:
            ctx.close()
    return a.value, ave.value, mx.value


def eval_orthogonality_blacs(index1: int, index2: int, eigenpairs, matrix_B: SparseMat | None = None, *,
                             ctx: Context | None = None, device: int = 0) -> float:
    """verifier.f90:233-330 -> orthogonality of eigenvectors index1..index2 (1-based, inclusive)."""
    X, n, nvec = _vectors(eigenpairs)
    if int(eigenpairs.blacs.desc[4]) != int(eigenpairs.blacs.desc[5]):
        raise TerminateError("eval_orthogonality_blacs: anisotropic block size not supported", 1)
    nnzB, pijB, pvB, _ka, _kb = _coo(matrix_B)
    o = ctypes.c_double()
    own = ctx is None
    if own:
        ctx = Context(device)
    try:
        ctx.call("ekb200_eval_orthogonality", n, nvec, int(index1), int(index2), nnzB, pijB, pvB, X.ctypes.data,
                 max(X.shape[0], 1), ctypes.byref(o))
    finally:
        if own:
            ctx.close()
    return o.value


def get_ipratios(proc, V: np.ndarray, V_desc, S_sparse: SparseMat | None = None, *, ctx: Context | None = None):
    """distribute_matrix.f90:18-78 -> ipratios(V_desc(cols_)); S_sparse switches the overlap (generalized) mode."""
    n, nvec = int(V_desc[2]), int(V_desc[3])
    if S_sparse is not None and S_sparse.size != n:
        raise TerminateError("inconsistent matrix dimension", 1)
    X = np.asfortranarray(V, dtype=np.float64)
    nnzB, pijB, pvB, _ka, _kb = _coo(S_sparse)
    out = np.zeros(nvec)
    own = ctx is None
    if own:
        ctx = Context(getattr(proc, "device", 0))
    try:
        ctx.call("ekb200_get_ipratios", n, nvec, nnzB, pijB, pvB, X.ctypes.data, max(X.shape[0], 1), out.ctypes.data)
    finally:
        if own:
            ctx.close()
    return out
