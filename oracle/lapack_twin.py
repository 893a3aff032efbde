"""ORACLE (test infrastructure, NOT product code) -- serial-LAPACK twin of EigenKernel's hot path.

The reference (`/root/reference`, Fortran + MPI + ScaLAPACK) cannot be built in this
image (no gfortran / MPI / ScaLAPACK), and all of its arithmetic lives in un-vendored,
unpinned ScaLAPACK.  This module restates the reference's call sequence with the serial
LAPACK routines of the same name minus the leading `p`, taken as compiled code from the
OpenBLAS 0.3.30 that SciPy bundles (`scipy.libs/libscipy_openblas-*.so`, symbols
`scipy_LAPACKE_*`), argument for argument:

    reduce_generalized   generalized_to_standard.f90:24   pdpotrf('L')          -> dpotrf('L')
                         generalized_to_standard.f90:37   pdsygst(1,'L')        -> dsygst(1,'L')
    eigen_solver_scalapack_all
                         solver_scalapack_all.f90:59      pdsytrd('L')          -> dsytrd('L')
                         solver_scalapack_all.f90:96      pdstedc('I')          -> dstedc('I')
                         solver_scalapack_all.f90:115     pdormtr('L','L','N')  -> dormtr('L','L','N')
    recovery_generalized generalized_to_standard.f90:103  pdtrtrs('L','T','N')  -> dtrtrs('L','T','N')
    eigen_solver_scalapack_select
                         solver_scalapack_select.f90:52-60 pdsyevx('V','I','L',il=1,iu=n_vec,
                                                           abstol=2*pdlamch('S')) -> dsyevx(...)

Parity pinning: the reference ships no tests; the only golden data are the end-to-end
answer files `matrix/ELSES_MATRIX_BNZ30_ev.txt`, `_ipr.txt` and
`ELSES_MATRIX_VCNT400std_E.txt` (copied to tests/golden/).  tests/test_oracle.py pins this
oracle against all three.  Per-stage intermediates (L, reduced A, d/e, Z) are NOT pinned by
the reference ("per-stage parity unpinned"); they are pinned only by this twin.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product path (eigenkernel_b200/) never does.
"""
from __future__ import annotations

import ctypes
import glob
import os
import time

import numpy as np

LAPACK_COL_MAJOR = 102
_c_int = ctypes.c_int
_c_char = ctypes.c_char
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def _find_openblas() -> str:
    import scipy

    base = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
    hits = sorted(glob.glob(os.path.join(base, "libscipy_openblas-*.so")))
    if not hits:
        raise ImportError("oracle: SciPy's bundled OpenBLAS (scipy.libs/libscipy_openblas-*.so) not found")
    return hits[0]


_LIB = ctypes.CDLL(_find_openblas())


def set_num_threads(n: int) -> None:
    _LIB.scipy_openblas_set_num_threads(_c_int(int(n)))


def get_num_threads() -> int:
    _LIB.scipy_openblas_get_num_threads.restype = _c_int
    return int(_LIB.scipy_openblas_get_num_threads())


def _fn(name, argtypes):
    f = getattr(_LIB, "scipy_LAPACKE_" + name)
    f.restype = _c_int
    f.argtypes = argtypes
    return f


_dpotrf = _fn("dpotrf", [_c_int, _c_char, _c_int, _dp, _c_int])
_dsygst = _fn("dsygst", [_c_int, _c_int, _c_char, _c_int, _dp, _c_int, _dp, _c_int])
_dsytrd = _fn("dsytrd", [_c_int, _c_char, _c_int, _dp, _c_int, _dp, _dp, _dp])
_dstedc = _fn("dstedc", [_c_int, _c_char, _c_int, _dp, _dp, _dp, _c_int])
_dormtr = _fn("dormtr", [_c_int, _c_char, _c_char, _c_char, _c_int, _c_int, _dp, _c_int, _dp, _dp, _c_int])
_dtrtrs = _fn("dtrtrs", [_c_int, _c_char, _c_char, _c_char, _c_int, _c_int, _dp, _c_int, _dp, _c_int])
_dsyevx = _fn(
    "dsyevx",
    [_c_int, _c_char, _c_char, _c_char, _c_int, _dp, _c_int, ctypes.c_double, ctypes.c_double, _c_int, _c_int,
     ctypes.c_double, _ip, _dp, _dp, _c_int, _ip],
)
_dsyevd = _fn("dsyevd", [_c_int, _c_char, _c_char, _c_int, _dp, _c_int, _dp])


def _p(a: np.ndarray):
    return a.ctypes.data_as(_dp)


def _fcopy(a) -> np.ndarray:
    """Fortran-ordered float64 private copy."""
    return np.array(a, dtype=np.float64, order="F", copy=True)


# ---------------------------------------------------------------- stage twins (in place on F-ordered arrays)
def potrf_lower(B: np.ndarray) -> int:
    """generalized_to_standard.f90:24 -- B = L L^T, lower, in place. info>0: not SPD."""
    n = B.shape[0]
    return _dpotrf(LAPACK_COL_MAJOR, b"L", n, _p(B), B.strides[1] // 8)


def sygst_lower(A: np.ndarray, L: np.ndarray) -> int:
    """generalized_to_standard.f90:37 -- A <- L^-1 A L^-T (itype=1, 'L'); only the lower triangle is defined."""
    n = A.shape[0]
    return _dsygst(LAPACK_COL_MAJOR, 1, b"L", n, _p(A), A.strides[1] // 8, _p(L), L.strides[1] // 8)


def sytrd_lower(A: np.ndarray):
    """solver_scalapack_all.f90:59 -- A = Q T Q^T; returns (d, e, tau, info)."""
    n = A.shape[0]
    d = np.empty(n)
    e = np.empty(max(n - 1, 1))
    tau = np.empty(max(n - 1, 1))
    info = _dsytrd(LAPACK_COL_MAJOR, b"L", n, _p(A), A.strides[1] // 8, _p(d), _p(e), _p(tau))
    return d, e[: n - 1], tau[: n - 1], info


def stedc_I(d: np.ndarray, e: np.ndarray):
    """solver_scalapack_all.f90:96 -- pdstedc('I'): eigen-decomposition of tridiagonal (d,e)."""
    n = d.shape[0]
    w = np.array(d, dtype=np.float64, copy=True)
    ee = np.zeros(max(n, 1))
    ee[: n - 1] = e
    Z = np.zeros((n, n), order="F")
    info = _dstedc(LAPACK_COL_MAJOR, b"I", n, _p(w), _p(ee), _p(Z), max(n, 1))
    return w, Z, info


def ormtr_LLN(A: np.ndarray, tau: np.ndarray, Z: np.ndarray) -> int:
    """solver_scalapack_all.f90:115 -- Z <- Q Z with the reflectors dsytrd('L') left in A."""
    n, k = Z.shape
    t = np.ascontiguousarray(tau, dtype=np.float64)
    if t.size == 0:
        t = np.zeros(1)
    return _dormtr(LAPACK_COL_MAJOR, b"L", b"L", b"N", n, k, _p(A), A.strides[1] // 8, _p(t), _p(Z), Z.strides[1] // 8)


def trtrs_LTN(L: np.ndarray, Z: np.ndarray) -> int:
    """generalized_to_standard.f90:103 -- Z <- L^-T Z."""
    n, k = Z.shape
    return _dtrtrs(LAPACK_COL_MAJOR, b"L", b"T", b"N", n, k, _p(L), L.strides[1] // 8, _p(Z), Z.strides[1] // 8)


# ---------------------------------------------------------------- workflow twins
def scalapack_twin(A, timings: dict | None = None):
    """`-s scalapack` (solver_main.f90:55-58 -> solver_scalapack_all.f90:19-124). Returns (w, Z)."""
    A = _fcopy(A)
    t0 = time.perf_counter()
    d, e, tau, info = sytrd_lower(A)
    if info:
        raise RuntimeError(f"info(pdsytrd): {info}")
    t1 = time.perf_counter()
    w, Z, info = stedc_I(d, e)
    if info:
        raise RuntimeError(f"info(pdstedc): {info}")
    t2 = time.perf_counter()
    info = ormtr_LLN(A, tau, Z)
    if info:
        raise RuntimeError(f"info(pdormtr): {info}")
    t3 = time.perf_counter()
    if timings is not None:
        timings["eigen_solver_scalapack_all:pdsytrd"] = t1 - t0
        timings["eigen_solver_scalapack_all:pdstedc"] = t2 - t1
        timings["eigen_solver_scalapack_all:pdormtr"] = t3 - t2
        timings["eigen_solver_scalapack_all"] = t3 - t0
    return w, Z


def general_scalapack_twin(A, B, timings: dict | None = None):
    """`-s general_scalapack` (solver_scalapack_all.f90:127-168). Returns (w, X, L), X^T B X = I."""
    A = _fcopy(A)
    L = _fcopy(B)
    t0 = time.perf_counter()
    info = potrf_lower(L)
    if info:
        raise RuntimeError(f"info(pdpotrf): {info}")
    t1 = time.perf_counter()
    info = sygst_lower(A, L)
    if info:
        raise RuntimeError(f"info(pdsygst): {info}")
    t2 = time.perf_counter()
    sub: dict = {}
    w, Z = scalapack_twin(A, sub)
    t3 = time.perf_counter()
    info = trtrs_LTN(L, Z)
    if info:
        raise RuntimeError(f"info(pdtrtrs): {info}")
    t4 = time.perf_counter()
    if timings is not None:
        timings["reduce_generalized:pdpotrf"] = t1 - t0
        timings["reduce_generalized:pdsygst"] = t2 - t1
        timings.update(sub)
        timings["recovery_generalized"] = t4 - t3
        timings["solve_with_general_scalapack"] = t4 - t0
    return w, Z, L


def scalapack_select_twin(A, n_vec: int):
    """`-s scalapack_select` (solver_scalapack_select.f90:52-60): lowest n_vec pairs by dsyevx.
    LAPACK dsyevx always re-orthogonalises clusters (no `orfac`), so it is at least as orthogonal as
    the reference's orfac=0 run."""
    A = _fcopy(A)
    n = A.shape[0]
    _LIB.scipy_dlamch_.restype = ctypes.c_double
    safmin = float(_LIB.scipy_dlamch_(ctypes.c_char_p(b"S"), 1))
    m = _c_int(0)
    w = np.zeros(n)
    Z = np.zeros((n, n_vec), order="F")
    ifail = np.zeros(n, dtype=np.int32)
    info = _dsyevx(LAPACK_COL_MAJOR, b"V", b"I", b"L", n, _p(A), n, 0.0, 0.0, 1, n_vec, 2.0 * safmin,
                   ctypes.byref(m), _p(w), _p(Z), n, ifail.ctypes.data_as(_ip))
    if info:
        raise RuntimeError(f"info(pdsyevx): {info}")
    return w[:n_vec], Z


def general_scalapack_select_twin(A, B, n_vec: int):
    """`-s general_scalapack_select` (solver_main.f90:66-75)."""
    A = _fcopy(A)
    L = _fcopy(B)
    if potrf_lower(L):
        raise RuntimeError("info(pdpotrf)")
    if sygst_lower(A, L):
        raise RuntimeError("info(pdsygst)")
    w, Z = scalapack_select_twin(A, n_vec)
    if trtrs_LTN(L, Z):
        raise RuntimeError("info(pdtrtrs)")
    return w, Z


def syevd(A):
    """dsyevd('V','L') -- independent cross-check of the twin chain (what README.md:65 calls PDSYEVD)."""
    A = _fcopy(A)
    n = A.shape[0]
    w = np.zeros(n)
    info = _dsyevd(LAPACK_COL_MAJOR, b"V", b"L", n, _p(A), n, _p(w))
    if info:
        raise RuntimeError(f"dsyevd info {info}")
    return w, A


# ---------------------------------------------------------------- inputs
def read_mtx_dense(path: str) -> np.ndarray:
    """MatrixMarket coordinate real symmetric -> dense symmetric (matrix_io.f90:91-144 +
    distribute_matrix.f90:401-422: every off-diagonal entry mirrored, last duplicate wins)."""
    with open(path) as f:
        line = f.readline()
        if not line.lower().startswith("%%matrixmarket"):
            raise ValueError("not a MatrixMarket file")
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        rows, cols, nnz = (int(x) for x in line.split())
        A = np.zeros((rows, cols), order="F")
        for _ in range(nnz):
            i, j, v = f.readline().split()
            i, j, v = int(i) - 1, int(j) - 1, float(v.replace("D", "E").replace("d", "e"))
            A[i, j] = v
            A[j, i] = v
    return A


_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def synthetic_u(seed: int, n: int) -> np.ndarray:
    """u(s,i,j) = splitmix64(s XOR (max(i,j)<<32 | min(i,j))) mapped to [-1,1); symmetric. SURVEY.md §8(d).
    Bit-identical to the device generator `ekb200_fill_synthetic` (eigenkernel_b200/csrc/fill.cu)."""
    i = np.arange(n, dtype=np.uint64)[:, None]
    j = np.arange(n, dtype=np.uint64)[None, :]
    hi = np.maximum(i, j)
    lo = np.minimum(i, j)
    key = np.uint64(seed) ^ ((hi << np.uint64(32)) | lo)
    z = _splitmix64(key)
    return np.asfortranarray((z >> np.uint64(11)).astype(np.float64) * (2.0 ** -52) - 1.0)


def synthetic_pair(n: int, seed: int, c_A: float = 0.0):
    """A: a_ij = u(seed,i,j), a_ii += c_A;  B: b_ij = u(seed+1,i,j)/n, b_ii = 2 (SPD, cond < 3)."""
    A = synthetic_u(seed, n)
    if c_A:
        A[np.diag_indices(n)] += c_A
    B = synthetic_u(seed + 1, n) / n
    B[np.diag_indices(n)] = 2.0
    return A, np.asfortranarray(B)


# ---------------------------------------------------------------- acceptance metrics
def residual_metrics(A, w, X, B=None):
    """max/avg over columns of ||A x - lambda B x||_2, in the two normalisations of verifier.f90
    (blacs variant :179,198-199 divides by ||A||_F only; local variant :61-63 also by ||x||)."""
    BX = X if B is None else B @ X
    R = A @ X - BX * w[None, : X.shape[1]]
    rn = np.linalg.norm(R, axis=0)
    an = np.linalg.norm(A, "fro")
    xn = np.linalg.norm(X, axis=0)
    return {
        "A_norm": an,
        "res_max_over_A": float(rn.max() / an),
        "res_avg_over_A": float(rn.mean() / an),
        "res_max_over_A_x": float((rn / xn).max() / an),
    }


def orthogonality_metrics(X, B=None):
    """||X^T B X - I||_F and verifier.f90:310-325's criterion (Gram normalised to unit diagonal,
    diagonal zeroed, Frobenius norm)."""
    G = X.T @ (X if B is None else B @ X)
    k = G.shape[0]
    dg = np.sqrt(np.abs(np.diag(G)))
    Gn = G / dg[:, None] / dg[None, :]
    np.fill_diagonal(Gn, 0.0)
    return {
        "orth_fro": float(np.linalg.norm(G - np.eye(k), "fro")),
        "orth_max": float(np.abs(G - np.eye(k)).max()),
        "verifier_orthogonality": float(np.linalg.norm(Gn, "fro")),
    }


def ipratios(X, B=None):
    """distribute_matrix.f90:18-78: IPR_j = sum_i v_ij^4 / (sum_i v_ij (Bv)_ij)^2 (generalized)
    or / (sum_i v_ij^2)^2 (standard)."""
    SV = X if B is None else B @ X
    return (X ** 4).sum(axis=0) / ((X * SV).sum(axis=0) ** 2)
