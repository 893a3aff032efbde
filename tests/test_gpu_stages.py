"""Per-stage parity of the CUDA path (through the C-ABI) against the LAPACK-twin oracle. GPU only."""
import numpy as np
import pytest

from oracle import lapack_twin as lt

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ta,tb,m,n,k", [
    ("N", "N", 256, 256, 256), ("N", "T", 300, 200, 100), ("T", "N", 129, 257, 65), ("T", "T", 64, 64, 64),
    ("N", "N", 1, 1, 1), ("N", "N", 1000, 37, 513), ("T", "N", 37, 1000, 2049), ("N", "T", 511, 513, 17),
    # interior fast path (full 128x128 / 128x64 / 64x128 tiles, full k-tiles) next to ragged edges and k tails
    ("N", "N", 1024, 768, 512), ("N", "T", 1024, 768, 512), ("T", "N", 1024, 768, 512), ("T", "T", 1024, 768, 512),
    ("N", "N", 1100, 700, 500), ("N", "T", 1100, 700, 500), ("T", "N", 1100, 700, 500), ("T", "T", 1100, 700, 500),
    ("N", "N", 2048, 64, 1000), ("T", "N", 64, 2048, 1000), ("N", "T", 2048, 64, 136), ("T", "T", 50, 640, 136),
    # few tiles, deep k: the TMA-fed kernel splits k (partial sums + reduce pass), ragged tiles / odd m / k tail included
    ("N", "N", 2048, 64, 4096), ("T", "N", 512, 1024, 4100), ("N", "T", 1100, 700, 3000), ("N", "N", 1001, 64, 2051),
    ("T", "T", 384, 200, 5000),
])
def test_dgemm_matches_numpy(ctx, ta, tb, m, n, k):
    rng = np.random.default_rng(m * 7 + n * 3 + k)
    A = rng.standard_normal((k, m) if ta == "T" else (m, k))
    B = rng.standard_normal((n, k) if tb == "T" else (k, n))
    C = rng.standard_normal((m, n))
    dA, dB, dC = ctx.from_numpy(A), ctx.from_numpy(B), ctx.from_numpy(C)
    ctx.dgemm(ta, tb, 1.5, dA, dB, -0.5, dC)
    ref = 1.5 * (A.T if ta == "T" else A) @ (B.T if tb == "T" else B) - 0.5 * C
    got = dC.download()
    assert np.max(np.abs(got - ref)) <= 1e-13 * k * max(1.0, np.abs(ref).max())
    for d in (dA, dB, dC):
        d.free()


def test_dgemm_unaligned_submatrix(ctx):
    rng = np.random.default_rng(5)
    A = rng.standard_normal((301, 203))
    B = rng.standard_normal((203, 155))
    dA, dB = ctx.from_numpy(A), ctx.from_numpy(B)
    dC = ctx.matrix(300, 154)
    # odd offsets: A(1:, 1:), B(1:, 1:)  -> 8-byte-aligned only
    ctx.call("ekb200_dgemm", b"N", b"N", 300, 154, 202, 1.0, dA.addr(1, 1), dA.ld, dB.addr(1, 1), dB.ld, 0.0,
             dC.ptr, dC.ld)
    ref = A[1:, 1:] @ B[1:, 1:]
    assert np.max(np.abs(dC.download() - ref)) <= 1e-11


@pytest.mark.parametrize("n", [1, 30, 64, 65, 200, 513, 1024])
def test_potrf_matches_oracle(ctx, n):
    _, B = lt.synthetic_pair(n, 11 + n)
    L_ref = np.array(B, order="F")
    assert lt.potrf_lower(L_ref) == 0
    dB = ctx.from_numpy(B)
    assert ctx.call("ekb200_potrf", n, dB.ptr, dB.ld) == 0
    L = np.tril(dB.download())
    assert np.max(np.abs(L - np.tril(L_ref))) <= 1e-13 * n
    dB.free()


def test_potrf_reports_non_spd(ctx):
    n = 200
    _, B = lt.synthetic_pair(n, 3)
    B[130, 130] = -1.0
    L_ref = np.array(B, order="F")
    info_ref = lt.potrf_lower(L_ref)
    dB = ctx.from_numpy(B)
    assert ctx.call("ekb200_potrf", n, dB.ptr, dB.ld) == info_ref == 131
    dB.free()


@pytest.mark.parametrize("n", [30, 200, 777])
def test_sygst_and_trtrs_match_oracle(ctx, n):
    A, B = lt.synthetic_pair(n, 100 + n)
    L = np.array(B, order="F")
    lt.potrf_lower(L)
    Ar = np.array(A, order="F")
    lt.sygst_lower(Ar, L)
    dB, dA = ctx.from_numpy(B), ctx.from_numpy(A)
    assert ctx.call("ekb200_potrf", n, dB.ptr, dB.ld) == 0
    assert ctx.call("ekb200_sygst", n, dA.ptr, dA.ld, dB.ptr, dB.ld) == 0
    got = dA.download()
    scale = np.abs(Ar).max()
    assert np.max(np.abs(np.tril(got) - np.tril(Ar))) <= 1e-13 * n * scale
    assert np.max(np.abs(got - got.T)) <= 1e-13 * n * scale
    Z = np.random.default_rng(n).standard_normal((n, 50))
    Zr = np.array(Z, order="F")
    lt.trtrs_LTN(L, Zr)
    dZ = ctx.from_numpy(Z)
    assert ctx.call("ekb200_trtrs_lt", n, 50, dB.ptr, dB.ld, dZ.ptr, dZ.ld) == 0
    assert np.max(np.abs(dZ.download() - Zr)) <= 1e-13 * n * np.abs(Zr).max()


def test_synthetic_fill_is_bit_identical_to_oracle(ctx):
    n = 300
    A, B = lt.synthetic_pair(n, 20240601)
    dA, dB = ctx.matrix(n, n), ctx.matrix(n, n)
    ctx.call("ekb200_fill_synthetic", n, 20240601, 1.0, 0, 0.0, dA.ptr, dA.ld)
    ctx.call("ekb200_fill_synthetic", n, 20240602, float(n), 1, 2.0, dB.ptr, dB.ld)
    assert np.array_equal(dA.download(), A)
    assert np.array_equal(dB.download(), B)


def test_coo_scatter_last_duplicate_wins_in_both_triangles(ctx):
    """distribute_global_sparse_matrix (distribute_matrix.f90:411-418) is a sequential pdelset loop: when an element
    occurs more than once -- repeated, or as (i,j) and later as (j,i) -- the LAST entry wins, in both triangles.  The
    device scatter must reproduce that independently of thread order (run twice, bit-identical)."""
    n, nnz = 97, 20000
    rng = np.random.default_rng(5)
    ij = rng.integers(1, n + 1, size=(nnz, 2)).astype(np.int32)   # heavy duplication, both orientations
    v = rng.standard_normal(nnz)
    ref = np.zeros((n, n), order="F")
    for (i, j), x in zip(ij, v):   # the reference's loop
        ref[i - 1, j - 1] = x
        ref[j - 1, i - 1] = x
    dA = ctx.matrix(n, n)
    for _ in range(2):
        assert ctx.call("ekb200_coo_to_dense", n, nnz, ij.ctypes.data, v.ctypes.data, dA.ptr, dA.ld) == 0
        got = dA.download()
        assert np.array_equal(got, ref)
        assert np.array_equal(got, got.T)
    # out-of-range entries are ignored, an empty list gives the zero matrix
    bad = np.array([[0, 1], [n + 1, 2], [3, 3]], dtype=np.int32)
    bv = np.array([1.0, 2.0, 3.0])
    assert ctx.call("ekb200_coo_to_dense", n, 3, bad.ctypes.data, bv.ctypes.data, dA.ptr, dA.ld) == 0
    got = dA.download()
    assert got[2, 2] == 3.0 and np.count_nonzero(got) == 1
    dA.free()


def test_fp64_peak_probe(ctx):
    p = ctx.fp64_peak()
    print("FP64 peak:", p)
    assert p["dmma_tflops"] > 1.0 and p["dfma_tflops"] > 1.0
