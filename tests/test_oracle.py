"""Pins the oracle (oracle/lapack_twin.py) against every golden file the reference ships
(BASELINE.md §3).  CPU only."""
import os

import numpy as np

from oracle import lapack_twin as lt


def _read_indexed(path):
    return np.array([float(l.split()[1]) for l in open(path) if l.strip()])


def test_bnz30_eigenvalues_match_shipped_answer(golden_dir):
    A = lt.read_mtx_dense(os.path.join(golden_dir, "ELSES_MATRIX_BNZ30_A.mtx"))
    B = lt.read_mtx_dense(os.path.join(golden_dir, "ELSES_MATRIX_BNZ30_B.mtx"))
    ev = _read_indexed(os.path.join(golden_dir, "ELSES_MATRIX_BNZ30_ev.txt"))
    w, X, L = lt.general_scalapack_twin(A, B)
    assert w.shape == (30,)
    assert np.max(np.abs(w - ev) / np.abs(ev)) <= 1e-12
    r = lt.residual_metrics(A, w, X, B)
    o = lt.orthogonality_metrics(X, B)
    assert r["res_max_over_A"] <= 1e-12 * 30
    assert o["orth_fro"] <= 1e-12 * 30


def test_bnz30_ipratios_match_shipped_answer(golden_dir):
    A = lt.read_mtx_dense(os.path.join(golden_dir, "ELSES_MATRIX_BNZ30_A.mtx"))
    B = lt.read_mtx_dense(os.path.join(golden_dir, "ELSES_MATRIX_BNZ30_B.mtx"))
    ipr_ref = _read_indexed(os.path.join(golden_dir, "ELSES_MATRIX_BNZ30_ipr.txt"))
    w, X, _ = lt.general_scalapack_twin(A, B)
    ipr = lt.ipratios(X, B)
    # near-degenerate pairs (gaps 3e-9..2e-7) limit reproducibility to ~5e-9
    assert np.max(np.abs(ipr - ipr_ref) / ipr_ref) <= 1e-7


def test_vcnt400_eigenvalues_match_shipped_answer(golden_dir):
    A = lt.read_mtx_dense(os.path.join(golden_dir, "ELSES_MATRIX_VCNT400std_A.mtx"))
    E = _read_indexed(os.path.join(golden_dir, "ELSES_MATRIX_VCNT400std_E.txt"))
    w, Z = lt.scalapack_twin(A)
    assert np.max(np.abs(w - E)) <= 6e-13  # file is rounded to 12 decimals
    w2, _ = lt.syevd(A)
    assert np.max(np.abs(w - w2)) <= 1e-12


def test_select_twin_matches_full(golden_dir):
    A = lt.read_mtx_dense(os.path.join(golden_dir, "ELSES_MATRIX_VCNT400std_A.mtx"))
    w, _ = lt.scalapack_twin(A)
    ws, Zs = lt.scalapack_select_twin(A, 40)
    assert np.max(np.abs(ws - w[:40])) <= 1e-12
    assert lt.residual_metrics(A, ws, Zs)["res_max_over_A"] <= 1e-12 * 400


def test_general_select_twin(golden_dir):
    A = lt.read_mtx_dense(os.path.join(golden_dir, "ELSES_MATRIX_BNZ30_A.mtx"))
    B = lt.read_mtx_dense(os.path.join(golden_dir, "ELSES_MATRIX_BNZ30_B.mtx"))
    ev = _read_indexed(os.path.join(golden_dir, "ELSES_MATRIX_BNZ30_ev.txt"))
    w, X = lt.general_scalapack_select_twin(A, B, 7)
    assert np.max(np.abs(w - ev[:7]) / np.abs(ev[:7])) <= 1e-12


def test_synthetic_generator_is_symmetric_and_spd():
    A, B = lt.synthetic_pair(64, 20240601)
    assert np.array_equal(A, A.T) and np.array_equal(B, B.T)
    assert np.all(np.abs(A) <= 1.0)
    assert np.linalg.eigvalsh(B).min() > 1.0
    w, X, _ = lt.general_scalapack_twin(A, B)
    assert lt.residual_metrics(A, w, X, B)["res_max_over_A"] <= 1e-12 * 64
