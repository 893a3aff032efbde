"""GPU tests of the C++ host twin (app/bin/ekb200_app) and of the explicit-inverse reduction variant
(-s general_b200inv, option "reduction" = 1): the reference's own CPU-runnable cases (BASELINE.json configs[0] and
[1]) driven exactly as a user of EigenKernel_App would, checked against the shipped answer files.

Tolerances are BASELINE.json's: eigenvalue relative difference <= 1e-10 (1e-12 on the fixtures), residual
||A x - lambda B x|| / ||A||_F <= 1e-12 n, B-orthogonality <= 1e-12 n."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

from eigenkernel_b200 import app_io
from oracle import lapack_twin as lt

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
APP = os.path.join(ROOT, "app", "bin", "ekb200_app")


_LAUNCHER_VARS = ("OMPI_COMM_WORLD_RANK", "OMPI_COMM_WORLD_SIZE", "PMI_RANK", "PMI_SIZE", "PMIX_RANK", "PMIX_SIZE",
                  "SLURM_PROCID", "SLURM_NTASKS", "RANK", "WORLD_SIZE")


def _clean_env():
    """The app takes its ranks from a launcher's environment when one is set; the tests start single processes."""
    return {k: v for k, v in os.environ.items() if k not in _LAUNCHER_VARS}


def _run(args, cwd, timeout=600):
    assert os.path.exists(APP), "app/bin/ekb200_app is missing: run __graft_entry__.build()"
    r = subprocess.run([APP] + args, cwd=cwd, env=_clean_env(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r


def _log(path):
    txt = open(path).read()
    return json.loads(re.sub(r"(\d)E([+-])0*(\d)", r"\1E\2\3", txt))


def _printed(stdout, label):
    m = re.search(re.escape(label) + r"\s*([-+0-9.E]+)", stdout)
    assert m, (label, stdout)
    return float(m.group(1))


@pytest.mark.parametrize("solver", ["general_b200", "general_b200inv"])
def test_app_bnz30_generalized_matches_shipped_answer_files(golden_dir, tmp_path, solver):
    fa, fb = (os.path.join(golden_dir, f"ELSES_MATRIX_BNZ30_{x}.mtx") for x in "AB")
    r = _run(["-s", solver, "-c", "-1", "-t", "1,30", "-p", "1-2,30", "-d", str(tmp_path), fa, fb], tmp_path)
    ev = app_io.read_indexed_values(os.path.join(golden_dir, "ELSES_MATRIX_BNZ30_ev.txt"))
    w = app_io.read_indexed_values(tmp_path / "eigenvalues.dat")
    assert w.shape == ev.shape and np.max(np.abs(w - ev) / np.abs(ev)) <= 1e-12
    ipr_ref = app_io.read_indexed_values(os.path.join(golden_dir, "ELSES_MATRIX_BNZ30_ipr.txt"))
    ipr = app_io.read_indexed_values(tmp_path / "ipratios.dat")
    assert ipr.shape == ipr_ref.shape and np.max(np.abs(ipr - ipr_ref) / ipr_ref) <= 1e-6
    # line format of the output files is the reference's '(I8, " ", E26.16e3)'
    for ln in open(tmp_path / "eigenvalues.dat").read().splitlines():
        assert re.fullmatch(r" {0,7}\d{1,8} [ -]{1,3}0\.\d{16}E[+-]\d{3}", ln), ln
    # the checker block (main.f90:146-172)
    assert "----- Checker Call -----" in r.stdout
    assert abs(_printed(r.stdout, "A norm:") - 5.348) < 1e-2
    assert _printed(r.stdout, "residual norm (max):") <= 1e-12 * 30
    assert _printed(r.stdout, "orthogonality criterion:") <= 1e-12 * 30
    # eigenvector files: <dir>/<j:08d>.dat, lines '(I8," ",I8," ",E26.16e3)'; they satisfy A x = lambda B x
    A = app_io.sparse_to_dense(app_io.read_matrix_file(fa))
    B = app_io.sparse_to_dense(app_io.read_matrix_file(fb))
    for j in (1, 2, 30):
        rows = [ln.split() for ln in open(tmp_path / f"{j:08d}.dat").read().splitlines()]
        assert len(rows) == 30 and all(int(x[1]) == j for x in rows) and [int(x[0]) for x in rows] == list(range(1, 31))
        x = np.array([float(t[2]) for t in rows])
        assert np.linalg.norm(A @ x - w[j - 1] * (B @ x)) <= 1e-12 * 30 * np.linalg.norm(A)
        assert abs(x @ B @ x - 1.0) <= 1e-12 * 30
    assert not os.path.exists(tmp_path / "00000003.dat")
    # log.json: setting block + the main:* events of main.f90 + the backend's stage events
    doc = _log(tmp_path / "log.json")
    assert doc["setting"]["solver"] == solver and doc["setting"]["dimension"] == 30
    names = [e["name"] for e in doc["events"]]
    for must in ("main", "main:eigen_solver", "main:read_matrix_files", "main:compute_and_print_ipratios",
                 "read_matrix_file", "reduce_generalized_b200:potrf", "eigen_solver_b200:sy2sb", "eigen_solver_b200:sb2st",
                 "eigen_solver_b200:stedc", "recovery_generalized_b200"):
        assert must in names, must
    assert ("reduce_generalized_b200:trtri" in names) == (solver == "general_b200inv")
    assert ("reduce_generalized_b200:sygst" in names) == (solver == "general_b200")
    assert names[0] == "main"  # newest name first (event_logger.f90:56-63)


def test_app_vcnt400_standard_and_selecting(golden_dir, tmp_path):
    fa = os.path.join(golden_dir, "ELSES_MATRIX_VCNT400std_A.mtx")
    ev = np.array([float(l.split()[-1]) for l in open(os.path.join(golden_dir, "ELSES_MATRIX_VCNT400std_E.txt")) if l.strip()])
    _run(["-s", "b200", "-c", "-1", "-o", "all.dat", "-i", "ipr_all.dat", "-l", "log_all.json", fa], tmp_path)
    w = app_io.read_indexed_values(tmp_path / "all.dat")
    assert w.shape == (400,) and np.max(np.abs(w - ev[:400])) <= 6e-13 * max(1.0, np.abs(ev).max())
    r = _run(["-s", "b200_select", "-n", "40", "-c", "40", "-t", "1,40", "--binary", "-p", "40", "-d", str(tmp_path),
              "-o", "sel.dat", fa], tmp_path)
    ws = app_io.read_indexed_values(tmp_path / "sel.dat")
    assert ws.shape == (40,) and np.max(np.abs(ws - w[:40])) <= 1e-12 * np.abs(w).max()
    assert len(open(tmp_path / "ipratios.dat").read().splitlines()) == 40   # desc(cols_) lines (main.f90:138)
    assert _printed(r.stdout, "residual norm (max):") <= 1e-12 * 400
    assert _printed(r.stdout, "orthogonality criterion:") <= 1e-12 * 400
    raw = open(tmp_path / "00000040.dat", "rb").read()                      # Fortran unformatted record
    assert len(raw) == 4 + 400 * 8 + 4 and int.from_bytes(raw[:4], "little") == 3200
    x = np.frombuffer(raw[4:-4])
    A = app_io.sparse_to_dense(app_io.read_matrix_file(fa))
    assert np.linalg.norm(A @ x - ws[39] * x) <= 1e-12 * 400 * np.linalg.norm(A)


def test_app_synthetic_generalized_runs_device_resident(tmp_path):
    """`synthetic:<n>:<seed>` stands in for the MatrixMarket files; the checks run on the device."""
    n, seed = 768, 20240601
    r = _run(["-s", "general_b200", "-c", "-1", "-t", f"1,{n}", f"synthetic:{n}:{seed}", f"synthetic:{n}:{seed + 1}"], tmp_path)
    w = app_io.read_indexed_values(tmp_path / "eigenvalues.dat")
    A, B = lt.synthetic_pair(n, seed)
    w_ref, _, _ = lt.general_scalapack_twin(A, B)
    assert np.max(np.abs(w - w_ref)) <= 1e-12 * np.abs(w_ref).max()
    assert _printed(r.stdout, "residual norm (max):") <= 1e-12 * n
    assert _printed(r.stdout, "orthogonality criterion:") <= 1e-12 * n


def test_explicit_inverse_reduction_matches_blocked_reduction(ctx):
    """Option "reduction" = 1 (X = L^-1, A <- X A X^T, Z <- X^T Z) gives the same eigenpairs as the pdsygst-style path
    at a size that exercises several 2048-blocks of the k-range cuts and an odd order."""
    for n in (1500, 4500):
        A, B = lt.synthetic_pair(n, 20240611 + n)
        out = {}
        for red in (0, 1):
            ctx.set_option("reduction", red)
            w, X = np.zeros(n), np.zeros((n, n), order="F")
            info = ctx.call("ekb200_sygvd", n, n, A.ctypes.data, n, B.ctypes.data, n, w.ctypes.data, X.ctypes.data, n)
            assert info == 0
            out[red] = (w, X)
        ctx.set_option("reduction", 0)
        w0, w1 = out[0][0], out[1][0]
        assert np.max(np.abs(w0 - w1)) <= 1e-12 * np.abs(w0).max()
        r = lt.residual_metrics(A, w1, out[1][1], B)
        o = lt.orthogonality_metrics(out[1][1], B)
        assert r["res_max_over_A"] <= 1e-12 * n and o["orth_fro"] <= 1e-12 * n
