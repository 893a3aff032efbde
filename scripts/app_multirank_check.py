"""Rank-per-GPU check of the C++ host twin (app/bin/ekb200_app --ngpu P): runs the reference's two
fixtures and one synthetic generalized problem on P forked ranks and compares the files the app writes with the
shipped answers / the LAPACK twin.  Usage (on a box with P GPUs): python scripts/app_multirank_check.py [P]"""
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eigenkernel_b200 import app_io  # noqa: E402
from oracle import lapack_twin as lt  # noqa: E402

P = sys.argv[1] if len(sys.argv) > 1 else "2"
APP = os.path.join(ROOT, "app", "bin", "ekb200_app")
G = os.path.join(ROOT, "tests", "golden")
ok = True


def run(args, cwd):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "SLURM_PROCID", "SLURM_NTASKS", "PMI_RANK", "PMI_SIZE")}
    r = subprocess.run([APP, "--ngpu", P] + args, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=150)
    if r.returncode != 0:
        print("FAILED rc", r.returncode, r.stdout[-1500:], r.stderr[-1500:])
    return r


def printed(stdout, label):
    m = re.search(re.escape(label) + r"\s*([-+0-9.E]+)", stdout)
    return float(m.group(1)) if m else float("nan")


def verdict(name, cond, **info):
    global ok
    ok = ok and bool(cond)
    print(json.dumps({"case": name, "ranks": int(P), "pass": bool(cond), **info}))


with tempfile.TemporaryDirectory() as t:
    fa, fb = (os.path.join(G, f"ELSES_MATRIX_BNZ30_{x}.mtx") for x in "AB")
    r = run(["-s", "general_b200", "-c", "-1", "-t", "1,30", "-p", "30", "-d", t, fa, fb], t)
    if r.returncode == 0:
        w = app_io.read_indexed_values(os.path.join(t, "eigenvalues.dat"))
        ev = app_io.read_indexed_values(os.path.join(G, "ELSES_MATRIX_BNZ30_ev.txt"))
        ipr = app_io.read_indexed_values(os.path.join(t, "ipratios.dat"))
        ipr_ref = app_io.read_indexed_values(os.path.join(G, "ELSES_MATRIX_BNZ30_ipr.txt"))
        verdict("bnz30 general_b200", np.max(np.abs(w - ev) / np.abs(ev)) <= 1e-12 and
                np.max(np.abs(ipr - ipr_ref) / ipr_ref) <= 1e-6 and printed(r.stdout, "residual norm (max):") <= 3e-11 and
                printed(r.stdout, "orthogonality criterion:") <= 3e-11 and os.path.exists(os.path.join(t, "00000030.dat")),
                dlambda=float(np.max(np.abs(w - ev) / np.abs(ev))), res=printed(r.stdout, "residual norm (max):"),
                orth=printed(r.stdout, "orthogonality criterion:"))
    else:
        verdict("bnz30 general_b200", False)

with tempfile.TemporaryDirectory() as t:
    fa = os.path.join(G, "ELSES_MATRIX_VCNT400std_A.mtx")
    r = run(["-s", "b200", "-c", "-1", "-t", "1,400", "-p", "1,400", "-d", t, fa], t)
    if r.returncode == 0:
        w = app_io.read_indexed_values(os.path.join(t, "eigenvalues.dat"))
        E = app_io.read_indexed_values(os.path.join(G, "ELSES_MATRIX_VCNT400std_E.txt"))
        files = os.path.exists(os.path.join(t, "00000001.dat")) and os.path.exists(os.path.join(t, "00000400.dat"))
        verdict("vcnt400 b200", np.max(np.abs(w - E)) <= 6e-13 and printed(r.stdout, "residual norm (max):") <= 4e-10 and
                printed(r.stdout, "orthogonality criterion:") <= 4e-10 and files, dlambda=float(np.max(np.abs(w - E))),
                res=printed(r.stdout, "residual norm (max):"), orth=printed(r.stdout, "orthogonality criterion:"),
                vector_files_from_both_slabs=files)
    else:
        verdict("vcnt400 b200", False)

with tempfile.TemporaryDirectory() as t:
    n, seed = 3000, 20240601
    r = run(["-s", "general_b200", "-c", "-1", "-t", f"1,{n}", f"synthetic:{n}:{seed}", f"synthetic:{n}:{seed + 1}"], t)
    if r.returncode == 0:
        w = app_io.read_indexed_values(os.path.join(t, "eigenvalues.dat"))
        A, B = lt.synthetic_pair(n, seed)
        w_ref, _, _ = lt.general_scalapack_twin(A, B)
        log = open(os.path.join(t, "log.json")).read()
        verdict("synthetic general_b200 n=3000", np.max(np.abs(w - w_ref)) <= 1e-12 * np.abs(w_ref).max() and
                printed(r.stdout, "residual norm (max):") <= 1e-12 * n and
                printed(r.stdout, "orthogonality criterion:") <= 1e-12 * n and "eigen_solver_b200:sy2sb" in log,
                dlambda=float(np.max(np.abs(w - w_ref)) / np.abs(w_ref).max()), res=printed(r.stdout, "residual norm (max):"),
                orth=printed(r.stdout, "orthogonality criterion:"))
    else:
        verdict("synthetic general_b200 n=3000", False)
print("ALL PASS" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
