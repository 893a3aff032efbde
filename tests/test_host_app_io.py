"""CPU tests of the host-side contract (formats, argument validation, C-ABI surface)."""
import ctypes
import os
import re

import numpy as np
import pytest

from eigenkernel_b200 import _lib, app_io
from eigenkernel_b200.solver import Argument, validate_argument

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fortran_e_format_known_values():
    assert app_io.fortran_e(-1.121921212197622) == "  -0.1121921212197622E+001"
    assert app_io.fortran_e(0.0) == "   0.0000000000000000E+000"
    assert app_io.fortran_e(1.0) == "   0.1000000000000000E+001"
    assert app_io.fortran_e(9.9999999999999999e-5) == "   0.1000000000000000E-003"
    assert app_io.fortran_e(4.36, 24, 16, 3) == " 0.4360000000000000E+001"
    assert len(app_io.fortran_e(-1e-300)) == 26


@pytest.mark.parametrize("name", ["ELSES_MATRIX_BNZ30_ev.txt", "ELSES_MATRIX_BNZ30_ipr.txt"])
def test_shipped_answer_files_roundtrip_byte_for_byte(golden_dir, name):
    """The shipped answer files are literally eigenvalues.dat / ipratios.dat (main.f90:115-117,139-141):
    parsing and re-formatting them must reproduce every byte."""
    path = os.path.join(golden_dir, name)
    vals = app_io.read_indexed_values(path)
    assert app_io.format_indexed_values(vals) == open(path).read()


def test_matrix_market_reader_matches_fixture_facts(golden_dir):
    fa = os.path.join(golden_dir, "ELSES_MATRIX_BNZ30_A.mtx")
    info = app_io.read_matrix_info(fa)
    assert (info.rep, info.rows, info.cols, info.entries) == ("coordinate", 30, 30, 303)
    m = app_io.read_matrix_file(fa, info)
    assert m.suffix.shape == (303, 2) and m.suffix.min() == 1 and m.suffix.max() == 30
    A = app_io.sparse_to_dense(m)
    assert np.array_equal(A, A.T) and abs(np.linalg.norm(A, "fro") - 5.348) < 1e-3


def test_matrix_market_reader_rejects_out_of_range(tmp_path):
    p = tmp_path / "bad.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real symmetric\n% c\n3 3 2\n1 1 1.0\n4 1 2.0\n")
    with pytest.raises(app_io.TerminateError):
        app_io.read_matrix_file(str(p))


def test_event_logger_accumulates_and_prepends():
    lg = app_io.EventLogger(echo=False)
    lg.add_event("a", 1.0)
    lg.add_event("b", 2.0)
    lg.add_event("a", 0.5)
    assert [e.name for e in lg.events] == ["b", "a"]
    assert lg.find("a").num_repeated == 2 and lg.find("a").val == 1.5
    txt = app_io.log_json_text({"version": "20160808", "dimension": 30}, lg.events)
    assert '"val":  0.2000000000000000E+001' in txt
    import json
    doc = json.loads(re.sub(r"(\d)E([+-])0*(\d)", r"\1E\2\3", txt))
    assert doc["events"][0]["name"] == "b" and doc["setting"]["dimension"] == 30


def _arg(solver, n=10, gen=False, n_vec=-1):
    a = Argument(solver_type=solver, matrix_A_info=app_io.MatrixInfo("coordinate", "real", "symmetric", n, n, n),
                 n_vec=n_vec)
    if gen:
        a.is_generalized_problem = True
        a.matrix_B_info = app_io.MatrixInfo("coordinate", "real", "symmetric", n, n, n)
    return a.finalize()


def test_validate_argument_mirrors_reference_messages():
    validate_argument(_arg("b200"))
    validate_argument(_arg("general_b200", gen=True))
    validate_argument(_arg("b200_select", n_vec=3))
    with pytest.raises(app_io.TerminateError, match="is not for generalized eigenvalue problem"):
        validate_argument(_arg("b200", gen=True))
    with pytest.raises(app_io.TerminateError, match="is not for standard eigenvalue problem"):
        validate_argument(_arg("general_b200"))
    with pytest.raises(app_io.TerminateError, match="does not support partial eigenvalue computation"):
        validate_argument(_arg("b200", n_vec=3))
    with pytest.raises(app_io.TerminateError, match="Unknown solver 'nope'"):
        validate_argument(_arg("nope"))
    bad = _arg("general_b200", gen=True)
    bad.matrix_B_info.rows = 11
    with pytest.raises(app_io.TerminateError, match="Matrix dimension mismatch"):
        validate_argument(bad)


def test_cabi_library_exports_every_declared_symbol():
    """libekb200.so loads without a GPU and exports exactly what include/ekb200.h declares."""
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "ekb200.h")).read()
    declared = sorted(set(re.findall(r"\b(ekb200_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == _lib.exported_symbols()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ekb200_version() >= 100


def test_every_option_key_is_documented_in_the_header():
    """ekb200_set_option's keys (csrc/api.cu) and the list in include/ekb200.h stay in step, both ways."""
    api = open(os.path.join(ROOT, "eigenkernel_b200", "csrc", "api.cu")).read()
    accepted = set(re.findall(r'strcmp\(key, "([a-z0-9_]+)"\)', api))
    hdr = open(os.path.join(ROOT, "include", "ekb200.h")).read()
    decl = hdr[hdr.index("int ekb200_set_option("):hdr.index("int ekb200_version(")]
    documented = set(re.findall(r'"([a-z0-9_]+)"', decl))
    assert accepted == documented, (sorted(accepted - documented), sorted(documented - accepted))


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from eigenkernel_b200.device import Context
    from eigenkernel_b200._lib import Ekb200Error
    with pytest.raises(Ekb200Error):
        Context(0)


def test_built_library_carries_the_blackwell_fp64_path():
    """Static evidence (hard part 1): the sm_100a cubin uses the FP64 tensor path (DMMA.8x8x4), cp.async
    pipelines (LDGSTS) and TMA bulk copies with mbarrier transactions (UBLKCP / SYNCS) -- and nothing was silently
    compiled for another architecture."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    elf = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", elf))
    assert archs == {"100a"}, archs
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert sass.count("DMMA.8x8x4") > 1000
    assert sass.count("LDGSTS") > 100
    assert sass.count("UBLKCP") > 0 and sass.count("SYNCS.ARRIVE.TRANS") > 0


def test_solver_status_codes_are_reported_by_range(capsys):
    """include/ekb200.h: 1..n = info(pdpotrf) (generalized_to_standard.f90:25-30); EKB200_FAIL_STEDC + k =
    info(pdstedc); EKB200_WARN_STEIN + k is a warning only -- the reference reports pdsyevx's IFAIL and carries on
    (solver_scalapack_select.f90:61-67)."""
    from eigenkernel_b200 import solver
    from eigenkernel_b200.app_io import TerminateError

    solver.interpret_info(0, 100, 100, True)
    solver.interpret_info(solver.WARN_STEIN + 3, 100, 40, False)          # no exception
    assert "did not converge for 3 of 40" in capsys.readouterr().out
    with pytest.raises(TerminateError) as e:
        solver.interpret_info(57, 100, 100, True)
    assert e.value.code == 57 and "info(pdpotrf): 57" in capsys.readouterr().out
    with pytest.raises(TerminateError) as e:
        solver.interpret_info(solver.FAIL_STEDC + 2, 100, 100, True)
    assert e.value.code == 2 and "info(pdstedc): 2" in capsys.readouterr().out
    with pytest.raises(TerminateError) as e:
        solver.interpret_info(57, 100, 100, False)                        # a standard solve has no Cholesky step
    assert "info(ekb200_sygvd_coo): 57" in capsys.readouterr().out
    with pytest.raises(TerminateError):
        solver.interpret_info(1000001, 100, 100, True)
    hdr = open(os.path.join(ROOT, "include", "ekb200.h")).read()
    assert f"#define EKB200_WARN_STEIN {solver.WARN_STEIN}" in hdr and f"#define EKB200_FAIL_STEDC {solver.FAIL_STEDC}" in hdr
