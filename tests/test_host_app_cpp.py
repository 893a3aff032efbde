"""CPU tests of the C++ host twin of EigenKernel_App (app/): formats, MatrixMarket reader, argument handling and the
command-line flow up to the solver call.  The formats are checked against the shipped answer files and against the
Python restatement (eigenkernel_b200/app_io.py), the reader against the fixtures of the reference."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

from eigenkernel_b200 import app_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
APP = os.path.join(ROOT, "app", "bin", "ekb200_app")
HOOKS = os.path.join(ROOT, "app", "libekb200_apphooks.so")


@pytest.fixture(scope="module")
def hooks():
    if not (os.path.exists(APP) and os.path.exists(HOOKS)):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(HOOKS)
    lib.ekapp_fortran_e.argtypes = [ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
    lib.ekapp_log_json.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_double),
                                   ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_longlong,
                                   ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
    lib.ekapp_read_matrix.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong),
                                      ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong), ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_longlong, ctypes.c_char_p, ctypes.c_int]
    lib.ekapp_parse_ranges.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_longlong), ctypes.c_int, ctypes.c_char_p,
                                       ctypes.c_int]
    return lib


def _fe(lib, x, w=26, d=16, e=3):
    buf = ctypes.create_string_buffer(128)
    lib.ekapp_fortran_e(x, w, d, e, buf, 128)
    return buf.value.decode()


def _read(lib, path, threads=1, with_body=True):
    rows, cols, ent = ctypes.c_longlong(), ctypes.c_longlong(), ctypes.c_longlong()
    msg = ctypes.create_string_buffer(256)
    rc = lib.ekapp_read_matrix(path.encode(), threads, rows, cols, ent, None, None, 0, msg, 256)
    if rc or not with_body:
        return rc, (rows.value, cols.value, ent.value), None, None, msg.value.decode()
    ij = np.zeros((ent.value, 2), dtype=np.int32)
    v = np.zeros(ent.value)
    rc = lib.ekapp_read_matrix(path.encode(), threads, rows, cols, ent, ij.ctypes.data, v.ctypes.data, ent.value, msg, 256)
    return rc, (rows.value, cols.value, ent.value), ij, v, msg.value.decode()


def test_edit_descriptors_match_known_values_and_python_restatement(hooks):
    assert _fe(hooks, -1.121921212197622) == "  -0.1121921212197622E+001"
    assert _fe(hooks, 0.0) == "   0.0000000000000000E+000"
    assert _fe(hooks, 9.9999999999999999e-5) == "   0.1000000000000000E-003"
    assert _fe(hooks, 4.36, 24, 16, 3) == " 0.4360000000000000E+001"
    assert _fe(hooks, 0.53481162e1, 15, 8, 0) == " 0.53481162E+01"   # the e15.8 of main.f90:153-155
    assert _fe(hooks, 1.5e-120, 15, 8, 0) == " 0.15000000-119"       # three-digit exponent drops the E
    rng = random.Random(7)
    for _ in range(2000):
        x = rng.uniform(-1, 1) * 10.0 ** rng.randint(-300, 300)
        assert _fe(hooks, x) == app_io.fortran_e(x)
        assert _fe(hooks, x, 24, 16, 3) == app_io.fortran_e(x, 24, 16, 3)


@pytest.mark.parametrize("name", ["ELSES_MATRIX_BNZ30_ev.txt", "ELSES_MATRIX_BNZ30_ipr.txt"])
def test_answer_files_reformat_byte_for_byte(hooks, golden_dir, name):
    """eigenvalues.dat / ipratios.dat lines are '(I8, " ", E26.16e3)' (main.f90:115-117,139-141)."""
    lines = open(os.path.join(golden_dir, name)).read().splitlines()
    for ln in lines:
        j, val = ln.split()
        assert f"{int(j):8d} {_fe(hooks, float(val))}" == ln


def test_log_json_matches_python_restatement(hooks):
    names = [b"main:read_command_argument", b"read_matrix_file", b"read_matrix_file", b"main"]
    vals = [1.5e-3, 0.25, 0.5, 12.0]
    arr_n = (ctypes.c_char_p * 4)(*names)
    arr_v = (ctypes.c_double * 4)(*vals)
    buf = ctypes.create_string_buffer(1 << 16)
    hooks.ekapp_log_json(4, arr_n, arr_v, b"bin/eigenkernel_app -s general_b200 A.mtx B.mtx", b"A.mtx", b"B.mtx",
                         b"general_b200", 30, 0, buf, 1 << 16)
    lg = app_io.EventLogger(echo=False)
    for n, v in zip(names, vals):
        lg.add_event(n.decode(), v)
    setting = {"version": "20160808", "command": "bin/eigenkernel_app -s general_b200 A.mtx B.mtx",
               "matrix_A_filename": "A.mtx", "matrix_B_filename": "B.mtx", "log_filename": "log.json", "dimension": 30,
               "solver": "general_b200", "g_block_size": 64, "block_size": 0}
    assert buf.value.decode() == app_io.log_json_text(setting, lg.events)


def test_matrix_market_reader_matches_python_reader_on_the_fixtures(hooks, golden_dir):
    for name, dims in (("ELSES_MATRIX_BNZ30_A.mtx", (30, 30, 303)), ("ELSES_MATRIX_BNZ30_B.mtx", (30, 30, 303)),
                       ("ELSES_MATRIX_VCNT400std_A.mtx", None)):
        path = os.path.join(golden_dir, name)
        rc, got, ij, v, _ = _read(hooks, path)
        assert rc == 0
        ref = app_io.read_matrix_file(path)
        if dims:
            assert got == dims
        assert got == (ref.size, ref.size, ref.num_non_zeros)
        assert np.array_equal(ij, ref.suffix) and np.array_equal(v, ref.value)


def test_parallel_parser_is_identical_to_serial(hooks, tmp_path):
    """the body is split at record boundaries and parsed by several threads."""
    rng = np.random.default_rng(3)
    n = 420
    ii, jj = np.tril_indices(n)
    vals = rng.standard_normal(ii.size)
    p = tmp_path / "big.mtx"
    with open(p, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real symmetric\n% generated\n%\n")
        f.write(f"{n} {n} {ii.size}\n")
        for q, (i, j, x) in enumerate(zip(ii, jj, vals)):
            if q % 1000 == 7:
                f.write("\n")                                            # empty records are skipped by list-directed input
            if q % 3 == 0:
                f.write(f"  {i + 1}   {j + 1}  {x:.17e}\n")
            elif q % 3 == 1:
                f.write(f"{i + 1},{j + 1},{x:.17e}".replace("e", "D") + "\n")  # commas and D exponents
            else:
                f.write(f"{i + 1}\t{j + 1}\t{float(x)!r} trailing ignored\n")
    assert os.path.getsize(p) > (1 << 20)
    rc1, d1, ij1, v1, _ = _read(hooks, str(p), threads=1)
    rc4, d4, ij4, v4, _ = _read(hooks, str(p), threads=5)
    assert rc1 == 0 and rc4 == 0 and d1 == d4 == (n, n, ii.size)
    assert np.array_equal(ij1, ij4) and np.array_equal(v1, v4)
    assert np.array_equal(ij1[:, 0], ii + 1) and np.array_equal(ij1[:, 1], jj + 1) and np.array_equal(v1, vals)


def test_reader_error_behaviour(hooks, tmp_path):
    bad = tmp_path / "range.mtx"
    bad.write_text("%%MatrixMarket matrix coordinate real symmetric\n% c\n3 3 2\n1 1 1.0\n4 1 2.0\n")
    rc, _, _, _, msg = _read(hooks, str(bad))
    assert rc >= 1000 and "index of matrix out of range" in msg
    short = tmp_path / "short.mtx"
    short.write_text("%%MatrixMarket matrix coordinate real symmetric\n3 3 3\n1 1 1.0\n2 1 2.0\n")
    rc, _, _, _, msg = _read(hooks, str(short))
    assert rc >= 1000 and "invalid format of matrix value" in msg
    junk = tmp_path / "junk.mtx"
    junk.write_text("%%MatrixMarket matrix coordinate real symmetric\n3 3 1\n1 x 1.0\n")
    rc, _, _, _, msg = _read(hooks, str(junk))
    assert rc >= 1000 and "invalid format of matrix value" in msg
    # mminfo's own codes (mmio.f:413-583)
    for text, code in (("%%NotMatrixMarket matrix coordinate real symmetric\n1 1 1\n", 7),
                       ("%%MatrixMarket tensor coordinate real symmetric\n1 1 1\n", 1),
                       ("%%MatrixMarket matrix sparse real symmetric\n1 1 1\n", 8),
                       ("%%MatrixMarket matrix coordinate quaternion symmetric\n1 1 1\n", 9),
                       ("%%MatrixMarket matrix coordinate real diagonal\n1 1 1\n", 11),
                       ("%%MatrixMarket matrix coordinate real symmetric\n3 3\n", 6),
                       ("%%MatrixMarket matrix coordinate real symmetric\n% only comments\n", 4)):
        p = tmp_path / "hdr.mtx"
        p.write_text(text)
        rc, _, _, _, _ = _read(hooks, str(p), with_body=False)
        assert rc == code, (text, rc)
    rc, _, _, _, _ = _read(hooks, str(tmp_path / "missing.mtx"), with_body=False)
    assert rc != 0


def test_printed_vecs_ranges_parser(hooks):
    out = (ctypes.c_longlong * 400)()
    msg = ctypes.create_string_buffer(256)
    for spec in ("3", "1-4", "1-4,7,9-12", "5,6"):
        n = hooks.ekapp_parse_ranges(spec.encode(), out, 200, msg, 256)
        assert [(out[2 * i], out[2 * i + 1]) for i in range(n)] == app_io.parse_printed_vecs_ranges(spec)
    assert hooks.ekapp_parse_ranges(b",3", out, 200, msg, 256) == -1 and b"invalid comma placement" in msg.value
    assert hooks.ekapp_parse_ranges(b"-3", out, 200, msg, 256) == -1 and b"invalid hyphen placement" in msg.value
    many = ",".join(str(i) for i in range(1, 102))
    assert hooks.ekapp_parse_ranges(many.encode(), out, 200, msg, 256) == -1 and b"too many ranges" in msg.value


def test_eigenvector_files_match_python_restatement(hooks, tmp_path):
    """print_eigenvectors (matrix_io.f90:173-285): <dir>/<j:08d>.dat, text '(I8," ",I8," ",E26.16e3)' or one Fortran
    unformatted record; the threaded writer produces the same bytes as the Python restatement."""
    hooks.ekapp_print_eigenvectors.argtypes = [ctypes.c_char_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_void_p,
                                               ctypes.c_longlong, ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_char_p, ctypes.c_int]
    rng = np.random.default_rng(2)
    n, k = 257, 12
    X = np.asfortranarray(rng.standard_normal((n, k)) * 10.0 ** rng.integers(-200, 200, (n, k)))
    msg = ctypes.create_string_buffer(256)
    for binary, threads in ((0, 1), (0, 4), (1, 3)):
        d = tmp_path / f"out_{binary}_{threads}"
        d.mkdir()
        rc = hooks.ekapp_print_eigenvectors(str(d).encode(), n, k, X.ctypes.data, n, b"1-3,7,12", binary, threads, msg, 256)
        assert rc == 0, msg.value
        assert sorted(os.listdir(d)) == [f"{j:08d}.dat" for j in (1, 2, 3, 7, 12)]
        ref = tmp_path / f"ref_{binary}_{threads}"
        ref.mkdir()
        for j in (1, 2, 3, 7, 12):
            app_io.write_eigenvector(str(ref), j, X[:, j - 1], binary=bool(binary))
            assert open(d / f"{j:08d}.dat", "rb").read() == open(ref / f"{j:08d}.dat", "rb").read()
    rc = hooks.ekapp_print_eigenvectors(str(tmp_path / "no_such_dir").encode(), n, k, X.ctypes.data, n, b"1", 0, 1, msg, 256)
    assert rc >= 1000 and b"print_eigenvectors: cannot open" in msg.value


def test_parser_fuzz_against_python_reader(hooks, tmp_path):
    """Random small MatrixMarket files (mixed separators, signs, exponents, comment blocks): the C++ reader and the
    Python restatement of read_matrix_file agree entry for entry."""
    rng = np.random.default_rng(99)
    for trial in range(40):
        n = int(rng.integers(1, 40))
        nnz = int(rng.integers(1, 3 * n + 2))
        ii = rng.integers(1, n + 1, nnz)
        jj = rng.integers(1, n + 1, nnz)
        vv = rng.standard_normal(nnz) * 10.0 ** rng.integers(-30, 30, nnz)
        p = tmp_path / f"f{trial}.mtx"
        with open(p, "w") as f:
            f.write("%%MatrixMarket matrix coordinate real symmetric\n")
            for _ in range(int(rng.integers(0, 4))):
                f.write("% comment line\n")
            f.write(f" {n}  {n} {nnz}\n")
            for i, j, v in zip(ii, jj, vv):
                style = int(rng.integers(0, 4))
                if style == 0:
                    f.write(f"{i} {j} {v:.17g}\n")
                elif style == 1:
                    f.write(f"   {i}\t{j}   {v:+.16e}\n")
                elif style == 2:
                    f.write(f"{i} {j} " + f"{v:.16e}".replace("e", "D") + "\n")
                else:
                    f.write(f"{i} {j} {v:.17g}   \r\n")
        rc, dims, ij, v, msg = _read(hooks, str(p), threads=int(rng.integers(1, 4)))
        assert rc == 0, msg
        ref = app_io.read_matrix_file(str(p))
        assert dims == (n, n, nnz) and np.array_equal(ij, ref.suffix) and np.array_equal(v, ref.value)


_LAUNCHER_VARS = ("OMPI_COMM_WORLD_RANK", "OMPI_COMM_WORLD_SIZE", "PMI_RANK", "PMI_SIZE", "PMIX_RANK", "PMIX_SIZE",
                  "SLURM_PROCID", "SLURM_NTASKS", "RANK", "WORLD_SIZE")


def _clean_env():
    """The app takes its ranks from a launcher's environment when one is set; the tests start single processes."""
    return {k: v for k, v in os.environ.items() if k not in _LAUNCHER_VARS}


def _run(args, cwd):
    return subprocess.run([APP] + args, cwd=cwd, env=_clean_env(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                          timeout=120)


def test_cli_dry_run_prints_the_reference_configuration_block(hooks, golden_dir, tmp_path):
    fa, fb = (os.path.join(golden_dir, f"ELSES_MATRIX_BNZ30_{x}.mtx") for x in "AB")
    r = _run(["-s", "general_b200", "-c", "-1", "--dry-run", fa, fb], tmp_path)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    for line in ("---------- Eigen Test start ----------", "----- Configurations -----", "problem type: generalized",
                 f"matrix A file: {fa}", "matrix A field: real", "matrix A symm: symmetric", "matrix A rows: 30",
                 "matrix B entries: 303", "solver: general_b200", "eigenvalues output file: eigenvalues.dat",
                 "ipratios output file: ipratios.dat", "required eigenpairs: 30", "verified eigenpairs: 30",
                 "log output file: log.json", "MPI processes: 1", "dry run mode, exit"):
        assert line in out, line
    assert "[Event" in r.stderr and "main:read_matrix_files" in r.stderr
    assert not os.path.exists(tmp_path / "eigenvalues.dat")


def test_cli_forked_ranks_dry_run(hooks, golden_dir, tmp_path):
    """--ngpu P forks the ranks before CUDA is touched; only the master prints (check_master)."""
    fa = os.path.join(golden_dir, "ELSES_MATRIX_VCNT400std_A.mtx")
    r = _run(["-s", "b200", "--ngpu", "2", "--dry-run", fa], tmp_path)
    assert r.returncode == 0, r.stderr
    assert r.stdout.count("Eigen Test start") == 1 and "MPI processes: 2" in r.stdout


def test_cli_ranks_from_an_external_launcher(hooks, golden_dir, tmp_path):
    """mpirun / srun / torchrun --no-python export the rank and the world size; the ranks then meet on a file-backed
    board (EKB200_RENDEZVOUS, by default /dev/shm/ekb200_<job id>_<launcher pid>) instead of being forked."""
    fa = os.path.join(golden_dir, "ELSES_MATRIX_VCNT400std_A.mtx")
    board = tmp_path / "board.bin"
    procs = []
    for var_r, var_s, r in (("RANK", "WORLD_SIZE", 0), ("OMPI_COMM_WORLD_RANK", "OMPI_COMM_WORLD_SIZE", 1)):
        env = dict(_clean_env(), EKB200_RENDEZVOUS=str(board))
        env[var_r], env[var_s] = str(r), "2"
        procs.append(subprocess.Popen([APP, "-s", "b200", "--dry-run", fa], cwd=tmp_path, env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=60) for p in procs]
    assert [p.returncode for p in procs] == [0, 0], outs
    assert outs[0][0].count("Eigen Test start") == 1 and "MPI processes: 2" in outs[0][0]
    assert outs[1][0].strip() == ""                       # only the master prints (check_master)
    assert not board.exists()                             # rank 0 removes the board of the launch


def test_cli_launcher_variables_alone_do_not_make_a_multi_rank_job(hooks, golden_dir, tmp_path):
    """An sbatch script that runs the binary once (no srun) still has SLURM_PROCID / SLURM_NTASKS in its environment;
    so does anything started under torchrun.  Attachment to an external launcher is opt-in (EKB200_EXTERNAL_RANKS=1 or
    EKB200_RENDEZVOUS): without it the run is a plain single-rank run and must not wait for peers that do not exist."""
    fa = os.path.join(golden_dir, "ELSES_MATRIX_VCNT400std_A.mtx")
    env = dict(_clean_env(), SLURM_PROCID="0", SLURM_NTASKS="4", RANK="0", WORLD_SIZE="4")
    r = subprocess.run([APP, "-s", "b200", "--dry-run", fa], cwd=tmp_path, env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=30)
    assert r.returncode == 0, r.stderr
    assert "MPI processes: 1" in r.stdout


def test_cli_external_launcher_missing_peer_times_out_with_a_message(hooks, golden_dir, tmp_path):
    """Opted in, but the peer never shows up: the barrier has a deadline, the job ends with a clear message and the
    board of the launch is removed."""
    fa = os.path.join(golden_dir, "ELSES_MATRIX_VCNT400std_A.mtx")
    board = tmp_path / "board.bin"
    env = dict(_clean_env(), EKB200_RENDEZVOUS=str(board), EKB200_RENDEZVOUS_TIMEOUT="2", RANK="0", WORLD_SIZE="2")
    r = subprocess.run([APP, "-s", "b200", "--dry-run", fa], cwd=tmp_path, env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=30)
    assert r.returncode == 1
    assert "rendezvous of 2 ranks timed out" in r.stderr
    assert not board.exists()
    # a rank > 0 whose rank 0 never creates the board gives up the same way
    env["RANK"] = "1"
    r = subprocess.run([APP, "-s", "b200", "--dry-run", fa], cwd=tmp_path, env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=30)
    assert r.returncode == 1 and "no board from rank 0" in r.stderr


def test_cli_external_launcher_ignores_stale_board_and_symlinks(hooks, golden_dir, tmp_path):
    """A crashed launch leaves id_ready = 1, an old NCCL id and a non-zero barrier count behind; a hostile user may plant
    a symlink at the predictable name.  Rank 0 creates the board exclusively (O_EXCL | O_NOFOLLOW, stale files
    unlinked), the other ranks only accept a board whose creator is alive."""
    import struct

    fa = os.path.join(golden_dir, "ELSES_MATRIX_VCNT400std_A.mtx")
    board = tmp_path / "board.bin"
    victim = tmp_path / "victim.txt"
    victim.write_text("do not clobber")
    for plant in ("stale", "symlink"):
        if plant == "stale":  # magic ok, dead creator (pid 2^22 + 5 does not exist), id_ready = 1, barrier mid-way
            board.write_bytes(struct.pack("<Iiiiii", 0x454B4232, 4194309, 2, 1, 7, 1) + b"\xff" * 200)
        else:
            os.symlink(victim, board)
        procs = []
        for r in (1, 0):  # rank 1 first: it must not latch onto the leftover
            env = dict(_clean_env(), EKB200_RENDEZVOUS=str(board), EKB200_RENDEZVOUS_TIMEOUT="20", RANK=str(r),
                       WORLD_SIZE="2")
            procs.append(subprocess.Popen([APP, "-s", "b200", "--dry-run", fa], cwd=tmp_path, env=env,
                                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
            if r == 1:
                import time
                time.sleep(0.3)
        outs = [p.communicate(timeout=60) for p in procs]
        assert [p.returncode for p in procs] == [0, 0], (plant, outs)
        assert "MPI processes: 2" in outs[1][0]
        assert not board.exists() and not os.path.islink(board)
        assert victim.read_text() == "do not clobber"


@pytest.mark.parametrize("args,code,needle", [
    (["-s", "nope", "A"], 1, "[Error] validate_argument: Unknown solver 'nope'"),
    (["-s", "b200", "A", "B"], 1, "[Error] validate_argument: solver 'b200' is not for generalized eigenvalue problem"),
    (["-s", "general_b200", "A"], 1, "[Error] validate_argument: solver 'general_b200' is not for standard eigenvalue problem"),
    (["-s", "general_b200", "-n", "3", "A", "B"], 1, "does not support partial eigenvalue computation"),
    (["-s", "general_scalapack", "A", "B"], 1, "is not supported in this build"),
    (["-s", "general_b200", "-c", "31", "A", "B"], 1, "Specified numbers with -c option are not valid"),
    (["-s", "general_b200", "-t", "5", "A", "B"], 1, "wrong format for -t option"),
    (["-s", "general_b200", "-t", "5,31", "A", "B"], 1, "Specified numbers with -t option are not valid"),
    (["-s", "general_b200", "-p", "1-31", "A", "B"], 1, "Specified numbers with -p option are not valid"),
    (["-s", "general_b200", "-z", "A", "B"], 1, "[Error] read_command_argument: unknown option-z"),
    (["-s", "general_b200"], 1, "[Error] read_command_argument: Matrix A file not specified"),
    (["-h"], 0, "[Info] read_command_argument: help printed"),
    (["-s", "b200", "missing.mtx"], None, "[Error] mminfo missing.mtx failed"),
])
def test_cli_error_behaviour_follows_terminate(hooks, golden_dir, tmp_path, args, code, needle):
    fa, fb = (os.path.join(golden_dir, f"ELSES_MATRIX_BNZ30_{x}.mtx") for x in "AB")
    args = [fa if a == "A" else fb if a == "B" else a for a in args]
    r = _run(args + ["--dry-run"], tmp_path)
    if code is None:
        assert r.returncode != 0
    else:
        assert r.returncode == code
    assert needle in r.stderr, r.stderr
    if "-h" in args:
        assert "Usage: eigen_test -s <solver_type> <options> <matrix_A> [<matrix_B>]" in r.stdout


def test_cli_synthetic_spec_and_dimension_mismatch(hooks, golden_dir, tmp_path):
    r = _run(["-s", "general_b200", "--dry-run", "synthetic:512:20240601", "synthetic:512:20240602"], tmp_path)
    assert r.returncode == 0 and "matrix A rows: 512" in r.stdout and "required eigenpairs: 512" in r.stdout
    r = _run(["-s", "general_b200", "--dry-run", "synthetic:512:1", "synthetic:256:2"], tmp_path)
    assert r.returncode == 1 and "Matrix dimension mismatch" in r.stderr


def test_cli_solver_call_fails_loudly_without_a_gpu(hooks, golden_dir, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    fa, fb = (os.path.join(golden_dir, f"ELSES_MATRIX_BNZ30_{x}.mtx") for x in "AB")
    r = _run(["-s", "general_b200", fa, fb], tmp_path)
    assert r.returncode != 0
    assert "[Error] solver_b200: no usable CUDA device (there is no CPU fallback)" in r.stderr
    assert " main:read_matrix_files" in r.stdout          # terminate prints the events first (processes.f90:128-131)
    assert not os.path.exists(tmp_path / "eigenvalues.dat")
