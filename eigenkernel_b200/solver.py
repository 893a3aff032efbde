"""Host-side mirror of EigenKernel's solver boundary for the B200 solvers.

`eigen_solver(arg, matrix_A, matrix_B)` is reference src/solver_main.f90:22-100 with five new cases
(`b200`, `b200_select`, `general_b200`, `general_b200_select`, `general_b200inv`) next to the existing names; the argument
checks are `validate_argument` (src/command_argument.f90:121-219) extended by the same names; the result is
the type-2 `eigenpairs` container of src/eigenpairs_types.f90:7-11 on a 1x1 grid (the local array IS the
matrix).  All arithmetic happens in libekb200.so (hand-written CUDA); there is no CPU fallback: without the
library or without a GPU every solver raises.
"""
from __future__ import annotations

import ctypes
import time
from dataclasses import dataclass, field

import numpy as np

from .app_io import EventLogger, MatrixInfo, SparseMat, TerminateError
from .device import Context

G_BLOCK_SIZE = 64          # global_variables.f90:5
WARN_STEIN, FAIL_STEDC = 500000, 600000  # EKB200_WARN_STEIN / EKB200_FAIL_STEDC (include/ekb200.h)
G_VERSION = "20160808"     # global_variables.f90:6

STANDARD_SOLVERS = ("b200", "b200_select")
GENERALIZED_SOLVERS = ("general_b200", "general_b200_select", "general_b200inv")
SELECT_SOLVERS = ("b200_select", "general_b200_select")
# names of the reference that this build does not provide (they abort like the *_dummy.f90 twins)
REFERENCE_ONLY_SOLVERS = (
    "lapack", "scalapack", "scalapack_select", "general_scalapack", "general_scalapack_select", "eigensx",
    "general_scalapack_eigensx", "general_scalapack_eigens", "general_elpa_scalapack", "general_elpa1",
    "general_elpa2", "general_elpa_eigensx", "general_elpa_eigens", "general_scalapacknew_eigens")


@dataclass
class Argument:
    """ek_argument_t (command_argument.f90:20-44), the fields the solver boundary reads."""
    solver_type: str = ""
    matrix_A_info: MatrixInfo = field(default_factory=MatrixInfo)
    matrix_B_info: MatrixInfo = field(default_factory=MatrixInfo)
    is_generalized_problem: bool = False
    block_size: int = 0
    n_vec: int = -1
    n_check_vec: int = 0
    ortho_check_index_start: int = 0
    ortho_check_index_end: int = 0
    printed_vecs_ranges: list = field(default_factory=list)

    def finalize(self) -> "Argument":
        """read_command_argument's defaults (command_argument.f90:446-452)."""
        if self.n_vec == -1:
            self.n_vec = self.matrix_A_info.rows
        if self.n_check_vec == -1:
            self.n_check_vec = self.n_vec
        return self


@dataclass
class Process:
    """ek_process_t (processes.f90:6-9): the grid is 1x1, one B200 behind it."""
    my_rank: int = 0
    n_procs: int = 1
    n_procs_row: int = 1
    n_procs_col: int = 1
    context: int = 0
    device: int = 0


@dataclass
class EigenpairsBlacs:
    values: np.ndarray | None = None
    desc: np.ndarray | None = None
    Vectors: np.ndarray | None = None


@dataclass
class Eigenpairs:
    """ek_eigenpairs_types_union_t (eigenpairs_types.f90:13-17); the B200 solvers fill type 2."""
    type_number: int = 0
    blacs: EigenpairsBlacs = field(default_factory=EigenpairsBlacs)


def validate_argument(arg: Argument) -> None:
    """command_argument.f90:121-219 with the b200 names added (same messages, same order of checks)."""
    dim = arg.matrix_A_info.rows
    ok = dim == arg.matrix_A_info.cols
    if arg.is_generalized_problem:
        ok = ok and dim == arg.matrix_B_info.rows and dim == arg.matrix_B_info.cols
    if not ok:
        raise TerminateError("validate_argument: Matrix dimension mismatch", 1)
    st = arg.solver_type.strip()
    if st in STANDARD_SOLVERS:
        valid = not arg.is_generalized_problem
    elif st in GENERALIZED_SOLVERS:
        valid = arg.is_generalized_problem
    elif st in REFERENCE_ONLY_SOLVERS:
        raise TerminateError(f"eigen_solver: solver '{st}' is not supported in this build", 1)
    else:
        raise TerminateError(f"validate_argument: Unknown solver '{st}'", 1)
    if not valid:
        kind = "generalized" if arg.is_generalized_problem else "standard"
        raise TerminateError(f"validate_argument: solver '{st}' is not for {kind} eigenvalue problem", 1)
    if st not in SELECT_SOLVERS and arg.n_vec != dim:
        raise TerminateError(f"validate_argument: Solver '{st}' does not support partial eigenvalue computation", 1)
    if st in SELECT_SOLVERS and not (0 < arg.n_vec <= dim):
        raise TerminateError("validate_argument: Specified number with -n option is not valid", 1)
    for a, b in arg.printed_vecs_ranges:
        if a < 0 or b < 0 or b > arg.n_vec or a > b:
            raise TerminateError("validate_argument: Specified numbers with -p option are not valid", 1)
    if arg.n_check_vec < 0 or arg.n_check_vec > arg.n_vec:
        raise TerminateError("validate_argument: Specified numbers with -c option are not valid", 1)
    if (arg.ortho_check_index_start < 0 or arg.ortho_check_index_end < 0 or
            arg.ortho_check_index_end > arg.n_vec or arg.ortho_check_index_start > arg.ortho_check_index_end):
        raise TerminateError("validate_argument: Specified numbers with -t option are not valid", 1)


def _coo_ptrs(m: SparseMat):
    ij = np.ascontiguousarray(m.suffix, dtype=np.int32)
    v = np.ascontiguousarray(m.value, dtype=np.float64)
    return ij, v


def interpret_info(info: int, n: int, n_vec: int, generalized: bool) -> None:
    """Status code of a whole-solve entry point -> the reference's reporting (generalized_to_standard.f90:25-30,
    solver_scalapack_select.f90:61-67).  The routine name follows the RANGE of the code (include/ekb200.h), not the
    kind of problem: 1..n = info(pdpotrf); FAIL_STEDC + k = info(pdstedc); WARN_STEIN + k is only a warning (k
    eigenvectors did not converge in inverse iteration; the results are complete), like pdsyevx's IFAIL report."""
    if WARN_STEIN < info < FAIL_STEDC:
        print(f"[Warning] eigen_solver_b200_select: inverse iteration did not converge for {info - WARN_STEIN} "
              f"of {n_vec} requested eigenvectors")
        return
    if info == 0:
        return
    if FAIL_STEDC < info < 1000000:
        routine, info = "pdstedc", info - FAIL_STEDC
    elif generalized and 0 < info <= n:
        routine = "pdpotrf"
    else:
        routine = "ekb200_sygvd_coo"
    print(f"info({routine}): {info}")
    raise TerminateError(f"eigen_solver: {routine} failed", info)


def eigen_solver(arg: Argument, matrix_A: SparseMat, matrix_B: SparseMat | None = None, *,
                 logger: EventLogger | None = None, ctx: Context | None = None, device: int = 0):
    """solver_main.f90:22-100 for the b200 cases.  Returns (eigenpairs, proc).

    Errors follow the reference: unknown names -> terminate('eigen_solver: Unknown solver', 1);
    library status codes -> `info(<routine>): N` + terminate (generalized_to_standard.f90:25-30)."""
    st = arg.solver_type.strip()
    n = arg.matrix_A_info.rows
    generalized = st in GENERALIZED_SOLVERS
    if st not in STANDARD_SOLVERS + GENERALIZED_SOLVERS:
        raise TerminateError("eigen_solver: Unknown solver", 1)
    if generalized and matrix_B is None:
        raise TerminateError(f"eigen_solver: solver '{st}' needs matrix B", 1)
    n_vec = arg.n_vec if st in SELECT_SOLVERS else n
    own = ctx is None
    if own:
        ctx = Context(device)  # raises loudly without the CUDA library / a GPU
    try:
        ctx.clear_events()
        if arg.block_size > 0:
            # g_block_size := --block-size (solver_main.f90:44-46); the device band width follows it when legal
            if arg.block_size in (32, 64):
                ctx.set_option("band", arg.block_size)
        # general_b200inv: explicit-inverse reduction (the ELPA-style workflow, solver_elpa_eigenexa.f90:110-150)
        ctx.set_option("reduction", 1 if st == "general_b200inv" else 0)
        t0 = time.perf_counter()
        ijA, vA = _coo_ptrs(matrix_A)
        if generalized:
            ijB, vB = _coo_ptrs(matrix_B)
            nnzB, pijB, pvB = matrix_B.num_non_zeros, ijB.ctypes.data, vB.ctypes.data
        else:
            nnzB, pijB, pvB = 0, None, None
        values = np.zeros(max(n, 1))
        vectors = np.zeros((n, max(n_vec, 1)), order="F")
        info = ctx.call("ekb200_sygvd_coo", n, n_vec, matrix_A.num_non_zeros, ijA.ctypes.data, vA.ctypes.data,
                        nnzB, pijB, pvB, values.ctypes.data, vectors.ctypes.data, max(n, 1))
        wall = time.perf_counter() - t0
        if logger is not None:
            for name, sec, rep in ctx.events():
                for _ in range(max(rep, 1) - 1):
                    logger.add_event(name, 0.0, to_print=False)
                logger.add_event(name, sec)
            logger.add_event("eigen_solver_b200:wall", wall)
        interpret_info(info, n, n_vec, generalized)
    finally:
        if own:
            ctx.close()
    ep = Eigenpairs(type_number=2)
    ep.blacs.values = values[:n]
    nb = arg.block_size if arg.block_size > 0 else G_BLOCK_SIZE
    nb = max(min(nb, n), 1)  # setup_distributed_matrix clamps the block size (distribute_matrix.f90:114-120)
    # desc = [dtype_=1, ctxt, M, N, MB, NB, RSRC, CSRC, LLD] (descriptor_parameters.f90:2-4)
    ep.blacs.desc = np.array([1, 0, n, n_vec, nb, nb, 0, 0, max(n, 1)], dtype=np.int32)
    ep.blacs.Vectors = vectors[:, :n_vec]
    return ep, Process(device=device)
