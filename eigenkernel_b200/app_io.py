"""Host-side I/O contract of EigenKernel_App, restated for the B200 solvers (formats must not change):

  MatrixMarket input      reference src/matrix_io.f90:72-144, src/mmio.f:341-585 (header probe only)
  eigenvalues.dat         src/main.f90:111-121          '(I8, " ", E26.16e3)'
  ipratios.dat            src/main.f90:131-143          same format, desc(cols_) lines
  eigenvector files       src/matrix_io.f90:173-285     <dir>/<j:08d>.dat, '(I8," ",I8," ",E26.16e3)'
  event logger / log.json src/event_logger.f90:23-141, src/fson.f90:454-553 (printer byte format)

Pure host code (numpy only); no GPU needed.
"""
from __future__ import annotations

import math
import os
import sys
import time
from dataclasses import dataclass, field

import numpy as np


class TerminateError(RuntimeError):
    """What `terminate(msg, code)` (src/processes.f90:122-139) is on the reference: the run stops with
    `[Error] msg`; here it is an exception carrying the same message and code."""

    def __init__(self, msg: str, code: int = 1):
        self.msg, self.code = msg, int(code)
        super().__init__(f"[Error] {msg}")


# ------------------------------------------------------------------------------------------ Fortran edit descriptors
def fortran_e(x: float, width: int = 26, digits: int = 16, expw: int = 3) -> str:
    """Fortran `Ew.dEe` as gfortran prints it: mantissa in [0.1, 1), `digits` decimals, `expw` exponent
    digits, right-aligned in `width` (E26.16e3 -> '   -0.1121921212197622E+001')."""
    x = float(x)
    if math.isnan(x):
        return "NaN".rjust(width)
    if math.isinf(x):
        return ("-Infinity" if x < 0 else "Infinity").rjust(width)
    if x == 0.0:
        body = "0." + "0" * digits + "E+" + "0" * expw
        if math.copysign(1.0, x) < 0:
            body = "-" + body
        return body.rjust(width)
    s = f"{abs(x):.{digits - 1}e}"  # d.ddd..e+XX with `digits` significant digits
    mant, ex = s.split("e")
    e10 = int(ex) + 1
    body = "0." + mant.replace(".", "") + "E" + ("+" if e10 >= 0 else "-") + f"{abs(e10):0{expw}d}"
    if x < 0:
        body = "-" + body
    return body.rjust(width)


def format_indexed_values(values) -> str:
    """The body of eigenvalues.dat / ipratios.dat (main.f90:115-117,139-141)."""
    return "".join(f"{j:8d} {fortran_e(v)}\n" for j, v in enumerate(values, start=1))


def write_eigenvalues(path: str, values, n_vec: int) -> None:
    with open(path, "w") as f:
        f.write(format_indexed_values(np.asarray(values)[:n_vec]))


def write_ipratios(path: str, ipratios) -> None:
    with open(path, "w") as f:
        f.write(format_indexed_values(ipratios))


def read_indexed_values(path: str) -> np.ndarray:
    return np.array([float(l.split()[1]) for l in open(path) if l.strip()])


def write_eigenvector(dirname: str, j: int, vec, binary: bool = False) -> str:
    """print_vector (matrix_io.f90:233-285): one file per eigenvector index j (1-based)."""
    path = os.path.join(dirname, f"{j:08d}.dat")
    vec = np.asarray(vec, dtype=np.float64)
    if binary:
        # Fortran unformatted sequential record: 4-byte length markers around the payload (gfortran)
        payload = vec.tobytes()
        mark = np.int32(len(payload)).tobytes()
        with open(path, "wb") as f:
            f.write(mark + payload + mark)
    else:
        with open(path, "w") as f:
            f.write("".join(f"{i:8d} {j:8d} {fortran_e(v)}\n" for i, v in enumerate(vec, start=1)))
    return path


def parse_printed_vecs_ranges(spec: str):
    """`-p a[-b][,c[-d]]...` (command_argument.f90:271-315); at most 100 ranges."""
    out = []
    for part in spec.split(","):
        if "-" in part:
            a, b = part.split("-", 1)
            out.append((int(a), int(b)))
        else:
            out.append((int(part), int(part)))
    if len(out) > 100:
        raise TerminateError("read_command_argument: too many ranges with -p option", 1)
    return out


# ------------------------------------------------------------------------------------------ MatrixMarket
@dataclass
class MatrixInfo:
    """ek_matrix_info_t (command_argument.f90:12-18)."""
    rep: str = ""
    field: str = ""
    symm: str = ""
    rows: int = 0
    cols: int = 0
    entries: int = 0


@dataclass
class SparseMat:
    """ek_sparse_mat_t (matrix_io.f90:11-15): replicated COO, 1-based, one triangle stored.
    `suffix` is (2, nnz) int32 in Fortran order, i.e. C-contiguous (nnz, 2) pairs (i, j)."""
    size: int
    num_non_zeros: int
    value: np.ndarray
    suffix: np.ndarray


def read_matrix_info(path: str) -> MatrixInfo:
    """mminfo (mmio.f:341-585 via command_argument.f90:89-103): banner + size line."""
    with open(path) as f:
        banner = f.readline().split()
        if len(banner) < 5 or banner[0].lower() != "%%matrixmarket" or banner[1].lower() != "matrix":
            raise TerminateError(f"read_command_argument: mminfo failed for {path}", 1)
        info = MatrixInfo(rep=banner[2].lower(), field=banner[3].lower(), symm=banner[4].lower())
        line = f.readline()
        while line and (line.startswith("%") or not line.strip()):
            line = f.readline()
        words = line.split()
        if len(words) != 3:
            raise TerminateError(f"read_command_argument: mminfo failed for {path}", 1)
        info.rows, info.cols, info.entries = (int(x) for x in words)
    return info


def read_matrix_file(path: str, info: MatrixInfo | None = None) -> SparseMat:
    """read_matrix_file (matrix_io.f90:22-144): coordinate body, list-directed `i j value`, range-checked.
    Symmetry is assumed, not read from the header (distribute_matrix.f90:411-418 mirrors every entry)."""
    info = info or read_matrix_info(path)
    if info.rep != "coordinate":
        raise TerminateError("read_matrix_file: only coordinate format is supported", 1)
    ij = np.empty((info.entries, 2), dtype=np.int32)
    v = np.empty(info.entries, dtype=np.float64)
    with open(path) as f:
        f.readline()
        line = f.readline()
        while line.startswith("%") or not line.strip():
            line = f.readline()
        for t in range(info.entries):
            w = f.readline().replace(",", " ").split()
            if len(w) < 3:
                raise TerminateError("read_matrix_file: unexpected end of file", 1)
            i, j = int(w[0]), int(w[1])
            if i < 1 or i > info.rows or j < 1 or j > info.cols:
                raise TerminateError("read_matrix_file_value: index of matrix out of range", 1)
            ij[t, 0], ij[t, 1] = i, j
            v[t] = float(w[2].replace("D", "E").replace("d", "e"))
    return SparseMat(size=info.rows, num_non_zeros=info.entries, value=v, suffix=ij)


def sparse_to_dense(m: SparseMat) -> np.ndarray:
    """convert_sparse_matrix_to_dense (distribute_matrix.f90:151-182) -- host-side, for checks only."""
    A = np.zeros((m.size, m.size), order="F")
    for (i, j), x in zip(m.suffix, m.value):
        A[i - 1, j - 1] = x
        A[j - 1, i - 1] = x
    return A


# ------------------------------------------------------------------------------------------ event logger
@dataclass
class _Event:
    name: str
    num_repeated: int
    val: float


@dataclass
class EventLogger:
    """add_event (event_logger.f90:23-65): accumulate by name, NEW names are prepended; every call prints
    `[Event<t F16.6>] <name>,<val E24.16e3>` to stderr."""
    t_init: float = field(default_factory=time.perf_counter)
    events: list = field(default_factory=list)
    echo: bool = True

    def add_event(self, name: str, seconds: float, to_print: bool = True) -> None:
        if self.echo and to_print:
            t = time.perf_counter() - self.t_init
            sys.stderr.write(f"[Event{t:16.6f}] {name},{fortran_e(seconds, 24, 16, 3)}\n")
        for e in self.events:
            if e.name == name:
                e.num_repeated += 1
                e.val += seconds
                return
        self.events.insert(0, _Event(name, 1, float(seconds)))

    def find(self, name: str):
        for e in self.events:
            if e.name == name:
                return e
        return None


def log_json_text(setting: dict, events: list) -> str:
    """fson_value_print (fson.f90:454-553) for the two-object tree main.f90:58-60,185-190 builds:
    2-space indent, `"name": value`, integers I0, reals E24.16e3 (unquoted), strings quoted unescaped."""

    def scalar(v):
        if isinstance(v, bool):
            return "true" if v else "false"
        if isinstance(v, (int, np.integer)):
            return str(int(v))
        if isinstance(v, (float, np.floating)):
            return fortran_e(v, 24, 16, 3)
        return '"' + str(v) + '"'

    out = ["{"]
    out.append('  "setting": {')
    items = list(setting.items())
    for q, (k, v) in enumerate(items):
        out.append(f'    "{k}": {scalar(v)}' + ("," if q + 1 < len(items) else ""))
    out.append("  },")
    out.append('  "events": [')
    for q, e in enumerate(events):
        out.append("    {")
        out.append(f'      "name": {scalar(e.name)},')
        out.append(f'      "num_repeated": {scalar(e.num_repeated)},')
        out.append(f'      "val": {scalar(float(e.val))}')
        out.append("    }" + ("," if q + 1 < len(events) else ""))
    out.append("  ]")
    out.append("}")
    return "\n".join(out) + "\n"
