"""Thin object layer over the C-ABI: a context (one GPU) and device-resident column-major matrices.

Mirrors the roles of `ek_process_t` / `setup_distribution` (reference src/processes.f90:6-36: the "grid"
is one B200) and `setup_distributed_matrix` (src/distribute_matrix.f90:92-148: allocation + descriptor).
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_char_p, c_double, c_int, c_void_p

import numpy as np

from . import _lib
from ._lib import Ekb200Error


class DevMatrix:
    """Column-major FP64 matrix in HBM owned by a Context. ld is padded to a multiple of 8 elements."""

    def __init__(self, ctx: "Context", m: int, n: int, ld: int | None = None):
        self.ctx, self.m, self.n = ctx, int(m), int(n)
        self.ld = int(ld) if ld is not None else max(8, (self.m + 7) // 8 * 8)
        self.ptr = ctx.alloc(self.ld * max(self.n, 1) * 8)

    def addr(self, i: int = 0, j: int = 0) -> int:
        return self.ptr + 8 * (i + j * self.ld)

    def upload(self, a: np.ndarray) -> "DevMatrix":
        a = np.asfortranarray(a, dtype=np.float64)
        assert a.shape == (self.m, self.n)
        self.ctx.call("ekb200_h2d_matrix", self.ptr, self.ld, a.ctypes.data, a.strides[1] // 8 if a.ndim == 2 and self.n > 1 else max(self.m, 1), self.m, self.n)
        return self

    def download(self) -> np.ndarray:
        out = np.empty((self.m, self.n), order="F")
        self.ctx.call("ekb200_d2h_matrix", out.ctypes.data, max(self.m, 1), self.ptr, self.ld, self.m, self.n)
        return out

    def free(self):
        if self.ptr:
            self.ctx.free(self.ptr)
            self.ptr = 0


class Context:
    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = c_void_p()
        info = self.lib.ekb200_create(byref(h), int(device))
        if info != 0:
            raise Ekb200Error("ekb200_create", info, self.lib.ekb200_strerror(info).decode() +
                              " -- a CUDA device is required; there is no CPU fallback")
        self.h = h
        self.device = device
        # EKB200_OPTIONS="key=value,...": library tuning options for experiments (ekb200_set_option), e.g. running the
        # whole test suite with an alternative kernel selected
        import os
        for kv in filter(None, os.environ.get("EKB200_OPTIONS", "").split(",")):
            k, v = kv.split("=")
            self.set_option(k.strip(), int(v))

    # -- plumbing
    def call(self, name: str, *args) -> int:
        info = getattr(self.lib, name)(self.h, *args)
        if info < 0 or info >= 1000000:
            raise Ekb200Error(name, info, self.lib.ekb200_strerror(info).decode() + "; " +
                              self.lib.ekb200_last_error(self.h).decode())
        return info

    def alloc(self, nbytes: int) -> int:
        p = c_void_p()
        self.call("ekb200_dev_alloc", int(nbytes), byref(p))
        return p.value

    def free(self, ptr: int):
        self.call("ekb200_dev_free", c_void_p(ptr))

    def matrix(self, m: int, n: int) -> DevMatrix:
        return DevMatrix(self, m, n)

    def from_numpy(self, a: np.ndarray) -> DevMatrix:
        a = np.asarray(a, dtype=np.float64)
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        return DevMatrix(self, a.shape[0], a.shape[1]).upload(a)

    def sync(self):
        self.call("ekb200_sync")

    def set_option(self, key: str, value: int):
        self.call("ekb200_set_option", key.encode(), int(value))

    # -- timing table (event_logger.f90:23-65)
    def events(self):
        out = []
        for i in range(self.lib.ekb200_num_events(self.h)):
            name, sec, rep = c_char_p(), c_double(), c_int()
            self.lib.ekb200_get_event(self.h, i, byref(name), byref(sec), byref(rep))
            out.append((name.value.decode(), sec.value, rep.value))
        return out

    def clear_events(self):
        self.lib.ekb200_clear_events(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.ekb200_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- stage wrappers on DevMatrix
    def dgemm(self, ta: str, tb: str, alpha: float, A: DevMatrix, B: DevMatrix, beta: float, C: DevMatrix):
        m, n = C.m, C.n
        k = A.m if ta.upper() == "T" else A.n
        return self.call("ekb200_dgemm", ta.encode(), tb.encode(), m, n, k, float(alpha), A.ptr, A.ld, B.ptr, B.ld,
                         float(beta), C.ptr, C.ld)

    def fp64_peak(self):
        a, b = c_double(), c_double()
        self.call("ekb200_measure_fp64_peak", byref(a), byref(b))
        return {"dmma_tflops": a.value, "dfma_tflops": b.value}
