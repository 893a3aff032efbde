"""Host-side decisions of the GEMM engine (csrc/layout.h, used by csrc/gemm.cu), run on the CPU through the host-check
library: the split-K factor of the TMA-fed kernel (round-count model) and its L2-aware tile order.  The engine is the
B200 counterpart of the PBLAS products behind pdpotrf / pdsygst / pdsytrd / pdormtr / pdtrtrs (reference
src/generalized_to_standard.f90:24,37,103; src/solver_scalapack_all.f90:59,115)."""
import ctypes
import math
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HC = os.path.join(ROOT, "eigenkernel_b200", "libekb200_hostcheck.so")

SMS = 148


def _lib():
    lib = ctypes.CDLL(HC)
    lib.ekb200_host_gemm_autosplit.restype = ctypes.c_int
    lib.ekb200_host_gemm_autosplit.argtypes = [ctypes.c_longlong] * 3 + [ctypes.c_int]
    lib.ekb200_host_gemm_raster_tile.restype = None
    lib.ekb200_host_gemm_raster_tile.argtypes = [ctypes.c_int] * 4 + [ctypes.POINTER(ctypes.c_int)] * 2
    return lib


def _rounds(m, n, s, sms=SMS):
    bn = 128 if n > 64 else 64
    return math.ceil(math.ceil(m / 128) * math.ceil(n / bn) * s / sms)


@pytest.mark.parametrize("gm,gn", [(1, 1), (1, 7), (5, 1), (12, 12), (13, 3), (24, 5), (25, 64), (64, 64), (250, 1),
                                   (128, 257), (11, 40)])
def test_raster_order_is_a_bijection_and_blocks_the_resident_set(gm, gn):
    lib = _lib()
    tm, tn = ctypes.c_int(), ctypes.c_int()
    seen = np.zeros((gm, gn), dtype=np.int64)
    order = []
    for pid in range(gm * gn):
        lib.ekb200_host_gemm_raster_tile(pid, gm, gn, 12, ctypes.byref(tm), ctypes.byref(tn))
        assert 0 <= tm.value < gm and 0 <= tn.value < gn
        seen[tm.value, tn.value] += 1
        order.append((tm.value, tn.value))
    assert (seen == 1).all()                       # every tile exactly once: no product element skipped or doubled
    # what the order is for: any 148 consecutive CTAs (the resident set) stay inside the groups of 12 row tiles they
    # span (one more than ceil(148 / (12 gn)) of them), not all gm row tiles
    if gm * gn >= SMS:
        bound = 12 * (math.ceil(SMS / (12 * gn)) + 1)
        for start in range(0, gm * gn - SMS, max(1, (gm * gn) // 50)):
            rows = {t[0] for t in order[start:start + SMS]}
            assert len(rows) <= min(gm, bound)


def test_autosplit_never_splits_shallow_or_already_full_products():
    lib = _lib()
    # k below 1024 cannot be halved without leaving a CTA less than 512 of it
    for k in (64, 128, 512, 1023):
        assert lib.ekb200_host_gemm_autosplit(32000, 64, k, SMS) == 1
    # many rounds already: quantisation loss is small, the partial-sum pass is not worth it
    assert lib.ekb200_host_gemm_autosplit(16000, 16384, 512, SMS) == 1
    assert lib.ekb200_host_gemm_autosplit(8192, 8192, 8192, SMS) == 1
    assert lib.ekb200_host_gemm_autosplit(32768, 32768, 16384, SMS) == 1


@pytest.mark.parametrize("m", [32000, 24000, 18944, 16000, 12000, 8000, 4096, 2048])
def test_autosplit_fills_the_rounds_of_the_panel_product(m):
    """The m x 64 panel product of the dense-to-band reduction (W = A V, k = m): T = m / 128 tiles on 148 SMs."""
    lib = _lib()
    s = lib.ekb200_host_gemm_autosplit(m, 64, m, SMS)
    assert 1 <= s <= 8 and (s == 1 or m // s >= 512)
    tiles = math.ceil(m / 128)
    eff = tiles * s / (_rounds(m, 64, s) * SMS)     # fraction of the CTA slots of the rounds that do work
    eff1 = tiles / (_rounds(m, 64, 1) * SMS)
    assert eff >= eff1 - 1e-12
    if m >= 8000:
        assert eff >= 0.84, (m, s, eff)            # unsplit: 0.63 at m = 24000, 0.84 at 32000
    # the workspace it asks for stays bounded
    assert s * m * 64 * 8 <= 512e6


def test_autosplit_respects_the_workspace_cap_and_the_sm_count():
    lib = _lib()
    # 2 x 512 x 16384 doubles = 134 MB: allowed; the round count goes 4 -> 7 half-length rounds
    assert lib.ekb200_host_gemm_autosplit(512, 16384, 16000, SMS) == 2
    # a product whose partial sums would not fit 512 MB is never split
    assert lib.ekb200_host_gemm_autosplit(8192, 8192, 2048, 4096) == 1
    # with more SMs than tiles x 8 every extra split still helps until k / s < 512
    s = lib.ekb200_host_gemm_autosplit(256, 64, 4096, SMS)
    assert s == 8
