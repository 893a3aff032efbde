"""Host model of the bulge-chasing schedule of csrc/sb2st.cu (band -> tridiagonal, second half of pdsytrd,
reference src/solver_scalapack_all.f90:59).

The CUDA kernel runs one sweep per CTA and lets sweep s start task t as soon as sweep s-1 has COMPLETED task t+1
("lag 2"), which is only legal because every task also computes the NEXT reflector of its sweep and writes that
reflector's beta into the band before it publishes its completion (look-ahead): the one element of task t+2's
block that sweep s+1 needs early.  This file replays exactly that schedule on the CPU with numpy, wavefront by
wavefront, every task of a wavefront reading a SNAPSHOT of the band taken when the wavefront starts (so a task
can never profit from a write of a task it is not ordered after), and checks that
  * the result is tridiagonal with the spectrum of the band matrix,
  * the reflectors (V2, TAU2 in the kernel's layout) reproduce it: Q2^T B Q2 = T,
  * the deferred part of the D block (every column but the first is stored only AFTER the task's completion count
    has been published, so that the release does not wait for it) is never read before the next count covers it,
  * concurrent tasks write disjoint elements (the L block of task t is stored WITHOUT its (0, 0) element: that beta
    went to the band with the look-ahead of task t-1 and may already have been consumed and overwritten).
It is a test of the ALGORITHM (index ranges, dependency distance); the kernel itself is checked on the GPU by
tests/test_gpu_twostage.py.
"""
import numpy as np
import pytest


def num_tasks(n, b, s):
    return 0 if s > n - 3 else (n - 3 - s) // b + 1


def house(x):
    """LAPACK dlarfg convention: H x = beta e1, H = I - tau v v^T, v[0] = 1."""
    alpha, sq = x[0], float(np.dot(x[1:], x[1:]))
    v = np.zeros_like(x)
    v[0] = 1.0
    if sq == 0.0:
        return v, 0.0, alpha
    beta = -np.copysign(np.sqrt(alpha * alpha + sq), alpha)
    v[1:] = x[1:] / (alpha - beta)
    return v, (beta - alpha) / beta, beta


class Band:
    """Lower band storage AB[i - j, j] = A[i, j] with 2b rows (b+1.. hold the bulges), as in the kernel."""

    def __init__(self, A, b):
        n = A.shape[0]
        self.n, self.b = n, b
        self.AB = np.zeros((2 * b, n))
        for j in range(n):
            for i in range(j, min(n, j + b + 1)):
                self.AB[i - j, j] = A[i, j]

    def dense(self):
        n, b = self.n, self.b
        M = np.zeros((n, n))
        for j in range(n):
            for d in range(min(2 * b, n - j)):
                M[j + d, j] = self.AB[d, j]
                M[j, j + d] = self.AB[d, j]
        return M


def run_schedule(A, b, lag, defer_d=True):
    """Replay the kernel's schedule.  Returns (band object, V2, TAU2, max writers per element per wavefront)."""
    n = A.shape[0]
    band = Band(A, b)
    AB = band.AB
    V2 = np.zeros((n, n))
    TAU2 = np.zeros((n // b + 2, n))
    # per-sweep state carried between tasks in registers / shared memory: the reflector of the next task and the
    # L block (= right-applied B block of the previous task, column 0 already replaced by beta e1)
    state = {}
    nsweep = max(n - 2, 0)
    # wavefront k runs every task (s, t) with lag * s + t == k - 1 (t = -1 is the sweep's prologue: reflector 0)
    kmax = lag * (nsweep - 1) + num_tasks(n, b, 0) + 2 if nsweep else 0
    worst_overlap = 0
    acc = {}  # element -> list of (sweep, task, is_write)
    for k in range(kmax + 1):
        snap_arr = AB.copy()
        writes = {}
        cur = [None]

        class _Snap:
            """reads of the band by the task in cur[0], logged per element"""

            def __getitem__(self, key):
                d, j = key
                if isinstance(d, slice):
                    for dd in range(d.start, d.stop):
                        acc.setdefault((j + dd, j), []).append((*cur[0], False))
                else:
                    acc.setdefault((j + d, j), []).append((*cur[0], False))
                return snap_arr[key]

        snap = _Snap()

        def put(i, j, val, tag, visible_with=None):
            """visible_with: the task whose completion count first covers this store (deferred stores: the columns
            of D other than the first go to the band only after the task's own count has been published, so for the
            ordering check they belong to the NEXT task of the sweep)."""
            key = (i, j)
            if key in writes:
                raise AssertionError(f"wavefront {k}: element {key} written by {writes[key]} and {tag}")
            writes[key] = tag
            acc.setdefault(key, []).append((*(visible_with or tag), True))
            AB[i - j, j] = val

        for s in range(nsweep):
            t = k - 1 - lag * s
            nt = num_tasks(n, b, s)
            cur[0] = (s, t)
            if t == -1:
                # prologue: needs prog[s-1] >= 1, i.e. (s-1, 0) complete, which ran in wavefront k - lag <= k - 1
                r0 = s + 1
                nr = min(b, n - r0)
                x = snap[1:1 + nr, s].copy()
                v, tau, beta = house(x)
                V2[r0:r0 + nr, s] = v
                TAU2[0, s] = tau
                put(s + 1, s, beta, (s, -1))
                for i in range(1, nr):
                    put(s + 1 + i, s, 0.0, (s, -1))
                state[s] = {"v": v, "tau": tau, "L": None}
            elif 0 <= t < nt:
                st = state[s]
                v, tau = st["v"], st["tau"]
                r0 = s + 1 + t * b
                nr = min(b, n - r0)
                nr2 = max(0, min(b, n - (r0 + b)))
                # --- L block (registers): left-apply, column 0 is already beta e1; store
                if t > 0:
                    L = st["L"]  # nr x b, columns r0-b .. r0-1
                    y = tau * (v[:nr] @ L)
                    y[0] = 0.0
                    L = L - np.outer(v[:nr], y)
                    for jj in range(b):
                        for ii in range(nr):
                            # element (0, 0) = beta went to the band with the look-ahead of task t-1; by now a later
                            # sweep may already have consumed AND overwritten it: it must not be stored again
                            if ii or jj:
                                put(r0 + ii, r0 - b + jj, L[ii, jj], (s, t))
                # --- D and B blocks from the snapshot (needs (s-1, t+1) complete: wavefront k-lag+... <= k-1)
                D = np.zeros((nr, nr))
                for jj in range(nr):
                    for ii in range(jj, nr):
                        D[ii, jj] = D[jj, ii] = snap[ii - jj, r0 + jj]
                Bk = np.zeros((nr2, nr))
                for jj in range(nr):
                    for ii in range(nr2):
                        Bk[ii, jj] = snap[b + ii - jj, r0 + jj]
                p = tau * (D @ v[:nr])
                w = p - 0.5 * tau * float(p @ v[:nr]) * v[:nr]
                D = D - np.outer(v[:nr], w) - np.outer(w, v[:nr])
                for jj in range(nr):
                    for ii in range(jj, nr):
                        late = defer_d and jj > 0 and t + 1 < nt
                        put(r0 + ii, r0 + jj, D[ii, jj], (s, t), (s, t + 1) if late else None)
                u = tau * (Bk @ v[:nr])
                Bk = Bk - np.outer(u, v[:nr])
                if t + 1 >= nt:
                    for jj in range(nr):
                        for ii in range(nr2):
                            put(r0 + b + ii, r0 + jj, Bk[ii, jj], (s, t))
                    state.pop(s)
                else:
                    vn, taun, betan = house(Bk[:, 0].copy())
                    V2[r0 + b:r0 + b + nr2, s] = vn
                    TAU2[t + 1, s] = taun
                    Bk[:, 0] = 0.0
                    Bk[0, 0] = betan
                    put(r0 + b, r0, betan, (s, t))  # look-ahead: the one element sweep s+1 needs early
                    vfull = np.zeros(b)
                    vfull[:nr2] = vn
                    Lfull = np.zeros((nr2, b))
                    Lfull[:, :nr] = Bk
                    state[s] = {"v": vfull, "tau": taun, "L": Lfull}
        worst_overlap = max(worst_overlap, len(writes))
    # Every pair of accesses to one element with at least one write must be ordered by the kernel's happens-before
    # relation: same sweep in task order, or (s', t') before (s, t) for s' < s iff t' <= t + (s - s') * (lag - 1)
    # (sweep s waits for prog[s-1] >= t + lag, i.e. tasks 0 .. t+lag-1 of sweep s-1).  This covers EVERY legal
    # interleaving, not only the as-soon-as-possible one replayed above.
    def hb(a, b_):
        (s1, t1), (s2, t2) = a, b_
        if s1 == s2:
            return t1 < t2
        return s1 < s2 and t1 <= t2 + (s2 - s1) * (lag - 1)

    for key, lst in acc.items():
        for x in range(len(lst)):
            for y in range(x + 1, len(lst)):
                a, b_ = lst[x], lst[y]
                if not (a[2] or b_[2]) or a[:2] == b_[:2]:
                    continue
                assert hb(a[:2], b_[:2]) or hb(b_[:2], a[:2]), f"unordered conflicting accesses to {key}: {a} {b_}"
    return band, V2, TAU2


def q2_from_reflectors(V2, TAU2, n, b):
    Q = np.eye(n)
    for s in range(n - 2):
        for t in range(num_tasks(n, b, s)):
            r0 = s + 1 + t * b
            nr = min(b, n - r0)
            v, tau = V2[r0:r0 + nr, s], TAU2[t, s]
            Q[:, r0:r0 + nr] -= tau * np.outer(Q[:, r0:r0 + nr] @ v, v)
    return Q


def random_band(n, b, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n))
    A = A + A.T
    for i in range(n):
        for j in range(n):
            if abs(i - j) > b:
                A[i, j] = 0.0
    return A


@pytest.mark.parametrize("n,b", [(3, 2), (9, 2), (17, 4), (23, 4), (24, 4), (25, 4), (40, 6), (41, 8), (64, 8)])
def test_lag2_schedule_with_lookahead_is_a_valid_bulge_chase(n, b):
    A = random_band(n, b, 100 * n + b)
    band, V2, TAU2 = run_schedule(A, b, lag=2)
    T = band.dense()
    off = T - np.diag(np.diag(T)) - np.diag(np.diag(T, 1), 1) - np.diag(np.diag(T, -1), -1)
    scale = np.abs(A).max()
    assert np.abs(off).max() <= 1e-13 * scale, "result is not tridiagonal"
    assert np.max(np.abs(np.linalg.eigvalsh(T) - np.linalg.eigvalsh(A))) <= 1e-12 * scale * n
    Q = q2_from_reflectors(V2, TAU2, n, b)
    assert np.max(np.abs(Q.T @ Q - np.eye(n))) <= 1e-13 * n
    assert np.max(np.abs(Q.T @ A @ Q - T)) <= 1e-12 * scale * n


def test_lag1_schedule_is_not_valid():
    """The dependency distance is tight: with lag 1 concurrent tasks collide (the model raises) or the result is
    wrong."""
    n, b = 25, 4
    A = random_band(n, b, 7)
    try:
        band, _, _ = run_schedule(A, b, lag=1)
    except AssertionError:
        return
    T = band.dense()
    assert np.max(np.abs(np.linalg.eigvalsh(T) - np.linalg.eigvalsh(A))) > 1e-8
