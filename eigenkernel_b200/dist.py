"""One process per B200: attach a Context to the ranks of a torch.distributed job.

B200 counterpart of `setup_distribution` / `layout_procs` (reference src/processes.f90:17-65) and of the
column bookkeeping of `get_local_cols` (src/distribute_matrix.f90:81-89).  torch.distributed is plumbing only:
it carries the 128-byte NCCL id from rank 0 to the others (any backend: gloo on CPU, nccl on GPUs) and the
host-side gathers of small results; the data-path exchanges (all-gathers of the sharded pdsygst, panel
exchanges of the sharded dense-to-band reduction) are NCCL calls issued inside libekb200.so on its own stream.
"""
from __future__ import annotations

import ctypes

import numpy as np

SLAB_GRAN = 128  # eigenvector slabs start on multiples of 128 columns (csrc/layout.h)


def slab_bounds(ncols: int, nranks: int, gran: int = SLAB_GRAN) -> list[int]:
    """Python mirror of ekb::slab_bounds (csrc/layout.h): rank r owns columns [b[r], b[r+1])."""
    chunk = max(1, -(-ncols // nranks))
    chunk = -(-chunk // gran) * gran
    return [min(r * chunk, ncols) for r in range(nranks + 1)]


def local_slab(ncols: int, nranks: int, rank: int, gran: int = SLAB_GRAN) -> tuple[int, int]:
    b = slab_bounds(ncols, nranks, gran)
    return b[rank], b[rank + 1] - b[rank]


def _bcast_bytes(buf: bytearray, src: int, group=None) -> bytes:
    """Broadcast a small byte string with whatever backend the default group has."""
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=src, group=group)
    return bytes(t.cpu().tolist())


def exchange_unique_id(make_id, group=None) -> bytes:
    """Rank 0 calls make_id() -> 128 bytes; every rank returns the same 128 bytes."""
    import torch.distributed as dist

    rank = dist.get_rank(group)
    raw = bytearray(make_id()) if rank == 0 else bytearray(128)
    if len(raw) != 128:
        raise ValueError("NCCL unique id must be 128 bytes")
    return _bcast_bytes(raw, 0, group)


def attach(ctx, group=None) -> tuple[int, int]:
    """Give `ctx` (eigenkernel_b200.device.Context) its NCCL communicator over the ranks of `group`.

    Returns (rank, world).  With world == 1 this is a no-op."""
    import torch.distributed as dist

    if not dist.is_initialized():
        return 0, 1
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return 0, 1

    def make_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        info = ctx.lib.ekb200_comm_unique_id(buf)
        if info != 0:
            raise RuntimeError(f"ekb200_comm_unique_id: info = {info}")
        return buf.raw

    uid = exchange_unique_id(make_id, group)
    ctx.call("ekb200_comm_init", world, rank, ctypes.c_char_p(uid))
    return rank, world


def gather_columns(local: np.ndarray, ncols: int, group=None, dst: int = 0):
    """Assemble the n x ncols matrix from the per-rank column slabs (host arrays) on rank `dst`.

    This is the host-side counterpart of the reference's eigenvector collection for printing
    (src/matrix_io.f90:187-226); it is never on the solve path."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    b = slab_bounds(ncols, world)
    assert local.shape[1] == b[rank + 1] - b[rank], (local.shape, b, rank)
    parts = [None] * world if rank == dst else None
    dist.gather_object(np.ascontiguousarray(local.T), parts, dst=dst, group=group)
    if rank != dst:
        return None
    n = local.shape[0]
    out = np.zeros((n, ncols), order="F")
    for r in range(world):
        if b[r + 1] > b[r]:
            out[:, b[r]:b[r + 1]] = parts[r].T
    return out


def max_over_ranks(x: float, group=None) -> float:
    """Multi-GPU timings are the max over ranks (the same rule the reference's timers follow implicitly:
    every rank waits at the next collective)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(x)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
