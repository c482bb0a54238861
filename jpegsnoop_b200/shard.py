"""Multi-GPU plumbing of the batch path: images of a batch are independent, so the batch is split
contiguously across ranks (no data-path collective); the only exchange is ONE broadcast of the
shared table blob (jsgpu_tables sets) from rank 0 — NCCL over NVLink on GPUs, gloo in CPU tests."""
import ctypes as C
import numpy as np
from . import _lib as B


def shard_range(n, world, rank):
    """Contiguous shard [lo, hi) of n images for `rank` of `world` (sizes differ by at most 1)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def tables_to_bytes(tarr):
    return bytes(tarr)


def tables_from_bytes(raw):
    n = len(raw) // C.sizeof(B.jsgpu_tables)
    return (B.jsgpu_tables * n).from_buffer_copy(raw)


def broadcast_tables(tarr, src=0, device=None):
    """Broadcast rank `src`'s table sets to every rank with torch.distributed (backend-agnostic).
    Ranks other than src may pass None.  Returns a (jsgpu_tables * n) array on every rank."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    dev = device if device is not None else "cpu"
    n = torch.tensor([len(tarr) if rank == src else 0], dtype=torch.int64, device=dev)
    dist.broadcast(n, src=src)
    nbytes = int(n.item()) * C.sizeof(B.jsgpu_tables)
    if rank == src:
        blob = torch.frombuffer(bytearray(tables_to_bytes(tarr)), dtype=torch.uint8).to(dev)
    else:
        blob = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dist.broadcast(blob, src=src)
    return tables_from_bytes(blob.cpu().numpy().tobytes())
