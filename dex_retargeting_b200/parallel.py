"""Multi-GPU plumbing: one process per GPU, frames sharded contiguously, no per-step collective.

The reference is a single-process, single-thread CPU program: there is nothing
to port.  Hand-frames (and streams) are independent, so a batch splits embarrassingly across ranks;
the only exchange the path has is making sure every rank solves with the SAME robot table, which is
one broadcast of 8 KB at init time (NCCL over NVLink when the process group is NCCL).
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np

from . import _native as N
from .table import table_bytes, table_from_bytes


def shard_range(num_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous shard [begin, end) of rank; sizes differ by at most one, earlier ranks get the extra."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world_size")
    base, extra = divmod(int(num_items), world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def broadcast_table(optimizer, src: int = 0, group=None, device=None):
    """Broadcast rank `src`'s compiled robot table to every rank and make the optimizer use it.

    With an NCCL group the table travels GPU->GPU (NVLink / NVSwitch) and the solver handle adopts the
    received device copy directly (`dexr_robot_create_from_device`).  With a CPU group (gloo; used by the
    CPU tests) only the bytes are exchanged and checked.  Returns the table bytes every rank now holds."""
    import torch
    import torch.distributed as dist

    backend = dist.get_backend(group)
    nbytes = C.sizeof(N.DexrTable)
    local = np.frombuffer(table_bytes(optimizer.build_table()), dtype=np.uint8).copy()
    if backend == "nccl":
        dev = torch.device("cuda", optimizer.device_index if device is None else int(device))
        buf = torch.from_numpy(local).to(dev)
        dist.broadcast(buf, src=src, group=group)
        torch.cuda.current_stream(dev).synchronize()
        data = bytes(buf.cpu().numpy())
        optimizer.adopt_device_table(table_from_bytes(data), buf.data_ptr(), dev.index)
        return data
    buf = torch.from_numpy(local)
    dist.broadcast(buf, src=src, group=group)
    data = bytes(buf.numpy())
    if len(data) != nbytes:
        raise RuntimeError("robot table broadcast returned the wrong size")
    return data


def all_gather_qpos(qpos_shard, num_total: int, group=None):
    """Optional: gather per-rank results [B_r, n] into one [B, n] tensor on every rank (uneven shards)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = [e - b for b, e in (shard_range(num_total, r, world) for r in range(world))]
    width = max(sizes)  # all_gather needs equal shapes: pad the short shards by at most one row
    padded = torch.zeros((width, qpos_shard.shape[1]), dtype=qpos_shard.dtype, device=qpos_shard.device)
    padded[: qpos_shard.shape[0]] = qpos_shard
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)
