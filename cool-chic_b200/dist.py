"""Multi-GPU plumbing: frames (and the Cool-chics inside them) are independent units
(SURVEY 8e), so the path shards by stream with NO data-path collective.  The only
communication is a broadcast of the bitstream bytes from the rank that read the file
(torch.distributed, NCCL on GPUs / gloo in the CPU tests)."""
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment: item i (coding order) -> rank i mod world."""
    return list(range(rank, n_items, world))


def broadcast_byte_strings(items: Optional[Sequence[bytes]], src: int = 0, device=None) -> List[bytes]:
    """Every rank returns the list of byte strings held by ``src``."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return list(items)
    dev = device if device is not None else torch.device("cpu")
    rank = dist.get_rank()
    meta = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == src:
        meta[0] = len(items)
    dist.broadcast(meta, src)
    n = int(meta.item())
    lens = torch.zeros(n, dtype=torch.int64, device=dev)
    if rank == src:
        lens.copy_(torch.tensor([len(b) for b in items], dtype=torch.int64))
    dist.broadcast(lens, src)
    total = int(lens.sum().item())
    buf = torch.empty(total, dtype=torch.uint8, device=dev)
    if rank == src:
        buf.copy_(torch.frombuffer(bytearray(b"".join(items)), dtype=torch.uint8))
    dist.broadcast(buf, src)
    raw = buf.cpu().numpy().tobytes()
    out, p = [], 0
    for ln in lens.tolist():
        out.append(raw[p:p + ln])
        p += ln
    return out


def gather_sharded(local: dict, world: int) -> dict:
    """Merge per-rank {index: object} dictionaries on every rank (all_gather_object)."""
    if not dist.is_initialized() or world == 1:
        return dict(local)
    parts = [None] * world
    dist.all_gather_object(parts, local)
    out = {}
    for p in parts:
        out.update(p)
    return out
