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


def broadcast_byte_strings(items: Optional[Sequence[bytes]], src: int = 0, device=None,
                           want: Optional[Sequence[int]] = None) -> List[Optional[bytes]]:
    """Every rank returns the list of byte strings held by ``src``.  ``want`` (indices) limits what a rank turns back into
    Python byte strings -- the other entries of the returned list are None: the whole buffer still travels (one broadcast),
    but a rank that decodes 1 / N of a batch does not cut N / N of it into pieces."""
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
    lens_l = lens.tolist()
    total = sum(lens_l)
    buf = torch.empty(total, dtype=torch.uint8, device=dev)
    if rank == src:
        staged = bytearray(total)  # (one copy of every item, no intermediate joined bytes object)
        p = 0
        for b in items:
            staged[p:p + len(b)] = b
            p += len(b)
        buf.copy_(torch.frombuffer(staged, dtype=torch.uint8))
    dist.broadcast(buf, src)
    wanted = range(n) if want is None else want
    if rank == src:  # (the source already holds the strings)
        out: List[Optional[bytes]] = [None] * n
        for i in wanted:
            out[i] = items[i]
        return out
    offs = [0] * (n + 1)
    for i, ln in enumerate(lens_l):
        offs[i + 1] = offs[i] + ln
    out = [None] * n
    if want is not None and 2 * sum(lens_l[i] for i in wanted) <= total:
        # a small share of the buffer: one gather where the buffer lives, one copy of that share to the host
        idx = list(wanted)
        if idx:
            share = torch.cat([buf[offs[i]:offs[i + 1]] for i in idx]).cpu().numpy()
            p = 0
            for i in idx:
                out[i] = share[p:p + lens_l[i]].tobytes()
                p += lens_l[i]
        return out
    raw = memoryview(buf.cpu().numpy())
    for i in wanted:
        out[i] = bytes(raw[offs[i]:offs[i + 1]])
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
