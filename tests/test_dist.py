"""world_size-2 gloo test (CPU) of the frame-sharding plumbing used by bench.py --gpus N."""
import os
import socket
import sys

import torch.multiprocessing as mp
from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    sys.path.insert(0, os.path.abspath(ROOT))
    import torch.distributed as dist

    import coolchic_b200  # noqa: F401
    from coolchic_b200.dist import broadcast_byte_strings, gather_sharded, shard_indices

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    items = [bytes([i]) * (10 + 7 * i) for i in range(5)] if rank == 0 else None
    got = broadcast_byte_strings(items, src=0)
    mine = shard_indices(len(got), rank, world)
    merged = gather_sharded({i: len(got[i]) for i in mine}, world)
    ret[rank] = (got == [bytes([i]) * (10 + 7 * i) for i in range(5)], mine, merged)
    dist.destroy_process_group()


def test_broadcast_and_shard_two_ranks():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][0] and ret[1][0]
    assert ret[0][1] == [0, 2, 4] and ret[1][1] == [1, 3]
    assert ret[0][2] == ret[1][2] == {i: 10 + 7 * i for i in range(5)}


def test_shard_indices_cover_everything():
    sys.path.insert(0, os.path.abspath(ROOT))
    import coolchic_b200  # noqa: F401
    from coolchic_b200.dist import shard_indices

    for n in (1, 7, 24, 32):
        for w in (1, 2, 4, 8):
            allidx = sorted(i for r in range(w) for i in shard_indices(n, r, w))
            assert allidx == list(range(n))
