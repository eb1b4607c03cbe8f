"""BASELINE.json configs[4] check: a 1080p YUV420 GOP decoded with its Cool-chics sharded over the ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/gpu_gop_sharded.py [n_frames]

Rank 0 fabricates the stream and first decodes it alone (reference result), then every rank decodes it through
decode_video_bytes under the NCCL process group; all ranks must reproduce rank 0's frames bit for bit."""
import contextlib
import hashlib
import io
import json
import os
import sys
import time

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import coolchic_b200  # noqa: E402,F401
from coolchic_b200 import _native, synth  # noqa: E402
from coolchic_b200.bitstream.decode import decode_video_bytes  # noqa: E402
from coolchic_b200.dist import broadcast_byte_strings  # noqa: E402


def digest(frames):
    h = hashlib.sha256()
    for k in sorted(frames, key=int):
        d = frames[k].data
        for t in ([d[c] for c in "yuv"] if isinstance(d, dict) else [d]):
            h.update(t.contiguous().cpu().numpy().tobytes())
    return h.hexdigest()


def install_timers(acc):
    """wall-clock split of decode_video_bytes: pass 1 (all Cool-chics) vs per-frame reconstruction"""
    from coolchic_b200.bitstream import decode as dec

    orig_all, orig_rec = dec._decode_all_coolchics, dec._reconstruct

    def timed_all(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter(); r = orig_all(*a, **k); torch.cuda.synchronize()
        acc["pass1_wall_s"] = time.perf_counter() - t
        return r

    def timed_rec(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter(); r = orig_rec(*a, **k); torch.cuda.synchronize()
        acc["pass2_wall_s"] = acc.get("pass2_wall_s", 0.0) + time.perf_counter() - t
        return r

    dec._decode_all_coolchics, dec._reconstruct = timed_all, timed_rec


def main():
    n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    ctx = _native.get_context(local)
    H, W = 1080, 1920
    data, want, t_single = None, None, None
    if rank == 0:
        ss = synth.SeedStream(ctx)
        data = synth.make_video_stream(ctx, ss, H, W, n_frames, "yuv420", 8, 8, seed=5)
        with contextlib.redirect_stdout(io.StringIO()):
            decode_video_bytes(data, device=local, output_device="cuda")  # warm-up (allocations, table build)
            torch.cuda.synchronize()
            acc = {}
            install_timers(acc)
            t0 = time.perf_counter()
            frames = decode_video_bytes(data, device=local, output_device="cuda")
            torch.cuda.synchronize()
            t_single = time.perf_counter() - t0
            sys.stderr.write("split " + json.dumps(acc) + "\n")
        want = digest(frames)
    import torch.distributed as dist

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        data = broadcast_byte_strings([data] if rank == 0 else None, src=0, device=torch.device("cuda", local))[0]
    with contextlib.redirect_stdout(io.StringIO()):
        decode_video_bytes(data, device=local, output_device="cuda")
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        frames = decode_video_bytes(data, device=local, output_device="cuda")
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t_sharded = time.perf_counter() - t0
    got = digest(frames)
    if world > 1:
        allg = [None] * world
        dist.all_gather_object(allg, got)
    else:
        allg = [got]
    if rank == 0:
        tm = ctx.last_timing()
        print(json.dumps({"pass1_entropy_ms": tm["entropy_ms"], "pass1_synthesis_ms": tm["synthesis_ms"], "pass1_upload_ms": tm["upload_ms"]}))
        print(json.dumps({"frames": n_frames, "bytes": len(data), "world": world, "all_ranks_equal_single_process": all(g == want for g in allg),
                          "single_gpu_s": t_single, "sharded_s": t_sharded,
                          "single_gpu_mpixel_s": n_frames * H * W / t_single / 1e6, "sharded_mpixel_s": n_frames * H * W / t_sharded / 1e6}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
