"""Development tool: throughput of N concurrent Kodak-size streams on one GPU (decode_many)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import coolchic_b200
from coolchic_b200 import _native, synth
from coolchic_b200._desc import desc_from_header
ctx = _native.get_context(0)
ss = synth.SeedStream(ctx)
items = []
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 24):
    cc, h, lat = synth.make_coolchic(ctx, ss, (512, 768), (0, 6), (4, 6), seed=i)
    h2 = type(h)(); rest = h2.read_header(cc)
    items.append((desc_from_header(h2), rest[:h2.get_value("nn_n_bytes")], rest[h2.get_value("nn_n_bytes"):][:h2.get_value("n_bytes_latent")]))
for n in (1, 2, 4, 8, 16, 24, 48, 96, 148):
    sub = (items * ((n + len(items) - 1) // len(items)))[:n]
    for it in range(2):
        t = time.time(); outs, _ = ctx.decode_many([x[0] for x in sub], [x[1] for x in sub], [x[2] for x in sub]); torch.cuda.synchronize(); dt = time.time() - t
    tm = ctx.last_timing()
    print(f"{n:4d} streams: wall {dt*1e3:8.2f} ms  entropy {tm['entropy_ms']:8.2f} ms  synth {tm['synthesis_ms']:7.2f} ms  -> {n*512*768/dt/1e6:8.1f} Mpixel/s")
