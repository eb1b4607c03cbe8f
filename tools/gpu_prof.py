"""Development tool: run the instrumented entropy kernel on kodim14 and print cycle counters."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import coolchic_b200
from coolchic_b200.bitstream.header import VideoHeader, FrameHeader, CoolChicHeader
from coolchic_b200._desc import desc_from_header
from coolchic_b200 import _native
data = open(os.path.join(ROOT, "tests/golden/kodim14.cool"), "rb").read()
v = VideoHeader(); rest = v.read_header(data); f = FrameHeader(); rest = f.read_header(rest); c = CoolChicHeader(); rest = c.read_header(rest)
d = desc_from_header(c)
nnb = rest[:c.get_value("nn_n_bytes")]; lb = rest[c.get_value("nn_n_bytes"):][:c.get_value("n_bytes_latent")]
ctx = _native.get_context(0)
nn = _native.decode_nn(d, nnb)
g = np.load(os.path.join(ROOT, "tests/golden/kodim14_latents.npz"))["latents"]
for it in range(3):
    lat = ctx.decode_latents(d, nn, lb); torch.cuda.synchronize()
    st = ctx.last_status()
    print(ctx.last_timing(), "ok" if np.array_equal(lat.cpu().numpy(), g) else "MISMATCH")
    print(" status", st[:4], "| coder wait %d total %d kcyc | producers(sum of 15 warps) wait %d arm %d win %d total %d kcyc" % tuple(st[4:10]), "ext", st[10:16])
