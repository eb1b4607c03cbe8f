"""Development tool: the kernels of one 1080p frame decode, for ncu (see profiles/):
   ncu --set full --import-source on -k regex:"k_entropy|k_tail_syn|k_ups_level_b" -s 22 -c 7 python tools/gpu_ncu_target.py
(kernel launches before the profiled decode: seed-stream decode 7, range encode 1, two warm decodes 14)."""
import os, sys
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import coolchic_b200  # noqa: E402,F401
from coolchic_b200 import _native, synth  # noqa: E402
from coolchic_b200._desc import desc_from_header  # noqa: E402
ctx = _native.get_context(0)
ss = synth.SeedStream(ctx)
data = synth.make_image_stream(ctx, ss, 1080, 1920, "rgb", 8, (0, 6), None, seed=0)
_, f, c, nnb, lb = synth.parse_single_image(data)
d = desc_from_header(c)
for it in range(3):
    outs, _ = ctx.decode_many([d], [nnb], [lb], finish=[(8, "rgb")])
    torch.cuda.synchronize()
    print(ctx.last_timing())
