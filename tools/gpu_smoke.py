"""First-light check on a B200: decode kodim14.cool through the C-ABI and compare with the
oracle / golden fixtures.  (Development tool; the real checks live in tests/.)"""
import os, sys, time, ctypes
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import coolchic_b200
from coolchic_b200.bitstream.header import VideoHeader, FrameHeader, CoolChicHeader
from coolchic_b200._desc import desc_from_header
from coolchic_b200 import _native
import ccoracle

data = open(os.path.join(ROOT, "tests/golden/kodim14.cool"), "rb").read()
v = VideoHeader(); rest = v.read_header(data)
f = FrameHeader(); rest = f.read_header(rest)
c = CoolChicHeader(); rest = c.read_header(rest)
d = desc_from_header(c)
nnb = rest[:c.get_value("nn_n_bytes")]; lb = rest[c.get_value("nn_n_bytes"):][:c.get_value("n_bytes_latent")]
ctx = _native.get_context(0)
nn = _native.decode_nn(d, nnb)
nn_o = ccoracle.decode_nn(d, nnb)
print("nn ints equal oracle:", np.array_equal(nn, nn_o))
g = np.load(os.path.join(ROOT, "tests/golden/kodim14_latents.npz"))["latents"]
for it in range(3):
    t = time.time(); lat = ctx.decode_latents(d, nn, lb); torch.cuda.synchronize(); dt = time.time() - t
    print("entropy wall %.2f ms" % (dt * 1e3), ctx.last_timing())
latc = lat.cpu().numpy()
print("latents equal golden:", np.array_equal(latc, g), "mismatch", int((latc != g).sum()))
if not np.array_equal(latc, g):
    idx = np.nonzero(latc != g)[0][:10]; print(idx, latc[idx], g[idx])
raw_o = ccoracle.synthesize(d, nn, g)
for it in range(3):
    t = time.time(); out, lat2 = ctx.decode_coolchic(d, nnb, lb, want_latents=True); torch.cuda.synchronize(); dt = time.time() - t
    print("full decode wall %.2f ms" % (dt * 1e3), ctx.last_timing())
raw = out[0].cpu().numpy()
print("raw bit-exact vs oracle:", np.array_equal(raw, raw_o), "max abs diff", float(np.abs(raw - raw_o).max()))
img = ctx.finish_frame(out, 8, "rgb")[0].cpu().numpy()
img_o = ccoracle.finish_frame(raw_o, 8, "rgb")
print("frame bit-exact vs oracle:", np.array_equal(img, img_o))
gi = np.load(os.path.join(ROOT, "tests/golden/kodim14_image_u8.npz"))["image"]
u8 = np.round(img * 255).astype(np.uint8).transpose(1, 2, 0)
print("uint8 diffs vs reference golden:", int((u8 != gi).sum()))
# device encoder: re-encode the latents -> shipped bytes
lat_e, payload, slow = ctx.encode_latents(d, nn, latents=torch.from_numpy(g).cuda())
print("re-encode equals shipped payload:", payload == lb, len(payload), len(lb), "slow", slow)
lat_s, payload_s, slow_s = ctx.encode_latents(d, nn, seed=1234)
lat_so, payload_so = ccoracle.sample_latents(d, nn, 1234)
print("sampled latents equal oracle:", np.array_equal(lat_s.cpu().numpy(), lat_so), "payload equal:", payload_s == payload_so, "slow", slow_s)
