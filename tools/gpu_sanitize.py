"""Development tool: decode fixtures under compute-sanitizer (memcheck / racecheck) and check them.
usage: compute-sanitizer --tool racecheck python tools/gpu_sanitize.py [kodim14] [small]"""
import os, sys
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import coolchic_b200
from coolchic_b200.bitstream.decode import decode_video
G = os.path.join(ROOT, "tests", "golden")
ok = True
if "small" in sys.argv:
    for name in ("img_48x64_rgb_arm8_1hidden", "img_50x70_rgb_final_bicubic", "img_48x72_rgb_common_randomness"):
        fr = decode_video(os.path.join(G, name + ".cool"))["0"]
        ref = np.load(os.path.join(G, name + ".npz"))["rgb"]
        got = np.round(fr.data[0].numpy() * (2 ** fr.bitdepth - 1)).astype(np.uint16)
        nd = int((got != ref).sum())
        print(name, "differing samples", nd)
        ok &= nd <= 4
    fr = decode_video(os.path.join(G, "gop5_64x96_yuv420.cool"))
    ref = np.load(os.path.join(G, "gop5_64x96_yuv420_frames.npz"))
    nd = 0
    for k, f in fr.items():
        for p in "yuv":
            nd += int((np.round(f.data[p][0, 0].numpy() * 255).astype(np.uint16) != ref[f"{k}_{p}"]).sum())
    print("gop5_64x96_yuv420 differing samples", nd)
    ok &= nd <= 40
if "kodim14" in sys.argv:
    fr = decode_video(os.path.join(G, "kodim14.cool"))["0"]
    ref = np.load(os.path.join(G, "kodim14_image_u8.npz"))["image"]
    got = np.round(fr.data[0].numpy() * 255).astype(np.uint8).transpose(1, 2, 0)
    nd = int((got != ref).sum())
    print("kodim14 differing samples", nd)
    ok &= nd <= 32
torch.cuda.synchronize()
print("SANITIZE_RUN_OK" if ok else "SANITIZE_RUN_MISMATCH")
