"""Generate tests/golden/gop5_64x96.* : a synthetic 5-frame GOP (I, P, 3 hierarchical B,
YUV420 8-bit, sinc-8 warps, mop residue / motion Cool-chics) written by coolchic_b200.synth
with the oracle as range encoder, and DECODED BY THE UNMODIFIED REFERENCE (through
oracle/refshim).  Runs only in the authoring container.  TEST INFRASTRUCTURE ONLY.

    python oracle/gen_golden_gop.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, "/root/reference")

import torch  # noqa: E402

torch.set_num_threads(1)
import coolchic_b200  # noqa: E402,F401
import pipeline  # noqa: E402
from coolchic_b200 import synth  # noqa: E402

GOLD = os.path.join(HERE, "..", "tests", "golden")


def main():
    be = pipeline.OracleBackend()
    ss = synth.SeedStream(be)
    # (name, (h, w, frames, format, bitdepth, warp_filter_size)); 2 / 4 = torch grid_sample bilinear / bicubic
    configs = [("gop5_64x96_yuv420", (64, 96, 5, "yuv420", 8, 8)), ("gop3_40x56_rgb", (40, 56, 3, "rgb", 8, 8)),
               ("gop3_48x72_yuv420_bilinear", (48, 72, 3, "yuv420", 8, 2)),
               ("gop3_40x56_rgb_bicubic", (40, 56, 3, "rgb", 8, 4)),
               ("gop3_48x72_yuv420_sinc6", (48, 72, 3, "yuv420", 8, 6)),
               ("gop4_40x56_rgb_sinc12", (40, 56, 4, "rgb", 8, 12))]
    for name, (h, w, n, fmt, bd, fs) in configs:
        data = synth.make_video_stream(be, ss, h, w, n, fmt, bd, fs, seed=1)
        path = os.path.join(GOLD, name + ".cool")
        with open(path, "wb") as f:
            f.write(data)
        from coolchic.bitstream.decode import decode_video as ref_decode_video

        frames = ref_decode_video(path, None, verbosity=0)
        M = 2**bd - 1
        out = {}
        for k, fd in frames.items():
            if fmt == "yuv420":
                for c in "yuv":
                    out[f"{k}_{c}"] = torch.round(fd.data[c][0, 0] * M).to(torch.int32).numpy().astype(np.uint16)
            else:
                out[k] = torch.round(fd.data[0] * M).to(torch.int32).numpy().astype(np.uint16)
        np.savez_compressed(os.path.join(GOLD, name + "_frames.npz"), **out)
        # how close is the oracle?
        mine = pipeline.decode_video(data)
        worst, nbad, ntot = 0, 0, 0
        for k, fd in frames.items():
            _, _, d = mine[int(k)]
            for c in ("yuv" if fmt == "yuv420" else [None]):
                a = (np.round((d[c] if c else d) * M)).astype(np.int32)
                b = out[f"{k}_{c}" if c else k].astype(np.int32)
                worst = max(worst, int(np.abs(a - b).max()))
                nbad += int((a != b).sum())
                ntot += a.size
        print(name, "bytes", len(data), "oracle vs reference: max level diff", worst, "differing samples", nbad, "of", ntot)


if __name__ == "__main__":
    main()
