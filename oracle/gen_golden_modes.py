"""Generate tests/golden/img_*.{cool,npz}: small synthetic intra streams exercising the optional
branches of the synthesis input / output (common randomness, bilinear / bicubic final resize),
written by coolchic_b200.synth with the oracle as range encoder and DECODED BY THE UNMODIFIED
REFERENCE (through oracle/refshim).  Runs only in the authoring container.  TEST INFRASTRUCTURE ONLY.

    python oracle/gen_golden_modes.py
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

# name -> (h, w, format, latent_resolution, final_upsampling_type, header overrides)
CONFIGS = {
    "img_48x72_rgb_common_randomness": (48, 72, "rgb", (0, 3), None, {"flag_common_randomness": 1}),
    "img_50x70_rgb_final_bicubic": (50, 70, "rgb", (1, 4), "bicubic", None),
    "img_44x60_yuv420_final_bilinear": (44, 60, "yuv420", (2, 5), "bilinear", None),
    # other ARM shapes (the all-int64 kernel on the device) and a 10-bit 4:4:4 frame
    "img_48x64_rgb_arm8_1hidden": (48, 64, "rgb", (0, 4), None, {"spatial_context_arm": 8, "n_hidden_layers_arm": 1}),
    "img_40x56_yuv444_10bit_arm24": (40, 56, "yuv444", (0, 3), None, {"spatial_context_arm": 24}),
}
BITDEPTH = {"img_40x56_yuv444_10bit_arm24": 10}


def main():
    from coolchic.bitstream.decode import decode_video as ref_decode_video

    be = pipeline.OracleBackend()
    ss = synth.SeedStream(be)
    for name, (h, w, fmt, lat, fin, ov) in CONFIGS.items():
        bd = BITDEPTH.get(name, 8)
        M = 2**bd - 1
        data = synth.make_image_stream(be, ss, h, w, fmt, bd, lat, None, seed=3, final_upsampling_type=fin, overrides=ov)
        path = os.path.join(GOLD, name + ".cool")
        with open(path, "wb") as f:
            f.write(data)
        fd = ref_decode_video(path, None, verbosity=0)["0"]
        out = {}
        if fmt == "yuv420":
            for c in "yuv":
                out[c] = torch.round(fd.data[c][0, 0] * M).to(torch.int32).numpy().astype(np.uint16)
        else:
            out["rgb"] = torch.round(fd.data[0] * M).to(torch.int32).numpy().astype(np.uint16)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
        _, _, d = pipeline.decode_video(data)[0]
        nbad = ntot = worst = 0
        for k, b in out.items():
            a = np.round((d[k] if fmt == "yuv420" else d) * M).astype(np.int32)
            worst = max(worst, int(np.abs(a - b.astype(np.int32)).max()))
            nbad += int((a != b).sum())
            ntot += a.size
        print(name, "bytes", len(data), "oracle vs reference: max level diff", worst, "differing samples", nbad, "of", ntot)


if __name__ == "__main__":
    main()
