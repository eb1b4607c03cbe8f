"""Generate tests/golden/kodim14_* from the UNMODIFIED reference decode path.

Runs only in the authoring container (needs /root/reference).  The reference is imported
from where it lies, with the two stand-ins of oracle/refshim (fvcore, constriction).
TEST INFRASTRUCTURE ONLY.

    python oracle/gen_golden_kodim14.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, "/root/reference")

import torch  # noqa: E402

torch.set_num_threads(1)

import coolchic.bitstream.component.coolchic as ref_cc  # noqa: E402
import coolchic.bitstream.decode as ref_dec  # noqa: E402

GOLD = os.path.join(HERE, "..", "tests", "golden")
SRC = "/root/reference/samples/bitstreams/kodim14.cool"

captured = {"latents": [], "raw": []}

_orig_entropy = ref_cc.entropy_coding_latent_arm


def _rec_entropy(*a, **k):
    out = _orig_entropy(*a, **k)
    captured["latents"].append(out.to(torch.int8).numpy().copy().reshape(out.shape[-2:]))
    return out


ref_cc.entropy_coding_latent_arm = _rec_entropy

_orig_edc = ref_dec.encode_decode_coolchic


def _rec_edc(*a, **k):
    out, b = _orig_edc(*a, **k)
    captured["raw"].append(out.numpy().copy())
    return out, b


ref_dec.encode_decode_coolchic = _rec_edc

frames = ref_dec.decode_video(SRC, None, verbosity=0)
img = frames["0"].data  # [1,3,H,W] float in [0,1] on the k/255 grid
u8 = torch.round(img[0] * 255).to(torch.uint8).permute(1, 2, 0).contiguous().numpy()

lat_flat = np.concatenate([g.reshape(-1) for g in captured["latents"]])
raw = captured["raw"][0][0]  # [3,H,W]

print("n symbols", lat_flat.size, "sum", int(lat_flat.astype(np.int64).sum()),
      "sumabs", int(np.abs(lat_flat.astype(np.int64)).sum()))
print("latent sha256", hashlib.sha256(lat_flat.tobytes()).hexdigest())
print("image  sha256", hashlib.sha256(u8.tobytes()).hexdigest(), "sum", int(u8.astype(np.int64).sum()))
print("raw range", float(raw.min()), float(raw.max()))

with open(SRC, "rb") as f:
    data = f.read()
with open(os.path.join(GOLD, "kodim14.cool"), "wb") as f:
    f.write(data)
np.savez_compressed(
    os.path.join(GOLD, "kodim14_latents.npz"),
    latents=lat_flat,
    shapes=np.array([g.shape for g in captured["latents"]], dtype=np.int32),
)
np.savez_compressed(os.path.join(GOLD, "kodim14_image_u8.npz"), image=u8)
# raw synthesis output: every 8th row + first/last 8 rows, and first/last 8 columns
H, W = raw.shape[-2:]
rows = sorted(set(list(range(0, H, 8)) + list(range(8)) + list(range(H - 8, H))))
np.savez_compressed(
    os.path.join(GOLD, "kodim14_raw_rows.npz"),
    rows=np.array(rows, dtype=np.int32),
    data=raw[:, rows, :],
    left=raw[:, :, :8],
    right=raw[:, :, W - 8:],
)
