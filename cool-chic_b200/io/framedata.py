"""Decoded-frame container -- mirror of the reference's ``coolchic/io/framedata.py:16-54``.

``data`` is a float32 tensor ``[1, 3, H, W]`` (rgb / yuv444) or a dict
``{"y": [1,1,H,W], "u": [1,1,H/2,W/2], "v": ...}`` (yuv420, io/format/yuv.py:21-38) with
values on the k / (2^bitdepth - 1) grid in [0, 1].  Tensors may live on the GPU.
"""
from dataclasses import dataclass, field
from typing import Any, Tuple


@dataclass
class FrameData:
    bitdepth: int
    frame_data_type: str  # "rgb" | "yuv420" | "yuv444"
    data: Any

    img_size: Tuple[int, int] = field(init=False)
    n_pixels: int = field(init=False)

    def __post_init__(self):
        ref = self.data.get("y") if self.frame_data_type == "yuv420" else self.data
        self.img_size = tuple(ref.size()[-2:])
        self.n_pixels = self.img_size[0] * self.img_size[1]

    def to_string(self) -> str:
        s = "Frame data information:\n"
        s += "-----------------------\n"
        s += f"{'Resolution (H, W)':<26}: {self.img_size[0]}, {self.img_size[1]}\n"
        s += f"{'Bitdepth':<26}: {self.bitdepth}\n"
        s += f"{'Data type':<26}: {self.frame_data_type}"
        return s

    def to(self, device) -> "FrameData":
        if self.frame_data_type == "yuv420":
            data = {k: v.to(device) for k, v in self.data.items()}
        else:
            data = self.data.to(device)
        return FrameData(self.bitdepth, self.frame_data_type, data)
