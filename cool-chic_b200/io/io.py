"""Frame writers: PNG (8-bit rgb), PPM (rgb), planar YUV (420 / 444, appended per frame).

Mirror of the reference's ``coolchic/io/io.py:53-105`` (dispatch on the file extension, same
assertions), ``io/format/png.py:44-62``, ``io/format/ppm.py:160-203`` and
``io/format/yuv.py:124-165``.  Samples are ``round(x * (2^b - 1))``; the YUV writer uses
uint16 only when bitdepth == 10 (yuv.py:157, kept as is), PPM stores 2-byte samples MSB first.
"""
import os

import numpy as np
import torch

from .framedata import FrameData

POSSIBLE_EXT = [".yuv", ".png", ".ppm"]


def _levels(t: torch.Tensor, bitdepth: int) -> np.ndarray:
    return torch.round(t.detach().float().cpu() * (2**bitdepth - 1)).numpy()


def write_png(data: torch.Tensor, file_path: str) -> None:
    from PIL import Image

    arr = data.detach().float().cpu().numpy()
    assert arr.ndim == 4 and arr.shape[0] == 1 and arr.shape[1] == 3, (
        f"Data shape must be [1, 3, H, W], found {arr.shape}"
    )
    arr = np.round(np.clip(arr[0].transpose(1, 2, 0), 0.0, 1.0) * 255).astype(np.uint8)
    Image.fromarray(arr, mode="RGB").save(file_path)


def write_ppm(data: torch.Tensor, bitdepth: int, file_path: str, norm: bool = True) -> None:
    c, h, w = data.size()[-3:]
    t = data.detach().float().cpu().reshape(c, h, w)
    max_val = 2**bitdepth - 1
    vals = torch.round(t * max_val) if norm else t
    hwc = vals.permute(1, 2, 0).contiguous().numpy()
    body = hwc.astype(np.uint8 if max_val <= 255 else ">u2").tobytes()
    with open(file_path, "wb") as f_out:
        f_out.write(f"P6\n{w} {h}\n{max_val}\n".encode("ascii"))
        f_out.write(body)


def write_yuv(data, bitdepth: int, frame_data_type: str, file_path: str, norm: bool = True,
              append: bool = False) -> None:
    assert frame_data_type in ["yuv420", "yuv444"], (
        f"Found incorrect datatype in write_yuv() function: {frame_data_type}. "
        'Data type should be "yuv420" or "yuv444".'
    )
    if frame_data_type == "yuv420":
        raw = torch.cat([data[k].detach().float().cpu().flatten() for k in ("y", "u", "v")])
    else:
        raw = data.detach().float().cpu().flatten()
    if norm:
        raw = raw * (2**bitdepth - 1)
    dtype = np.uint16 if bitdepth == 10 else np.uint8
    out = torch.round(raw).numpy().astype(np.int64).astype(dtype)
    with open(file_path, "ab" if append else "wb") as f_out:
        out.tofile(f_out)


# ---- writers of samples already packed on the device (ccd_pack_frame): no arithmetic left on the host ----------
def write_packed_yuv(planar: np.ndarray, bitdepth: int, frame_data_type: str, file_path: str, append: bool) -> None:
    assert frame_data_type in ["yuv420", "yuv444"], (
        f"Found incorrect datatype in write_yuv() function: {frame_data_type}. "
        'Data type should be "yuv420" or "yuv444".'
    )
    # io/format/yuv.py:157: 2-byte samples only when bitdepth == 10, 1-byte otherwise (kept as is)
    dtype = np.uint16 if bitdepth == 10 else np.uint8
    out = planar if planar.dtype == dtype else planar.astype(np.int64).astype(dtype)
    with open(file_path, "ab" if append else "wb") as f_out:
        out.tofile(f_out)


def write_packed_ppm(hwc: np.ndarray, bitdepth: int, file_path: str) -> None:
    h, w = hwc.shape[:2]
    max_val = 2**bitdepth - 1
    body = hwc.astype(np.uint8 if max_val <= 255 else ">u2").tobytes()
    with open(file_path, "wb") as f_out:
        f_out.write(f"P6\n{w} {h}\n{max_val}\n".encode("ascii"))
        f_out.write(body)


def write_packed_png(hwc_u8: np.ndarray, file_path: str) -> None:
    from PIL import Image

    Image.fromarray(np.ascontiguousarray(hwc_u8), mode="RGB").save(file_path)


def save_frame_data_to_file(frame_data: FrameData, file_path: str, append: bool = False) -> None:
    ext = os.path.splitext(file_path)[1]
    assert ext in POSSIBLE_EXT, (
        f"The function save_frame_data_to_file() expects a file ending with {POSSIBLE_EXT}. Found {file_path}"
    )
    if ext == ".png":
        assert frame_data.frame_data_type == "rgb", (
            "The function save_frame_data_to_file() can only save a RGB data "
            f"into a PNG file. Found frame_data_type = {frame_data.frame_data_type}."
        )
        assert frame_data.bitdepth == 8, (
            "The function save_frame_data_to_file() can only write 8-bit data "
            f"into a PNG file. Found bitdepth = {frame_data.bitdepth}."
        )
        write_png(frame_data.data, file_path)
    elif ext == ".ppm":
        assert frame_data.frame_data_type == "rgb", (
            "The function save_frame_data_to_file() can only save a RGB data "
            f"into a PPM file. Found frame_data_type = {frame_data.frame_data_type}."
        )
        write_ppm(frame_data.data, frame_data.bitdepth, file_path, norm=True)
    else:
        assert frame_data.frame_data_type in ["yuv420", "yuv444"], (
            "The function save_frame_data_to_file() can only save a YUV data "
            f"into a YUV file. Found frame_data_type = {frame_data.frame_data_type}."
        )
        write_yuv(frame_data.data, frame_data.bitdepth, frame_data.frame_data_type, file_path,
                  norm=True, append=append)
