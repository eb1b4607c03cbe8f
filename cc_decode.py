"""Cool-chic decoder command line -- same flags as the reference's ``cc_decode.py:9-20``:

    python cc_decode.py -i <bitstream.cool> -o <decoded.png|.ppm|.yuv> [--verbosity N]

Decoding runs on a B200 (sm_100a) through ``cool-chic_b200/csrc/libccdec.so``.
"""
import argparse

import coolchic_b200  # noqa: F401  (import alias of the cool-chic_b200/ package)
from coolchic_b200.bitstream.decode import decode_video

if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--input", "-i", type=str, help="Bitstream path.")
    parser.add_argument("--output", "-o", type=str, help="Decoded file path.")
    parser.add_argument("--verbosity", type=int, help="Verbosity level.", default=0)
    parser.add_argument("--device", type=int, help="CUDA device ordinal.", default=0)
    args = parser.parse_args()

    decode_video(args.input, decoded_path=args.output, verbosity=args.verbosity, device=args.device)
