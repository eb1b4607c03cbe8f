"""Bit-packed headers of a Cool-chic 5.0 bitstream: video / frame / cool-chic.

Host-side mirror of the reference's ``coolchic/bitstream/header/{header,element}.py``
(``AbstractHeader.read_header`` header.py:72-88, ``to_bytes`` :90-105, ``VideoHeader``
:129-167, ``FrameHeader`` :171-239, ``CoolChicHeader`` :243-377; element syntax
element.py:46-85,302-373).  Same class names, ``read_header`` / ``get_value`` /
``set_value`` / ``to_bytes`` / ``pretty_string`` behaviour and ``ValueError`` on
out-of-range values -- but implemented as a table-driven integer bit codec instead of
strings of '0'/'1'.

Syntax rules: fields are MSB-first, concatenated, zero-padded at the END to a byte
(header.py:90-105).  Signed fields are sign-magnitude (element.py:46-85).  Every header
carries its own byte length in the 16-bit ``n_bytes_header`` field, which closes the
fixed-length part (header.py:33-34).
"""
from __future__ import annotations

import functools
import math
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

FRAME_TYPES = ("I", "P", "B")                       # utils/codingstructure.py:20
FRAME_DATA_TYPES = ("rgb", "yuv420", "yuv444", "flow")  # io/types.py:12
BITDEPTHS = (8, 9, 10, 11, 12, 13, 14, 15, 16)      # io/types.py:13
FINAL_UPSAMPLING = ("nearest", "bilinear", "bicubic")  # header.py:271-276
SYN_MODES = ("linear", "residual")                  # core/synthesis.py:171
SYN_NONLIN = ("none", "relu")                       # core/synthesis.py:166-169
NN_MODULES = ("arm", "ifce", "upsampling", "synthesis")  # core/types.py:99-102 (bitstream order)
NN_KINDS = ("weight", "bias")                       # core/types.py:18-19

# Index -> log2(q_step) tables, nnquant/quantstep.py:19-44.
Q_SHIFT_TABLE: Dict[Tuple[str, str], Tuple[int, ...]] = {
    ("arm", "weight"): tuple(range(-8, 1)),
    ("arm", "bias"): tuple(range(-16, 1)),
    ("ifce", "weight"): tuple(range(-8, 1)),
    ("ifce", "bias"): tuple(range(-16, 1)),
    ("upsampling", "weight"): tuple(range(-12, 1)),
    ("upsampling", "bias"): (0,),
    ("synthesis", "weight"): tuple(range(-12, 1)),
    ("synthesis", "bias"): tuple(range(-24, 1)),
}
# Index -> exp-Golomb order, nnquant/expgolomb.py:20-37.
EXP_GOL_TABLE = tuple(range(0, 13))


class BitReader:
    """MSB-first bit reader over a bytes object."""

    # a header is at most 65 535 bytes (n_bytes_header is a 16-bit field): never turn the whole remaining
    # bitstream (tens of MB for a GOP) into one Python integer
    MAX_BYTES = 65536 + 16

    def __init__(self, data):
        # headers are tens of bytes: start with a small window (shifting a 64 KB integer for every field is what made
        # parsing a batch of streams slow) and widen it on demand
        self._data = data
        self._limit = min(len(data), self.MAX_BYTES)
        self._load(min(self._limit, 256))
        self.pos = 0

    def _load(self, n_bytes: int) -> None:
        self._v = int.from_bytes(self._data[:n_bytes], "big")
        self._n = 8 * n_bytes

    def read(self, n_bits: int, signed: bool = False) -> int:
        if n_bits == 0:
            return 0
        if self.pos + n_bits > self._n:
            if self.pos + n_bits > 8 * self._limit:
                raise ValueError(
                    f"Header truncated: need {n_bits} bits at position {self.pos}, have {8 * self._limit}."
                )
            self._load(min(self._limit, max(2 * (self._n // 8), (self.pos + n_bits + 7) // 8)))
        shift = self._n - self.pos - n_bits
        raw = (self._v >> shift) & ((1 << n_bits) - 1)
        self.pos += n_bits
        if signed:
            mag = raw & ((1 << (n_bits - 1)) - 1)
            return -mag if raw >> (n_bits - 1) else mag
        return raw


class BitWriter:
    def __init__(self):
        self._v = 0
        self._n = 0

    def write(self, value: int, n_bits: int, signed: bool = False, name: str = "") -> None:
        # range rules of element.py:18-43 (note: sign-magnitude cannot hold -(2^(n-1)))
        if signed:
            lo, hi = -(2 ** (n_bits - 1)), 2 ** (n_bits - 1) - 1
        else:
            lo, hi = 0, 2**n_bits - 1
        if value > hi or value < lo or (signed and abs(value) > hi):
            raise ValueError(
                f"Trying to convert value {name}={value} to bytes using  {n_bits} bits with "
                f"signed={signed}. Value should be in [{lo}, {hi}]."
            )
        if signed:
            raw = (int(value < 0) << (n_bits - 1)) | abs(value)
        else:
            raw = value
        self._v = (self._v << n_bits) | raw
        self._n += n_bits

    def to_bytes(self) -> bytes:
        pad = (8 - self._n % 8) % 8
        return ((self._v << pad)).to_bytes((self._n + pad) // 8, "big")

    @property
    def n_bits(self) -> int:
        return self._n


@dataclass
class DescriptorNN:
    """(weight, bias) pair -- mirror of core/types.py:13-19."""

    weight: Any = None
    bias: Any = None

    def get_value(self, kind: str) -> Any:
        if kind not in NN_KINDS:
            raise ValueError(f"Can not get value for weight_or_bias={kind}. Available names are {list(NN_KINDS)}")
        return getattr(self, kind)

    def pretty_string(self) -> str:
        return f"weight={self.weight:<10}; bias={self.bias:<10}"


@dataclass
class DescriptorCoolChic:
    """Per-module descriptors in bitstream order -- mirror of core/types.py:92-102."""

    arm: DescriptorNN = field(default_factory=DescriptorNN)
    ifce: DescriptorNN = field(default_factory=DescriptorNN)
    upsampling: DescriptorNN = field(default_factory=DescriptorNN)
    synthesis: DescriptorNN = field(default_factory=DescriptorNN)

    def get_value(self, module: str, kind: Optional[str] = None) -> Any:
        d = getattr(self, module)
        return d if kind is None else d.get_value(kind)

    def set_value(self, value: Any, module: str, weight_or_bias: str) -> None:
        setattr(getattr(self, module), weight_or_bias, value)

    def pretty_string(self) -> str:
        return "".join(f"{m} {getattr(self, m).pretty_string()} " for m in NN_MODULES)


# field kinds: ("u", name, bits) | ("idx", name, bits, table) | ("list", name, bits, count_key|int, signed)
_Field = tuple


class _Header:
    """Common machinery: ordered fields, values dict, read/write."""

    _FIXED: Sequence[_Field] = ()

    def __init__(self):
        self._values: Dict[str, Any] = {}
        self._order: List[str] = []

    # ---- to be overridden -------------------------------------------------------------
    def _variable_fields(self) -> Sequence[_Field]:
        return ()

    # ---- generic ----------------------------------------------------------------------
    def _all_fields(self) -> List[_Field]:
        return list(self._FIXED) + [("u", "n_bytes_header", 16)] + list(self._variable_fields())

    def get_value(self, key: str) -> Optional[Any]:
        return self._values.get(key)

    def set_value(self, key: str, val: Any) -> None:
        names = [f[1] for f in self._all_fields()]
        if key not in names:
            raise ValueError(f"Can not set value {val}. Key {key} can not be found in the header.")
        self._values[key] = val

    def _read_field(self, br: BitReader, f: _Field) -> Any:
        kind = f[0]
        if kind == "u":
            return br.read(f[2])
        if kind == "idx":
            i = br.read(f[2])
            table = f[3]
            if i >= len(table):
                raise ValueError(f"Try to read list of length {len(table)} at index {i}.Variable name is {f[1]}")
            return table[i]
        if kind == "list":
            n = f[3] if isinstance(f[3], int) else self._values[f[3]]
            mult = f[5] if len(f) > 5 else 1
            return [br.read(f[2], signed=f[4]) for _ in range(n * mult)]
        if kind == "desc":
            out = DescriptorCoolChic()
            for m in NN_MODULES:
                for k in NN_KINDS:
                    i = br.read(f[2])
                    table = f[3](m, k)
                    if i >= len(table):
                        raise ValueError(
                            f"Try to read list of length {len(table)} at index {i}.Variable name is {f[1]}"
                        )
                    out.set_value(table[i], m, k)
            return out
        if kind == "syn":
            out_ft = br.read(7)
            k_size = br.read(4)
            mode = SYN_MODES[br.read(1)]
            nl = SYN_NONLIN[br.read(1)]
            return f"{out_ft}-{k_size}-{mode}-{nl}"
        raise AssertionError(kind)

    def _write_field(self, bw: BitWriter, f: _Field) -> None:
        kind, name = f[0], f[1]
        v = self._values.get(name)
        if kind == "u":
            bw.write(int(v or 0), f[2], name=name)
        elif kind == "idx":
            bw.write(list(f[3]).index(v), f[2], name=name)
        elif kind == "list":
            n = f[3] if isinstance(f[3], int) else self._values[f[3]]
            mult = f[5] if len(f) > 5 else 1
            v = list(v or [])
            if len(v) != n * mult:
                raise ValueError(f"{name}: expected {n * mult} values, found {len(v)}")
            for x in v:
                bw.write(int(x), f[2], signed=f[4], name=name)
        elif kind == "desc":
            for m in NN_MODULES:
                for k in NN_KINDS:
                    table = list(f[3](m, k))
                    bw.write(table.index(v.get_value(m, k)), f[2], name=f"{name}-{m}-{k}")
        elif kind == "syn":
            out_ft, k_size, mode, nl = str(v).split("-")
            bw.write(int(out_ft), 7, name="out_ft")
            bw.write(int(k_size), 4, name="k_size")
            bw.write(SYN_MODES.index(mode), 1, name="mode")
            bw.write(SYN_NONLIN.index(nl), 1, name="non_linearity")
        else:
            raise AssertionError(kind)

    def read_header(self, raw_data: bytes) -> bytes:
        """Parse the header at the start of ``raw_data``; return the bytes that follow it."""
        br = BitReader(raw_data)
        self._values = {}
        for f in list(self._FIXED) + [("u", "n_bytes_header", 16)]:
            self._values[f[1]] = self._read_field(br, f)
        for f in self._variable_fields():
            self._values[f[1]] = self._read_field(br, f)
        n = self._values["n_bytes_header"]
        if n * 8 < br.pos or n > len(raw_data):
            raise ValueError(f"Inconsistent n_bytes_header={n} (parsed {br.pos} bits, have {len(raw_data)} bytes).")
        return raw_data[n:]

    def to_bytes(self) -> bytes:
        fields = self._all_fields()
        n_bits = 0
        probe = BitWriter()
        self._values.setdefault("n_bytes_header", 0)
        for f in fields:
            self._write_field(probe, f)
        n_bits = probe.n_bits
        self._values["n_bytes_header"] = math.ceil(n_bits / 8)
        bw = BitWriter()
        for f in fields:
            self._write_field(bw, f)
        return bw.to_bytes()

    def pretty_string(self) -> str:
        msg = ""
        for f in self._all_fields():
            v = self._values.get(f[1])
            s = v.pretty_string() if isinstance(v, DescriptorCoolChic) else f"{v}"
            msg += f"{f[1]:<30}{s:<40}\n"
        return msg


class VideoHeader(_Header):
    """header.py:129-167.  n_frames:12 n_intras:12 n_p_frames:12 | intra_pos p_pos (12 each)."""

    _FIXED = (("u", "n_frames", 12), ("u", "n_intras", 12), ("u", "n_p_frames", 12))

    def _variable_fields(self):
        return (
            ("list", "intra_pos", 12, "n_intras", False),
            ("list", "p_pos", 12, "n_p_frames", False),
        )

    def set_header(self, n_frames: int, intra_pos: List[int], p_pos: List[int]) -> None:
        self._values.update(n_frames=n_frames, n_intras=len(intra_pos), n_p_frames=len(p_pos),
                            intra_pos=list(intra_pos), p_pos=list(p_pos))

    def get_coding_structure(self):
        from ..utils.codingstructure import CodingStructure

        return CodingStructure(
            n_frames=self.get_value("n_frames"),
            intra_pos=self.get_value("intra_pos"),
            p_pos=self.get_value("p_pos"),
        )


class FrameHeader(_Header):
    """header.py:171-239."""

    _FIXED = (
        ("u", "display_index", 12),
        ("idx", "frame_type", 2, FRAME_TYPES),
        ("idx", "frame_data_type", 2, FRAME_DATA_TYPES),
        ("idx", "bitdepth", 4, BITDEPTHS),
    )

    def _n_refs(self) -> int:
        return {"I": 0, "P": 1, "B": 2}[self._values.get("frame_type", "I")]

    def _variable_fields(self):
        n = self._n_refs()
        out = [
            ("list", "index_references", 12, n, False),
            ("list", "global_flow", 14, 2 * n, True),  # (x, y) per reference
        ]
        if n:
            out.append(("u", "warp_filter_size", 4))
        return out


@functools.lru_cache(maxsize=None)
def _q_table(module: str, kind: str):
    return tuple(2.0**s for s in Q_SHIFT_TABLE[(module, kind)])


def _eg_table(module: str, kind: str):
    return EXP_GOL_TABLE


class CoolChicHeader(_Header):
    """header.py:243-377."""

    _FIXED = (
        ("u", "linear_stabiliser_synth", 1),
        ("u", "n_layer_synthesis", 3),
        ("u", "ups_k_size", 4),
        ("u", "ups_preconcat_k_size", 4),
        ("u", "output_feature_ifce", 5),
        ("u", "spatial_context_arm", 6),
        ("u", "linear_stabiliser_arm", 1),
        ("u", "n_hidden_layers_arm", 3),
        ("list", "img_size", 14, 2, False),
        ("list", "latent_resolution", 4, 2, False),
        ("u", "n_latent_grids", 5),
        ("u", "flag_hyperlatent", 1),
        ("u", "flag_common_randomness", 1),
        ("idx", "final_upsampling_type", 2, FINAL_UPSAMPLING),
        ("desc", "nn_q_step", 5, _q_table),
        ("desc", "nn_expgol_cnt", 4, _eg_table),
        ("u", "nn_n_bytes", 14),
        ("u", "nn_n_bit_pad", 3),
        ("u", "n_bytes_latent", 28),
    )

    def _variable_fields(self):
        out = []
        if self._values.get("output_feature_ifce", 0) > 0:
            out.append(("list", "ifce_resolution", 4, 2, False))
        if self._values.get("flag_hyperlatent", 0):
            out.append(("list", "hyperlatent_resolution", 4, 2, False))
        for i in range(self._values.get("n_layer_synthesis", 0)):
            out.append(("syn", f"syn_layer_{i}"))
        return out

    def get_coolchic_parameter(self) -> "CoolChicParameter":
        """header.py:354-377.  A corrupt header raises ValueError here (the reference fails later with whatever
        exception the first impossible value triggers: ZeroDivisionError, IndexError ...)."""
        g = self.get_value
        img = tuple(g("img_size"))
        lat = tuple(g("latent_resolution"))
        if len(img) != 2 or img[0] < 1 or img[1] < 1:
            raise ValueError(f"Corrupt Cool-chic header: img_size = {img}")
        if g("n_layer_synthesis") < 1:
            raise ValueError("Corrupt Cool-chic header: n_layer_synthesis = 0")
        if len(lat) != 2 or lat[0] > lat[1]:
            raise ValueError(f"Corrupt Cool-chic header: latent_resolution = {lat}")
        for key in ("ifce_resolution", "hyperlatent_resolution"):
            r = g(key)
            if r is not None and (len(r) != 2 or r[0] > r[1]):
                raise ValueError(f"Corrupt Cool-chic header: {key} = {tuple(r)}")
        return CoolChicParameter(
            layers_synthesis=[g(f"syn_layer_{i}") for i in range(g("n_layer_synthesis"))],
            linear_stabiliser_synth=bool(g("linear_stabiliser_synth")),
            ups_k_size=g("ups_k_size"),
            ups_preconcat_k_size=g("ups_preconcat_k_size"),
            ifce_resolution=tuple(g("ifce_resolution")) if g("ifce_resolution") is not None else None,
            output_feature_ifce=g("output_feature_ifce"),
            spatial_context_arm=g("spatial_context_arm"),
            linear_stabiliser_arm=bool(g("linear_stabiliser_arm")),
            n_hidden_layers_arm=g("n_hidden_layers_arm"),
            latent_resolution=tuple(g("latent_resolution")),
            hyperlatent_resolution=(
                tuple(g("hyperlatent_resolution")) if g("hyperlatent_resolution") is not None else None
            ),
            flag_common_randomness=bool(g("flag_common_randomness")),
            img_size=tuple(g("img_size")),
            final_upsampling_type=g("final_upsampling_type"),
        )


@dataclass
class CoolChicParameter:
    """Architecture of one Cool-chic as derived from its header -- the decode-relevant part
    of the reference's ``CoolChicEncoderParameter`` (core/coolchic.py:51-225)."""

    layers_synthesis: List[str]
    linear_stabiliser_synth: bool
    ups_k_size: int
    ups_preconcat_k_size: int
    ifce_resolution: Optional[Tuple[int, int]]
    output_feature_ifce: int
    spatial_context_arm: int
    linear_stabiliser_arm: bool
    n_hidden_layers_arm: int
    latent_resolution: Tuple[int, int]
    hyperlatent_resolution: Optional[Tuple[int, int]]
    flag_common_randomness: bool
    img_size: Tuple[int, int]
    final_upsampling_type: str

    # derived (core/coolchic.py:149-225)
    size_per_latent: List[Tuple[int, int, int, int]] = field(init=False, default_factory=list)
    flag_is_hyperlatent: List[bool] = field(init=False, default_factory=list)
    input_features_ifce: List[int] = field(init=False, default_factory=list)
    n_latent_grids: int = field(init=False, default=0)
    flag_ifce: bool = field(init=False, default=False)
    flag_hyperlatent: bool = field(init=False, default=False)
    total_context_arm: int = field(init=False, default=0)
    input_feature_synthesis: int = field(init=False, default=0)

    def __post_init__(self):
        # post_init_latent, core/coolchic.py:160-190
        self.flag_hyperlatent = self.hyperlatent_resolution is not None
        lo, hi = self.latent_resolution
        if self.flag_hyperlatent:
            all_res = tuple(self.latent_resolution) + tuple(self.hyperlatent_resolution)
            min_ds, max_ds = min(all_res), max(all_res)
        else:
            min_ds, max_ds = lo, hi
        for i in range(min_ds, max_ds + 1):
            h_grid, w_grid = [int(math.ceil(x / (2**i))) for x in self.img_size]
            if lo <= i <= hi:
                self.size_per_latent.append((1, 1, h_grid, w_grid))
                self.flag_is_hyperlatent.append(False)
            if self.flag_hyperlatent:
                if self.hyperlatent_resolution[0] <= i <= self.hyperlatent_resolution[1]:
                    self.size_per_latent.append((1, 1, h_grid, w_grid))
                    self.flag_is_hyperlatent.append(True)
        self.n_latent_grids = len(self.size_per_latent)
        # post_init_arm :192-193
        self.total_context_arm = self.spatial_context_arm + self.output_feature_ifce
        # post_init_synthesis :202-206
        self.input_feature_synthesis = hi - lo + 1
        if self.flag_common_randomness:
            self.input_feature_synthesis *= 2
        # post_init_ifce :208-225
        self.flag_ifce = self.ifce_resolution is not None
        for i, size in enumerate(self.size_per_latent):
            ratio = int(math.ceil(math.log2(self.img_size[0] / size[-2])))
            if not self.flag_ifce:
                self.input_features_ifce.append(0)
            elif self.ifce_resolution[0] <= ratio <= self.ifce_resolution[1]:
                self.input_features_ifce.append(max(self.n_latent_grids - 1 - i, 1))
            else:
                self.input_features_ifce.append(0)

    def parsed_synthesis_layers(self) -> List[Tuple[int, int, bool, bool]]:
        """[(out_ft, k, residual, relu)] -- Synthesis._parse_layer_syntax synthesis.py:243-270."""
        out = []
        for lay in self.layers_synthesis:
            o, k, mode, nl = lay.split("-")
            if mode not in SYN_MODES:
                raise ValueError(f"Unknown mode. Found {mode}. Should be in {list(SYN_MODES)}")
            if nl not in SYN_NONLIN:
                raise ValueError(f"Unknown non linearity. Found {nl}. Should be in {list(SYN_NONLIN)}")
            out.append((int(o), int(k), mode == "residual", nl == "relu"))
        return out
