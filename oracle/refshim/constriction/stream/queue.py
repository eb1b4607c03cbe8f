"""Range coder of constriction 0.4.2 (`stream.queue`), Word=u32, State=u64,
PRECISION=24.  SURVEY.md Appendix C.2 (decoder) and C.4 (encoder).  If the compiled
oracle (oracle/_build/libccoracle.so) is present it is used for speed; the pure-Python
code below is the specification and the fallback."""
import ctypes
import os

import numpy as np

from .model import PRECISION

M64 = (1 << 64) - 1

_lib = None
_so = os.path.join(os.path.dirname(__file__), "..", "..", "..", "_build", "libccoracle.so")
if os.path.exists(_so) and not os.environ.get("CCSHIM_PURE_PYTHON"):
    try:
        _lib = ctypes.CDLL(os.path.abspath(_so))
        _lib.cco_rc_decode_block.restype = ctypes.c_int
        _lib.cco_rc_decode_block.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        _lib.cco_rc_dec_new.restype = ctypes.c_void_p
        _lib.cco_rc_dec_new.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        _lib.cco_rc_dec_free.argtypes = [ctypes.c_void_p]
    except (OSError, AttributeError):
        _lib = None


class RangeDecoder:
    def __init__(self, compressed: np.ndarray):
        self.words = np.ascontiguousarray(compressed, dtype=np.uint32)
        self.n_calls = 0
        if _lib is not None:
            self._h = _lib.cco_rc_dec_new(self.words.ctypes.data, self.words.size)
            return
        self._h = None
        self.pos = 0
        self.lower = 0
        self.range = M64
        self.point = (self._next() << 32) | self._next()

    def __del__(self):
        if getattr(self, "_h", None) is not None and _lib is not None:
            _lib.cco_rc_dec_free(self._h)
            self._h = None

    def _next(self) -> int:
        w = int(self.words[self.pos]) if self.pos < self.words.size else 0
        self.pos += 1
        return w

    def decode(self, model, mu: np.ndarray, scale: np.ndarray) -> np.ndarray:
        self.n_calls += 1
        mu = np.ascontiguousarray(mu, dtype=np.float32)
        scale = np.ascontiguousarray(scale, dtype=np.float32)
        out = np.empty(mu.size, dtype=np.int32)
        if self._h is not None:
            rc = _lib.cco_rc_decode_block(
                self._h, mu.ctypes.data, scale.ctypes.data, mu.size, out.ctypes.data)
            if rc != 0:
                raise ValueError("Tried to decode from compressed data that is invalid")
            return out
        for i in range(mu.size):
            m, b = float(mu[i]), float(scale[i])
            sc = self.range >> PRECISION
            q = ((self.point - self.lower) & M64) // sc
            if q >= (1 << PRECISION):
                raise ValueError("Tried to decode from compressed data that is invalid")
            lo, hi = model.lo, model.hi + 1  # invariant: left(lo) <= q < left(hi)
            while hi - lo > 1:
                mid = (lo + hi) // 2
                if model.left(mid, m, b) <= q:
                    lo = mid
                else:
                    hi = mid
            left = model.left(lo, m, b)
            prob = model.left(lo + 1, m, b) - left
            self.lower = (self.lower + sc * left) & M64
            self.range = sc * prob
            if self.range < (1 << 32):
                self.lower = (self.lower << 32) & M64
                self.range <<= 32
                self.point = ((self.point << 32) & M64) | self._next()
            out[i] = lo
        return out


class RangeEncoder:
    def __init__(self):
        self.lower = 0
        self.range = M64
        self.words = []
        self.n_symbols = 0

    def _carry(self):
        i = len(self.words) - 1
        while i >= 0:
            self.words[i] = (self.words[i] + 1) & 0xFFFFFFFF
            if self.words[i] != 0:
                break
            i -= 1

    def encode(self, symbols, model, mu, scale):
        symbols = np.asarray(symbols).reshape(-1)
        mu = np.asarray(mu, dtype=np.float32).reshape(-1)
        scale = np.asarray(scale, dtype=np.float32).reshape(-1)
        for s, m, b in zip(symbols.tolist(), mu.tolist(), scale.tolist()):
            s = int(s)
            if s < model.lo or s > model.hi:
                raise ValueError("symbol out of model support")
            left = model.left(s, m, b)
            prob = model.left(s + 1, m, b) - left
            sc = self.range >> PRECISION
            new = self.lower + sc * left
            if new > M64:
                new &= M64
                self._carry()
            self.lower = new
            self.range = sc * prob
            if self.range < (1 << 32):
                self.words.append(self.lower >> 32)
                self.lower = (self.lower << 32) & M64
                self.range <<= 32
            self.n_symbols += 1

    def get_compressed(self) -> np.ndarray:
        words = list(self.words)
        if self.n_symbols > 0:
            point = self.lower + (1 << 32) - 1
            if point > M64:
                point &= M64
                i = len(words) - 1
                while i >= 0:
                    words[i] = (words[i] + 1) & 0xFFFFFFFF
                    if words[i] != 0:
                        break
                    i -= 1
            words.append(point >> 32)
        return np.array(words, dtype=np.uint32)
