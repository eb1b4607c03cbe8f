"""`QuantizedLaplace(min, max)`: leakily quantised Laplace family with per-symbol
(mu, scale) parameters (f32 widened to f64).  SURVEY.md Appendix C.1."""
import math

PRECISION = 24


class QuantizedLaplace:
    def __init__(self, min_symbol_inclusive: int, max_symbol_inclusive: int):
        self.lo = int(min_symbol_inclusive)
        self.hi = int(max_symbol_inclusive)
        # free weight = (2^24 - 1) - (max - min), held as f64
        self.free_weight = float(((1 << PRECISION) - 1) - (self.hi - self.lo))

    def left(self, s: int, mu: float, b: float) -> int:
        """Left-sided cumulative of symbol s (s may be hi+1 -> 2^24)."""
        if s <= self.lo:
            return 0
        if s > self.hi:
            return 1 << PRECISION
        x = s - 0.5
        if x <= mu:
            c = 0.5 * math.exp((x - mu) / b)
        else:
            c = 1.0 - 0.5 * math.exp((mu - x) / b)
        return int(self.free_weight * c) + (s - self.lo)
