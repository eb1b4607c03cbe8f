"""Dummy symbols so that `from fvcore.nn import FlopCountAnalysis, flop_count_table`
(reference coolchic/component/core/coolchic.py:16) succeeds.  Never called on the decode path."""


class FlopCountAnalysis:  # pragma: no cover - encoder only
    def __init__(self, *a, **k):
        raise RuntimeError("fvcore stand-in: FlopCountAnalysis is not available")


def flop_count_table(*a, **k):  # pragma: no cover - encoder only
    raise RuntimeError("fvcore stand-in: flop_count_table is not available")
