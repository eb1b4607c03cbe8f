"""Stand-in for the third-party Rust wheel `constriction==0.4.2` (reference
requirements.txt:10), which is not installed in this image and has no source under
/root/reference.  Restates the published algorithm of
`constriction.stream.queue.Range{De,En}coder` (Word=u32, State=u64, PRECISION=24) and
`constriction.stream.model.QuantizedLaplace` (leaky quantiser over f64) — SURVEY.md
Appendix C — so that the *reference's own* decode path can run here
(call sites: coolchic/bitstream/component/rangecoder.py:30-34,62,82,93).

TEST INFRASTRUCTURE ONLY: used to generate tests/golden/ fixtures and as CPU baseline.
"""
from . import stream  # noqa: F401
