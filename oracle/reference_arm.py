"""CPU arm of the benchmark (bench.py --impl reference, bench.py cpu_baseline).  BENCHMARK / TEST INFRASTRUCTURE ONLY:
nothing in the product imports this file.

Two CPU implementations of the decode path are timed here, on the host cores of the box that runs the benchmark:

* kind "reference": the UNMODIFIED reference package (Orange-OpenSource/Cool-Chic 5.0.1), staged by
  ``oracle/make_ref.sh`` into ``oracle/_ref/coolchic`` (git-ignored, travels to the GPU box), called through its own
  public API ``coolchic.bitstream.decode.decode_video(path, None)`` (``cc_decode.py:9-20``, ``bitstream/decode.py:26-91``).
  Its two third-party imports that this image lacks come from ``oracle/refshim``: ``fvcore`` (dummy) and
  ``constriction`` (the range coder forwards to the compiled oracle, so the stand-in is not the bottleneck).
* kind "port": the C restatement ``oracle/ccoracle.c`` (entropy stage on one core -- a stream is one serial chain --,
  float tail on OpenMP threads).

Inputs are fabricated WITHOUT a GPU by ``coolchic_b200.synth`` on top of ``pipeline.OracleBackend`` (same bytes as
the device-side writer: both range encoders are byte-exact with constriction)."""
import contextlib
import io
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, "_ref")
UNIT = "Mpixel/s"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "coolchic", "bitstream", "decode.py"))


def _import_oracle():
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import ccoracle

    ccoracle.build()
    return ccoracle


def fabricate(workload: dict) -> bytes:
    """The benchmark's synthetic stream, written on the CPU (oracle range encoder).  ``workload`` has the keys of
    bench.WORKLOADS entries: kind 'image' (h, w, fmt, lat_res, hyp_res, seed) or 'video' (h, w, fmt, n_frames)."""
    _import_oracle()
    import coolchic_b200  # noqa: F401  (host logic only: headers, exp-Golomb, stream assembly)
    from coolchic_b200 import synth
    from pipeline import OracleBackend

    be = OracleBackend()
    ss = synth.SeedStream(be)
    if workload.get("kind", "image") == "video":
        return synth.make_video_stream(be, ss, workload["h"], workload["w"], workload["n_frames"], workload["fmt"], 8,
                                       warp_filter_size=workload.get("warp_filter_size", 8), seed=workload.get("seed", 0))
    return synth.make_image_stream(be, ss, workload["h"], workload["w"], workload["fmt"], 8, workload["lat_res"],
                                   workload["hyp_res"], seed=workload.get("seed", 0))


def _load_reference():
    """Import the staged reference (once).  Returns its decode_video."""
    shim = os.path.join(HERE, "refshim")
    for p in (REF_DIR, shim):
        if p not in sys.path:
            sys.path.insert(0, p)
    _import_oracle()  # builds oracle/_build/libccoracle.so, which the constriction stand-in forwards to
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # SyntaxWarning: invalid escape sequences in the reference's docstrings
        from coolchic.bitstream.decode import decode_video
    import coolchic

    if not os.path.abspath(coolchic.__file__).startswith(os.path.abspath(REF_DIR)):
        raise RuntimeError(f"imported coolchic from {coolchic.__file__}, not from oracle/_ref")
    return decode_video


def time_reference(data: bytes, n_pixels: int, threads=None) -> dict:
    """One decode of ``data`` by the reference's own decode_video (decoded_path=None), all frames, serially (the
    reference has no frame parallelism).  ``threads``: torch intra-op threads (None = the reference's default, i.e.
    whatever torch picks on this box)."""
    import torch

    decode_video = _load_reference()
    if threads is not None:
        torch.set_num_threads(int(threads))
    used = torch.get_num_threads()
    with tempfile.NamedTemporaryFile(suffix=".cool", delete=False) as f:
        f.write(data)
        path = f.name
    try:
        sink = io.StringIO()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(sink):  # the reference prints one line per frame (decode.py:79-81)
            frames = decode_video(path, None, verbosity=0)
        dt = time.perf_counter() - t0
    finally:
        os.unlink(path)
    return {"value": n_pixels / dt / 1e6, "unit": UNIT, "cores": used, "kind": "reference", "seconds": dt,
            "frames": len(frames),
            "sample": f"the whole workload ({len(frames)} frame(s)) once through the unmodified reference "
                      f"decode_video (oracle/_ref, torch {torch.__version__} CPU, {used} intra-op thread(s); constriction "
                      f"stand-in backed by the compiled oracle); {dt:.2f} s"}


def time_port(data: bytes, n_pixels: int, threads: int) -> dict:
    """The same stream through the C port (oracle), all frames."""
    ccoracle = _import_oracle()
    import pipeline

    cores = ccoracle.set_threads(threads)
    t0 = time.perf_counter()
    frames = pipeline.decode_video(data)
    dt = time.perf_counter() - t0
    return {"value": n_pixels / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port", "seconds": dt,
            "frames": len(frames),
            "sample": f"the whole workload ({len(frames)} frame(s)) once through the C port of the path (oracle/ccoracle.c; "
                      f"entropy stage of a stream on 1 core -- one serial chain --, float tail on {cores} OpenMP thread(s)); "
                      f"{dt:.2f} s"}
