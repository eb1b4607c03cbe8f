"""The C-ABI surface (no compute without a GPU): the library loads, exports every symbol that
include/ccdec.h declares, agrees on the descriptor layout, and fails loudly without a device."""
import ctypes
import os
import re

import pytest
from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "ccdec.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ccd_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_everything_declared():
    import coolchic_b200  # noqa: F401
    from coolchic_b200 import _native

    assert os.path.exists(_native.LIB_PATH), "libccdec.so not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ccdec.h but not exported"
    assert sorted(_native.EXPORTS) == names


def test_descriptor_layout_and_counts(kodim14, oracle):
    from coolchic_b200 import _native
    from coolchic_b200._desc import CcdCoolChicDesc

    lib = _native.load_library()
    assert lib.ccd_sizeof_desc() == ctypes.sizeof(CcdCoolChicDesc) == oracle.lib().cco_sizeof_desc()
    d = kodim14["desc"]
    assert lib.ccd_nn_count(ctypes.byref(d)) == 1881
    n, offs = _native.latent_layout(d)
    n_o, offs_o = oracle.latent_layout(d)
    assert n == n_o == 526272 and offs == offs_o.tolist()
    # host exp-Golomb decode of the NN payload == oracle
    import numpy as np

    assert np.array_equal(_native.decode_nn(d, kodim14["nn_bytes"]), oracle.decode_nn(d, kodim14["nn_bytes"]))
    with pytest.raises(_native.CcdError) as e:
        _native.decode_nn(d, kodim14["nn_bytes"][:100])
    assert e.value.code == -2  # CCD_ERR_NN_TRUNCATED
    bad = CcdCoolChicDesc.from_buffer_copy(d)
    bad.n_grids = 0
    assert lib.ccd_nn_count(ctypes.byref(bad)) == -1  # CCD_ERR_ARG


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run (there is no CPU path)."""
    import torch

    from coolchic_b200 import _native

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _native.load_library()
    h = ctypes.c_void_p()
    rc = lib.ccd_create(0, ctypes.byref(h))
    assert rc == -7 and b"no CPU fallback" in lib.ccd_last_error(None)
    with pytest.raises(RuntimeError):
        _native.Context(0)
    from coolchic_b200.bitstream.decode import decode_video

    with pytest.raises(RuntimeError):
        decode_video(os.path.join(ROOT, "tests", "golden", "kodim14.cool"))


def test_product_does_not_import_oracle():
    """No product source may import / include / dlopen anything under oracle/ (comments may
    mention it)."""
    pkg = os.path.join(ROOT, "cool-chic_b200")
    pat = re.compile(r"(import\s+ccoracle|from\s+ccoracle|#include\s*[<\"][^>\"]*oracle|libccoracle|sys\.path[^\n]*oracle)")
    files = [os.path.join(ROOT, "cc_decode.py"), os.path.join(ROOT, "coolchic_b200.py")]
    for dirpath, _, names in os.walk(pkg):
        files += [os.path.join(dirpath, fn) for fn in names if fn.endswith((".py", ".cu", ".h", ".cuh", ".inc"))]
    for fn in files:
        assert not pat.search(open(fn).read()), fn
