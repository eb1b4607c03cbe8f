"""cool-chic_b200 -- B200-native decoder for Cool-chic 5.0 bitstreams.

Drop-in for the decode path of Orange-OpenSource/Cool-Chic (``cc_decode.py`` ->
``coolchic.bitstream.decode.decode_video``): Python host code (this package) calling
hand-written sm_100a CUDA kernels through the C-ABI of ``csrc/libccdec.so``
(``include/ccdec.h``).  PyTorch tensors are only the I/O container.

The directory name contains a hyphen; import it through the ``coolchic_b200`` alias module
at the repository root.
"""
__version__ = "0.1.0"
