"""lc2is_b200 - B200-native (sm_100a) implementation of the LC2IS segmentation-head hot path.

Host-side mirror of the reference's call surface (``model/text_patch.py``, ``model/decoder.py``,
``model/loss.py``, ``metrics.py``) over hand-written CUDA kernels behind a C ABI
(``include/lc2is_b200.h``).  No CPU fallback: importing the compute modules without the built
shared library raises.
"""
__version__ = "0.1.0"
