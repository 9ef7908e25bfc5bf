"""polus_b200 -- B200-native drop-in for the data-parallel training step of bioinformatics-ua/polus.

Host code is Python (as the reference is); all device work is hand-written CUDA for sm_100a in
libpolus_b200.so, bound through the C ABI of include/polus_b200.h.  There is no CPU fallback.
"""
__version__ = "0.1.0"
