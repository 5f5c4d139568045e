"""ctsm_b200 — B200-native CTSM biogeophysics hot path (host-side mirror).

The compute lives in ctsm_b200/lib/libctsm_b200.so (hand-written CUDA for
sm_100a behind the C ABI of include/ctsm_b200.h).  This package only marshals
pointers; it has no CPU fallback and raises abi.LibraryMissing when the CUDA
library has not been built.
"""
__version__ = "0.1.0"
