"""dissect_b200 — B200-native replacement for DISSECT's `--make-grm` hot path.

Only what the path needs lives here: `csrc/` (sm_100a CUDA kernels + the C ABI of
`include/dgrm.h`, built into `libdgrm.so`), `host/` (C++ mirror of the reference's
Kernel/Genotype/Matrix interface on top of the C ABI), and thin ctypes bindings used by the
tests, `bench.py` and the multi-GPU driver.  There is no CPU fallback: importing the engine
without the built CUDA library raises.
"""
from .engine import GrmEngine, DgrmError, lib_path  # noqa: F401
