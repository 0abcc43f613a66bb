"""spiral_b200 - B200 (sm_100a) implementation of Spiral's server-side query answering.

The product is libspiral_b200.so (hand-written CUDA behind a C-ABI, include/spiral_b200.h).
This package is the thin Python host side used by tests and bench.py: a ctypes binding
(`spiral_b200.lib`), a mirror of the reference's server entry points (`spiral_b200.server`) and the
face of the GPU client (`spiral_b200.client`).
There is no CPU fallback: importing works anywhere, every compute call needs a CUDA device.
"""
from .lib import load_library, SpiralParams, SB200Error  # noqa: F401
