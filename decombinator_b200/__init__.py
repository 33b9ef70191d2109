"""decombinator_b200: the `decombine` hot path of innate2adaptive/decombinator on NVIDIA B200.

Python host code mirrors the reference's module layout for this path (``decombine``, ``io``,
``pipeline``); the work is done by hand-written sm_100a CUDA kernels behind the C ABI of
``include/dcb.h`` (``libdcb.so``, built in-tree by ``decombinator_b200.build``).
"""
__version__ = "4.3.0+b200.r1"
