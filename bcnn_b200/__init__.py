"""bcnn_b200 -- a B200-native (sm_100a) implementation of bcnn's CNN layer hot path.

The product is `libbcnn_b200.so` (hand-written CUDA kernels in csrc/, C99 host runtime in
src/ that mirrors bcnn's net / node / tensor / layer interface; headers in /include).
This Python package is plumbing: the in-tree build (`build`), a ctypes binding of the C
API (`capi`) and the BASELINE workloads expressed through that API (`configs`).
There is no CPU fallback: importing is harmless, but creating a net without the built
library or without a CUDA device raises.
"""
from . import capi, configs  # noqa: F401

__version__ = "0.1.0"
