"""consistentid_b200 - B200-native (sm_100a) implementation of the ConsistentID denoising hot path.

Host side: Python/PyTorch (device memory, streams, torch.distributed).  Arithmetic: libcidb200.so (hand-written CUDA:
tcgen05 GEMM / implicit-GEMM conv / attention fed by TMA, fused HBM-bound kernels) behind the C ABI in include/cidb200.h.
Importing the package requires the built library - there is no fallback path.
"""
from . import lib  # noqa: F401  (raises ImportError with build instructions if the .so is missing)

__all__ = ["lib"]
