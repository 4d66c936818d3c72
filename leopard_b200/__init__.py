"""leopard_b200 — B200-native (sm_100a) implementation of Leopard's multi-image prefill hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all arithmetic on the path runs in the
hand-written CUDA library `lib/libleopard_b200.so` behind the C ABI in `include/leopard_b200.h`.
See DESIGN.md for the scope and INTEGRATION.md for how the reference scripts bind to it.
"""
__version__ = "0.1.0"
