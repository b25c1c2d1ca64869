"""versband_b200 — Blackwell-native (sm_100a) AccompBand inference hot path.

Hand-written CUDA kernels behind a C-ABI library (include/versband.h), driven by PyTorch
modules that mirror the reference's nn.Module surfaces. No fallback path exists.
"""
__version__ = "0.1.0"
