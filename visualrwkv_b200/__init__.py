"""visualrwkv_b200 — B200-native (sm_100a) implementation of the VisualRWKV data-parallel hot path.

Host side mirrors the reference interface (VisualRWKV-v7/v7.00/src/model.py); all arithmetic on
the path runs in hand-written CUDA behind the C ABI of include/vrwkv_b200.h.
"""
__version__ = "0.1.0"
