"""sniper_b200: the SNIPER per-chip training hot path as hand-written sm_100a CUDA behind a C-ABI."""
__version__ = "0.1.0"
