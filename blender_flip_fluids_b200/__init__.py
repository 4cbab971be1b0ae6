"""B200-native particle<->grid substep (P2G, G2P, RK3 advection) of the FLIP Fluids engine.

The compute path is libffb200.so (hand-written CUDA for sm_100a behind the C ABI of
include/ffb200.h); ``engine`` is its ctypes host side, ``scenes`` builds synthetic inputs,
``slab`` runs the z-slab multi-GPU decomposition. There is no CPU fallback.
"""
__all__ = ["engine", "scenes", "build"]
