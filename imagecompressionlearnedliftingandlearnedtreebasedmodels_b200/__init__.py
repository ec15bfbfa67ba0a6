"""B200-native (sm_100a) learned-lifting DWT + tree-based entropy-model hot path.

Drop-in for the reference's ``graphs`` modules (same class names, constructor
signatures, ``state_dict`` layout); the arithmetic runs in hand-written CUDA behind
the C ABI of ``include/ll_api.h`` (``libll_b200.so``).  There is no CPU fallback.
"""
__all__ = ["build", "ops"]
