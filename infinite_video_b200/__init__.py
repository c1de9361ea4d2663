"""B200-native long-term-memory (LTM) consolidation path of infinity-Video.

Public surface:
  LongTermAttention  -- drop-in for the reference module (same ctor / forward signature)
  BatchedLTM         -- the same path for Bv independent videos with explicit state and uniforms
"""
__version__ = "0.1.0"

from . import tables  # noqa: F401


def __getattr__(name):
    # lazy: the CUDA library is only needed once a module is constructed
    if name in ("LongTermAttention",):
        from .ltm import LongTermAttention
        return LongTermAttention
    if name in ("BatchedLTM", "BatchedRectLTM", "BatchedGaussLTM"):
        from . import batched
        return getattr(batched, name)
    raise AttributeError(name)
