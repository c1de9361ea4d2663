"""B200-native long-term-memory (LTM) consolidation path of infinity-Video.

Public surface:
  LongTermAttention  -- drop-in for the reference module (same ctor / forward signature)
  BatchedRectLTM / BatchedGaussLTM -- the same path for Bv independent videos with explicit state and uniforms
  CrossAttentionLTM  -- the caller's whole cross-attention branch (short-term attention + LTM + alpha blend)
"""
__version__ = "0.1.0"

from . import tables  # noqa: F401


def __getattr__(name):
    # lazy: the CUDA library is only needed once a module is constructed
    if name in ("LongTermAttention",):
        from .ltm import LongTermAttention
        return LongTermAttention
    if name in ("BatchedRectLTM", "BatchedGaussLTM"):
        from . import batched
        return getattr(batched, name)
    if name == "BatchedLTM":                       # the live (rect / "gibbs") variant
        from .batched import BatchedRectLTM
        return BatchedRectLTM
    if name == "CrossAttentionLTM":
        from .cross_attention import CrossAttentionLTM
        return CrossAttentionLTM
    raise AttributeError(name)
