"""graphflow_b200 -- B200-native (sm_100a) second-order CCN message passing: StackTensor3D -> RisiContraction_18
(forward + backward) -> feature mix, behind the C-ABI of include/ccn_b200.h.

    from graphflow_b200 import Context
    ctx = Context(device=0)
    out = ctx.contract18_forward(T, adj)            # T [B,N,N,N,C], adj [B,N,N] -> [B,N,N,18C]
    gT  = ctx.contract18_backward(gout, adj)        # transpose

The header-compatible C++ classes (RisiContraction_18_gpu, StackTensor3D, MatMul_gpu replacements) are in
include/graphflow_b200/.  There is no CPU or eager fallback: without the built library every call raises CCNError.
"""
from ._lib import ADJ_POSITIVE_PART, ADJ_RAW, CCNError, LIB_PATH, load  # noqa: F401
from .ops import NUM_CONTRACTIONS, Context  # noqa: F401

__all__ = ["Context", "CCNError", "ADJ_POSITIVE_PART", "ADJ_RAW", "NUM_CONTRACTIONS", "load", "LIB_PATH"]
