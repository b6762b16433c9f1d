"""cuembed_b200 -- B200-native (sm_100a) embedding lookup kernels behind the
cuEmbed host API: forward gather-pool, index transpose, deterministic backward.

The compute path is libcuembed_b200.so (hand-written CUDA, C ABI declared in
include/cuembed_b200.h); this package is the Python mirror of the reference's
operator interface plus the synthetic-workload generator used by the
benchmark.  Importing the package does not touch the GPU; the first call
loads (and if necessary builds) the library and fails loudly if it cannot.
"""
from .api import (OPT_ADAGRAD, OPT_SGD, CombineMode, ComputeCompressedGradIndices, CuEmbedError,
                  EmbeddingBackward, EmbeddingBackwardUpdate, EmbeddingForward,
                  EmbeddingForwardMulti, ExtractRowIdsForConcat,
                  ExtractRowIdsFromCSR, ExtractRowIdsFromFixed, ShardFinalize,
                  ShardSelect, Transpose, backward_workspace_bytes, launch_count,
                  EmbeddingForwardHot, HotRowsFromSorted, forward_hot_capacity,
                  TransposeFixed, EmbeddingForwardMapped, DebugCheckLookup)

__all__ = [
    "CombineMode", "CuEmbedError", "EmbeddingForward", "EmbeddingBackward", "EmbeddingBackwardUpdate", "EmbeddingForwardMulti", "OPT_SGD", "OPT_ADAGRAD",
    "ExtractRowIdsFromFixed", "ExtractRowIdsFromCSR", "ExtractRowIdsForConcat",
    "Transpose", "ComputeCompressedGradIndices", "backward_workspace_bytes",
    "launch_count", "ShardSelect", "ShardFinalize", "EmbeddingForwardHot", "HotRowsFromSorted",
    "forward_hot_capacity", "TransposeFixed", "EmbeddingForwardMapped", "DebugCheckLookup",
]
