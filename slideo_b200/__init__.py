"""slideo_b200 -- B200-native (sm_100a) drop-in for the per-frame hot path of hediet/slideo's
`crates/matching-opencv`: ORB extraction -> brute-force k-NN against the pooled slide descriptors -> 1.05-ratio vote.

The product is `libslideo_b200.so` (hand-written CUDA behind the C ABI of include/slideo_b200.h).  This package is
the host-side mirror of the reference's plugin interface plus a ctypes binding; it contains no arithmetic and no
fallback path -- without the built library or without an sm_100 GPU every call raises.
"""
from . import ffi
from .context import Context, PinnedBuffer, SlideoError, default_config
from .matcher import (B200ImageVideoMatcher, B200VideoMatcher, B200VideoMatcherTask, Matching, ProgressReporter)

__all__ = ["ffi", "Context", "PinnedBuffer", "SlideoError", "default_config", "B200ImageVideoMatcher", "B200VideoMatcher",
           "B200VideoMatcherTask", "Matching", "ProgressReporter"]
