"""ctypes binding of libslideo_b200.so -- 1:1 with include/slideo_b200.h (the drop-in C ABI).

The library is the product; this file only declares its symbols.  There is no fallback of any kind: if the
shared object is missing, `load()` raises with the build command, and every compute entry point fails with
SLIDEO_B200_E_CUDA when no sm_100 device is present.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libslideo_b200.so")

OK = 0
E_INVALID_ARG, E_CUDA, E_OOM, E_NOTIMPL, E_STATE, E_CAPACITY, E_INTERNAL = -1, -2, -3, -4, -5, -6, -7
STATUS_NAMES = {0: "OK", -1: "E_INVALID_ARG", -2: "E_CUDA", -3: "E_OOM", -4: "E_NOTIMPL", -5: "E_STATE",
                -6: "E_CAPACITY", -7: "E_INTERNAL"}
DESC_ORB256, DESC_SIFT128 = 0, 1
ABI_VERSION = 1

c_i32, c_f32, c_sz, c_vp = ctypes.c_int32, ctypes.c_float, ctypes.c_size_t, ctypes.c_void_p
c_i32p = ctypes.POINTER(ctypes.c_int32)


class Config(ctypes.Structure):
    """slideo_b200_config (defaults = the reference's literals, feature_extractor.rs:13-23, lib.rs:266,275)."""
    _fields_ = [("abi_version", c_i32), ("device", c_i32), ("nfeatures", c_i32), ("scale_factor", c_f32),
                ("nlevels", c_i32), ("edge_threshold", c_i32), ("patch_size", c_i32), ("fast_threshold", c_i32),
                ("knn_k", c_i32), ("vote_ratio", c_f32), ("descriptor_kind", c_i32), ("max_batch", c_i32),
                ("keep_matches", c_i32), ("geometric_verification", c_i32), ("knn_impl", c_i32), ("reserved", c_i32 * 1)]


class FrameResult(ctypes.Structure):
    _fields_ = [("best_slide", c_i32), ("votes", c_i32), ("n_keypoints", c_i32)]


class Match(ctypes.Structure):
    """Field for field the reference's KeyedDMatch (flann.rs:51-59)."""
    _fields_ = [("query_idx", c_i32), ("train_idx", c_i32), ("source", c_i32), ("distance", c_f32)]


TOP_SLIDES, TOP_RATED = 40, 10


class VerifyResult(ctypes.Structure):
    """slideo_b200_verify_result: RANSAC gate of one frame (lib.rs:284-333)."""
    _fields_ = [("n_candidates", c_i32), ("n_survivors", c_i32), ("cand_page", c_i32 * TOP_SLIDES), ("cand_votes", c_i32 * TOP_SLIDES),
                ("cand_rating", c_i32 * TOP_SLIDES), ("survivor_page", c_i32 * TOP_RATED), ("survivor_rating", c_i32 * TOP_RATED)]


class Decision(ctypes.Structure):
    """slideo_b200_decision: warp + similarity gate of one frame (lib.rs:335-389)."""
    _fields_ = [("image", c_i32), ("n_rated", c_i32), ("rated_page", c_i32 * TOP_RATED), ("rated_similarity", c_f32 * TOP_RATED),
                ("refined_matrix", (ctypes.c_double * 4) * TOP_RATED)]


class Timings(ctypes.Structure):
    _fields_ = [("ms_detect", c_f32), ("ms_knn", c_f32), ("ms_vote", c_f32), ("ms_h2d", c_f32),
                ("knn_pairs", ctypes.c_int64), ("knn_launches", ctypes.c_int64), ("kernel_launches", ctypes.c_int64),
                ("frames", ctypes.c_int64), ("ms_total", c_f32), ("ms_verify", c_f32)]


# name -> (restype, argtypes); every symbol include/slideo_b200.h declares
SYMBOLS = {
    "slideo_b200_default_config": (c_i32, [ctypes.POINTER(Config)]),
    "slideo_b200_create": (c_i32, [ctypes.POINTER(Config), ctypes.POINTER(c_vp)]),
    "slideo_b200_destroy": (c_i32, [c_vp]),
    "slideo_b200_last_error": (ctypes.c_char_p, [c_vp]),
    "slideo_b200_version": (ctypes.c_char_p, []),
    "slideo_b200_add_page_gray8": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32p]),
    "slideo_b200_add_page_descriptors": (c_i32, [c_vp, c_vp, c_i32]),
    "slideo_b200_add_page_features": (c_i32, [c_vp, c_vp, c_vp, c_i32]),
    "slideo_b200_finalize_pool": (c_i32, [c_vp]),
    "slideo_b200_pool_info": (c_i32, [c_vp, c_i32p, c_i32p]),
    "slideo_b200_pool_export": (c_i32, [c_vp, c_vp, c_vp]),
    "slideo_b200_pool_import": (c_i32, [c_vp, c_vp, c_i32, c_vp, c_i32]),
    "slideo_b200_pool_reserve": (c_i32, [c_vp, c_i32, c_i32]),
    "slideo_b200_pool_device_view": (c_i32, [c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_sz), ctypes.POINTER(c_vp),
                                             ctypes.POINTER(c_sz)]),
    "slideo_b200_pool_commit": (c_i32, [c_vp]),
    "slideo_b200_pool_points_device_view": (c_i32, [c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_sz), c_i32p, c_i32]),
    "slideo_b200_pool_pages_device_view": (c_i32, [c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_sz), c_i32p, c_i32p, c_i32, c_i32]),
    "slideo_b200_match_frames_bgr8": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_sz, c_vp]),
    "slideo_b200_match_frames_bgr8_device": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_sz, c_vp]),
    "slideo_b200_submit_frames_bgr8": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_sz, ctypes.POINTER(ctypes.c_int64)]),
    "slideo_b200_submit_frames_bgr8_device": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_sz, ctypes.POINTER(ctypes.c_int64)]),
    "slideo_b200_collect": (c_i32, [c_vp, ctypes.c_int64, c_vp, c_i32, c_i32p]),
    "slideo_b200_match_descriptors": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_vp]),
    "slideo_b200_get_matches": (c_i32, [c_vp, c_i32, c_vp, c_i32, c_i32p]),
    "slideo_b200_get_verification": (c_i32, [c_vp, c_i32, c_i32, c_vp]),
    "slideo_b200_get_decisions": (c_i32, [c_vp, c_i32, c_i32, c_vp]),
    "slideo_b200_mark_changed_bgr8": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_sz, c_i32, c_vp, c_vp]),
    "slideo_b200_mark_changed_bgr8_device": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_sz, c_i32, c_vp, c_vp]),
    "slideo_b200_extract_orb": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_i32, c_i32p]),
    "slideo_b200_debug_fetch": (c_i32, [c_vp, c_i32, c_i32, c_vp, c_sz, c_i32p, c_i32p]),
    "slideo_b200_extract_sift": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_i32, c_i32p]),
    "slideo_b200_debug_fetch_sift": (c_i32, [c_vp, c_i32, c_i32, c_vp, c_sz, c_i32p, c_i32p, c_i32p]),
    "slideo_b200_bf_knn_hamming": (c_i32, [c_vp, c_vp, c_i32, c_vp, c_i32, c_i32, c_vp, c_vp]),
    "slideo_b200_bf_knn_hamming_device": (c_i32, [c_vp, c_vp, c_i32, c_vp, c_i32, c_i32, c_vp]),
    "slideo_b200_bf_knn_l2": (c_i32, [c_vp, c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "slideo_b200_bf_knn_l2_device": (c_i32, [c_vp, c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "slideo_b200_host_alloc": (c_i32, [ctypes.POINTER(c_vp), c_sz]),
    "slideo_b200_host_free": (c_i32, [c_vp]),
    "slideo_b200_get_timings": (c_i32, [c_vp, ctypes.POINTER(Timings), c_i32]),
    "slideo_b200_set_progress_callback": (c_i32, [c_vp, c_vp, c_vp]),
    "slideo_b200_microbench": (c_i32, [c_vp, c_i32, ctypes.POINTER(ctypes.c_double)]),
    "slideo_b200_synchronize": (c_i32, [c_vp]),
}

PROGRESS_FN = ctypes.CFUNCTYPE(None, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_void_p)

_LIB = None


def load() -> ctypes.CDLL:
    """Loads the in-tree shared library (built by `make -C slideo_b200/csrc` or `__graft_entry__.build()`)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                              "(there is no CPU or PyTorch fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB
