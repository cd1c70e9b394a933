"""`Context`: a thin object wrapper over one `slideo_b200_ctx*` (one per GPU, not re-entrant).

Every method is one C-ABI call on numpy HOST buffers (or raw device pointers for the `*_device` variants);
status codes become `SlideoError`.  Mirrors what a Rust `matching-b200` crate does over the same symbols
(INTEGRATION.md).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np

from . import ffi


class SlideoError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{ffi.STATUS_NAMES.get(status, status)}: {message}")
        self.status = status
        self.message = message


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def default_config(**overrides) -> ffi.Config:
    cfg = ffi.Config()
    st = ffi.load().slideo_b200_default_config(ctypes.byref(cfg))
    if st != ffi.OK:
        raise SlideoError(st, "default_config failed")
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise TypeError(f"unknown config field {k!r}")
        setattr(cfg, k, v)
    return cfg


class Context:
    def __init__(self, cfg: Optional[ffi.Config] = None, **overrides):
        self._lib = ffi.load()
        self._h = ctypes.c_void_p()
        self.cfg = cfg if cfg is not None else default_config(**overrides)
        st = self._lib.slideo_b200_create(ctypes.byref(self.cfg), ctypes.byref(self._h))
        if st != ffi.OK:
            raise SlideoError(st, (self._lib.slideo_b200_last_error(None) or b"").decode())

    # ---- lifecycle -------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.slideo_b200_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st: int):
        if st != ffi.OK:
            raise SlideoError(st, (self._lib.slideo_b200_last_error(self._h) or b"").decode())

    @property
    def k(self) -> int:
        return int(self.cfg.knn_k)

    @property
    def desc_kind(self) -> int:
        return int(self.cfg.descriptor_kind)

    def _desc_array(self, desc) -> np.ndarray:
        if self.desc_kind == ffi.DESC_ORB256:
            a = np.ascontiguousarray(desc, np.uint8)
            if a.ndim != 2 or a.shape[1] != 32:
                a = a.reshape(-1, 32)
        else:
            a = np.ascontiguousarray(desc, np.float32)
            if a.ndim != 2 or a.shape[1] != 128:
                a = a.reshape(-1, 128)
        return a

    # ---- page pool -------------------------------------------------------------------------------------
    def add_page_gray8(self, gray: np.ndarray) -> int:
        """One page as the 8-bit gray `imread(path, 0)` yields (lib.rs:98).  Returns its keypoint count."""
        if gray.ndim != 2 or gray.dtype != np.uint8:
            raise TypeError("gray must be a 2-d uint8 array")
        if gray.strides[1] != 1:
            gray = np.ascontiguousarray(gray)
        n = ctypes.c_int32()
        self._ck(self._lib.slideo_b200_add_page_gray8(self._h, _ptr(gray), gray.shape[1], gray.shape[0], gray.strides[0],
                                                      ctypes.byref(n)))
        return n.value

    def add_page_descriptors(self, desc) -> None:
        a = self._desc_array(desc)
        self._ck(self._lib.slideo_b200_add_page_descriptors(self._h, _ptr(a), len(a)))

    def add_page_features(self, desc, pts) -> None:
        """Descriptors + KeyPoint.pt (n x 2 float32, level-0 coordinates) of a page extracted elsewhere."""
        a = self._desc_array(desc)
        p = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        if len(p) != len(a):
            raise ValueError("desc and pts disagree")
        self._ck(self._lib.slideo_b200_add_page_features(self._h, _ptr(a), _ptr(p), len(a)))

    def finalize_pool(self) -> None:
        self._ck(self._lib.slideo_b200_finalize_pool(self._h))

    def pool_info(self):
        n, p = ctypes.c_int32(), ctypes.c_int32()
        self._ck(self._lib.slideo_b200_pool_info(self._h, ctypes.byref(n), ctypes.byref(p)))
        return n.value, p.value

    def pool_export(self):
        n, p = self.pool_info()
        desc = np.empty((n, 32), np.uint8) if self.desc_kind == ffi.DESC_ORB256 else np.empty((n, 128), np.float32)
        offs = np.empty(p + 1, np.int32)
        self._ck(self._lib.slideo_b200_pool_export(self._h, _ptr(desc), _ptr(offs)))
        return desc, offs

    def pool_import(self, desc, page_offsets) -> None:
        a = self._desc_array(desc)
        offs = np.ascontiguousarray(page_offsets, np.int32)
        self._ck(self._lib.slideo_b200_pool_import(self._h, _ptr(a), len(a), _ptr(offs), len(offs) - 1))

    def pool_reserve(self, n_desc: int, n_pages: int) -> None:
        self._ck(self._lib.slideo_b200_pool_reserve(self._h, n_desc, n_pages))

    def pool_device_view(self):
        """(d_desc ptr, desc_bytes, d_page_offsets ptr, offsets_bytes) -- raw device pointers for an NCCL broadcast."""
        d, db, o, ob = ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_void_p(), ctypes.c_size_t()
        self._ck(self._lib.slideo_b200_pool_device_view(self._h, ctypes.byref(d), ctypes.byref(db), ctypes.byref(o),
                                                        ctypes.byref(ob)))
        return d.value or 0, db.value, o.value or 0, ob.value

    def pool_points_device_view(self, received: bool = False):
        """(d_pt ptr, bytes, has_points) of the pooled keypoint coordinates (geometric verification across GPUs)."""
        d, b, h = ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_int32()
        self._ck(self._lib.slideo_b200_pool_points_device_view(self._h, ctypes.byref(d), ctypes.byref(b), ctypes.byref(h), int(received)))
        return d.value or 0, b.value, bool(h.value)

    def pool_pages_device_view(self, set_w: int = 0, set_h: int = 0):
        """(d_small ptr, bytes, page_w, page_h) of the pages' small images (warp + similarity gate across GPUs); on a reserved ctx
        pass the sender's page size to allocate the receiving buffer."""
        d, b, w, h = ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_int32(), ctypes.c_int32()
        self._ck(self._lib.slideo_b200_pool_pages_device_view(self._h, ctypes.byref(d), ctypes.byref(b), ctypes.byref(w), ctypes.byref(h),
                                                              set_w, set_h))
        return d.value or 0, b.value, w.value, h.value

    def pool_commit(self) -> None:
        self._ck(self._lib.slideo_b200_pool_commit(self._h))

    # ---- the per-frame hot path ------------------------------------------------------------------------
    @staticmethod
    def _results(res, n):
        a = np.frombuffer(res, dtype=np.int32).reshape(n, 3).copy() if n else np.zeros((0, 3), np.int32)
        return a

    def match_frames_bgr8(self, frames: np.ndarray) -> np.ndarray:
        """frames: [n, h, w, 3] uint8 BGR on the HOST.  Returns int32 [n, 3] = (best_slide, votes, n_keypoints)."""
        if frames.ndim == 3:
            frames = frames[None]
        if frames.ndim != 4 or frames.shape[3] != 3 or frames.dtype != np.uint8:
            raise TypeError("frames must be [n, h, w, 3] uint8")
        if frames.strides[3] != 1 or frames.strides[2] != 3:
            frames = np.ascontiguousarray(frames)
        n, h, w, _ = frames.shape
        res = (ffi.FrameResult * max(n, 1))()
        frame_stride = frames.strides[0] if n > 1 else frames.strides[1] * h   # a length-1 axis may carry any stride
        self._ck(self._lib.slideo_b200_match_frames_bgr8(self._h, _ptr(frames), n, w, h, frames.strides[1], frame_stride, res))
        return self._results(res, n)

    def match_frames_bgr8_ptr(self, host_ptr: int, n: int, w: int, h: int, stride: Optional[int] = None,
                              frame_stride: Optional[int] = None) -> np.ndarray:
        """Same on a raw HOST pointer (e.g. pinned memory from `host_alloc`)."""
        stride = stride or 3 * w
        frame_stride = frame_stride or stride * h
        res = (ffi.FrameResult * max(n, 1))()
        self._ck(self._lib.slideo_b200_match_frames_bgr8(self._h, ctypes.c_void_p(host_ptr), n, w, h, stride, frame_stride, res))
        return self._results(res, n)

    def match_frames_bgr8_device(self, dev_ptr: int, n: int, w: int, h: int, stride: Optional[int] = None,
                                 frame_stride: Optional[int] = None) -> np.ndarray:
        """Frames already resident in device memory (raw pointer, e.g. torch.Tensor.data_ptr())."""
        stride = stride or 3 * w
        frame_stride = frame_stride or stride * h
        res = (ffi.FrameResult * max(n, 1))()
        self._ck(self._lib.slideo_b200_match_frames_bgr8_device(self._h, ctypes.c_void_p(dev_ptr), n, w, h, stride,
                                                                frame_stride, res))
        return self._results(res, n)

    # asynchronous variant: submit returns a ticket at once, collect blocks until that ticket's results are complete
    def submit_frames_bgr8_ptr(self, host_ptr: int, n: int, w: int, h: int, stride: Optional[int] = None,
                               frame_stride: Optional[int] = None) -> int:
        """HOST frames (pinned memory makes the uploads asynchronous; the buffer must stay valid until `collect`)."""
        stride = stride or 3 * w
        frame_stride = frame_stride or stride * h
        t = ctypes.c_int64()
        self._ck(self._lib.slideo_b200_submit_frames_bgr8(self._h, ctypes.c_void_p(host_ptr), n, w, h, stride, frame_stride, ctypes.byref(t)))
        return t.value

    def submit_frames_bgr8(self, frames: np.ndarray) -> int:
        if frames.ndim != 4 or frames.shape[3] != 3 or frames.dtype != np.uint8 or not frames.flags["C_CONTIGUOUS"]:
            raise TypeError("frames must be a C-contiguous [n, h, w, 3] uint8 array (kept alive by the caller until collect)")
        n, h, w, _ = frames.shape
        return self.submit_frames_bgr8_ptr(frames.ctypes.data, n, w, h)

    def submit_frames_bgr8_device(self, dev_ptr: int, n: int, w: int, h: int, stride: Optional[int] = None,
                                  frame_stride: Optional[int] = None) -> int:
        stride = stride or 3 * w
        frame_stride = frame_stride or stride * h
        t = ctypes.c_int64()
        self._ck(self._lib.slideo_b200_submit_frames_bgr8_device(self._h, ctypes.c_void_p(dev_ptr), n, w, h, stride, frame_stride,
                                                                 ctypes.byref(t)))
        return t.value

    def collect(self, ticket: int, n: int) -> np.ndarray:
        """Results of one ticket: int32 [n, 3] = (best_slide, votes, n_keypoints)."""
        res = (ffi.FrameResult * max(n, 1))()
        got = ctypes.c_int32()
        self._ck(self._lib.slideo_b200_collect(self._h, ticket, res, n, ctypes.byref(got)))
        return self._results(res, got.value)

    def match_descriptors(self, desc, frame_offsets) -> np.ndarray:
        a = self._desc_array(desc)
        fo = np.ascontiguousarray(frame_offsets, np.int32)
        n = len(fo) - 1
        res = (ffi.FrameResult * max(n, 1))()
        self._ck(self._lib.slideo_b200_match_descriptors(self._h, _ptr(a), _ptr(fo), n, res))
        return self._results(res, n)

    def get_matches(self, frame_i: int):
        """k-NN rows of frame `frame_i` of the last match call as a structured array [n_query, k] (needs keep_matches)."""
        rows = ctypes.c_int32()
        self._ck(self._lib.slideo_b200_get_matches(self._h, frame_i, None, 0, ctypes.byref(rows)))
        dt = np.dtype([("query_idx", np.int32), ("train_idx", np.int32), ("source", np.int32), ("distance", np.float32)])
        out = np.empty((rows.value, self.k), dt)
        if rows.value:
            self._ck(self._lib.slideo_b200_get_matches(self._h, frame_i, _ptr(out), rows.value, ctypes.byref(rows)))
        return out

    def mark_changed_bgr8(self, frames: np.ndarray, reset: bool = False):
        """MarkSimilarIter (video_capture.rs:60-103) on consecutive SAMPLED frames [n, h, w, 3] (HOST): (changed bool[n], similarity f32[n])."""
        if frames.ndim == 3:
            frames = frames[None]
        frames = np.ascontiguousarray(frames, np.uint8)
        n, h, w, _ = frames.shape
        ch = np.zeros(max(n, 1), np.uint8)
        sim = np.zeros(max(n, 1), np.float32)
        self._ck(self._lib.slideo_b200_mark_changed_bgr8(self._h, _ptr(frames), n, w, h, 3 * w, 3 * w * h, int(reset), _ptr(ch), _ptr(sim)))
        return ch[:n].astype(bool), sim[:n]

    def mark_changed_bgr8_device(self, dev_ptr: int, n: int, w: int, h: int, reset: bool = False):
        ch = np.zeros(max(n, 1), np.uint8)
        sim = np.zeros(max(n, 1), np.float32)
        self._ck(self._lib.slideo_b200_mark_changed_bgr8_device(self._h, ctypes.c_void_p(dev_ptr), n, w, h, 3 * w, 3 * w * h, int(reset),
                                                                _ptr(ch), _ptr(sim)))
        return ch[:n].astype(bool), sim[:n]

    def get_verification(self, frame0: int, n: int):
        """RANSAC gate records of frames [frame0, frame0+n) of the last match_frames call (cfg.geometric_verification)."""
        res = (ffi.VerifyResult * max(n, 1))()
        self._ck(self._lib.slideo_b200_get_verification(self._h, frame0, n, res))
        out = []
        for i in range(n):
            r = res[i]
            out.append(dict(cand=[(r.cand_page[j], r.cand_votes[j], r.cand_rating[j]) for j in range(r.n_candidates)],
                            survivors=[(r.survivor_page[j], r.survivor_rating[j]) for j in range(r.n_survivors)]))
        return out

    def get_decisions(self, frame0: int, n: int):
        """Warp + similarity gate of frames [frame0, frame0+n) of the last match_frames call (cfg.geometric_verification == 2)."""
        res = (ffi.Decision * max(n, 1))()
        self._ck(self._lib.slideo_b200_get_decisions(self._h, frame0, n, res))
        out = []
        for i in range(n):
            r = res[i]
            out.append(dict(image=r.image, rated=[(r.rated_page[j], np.float32(r.rated_similarity[j])) for j in range(r.n_rated)],
                            refined=[[r.refined_matrix[j][c] for c in range(4)] for j in range(ffi.TOP_RATED)]))
        return out

    # ---- stage-level entry points ------------------------------------------------------------------------
    def extract_orb(self, img: np.ndarray, cap: int = 16384):
        """ORB::detectAndCompute (feature_extractor.rs:29-46) on one host image (gray [h,w] or BGR [h,w,3]).

        Returns kp_i [n,4] int32 {x_level,y_level,octave,score}, kp_f [n,4] f32 {pt.x,pt.y,size,angle}, desc [n,32] u8,
        canonical order (octave, y, x).
        """
        if img.dtype != np.uint8 or img.ndim not in (2, 3):
            raise TypeError("img must be uint8 [h,w] or [h,w,3]")
        ch = 1 if img.ndim == 2 else img.shape[2]
        img = np.ascontiguousarray(img)
        h, w = img.shape[:2]
        kp_i = np.empty((cap, 4), np.int32)
        kp_f = np.empty((cap, 4), np.float32)
        desc = np.empty((cap, 32), np.uint8)
        n = ctypes.c_int32()
        self._ck(self._lib.slideo_b200_extract_orb(self._h, _ptr(img), w, h, img.strides[0], ch, _ptr(kp_i), _ptr(kp_f),
                                                   _ptr(desc), cap, ctypes.byref(n)))
        return kp_i[:n.value].copy(), kp_f[:n.value].copy(), desc[:n.value].copy()

    def debug_fetch(self, what: int, level: int) -> np.ndarray:
        w, h = ctypes.c_int32(), ctypes.c_int32()
        self._ck(self._lib.slideo_b200_debug_fetch(self._h, what, level, None, 0, ctypes.byref(w), ctypes.byref(h)))
        if what == 2:
            out = np.empty(w.value, np.uint32)
            self._ck(self._lib.slideo_b200_debug_fetch(self._h, what, level, _ptr(out), out.nbytes, ctypes.byref(w), ctypes.byref(h)))
            return out
        out = np.empty((h.value, w.value), np.uint8)
        self._ck(self._lib.slideo_b200_debug_fetch(self._h, what, level, _ptr(out), out.nbytes, ctypes.byref(w), ctypes.byref(h)))
        return out

    def extract_sift(self, img: np.ndarray, cap: int = 65536):
        """SIFT::detectAndCompute (cv::SIFT::create() defaults) on one host image (gray [h,w] or BGR [h,w,3]).

        Returns kp_f [n,5] f32 {pt.x, pt.y, size, angle, response}, octave [n] int32 (packed like cv::KeyPoint::octave),
        desc [n,128] f32 (integer-valued), in OpenCV's output order.
        """
        if img.dtype != np.uint8 or img.ndim not in (2, 3):
            raise TypeError("img must be uint8 [h,w] or [h,w,3]")
        ch = 1 if img.ndim == 2 else img.shape[2]
        img = np.ascontiguousarray(img)
        h, w = img.shape[:2]
        n = ctypes.c_int32()
        self._ck(self._lib.slideo_b200_extract_sift(self._h, _ptr(img), w, h, img.strides[0], ch, None, None, None, 0, ctypes.byref(n)))
        cap = n.value
        kp_f = np.empty((cap, 5), np.float32)
        octv = np.empty(cap, np.int32)
        desc = np.empty((cap, 128), np.float32)
        self._ck(self._lib.slideo_b200_extract_sift(self._h, _ptr(img), w, h, img.strides[0], ch, _ptr(kp_f), _ptr(octv), _ptr(desc), cap,
                                                    ctypes.byref(n)))
        return kp_f[:n.value], octv[:n.value], desc[:n.value]

    def debug_fetch_sift(self, octave: int, layer: int) -> np.ndarray:
        """Gaussian layer of the scale space of the last SIFT call (image 0)."""
        w, h, no = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        self._ck(self._lib.slideo_b200_debug_fetch_sift(self._h, octave, layer, None, 0, ctypes.byref(w), ctypes.byref(h), ctypes.byref(no)))
        out = np.empty((h.value, w.value), np.float32)
        self._ck(self._lib.slideo_b200_debug_fetch_sift(self._h, octave, layer, _ptr(out), out.nbytes, ctypes.byref(w), ctypes.byref(h), ctypes.byref(no)))
        return out

    def bf_knn_hamming(self, q, t, k: int = 30):
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
        idx = np.empty((len(q), k), np.int32)
        dist = np.empty((len(q), k), np.int32)
        self._ck(self._lib.slideo_b200_bf_knn_hamming(self._h, _ptr(q), len(q), _ptr(t), len(t), k, _ptr(idx), _ptr(dist)))
        return idx, dist

    def bf_knn_hamming_device(self, d_q: int, nq: int, d_t: int, nt: int, k: int, d_keys_out: int) -> None:
        self._ck(self._lib.slideo_b200_bf_knn_hamming_device(self._h, ctypes.c_void_p(d_q), nq, ctypes.c_void_p(d_t), nt, k,
                                                             ctypes.c_void_p(d_keys_out)))

    def bf_knn_l2(self, q, t, k: int = 30):
        q = np.ascontiguousarray(q, np.float32)
        t = np.ascontiguousarray(t, np.float32)
        idx = np.empty((len(q), k), np.int32)
        dist = np.empty((len(q), k), np.float32)
        self._ck(self._lib.slideo_b200_bf_knn_l2(self._h, _ptr(q), len(q), _ptr(t), len(t), q.shape[1], k, _ptr(idx), _ptr(dist)))
        return idx, dist

    def bf_knn_l2_device(self, d_q: int, nq: int, d_t: int, nt: int, dim: int, k: int, d_idx: int, d_dist: int) -> None:
        self._ck(self._lib.slideo_b200_bf_knn_l2_device(self._h, ctypes.c_void_p(d_q), nq, ctypes.c_void_p(d_t), nt, dim, k,
                                                        ctypes.c_void_p(d_idx), ctypes.c_void_p(d_dist)))

    # ---- utilities ---------------------------------------------------------------------------------------
    def timings(self, reset: bool = False) -> dict:
        t = ffi.Timings()
        self._ck(self._lib.slideo_b200_get_timings(self._h, ctypes.byref(t), int(reset)))
        return {f: getattr(t, f) for f, _ in ffi.Timings._fields_}

    def set_progress_callback(self, handler=None) -> None:
        """handler(processed, total, message) or None -- the C form of crates/matching/src/progress.rs:3-17."""
        if handler is None:
            self._progress_cb = None
            self._ck(self._lib.slideo_b200_set_progress_callback(self._h, None, None))
            return
        self._progress_cb = ffi.PROGRESS_FN(lambda a, b, m, _u: handler(int(a), int(b), (m or b"").decode()))   # keep the thunk alive
        self._ck(self._lib.slideo_b200_set_progress_callback(self._h, ctypes.cast(self._progress_cb, ctypes.c_void_p), None))

    def microbench(self, which: int) -> float:
        v = ctypes.c_double()
        self._ck(self._lib.slideo_b200_microbench(self._h, which, ctypes.byref(v)))
        return v.value

    def synchronize(self) -> None:
        self._ck(self._lib.slideo_b200_synchronize(self._h))


class PinnedBuffer:
    """Pinned host memory from the library (makes the frame uploads asynchronous)."""

    def __init__(self, nbytes: int):
        self._lib = ffi.load()
        p = ctypes.c_void_p()
        st = self._lib.slideo_b200_host_alloc(ctypes.byref(p), nbytes)
        if st != ffi.OK:
            raise SlideoError(st, (self._lib.slideo_b200_last_error(None) or b"").decode())
        self.ptr = p.value
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(self.ptr))

    def close(self):
        if self.ptr:
            self.array = None
            self._lib.slideo_b200_host_free(ctypes.c_void_p(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
