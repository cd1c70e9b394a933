"""Host-side mirror of the reference's plugin interface `crates/matching` (crates/matching/src/lib.rs:7-40),
implemented over the C ABI the way `crates/matching-opencv` implements it over OpenCV
(crates/matching-opencv/src/lib.rs:33-245).  Same names, same call order, same argument meaning:

    matcher = B200ImageVideoMatcher()                              # OpenCVImageVideoMatcher::default()   lib.rs:33-35
    vm      = matcher.create_video_matcher(images, reporter)       # ImageVideoMatcher                    lib.rs:37-64
    task    = vm.match_images_with_video(video_path, reporter)     # VideoMatcher                         lib.rs:140-158
    result  = task.process()                                       # VideoMatcherTask -> [Matching]       lib.rs:169-245

What differs, and why (DESIGN.md "Boundary"):
  * the per-frame decision: with `geometric_verification=True` (default) the reference's complete tail runs on the GPU --
    top-40 by votes, RANSAC rating gate (lib.rs:284-333), warp + similarity gate (lib.rs:335-389) -- and the frame maps to the
    most similar survivor or to `image=None`, exactly like `Matching.image`.  With `geometric_verification=False` the decision
    is the head of the vote ranking (argmax-votes slide, lib.rs:268-295) and `min_votes` stands in for the gates.
  * frames are matched in batches on the GPU instead of one rayon task per frame (lib.rs:213-214).
  * decoding stays on the host (the reference uses OpenCV's FFmpeg VideoCapture, video_capture.rs:15-57); any
    iterable of (frame_bgr, seconds, frame_idx) can be passed instead of a path.
Errors: the reference panics (unwrap); this mirror raises SlideoError / ValueError.
"""
from __future__ import annotations

import dataclasses
import math
import os
from typing import Any, Callable, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from . import ffi
from .context import Context, default_config

SAMPLE_INTERVAL_S = 5.0        # lib.rs:145,175
CHANGED_THRESHOLD = 0.98       # video_capture.rs:98
SMALL_IMAGE_AREA = 300 * 400   # image_utils.rs:11


class ProgressReporter:
    """crates/matching/src/progress.rs:3-17 -- a (processed, total, message) callback."""

    def __init__(self, handler: Optional[Callable[[int, int, str], None]] = None):
        self._handler = handler

    def report(self, processed_count: int, total_count: int, message: str) -> None:
        if self._handler is not None:
            self._handler(int(processed_count), int(total_count), message)


@dataclasses.dataclass
class Matching:
    """crates/matching/src/lib.rs:35-40 (video_time in seconds instead of a Duration)."""
    video_time: float
    video_frame_idx: int
    image: Optional[Any]
    votes: int = 0               # extra: the hot path's match-count for the chosen slide


def _image_gray(image) -> np.ndarray:
    """A MatchableImage is anything with get_path() (lib.rs:31-33); arrays / objects with .gray are accepted too
    so synthetic pages need no files.  Pages are read as 8-bit gray exactly like lib.rs:98 (imread(path, 0))."""
    if isinstance(image, np.ndarray):
        return image
    g = getattr(image, "gray", None)
    if g is not None:
        return g() if callable(g) else g
    import cv2  # host-side PNG decode only (the reference decodes with OpenCV's imread too)
    path = os.fspath(image.get_path())
    img = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
    if img is None:
        raise ValueError(f"Could not read image file '{path}'")   # lib.rs:99-101 panics with the same text
    return img


class B200ImageVideoMatcher:
    """Drop-in for OpenCVImageVideoMatcher (lib.rs:33-75)."""

    def __init__(self, device: int = 0, min_votes: int = 1, geometric_verification: bool = True, **config_overrides):
        self._device = device
        self._min_votes = min_votes
        self._overrides = dict(config_overrides, geometric_verification=2 if geometric_verification else 0)

    def create_video_matcher(self, images: Sequence[Any], progress_reporter: Optional[ProgressReporter] = None
                             ) -> "B200VideoMatcher":
        reporter = progress_reporter or ProgressReporter()
        ctx = Context(default_config(device=self._device, **self._overrides))
        total = len(images)
        for i, im in enumerate(images):                     # lib.rs:45-56 (rayon par_iter there; GPU-serial here)
            ctx.add_page_gray8(_image_gray(im))
            reporter.report(i + 1, total, "Preprocessing pdf pages...")
        ctx.finalize_pool()
        return B200VideoMatcher(ctx, list(images), self._min_votes)


class B200VideoMatcher:
    """Drop-in for OpenCVVideoMatcher (lib.rs:134-158)."""

    def __init__(self, ctx: Context, images: List[Any], min_votes: int):
        self.ctx = ctx
        self.images = images
        self.min_votes = min_votes

    def match_images_with_video(self, video, progress_reporter: Optional[ProgressReporter] = None
                                ) -> "B200VideoMatcherTask":
        reporter = progress_reporter or ProgressReporter()
        task = B200VideoMatcherTask(self, video, reporter)
        reporter.report(0, task.frames_to_process_hint(), "")     # "Reports immediate progress" lib.rs:150
        return task


def sampled_frames(path: str, interval_s: float = SAMPLE_INTERVAL_S) -> Iterator[Tuple[np.ndarray, float, int]]:
    """VideoCaptureIter (video_capture.rs:9-57): grab every frame, retrieve when frame_idx % floor(fps*interval) < 1."""
    import cv2
    cap = cv2.VideoCapture(os.fspath(path))
    if not cap.isOpened():
        raise ValueError(f"Could not open video '{path}'")
    fps = cap.get(cv2.CAP_PROP_FPS)
    step = math.floor(fps * interval_s)
    try:
        while True:
            frame_idx = cap.get(cv2.CAP_PROP_POS_FRAMES)
            if not cap.grab():
                return
            if step <= 0 or math.fmod(frame_idx, step) < 1.0:
                ok, frame = cap.retrieve()
                if ok:
                    yield frame, frame_idx / fps, int(frame_idx)
    finally:
        cap.release()


def video_totals(path: str) -> Tuple[float, float]:
    """(total_frames, total_seconds) as VideoCaptureIter::total_frames/total_time (video_capture.rs:30-37)."""
    import cv2
    cap = cv2.VideoCapture(os.fspath(path))
    n, fps = cap.get(cv2.CAP_PROP_FRAME_COUNT), cap.get(cv2.CAP_PROP_FPS)
    cap.release()
    return n, (n / fps if fps else 0.0)


def to_small_image(img: np.ndarray) -> np.ndarray:
    """image_utils.rs:8-19: INTER_AREA resize to ~300x400 area, aspect preserved (host side, prefilter only)."""
    import cv2
    h, w = img.shape[:2]
    f = np.sqrt(np.float32(SMALL_IMAGE_AREA) / np.float32(w * h), dtype=np.float32)     # f32 like the reference
    return cv2.resize(img, (int(np.float32(w) * f), int(np.float32(h) * f)), interpolation=cv2.INTER_AREA)   # `as i32` truncates


def compute_similarity(a: np.ndarray, b: np.ndarray) -> float:
    """image_utils.rs:21-27: 1 - ||a-b||_2 / sqrt(255^2 * 3 * pixels)."""
    d = a.astype(np.float64) - b.astype(np.float64)
    return 1.0 - math.sqrt(float((d * d).sum())) / math.sqrt(255.0 * 255.0 * 3.0 * a.shape[0] * a.shape[1])


def mark_similar_gpu(ctx: Context, frames: Iterable[Tuple[np.ndarray, float, int]], batch: int = 32
                     ) -> Iterator[Tuple[bool, np.ndarray, float, int]]:
    """MarkSimilarIter (video_capture.rs:60-103) through the library (K13, slideo_b200_mark_changed_bgr8)."""
    buf: List[Tuple[np.ndarray, float, int]] = []
    first = True

    def flush():
        nonlocal first
        if not buf:
            return
        changed, _ = ctx.mark_changed_bgr8(np.stack([b[0] for b in buf]), reset=first)
        first = False
        for c, (f, t, i) in zip(changed, buf):
            yield bool(c), f, t, i
        buf.clear()

    for item in frames:
        if buf and item[0].shape != buf[0][0].shape:
            yield from flush()
            first = True
        buf.append(item)
        if len(buf) >= batch:
            yield from flush()
    yield from flush()


def mark_similar(frames: Iterable[Tuple[np.ndarray, float, int]]) -> Iterator[Tuple[bool, np.ndarray, float, int]]:
    """MarkSimilarIter (video_capture.rs:60-103) on the host with cv2 (kept for environments that prefilter before upload)."""
    last = None
    for frame, t, idx in frames:
        small = to_small_image(frame)
        sim = compute_similarity(last, small) if last is not None and last.shape == small.shape else 0.0
        last = small
        yield sim < CHANGED_THRESHOLD, frame, t, idx


class B200VideoMatcherTask:
    """Drop-in for OpenCVVideoMatcherTask (lib.rs:160-245)."""

    def __init__(self, vm: B200VideoMatcher, video, reporter: ProgressReporter, prefilter: bool = True):
        self.vm = vm
        self.video = video
        self.reporter = reporter
        self.prefilter = prefilter

    def _is_path(self) -> bool:
        return isinstance(self.video, (str, os.PathLike))

    def frames_to_process_hint(self) -> int:
        if self._is_path():
            return int(video_totals(self.video)[1] / SAMPLE_INTERVAL_S)     # lib.rs:146-148
        try:
            return len(self.video)
        except TypeError:
            return 0

    def process(self) -> List[Matching]:
        ctx, images = self.vm.ctx, self.vm.images
        results: List[Matching] = []
        name = os.path.basename(os.fspath(self.video)) if self._is_path() else "frames"
        if self._is_path():
            total_frames, total_time = video_totals(self.video)
            frames_to_process = int(total_time / SAMPLE_INTERVAL_S)
            results.append(Matching(total_time, int(total_frames), None))    # lib.rs:186-190 end marker
            source = sampled_frames(self.video)
        else:
            source = iter(self.video)
            frames_to_process = self.frames_to_process_hint()
        stream = mark_similar_gpu(ctx, source) if self.prefilter else ((True, f, t, i) for f, t, i in source)

        done = 0
        batch: List[Tuple[np.ndarray, float, int]] = []
        B = int(ctx.cfg.max_batch)

        def flush():
            nonlocal done
            if not batch:
                return
            # one geometry per call: group by frame shape (videos have one)
            shapes = {}
            for j, (f, _, _) in enumerate(batch):
                shapes.setdefault(f.shape, []).append(j)
            for shape, idxs in shapes.items():
                frames = np.stack([batch[j][0] for j in idxs])
                res = ctx.match_frames_bgr8(frames)
                dec = ctx.get_decisions(0, len(idxs)) if ctx.cfg.geometric_verification >= 2 else None
                ver = ctx.get_verification(0, len(idxs)) if ctx.cfg.geometric_verification == 1 else None
                for n, (j, (best, votes, _nkp)) in enumerate(zip(idxs, res)):
                    if dec is not None:      # lib.rs:383-389: the most similar survivor of both gates
                        img = images[dec[n]["image"]] if dec[n]["image"] >= 0 else None
                    elif ver is not None:    # lib.rs:329-333 only: best-rated survivor of the RANSAC gate
                        surv = ver[n]["survivors"]
                        img = images[surv[0][0]] if surv else None
                    else:
                        img = images[best] if best >= 0 and votes >= self.vm.min_votes else None
                    results.append(Matching(batch[j][1], batch[j][2], img, int(votes)))
                    done += 1
                    self.reporter.report(done, frames_to_process, f"Processing frames of '{name}'...")
            batch.clear()

        for changed, frame, t, idx in stream:
            if not changed:                       # lib.rs:207-210
                done += 1
                self.reporter.report(done, frames_to_process, f"Processing frames of '{name}'...")
                continue
            batch.append((frame, t, idx))
            if len(batch) >= B:
                flush()
        flush()
        self.reporter.report(frames_to_process, frames_to_process, "Finished!")    # lib.rs:223-227

        # lib.rs:229-244: sort by time, drop consecutive equal images
        results.sort(key=lambda m: m.video_time)
        cleaned: List[Matching] = []
        last: Optional[Matching] = None
        for m in results:
            if last is not None and _same_image(last.image, m.image):
                continue
            last = m
            cleaned.append(m)
        return cleaned


def _same_image(a, b) -> bool:
    if a is None or b is None:
        return a is None and b is None
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        return a is b
    return a == b
