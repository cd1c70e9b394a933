//! `matching-b200`: drop-in replacement for `matching-opencv` (crates/matching-opencv/src/lib.rs) behind the
//! `matching` traits (crates/matching/src/lib.rs:7-40).  SOURCE ONLY in this repository (no Rust toolchain in the
//! build image); it is the binding a maintainer adds next to `crates/matching-opencv`.
//!
//! Swap in `crates/app`: `Cargo.toml:25` `matching-opencv` -> `matching-b200`; `main.rs:14,69`
//! `OpenCVImageVideoMatcher::default()` -> `B200ImageVideoMatcher::default()`.
//!
//! What runs where: PNG read (`imread(path, 0)`, lib.rs:98) and video decode (`VideoCaptureIter`,
//! video_capture.rs:15-57) stay on the host exactly as in the reference; ORB, k-NN and the vote run on the GPU.
//! The reference's per-frame tail (RANSAC + warp gates, lib.rs:297-389) is not part of this hot path yet:
//! `min_votes` stands in for it.
mod ffi;

use ffi::*;
use matching::{ImageVideoMatcher, MatchableImage, Matching, ProgressReporter, VideoMatcher, VideoMatcherTask};
use opencv::{core::Mat, imgcodecs::imread, prelude::*};
use std::{ffi::CStr, path::{Path, PathBuf}, ptr, sync::Arc, time::Duration};

/// Owns one `slideo_b200_ctx*`.  The ctx is not re-entrant: `process()` is called from one thread per video
/// (crates/app/src/main.rs:87-93), which is exactly the contract.
struct Ctx(*mut slideo_b200_ctx);
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { slideo_b200_destroy(self.0) };
    }
}
impl Ctx {
    /// The reference unwraps every OpenCV call (lib.rs:95-101, flann.rs:16-46); non-zero status -> panic keeps the
    /// trait signatures unchanged.
    fn check(&self, st: i32) {
        if st != SLIDEO_B200_OK {
            let msg = unsafe { CStr::from_ptr(slideo_b200_last_error(self.0)) }.to_string_lossy().into_owned();
            panic!("slideo_b200 error {}: {}", st, msg);
        }
    }
}

pub struct B200ImageVideoMatcher {
    pub device: i32,
    pub min_votes: i32,
}
impl Default for B200ImageVideoMatcher {
    fn default() -> Self {
        Self { device: 0, min_votes: 1 }
    }
}

impl<'i> ImageVideoMatcher<'i> for B200ImageVideoMatcher {
    fn create_video_matcher<I: MatchableImage + Send + Sync + Copy + Eq + 'i>(
        &self,
        images: Vec<I>,
        progress_reporter: ProgressReporter,
    ) -> Box<dyn VideoMatcher<'i, I> + 'i> {
        let mut cfg = unsafe { std::mem::zeroed::<slideo_b200_config>() };
        unsafe { slideo_b200_default_config(&mut cfg) };
        cfg.device = self.device;
        let mut raw = ptr::null_mut();
        let st = unsafe { slideo_b200_create(&cfg, &mut raw) };
        if st != SLIDEO_B200_OK {
            let msg = unsafe { CStr::from_ptr(slideo_b200_last_error(ptr::null())) }.to_string_lossy().into_owned();
            panic!("slideo_b200_create failed ({}): {}", st, msg);
        }
        let ctx = Ctx(raw);
        let total = images.len() as u64;
        for (i, img) in images.iter().enumerate() {
            // lib.rs:93-104: imread(path, IMREAD_GRAYSCALE); the gray->BGR->gray round trip ORB performs is exact
            let gray: Mat = imread(&img.get_path().to_string_lossy(), 0).unwrap();
            if gray.size().unwrap().width == 0 {
                panic!("Could not read image file '{}'", img.get_path().to_string_lossy());
            }
            let (w, h) = (gray.cols(), gray.rows());
            let stride = gray.mat_step().get(0) as i32;
            ctx.check(unsafe { slideo_b200_add_page_gray8(ctx.0, gray.data().unwrap(), w, h, stride, ptr::null_mut()) });
            progress_reporter.report(i as u64 + 1, total, "Preprocessing pdf pages...");
        }
        ctx.check(unsafe { slideo_b200_finalize_pool(ctx.0) });
        Box::new(B200VideoMatcher { ctx: Arc::new(ctx), images: Arc::new(images), min_votes: self.min_votes })
    }
}

struct B200VideoMatcher<I> {
    ctx: Arc<Ctx>,
    images: Arc<Vec<I>>,
    min_votes: i32,
}

impl<'i, I: MatchableImage + Send + Sync + Copy + Eq + 'i> VideoMatcher<'i, I> for B200VideoMatcher<I> {
    fn match_images_with_video(&self, video_path: &Path, progress_reporter: ProgressReporter) -> Box<dyn VideoMatcherTask<I> + 'i> {
        // lib.rs:145-150: report (0, total_time / 5 s) immediately
        let interval = Duration::from_secs(5);
        let vid = video_capture::VideoCaptureIter::open(video_path, interval);
        let frames_to_process = (vid.total_time().as_secs_f64() / interval.as_secs_f64()) as u64;
        progress_reporter.report(0, frames_to_process, "");
        Box::new(B200VideoMatcherTask {
            ctx: self.ctx.clone(),
            images: self.images.clone(),
            min_votes: self.min_votes,
            video_path: video_path.to_owned(),
            progress_reporter,
        })
    }
}

struct B200VideoMatcherTask<I> {
    ctx: Arc<Ctx>,
    images: Arc<Vec<I>>,
    min_votes: i32,
    video_path: PathBuf,
    progress_reporter: ProgressReporter,
}

const BATCH: usize = 148;

impl<I: MatchableImage + Send + Sync + Copy + Eq> VideoMatcherTask<I> for B200VideoMatcherTask<I> {
    fn process(&self) -> Vec<Matching<I>> {
        let interval = Duration::from_secs(5);
        let vid = video_capture::VideoCaptureIter::open(&self.video_path, interval);
        let (total_time, total_frames) = (vid.total_time(), vid.total_frames());
        let frames_to_process = (total_time.as_secs_f64() / interval.as_secs_f64()) as u64;
        let mut results = vec![Matching { image: None, video_frame_idx: total_frames as usize, video_time: total_time }]; // lib.rs:186-190

        // Instead of one rayon task per changed frame (lib.rs:213-214) the changed frames are copied into one pinned
        // batch buffer and matched by a single call.
        let mut pending: Vec<(Duration, usize)> = Vec::with_capacity(BATCH);
        let mut pinned: *mut libc::c_void = ptr::null_mut();
        let (mut w, mut h) = (0i32, 0i32);
        let mut done = 0u64;
        let mut flush = |pending: &mut Vec<(Duration, usize)>, pinned: *mut libc::c_void, w: i32, h: i32, results: &mut Vec<Matching<I>>| {
            if pending.is_empty() {
                return;
            }
            let mut out = vec![slideo_b200_frame_result::default(); pending.len()];
            self.ctx.check(unsafe {
                slideo_b200_match_frames_bgr8(self.ctx.0, pinned as *const u8, pending.len() as i32, w, h, 3 * w, (3 * w * h) as usize, out.as_mut_ptr())
            });
            for ((t, idx), r) in pending.drain(..).zip(out) {
                let image = if r.best_slide >= 0 && r.votes >= self.min_votes { Some(self.images[r.best_slide as usize]) } else { None };
                results.push(Matching { video_time: t, video_frame_idx: idx, image });
            }
        };
        for (changed, frame, frame_time, frame_idx) in video_capture::MarkSimilarIter::new(vid) {
            done += 1;
            self.progress_reporter.report(done, frames_to_process, &format!("Processing frames of '{}'...", self.video_path.file_name().unwrap().to_string_lossy()));
            if !changed {
                continue; // lib.rs:207-210
            }
            if pinned.is_null() {
                w = frame.cols();
                h = frame.rows();
                self.ctx.check(unsafe { slideo_b200_host_alloc(&mut pinned, BATCH * (3 * w * h) as usize) });
            }
            unsafe {
                let dst = (pinned as *mut u8).add(pending.len() * (3 * w * h) as usize);
                for y in 0..h {
                    ptr::copy_nonoverlapping(frame.ptr(y).unwrap(), dst.add((y * 3 * w) as usize), (3 * w) as usize);
                }
            }
            pending.push((frame_time, frame_idx));
            if pending.len() == BATCH {
                flush(&mut pending, pinned, w, h, &mut results);
            }
        }
        flush(&mut pending, pinned, w, h, &mut results);
        unsafe { slideo_b200_host_free(pinned) };
        self.progress_reporter.report(frames_to_process, frames_to_process, "Finished!"); // lib.rs:223-227

        // lib.rs:229-244: sort by time, drop consecutive equal images
        results.sort_by_key(|m| m.video_time);
        let mut cleaned: Vec<Matching<I>> = Vec::new();
        for m in results {
            if let Some(last) = cleaned.last() {
                if last.image == m.image {
                    continue;
                }
            }
            cleaned.push(m);
        }
        cleaned
    }
}

/// `video_capture.rs` and `image_utils.rs` of `matching-opencv` are reused unchanged (host-side decode + the
/// changed-frame prefilter): copy crates/matching-opencv/src/{video_capture,image_utils}.rs next to this file.
mod video_capture;
mod image_utils;
