//! `matching-b200`: drop-in replacement for `matching-opencv` (crates/matching-opencv/src/lib.rs) behind the
//! `matching` traits (crates/matching/src/lib.rs:7-40).  SOURCE ONLY in this repository (no Rust toolchain in the
//! build image); it is the binding a maintainer adds next to `crates/matching-opencv`.
//!
//! Swap in `crates/app`: `Cargo.toml:25` `matching-opencv` -> `matching-b200`; `main.rs:14,69`
//! `OpenCVImageVideoMatcher::default()` -> `B200ImageVideoMatcher::default()`.
//!
//! What runs where.  Host: PNG read (`imread(path, 0)`, lib.rs:98) and video decode + the 5 s sampling rule
//! (`VideoCaptureIter`, video_capture.rs:15-57; `sampler` below).  GPU, through `libslideo_b200.so`: the changed-frame
//! prefilter (`MarkSimilarIter`, video_capture.rs:60-103 -> `slideo_b200_mark_changed_bgr8`), ORB, the k-NN, the vote and
//! the reference's complete decision tail -- top-40 by votes, RANSAC rating gate (lib.rs:284-333), warp + similarity gate
//! (lib.rs:335-389) -- with `cfg.geometric_verification = 2`; `Matching.image` is `slideo_b200_decision.image`
//! (`-1` -> `None`), exactly the reference's `result.get(0)` (lib.rs:383-389).
mod ffi;
mod sampler;

use ffi::*;
use matching::{ImageVideoMatcher, MatchableImage, Matching, ProgressReporter, VideoMatcher, VideoMatcherTask};
use opencv::{core::Mat, imgcodecs::imread, prelude::*};
use std::{ffi::CStr, os::raw::c_void, path::{Path, PathBuf}, ptr, sync::Arc, time::Duration};

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
    /// frames per library call (the library batches internally; this bounds the pinned staging buffer)
    pub batch: usize,
}
impl Default for B200ImageVideoMatcher {
    fn default() -> Self {
        Self { device: 0, batch: 64 }
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
        cfg.geometric_verification = 2; // the reference's whole gate chain, lib.rs:284-389
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
        Box::new(B200VideoMatcher { ctx: Arc::new(ctx), images: Arc::new(images), batch: self.batch.max(1) })
    }
}

struct B200VideoMatcher<I> {
    ctx: Arc<Ctx>,
    images: Arc<Vec<I>>,
    batch: usize,
}

const SAMPLE_INTERVAL: Duration = Duration::from_secs(5); // lib.rs:145,175

impl<'i, I: MatchableImage + Send + Sync + Copy + Eq + 'i> VideoMatcher<'i, I> for B200VideoMatcher<I> {
    fn match_images_with_video(&self, video_path: &Path, progress_reporter: ProgressReporter) -> Box<dyn VideoMatcherTask<I> + 'i> {
        // lib.rs:145-150: report (0, total_time / 5 s) immediately
        let vid = sampler::SampledVideo::open(video_path, SAMPLE_INTERVAL);
        let frames_to_process = (vid.total_time().as_secs_f64() / SAMPLE_INTERVAL.as_secs_f64()) as u64;
        progress_reporter.report(0, frames_to_process, "");
        Box::new(B200VideoMatcherTask {
            ctx: self.ctx.clone(),
            images: self.images.clone(),
            batch: self.batch,
            video_path: video_path.to_owned(),
            progress_reporter,
        })
    }
}

struct B200VideoMatcherTask<I> {
    ctx: Arc<Ctx>,
    images: Arc<Vec<I>>,
    batch: usize,
    video_path: PathBuf,
    progress_reporter: ProgressReporter,
}

/// Pinned staging buffer for one batch of sampled frames (freed on drop).
struct Pinned(*mut c_void);
impl Drop for Pinned {
    fn drop(&mut self) {
        unsafe { slideo_b200_host_free(self.0 as *mut _) };
    }
}

impl<I: MatchableImage + Send + Sync + Copy + Eq> VideoMatcherTask<I> for B200VideoMatcherTask<I> {
    fn process(&self) -> Vec<Matching<I>> {
        let mut vid = sampler::SampledVideo::open(&self.video_path, SAMPLE_INTERVAL);
        let (total_time, total_frames) = (vid.total_time(), vid.total_frames());
        let frames_to_process = (total_time.as_secs_f64() / SAMPLE_INTERVAL.as_secs_f64()) as u64;
        let name = self.video_path.file_name().unwrap().to_string_lossy().into_owned();
        let mut results = vec![Matching { image: None, video_frame_idx: total_frames as usize, video_time: total_time }]; // lib.rs:186-190

        // Instead of one rayon task per changed frame (lib.rs:213-214): sampled frames are copied into one pinned batch buffer; one
        // call flags the changed ones (the chain state lives in the ctx, so batch boundaries do not matter), the changed frames are
        // compacted in place and matched -- with the whole gate chain -- by a second call.
        let mut meta: Vec<(Duration, usize)> = Vec::with_capacity(self.batch);
        let mut pinned: Option<Pinned> = None;
        let (mut w, mut h) = (0i32, 0i32);
        let mut first_call = true;
        let mut done = 0u64;
        loop {
            let next = vid.next_sampled();
            if let Some((frame, frame_time, frame_idx)) = &next {
                if pinned.is_none() {
                    w = frame.cols();
                    h = frame.rows();
                    let mut p = ptr::null_mut();
                    self.ctx.check(unsafe { slideo_b200_host_alloc(&mut p, self.batch * (3 * w * h) as usize) });
                    pinned = Some(Pinned(p as *mut c_void));
                }
                let frame_bytes = (3 * w * h) as usize;
                unsafe {
                    let dst = (pinned.as_ref().unwrap().0 as *mut u8).add(meta.len() * frame_bytes);
                    for y in 0..h {
                        ptr::copy_nonoverlapping(frame.ptr(y).unwrap(), dst.add((y * 3 * w) as usize), (3 * w) as usize);
                    }
                }
                meta.push((*frame_time, *frame_idx));
            }
            if meta.len() == self.batch || (next.is_none() && !meta.is_empty()) {
                let n = meta.len();
                let frame_bytes = (3 * w * h) as usize;
                let base = pinned.as_ref().unwrap().0 as *mut u8;
                // MarkSimilarIter (video_capture.rs:86-102): changed iff similarity to the previous sampled frame < 0.98
                let mut changed = vec![0u8; n];
                self.ctx.check(unsafe {
                    slideo_b200_mark_changed_bgr8(self.ctx.0, base, n as i32, w, h, 3 * w, frame_bytes, first_call as i32, changed.as_mut_ptr(), ptr::null_mut())
                });
                first_call = false;
                let mut kept: Vec<(Duration, usize)> = Vec::with_capacity(n);
                for i in 0..n {
                    if changed[i] == 0 {
                        continue; // lib.rs:207-210
                    }
                    if kept.len() != i {
                        unsafe { ptr::copy(base.add(i * frame_bytes), base.add(kept.len() * frame_bytes), frame_bytes) };
                    }
                    kept.push(meta[i]);
                }
                if !kept.is_empty() {
                    let m = kept.len();
                    let mut out = vec![slideo_b200_frame_result::default(); m];
                    self.ctx.check(unsafe { slideo_b200_match_frames_bgr8(self.ctx.0, base, m as i32, w, h, 3 * w, frame_bytes, out.as_mut_ptr()) });
                    let mut dec: Vec<slideo_b200_decision> = vec![unsafe { std::mem::zeroed() }; m];
                    self.ctx.check(unsafe { slideo_b200_get_decisions(self.ctx.0, 0, m as i32, dec.as_mut_ptr()) });
                    for ((t, idx), d) in kept.into_iter().zip(dec) {
                        // lib.rs:383-389: the most similar survivor of both gates, or None
                        let image = if d.image >= 0 { Some(self.images[d.image as usize]) } else { None };
                        results.push(Matching { video_time: t, video_frame_idx: idx, image });
                    }
                }
                done += n as u64;
                self.progress_reporter.report(done, frames_to_process, &format!("Processing frames of '{}'...", name));
                meta.clear();
            }
            if next.is_none() {
                break;
            }
        }
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
