//! Host-side frame source: decodes a video with OpenCV's VideoCapture and yields one frame per sampling interval -- the job of
//! `VideoCaptureIter` in the reference (crates/matching-opencv/src/video_capture.rs:9-57), written for this crate: every frame is
//! grabbed, a frame is retrieved (decoded to BGR) only when `frame_idx % floor(fps * interval) < 1`.
use opencv::{core::Mat, prelude::*, videoio};
use std::{path::Path, time::Duration};

pub struct SampledVideo {
    cap: videoio::VideoCapture,
    fps: f64,
    step: f64,
}

impl SampledVideo {
    pub fn open(path: &Path, interval: Duration) -> Self {
        let cap = videoio::VideoCapture::from_file(&path.to_string_lossy(), videoio::CAP_ANY).unwrap();
        if !cap.is_opened().unwrap() {
            panic!("Could not open video '{}'", path.to_string_lossy());
        }
        let fps = cap.get(videoio::CAP_PROP_FPS).unwrap();
        SampledVideo { cap, fps, step: (fps * interval.as_secs_f64()).floor() }
    }

    /// video_capture.rs:30-37
    pub fn total_frames(&self) -> f64 {
        self.cap.get(videoio::CAP_PROP_FRAME_COUNT).unwrap()
    }
    pub fn total_time(&self) -> Duration {
        Duration::from_secs_f64(self.total_frames() / self.fps)
    }

    /// The next sampled frame as (BGR image, time, frame index), or None at the end of the stream.
    pub fn next_sampled(&mut self) -> Option<(Mat, Duration, usize)> {
        loop {
            let frame_idx = self.cap.get(videoio::CAP_PROP_POS_FRAMES).unwrap();
            if !self.cap.grab().unwrap() {
                return None;
            }
            if self.step <= 0.0 || frame_idx % self.step < 1.0 {
                let mut frame = Mat::default();
                if self.cap.retrieve(&mut frame, 0).unwrap() {
                    return Some((frame, Duration::from_secs_f64(frame_idx / self.fps), frame_idx as usize));
                }
            }
        }
    }
}
