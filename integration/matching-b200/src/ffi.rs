//! 1:1 binding of include/slideo_b200.h (ABI version 1).  Keep in sync with slideo_b200/ffi.py.
#![allow(non_camel_case_types)]
use libc::{c_char, c_void, size_t};

pub const SLIDEO_B200_ABI_VERSION: i32 = 1;
pub const SLIDEO_B200_OK: i32 = 0;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct slideo_b200_config {
    pub abi_version: i32,
    pub device: i32,
    pub nfeatures: i32,
    pub scale_factor: f32,
    pub nlevels: i32,
    pub edge_threshold: i32,
    pub patch_size: i32,
    pub fast_threshold: i32,
    pub knn_k: i32,
    pub vote_ratio: f32,
    pub descriptor_kind: i32,
    pub max_batch: i32,
    pub keep_matches: i32,
    /// 1: also the RANSAC gate (lib.rs:284-333); 2: plus the warp + similarity gate (lib.rs:335-389)
    pub geometric_verification: i32,
    pub reserved: [i32; 2],
}

/// `descriptor_kind`: the reference's ORB / Hamming path, or the SIFT-128 / L2 variant (K11 + K10)
pub const SLIDEO_B200_DESC_ORB256: i32 = 0;
pub const SLIDEO_B200_DESC_SIFT128: i32 = 1;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct slideo_b200_frame_result {
    pub best_slide: i32,
    pub votes: i32,
    pub n_keypoints: i32,
}

/// Field for field `KeyedDMatch` (crates/matching-opencv/src/flann.rs:51-59).
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct slideo_b200_match {
    pub query_idx: i32,
    pub train_idx: i32,
    pub source: i32,
    pub distance: f32,
}

pub enum slideo_b200_ctx {}

#[link(name = "slideo_b200")]
extern "C" {
    pub fn slideo_b200_default_config(cfg: *mut slideo_b200_config) -> i32;
    pub fn slideo_b200_create(cfg: *const slideo_b200_config, out: *mut *mut slideo_b200_ctx) -> i32;
    pub fn slideo_b200_destroy(ctx: *mut slideo_b200_ctx) -> i32;
    pub fn slideo_b200_last_error(ctx: *const slideo_b200_ctx) -> *const c_char;
    pub fn slideo_b200_add_page_gray8(ctx: *mut slideo_b200_ctx, px: *const u8, w: i32, h: i32, stride: i32, out_n: *mut i32) -> i32;
    pub fn slideo_b200_finalize_pool(ctx: *mut slideo_b200_ctx) -> i32;
    pub fn slideo_b200_match_frames_bgr8(
        ctx: *mut slideo_b200_ctx, frames: *const u8, n: i32, w: i32, h: i32, stride: i32, frame_stride: size_t,
        out: *mut slideo_b200_frame_result,
    ) -> i32;
    pub fn slideo_b200_get_matches(ctx: *mut slideo_b200_ctx, frame_i: i32, out: *mut slideo_b200_match, cap_rows: i32, out_rows: *mut i32) -> i32;
    /// SIFT::detectAndCompute (cv::SIFT::create() defaults): kp_f n x 5, kp_octave n, desc n x 128 floats; any pointer may be null
    pub fn slideo_b200_extract_sift(
        ctx: *mut slideo_b200_ctx, img: *const u8, w: i32, h: i32, stride: i32, channels: i32, kp_f: *mut f32, kp_octave: *mut i32,
        desc: *mut f32, cap: i32, out_n: *mut i32,
    ) -> i32;
    pub fn slideo_b200_host_alloc(out: *mut *mut c_void, bytes: size_t) -> i32;
    pub fn slideo_b200_host_free(p: *mut c_void) -> i32;
}
