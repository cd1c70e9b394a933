//! 1:1 binding of include/slideo_b200.h (ABI version 1): every function, every struct, every constant.
//! tests/test_abi.py parses this file and the header and compares names, argument counts and struct layouts.
#![allow(non_camel_case_types, dead_code)]
use libc::{c_char, c_void, size_t};

pub const SLIDEO_B200_ABI_VERSION: i32 = 1;

// slideo_b200_status
pub const SLIDEO_B200_OK: i32 = 0;
pub const SLIDEO_B200_E_INVALID_ARG: i32 = -1;
pub const SLIDEO_B200_E_CUDA: i32 = -2;
pub const SLIDEO_B200_E_OOM: i32 = -3;
pub const SLIDEO_B200_E_NOTIMPL: i32 = -4;
pub const SLIDEO_B200_E_STATE: i32 = -5;
pub const SLIDEO_B200_E_CAPACITY: i32 = -6;
pub const SLIDEO_B200_E_INTERNAL: i32 = -7;

// slideo_b200_descriptor_kind: the reference's ORB / Hamming path, or the SIFT-128 / L2 variant (K11 + K10)
pub const SLIDEO_B200_DESC_ORB256: i32 = 0;
pub const SLIDEO_B200_DESC_SIFT128: i32 = 1;

pub const SLIDEO_B200_TOP_SLIDES: usize = 40;
pub const SLIDEO_B200_TOP_RATED: usize = 10;

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
    pub knn_impl: i32,
    pub reserved: [i32; 1],
}

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

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct slideo_b200_timings {
    pub ms_detect: f32,
    pub ms_knn: f32,
    pub ms_vote: f32,
    pub ms_h2d: f32,
    pub knn_pairs: i64,
    pub knn_launches: i64,
    pub kernel_launches: i64,
    pub frames: i64,
    pub ms_total: f32,
    pub ms_verify: f32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct slideo_b200_verify_result {
    pub n_candidates: i32,
    pub n_survivors: i32,
    pub cand_page: [i32; SLIDEO_B200_TOP_SLIDES],
    pub cand_votes: [i32; SLIDEO_B200_TOP_SLIDES],
    pub cand_rating: [i32; SLIDEO_B200_TOP_SLIDES],
    pub survivor_page: [i32; SLIDEO_B200_TOP_RATED],
    pub survivor_rating: [i32; SLIDEO_B200_TOP_RATED],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct slideo_b200_decision {
    /// rated_page[0] or -1 (`Matching.image = None`)
    pub image: i32,
    pub n_rated: i32,
    pub rated_page: [i32; SLIDEO_B200_TOP_RATED],
    pub rated_similarity: [f32; SLIDEO_B200_TOP_RATED],
    pub refined_matrix: [[f64; 4]; SLIDEO_B200_TOP_RATED],
}

pub enum slideo_b200_ctx {}

pub type slideo_b200_progress_fn = Option<unsafe extern "C" fn(processed: u64, total: u64, message: *const c_char, user: *mut c_void)>;

#[link(name = "slideo_b200")]
extern "C" {
    // ---- lifecycle
    pub fn slideo_b200_default_config(cfg: *mut slideo_b200_config) -> i32;
    pub fn slideo_b200_create(cfg: *const slideo_b200_config, out_ctx: *mut *mut slideo_b200_ctx) -> i32;
    pub fn slideo_b200_destroy(ctx: *mut slideo_b200_ctx) -> i32;
    pub fn slideo_b200_last_error(ctx: *const slideo_b200_ctx) -> *const c_char;
    pub fn slideo_b200_version() -> *const c_char;

    // ---- page pool
    pub fn slideo_b200_add_page_gray8(ctx: *mut slideo_b200_ctx, px: *const u8, w: i32, h: i32, stride: i32, out_n_keypoints: *mut i32) -> i32;
    pub fn slideo_b200_add_page_descriptors(ctx: *mut slideo_b200_ctx, desc: *const c_void, n: i32) -> i32;
    pub fn slideo_b200_add_page_features(ctx: *mut slideo_b200_ctx, desc: *const c_void, pt_xy: *const f32, n: i32) -> i32;
    pub fn slideo_b200_finalize_pool(ctx: *mut slideo_b200_ctx) -> i32;
    pub fn slideo_b200_pool_info(ctx: *const slideo_b200_ctx, out_n_descriptors: *mut i32, out_n_pages: *mut i32) -> i32;
    pub fn slideo_b200_pool_export(ctx: *const slideo_b200_ctx, desc: *mut c_void, page_offsets: *mut i32) -> i32;
    pub fn slideo_b200_pool_import(ctx: *mut slideo_b200_ctx, desc: *const c_void, n_desc: i32, page_offsets: *const i32, n_pages: i32) -> i32;
    pub fn slideo_b200_pool_reserve(ctx: *mut slideo_b200_ctx, n_desc: i32, n_pages: i32) -> i32;
    pub fn slideo_b200_pool_device_view(
        ctx: *mut slideo_b200_ctx, d_desc: *mut *mut c_void, desc_bytes: *mut size_t, d_page_offsets: *mut *mut c_void, offsets_bytes: *mut size_t,
    ) -> i32;
    pub fn slideo_b200_pool_commit(ctx: *mut slideo_b200_ctx) -> i32;
    pub fn slideo_b200_pool_points_device_view(ctx: *mut slideo_b200_ctx, d_pt: *mut *mut c_void, bytes: *mut size_t, has_points: *mut i32, received: i32) -> i32;
    pub fn slideo_b200_pool_pages_device_view(
        ctx: *mut slideo_b200_ctx, d_small: *mut *mut c_void, bytes: *mut size_t, page_w: *mut i32, page_h: *mut i32, set_w: i32, set_h: i32,
    ) -> i32;

    // ---- the per-frame hot path
    pub fn slideo_b200_match_frames_bgr8(
        ctx: *mut slideo_b200_ctx, frames: *const u8, n: i32, w: i32, h: i32, stride: i32, frame_stride: size_t, out: *mut slideo_b200_frame_result,
    ) -> i32;
    pub fn slideo_b200_match_frames_bgr8_device(
        ctx: *mut slideo_b200_ctx, d_frames: *const c_void, n: i32, w: i32, h: i32, stride: i32, frame_stride: size_t, out: *mut slideo_b200_frame_result,
    ) -> i32;
    pub fn slideo_b200_submit_frames_bgr8(
        ctx: *mut slideo_b200_ctx, frames: *const u8, n: i32, w: i32, h: i32, stride: i32, frame_stride: size_t, out_ticket: *mut i64,
    ) -> i32;
    pub fn slideo_b200_submit_frames_bgr8_device(
        ctx: *mut slideo_b200_ctx, d_frames: *const c_void, n: i32, w: i32, h: i32, stride: i32, frame_stride: size_t, out_ticket: *mut i64,
    ) -> i32;
    pub fn slideo_b200_collect(ctx: *mut slideo_b200_ctx, ticket: i64, out: *mut slideo_b200_frame_result, cap: i32, out_n: *mut i32) -> i32;
    pub fn slideo_b200_match_descriptors(ctx: *mut slideo_b200_ctx, desc: *const c_void, frame_offsets: *const i32, n: i32, out: *mut slideo_b200_frame_result) -> i32;
    pub fn slideo_b200_get_matches(ctx: *mut slideo_b200_ctx, frame_i: i32, out: *mut slideo_b200_match, cap_rows: i32, out_rows: *mut i32) -> i32;
    pub fn slideo_b200_get_verification(ctx: *mut slideo_b200_ctx, frame0: i32, n: i32, out: *mut slideo_b200_verify_result) -> i32;
    pub fn slideo_b200_get_decisions(ctx: *mut slideo_b200_ctx, frame0: i32, n: i32, out: *mut slideo_b200_decision) -> i32;

    // ---- changed-frame prefilter
    pub fn slideo_b200_mark_changed_bgr8(
        ctx: *mut slideo_b200_ctx, frames: *const u8, n: i32, w: i32, h: i32, stride: i32, frame_stride: size_t, reset: i32, out_changed: *mut u8,
        out_similarity: *mut f32,
    ) -> i32;
    pub fn slideo_b200_mark_changed_bgr8_device(
        ctx: *mut slideo_b200_ctx, d_frames: *const c_void, n: i32, w: i32, h: i32, stride: i32, frame_stride: size_t, reset: i32, out_changed: *mut u8,
        out_similarity: *mut f32,
    ) -> i32;

    // ---- stage-level entry points
    pub fn slideo_b200_extract_orb(
        ctx: *mut slideo_b200_ctx, img: *const u8, w: i32, h: i32, stride: i32, channels: i32, kp_i: *mut i32, kp_f: *mut f32, desc: *mut u8, cap: i32,
        out_n: *mut i32,
    ) -> i32;
    pub fn slideo_b200_extract_sift(
        ctx: *mut slideo_b200_ctx, img: *const u8, w: i32, h: i32, stride: i32, channels: i32, kp_f: *mut f32, kp_octave: *mut i32, desc: *mut f32,
        cap: i32, out_n: *mut i32,
    ) -> i32;
    pub fn slideo_b200_debug_fetch_sift(
        ctx: *mut slideo_b200_ctx, octave: i32, layer: i32, out: *mut f32, cap_bytes: size_t, out_w: *mut i32, out_h: *mut i32, out_n_octaves: *mut i32,
    ) -> i32;
    pub fn slideo_b200_debug_fetch(ctx: *mut slideo_b200_ctx, what: i32, level: i32, out: *mut c_void, cap_bytes: size_t, out_w: *mut i32, out_h: *mut i32) -> i32;
    pub fn slideo_b200_bf_knn_hamming(ctx: *mut slideo_b200_ctx, q: *const u8, nq: i32, t: *const u8, nt: i32, k: i32, idx: *mut i32, dist: *mut i32) -> i32;
    pub fn slideo_b200_bf_knn_hamming_device(ctx: *mut slideo_b200_ctx, d_q: *const c_void, nq: i32, d_t: *const c_void, nt: i32, k: i32, d_keys_out: *mut c_void) -> i32;
    pub fn slideo_b200_bf_knn_l2(ctx: *mut slideo_b200_ctx, q: *const f32, nq: i32, t: *const f32, nt: i32, dim: i32, k: i32, idx: *mut i32, dist: *mut f32) -> i32;
    pub fn slideo_b200_bf_knn_l2_device(
        ctx: *mut slideo_b200_ctx, d_q: *const c_void, nq: i32, d_t: *const c_void, nt: i32, dim: i32, k: i32, d_idx: *mut c_void, d_dist: *mut c_void,
    ) -> i32;

    // ---- utilities
    pub fn slideo_b200_host_alloc(out: *mut *mut c_void, bytes: size_t) -> i32;
    pub fn slideo_b200_host_free(p: *mut c_void) -> i32;
    pub fn slideo_b200_get_timings(ctx: *mut slideo_b200_ctx, out: *mut slideo_b200_timings, reset: i32) -> i32;
    pub fn slideo_b200_set_progress_callback(ctx: *mut slideo_b200_ctx, fn_: slideo_b200_progress_fn, user: *mut c_void) -> i32;
    pub fn slideo_b200_microbench(ctx: *mut slideo_b200_ctx, which: i32, out_per_second: *mut f64) -> i32;
    pub fn slideo_b200_synchronize(ctx: *mut slideo_b200_ctx) -> i32;
}
