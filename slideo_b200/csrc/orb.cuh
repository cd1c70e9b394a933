// orb.cuh -- K1-K7: ORB::detectAndCompute as the reference configures it (feature_extractor.rs:12-46), batched.
#pragma once
#include <vector>

#include "common.cuh"

namespace slideo {

constexpr int ORB_MAX_LEVELS = 12;
constexpr int ORB_MAX_DIM = 4095;  // candidate packing: score << 24 | y << 12 | x

struct OrbConfig {
    int nfeatures = 2000;
    float scale_factor = 1.2f;
    int nlevels = 8;
    int edge_threshold = 62;
    int patch_size = 62;
    int fast_threshold = 20;
};

struct OrbLevelGeom {
    int w, h, pitch;      // pitch in bytes (multiple of 16)
    int quota;            // retainBest(n)
    int cand_cap;         // raw FAST candidates kept per image
    int sel_cap;          // selected keypoints per image (power of two, sort width)
    int tiles_x, tiles_y; // 64x32 tiles
    int tile_base;        // first flat tile index of this level
    int btiles_x, btiles_y, btile_base;   // the blur's own (larger) tiles
    float scale, inv_scale;
    size_t img_off;       // byte offset of the level inside one image's pyramid
    size_t cand_off;      // uint32 offset inside one image's candidate area
    size_t sel_off;       // uint32 offset inside one image's selected area
    size_t tab_off;       // int32 offset of the resize tables of this level (xofs, xc1, yofs, yc1)
};

// Device-resident workspace + launcher for a fixed image geometry (w, h) and batch capacity.
class OrbExtractor {
public:
    OrbExtractor(const OrbConfig& cfg, int w, int h, int batch_cap);
    ~OrbExtractor();
    OrbExtractor(const OrbExtractor&) = delete;
    OrbExtractor& operator=(const OrbExtractor&) = delete;

    int width() const { return w_; }
    int height() const { return h_; }
    int batch_cap() const { return batch_cap_; }
    int nlevels() const { return cfg_.nlevels; }
    const OrbLevelGeom& level(int l) const { return lv_[l]; }

    // Runs the whole extractor on n device images (channels 1 or 3) and returns the total keypoint count of the batch (one
    // stream synchronisation: meant for single images -- pages, stage-level extraction).  Throws CapacityError on overflow.
    int run(const uint8_t* d_src, int n, int stride, size_t frame_stride, int channels, cudaStream_t stream, int* launches,
            int num_sms = 148);
    // The frame path: no host synchronisation at all.  The batch is appended to a query stream whose counters live on the
    // device (KnnStream, common.cuh): descriptors / frame ids / KeyPoint.pt go to the stream position of the batch, the frame
    // table of the stream is extended, capacity problems raise sticky bits in st->flags.
    struct StreamSink {
        uint8_t* desc;        // [q_cap x 32]
        int32_t* q_frame;     // [q_cap] frame (within the stream) of each query
        float2* pt;           // [q_cap] KeyPoint.pt (may be null)
        int32_t* frame_q0;    // [f_cap + 1] first query of each frame; frame_q0[0] = 0 is set by the owner
        int32_t* frame_nkp;   // [f_cap]
        KnnStream* st;
        int q_cap, f_cap;
    };
    void run_stream(const uint8_t* d_src, int n, int stride, size_t frame_stride, int channels, cudaStream_t stream, int* launches,
                    const StreamSink& sink, int num_sms = 148);

    // results of the last run (device pointers; canonical order frame, octave, y, x)
    const uint8_t* d_desc() const { return d_desc_; }          // total x 32
    const int32_t* d_kp_i() const { return d_kp_i_; }          // total x 4 {x_level, y_level, octave, score}
    const float* d_kp_f() const { return d_kp_f_; }            // total x 4 {pt.x, pt.y, size, angle}
    const int32_t* d_q_frame() const { return d_q_frame_; }    // total: image index within the batch
    const int32_t* d_frame_off() const { return d_frame_off_; }  // n + 1
    const int32_t* d_frame_nkp() const { return d_frame_nkp_; }  // n
    const std::vector<int32_t>& h_frame_off() const { return h_frame_off_; }
    size_t kp_cap() const { return kp_cap_; }

    // debug access for stage-level parity tests (image 0 of the last batch)
    const uint8_t* d_pyramid(int img) const { return d_pyr_ + (size_t)img * pyr_img_bytes_; }
    const uint8_t* d_blurred(int img) const { return d_blur_ + (size_t)img * pyr_img_bytes_; }
    const uint32_t* d_candidates(int img) const { return d_cand_ + (size_t)img * cand_img_words_; }
    const int32_t* d_cand_count() const { return d_cand_cnt_; }  // [batch][nlevels]

private:
    void enqueue(const uint8_t* d_src, int n, int stride, size_t frame_stride, int channels, cudaStream_t stream, int* launches,
                 const StreamSink* sink, int num_sms);
    OrbConfig cfg_;
    bool safe_ = false;          // see describe_kernel<SAFE> (orb.cu)
    int w_, h_, batch_cap_;
    std::vector<OrbLevelGeom> lv_;
    int total_tiles_ = 0, total_btiles_ = 0;
    size_t pyr_img_bytes_ = 0, cand_img_words_ = 0, sel_img_words_ = 0, kp_cap_ = 0;
    uint8_t *d_pyr_ = nullptr, *d_blur_ = nullptr, *d_desc_ = nullptr;
    uint32_t *d_cand_ = nullptr, *d_sel_ = nullptr, *d_kp_src_ = nullptr;
    int32_t *d_cand_cnt_ = nullptr, *d_sel_cnt_ = nullptr, *d_kp_off_ = nullptr, *d_frame_off_ = nullptr,
            *d_frame_nkp_ = nullptr, *d_q_frame_ = nullptr, *d_kp_i_ = nullptr, *d_tables_ = nullptr, *d_flags_ = nullptr,
            *d_info_ = nullptr;   // batch header {total, stream query base, stream frame base, flags}
    float* d_kp_f_ = nullptr;
    void* d_geom_ = nullptr;     // OrbLevelGeom[nlevels] on the device
    void* fast_maps_ = nullptr;  // per-level TMA tensor maps of the pyramid (device memory): FAST's boxes
    void* blur_maps_ = nullptr;  // the same tensors with the blur's boxes
    uint32_t *d_tiles_fast_ = nullptr, *d_tiles_blur_ = nullptr;   // flat tile index -> level | tx0 << 4 | ty0 << 16
    int8_t* d_pattern_ = nullptr;  // 512 x 2 int8
    int32_t* h_pinned_ = nullptr;  // [0] total, [1] flags, [2..] frame offsets
    std::vector<int32_t> h_frame_off_;
    int umax_[40];
};

}  // namespace slideo
