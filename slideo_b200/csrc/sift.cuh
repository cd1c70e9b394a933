// sift.cuh -- K11: SIFT::detectAndCompute with cv::SIFT::create()'s defaults, batched (the north_star's "SIFT DoG pyramid +
// 128-d float descriptor variant"; it sits where feature_extractor.rs:29-46 calls ORB::detectAndCompute).  The algorithm is
// the one restated in oracle/sift_oracle.c (pinned against cv2 4.13.0).
#pragma once
#include <vector>

#include "common.cuh"

namespace slideo {

constexpr int SIFT_LAYERS = 3;               // nOctaveLayers
constexpr int SIFT_GAUSS = SIFT_LAYERS + 3;  // Gaussian images per octave
constexpr int SIFT_MAX_OCT = 14;
constexpr int SIFT_MAX_DIM = 8191;           // doubled image; candidates pack (row, col) into 13 bits each

struct SiftOctave {
    int w, h, pitch;      // floats
    int diag;             // (int)sqrt(w^2 + h^2): descriptor radius clamp
    int tiles_x, tiles_y; // extrema tiles
    int tile_base;
    int pad_;
    size_t off;           // float offset of Gaussian layer 0 inside one image's pyramid
    size_t layer_stride;  // floats between layers
    double ifx, ify;      // INTER_NEAREST source step from the previous octave
};

struct SiftGeo {
    int n_oct, total_tiles, kp_cap, cand_cap;
    size_t img_floats;
    SiftOctave oc[SIFT_MAX_OCT];
};

// Device-resident workspace + launcher for one image geometry (w, h) and batch capacity.
class SiftExtractor {
public:
    SiftExtractor(int w, int h, int batch_cap, int kp_cap_per_image = 0);
    ~SiftExtractor();
    SiftExtractor(const SiftExtractor&) = delete;
    SiftExtractor& operator=(const SiftExtractor&) = delete;

    int width() const { return w_; }
    int height() const { return h_; }
    int batch_cap() const { return batch_cap_; }
    int kp_cap_per_image() const { return geo_.kp_cap; }
    const SiftGeo& geo() const { return geo_; }

    // n device images (channels 1 or 3, 8 bit).  Returns the keypoint count of the batch; one small D2H + stream sync after
    // the duplicate removal.  Throws CapacityError when a fixed capacity is exceeded (never truncates silently).
    int run(const uint8_t* d_src, int n, int stride, size_t frame_stride, int channels, cudaStream_t stream, int* launches);

    // results of the last run: images back to back, inside an image OpenCV's order (KeyPoint_LessThan after removeDuplicated)
    const float* d_kp_f() const { return d_kp_f_; }            // total x 5 {pt.x, pt.y, size, angle, response}
    const int32_t* d_kp_octave() const { return d_kp_oct_; }   // total, packed like cv::KeyPoint::octave
    const float* d_desc() const { return d_desc_; }            // total x 128, integer-valued 0..255
    const int32_t* d_q_frame() const { return d_q_frame_; }    // total: image index within the batch
    const int32_t* d_frame_off() const { return d_frame_off_; }  // n + 1
    const int32_t* d_frame_nkp() const { return d_frame_nkp_; }  // n
    const std::vector<int32_t>& h_frame_off() const { return h_frame_off_; }
    int last_candidates() const { return last_cand_; }           // raw scale-space extrema of the last batch (diagnostics)

    // stage-level parity: Gaussian layer `layer` of octave `octave` of image `img` of the last batch
    const float* d_gauss(int img, int octave, int layer) const {
        return d_pyr_ + (size_t)img * geo_.img_floats + geo_.oc[octave].off + (size_t)layer * geo_.oc[octave].layer_stride;
    }

private:
    int w_, h_, batch_cap_;
    SiftGeo geo_{};
    size_t total_cap_ = 0;
    float* d_pyr_ = nullptr;
    uint8_t* d_gray_ = nullptr;
    uint32_t* d_cand_ = nullptr;
    int32_t* d_cnt_ = nullptr;        // [4][batch]: candidates, raw keypoints, unique keypoints, refined candidates; radius buckets
    void* d_refined_ = nullptr;       // [batch][kp_cap] candidates that survived adjustLocalExtrema
    uint32_t* d_perm_ = nullptr;      // descriptor scheduling order (image << 16 | index), by window radius
    float* d_raw_ = nullptr;          // [batch][kp_cap][6] unsorted keypoints (x, y, size, angle, response, octave bits)
    int32_t* d_order_ = nullptr;      // [batch][kp_cap]
    float* d_uniq_ = nullptr;         // [batch][kp_cap][6] sorted + unique
    float *d_kp_f_ = nullptr, *d_desc_ = nullptr;
    int32_t *d_kp_oct_ = nullptr, *d_q_frame_ = nullptr, *d_frame_off_ = nullptr, *d_frame_nkp_ = nullptr;
    int32_t* h_pinned_ = nullptr;
    std::vector<int32_t> h_frame_off_;
    int last_cand_ = 0;
};

}  // namespace slideo
