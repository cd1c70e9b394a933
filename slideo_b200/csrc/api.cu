// api.cu -- the extern "C" boundary of libslideo_b200.so (include/slideo_b200.h) and the per-ctx host logic:
// device-resident page pool, the streaming frame path (uploads -> K1-K7 -> K8/K9 -> per-frame result, no host synchronisation
// between submit and collect: keypoint counts, K8 ranges and finished frames are all tracked on the device).
// Host-side counterpart of the reference's orchestration in crates/matching-opencv/src/lib.rs:37-64 (page pool),
// lib.rs:249-295 (per-frame path) and flann.rs:64-89 (matcher); all arithmetic runs in the CUDA kernels of
// orb.cu / knn_hamming.cu / knn_l2.cu.  There is no CPU fallback: without a device every entry point fails.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/slideo_b200.h"
#include "common.cuh"
#include "knn_l2.cuh"
#include "orb.cuh"
#include "prefilter.cuh"
#include "sift.cuh"
#include "verify.cuh"

using namespace slideo;

namespace {

template <typename T>
struct DevBuf {  // grow-only device buffer
    T* p = nullptr;
    size_t cap = 0;
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 4 + 64;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            throw CudaError(e, std::string("cudaMalloc of ") + std::to_string(want * sizeof(T)) + " bytes failed: " + cudaGetErrorString(e));
        }
        cap = want;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    ~DevBuf() { release(); }
};

struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    int kind = 0;  // 0 detect, 1 knn, 2 h2d
};

thread_local std::string g_create_error;

}  // namespace

struct slideo_b200_ctx {
    slideo_b200_config cfg{};
    int device = 0, num_sms = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr, knn_stream = nullptr;   // extraction | uploads | K8 of the frame path
    mutable std::string err;

    // ---- page pool --------------------------------------------------------------------------------------
    int desc_bytes = 32;                 // 32 (ORB256) or 512 (SIFT128 as float)
    std::vector<uint8_t> h_pool;         // host staging until finalize
    std::vector<int32_t> page_off{0};    // prefix offsets, size n_pages + 1
    std::vector<float> h_pool_pt;        // KeyPoint.pt of every pooled descriptor (x, y); empty when a page came without points
    bool pool_pts_valid = true;
    bool pts_received = false;           // a reserved pool got its coordinates through pool_points_device_view
    DevBuf<float2> d_pool_pt;
    bool finalized = false;
    DevBuf<uint8_t> d_pool;              // nt x 32 (ORB) / nt x 128 bf16 (SIFT)
    DevBuf<uint8_t> d_pool5;             // ORB: the pool as bit-sliced slabs, what K8 v5 streams (knn5_pool_prepare_launch)
    DevBuf<uint8_t> d_t5;                // bit-sliced copy of a caller-provided pool (stage-level k-NN)
    DevBuf<uint8_t> d_t48;               // 48 B expanded copy of a caller-provided pool (stage-level k-NN, cfg.knn_impl == 4)
    DevBuf<uint32_t> d_partial5;         // K8 v5 partial rows (split tiles): static launches on `stream`
    DevBuf<uint32_t> d_partial5_stream;  // ... and the frame path's launches on `knn_stream` (the two may be in flight together)
    DevBuf<uint8_t> d_mark_frames;       // upload buffer of mark_changed_bgr8 (never the staging ring: tickets may be in flight)
    DevBuf<uint8_t> d_pool_tail;         // SIFT: bf16 norm tails of the pool (knn_l2.cu)
    DevBuf<uint8_t> d_pool_f32;          // SIFT: the pooled descriptors as fp32 (what the NCCL broadcast moves; bf16 operands are derived)
    DevBuf<uint16_t> d_page_of;          // nt
    DevBuf<int32_t> d_page_off;          // n_pages + 1
    int nt = 0, n_pages = 0;
    bool reserved = false;               // pool_reserve called, waiting for commit

    // ---- workspaces -------------------------------------------------------------------------------------
    std::map<std::tuple<int, int, int>, std::unique_ptr<OrbExtractor>> extractors;
    OrbExtractor* last_ext = nullptr;
    std::map<std::tuple<int, int, int>, std::unique_ptr<SiftExtractor>> sift_extractors;   // K11 workspaces (SIFT128 variant)
    SiftExtractor* last_sift = nullptr;
    static constexpr int SIFT_BATCH = 16; // upper bound of the frames per K11 batch (265 MB of fp32 scale space per 1080p frame)
    static constexpr int N_STAGING = 4;  // frame staging buffers (uploads run this many batches ahead)
    DevBuf<uint8_t> d_frames[N_STAGING];
    DevBuf<uint8_t> d_img;               // single-image upload (pages, extract_orb)
    DevBuf<uint32_t> d_scratch, d_partial, d_keys;
    DevBuf<int32_t> d_votes, d_results, d_q_frame, d_idx, d_dist;
    DevBuf<uint8_t> d_q;                 // uploaded query descriptors
    DevBuf<uint8_t> d_t;                 // uploaded train descriptors (stage-level knn)
    L2Workspace l2ws;
    DevBuf<uint8_t> d_l2_pool, d_l2_tail; // bf16 operands of a caller-provided pool (stage-level L2 k-NN)
    int32_t* h_results = nullptr;        // pinned, max_batch x 3 x 2
    size_t h_results_cap = 0;
    cudaEvent_t ev_copy[N_STAGING] = {}, ev_free[N_STAGING] = {};
    cudaEvent_t ev_detect = nullptr;     // extraction of everything appended to the query stream so far is done

    // ---- kept matches of the last match call -------------------------------------------------------------
    std::vector<uint32_t> kept_keys;     // total_q x k
    std::vector<float> kept_l2;          // SIFT: distances
    std::vector<int32_t> kept_l2_idx;
    std::vector<int32_t> kept_frame_off{0};

    // ---- timings ----------------------------------------------------------------------------------------
    slideo_b200_timings tm{};
    std::vector<EventPair> ev_pending, ev_free_list;

    ~slideo_b200_ctx() {
        cudaSetDevice(device);
        if (stream) cudaStreamSynchronize(stream);
        if (copy_stream) cudaStreamSynchronize(copy_stream);
        if (knn_stream) cudaStreamSynchronize(knn_stream);
        extractors.clear();
        sift_extractors.clear();
        engine_destroy();
        for (auto& e : ev_pending) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
        for (auto& e : ev_free_list) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
        for (int i = 0; i < N_STAGING; ++i) {
            if (ev_copy[i]) cudaEventDestroy(ev_copy[i]);
            if (ev_free[i]) cudaEventDestroy(ev_free[i]);
        }
        if (h_results) cudaFreeHost(h_results);
        if (stream) cudaStreamDestroy(stream);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (knn_stream) cudaStreamDestroy(knn_stream);
        if (ev_detect) cudaEventDestroy(ev_detect);
    }

    // -------------------------------------------------------------------------------------------------------
    EventPair begin_timing(int kind, cudaStream_t s) {
        EventPair e;
        if (!ev_free_list.empty()) {
            e = ev_free_list.back();
            ev_free_list.pop_back();
        } else {
            SLIDEO_CUDA(cudaEventCreate(&e.a));
            SLIDEO_CUDA(cudaEventCreate(&e.b));
        }
        e.kind = kind;
        SLIDEO_CUDA(cudaEventRecord(e.a, s));
        return e;
    }
    void end_timing(EventPair e, cudaStream_t s) {
        SLIDEO_CUDA(cudaEventRecord(e.b, s));
        ev_pending.push_back(e);
    }
    void collect_timings() {  // after a synchronize
        for (auto& e : ev_pending) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
                if (e.kind == 0) tm.ms_detect += ms;
                else if (e.kind == 1) tm.ms_knn += ms;
                else if (e.kind == 2) tm.ms_h2d += ms;
                else if (e.kind == 4) tm.ms_total += ms;
                else if (e.kind == 5) tm.ms_verify += ms;
                else tm.ms_vote += ms;
            }
            ev_free_list.push_back(e);
        }
        ev_pending.clear();
    }

    OrbExtractor& extractor(int w, int h, int cap) {
        auto key = std::make_tuple(w, h, cap);
        auto it = extractors.find(key);
        if (it == extractors.end()) {
            OrbConfig oc;
            oc.nfeatures = cfg.nfeatures;
            oc.scale_factor = cfg.scale_factor;
            oc.nlevels = cfg.nlevels;
            oc.edge_threshold = cfg.edge_threshold;
            oc.patch_size = cfg.patch_size;
            oc.fast_threshold = cfg.fast_threshold;
            if (extractors.size() >= 8) {  // bounded cache of geometries
                SLIDEO_CUDA(cudaStreamSynchronize(stream));
                extractors.clear();
                last_ext = nullptr;
            }
            it = extractors.emplace(key, std::make_unique<OrbExtractor>(oc, w, h, cap)).first;
        }
        last_ext = it->second.get();
        return *it->second;
    }

    SiftExtractor& sift_extractor(int w, int h, int cap) {
        auto key = std::make_tuple(w, h, cap);
        auto it = sift_extractors.find(key);
        if (it == sift_extractors.end()) {
            if (sift_extractors.size() >= 4) {  // bounded cache of geometries
                SLIDEO_CUDA(cudaStreamSynchronize(stream));
                sift_extractors.clear();
                last_sift = nullptr;
            }
            it = sift_extractors.emplace(key, std::make_unique<SiftExtractor>(w, h, cap)).first;
        }
        last_sift = it->second.get();
        return *it->second;
    }

    // K10 + vote + argmax on nq device-resident fp32 descriptors of nb frames (SIFT128 variant of lib.rs:266-295)
    void knn_vote_l2(const float* dq, int nq, const int32_t* d_qf, int nb, const int32_t* d_nkp, const std::vector<int32_t>& fo) {
        d_votes.reserve((size_t)nb * std::max(n_pages, 1));
        d_results.reserve((size_t)nb * 3);
        d_idx.reserve((size_t)std::max(nq, 1) * cfg.knn_k);
        d_dist.reserve((size_t)std::max(nq, 1) * cfg.knn_k);
        SLIDEO_CUDA(cudaMemsetAsync(d_votes.p, 0, (size_t)nb * std::max(n_pages, 1) * 4, stream));
        if (nq > 0) {
            EventPair tk = begin_timing(1, stream);
            int nl = 0;
            l2_knn_launch(l2ws, dq, nq, d_pool.p, d_pool_tail.p, nt, cfg.knn_k, d_idx.p, (float*)d_dist.p, num_sms, stream, &nl);
            l2_vote_launch(d_idx.p, (const float*)d_dist.p, nq, cfg.knn_k, d_qf, d_page_of.p, d_votes.p, n_pages, cfg.vote_ratio, stream);
            end_timing(tk, stream);
            tm.knn_launches += nl;
            tm.kernel_launches += nl + 1;
            tm.knn_pairs += (int64_t)nq * nt;
        }
        vote_argmax_launch(d_votes.p, nb, n_pages, d_nkp, d_results.p, stream);
        tm.kernel_launches += 1;
        if (cfg.keep_matches) {
            const size_t base = kept_l2.size();
            kept_l2.resize(base + (size_t)nq * cfg.knn_k);
            kept_l2_idx.resize(base + (size_t)nq * cfg.knn_k);
            if (nq > 0) {
                SLIDEO_CUDA(cudaMemcpyAsync(kept_l2.data() + base, d_dist.p, (size_t)nq * cfg.knn_k * 4, cudaMemcpyDeviceToHost, stream));
                SLIDEO_CUDA(cudaMemcpyAsync(kept_l2_idx.data() + base, d_idx.p, (size_t)nq * cfg.knn_k * 4, cudaMemcpyDeviceToHost, stream));
            }
            const int32_t qb = kept_frame_off.back();
            for (size_t i = 1; i < fo.size(); ++i) kept_frame_off.push_back(qb + fo[i]);
            SLIDEO_CUDA(cudaStreamSynchronize(stream));
        }
    }

    // the per-frame path of the SIFT128 variant: K11 on batches of frames -> K10 against the pool -> vote -> argmax
    void match_frames_sift(const uint8_t* frames, bool on_device, int n, int w, int h, int stride, size_t frame_stride) {
        const int B = std::min(cfg.max_batch, SIFT_BATCH);
        SiftExtractor& ex = sift_extractor(w, h, B);
        const size_t img_bytes = (size_t)3 * w * h;
        const int n_batches = cdiv(n, B);
        if (!on_device)
            for (int i = 0; i < 2; ++i) d_frames[i].reserve((size_t)std::min(B, n) * img_bytes);
        auto issue_copy = [&](int b) {
            const int buf = b & 1, f0 = b * B, nb = std::min(B, n - f0);
            SLIDEO_CUDA(cudaStreamWaitEvent(copy_stream, ev_free[buf], 0));
            EventPair t = begin_timing(2, copy_stream);
            upload_images(d_frames[buf].p, frames + (size_t)f0 * frame_stride, nb, 3 * w, h, stride, frame_stride, copy_stream);
            end_timing(t, copy_stream);
            SLIDEO_CUDA(cudaEventRecord(ev_copy[buf], copy_stream));
        };
        if (!on_device) issue_copy(0);
        for (int b = 0; b < n_batches; ++b) {
            const int f0 = b * B, nb = std::min(B, n - f0);
            const uint8_t* src = frames + (size_t)f0 * frame_stride;
            int st = stride;
            size_t fst = frame_stride;
            if (!on_device) {
                if (b + 1 < n_batches) issue_copy(b + 1);
                SLIDEO_CUDA(cudaStreamWaitEvent(stream, ev_copy[b & 1], 0));
                src = d_frames[b & 1].p;
                st = 3 * w;
                fst = img_bytes;
            }
            static const bool trace = getenv("SLIDEO_TRACE") != nullptr;   // developer knob, read once per process
            auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
            const double h0 = trace ? now() : 0;
            EventPair t = begin_timing(0, stream);
            int nl = 0;
            const int total = ex.run(src, nb, st, fst, 3, stream, &nl);
            end_timing(t, stream);
            if (!on_device) SLIDEO_CUDA(cudaEventRecord(ev_free[b & 1], stream));
            tm.kernel_launches += nl;
            tm.frames += nb;
            const double h1 = trace ? now() : 0;
            knn_vote_l2(ex.d_desc(), total, ex.d_q_frame(), nb, ex.d_frame_nkp(), ex.h_frame_off());
            const double h2 = trace ? now() : 0;
            SLIDEO_CUDA(cudaMemcpyAsync(h_results + (size_t)f0 * 3, d_results.p, (size_t)nb * 3 * 4, cudaMemcpyDeviceToHost, stream));
            if (trace) fprintf(stderr, "[trace] batch %d: K11 run (host) %.2f ms, knn_vote_l2 enqueue (host) %.2f ms, d2h enqueue %.2f ms, nq %d\n", b, h1 - h0, h2 - h1, now() - h2, total);
        }
    }

    void require_orb() const {
        if (cfg.descriptor_kind != SLIDEO_B200_DESC_ORB256) throw StateError("this entry point needs descriptor_kind ORB256");
    }
    void require_pool() const {
        if (!finalized) throw StateError("page pool not finalized (call slideo_b200_finalize_pool / pool_import / pool_commit first)");
    }

    // uploads `n_img` images (host) into dst as tightly packed rows of row_bytes
    void upload_images(uint8_t* dst, const uint8_t* src, int n_img, int row_bytes, int h, int stride, size_t frame_stride,
                       cudaStream_t s) {
        if (stride == row_bytes && frame_stride == (size_t)row_bytes * h) {
            SLIDEO_CUDA(cudaMemcpyAsync(dst, src, (size_t)n_img * row_bytes * h, cudaMemcpyHostToDevice, s));
        } else if (stride == row_bytes) {
            SLIDEO_CUDA(cudaMemcpy2DAsync(dst, (size_t)row_bytes * h, src, frame_stride, (size_t)row_bytes * h, n_img,
                                          cudaMemcpyHostToDevice, s));
        } else {
            for (int i = 0; i < n_img; ++i)
                SLIDEO_CUDA(cudaMemcpy2DAsync(dst + (size_t)i * row_bytes * h, row_bytes, src + (size_t)i * frame_stride, stride,
                                              row_bytes, h, cudaMemcpyHostToDevice, s));
        }
    }

    // K8 (+ fused K9) on nq device-resident query descriptors against the pool; fills d_results for n_frames
    void knn_vote_hamming(const uint8_t* dq, int nq, const int32_t* d_qf, int n_frames, const int32_t* d_frame_nkp,
                          bool want_keys) {
        d_votes.reserve((size_t)n_frames * std::max(n_pages, 1));
        d_results.reserve((size_t)n_frames * 3);
        SLIDEO_CUDA(cudaMemsetAsync(d_votes.p, 0, (size_t)n_frames * std::max(n_pages, 1) * 4, stream));
        if (nq > 0) {
            Knn5Plan plan = knn5_plan(nq, nt, cfg.knn_k, num_sms);
            if (plan.partial_bytes) d_partial5.reserve(plan.partial_bytes / 4);
            if (want_keys) d_keys.reserve((size_t)nq * cfg.knn_k);
            VoteArgs va{d_qf, d_page_of.p, d_votes.p, n_pages, cfg.vote_ratio};
            EventPair t = begin_timing(1, stream);
            int nl = 0;
            knn5_launch(plan, dq, d_pool5.p, want_keys ? d_keys.p : nullptr, d_partial5.p, &va, stream, &nl);
            end_timing(t, stream);
            tm.knn_launches += 1;
            tm.kernel_launches += nl;
            tm.knn_pairs += (int64_t)nq * nt;
        }
        vote_argmax_launch(d_votes.p, n_frames, n_pages, d_frame_nkp, d_results.p, stream);
        tm.kernel_launches += 1;
    }

    void ensure_host_results(size_t n) {
        if (n <= h_results_cap) return;
        if (h_results) cudaFreeHost(h_results);
        h_results = nullptr;
        SLIDEO_CUDA(cudaMallocHost(&h_results, n * 3 * sizeof(int32_t)));
        h_results_cap = n;
    }

    void keep_batch_keys(int nq, const std::vector<int32_t>& frame_off_local, cudaStream_t st = nullptr) {
        if (!st) st = stream;
        const size_t base = kept_keys.size();
        kept_keys.resize(base + (size_t)nq * cfg.knn_k);
        if (nq > 0)
            SLIDEO_CUDA(cudaMemcpyAsync(kept_keys.data() + base, d_keys.p, (size_t)nq * cfg.knn_k * 4, cudaMemcpyDeviceToHost, st));
        const int32_t q0 = kept_frame_off.back();
        for (size_t i = 1; i < frame_off_local.size(); ++i) kept_frame_off.push_back(q0 + frame_off_local[i]);
        SLIDEO_CUDA(cudaStreamSynchronize(st));
    }

    // ---- the per-frame hot path as a device-driven stream ------------------------------------------------------------------
    // Detection (K1-K7, `stream`) appends the descriptors of every batch to a query stream; the keypoint counts stay on the device
    // (KnnStream).  After each batch a one-thread plan kernel cuts the next K8 range (whole waves of 148 x 128 queries; the
    // remainder rides with the next launch), K8 v5 + the finalize kernels run on `knn_stream` and write finished frames straight
    // into a host-mapped result ring plus a progress word.  Nothing on the host ever waits for a count: submit() only enqueues,
    // collect() waits for launch events until the progress word covers the ticket.  The stream is continuous across submit
    // calls, so the upload + detection of call i+1 overlap K8 of call i (the reference's streaming loop, lib.rs:196-220).
    // Query-stream buffers come in two sets ("epochs" of EPOCH_FRAMES frames) that alternate.
    static constexpr int EPOCH_FRAMES = 2048;
    static constexpr int RING = 1 << 16;          // frames of the host result ring
    static constexpr int DYN_SLOTS = 256;         // K8 launch descriptors in flight
    struct Epoch {
        DevBuf<uint8_t> desc;                     // [q_cap x 32]
        DevBuf<int32_t> q_frame, frame_q0, frame_nkp, votes;
        DevBuf<float2> pt;                        // KeyPoint.pt of each query (geometric verification)
        DevBuf<uint32_t> keys;                    // k-NN rows (keep_matches / geometric verification)
        KnnStream* d_state = nullptr;
        cudaEvent_t ev_done = nullptr;            // K8 side of the epoch finished
        long long seq_base = 0;                   // global number of the epoch's frame 0
        int frames = 0, q_cap = 0;
        bool used = false;
    };
    Epoch ep[2];
    int cur_ep = 0;
    bool ep_open = false;
    int ep_w = 0, ep_h = 0;
    long long seq_submitted = 0;
    KnnDyn* d_dyn = nullptr;
    cudaEvent_t ev_plan[DYN_SLOTS] = {}, ev_knn[DYN_SLOTS] = {};
    long long launch_seq = 0, launch_synced = 0;  // K8 launch groups enqueued / known to be complete
    long long batch_seq = 0;                      // staging-ring position (uploads)
    unsigned long long pairs_seen = 0;            // device pair counters already folded into tm.knn_pairs
    int32_t* h_ring = nullptr;                    // pinned + mapped: RING x 3 results, then the progress word and the flags word
    volatile long long* h_progress = nullptr;
    volatile int* h_flags = nullptr;
    struct Ticket { long long id, seq0, seq1; };
    std::deque<Ticket> tickets;
    long long next_ticket = 1;
    bool span_open = false, span_closed = false;  // ms_total: first submit after a timings reset .. last collect
    cudaEvent_t ev_span0 = nullptr, ev_span1 = nullptr;

    // verification state of the last match call
    DevBuf<int32_t> d_v_cand_page, d_v_cand_votes, d_v_n_cand, d_v_rating;
    DevBuf<uint8_t> d_v_corr;
    DevBuf<uint8_t> d_v_pairs;                  // RANSAC sample pairs of every small correspondence count (built once, verify.cuh)
    bool v_pairs_built = false;
    DevBuf<VerifyRecord> d_v_out;
    std::vector<VerifyRecord> verify_results;   // one per frame of the last match_frames_* call
    // photometric stage (K14, cfg.geometric_verification == 2)
    std::vector<slideo_b200_decision> decisions;
    std::vector<uint8_t> h_page_small;          // gray small image of every page, back to back (page p at h_page_small_off[p])
    // page geometry: one class per distinct page size (the reference warps the frame to each slide's own size, lib.rs:339-348)
    struct PageClass { int w = 0, h = 0; AreaTables area; };
    std::vector<std::unique_ptr<PageClass>> page_classes;
    std::vector<int32_t> h_page_class;               // [n_pages]
    std::vector<unsigned long long> h_page_small_off; // [n_pages]
    DevBuf<PageGeom> d_page_geom;
    DevBuf<int32_t> d_page_class;
    DevBuf<unsigned long long> d_page_small_off;
    int max_small_w = 0, max_small_h = 0;
    bool page_geom_ok = true;                   // every page came with its image (add_page_gray8) or the images were replicated
    bool page_small_ready = false;              // d_page_small holds the small image of every page (built here or replicated)
    bool page_small_received = false;           // a reserved pool got its page images through pool_pages_device_view
    DevBuf<uint8_t> d_page_small, d_all_frames;
    DevBuf<int32_t> d_v_best_it, d_v_surv_cand;
    DevBuf<double> d_v_refined;
    DevBuf<unsigned long long> d_v_sumsq;
    const uint8_t* photo_frames = nullptr;      // device frames of the current group (frame i at + i * photo_frame_stride)
    int photo_w = 0, photo_h = 0, photo_row_stride = 0;
    size_t photo_frame_stride = 0;

    // ---- changed-frame prefilter (K13) --------------------------------------------------------------------
    AreaTables area;
    DevBuf<uint8_t> d_small;                    // [max_batch + 1] small images: slot 0 = last frame of the previous batch / call
    DevBuf<uint8_t> d_page_small_tmp;           // small image of the page being added
    DevBuf<unsigned long long> d_sumsq;
    bool have_prev_small = false;

    slideo_b200_progress_fn progress_fn = nullptr;   // optional (processed, total, message) callback, matching/src/progress.rs:3-17
    void* progress_user = nullptr;

    bool want_keys() const { return cfg.keep_matches != 0 || cfg.geometric_verification != 0; }
    size_t kp_per_frame_cap(int w, int h) { return extractor(w, h, cfg.max_batch).kp_cap() / (size_t)cfg.max_batch; }

    void engine_init() {
        SLIDEO_CUDA(cudaMalloc(&d_dyn, sizeof(KnnDyn) * DYN_SLOTS));
        SLIDEO_CUDA(cudaMemset(d_dyn, 0, sizeof(KnnDyn) * DYN_SLOTS));
        for (int i = 0; i < DYN_SLOTS; ++i) {
            SLIDEO_CUDA(cudaEventCreateWithFlags(&ev_plan[i], cudaEventDisableTiming));
            SLIDEO_CUDA(cudaEventCreateWithFlags(&ev_knn[i], cudaEventDisableTiming));
        }
        void* hp = nullptr;
        SLIDEO_CUDA(cudaHostAlloc(&hp, (size_t)RING * 12 + 64, cudaHostAllocMapped));
        std::memset(hp, 0, (size_t)RING * 12 + 64);
        h_ring = (int32_t*)hp;
        h_progress = (volatile long long*)((uint8_t*)hp + (size_t)RING * 12);
        h_flags = (volatile int*)((uint8_t*)hp + (size_t)RING * 12 + 16);
        for (int i = 0; i < 2; ++i) {
            SLIDEO_CUDA(cudaMalloc(&ep[i].d_state, sizeof(KnnStream)));
            SLIDEO_CUDA(cudaMemset(ep[i].d_state, 0, sizeof(KnnStream)));
            SLIDEO_CUDA(cudaEventCreateWithFlags(&ep[i].ev_done, cudaEventDisableTiming));
        }
        SLIDEO_CUDA(cudaEventCreate(&ev_span0));
        SLIDEO_CUDA(cudaEventCreate(&ev_span1));
        d_partial5_stream.reserve(knn5_dyn_partial_bytes(num_sms, KNN_MAX_K) / 4);
    }
    void engine_destroy() {
        if (d_dyn) cudaFree(d_dyn);
        for (int i = 0; i < DYN_SLOTS; ++i) {
            if (ev_plan[i]) cudaEventDestroy(ev_plan[i]);
            if (ev_knn[i]) cudaEventDestroy(ev_knn[i]);
        }
        if (h_ring) cudaFreeHost(h_ring);
        for (int i = 0; i < 2; ++i) {
            if (ep[i].d_state) cudaFree(ep[i].d_state);
            if (ep[i].ev_done) cudaEventDestroy(ep[i].ev_done);
        }
        if (ev_span0) cudaEventDestroy(ev_span0);
        if (ev_span1) cudaEventDestroy(ev_span1);
    }

    // one K8 launch group over whatever part of the current epoch's stream is ready (flush: all of it)
    void launch_group(bool flush) {
        const int slot = (int)(launch_seq % DYN_SLOTS);
        if (launch_seq >= DYN_SLOTS) {   // the descriptor slot is reused: its previous launch must be complete
            SLIDEO_CUDA(cudaEventSynchronize(ev_knn[slot]));
            launch_synced = std::max(launch_synced, launch_seq - DYN_SLOTS + 1);
        }
        Epoch& e = ep[cur_ep];
        knn5_plan_launch(e.d_state, d_dyn + slot, nt, num_sms, flush ? 1 : 0, stream);
        SLIDEO_CUDA(cudaEventRecord(ev_plan[slot], stream));
        SLIDEO_CUDA(cudaStreamWaitEvent(knn_stream, ev_plan[slot], 0));
        VoteArgs va{e.q_frame.p, d_page_of.p, e.votes.p, n_pages, cfg.vote_ratio};
        EventPair t = begin_timing(1, knn_stream);
        int nl = 0;
        knn5_launch_dyn(d_dyn + slot, e.q_cap, nt, cfg.knn_k, num_sms, e.desc.p, d_pool5.p, want_keys() ? e.keys.p : nullptr, d_partial5_stream.p,
                        &va, knn_stream, &nl);
        end_timing(t, knn_stream);
        stream_finalize_launch(e.d_state, d_dyn + slot, e.frame_q0.p, e.votes.p, n_pages, e.frame_nkp.p, h_ring, RING - 1, e.seq_base,
                               nullptr, h_progress, h_flags, knn_stream);
        SLIDEO_CUDA(cudaEventRecord(ev_knn[slot], knn_stream));
        ++launch_seq;
        tm.knn_launches += 1;
        tm.kernel_launches += nl + 3;
    }

    void close_epoch() {
        if (!ep_open) return;
        launch_group(true);
        SLIDEO_CUDA(cudaEventRecord(ep[cur_ep].ev_done, knn_stream));
        ep_open = false;
    }

    void open_epoch(int w, int h) {
        if (ep[cur_ep].used) cur_ep ^= 1;
        Epoch& e = ep[cur_ep];
        const size_t per_frame = kp_per_frame_cap(w, h);
        const size_t q_cap = (size_t)EPOCH_FRAMES * per_frame;
        const int np = std::max(n_pages, 1);
        const bool grow = q_cap * 32 + 64 > e.desc.cap || (size_t)EPOCH_FRAMES * np > e.votes.cap ||
                          (want_keys() && q_cap * cfg.knn_k > e.keys.cap) || (cfg.geometric_verification && q_cap + 64 > e.pt.cap);
        if (e.used) {
            if (grow) SLIDEO_CUDA(cudaEventSynchronize(e.ev_done));                  // the buffers are about to be reallocated
            else SLIDEO_CUDA(cudaStreamWaitEvent(stream, e.ev_done, 0));             // K8 of two epochs ago is done with this set
        }
        e.desc.reserve(q_cap * 32 + 64);
        e.q_frame.reserve(q_cap + 64);
        e.frame_q0.reserve((size_t)EPOCH_FRAMES + 2);
        e.frame_nkp.reserve((size_t)EPOCH_FRAMES + 1);
        e.votes.reserve((size_t)EPOCH_FRAMES * np);
        if (cfg.geometric_verification) e.pt.reserve(q_cap + 64);
        if (want_keys()) e.keys.reserve(q_cap * cfg.knn_k);
        SLIDEO_CUDA(cudaMemsetAsync(e.d_state, 0, 24, stream));                      // counters + flags; `pairs` keeps accumulating
        SLIDEO_CUDA(cudaMemsetAsync(e.frame_q0.p, 0, 4, stream));
        SLIDEO_CUDA(cudaMemsetAsync(e.votes.p, 0, (size_t)EPOCH_FRAMES * np * 4, stream));
        e.seq_base = seq_submitted;
        e.frames = 0;
        e.q_cap = (int)std::min<size_t>(q_cap, 0x7FFFFFFF);
        e.used = true;
        ep_open = true;
        ep_w = w;
        ep_h = h;
    }

    // enqueues uploads (host frames), detection and K8 launches of n frames; returns at once.  keep_dst != nullptr: the frames are
    // uploaded into that device buffer (and stay there) instead of the staging ring.
    void enqueue_frames(const uint8_t* frames, bool on_device, int n, int w, int h, int stride, size_t frame_stride, uint8_t* keep_dst) {
        const int B = cfg.max_batch;
        const size_t img_bytes = (size_t)3 * w * h;
        OrbExtractor& ex = extractor(w, h, B);
        if (!on_device && !keep_dst)
            for (int i = 0; i < N_STAGING; ++i) {
                if ((size_t)std::min(B, n) * img_bytes > d_frames[i].cap) {   // growing a staging buffer: nothing may still use it
                    SLIDEO_CUDA(cudaStreamSynchronize(copy_stream));
                    SLIDEO_CUDA(cudaStreamSynchronize(stream));
                    d_frames[i].reserve((size_t)B * img_bytes);
                }
            }
        if (!span_open) {
            SLIDEO_CUDA(cudaEventRecord(ev_span0, stream));
            span_open = true;
        }
        const int n_batches = cdiv(n, B);
        const long long b0 = batch_seq;
        auto issue_copy = [&](int b) {
            const long long bs = b0 + b;
            const int buf = (int)(bs % N_STAGING), f0 = b * B, nb = std::min(B, n - f0);
            uint8_t* dst = keep_dst ? keep_dst + (size_t)f0 * img_bytes : d_frames[buf].p;
            if (!keep_dst && bs >= N_STAGING) SLIDEO_CUDA(cudaStreamWaitEvent(copy_stream, ev_free[buf], 0));   // detection of batch bs - N_STAGING is done
            EventPair t = begin_timing(2, copy_stream);
            upload_images(dst, frames + (size_t)f0 * frame_stride, nb, 3 * w, h, stride, frame_stride, copy_stream);
            end_timing(t, copy_stream);
            SLIDEO_CUDA(cudaEventRecord(ev_copy[buf], copy_stream));
        };
        // uploads run N_STAGING batches ahead on the copy stream
        if (!on_device)
            for (int b = 0; b < std::min(N_STAGING - 1, n_batches); ++b) issue_copy(b);
        for (int b = 0; b < n_batches; ++b) {
            const int f0 = b * B, nb = std::min(B, n - f0);
            if (!ep_open || ep_w != w || ep_h != h || ep[cur_ep].frames + nb > EPOCH_FRAMES) {
                close_epoch();
                open_epoch(w, h);
            }
            Epoch& e = ep[cur_ep];
            const uint8_t* src = frames + (size_t)f0 * frame_stride;
            int st = stride;
            size_t fst = frame_stride;
            const int buf = (int)((b0 + b) % N_STAGING);
            if (!on_device) {
                if (b + N_STAGING - 1 < n_batches) issue_copy(b + N_STAGING - 1);
                SLIDEO_CUDA(cudaStreamWaitEvent(stream, ev_copy[buf], 0));
                src = keep_dst ? keep_dst + (size_t)f0 * img_bytes : d_frames[buf].p;
                st = 3 * w;
                fst = img_bytes;
            }
            OrbExtractor::StreamSink sink{e.desc.p, e.q_frame.p, cfg.geometric_verification ? e.pt.p : nullptr, e.frame_q0.p, e.frame_nkp.p,
                                          e.d_state, e.q_cap, EPOCH_FRAMES};
            EventPair t = begin_timing(0, stream);
            int nl = 0;
            ex.run_stream(src, nb, st, fst, 3, stream, &nl, sink, num_sms);
            end_timing(t, stream);
            if (!on_device && !keep_dst) SLIDEO_CUDA(cudaEventRecord(ev_free[buf], stream));
            tm.kernel_launches += nl;
            tm.frames += nb;
            e.frames += nb;
            seq_submitted += nb;
            launch_group(false);
        }
        if (!on_device) batch_seq += n_batches;
    }

    // waits until the results of frames [.., seq_end) are in the host ring; flushes the stream when its tail sits in a partial wave
    void wait_progress(long long seq_end, bool check_flags, long long seq_begin = -1) {
        bool flushed = false;
        while (*h_progress < seq_end) {
            if (launch_synced < launch_seq) {
                SLIDEO_CUDA(cudaEventSynchronize(ev_knn[launch_synced % DYN_SLOTS]));
                ++launch_synced;
                if (progress_fn && seq_begin >= 0) {
                    const long long done = std::min(std::max(*h_progress - seq_begin, 0LL), seq_end - seq_begin);
                    progress_fn((uint64_t)done, (uint64_t)(seq_end - seq_begin), "Processing frames...", progress_user);
                }
                continue;
            }
            if (!ep_open || flushed) {   // everything has run and the frames are still missing: a capacity bit dropped them
                if (*h_flags) break;
                throw std::runtime_error("frame stream lost results (progress word behind the submitted frames)");
            }
            launch_group(true);
            flushed = true;
        }
        if (progress_fn && seq_begin >= 0 && *h_progress >= seq_end)   // always one final report, also when nothing had to be waited for
            progress_fn((uint64_t)(seq_end - seq_begin), (uint64_t)(seq_end - seq_begin), "Processing frames...", progress_user);
        if (!check_flags) return;
        const int fl = *h_flags;
        if (fl) {
            *h_flags = 0;
            if (fl & 1) throw CapacityError("FAST candidate capacity exceeded on at least one frame");
            if (fl & 2) throw CapacityError("selected-keypoint capacity exceeded on at least one frame");
            throw CapacityError("query stream capacity exceeded");
        }
    }

    long long submit(const uint8_t* frames, bool on_device, int n, int w, int h, int stride, size_t frame_stride) {
        if (seq_submitted + n - (tickets.empty() ? seq_submitted : tickets.front().seq0) > RING)
            throw StateError("too many frames in flight: collect earlier tickets first");
        Ticket t{next_ticket++, seq_submitted, 0};
        enqueue_frames(frames, on_device, n, w, h, stride, frame_stride, nullptr);
        t.seq1 = seq_submitted;
        tickets.push_back(t);
        return t.id;
    }

    int collect(long long id, slideo_b200_frame_result* out, int cap) {
        size_t i = 0;
        while (i < tickets.size() && tickets[i].id != id) ++i;
        if (i == tickets.size()) throw ArgError("unknown or already collected ticket");
        const Ticket t = tickets[i];
        const int n = (int)(t.seq1 - t.seq0);
        if (out && cap < n) throw ArgError("cap smaller than the number of frames of the ticket");
        wait_progress(t.seq1, true, t.seq0);
        if (out)
            for (int f = 0; f < n; ++f) {
                const int32_t* r = h_ring + (size_t)((t.seq0 + f) & (RING - 1)) * 3;
                out[f].best_slide = r[0];
                out[f].votes = r[1];
                out[f].n_keypoints = r[2];
            }
        tickets.erase(tickets.begin() + (long)i);
        SLIDEO_CUDA(cudaEventRecord(ev_span1, knn_stream));
        span_closed = true;
        return n;
    }

    // drains the frame stream: every ticket stays collectable, all device work of the path is complete afterwards
    void drain() {
        if (ep_open && seq_submitted > 0) wait_progress(seq_submitted, false);
        SLIDEO_CUDA(cudaStreamSynchronize(knn_stream));
        SLIDEO_CUDA(cudaStreamSynchronize(stream));
        SLIDEO_CUDA(cudaStreamSynchronize(copy_stream));
        launch_synced = launch_seq;
    }

    // the synchronous variant with the reference's decision tail (cfg.geometric_verification) and / or kept k-NN rows: groups of
    // at most EPOCH_FRAMES frames, each one an epoch of its own that is verified before the next starts
    void match_verified(const uint8_t* frames, bool on_device, int n, int w, int h, int stride, size_t frame_stride,
                        slideo_b200_frame_result* out) {
        const size_t img_bytes = (size_t)3 * w * h;
        const bool keep_frames = cfg.geometric_verification >= 2 && !on_device;   // the warp gate reads the frames again at the end
        for (int s0 = 0; s0 < n; s0 += EPOCH_FRAMES) {
            const int ns = std::min(EPOCH_FRAMES, n - s0);
            close_epoch();
            if (keep_frames) {
                if ((size_t)ns * img_bytes > d_all_frames.cap) drain();
                d_all_frames.reserve((size_t)ns * img_bytes);
            }
            const long long seq0 = seq_submitted;
            enqueue_frames(frames + (size_t)s0 * frame_stride, on_device, ns, w, h, stride, frame_stride, keep_frames ? d_all_frames.p : nullptr);
            const int set = cur_ep;
            close_epoch();
            wait_progress(seq_submitted, true, seq0);
            SLIDEO_CUDA(cudaStreamSynchronize(knn_stream));
            for (int f = 0; f < ns; ++f) {
                const int32_t* r = h_ring + (size_t)((seq0 + f) & (RING - 1)) * 3;
                out[s0 + f].best_slide = r[0];
                out[s0 + f].votes = r[1];
                out[s0 + f].n_keypoints = r[2];
            }
            Epoch& e = ep[set];
            std::vector<int32_t> fo((size_t)ns + 1);
            SLIDEO_CUDA(cudaMemcpy(fo.data(), e.frame_q0.p, ((size_t)ns + 1) * 4, cudaMemcpyDeviceToHost));
            if (cfg.geometric_verification) {
                if (keep_frames) {
                    photo_frames = d_all_frames.p;
                    photo_w = w; photo_h = h; photo_row_stride = 3 * w; photo_frame_stride = img_bytes;
                } else {
                    photo_frames = on_device ? frames + (size_t)s0 * frame_stride : nullptr;
                    photo_w = w; photo_h = h; photo_row_stride = stride; photo_frame_stride = frame_stride;
                }
                verify_epoch(e, ns, fo.back());
            }
            if (cfg.keep_matches) {
                const size_t base = kept_keys.size();
                kept_keys.resize(base + (size_t)fo.back() * cfg.knn_k);
                if (fo.back() > 0)
                    SLIDEO_CUDA(cudaMemcpy(kept_keys.data() + base, e.keys.p, (size_t)fo.back() * cfg.knn_k * 4, cudaMemcpyDeviceToHost));
                const int32_t q0 = kept_frame_off.back();
                for (size_t i = 1; i < fo.size(); ++i) kept_frame_off.push_back(q0 + fo[i]);
            }
        }
        SLIDEO_CUDA(cudaEventRecord(ev_span1, knn_stream));
        span_closed = true;
    }

    // K12 on the frames of a finished epoch (lib.rs:284-333); appends one record per frame to verify_results
    void verify_epoch(Epoch& e, int n_frames, int total_q) {
        if (!pool_pts_valid) throw StateError("geometric verification needs the keypoint coordinates of every page (add_page_gray8 / add_page_features)");
        const size_t F = (size_t)n_frames;
        d_v_cand_page.reserve(F * VERIFY_TOP_SLIDES);
        d_v_cand_votes.reserve(F * VERIFY_TOP_SLIDES);
        d_v_rating.reserve(F * VERIFY_TOP_SLIDES);
        d_v_n_cand.reserve(F + 1);
        d_v_out.reserve(F + 1);
        d_v_corr.reserve(verify_corr_bytes((long long)total_q * cfg.knn_k));
        VerifyArgs a;
        a.n_frames = n_frames; a.n_pages = n_pages; a.k = cfg.knn_k; a.ratio = cfg.vote_ratio;
        a.d_votes = e.votes.p; a.d_keys = e.keys.p; a.d_frame_q0 = e.frame_q0.p; a.d_page_of = d_page_of.p;
        a.d_frame_pt = e.pt.p; a.d_pool_pt = d_pool_pt.p;
        a.d_cand_page = d_v_cand_page.p; a.d_cand_votes = d_v_cand_votes.p; a.d_n_cand = d_v_n_cand.p; a.d_rating = d_v_rating.p;
        a.d_corr = d_v_corr.p; a.d_out = d_v_out.p;
        if (!v_pairs_built) {
            d_v_pairs.reserve(verify_pairs_bytes());
            verify_pairs_build(d_v_pairs.p, knn_stream);
            v_pairs_built = true;
            tm.kernel_launches += 1;
        }
        a.d_pairs = reinterpret_cast<const uint2*>(d_v_pairs.p);
        const bool photo = cfg.geometric_verification >= 2;
        if (photo) {
            d_v_best_it.reserve(F * VERIFY_TOP_SLIDES);
            d_v_surv_cand.reserve(F * VERIFY_TOP_RATED);
        }
        a.d_best_it = photo ? d_v_best_it.p : nullptr;
        a.d_survivor_cand = photo ? d_v_surv_cand.p : nullptr;
        EventPair t = begin_timing(5, knn_stream);
        int nl = 0;
        verify_launch(a, knn_stream, &nl);
        end_timing(t, knn_stream);
        tm.kernel_launches += nl;
        const size_t base = verify_results.size();
        verify_results.resize(base + F);
        SLIDEO_CUDA(cudaMemcpyAsync(verify_results.data() + base, d_v_out.p, F * sizeof(VerifyRecord), cudaMemcpyDeviceToHost, knn_stream));
        if (photo) photometric_epoch(a, n_frames, base);
        SLIDEO_CUDA(cudaStreamSynchronize(knn_stream));
    }

    // K14 on the survivors of a finished epoch (lib.rs:335-389); appends one decision per frame
    void photometric_epoch(const VerifyArgs& va, int n_frames, size_t base) {
        if (!page_geom_ok || !page_small_ready)
            throw StateError("the warp + similarity gate needs every page as an image (add_page_gray8, or replicated page images)");
        if (!photo_frames) throw StateError("frames are not resident for the warp + similarity gate");
        const size_t F = (size_t)n_frames;
        d_v_refined.reserve(F * VERIFY_TOP_RATED * 4);
        d_v_sumsq.reserve(F * VERIFY_TOP_RATED);
        PhotoArgs p;
        p.n_frames = n_frames; p.k = cfg.knn_k;
        p.d_records = d_v_out.p; p.d_survivor_cand = d_v_surv_cand.p; p.d_best_it = d_v_best_it.p; p.d_cand_votes = va.d_cand_votes;
        p.d_corr = va.d_corr; p.d_frame_q0 = va.d_frame_q0; p.d_frame_pt = va.d_frame_pt; p.d_pool_pt = va.d_pool_pt;
        p.d_frames = photo_frames; p.frame_w = photo_w; p.frame_h = photo_h; p.frame_stride_row = photo_row_stride;
        p.frame_stride = photo_frame_stride;
        p.d_geom = d_page_geom.p; p.d_page_class = d_page_class.p; p.d_page_small_off = d_page_small_off.p;
        p.max_small_w = max_small_w; p.max_small_h = max_small_h;
        p.d_page_small = d_page_small.p; p.d_refined = d_v_refined.p; p.d_sumsq = d_v_sumsq.p;
        EventPair t = begin_timing(5, knn_stream);
        int nl = 0;
        photometric_launch(p, knn_stream, &nl);
        end_timing(t, knn_stream);
        tm.kernel_launches += nl;
        std::vector<unsigned long long> ss(F * VERIFY_TOP_RATED);
        std::vector<double> ref(F * VERIFY_TOP_RATED * 4);
        SLIDEO_CUDA(cudaMemcpyAsync(ss.data(), d_v_sumsq.p, ss.size() * 8, cudaMemcpyDeviceToHost, knn_stream));
        SLIDEO_CUDA(cudaMemcpyAsync(ref.data(), d_v_refined.p, ref.size() * 8, cudaMemcpyDeviceToHost, knn_stream));
        SLIDEO_CUDA(cudaStreamSynchronize(knn_stream));
        decisions.resize(base + F);
        for (size_t f = 0; f < F; ++f) {
            const VerifyRecord& r = verify_results[base + f];
            slideo_b200_decision& d = decisions[base + f];
            std::memset(&d, 0, sizeof d);
            struct Rated { int page; float sim; };
            std::vector<Rated> rated;
            for (int j = 0; j < r.n_survivors; ++j) {
                const double error_l2 = sqrt((double)ss[f * VERIFY_TOP_RATED + j]);
                const AreaTables& ar = page_classes[(size_t)h_page_class[(size_t)r.survivor_page[j]]]->area;
                const float max_error = sqrtf((255.0f * 255.0f * 3.0f) * (float)(ar.dw * ar.dh));   // image_utils.rs:24-25
                const float sim = 1.0f - (float)error_l2 / max_error;               // image_utils.rs:26
                rated.push_back({r.survivor_page[j], sim});
                for (int c = 0; c < 4; ++c) d.refined_matrix[j][c] = ref[(f * VERIFY_TOP_RATED + j) * 4 + c];
            }
            std::stable_sort(rated.begin(), rated.end(), [](const Rated& a, const Rated& b) { return a.sim > b.sim; });   // lib.rs:370
            d.image = -1;
            for (int j = 0; j < VERIFY_TOP_RATED; ++j) { d.rated_page[j] = -1; d.rated_similarity[j] = 0.f; }
            for (const Rated& x : rated)
                if (x.sim > 0.5f) {                                                   // lib.rs:381
                    d.rated_page[d.n_rated] = x.page;
                    d.rated_similarity[d.n_rated] = x.sim;
                    ++d.n_rated;
                }
            if (d.n_rated) d.image = d.rated_page[0];
        }
    }

    // the size class of a page (created on first use)
    int page_class_index(int w, int h) {
        for (size_t c = 0; c < page_classes.size(); ++c)
            if (page_classes[c]->w == w && page_classes[c]->h == h) return (int)c;
        page_classes.emplace_back(new PageClass());
        page_classes.back()->w = w;
        page_classes.back()->h = h;
        page_classes.back()->area.build(w, h);
        return (int)page_classes.size() - 1;
    }
    // device tables of the page geometry (classes, class and small-image offset of every page)
    void upload_page_geometry() {
        std::vector<PageGeom> g(page_classes.size());
        max_small_w = max_small_h = 0;
        for (size_t c = 0; c < page_classes.size(); ++c) {
            const PageClass& pc = *page_classes[c];
            g[c].page_w = pc.w; g[c].page_h = pc.h; g[c].small_w = pc.area.dw; g[c].small_h = pc.area.dh;
            g[c].d_xoff = pc.area.d_xoff; g[c].d_xsi = pc.area.d_xsi; g[c].d_xa = pc.area.d_xa;
            g[c].d_yoff = pc.area.d_yoff; g[c].d_ysi = pc.area.d_ysi; g[c].d_ya = pc.area.d_ya;
            max_small_w = std::max(max_small_w, pc.area.dw);
            max_small_h = std::max(max_small_h, pc.area.dh);
        }
        d_page_geom.reserve(std::max<size_t>(g.size(), 1));
        d_page_class.reserve(std::max<size_t>(h_page_class.size(), 1));
        d_page_small_off.reserve(std::max<size_t>(h_page_small_off.size(), 1));
        if (!g.empty()) SLIDEO_CUDA(cudaMemcpy(d_page_geom.p, g.data(), g.size() * sizeof(PageGeom), cudaMemcpyHostToDevice));
        if (!h_page_class.empty()) {
            SLIDEO_CUDA(cudaMemcpy(d_page_class.p, h_page_class.data(), h_page_class.size() * 4, cudaMemcpyHostToDevice));
            SLIDEO_CUDA(cudaMemcpy(d_page_small_off.p, h_page_small_off.data(), h_page_small_off.size() * 8, cudaMemcpyHostToDevice));
        }
    }

    void reset_kept() {
        verify_results.clear();
        decisions.clear();
        kept_keys.clear();
        kept_l2.clear();
        kept_l2_idx.clear();
        kept_frame_off.assign(1, 0);
    }

    void upload_pool_orb(const uint8_t* desc) {
        d_pool.reserve((size_t)std::max(nt, 1) * 32 + 64);
        d_page_of.reserve((size_t)std::max(nt, 1));
        d_page_off.reserve((size_t)n_pages + 1);
        if (nt > 0 && desc) SLIDEO_CUDA(cudaMemcpyAsync(d_pool.p, desc, (size_t)nt * 32, cudaMemcpyHostToDevice, stream));
        build_page_of();
    }
    void build_page_of() {
        std::vector<uint16_t> po((size_t)nt);
        for (int p = 0; p < n_pages; ++p)
            for (int i = page_off[p]; i < page_off[p + 1]; ++i) po[(size_t)i] = (uint16_t)p;
        d_page_of.reserve((size_t)std::max(nt, 1));
        d_page_off.reserve((size_t)n_pages + 1);
        if (nt > 0) SLIDEO_CUDA(cudaMemcpyAsync(d_page_of.p, po.data(), (size_t)nt * 2, cudaMemcpyHostToDevice, stream));
        SLIDEO_CUDA(cudaMemcpyAsync(d_page_off.p, page_off.data(), ((size_t)n_pages + 1) * 4, cudaMemcpyHostToDevice, stream));
        if (cfg.geometric_verification >= 2 && page_geom_ok && !h_page_small.empty() && (int)h_page_class.size() == n_pages) {
            d_page_small.reserve(h_page_small.size());
            SLIDEO_CUDA(cudaMemcpyAsync(d_page_small.p, h_page_small.data(), h_page_small.size(), cudaMemcpyHostToDevice, stream));
            upload_page_geometry();
            page_small_ready = true;
        }
        pool_pts_valid = pool_pts_valid && h_pool_pt.size() == (size_t)nt * 2;
        if (pool_pts_valid && !pts_received) {
            d_pool_pt.reserve((size_t)std::max(nt, 1));
            if (nt > 0) SLIDEO_CUDA(cudaMemcpyAsync(d_pool_pt.p, h_pool_pt.data(), (size_t)nt * 8, cudaMemcpyHostToDevice, stream));
        }
        if (cfg.descriptor_kind == SLIDEO_B200_DESC_ORB256) {
            d_pool5.reserve(knn5_pool_bytes(std::max(nt, 1)) + 64);
            knn5_pool_prepare_launch(d_pool.p, nt, d_pool5.p, stream);
            tm.kernel_launches += 1;
        }
        SLIDEO_CUDA(cudaStreamSynchronize(stream));
    }
    void check_pool_limits(int64_t n_desc, int64_t pages) const {
        if (n_desc > KNN_MAX_POOL) throw CapacityError("pool exceeds 8,388,608 descriptors");
        if (pages > 65535) throw CapacityError("pool exceeds 65535 pages");
    }
};

namespace {

template <typename F>
int32_t guarded(const slideo_b200_ctx* ctx, F&& f) {
    std::string* err = ctx ? &ctx->err : &g_create_error;
    try {
        if (ctx) {
            cudaError_t e = cudaSetDevice(ctx->device);
            if (e != cudaSuccess) throw CudaError(e, std::string("cudaSetDevice failed: ") + cudaGetErrorString(e));
        }
        err->clear();
        f();
        return SLIDEO_B200_OK;
    } catch (const ArgError& e) {
        *err = e.what();
        return SLIDEO_B200_E_INVALID_ARG;
    } catch (const StateError& e) {
        *err = e.what();
        return SLIDEO_B200_E_STATE;
    } catch (const CapacityError& e) {
        *err = e.what();
        return SLIDEO_B200_E_CAPACITY;
    } catch (const NotImplError& e) {
        *err = e.what();
        return SLIDEO_B200_E_NOTIMPL;
    } catch (const CudaError& e) {
        *err = e.what();
        cudaGetLastError();
        return e.code == cudaErrorMemoryAllocation ? SLIDEO_B200_E_OOM : SLIDEO_B200_E_CUDA;
    } catch (const std::bad_alloc&) {
        *err = "host allocation failed";
        return SLIDEO_B200_E_OOM;
    } catch (const std::exception& e) {
        *err = e.what();
        return SLIDEO_B200_E_INTERNAL;
    } catch (...) {
        *err = "unknown exception";
        return SLIDEO_B200_E_INTERNAL;
    }
}

#define REQUIRE_CTX(ctx) \
    if (!(ctx)) return SLIDEO_B200_E_INVALID_ARG

void arg(bool ok, const char* what) {
    if (!ok) throw ArgError(what);
}

}  // namespace

// ===================================================================================================================
extern "C" {

int32_t slideo_b200_default_config(slideo_b200_config* cfg) {
    if (!cfg) return SLIDEO_B200_E_INVALID_ARG;
    std::memset(cfg, 0, sizeof *cfg);
    cfg->abi_version = SLIDEO_B200_ABI_VERSION;
    cfg->device = 0;
    cfg->nfeatures = 2000;       // feature_extractor.rs:14
    cfg->scale_factor = 1.2f;    // :15
    cfg->nlevels = 8;            // :16
    cfg->edge_threshold = 62;    // :17
    cfg->patch_size = 62;        // :22
    cfg->fast_threshold = 20;    // :23
    cfg->knn_k = 30;             // lib.rs:266
    cfg->vote_ratio = 1.05f;     // lib.rs:275
    cfg->descriptor_kind = SLIDEO_B200_DESC_ORB256;
    cfg->max_batch = 32;
    cfg->keep_matches = 0;
    return SLIDEO_B200_OK;
}

int32_t slideo_b200_create(const slideo_b200_config* cfg, slideo_b200_ctx** out_ctx) {
    if (out_ctx) *out_ctx = nullptr;
    return guarded(nullptr, [&] {
        arg(cfg && out_ctx, "cfg and out_ctx must not be NULL");
        arg(cfg->abi_version == SLIDEO_B200_ABI_VERSION, "abi_version mismatch");
        arg(cfg->knn_k >= 1 && cfg->knn_k <= KNN_MAX_K, "knn_k must be in 1..32");
        arg(cfg->max_batch >= 1 && cfg->max_batch <= 1024, "max_batch must be in 1..1024");
        arg(cfg->nfeatures >= 1 && cfg->nfeatures <= 4096, "nfeatures must be in 1..4096");
        arg(cfg->nlevels >= 1 && cfg->nlevels <= ORB_MAX_LEVELS, "nlevels out of range");
        arg(cfg->scale_factor > 1.0f && cfg->scale_factor <= 2.0f, "scale_factor must be in (1, 2]");
        arg(cfg->edge_threshold >= 3 && cfg->edge_threshold < 1024, "edge_threshold out of range");
        arg(cfg->patch_size >= 2 && cfg->patch_size / 2 <= 38, "patch_size out of range");
        arg(cfg->fast_threshold >= 1 && cfg->fast_threshold <= 126, "fast_threshold must be in 1..126");
        arg(cfg->vote_ratio >= 1.0f && cfg->vote_ratio < 16.f, "vote_ratio out of range");
        arg(cfg->descriptor_kind == SLIDEO_B200_DESC_ORB256 || cfg->descriptor_kind == SLIDEO_B200_DESC_SIFT128,
            "unknown descriptor_kind");
        arg(cfg->knn_impl == 0 || cfg->knn_impl == 4 || cfg->knn_impl == 5, "knn_impl must be 0, 4 or 5");
        if (cfg->descriptor_kind == SLIDEO_B200_DESC_SIFT128 && cfg->geometric_verification)
            throw NotImplError("geometric verification is implemented for the ORB256 path only");
        int n_dev = 0;
        cudaError_t e = cudaGetDeviceCount(&n_dev);
        if (e != cudaSuccess || n_dev == 0)
            throw CudaError(e == cudaSuccess ? cudaErrorNoDevice : e,
                            std::string("no CUDA device (libslideo_b200 has no CPU fallback): ") +
                                cudaGetErrorString(e == cudaSuccess ? cudaErrorNoDevice : e));
        arg(cfg->device >= 0 && cfg->device < n_dev, "device ordinal out of range");
        SLIDEO_CUDA(cudaSetDevice(cfg->device));
        cudaDeviceProp prop;
        SLIDEO_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
        if (prop.major != 10)
            throw CudaError(cudaErrorInvalidDevice, std::string("device is sm_") + std::to_string(prop.major) +
                                                        std::to_string(prop.minor) + "; this library is built for sm_100a only");
        std::unique_ptr<slideo_b200_ctx> c(new slideo_b200_ctx());
        c->cfg = *cfg;
        c->device = cfg->device;
        c->num_sms = prop.multiProcessorCount;
        c->desc_bytes = cfg->descriptor_kind == SLIDEO_B200_DESC_ORB256 ? 32 : 512;
        SLIDEO_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        SLIDEO_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        {
            // K8 / K10 run on the stream of greatest priority: their persistent CTAs take whole SMs, and when a wave's CTAs queue
            // behind the many small CTAs of the image-scan kernels the launch is stretched for no gain in total work (measured:
            // K8 launches 9.42 -> 8.72 ms inside the step, +0.4 % frames/s)
            int pr_least = 0, pr_greatest = 0;
            SLIDEO_CUDA(cudaDeviceGetStreamPriorityRange(&pr_least, &pr_greatest));
            SLIDEO_CUDA(cudaStreamCreateWithPriority(&c->knn_stream, cudaStreamNonBlocking, pr_greatest));
        }
        SLIDEO_CUDA(cudaEventCreateWithFlags(&c->ev_detect, cudaEventDisableTiming));
        for (int i = 0; i < slideo_b200_ctx::N_STAGING; ++i) {
            SLIDEO_CUDA(cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming));
            SLIDEO_CUDA(cudaEventCreateWithFlags(&c->ev_free[i], cudaEventDisableTiming));
        }
        if (cfg->descriptor_kind == SLIDEO_B200_DESC_ORB256) c->engine_init();
        *out_ctx = c.release();
    });
}

int32_t slideo_b200_destroy(slideo_b200_ctx* ctx) {
    if (!ctx) return SLIDEO_B200_OK;
    try {
        delete ctx;
    } catch (...) {
        return SLIDEO_B200_E_INTERNAL;
    }
    return SLIDEO_B200_OK;
}

const char* slideo_b200_last_error(const slideo_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

const char* slideo_b200_version(void) { return "slideo_b200 0.1.0 (sm_100a, abi 1)"; }

// ---- page pool ------------------------------------------------------------------------------------------------
int32_t slideo_b200_add_page_gray8(slideo_b200_ctx* ctx, const uint8_t* px, int32_t w, int32_t h, int32_t stride,
                                   int32_t* out_n_keypoints) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        arg(px != nullptr, "px must not be NULL");
        arg(stride >= w, "stride < w");
        if (ctx->finalized || ctx->reserved) throw StateError("pool already finalized");
        ctx->check_pool_limits((int64_t)ctx->page_off.back(), (int64_t)ctx->page_off.size());
        if (ctx->cfg.descriptor_kind == SLIDEO_B200_DESC_SIFT128) {   // K11 on the page; descriptors pooled as n x 128 floats
            SiftExtractor& sx = ctx->sift_extractor(w, h, 1);
            ctx->d_img.reserve((size_t)w * h);
            ctx->upload_images(ctx->d_img.p, px, 1, w, h, stride, (size_t)stride * h, ctx->stream);
            int nl = 0;
            const int total = sx.run(ctx->d_img.p, 1, w, (size_t)w * h, 1, ctx->stream, &nl);
            ctx->tm.kernel_launches += nl;
            const size_t base = ctx->h_pool.size();
            ctx->h_pool.resize(base + (size_t)total * 512);
            std::vector<float> kpf((size_t)total * 5);
            if (total > 0) {
                SLIDEO_CUDA(cudaMemcpyAsync(ctx->h_pool.data() + base, sx.d_desc(), (size_t)total * 512, cudaMemcpyDeviceToHost, ctx->stream));
                SLIDEO_CUDA(cudaMemcpyAsync(kpf.data(), sx.d_kp_f(), (size_t)total * 20, cudaMemcpyDeviceToHost, ctx->stream));
            }
            SLIDEO_CUDA(cudaStreamSynchronize(ctx->stream));
            for (int i = 0; i < total; ++i) {
                ctx->h_pool_pt.push_back(kpf[(size_t)i * 5]);
                ctx->h_pool_pt.push_back(kpf[(size_t)i * 5 + 1]);
            }
            ctx->page_geom_ok = false;
            ctx->page_off.push_back(ctx->page_off.back() + total);
            ctx->check_pool_limits((int64_t)ctx->page_off.back(), (int64_t)ctx->page_off.size() - 1);
            if (out_n_keypoints) *out_n_keypoints = total;
            return;
        }
        OrbExtractor& ex = ctx->extractor(w, h, 1);
        ctx->d_img.reserve((size_t)w * h);
        ctx->upload_images(ctx->d_img.p, px, 1, w, h, stride, (size_t)stride * h, ctx->stream);
        int nl = 0;
        const int total = ex.run(ctx->d_img.p, 1, w, (size_t)w * h, 1, ctx->stream, &nl);
        ctx->tm.kernel_launches += nl;
        const size_t base = ctx->h_pool.size();
        ctx->h_pool.resize(base + (size_t)total * 32);
        std::vector<float> kpf((size_t)total * 4);
        if (total > 0) {
            SLIDEO_CUDA(cudaMemcpyAsync(ctx->h_pool.data() + base, ex.d_desc(), (size_t)total * 32, cudaMemcpyDeviceToHost, ctx->stream));
            SLIDEO_CUDA(cudaMemcpyAsync(kpf.data(), ex.d_kp_f(), (size_t)total * 16, cudaMemcpyDeviceToHost, ctx->stream));
        }
        SLIDEO_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->cfg.geometric_verification >= 2) {   // slide.small_img (lib.rs:105): to_small_image of the gray page replicated to BGR
            const int pc = ctx->page_class_index(w, h);
            const AreaTables& ar = ctx->page_classes[(size_t)pc]->area;
            const size_t sb = (size_t)ar.dw * ar.dh;
            ctx->d_page_small_tmp.reserve(sb);
            area_small_launch(ar, ctx->d_img.p, 1, w, (size_t)w * h, ctx->d_page_small_tmp.p, ctx->stream, 1);
            const size_t b0 = ctx->h_page_small.size();
            ctx->h_page_small.resize(b0 + sb);
            SLIDEO_CUDA(cudaMemcpyAsync(ctx->h_page_small.data() + b0, ctx->d_page_small_tmp.p, sb, cudaMemcpyDeviceToHost, ctx->stream));
            SLIDEO_CUDA(cudaStreamSynchronize(ctx->stream));
            ctx->tm.kernel_launches += 1;
            ctx->h_page_class.push_back(pc);
            ctx->h_page_small_off.push_back((unsigned long long)b0);
        }
        for (int i = 0; i < total; ++i) {   // KeyPoint.pt of the slide keypoints (lib.rs:299)
            ctx->h_pool_pt.push_back(kpf[(size_t)i * 4]);
            ctx->h_pool_pt.push_back(kpf[(size_t)i * 4 + 1]);
        }
        ctx->page_off.push_back(ctx->page_off.back() + total);
        ctx->check_pool_limits((int64_t)ctx->page_off.back(), (int64_t)ctx->page_off.size() - 1);
        if (out_n_keypoints) *out_n_keypoints = total;
    });
}

int32_t slideo_b200_add_page_descriptors(slideo_b200_ctx* ctx, const void* desc, int32_t n) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        arg(n >= 0, "n < 0");
        arg(desc != nullptr || n == 0, "desc must not be NULL");
        if (ctx->finalized || ctx->reserved) throw StateError("pool already finalized");
        ctx->check_pool_limits((int64_t)ctx->page_off.back() + n, (int64_t)ctx->page_off.size());
        const size_t base = ctx->h_pool.size();
        ctx->h_pool.resize(base + (size_t)n * ctx->desc_bytes);
        if (n) std::memcpy(ctx->h_pool.data() + base, desc, (size_t)n * ctx->desc_bytes);
        ctx->page_off.push_back(ctx->page_off.back() + n);
        if (n) ctx->pool_pts_valid = false;   // no keypoint coordinates for this page: geometric verification unavailable
        ctx->page_geom_ok = false;
    });
}

int32_t slideo_b200_add_page_features(slideo_b200_ctx* ctx, const void* desc, const float* pt_xy, int32_t n) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        ctx->require_orb();
        arg(n >= 0, "n < 0");
        arg((desc != nullptr && pt_xy != nullptr) || n == 0, "desc / pt_xy must not be NULL");
        if (ctx->finalized || ctx->reserved) throw StateError("pool already finalized");
        ctx->check_pool_limits((int64_t)ctx->page_off.back() + n, (int64_t)ctx->page_off.size());
        const size_t base = ctx->h_pool.size();
        ctx->h_pool.resize(base + (size_t)n * 32);
        if (n) std::memcpy(ctx->h_pool.data() + base, desc, (size_t)n * 32);
        ctx->h_pool_pt.insert(ctx->h_pool_pt.end(), pt_xy, pt_xy + (size_t)n * 2);
        ctx->page_geom_ok = false;   // no page image: the warp + similarity gate is unavailable
        ctx->page_off.push_back(ctx->page_off.back() + n);
    });
}

static void finalize_from_host(slideo_b200_ctx* ctx) {
    ctx->nt = ctx->page_off.back();
    ctx->n_pages = (int)ctx->page_off.size() - 1;
    if (ctx->cfg.descriptor_kind == SLIDEO_B200_DESC_ORB256) {
        ctx->upload_pool_orb(ctx->h_pool.data());
    } else {
        ctx->d_pool_f32.reserve((size_t)std::max(ctx->nt, 1) * 512);
        if (ctx->nt > 0)
            SLIDEO_CUDA(cudaMemcpyAsync(ctx->d_pool_f32.p, ctx->h_pool.data(), (size_t)ctx->nt * 512, cudaMemcpyHostToDevice, ctx->stream));
        ctx->d_pool.reserve(l2_main_bytes(ctx->nt));
        ctx->d_pool_tail.reserve(l2_tail_bytes(ctx->nt));
        l2_prepare_launch((const float*)ctx->d_pool_f32.p, ctx->nt, false, ctx->d_pool.p, ctx->d_pool_tail.p, ctx->stream);
        ctx->tm.kernel_launches += 1;
        ctx->build_page_of();
    }
    ctx->finalized = true;
}

int32_t slideo_b200_finalize_pool(slideo_b200_ctx* ctx) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        if (ctx->finalized) throw StateError("pool already finalized");
        if (ctx->reserved) throw StateError("pool_reserve pending: call pool_commit");
        finalize_from_host(ctx);
    });
}

int32_t slideo_b200_pool_info(const slideo_b200_ctx* ctx, int32_t* out_n_descriptors, int32_t* out_n_pages) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        if (out_n_descriptors) *out_n_descriptors = ctx->page_off.back();
        if (out_n_pages) *out_n_pages = (int32_t)ctx->page_off.size() - 1;
    });
}

int32_t slideo_b200_pool_export(const slideo_b200_ctx* ctx, void* desc, int32_t* page_offsets) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        if (ctx->h_pool.size() != (size_t)ctx->page_off.back() * ctx->desc_bytes)
            throw StateError("this ctx holds no host copy of the pool (it was imported through the device view)");
        if (desc && !ctx->h_pool.empty()) std::memcpy(desc, ctx->h_pool.data(), ctx->h_pool.size());
        if (page_offsets) std::memcpy(page_offsets, ctx->page_off.data(), ctx->page_off.size() * 4);
    });
}

int32_t slideo_b200_pool_import(slideo_b200_ctx* ctx, const void* desc, int32_t n_desc, const int32_t* page_offsets,
                                int32_t n_pages) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        arg(n_desc >= 0 && n_pages >= 0, "negative size");
        arg(page_offsets != nullptr, "page_offsets must not be NULL");
        arg(desc != nullptr || n_desc == 0, "desc must not be NULL");
        arg(page_offsets[0] == 0 && page_offsets[n_pages] == n_desc, "page_offsets must start at 0 and end at n_desc");
        for (int p = 0; p < n_pages; ++p) arg(page_offsets[p] <= page_offsets[p + 1], "page_offsets must be non-decreasing");
        ctx->check_pool_limits(n_desc, n_pages);
        ctx->drain();
        ctx->close_epoch();
        ctx->h_pool.assign((const uint8_t*)desc, (const uint8_t*)desc + (size_t)n_desc * ctx->desc_bytes);
        ctx->h_pool_pt.clear();
        ctx->pool_pts_valid = n_desc == 0;
        ctx->page_off.assign(page_offsets, page_offsets + n_pages + 1);
        ctx->reserved = false;
        finalize_from_host(ctx);
    });
}

int32_t slideo_b200_pool_reserve(slideo_b200_ctx* ctx, int32_t n_desc, int32_t n_pages) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        arg(n_desc >= 0 && n_pages >= 0, "negative size");
        ctx->check_pool_limits(n_desc, n_pages);
        ctx->drain();
        ctx->close_epoch();
        ctx->finalized = false;
        ctx->reserved = true;
        ctx->nt = n_desc;
        ctx->n_pages = n_pages;
        ctx->h_pool.clear();
        ctx->h_pool_pt.clear();
        ctx->pool_pts_valid = n_desc == 0;
        ctx->pts_received = false;
        ctx->page_small_received = false;
        ctx->page_small_ready = false;
        ctx->page_geom_ok = false;
        ctx->h_page_small.clear();
        ctx->page_classes.clear();
        ctx->h_page_class.clear();
        ctx->h_page_small_off.clear();
        ctx->page_off.assign((size_t)n_pages + 1, 0);
        if (ctx->cfg.descriptor_kind == SLIDEO_B200_DESC_ORB256) ctx->d_pool.reserve((size_t)std::max(n_desc, 1) * 32 + 64);
        else ctx->d_pool_f32.reserve((size_t)std::max(n_desc, 1) * 512);
        ctx->d_page_off.reserve((size_t)n_pages + 1);
    });
}

int32_t slideo_b200_pool_device_view(slideo_b200_ctx* ctx, void** d_desc, size_t* desc_bytes, void** d_page_offsets,
                                     size_t* offsets_bytes) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        if (!ctx->finalized && !ctx->reserved) throw StateError("no device pool yet (finalize_pool or pool_reserve first)");
        const bool orb = ctx->cfg.descriptor_kind == SLIDEO_B200_DESC_ORB256;
        if (d_desc) *d_desc = orb ? (void*)ctx->d_pool.p : (void*)ctx->d_pool_f32.p;
        if (desc_bytes) *desc_bytes = (size_t)ctx->nt * ctx->desc_bytes;
        if (d_page_offsets) *d_page_offsets = ctx->d_page_off.p;
        if (offsets_bytes) *offsets_bytes = ((size_t)ctx->n_pages + 1) * 4;
    });
}

int32_t slideo_b200_pool_points_device_view(slideo_b200_ctx* ctx, void** d_pt, size_t* bytes, int32_t* has_points, int32_t received) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        if (!ctx->finalized && !ctx->reserved) throw StateError("no device pool yet (finalize_pool or pool_reserve first)");
        if (ctx->reserved) {
            ctx->d_pool_pt.reserve((size_t)std::max(ctx->nt, 1));
            if (received) ctx->pts_received = true;
        }
        if (d_pt) *d_pt = ctx->d_pool_pt.p;
        if (bytes) *bytes = (size_t)ctx->nt * 8;
        if (has_points) *has_points = ctx->reserved ? (ctx->pts_received ? 1 : 0) : (ctx->pool_pts_valid ? 1 : 0);
    });
}

int32_t slideo_b200_pool_pages_device_view(slideo_b200_ctx* ctx, void** d_small, size_t* bytes, int32_t* page_w, int32_t* page_h,
                                           int32_t set_w, int32_t set_h) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        if (!ctx->finalized && !ctx->reserved) throw StateError("no device pool yet (finalize_pool or pool_reserve first)");
        if (ctx->reserved && set_w > 0 && set_h > 0) {   // the receiving side: size the buffer for the sender's (uniform) page geometry
            ctx->page_classes.clear();
            const int pc = ctx->page_class_index(set_w, set_h);
            const AreaTables& ar = ctx->page_classes[(size_t)pc]->area;
            const size_t sb = (size_t)ar.dw * ar.dh;
            ctx->d_page_small.reserve((size_t)std::max(ctx->n_pages, 1) * sb);
            ctx->h_page_class.assign((size_t)ctx->n_pages, 0);
            ctx->h_page_small_off.resize((size_t)ctx->n_pages);
            for (int p = 0; p < ctx->n_pages; ++p) ctx->h_page_small_off[(size_t)p] = (unsigned long long)p * sb;
            ctx->upload_page_geometry();
            ctx->page_geom_ok = true;
            ctx->page_small_received = true;
        }
        // replication moves ONE page size: a deck of mixed page sizes reports 0 x 0 here and keeps its gates on the rank that built it
        const bool uniform = ctx->page_classes.size() == 1;
        const bool have = uniform && (ctx->reserved ? ctx->page_small_received : (ctx->page_geom_ok && ctx->page_small_ready));
        const AreaTables* ar = uniform ? &ctx->page_classes[0]->area : nullptr;
        if (d_small) *d_small = have ? (void*)ctx->d_page_small.p : nullptr;
        if (bytes) *bytes = have ? (size_t)ctx->n_pages * ar->dw * ar->dh : 0;
        if (page_w) *page_w = have ? ctx->page_classes[0]->w : 0;
        if (page_h) *page_h = have ? ctx->page_classes[0]->h : 0;
    });
}

int32_t slideo_b200_pool_commit(slideo_b200_ctx* ctx) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        if (!ctx->reserved) throw StateError("pool_commit without pool_reserve");
        SLIDEO_CUDA(cudaDeviceSynchronize());  // the broadcast ran on a stream the library does not own
        SLIDEO_CUDA(cudaMemcpy(ctx->page_off.data(), ctx->d_page_off.p, ((size_t)ctx->n_pages + 1) * 4, cudaMemcpyDeviceToHost));
        if (ctx->page_off[0] != 0 || ctx->page_off[ctx->n_pages] != ctx->nt) throw ArgError("received page offsets do not match the reserved geometry");
        for (int p = 0; p < ctx->n_pages; ++p)
            if (ctx->page_off[p] > ctx->page_off[p + 1]) throw ArgError("received page offsets are not monotone");
        const bool got_pts = ctx->pts_received;
        if (ctx->cfg.descriptor_kind == SLIDEO_B200_DESC_SIFT128) {   // the broadcast moved fp32 rows; derive the K10 operands
            ctx->d_pool.reserve(l2_main_bytes(ctx->nt));
            ctx->d_pool_tail.reserve(l2_tail_bytes(ctx->nt));
            l2_prepare_launch((const float*)ctx->d_pool_f32.p, ctx->nt, false, ctx->d_pool.p, ctx->d_pool_tail.p, ctx->stream);
            ctx->tm.kernel_launches += 1;
        }
        ctx->build_page_of();              // (no host copy of the coordinates on this rank: the device copy filled by the caller stays)
        if (got_pts) ctx->pool_pts_valid = true;
        if (ctx->page_small_received) ctx->page_small_ready = true;
        ctx->reserved = false;
        ctx->finalized = true;
    });
}

// ---- the per-frame hot path -----------------------------------------------------------------------------------
namespace {

// shared body of the synchronous frame entry points: submit + collect on the device-driven stream (ORB256), or the K11 -> K10
// loop (SIFT128)
void match_frames_common(slideo_b200_ctx* ctx, const uint8_t* frames, bool on_device, int n, int w, int h, int stride, size_t frame_stride,
                         slideo_b200_frame_result* out) {
    ctx->reset_kept();
    if (ctx->cfg.descriptor_kind == SLIDEO_B200_DESC_SIFT128) {
        ctx->ensure_host_results((size_t)n);
        EventPair t_total = ctx->begin_timing(4, ctx->stream);
        ctx->match_frames_sift(frames, on_device, n, w, h, stride, frame_stride);
        ctx->end_timing(t_total, ctx->stream);
        SLIDEO_CUDA(cudaStreamSynchronize(ctx->stream));
        SLIDEO_CUDA(cudaStreamSynchronize(ctx->copy_stream));
        ctx->collect_timings();
        for (int i = 0; i < n; ++i) {
            out[i].best_slide = ctx->h_results[3 * i];
            out[i].votes = ctx->h_results[3 * i + 1];
            out[i].n_keypoints = ctx->h_results[3 * i + 2];
        }
        return;
    }
    if (ctx->cfg.geometric_verification || ctx->cfg.keep_matches) {
        ctx->match_verified(frames, on_device, n, w, h, stride, frame_stride, out);
        return;
    }
    for (int f0 = 0; f0 < n; f0 += slideo_b200_ctx::RING / 2) {   // (the host result ring bounds one ticket)
        const int nn = std::min(slideo_b200_ctx::RING / 2, n - f0);
        const long long id = ctx->submit(frames + (size_t)f0 * frame_stride, on_device, nn, w, h, stride, frame_stride);
        ctx->collect(id, out + f0, nn);
    }
}

}  // namespace

int32_t slideo_b200_match_frames_bgr8(slideo_b200_ctx* ctx, const uint8_t* frames, int32_t n, int32_t w, int32_t h,
                                      int32_t stride, size_t frame_stride, slideo_b200_frame_result* out) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        ctx->require_pool();
        arg(n >= 0, "n < 0");
        if (n == 0) return;
        arg(frames && out, "frames/out must not be NULL");
        arg(stride >= 3 * w, "stride < 3*w");
        arg(frame_stride >= (size_t)stride * (h - 1) + (size_t)3 * w, "frame_stride too small");
        match_frames_common(ctx, frames, false, n, w, h, stride, frame_stride, out);
    });
}

int32_t slideo_b200_match_frames_bgr8_device(slideo_b200_ctx* ctx, const void* d_frames, int32_t n, int32_t w, int32_t h,
                                             int32_t stride, size_t frame_stride, slideo_b200_frame_result* out) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        ctx->require_pool();
        arg(n >= 0, "n < 0");
        if (n == 0) return;
        arg(d_frames && out, "d_frames/out must not be NULL");
        arg(stride >= 3 * w, "stride < 3*w");
        match_frames_common(ctx, (const uint8_t*)d_frames, true, n, w, h, stride, frame_stride, out);
    });
}

// ---- the same path, asynchronous: submit returns as soon as everything is enqueued, collect waits for one ticket ---------
int32_t slideo_b200_submit_frames_bgr8(slideo_b200_ctx* ctx, const uint8_t* frames, int32_t n, int32_t w, int32_t h, int32_t stride,
                                       size_t frame_stride, int64_t* out_ticket) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        ctx->require_pool();
        ctx->require_orb();
        arg(n >= 1 && n <= slideo_b200_ctx::RING / 2, "n must be in 1..32768");
        arg(frames && out_ticket, "frames/out_ticket must not be NULL");
        arg(stride >= 3 * w, "stride < 3*w");
        arg(frame_stride >= (size_t)stride * (h - 1) + (size_t)3 * w, "frame_stride too small");
        if (ctx->cfg.geometric_verification || ctx->cfg.keep_matches)
            throw NotImplError("submit/collect carries the vote results only: use match_frames_* for the verification tail / kept matches");
        *out_ticket = ctx->submit(frames, false, n, w, h, stride, frame_stride);
    });
}

int32_t slideo_b200_submit_frames_bgr8_device(slideo_b200_ctx* ctx, const void* d_frames, int32_t n, int32_t w, int32_t h, int32_t stride,
                                              size_t frame_stride, int64_t* out_ticket) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        ctx->require_pool();
        ctx->require_orb();
        arg(n >= 1 && n <= slideo_b200_ctx::RING / 2, "n must be in 1..32768");
        arg(d_frames && out_ticket, "d_frames/out_ticket must not be NULL");
        arg(stride >= 3 * w, "stride < 3*w");
        if (ctx->cfg.geometric_verification || ctx->cfg.keep_matches)
            throw NotImplError("submit/collect carries the vote results only: use match_frames_* for the verification tail / kept matches");
        *out_ticket = ctx->submit((const uint8_t*)d_frames, true, n, w, h, stride, frame_stride);
    });
}

int32_t slideo_b200_collect(slideo_b200_ctx* ctx, int64_t ticket, slideo_b200_frame_result* out, int32_t cap, int32_t* out_n) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        ctx->require_orb();
        const int n = ctx->collect(ticket, out, cap);
        if (out_n) *out_n = n;
    });
}

int32_t slideo_b200_match_descriptors(slideo_b200_ctx* ctx, const void* desc, const int32_t* frame_offsets, int32_t n,
                                      slideo_b200_frame_result* out) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        ctx->require_pool();
        arg(n >= 0, "n < 0");
        if (n == 0) return;
        arg(frame_offsets && out, "frame_offsets/out must not be NULL");
        arg(frame_offsets[0] == 0, "frame_offsets[0] must be 0");
        for (int i = 0; i < n; ++i) arg(frame_offsets[i] <= frame_offsets[i + 1], "frame_offsets must be non-decreasing");
        arg(desc != nullptr || frame_offsets[n] == 0, "desc must not be NULL");
        ctx->reset_kept();
        EventPair t_total = ctx->begin_timing(4, ctx->stream);
        const int B = ctx->cfg.max_batch;
        const bool orb = ctx->cfg.descriptor_kind == SLIDEO_B200_DESC_ORB256;
        ctx->ensure_host_results((size_t)n);
        std::vector<int32_t> qf, nkp, fo;
        for (int f0 = 0; f0 < n; f0 += B) {
            const int nb = std::min(B, n - f0);
            const int q0 = frame_offsets[f0], nq = frame_offsets[f0 + nb] - q0;
            qf.resize((size_t)nq);
            nkp.resize((size_t)nb);
            fo.assign(1, 0);
            for (int i = 0; i < nb; ++i) {
                nkp[i] = frame_offsets[f0 + i + 1] - frame_offsets[f0 + i];
                for (int j = frame_offsets[f0 + i] - q0; j < frame_offsets[f0 + i + 1] - q0; ++j) qf[(size_t)j] = i;
                fo.push_back(frame_offsets[f0 + i + 1] - q0);
            }
            ctx->d_q.reserve((size_t)std::max(nq, 1) * ctx->desc_bytes);
            ctx->d_q_frame.reserve((size_t)nq + nb + 1);
            int32_t* d_nkp = ctx->d_q_frame.p + nq;
            EventPair t = ctx->begin_timing(2, ctx->stream);
            if (nq > 0) {
                SLIDEO_CUDA(cudaMemcpyAsync(ctx->d_q.p, (const uint8_t*)desc + (size_t)q0 * ctx->desc_bytes,
                                            (size_t)nq * ctx->desc_bytes, cudaMemcpyHostToDevice, ctx->stream));
                SLIDEO_CUDA(cudaMemcpyAsync(ctx->d_q_frame.p, qf.data(), (size_t)nq * 4, cudaMemcpyHostToDevice, ctx->stream));
            }
            SLIDEO_CUDA(cudaMemcpyAsync(d_nkp, nkp.data(), (size_t)nb * 4, cudaMemcpyHostToDevice, ctx->stream));
            ctx->end_timing(t, ctx->stream);
            if (orb) {
                ctx->knn_vote_hamming(ctx->d_q.p, nq, ctx->d_q_frame.p, nb, d_nkp, ctx->cfg.keep_matches != 0);
            } else {
                // SIFT128: tcgen05 L2 k-NN, then the same vote on (distance, index) rows
                ctx->d_votes.reserve((size_t)nb * std::max(ctx->n_pages, 1));
                ctx->d_results.reserve((size_t)nb * 3);
                ctx->d_idx.reserve((size_t)std::max(nq, 1) * ctx->cfg.knn_k);
                ctx->d_dist.reserve((size_t)std::max(nq, 1) * ctx->cfg.knn_k);
                SLIDEO_CUDA(cudaMemsetAsync(ctx->d_votes.p, 0, (size_t)nb * std::max(ctx->n_pages, 1) * 4, ctx->stream));
                if (nq > 0) {
                    EventPair tk = ctx->begin_timing(1, ctx->stream);
                    int nl = 0;
                    l2_knn_launch(ctx->l2ws, (const float*)ctx->d_q.p, nq, ctx->d_pool.p, ctx->d_pool_tail.p, ctx->nt,
                                  ctx->cfg.knn_k, ctx->d_idx.p, (float*)ctx->d_dist.p, ctx->num_sms, ctx->stream, &nl);
                    l2_vote_launch(ctx->d_idx.p, (const float*)ctx->d_dist.p, nq, ctx->cfg.knn_k, ctx->d_q_frame.p, ctx->d_page_of.p,
                                   ctx->d_votes.p, ctx->n_pages, ctx->cfg.vote_ratio, ctx->stream);
                    ctx->end_timing(tk, ctx->stream);
                    ctx->tm.knn_launches += nl;
                    ctx->tm.kernel_launches += nl + 1;
                    ctx->tm.knn_pairs += (int64_t)nq * ctx->nt;
                }
                vote_argmax_launch(ctx->d_votes.p, nb, ctx->n_pages, d_nkp, ctx->d_results.p, ctx->stream);
                ctx->tm.kernel_launches += 1;
            }
            SLIDEO_CUDA(cudaMemcpyAsync(ctx->h_results + (size_t)f0 * 3, ctx->d_results.p, (size_t)nb * 3 * 4, cudaMemcpyDeviceToHost,
                                        ctx->stream));
            if (ctx->cfg.keep_matches) {
                if (orb) {
                    ctx->keep_batch_keys(nq, fo);
                } else {
                    const size_t base = ctx->kept_l2.size();
                    ctx->kept_l2.resize(base + (size_t)nq * ctx->cfg.knn_k);
                    ctx->kept_l2_idx.resize(base + (size_t)nq * ctx->cfg.knn_k);
                    if (nq > 0) {
                        SLIDEO_CUDA(cudaMemcpyAsync(ctx->kept_l2.data() + base, ctx->d_dist.p, (size_t)nq * ctx->cfg.knn_k * 4,
                                                    cudaMemcpyDeviceToHost, ctx->stream));
                        SLIDEO_CUDA(cudaMemcpyAsync(ctx->kept_l2_idx.data() + base, ctx->d_idx.p, (size_t)nq * ctx->cfg.knn_k * 4,
                                                    cudaMemcpyDeviceToHost, ctx->stream));
                    }
                    const int32_t qb = ctx->kept_frame_off.back();
                    for (size_t i = 1; i < fo.size(); ++i) ctx->kept_frame_off.push_back(qb + fo[i]);
                }
            }
            SLIDEO_CUDA(cudaStreamSynchronize(ctx->stream));  // host staging vectors are reused by the next batch
            ctx->tm.frames += nb;
        }
        ctx->end_timing(t_total, ctx->stream);
        SLIDEO_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->collect_timings();
        for (int i = 0; i < n; ++i) {
            out[i].best_slide = ctx->h_results[3 * i];
            out[i].votes = ctx->h_results[3 * i + 1];
            out[i].n_keypoints = ctx->h_results[3 * i + 2];
        }
    });
}

int32_t slideo_b200_get_matches(slideo_b200_ctx* ctx, int32_t frame_i, slideo_b200_match* out, int32_t cap_rows,
                                int32_t* out_rows) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        if (!ctx->cfg.keep_matches) throw StateError("cfg.keep_matches was not set");
        arg(frame_i >= 0 && frame_i + 1 < (int)ctx->kept_frame_off.size(), "frame_i out of range of the last match call");
        const int q0 = ctx->kept_frame_off[frame_i], q1 = ctx->kept_frame_off[frame_i + 1], k = ctx->cfg.knn_k;
        if (out_rows) *out_rows = q1 - q0;
        if (!out) return;
        arg(cap_rows >= q1 - q0, "cap_rows too small");
        const bool orb = ctx->cfg.descriptor_kind == SLIDEO_B200_DESC_ORB256;
        for (int q = q0; q < q1; ++q)
            for (int j = 0; j < k; ++j) {
                slideo_b200_match& m = out[(size_t)(q - q0) * k + j];
                int gi;
                float dist;
                if (orb) {
                    const uint32_t key = ctx->kept_keys[(size_t)q * k + j];
                    gi = key == KEY_EMPTY ? -1 : (int)(key & KEY_IDX_MASK);
                    dist = (float)(key >> KEY_IDX_BITS);
                } else {
                    gi = ctx->kept_l2_idx[(size_t)q * k + j];
                    dist = ctx->kept_l2[(size_t)q * k + j];
                }
                m.query_idx = q - q0;
                if (gi < 0) {
                    m.train_idx = -1;
                    m.source = -1;
                    m.distance = -1.f;
                    continue;
                }
                const int page = (int)(std::upper_bound(ctx->page_off.begin(), ctx->page_off.end(), gi) - ctx->page_off.begin()) - 1;
                m.source = page;
                m.train_idx = gi - ctx->page_off[page];
                m.distance = dist;
            }
    });
}

int32_t slideo_b200_get_verification(slideo_b200_ctx* ctx, int32_t frame0, int32_t n, slideo_b200_verify_result* out) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        if (!ctx->cfg.geometric_verification) throw StateError("cfg.geometric_verification was not set");
        arg(frame0 >= 0 && n >= 0 && (size_t)frame0 + (size_t)n <= ctx->verify_results.size(), "frame range outside the last match call");
        arg(out != nullptr || n == 0, "out must not be NULL");
        static_assert(sizeof(slideo_b200_verify_result) == sizeof(VerifyRecord), "ABI struct and kernel record must agree");
        if (n) std::memcpy(out, ctx->verify_results.data() + frame0, (size_t)n * sizeof(VerifyRecord));
    });
}

int32_t slideo_b200_get_decisions(slideo_b200_ctx* ctx, int32_t frame0, int32_t n, slideo_b200_decision* out) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        if (ctx->cfg.geometric_verification < 2) throw StateError("cfg.geometric_verification is not 2");
        arg(frame0 >= 0 && n >= 0 && (size_t)frame0 + (size_t)n <= ctx->decisions.size(), "frame range outside the last match call");
        arg(out != nullptr || n == 0, "out must not be NULL");
        if (n) std::memcpy(out, ctx->decisions.data() + frame0, (size_t)n * sizeof(slideo_b200_decision));
    });
}

}  // extern "C"

namespace {

// MarkSimilarIter over n frames; frames are fetched batch by batch through `get_batch(f0, nb)` which returns a device pointer to
// nb packed frames (row stride `stride`, frame stride `frame_stride`) valid on ctx->stream
template <typename GetBatch>
void mark_changed_impl(slideo_b200_ctx* ctx, int n, int w, int h, bool reset, uint8_t* out_changed, float* out_similarity, GetBatch&& get_batch) {
    if (ctx->area.sw != w || ctx->area.sh != h) {
        SLIDEO_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->area.build(w, h);
        ctx->have_prev_small = false;
    }
    if (reset) ctx->have_prev_small = false;
    const int B = ctx->cfg.max_batch;
    const size_t small_bytes = (size_t)ctx->area.dw * ctx->area.dh * 3;
    ctx->d_small.reserve((size_t)(B + 1) * small_bytes);   // fixed per geometry: slot 0 (the chain state) must survive from call to call
    ctx->d_sumsq.reserve((size_t)B + 1);
    std::vector<unsigned long long> h_ss((size_t)B + 1);
    const int p = ctx->area.dw * ctx->area.dh;
    const float max_error = sqrtf((255.0f * 255.0f * 3.0f) * (float)p);   // image_utils.rs:24-25 (f32)
    for (int f0 = 0; f0 < n; f0 += B) {
        const int nb = std::min(B, n - f0);
        int stride = 0;
        size_t frame_stride = 0;
        const uint8_t* d_src = get_batch(f0, nb, &stride, &frame_stride);
        EventPair t = ctx->begin_timing(0, ctx->stream);
        area_small_launch(ctx->area, d_src, nb, stride, frame_stride, ctx->d_small.p + small_bytes, ctx->stream);
        // pairs (slot i, slot i + 1): slot 0 holds the previous batch's last small image
        small_sumsq_launch(ctx->d_small.p, nb, small_bytes, ctx->d_sumsq.p, ctx->stream);
        ctx->end_timing(t, ctx->stream);
        ctx->tm.kernel_launches += 2;
        SLIDEO_CUDA(cudaMemcpyAsync(h_ss.data(), ctx->d_sumsq.p, (size_t)nb * 8, cudaMemcpyDeviceToHost, ctx->stream));
        SLIDEO_CUDA(cudaMemcpyAsync(ctx->d_small.p, ctx->d_small.p + (size_t)nb * small_bytes, small_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        SLIDEO_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < nb; ++i) {
            float sim = 0.0f;   // video_capture.rs:92-96: no previous frame -> 0.0
            if (i > 0 || ctx->have_prev_small) {
                const double error_l2 = sqrt((double)h_ss[(size_t)i]);            // norm2(a, b, NORM_L2)
                sim = 1.0f - (float)error_l2 / max_error;                          // image_utils.rs:26
            }
            if (out_similarity) out_similarity[f0 + i] = sim;
            out_changed[f0 + i] = sim < 0.98f ? 1 : 0;                            // video_capture.rs:98
        }
        ctx->have_prev_small = true;
    }
    ctx->collect_timings();
}

}  // namespace

extern "C" {

int32_t slideo_b200_mark_changed_bgr8(slideo_b200_ctx* ctx, const uint8_t* frames, int32_t n, int32_t w, int32_t h, int32_t stride,
                                      size_t frame_stride, int32_t reset, uint8_t* out_changed, float* out_similarity) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        arg(n >= 0, "n < 0");
        if (n == 0) return;
        arg(frames && out_changed, "frames/out_changed must not be NULL");
        arg(w >= 16 && h >= 16 && w <= 16384 && h <= 16384, "frame size out of range");
        arg(stride >= 3 * w, "stride < 3*w");
        arg(frame_stride >= (size_t)stride * (h - 1) + (size_t)3 * w, "frame_stride too small");
        const size_t img_bytes = (size_t)3 * w * h;
        if ((size_t)std::min(ctx->cfg.max_batch, n) * img_bytes > ctx->d_mark_frames.cap) SLIDEO_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->d_mark_frames.reserve((size_t)std::min(ctx->cfg.max_batch, n) * img_bytes);
        mark_changed_impl(ctx, n, w, h, reset != 0, out_changed, out_similarity, [&](int f0, int nb, int* st, size_t* fst) -> const uint8_t* {
            EventPair t = ctx->begin_timing(2, ctx->stream);
            ctx->upload_images(ctx->d_mark_frames.p, frames + (size_t)f0 * frame_stride, nb, 3 * w, h, stride, frame_stride, ctx->stream);
            ctx->end_timing(t, ctx->stream);
            *st = 3 * w;
            *fst = img_bytes;
            return ctx->d_mark_frames.p;
        });
    });
}

int32_t slideo_b200_mark_changed_bgr8_device(slideo_b200_ctx* ctx, const void* d_frames, int32_t n, int32_t w, int32_t h, int32_t stride,
                                             size_t frame_stride, int32_t reset, uint8_t* out_changed, float* out_similarity) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        arg(n >= 0, "n < 0");
        if (n == 0) return;
        arg(d_frames && out_changed, "d_frames/out_changed must not be NULL");
        arg(w >= 16 && h >= 16 && w <= 16384 && h <= 16384, "frame size out of range");
        arg(stride >= 3 * w, "stride < 3*w");
        mark_changed_impl(ctx, n, w, h, reset != 0, out_changed, out_similarity, [&](int f0, int, int* st, size_t* fst) -> const uint8_t* {
            *st = stride;
            *fst = frame_stride;
            return (const uint8_t*)d_frames + (size_t)f0 * frame_stride;
        });
    });
}

// ---- stage-level entry points -----------------------------------------------------------------------------------
int32_t slideo_b200_extract_orb(slideo_b200_ctx* ctx, const uint8_t* img, int32_t w, int32_t h, int32_t stride, int32_t channels,
                                int32_t* kp_i, float* kp_f, uint8_t* desc, int32_t cap, int32_t* out_n) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        arg(img != nullptr, "img must not be NULL");
        arg(channels == 1 || channels == 3, "channels must be 1 or 3");
        arg(stride >= w * channels, "stride too small");
        OrbExtractor& ex = ctx->extractor(w, h, 1);
        ctx->d_img.reserve((size_t)w * h * channels);
        ctx->upload_images(ctx->d_img.p, img, 1, w * channels, h, stride, (size_t)stride * h, ctx->stream);
        int nl = 0;
        EventPair t = ctx->begin_timing(0, ctx->stream);
        const int total = ex.run(ctx->d_img.p, 1, w * channels, (size_t)w * h * channels, channels, ctx->stream, &nl);
        ctx->end_timing(t, ctx->stream);
        ctx->tm.kernel_launches += nl;
        if (out_n) *out_n = total;
        if (total > cap && (kp_i || kp_f || desc)) throw CapacityError("cap smaller than the number of keypoints");
        if (total > 0) {
            if (kp_i) SLIDEO_CUDA(cudaMemcpyAsync(kp_i, ex.d_kp_i(), (size_t)total * 16, cudaMemcpyDeviceToHost, ctx->stream));
            if (kp_f) SLIDEO_CUDA(cudaMemcpyAsync(kp_f, ex.d_kp_f(), (size_t)total * 16, cudaMemcpyDeviceToHost, ctx->stream));
            if (desc) SLIDEO_CUDA(cudaMemcpyAsync(desc, ex.d_desc(), (size_t)total * 32, cudaMemcpyDeviceToHost, ctx->stream));
        }
        SLIDEO_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->collect_timings();
    });
}

int32_t slideo_b200_debug_fetch(slideo_b200_ctx* ctx, int32_t what, int32_t level, void* out, size_t cap_bytes, int32_t* out_w,
                                int32_t* out_h) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        OrbExtractor* ex = ctx->last_ext;
        if (!ex) throw StateError("no extract/match call yet");
        arg(level >= 0 && level < ex->nlevels(), "level out of range");
        const OrbLevelGeom& L = ex->level(level);
        SLIDEO_CUDA(cudaStreamSynchronize(ctx->stream));
        if (what == 0 || what == 1) {
            if (out_w) *out_w = L.w;
            if (out_h) *out_h = L.h;
            if (!out) return;
            arg(cap_bytes >= (size_t)L.w * L.h, "cap_bytes too small");
            const uint8_t* src = (what == 0 ? ex->d_pyramid(0) : ex->d_blurred(0)) + L.img_off;
            SLIDEO_CUDA(cudaMemcpy2D(out, L.w, src, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost));
        } else if (what == 2) {
            int32_t cnt = 0;
            SLIDEO_CUDA(cudaMemcpy(&cnt, ex->d_cand_count() + level, 4, cudaMemcpyDeviceToHost));
            cnt = std::min(cnt, L.cand_cap);
            if (out_w) *out_w = cnt;
            if (out_h) *out_h = 1;
            if (!out) return;
            arg(cap_bytes >= (size_t)cnt * 4, "cap_bytes too small");
            if (cnt) SLIDEO_CUDA(cudaMemcpy(out, ex->d_candidates(0) + L.cand_off, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
        } else {
            throw ArgError("unknown `what`");
        }
    });
}

int32_t slideo_b200_extract_sift(slideo_b200_ctx* ctx, const uint8_t* img, int32_t w, int32_t h, int32_t stride, int32_t channels,
                                 float* kp_f, int32_t* kp_octave, float* desc, int32_t cap, int32_t* out_n) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        arg(img != nullptr, "img must not be NULL");
        arg(channels == 1 || channels == 3, "channels must be 1 or 3");
        arg(stride >= w * channels, "stride too small");
        SiftExtractor& sx = ctx->sift_extractor(w, h, 1);
        ctx->d_img.reserve((size_t)w * h * channels);
        ctx->upload_images(ctx->d_img.p, img, 1, w * channels, h, stride, (size_t)stride * h, ctx->stream);
        int nl = 0;
        EventPair t = ctx->begin_timing(0, ctx->stream);
        const int total = sx.run(ctx->d_img.p, 1, w * channels, (size_t)w * h * channels, channels, ctx->stream, &nl);
        ctx->end_timing(t, ctx->stream);
        ctx->tm.kernel_launches += nl;
        if (out_n) *out_n = total;
        if (total > cap && (kp_f || kp_octave || desc)) throw CapacityError("cap smaller than the number of keypoints");
        if (total > 0) {
            if (kp_f) SLIDEO_CUDA(cudaMemcpyAsync(kp_f, sx.d_kp_f(), (size_t)total * 20, cudaMemcpyDeviceToHost, ctx->stream));
            if (kp_octave) SLIDEO_CUDA(cudaMemcpyAsync(kp_octave, sx.d_kp_octave(), (size_t)total * 4, cudaMemcpyDeviceToHost, ctx->stream));
            if (desc) SLIDEO_CUDA(cudaMemcpyAsync(desc, sx.d_desc(), (size_t)total * 512, cudaMemcpyDeviceToHost, ctx->stream));
        }
        SLIDEO_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->collect_timings();
    });
}

int32_t slideo_b200_debug_fetch_sift(slideo_b200_ctx* ctx, int32_t octave, int32_t layer, float* out, size_t cap_bytes, int32_t* out_w,
                                     int32_t* out_h, int32_t* out_n_octaves) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        SiftExtractor* sx = ctx->last_sift;
        if (!sx) throw StateError("no SIFT extract/match call yet");
        const SiftGeo& g = sx->geo();
        if (out_n_octaves) *out_n_octaves = g.n_oct;
        arg(octave >= 0 && octave < g.n_oct, "octave out of range");
        arg(layer >= 0 && layer < SIFT_GAUSS, "layer out of range");
        const SiftOctave& oc = g.oc[octave];
        if (out_w) *out_w = oc.w;
        if (out_h) *out_h = oc.h;
        if (!out) return;
        arg(cap_bytes >= (size_t)oc.w * oc.h * 4, "cap_bytes too small");
        SLIDEO_CUDA(cudaStreamSynchronize(ctx->stream));
        SLIDEO_CUDA(cudaMemcpy2D(out, (size_t)oc.w * 4, sx->d_gauss(0, octave, layer), (size_t)oc.pitch * 4, (size_t)oc.w * 4, oc.h, cudaMemcpyDeviceToHost));
    });
}

int32_t slideo_b200_bf_knn_hamming_device(slideo_b200_ctx* ctx, const void* d_q, int32_t nq, const void* d_t, int32_t nt, int32_t k,
                                          void* d_keys_out) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        arg(nq >= 0 && nt >= 0, "negative size");
        arg(k >= 1 && k <= KNN_MAX_K, "k must be in 1..32");
        arg(nt <= KNN_MAX_POOL, "nt exceeds 8,388,608");
        if (nq == 0) return;
        arg(d_q && d_keys_out && (d_t || nt == 0), "NULL buffer");
        arg(((uintptr_t)d_q & 15) == 0 && ((uintptr_t)d_t & 15) == 0, "device buffers must be 16-byte aligned");
        int nl = 0;
        if (ctx->cfg.knn_impl == 4) {   // the XOR / POPC kernel of round 1, kept as an independent implementation
            KnnPlan plan = knn_hamming_plan(nq, nt, k, ctx->num_sms);
            ctx->d_scratch.reserve(plan.scratch_bytes / 4);
            if (plan.partial_bytes) ctx->d_partial.reserve(plan.partial_bytes / 4);
            ctx->d_t48.reserve(plan.pool_bytes + 64);
            knn_pool_expand_launch(d_t, nt, ctx->d_t48.p, ctx->stream);
            EventPair t = ctx->begin_timing(1, ctx->stream);
            knn_hamming_launch(plan, d_q, ctx->d_t48.p, (uint32_t*)d_keys_out, ctx->d_scratch.p, ctx->d_partial.p, nullptr, ctx->stream, &nl);
            ctx->end_timing(t, ctx->stream);
        } else {
            Knn5Plan plan = knn5_plan(nq, nt, k, ctx->num_sms);
            if (plan.partial_bytes) ctx->d_partial5.reserve(plan.partial_bytes / 4);
            ctx->d_t5.reserve(knn5_pool_bytes(std::max(nt, 1)) + 64);
            knn5_pool_prepare_launch(d_t, nt, ctx->d_t5.p, ctx->stream);   // once per pool in the frame path; per call here
            EventPair t = ctx->begin_timing(1, ctx->stream);
            knn5_launch(plan, d_q, ctx->d_t5.p, (uint32_t*)d_keys_out, ctx->d_partial5.p, nullptr, ctx->stream, &nl);
            ctx->end_timing(t, ctx->stream);
        }
        ctx->tm.knn_launches += nl;
        ctx->tm.kernel_launches += nl;
        ctx->tm.knn_pairs += (int64_t)nq * nt;
    });
}

int32_t slideo_b200_bf_knn_hamming(slideo_b200_ctx* ctx, const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt, int32_t k,
                                   int32_t* idx, int32_t* dist) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        arg(nq >= 0 && nt >= 0, "negative size");
        arg(k >= 1 && k <= KNN_MAX_K, "k must be in 1..32");
        arg(nt <= KNN_MAX_POOL, "nt exceeds 8,388,608");
        if (nq == 0) return;
        arg(q && idx && dist && (t || nt == 0), "NULL buffer");
        ctx->d_q.reserve((size_t)nq * 32);
        ctx->d_t.reserve((size_t)std::max(nt, 1) * 32);
        ctx->d_keys.reserve((size_t)nq * k);
        ctx->d_idx.reserve((size_t)nq * k);
        ctx->d_dist.reserve((size_t)nq * k);
        SLIDEO_CUDA(cudaMemcpyAsync(ctx->d_q.p, q, (size_t)nq * 32, cudaMemcpyHostToDevice, ctx->stream));
        if (nt) SLIDEO_CUDA(cudaMemcpyAsync(ctx->d_t.p, t, (size_t)nt * 32, cudaMemcpyHostToDevice, ctx->stream));
        int nl = 0;
        if (ctx->cfg.knn_impl == 4) {
            KnnPlan plan = knn_hamming_plan(nq, nt, k, ctx->num_sms);
            ctx->d_scratch.reserve(plan.scratch_bytes / 4);
            if (plan.partial_bytes) ctx->d_partial.reserve(plan.partial_bytes / 4);
            ctx->d_t48.reserve(plan.pool_bytes + 64);
            knn_pool_expand_launch(ctx->d_t.p, nt, ctx->d_t48.p, ctx->stream);
            EventPair tk = ctx->begin_timing(1, ctx->stream);
            knn_hamming_launch(plan, ctx->d_q.p, ctx->d_t48.p, ctx->d_keys.p, ctx->d_scratch.p, ctx->d_partial.p, nullptr, ctx->stream, &nl);
            ctx->end_timing(tk, ctx->stream);
        } else {
            Knn5Plan plan = knn5_plan(nq, nt, k, ctx->num_sms);
            if (plan.partial_bytes) ctx->d_partial5.reserve(plan.partial_bytes / 4);
            ctx->d_t5.reserve(knn5_pool_bytes(std::max(nt, 1)) + 64);
            knn5_pool_prepare_launch(ctx->d_t.p, nt, ctx->d_t5.p, ctx->stream);
            EventPair tk = ctx->begin_timing(1, ctx->stream);
            knn5_launch(plan, ctx->d_q.p, ctx->d_t5.p, ctx->d_keys.p, ctx->d_partial5.p, nullptr, ctx->stream, &nl);
            ctx->end_timing(tk, ctx->stream);
        }
        keys_to_idx_dist_launch(ctx->d_keys.p, (size_t)nq * k, ctx->d_idx.p, ctx->d_dist.p, ctx->stream);
        ctx->tm.knn_launches += nl;
        ctx->tm.kernel_launches += nl + 1;
        ctx->tm.knn_pairs += (int64_t)nq * nt;
        SLIDEO_CUDA(cudaMemcpyAsync(idx, ctx->d_idx.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
        SLIDEO_CUDA(cudaMemcpyAsync(dist, ctx->d_dist.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
        SLIDEO_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->collect_timings();
    });
}

int32_t slideo_b200_bf_knn_l2(slideo_b200_ctx* ctx, const float* q, int32_t nq, const float* t, int32_t nt, int32_t dim, int32_t k,
                              int32_t* idx, float* dist) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        arg(nq >= 0 && nt >= 0, "negative size");
        arg(dim == 128, "dim must be 128");
        arg(k >= 1 && k <= KNN_MAX_K, "k must be in 1..32");
        arg(nt <= (1 << 24), "nt exceeds 16,777,216");
        if (nq == 0) return;
        arg(q && idx && dist && (t || nt == 0), "NULL buffer");
        ctx->d_q.reserve((size_t)nq * 512);
        ctx->d_t.reserve((size_t)std::max(nt, 1) * 512);
        ctx->d_l2_pool.reserve(l2_main_bytes(nt));
        ctx->d_l2_tail.reserve(l2_tail_bytes(nt));
        ctx->d_idx.reserve((size_t)nq * k);
        ctx->d_dist.reserve((size_t)nq * k);
        SLIDEO_CUDA(cudaMemcpyAsync(ctx->d_q.p, q, (size_t)nq * 512, cudaMemcpyHostToDevice, ctx->stream));
        if (nt) SLIDEO_CUDA(cudaMemcpyAsync(ctx->d_t.p, t, (size_t)nt * 512, cudaMemcpyHostToDevice, ctx->stream));
        l2_prepare_launch((const float*)ctx->d_t.p, nt, false, ctx->d_l2_pool.p, ctx->d_l2_tail.p, ctx->stream);
        EventPair tk = ctx->begin_timing(1, ctx->stream);
        int nl = 0;
        l2_knn_launch(ctx->l2ws, (const float*)ctx->d_q.p, nq, ctx->d_l2_pool.p, ctx->d_l2_tail.p, nt, k, ctx->d_idx.p, (float*)ctx->d_dist.p,
                      ctx->num_sms, ctx->stream, &nl);
        ctx->end_timing(tk, ctx->stream);
        ctx->tm.knn_launches += nl;
        ctx->tm.kernel_launches += nl + 1;
        ctx->tm.knn_pairs += (int64_t)nq * nt;
        SLIDEO_CUDA(cudaMemcpyAsync(idx, ctx->d_idx.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
        SLIDEO_CUDA(cudaMemcpyAsync(dist, ctx->d_dist.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
        SLIDEO_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->collect_timings();
    });
}

int32_t slideo_b200_bf_knn_l2_device(slideo_b200_ctx* ctx, const void* d_q, int32_t nq, const void* d_t, int32_t nt, int32_t dim,
                                     int32_t k, void* d_idx, void* d_dist) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        arg(nq >= 0 && nt >= 0, "negative size");
        arg(dim == 128, "dim must be 128");
        arg(k >= 1 && k <= KNN_MAX_K, "k must be in 1..32");
        arg(nt <= (1 << 24), "nt exceeds 16,777,216");
        if (nq == 0) return;
        arg(d_q && d_idx && d_dist && (d_t || nt == 0), "NULL buffer");
        arg(((uintptr_t)d_q & 15) == 0 && ((uintptr_t)d_t & 15) == 0, "device buffers must be 16-byte aligned");
        // bf16 copy + norms of the pool (outside the K10 timing bracket)
        ctx->d_l2_pool.reserve(l2_main_bytes(nt));
        ctx->d_l2_tail.reserve(l2_tail_bytes(nt));
        l2_prepare_launch((const float*)d_t, nt, false, ctx->d_l2_pool.p, ctx->d_l2_tail.p, ctx->stream);
        EventPair tk = ctx->begin_timing(1, ctx->stream);
        int nl = 0;
        l2_knn_launch(ctx->l2ws, (const float*)d_q, nq, ctx->d_l2_pool.p, ctx->d_l2_tail.p, nt, k, (int32_t*)d_idx,
                      (float*)d_dist, ctx->num_sms, ctx->stream, &nl);
        ctx->end_timing(tk, ctx->stream);
        ctx->tm.knn_launches += nl;
        ctx->tm.kernel_launches += nl;
        ctx->tm.knn_pairs += (int64_t)nq * nt;
    });
}

// ---- utilities ------------------------------------------------------------------------------------------------
int32_t slideo_b200_host_alloc(void** out, size_t bytes) {
    if (!out) return SLIDEO_B200_E_INVALID_ARG;
    *out = nullptr;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaMallocHost failed: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? SLIDEO_B200_E_OOM : SLIDEO_B200_E_CUDA;
    }
    return SLIDEO_B200_OK;
}

int32_t slideo_b200_host_free(void* p) {
    if (!p) return SLIDEO_B200_OK;
    return cudaFreeHost(p) == cudaSuccess ? SLIDEO_B200_OK : SLIDEO_B200_E_CUDA;
}

int32_t slideo_b200_set_progress_callback(slideo_b200_ctx* ctx, slideo_b200_progress_fn fn, void* user) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        ctx->progress_fn = fn;
        ctx->progress_user = user;
    });
}

int32_t slideo_b200_get_timings(slideo_b200_ctx* ctx, slideo_b200_timings* out, int32_t reset) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        ctx->drain();
        ctx->collect_timings();
        if (ctx->d_dyn) {   // pairs handed to K8 by the plan kernels live on the device
            unsigned long long pairs = 0;
            for (int i = 0; i < 2; ++i) {
                KnnStream st;
                SLIDEO_CUDA(cudaMemcpy(&st, ctx->ep[i].d_state, sizeof st, cudaMemcpyDeviceToHost));
                pairs += st.pairs;
            }
            ctx->tm.knn_pairs += (int64_t)(pairs - ctx->pairs_seen);
            ctx->pairs_seen = pairs;
            if (ctx->span_open && ctx->span_closed) {
                float ms = 0.f;
                SLIDEO_CUDA(cudaEventElapsedTime(&ms, ctx->ev_span0, ctx->ev_span1));
                ctx->tm.ms_total += ms;
            }
            ctx->span_open = ctx->span_closed = false;
        }
        if (out) *out = ctx->tm;
        if (reset) ctx->tm = slideo_b200_timings{};
    });
}

int32_t slideo_b200_microbench(slideo_b200_ctx* ctx, int32_t which, double* out_per_second) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        arg(which >= 0 && which <= 3, "which must be 0, 1, 2 or 3");
        arg(out_per_second != nullptr, "out_per_second must not be NULL");
        *out_per_second = which == 3 ? knn5_microbench_run(ctx->num_sms, ctx->stream) : microbench_run(which, ctx->num_sms, ctx->stream);
    });
}

int32_t slideo_b200_synchronize(slideo_b200_ctx* ctx) {
    REQUIRE_CTX(ctx);
    return guarded(ctx, [&] {
        ctx->drain();
        ctx->collect_timings();
    });
}

}  // extern "C"
