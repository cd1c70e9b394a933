// knn_hamming.cu -- K8: exact brute-force Hamming k-NN (k <= 32) of 256-bit descriptors, K9: ratio vote.
//
// Replaces FlannMatcher::knn_match (crates/matching-opencv/src/flann.rs:73-89, called with k = 30 at
// crates/matching-opencv/src/lib.rs:266) with the exact semantics LSH approximates == cv2.BFMatcher(NORM_HAMMING)
// .knnMatch: the k smallest by (distance, pooled index), and the vote loop lib.rs:268-282.
//
// Design (integer-pipe bound, not HBM bound: the pool is <= 32 MB and lives in the 126 MB L2):
//   * every thread owns QR=4 query descriptors in registers for a whole pass over (a split of) the pool;
//   * the pool is streamed through shared memory in 8 KB chunks by 1-D TMA bulk copies (cp.async.bulk +
//     mbarrier, 3 stages); all lanes read the same pooled descriptor (LDS.128 broadcast);
//   * distance = popcount of 8 xor-ed words through a carry-save adder tree: 14 LOP3 + 4 POPC per pair instead of
//     8 LOP3 + 8 POPC, which balances the ALU pipe against the quarter-rate POPC pipe;
//   * selection is exact: key = dist << 23 | index is compared against the thread's running k-th best key; the
//     rare survivors are appended to a 64-slot per-query candidate buffer (L2-resident scratch) that is compacted
//     by a warp-cooperative 64-key bitonic sort whenever it fills up; the final sort emits rows in oracle order;
//   * small query counts split the pool across CTAs (partial rows + a merge kernel);
//   * K9 (vote, lib.rs:270-282) is fused into the final sort: lane m holds neighbour m of the row.
#include "common.cuh"

namespace slideo {

namespace {

constexpr int KNN_THREADS = 128;
constexpr int KNN_QR = 4;
constexpr int KNN_TILE = KNN_THREADS * KNN_QR;  // queries per work item
constexpr int KNN_SLOTS = 64;                   // candidate keys per query
constexpr int KNN_CHUNK = 256;                  // pooled descriptors per smem stage (8 KB)
constexpr int KNN_STAGES = 3;
constexpr int KNN_GROUP = 8;                    // pooled descriptors between overflow checks
constexpr int KNN_CTAS_PER_SM = 4;
constexpr unsigned FULL = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return r;
}
constexpr int LUT_XOR3 = 0x96;  // a ^ b ^ c
constexpr int LUT_MAJ = 0xE8;   // majority(a, b, c)
constexpr int LUT_CARRY = 0xD4; // majority(a, b, a ^ b ^ c): carry of a full adder given two inputs and the sum

// 64-key ascending bitonic sort across a warp: position p = lane (k0) and 32 + lane (k1).
__device__ __forceinline__ void warp_sort64(uint32_t& k0, uint32_t& k1, int lane) {
#pragma unroll
    for (int size = 2; size <= 64; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride == 32) {
                uint32_t lo = min(k0, k1), hi = max(k0, k1);
                k0 = lo;
                k1 = hi;
            } else {
                uint32_t o0 = __shfl_xor_sync(FULL, k0, stride), o1 = __shfl_xor_sync(FULL, k1, stride);
                bool lower = (lane & stride) == 0;
                bool up0 = size == 64 ? true : (size == 32 ? true : (lane & size) == 0);
                bool up1 = size == 64 ? true : (size == 32 ? false : (lane & size) == 0);
                k0 = (lower == up0) ? min(k0, o0) : max(k0, o0);
                k1 = (lower == up1) ? min(k1, o1) : max(k1, o1);
            }
        }
    }
}

struct KnnParams {
    const uint4* q;
    const uint4* pool;
    uint32_t* keys_out;   // [nq][k] (n_splits == 1) -- may be null
    uint32_t* partial;    // [nq][n_splits][k] (n_splits > 1)
    uint32_t* scratch;    // [grid][KNN_TILE][KNN_SLOTS]
    int nq, nt, k, n_tiles, n_splits, split_len;
    VoteArgs vote;        // vote.votes == nullptr -> no fused vote
};

// Row emission shared by K8 (n_splits == 1) and the merge kernel: lane m < k holds neighbour m (sorted).
__device__ __forceinline__ void emit_row(uint32_t key, int lane, int q, int k, uint32_t* keys_out, const VoteArgs& v) {
    if (keys_out != nullptr && lane < k) keys_out[(size_t)q * k + lane] = key;
    if (v.votes != nullptr) {
        uint32_t best = __shfl_sync(FULL, key, 0);
        if (lane < k && key != KEY_EMPTY) {
            // lib.rs:275  `dmatch.distance < best.distance * 1.05`  (f32; best == 0 -> no vote)
            float d = (float)(key >> KEY_IDX_BITS), b = (float)(best >> KEY_IDX_BITS);
            if (d < __fmul_rn(b, v.ratio)) {
                int page = v.page_of[key & KEY_IDX_MASK];
                atomicAdd(&v.votes[(size_t)v.q_frame[q] * v.n_pages + page], 1);
            }
        }
    }
}

__global__ void __launch_bounds__(KNN_THREADS, KNN_CTAS_PER_SM) knn_hamming_kernel(const KnnParams P) {
    __shared__ __align__(128) uint4 s_pool[KNN_STAGES][KNN_CHUNK * 2];
    __shared__ __align__(8) uint64_t s_full[KNN_STAGES];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < KNN_STAGES; ++s) mbar_init(&s_full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    uint32_t* my_scratch = P.scratch + (size_t)blockIdx.x * KNN_TILE * KNN_SLOTS;
    uint32_t gchunk = 0;  // chunks consumed so far by this CTA (stage = gchunk % STAGES, parity from gchunk / STAGES)
    const int n_items = P.n_tiles * P.n_splits;

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int tile = item / P.n_splits, split = item - tile * P.n_splits;
        const int qbase = tile * KNN_TILE;
        const int t0 = split * P.split_len;
        const int t1 = min(P.nt, t0 + P.split_len);
        const int n_chunks = (t1 - t0 + KNN_CHUNK - 1) / KNN_CHUNK;

        // producer prologue
        if (tid == 0) {
            for (int c = 0; c < min(n_chunks, KNN_STAGES); ++c) {
                const int s = (gchunk + c) % KNN_STAGES;
                const int n = min(KNN_CHUNK, t1 - (t0 + c * KNN_CHUNK));
                mbar_expect_tx(&s_full[s], (uint32_t)n * 32u);
                bulk_g2s(&s_pool[s][0], P.pool + (size_t)(t0 + c * KNN_CHUNK) * 2, (uint32_t)n * 32u, &s_full[s]);
            }
        }

        // this thread's queries
        uint32_t qw[KNN_QR][8], q012[KNN_QR], q345[KNN_QR], tau[KNN_QR];
        int cnt[KNN_QR];
#pragma unroll
        for (int i = 0; i < KNN_QR; ++i) {
            const int q = qbase + tid + i * KNN_THREADS;
            uint4 a = make_uint4(0, 0, 0, 0), b = a;
            if (q < P.nq) {
                a = __ldg(P.q + (size_t)q * 2);
                b = __ldg(P.q + (size_t)q * 2 + 1);
            }
            qw[i][0] = a.x; qw[i][1] = a.y; qw[i][2] = a.z; qw[i][3] = a.w;
            qw[i][4] = b.x; qw[i][5] = b.y; qw[i][6] = b.z; qw[i][7] = b.w;
            q012[i] = a.x ^ a.y ^ a.z;
            q345[i] = a.w ^ b.x ^ b.y;
            tau[i] = q < P.nq ? KEY_EMPTY : 0u;  // dummy queries never push
            cnt[i] = 0;
        }

        auto compact = [&](int i) {
            // warp-cooperative: every lane whose buffer i is nearly full gets it sorted and cut to the k best
            unsigned need = __ballot_sync(FULL, cnt[i] > KNN_SLOTS - KNN_GROUP);
            while (need) {
                const int L = __ffs(need) - 1;
                need &= need - 1;
                uint32_t* buf = my_scratch + (size_t)(warp * 32 + L + i * KNN_THREADS) * KNN_SLOTS;
                const int n = __shfl_sync(FULL, cnt[i], L);
                uint32_t k0 = lane < n ? __ldcg(buf + lane) : KEY_EMPTY;
                uint32_t k1 = lane + 32 < n ? __ldcg(buf + lane + 32) : KEY_EMPTY;
                warp_sort64(k0, k1, lane);
                if (lane < P.k) buf[lane] = k0;
                const uint32_t kth = __shfl_sync(FULL, k0, P.k - 1);
                if (lane == L) {
                    tau[i] = kth;
                    cnt[i] = min(n, P.k);
                }
            }
            __syncwarp();
        };

        for (int c = 0; c < n_chunks; ++c, ++gchunk) {
            const int s = gchunk % KNN_STAGES;
            const int n = min(KNN_CHUNK, t1 - (t0 + c * KNN_CHUNK));
            mbar_wait(&s_full[s], (gchunk / KNN_STAGES) & 1);
            const uint4* sp = &s_pool[s][0];
            const uint32_t gbase = (uint32_t)(t0 + c * KNN_CHUNK);

            auto process = [&](int j) {
                const uint4 a = sp[2 * j], b = sp[2 * j + 1];
                const uint32_t p012 = a.x ^ a.y ^ a.z, p345 = a.w ^ b.x ^ b.y;
                const uint32_t gidx = gbase + (uint32_t)j;
#pragma unroll
                for (int i = 0; i < KNN_QR; ++i) {
                    const uint32_t x0 = qw[i][0] ^ a.x, x1 = qw[i][1] ^ a.y;
                    const uint32_t s1 = q012[i] ^ p012;                 // x0 ^ x1 ^ x2
                    const uint32_t c1 = lop3<LUT_CARRY>(x0, x1, s1);    // maj(x0, x1, x2)
                    const uint32_t x3 = qw[i][3] ^ a.w, x4 = qw[i][4] ^ b.x;
                    const uint32_t s2 = q345[i] ^ p345;                 // x3 ^ x4 ^ x5
                    const uint32_t c2 = lop3<LUT_CARRY>(x3, x4, s2);
                    const uint32_t x6 = qw[i][6] ^ b.z, x7 = qw[i][7] ^ b.w;
                    const uint32_t s3 = lop3<LUT_XOR3>(s1, s2, x6);     // weight 1
                    const uint32_t c3 = lop3<LUT_MAJ>(s1, s2, x6);      // weight 2
                    const uint32_t s5 = lop3<LUT_XOR3>(c1, c2, c3);     // weight 2
                    const uint32_t c5 = lop3<LUT_MAJ>(c1, c2, c3);      // weight 4
                    const uint32_t ones = __popc(s3) + __popc(x7);
                    const uint32_t key = (ones << KEY_IDX_BITS) + ((uint32_t)__popc(s5) << (KEY_IDX_BITS + 1)) +
                                         ((uint32_t)__popc(c5) << (KEY_IDX_BITS + 2)) + gidx;
                    if (key < tau[i]) {
                        my_scratch[(size_t)(tid + i * KNN_THREADS) * KNN_SLOTS + cnt[i]] = key;
                        ++cnt[i];
                    }
                }
            };
            auto check = [&]() {
                bool need = false;
#pragma unroll
                for (int i = 0; i < KNN_QR; ++i) need |= cnt[i] > KNN_SLOTS - KNN_GROUP;
                if (__any_sync(FULL, need)) {
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < KNN_QR; ++i) compact(i);
                }
            };

            int j = 0;
            for (; j + KNN_GROUP <= n; j += KNN_GROUP) {
#pragma unroll
                for (int jj = 0; jj < KNN_GROUP; ++jj) process(j + jj);
                check();
            }
            if (j < n) {
                for (; j < n; ++j) process(j);
                check();
            }

            __syncthreads();  // every warp is done with stage s
            if (tid == 0 && c + KNN_STAGES < n_chunks) {
                const int cn = c + KNN_STAGES;
                const int nn = min(KNN_CHUNK, t1 - (t0 + cn * KNN_CHUNK));
                mbar_expect_tx(&s_full[s], (uint32_t)nn * 32u);
                bulk_g2s(&s_pool[s][0], P.pool + (size_t)(t0 + cn * KNN_CHUNK) * 2, (uint32_t)nn * 32u, &s_full[s]);
            }
        }

        // final sort + emission: 4 x 32 rows per warp, one row at a time, lane m <- neighbour m
        __syncwarp();
#pragma unroll
        for (int i = 0; i < KNN_QR; ++i) {
            for (int L = 0; L < 32; ++L) {
                const int q = qbase + warp * 32 + L + i * KNN_THREADS;
                if (q >= P.nq) break;  // warp-uniform
                const uint32_t* buf = my_scratch + (size_t)(warp * 32 + L + i * KNN_THREADS) * KNN_SLOTS;
                const int n = __shfl_sync(FULL, cnt[i], L);
                uint32_t k0 = lane < n ? __ldcg(buf + lane) : KEY_EMPTY;
                uint32_t k1 = lane + 32 < n ? __ldcg(buf + lane + 32) : KEY_EMPTY;
                warp_sort64(k0, k1, lane);
                if (P.n_splits == 1) {
                    emit_row(k0, lane, q, P.k, P.keys_out, P.vote);
                } else if (lane < P.k) {
                    P.partial[((size_t)q * P.n_splits + split) * P.k + lane] = k0;
                }
            }
        }
        __syncthreads();  // scratch + smem stages are reused by the next item
    }
}

// merge of per-split partial rows: one warp per query
__global__ void __launch_bounds__(128) knn_merge_kernel(const uint32_t* __restrict__ partial, int nq, int n_splits, int k,
                                                        uint32_t* keys_out, const VoteArgs vote) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    uint32_t k0 = KEY_EMPTY;
    for (int s = 0; s < n_splits; ++s) {
        uint32_t k1 = lane < k ? partial[((size_t)q * n_splits + s) * k + lane] : KEY_EMPTY;
        warp_sort64(k0, k1, lane);
    }
    emit_row(k0, lane, q, k, keys_out, vote);
}

__global__ void vote_argmax_kernel(const int32_t* __restrict__ votes, int n_frames, int n_pages,
                                   const int32_t* __restrict__ frame_nkp, int32_t* __restrict__ results) {
    const int lane = threadIdx.x & 31;
    const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= n_frames) return;
    int bv = 0, bp = -1;
    for (int p = lane; p < n_pages; p += 32) {
        int v = votes[(size_t)f * n_pages + p];
        if (v > bv) { bv = v; bp = p; }  // increasing p per lane: first max wins
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        int ov = __shfl_xor_sync(FULL, bv, o), op = __shfl_xor_sync(FULL, bp, o);
        if (ov > bv || (ov == bv && ov > 0 && op < bp)) { bv = ov; bp = op; }
    }
    if (lane == 0) {
        results[3 * f] = bp;
        results[3 * f + 1] = bv;
        results[3 * f + 2] = frame_nkp ? frame_nkp[f] : 0;
    }
}

__global__ void keys_to_idx_dist_kernel(const uint32_t* __restrict__ keys, size_t n, int32_t* __restrict__ idx,
                                        int32_t* __restrict__ dist) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t key = keys[i];
    idx[i] = key == KEY_EMPTY ? -1 : (int32_t)(key & KEY_IDX_MASK);
    dist[i] = key == KEY_EMPTY ? -1 : (int32_t)(key >> KEY_IDX_BITS);
}

// ---- integer-pipe micro-benchmarks: the roofline denominators for K8 ------------------------------------------
template <int WHICH>
__global__ void __launch_bounds__(256) microbench_kernel(uint32_t* out, int iters) {
    uint32_t a = threadIdx.x * 2654435761u + 1u, b = blockIdx.x * 40503u + 7u, c = a ^ 0x9E3779B9u, d = b + 0x7F4A7C15u;
    uint32_t e = a + 11u, f = b ^ 0x1234567u, g = c + 5u, h = d ^ 0xABCDEFu;
    if (WHICH == 0) {  // 8 independent LOP3 chains
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a = lop3<LUT_XOR3>(a, b, c); b = lop3<LUT_MAJ>(b, c, d); c = lop3<LUT_CARRY>(c, d, e); d = lop3<LUT_XOR3>(d, e, f);
                e = lop3<LUT_MAJ>(e, f, g); f = lop3<LUT_CARRY>(f, g, h); g = lop3<LUT_XOR3>(g, h, a); h = lop3<LUT_MAJ>(h, a, b);
            }
        }
    } else {  // 8 independent POPC chains (popc feeds an xor to keep the chain alive on another pipe)
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a = __popc(a) + 0x55555u; b = __popc(b) + 0x33333u; c = __popc(c) + 0x77777u; d = __popc(d) + 0x11111u;
                e = __popc(e) + 0x5a5a5u; f = __popc(f) + 0x3c3c3u; g = __popc(g) + 0x69696u; h = __popc(h) + 0x0f0f1u;
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

// which = 2: the K8 inner-loop instruction mix (14 LOP3 + 4 POPC + key assembly + compare per pair, 4 queries per
// thread) with no memory traffic and no selection -- the achievable ceiling of this formulation in pairs/s.
__global__ void __launch_bounds__(KNN_THREADS, KNN_CTAS_PER_SM) microbench_mix_kernel(uint32_t* out, int iters) {
    uint32_t qw[KNN_QR][8], q012[KNN_QR], q345[KNN_QR], best[KNN_QR];
    uint32_t seed = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 1u;
#pragma unroll
    for (int i = 0; i < KNN_QR; ++i) {
#pragma unroll
        for (int w = 0; w < 8; ++w) { seed = seed * 1664525u + 1013904223u; qw[i][w] = seed; }
        q012[i] = qw[i][0] ^ qw[i][1] ^ qw[i][2];
        q345[i] = qw[i][3] ^ qw[i][4] ^ qw[i][5];
        best[i] = KEY_EMPTY;
    }
    uint4 a = make_uint4(seed, seed * 3u, seed * 5u, seed * 7u), b = make_uint4(seed * 11u, seed * 13u, seed * 17u, seed * 19u);
    for (int it = 0; it < iters; ++it) {
        const uint32_t p012 = a.x ^ a.y ^ a.z, p345 = a.w ^ b.x ^ b.y;
#pragma unroll
        for (int i = 0; i < KNN_QR; ++i) {
            const uint32_t x0 = qw[i][0] ^ a.x, x1 = qw[i][1] ^ a.y;
            const uint32_t s1 = q012[i] ^ p012;
            const uint32_t c1 = lop3<LUT_CARRY>(x0, x1, s1);
            const uint32_t x3 = qw[i][3] ^ a.w, x4 = qw[i][4] ^ b.x;
            const uint32_t s2 = q345[i] ^ p345;
            const uint32_t c2 = lop3<LUT_CARRY>(x3, x4, s2);
            const uint32_t x6 = qw[i][6] ^ b.z, x7 = qw[i][7] ^ b.w;
            const uint32_t s3 = lop3<LUT_XOR3>(s1, s2, x6);
            const uint32_t c3 = lop3<LUT_MAJ>(s1, s2, x6);
            const uint32_t s5 = lop3<LUT_XOR3>(c1, c2, c3);
            const uint32_t c5 = lop3<LUT_MAJ>(c1, c2, c3);
            const uint32_t ones = __popc(s3) + __popc(x7);
            const uint32_t key = (ones << KEY_IDX_BITS) + ((uint32_t)__popc(s5) << (KEY_IDX_BITS + 1)) +
                                 ((uint32_t)__popc(c5) << (KEY_IDX_BITS + 2)) + (uint32_t)it;
            if (key < best[i]) best[i] = key;
        }
        // next "pooled descriptor": a cheap dependent update (2 IMAD-class ops per pooled descriptor, amortised over 4 pairs)
        a.x = a.x * 1664525u + best[0];
        b.w = b.w * 22695477u + a.x;
        uint32_t t = a.x; a.x = a.y; a.y = a.z; a.z = a.w; a.w = b.x; b.x = b.y; b.y = b.z; b.z = b.w; b.w = t;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = best[0] ^ best[1] ^ best[2] ^ best[3];
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
KnnPlan knn_hamming_plan(int nq, int nt, int k, int num_sms) {
    KnnPlan p;
    p.nq = nq; p.nt = nt; p.k = k;
    p.n_tiles = cdiv(nq, KNN_TILE);
    const int chunks = cdiv(nt > 0 ? nt : 1, KNN_CHUNK);
    const int target = num_sms * KNN_CTAS_PER_SM;
    // choose the number of pool splits: fill the machine for several waves while keeping splits long enough
    // that the per-item final sort (~ one pass over 350 pooled descriptors) stays a small fraction
    int best_ns = 1;
    double best_eff = -1.0;
    const int max_ns = chunks < 256 ? chunks : 256;
    for (int ns = 1; ns <= max_ns; ++ns) {
        const int len_chunks = cdiv(chunks, ns);
        const int real_ns = cdiv(chunks, len_chunks);
        if (real_ns != ns) continue;
        const long items = (long)p.n_tiles * ns;
        const int grid = (int)(items < target ? items : target);
        const long waves = (items + grid - 1) / grid;
        double eff = (double)items / (double)(waves * target);
        const double len = (double)len_chunks * KNN_CHUNK;
        eff *= len / (len + 350.0);
        if (eff > best_eff * 1.02) { best_eff = eff; best_ns = ns; }
    }
    p.n_splits = best_ns;
    p.split_len = cdiv(chunks, best_ns) * KNN_CHUNK;
    const long items = (long)p.n_tiles * p.n_splits;
    p.grid = (int)(items < target ? items : target);
    if (p.grid < 1) p.grid = 1;
    p.scratch_bytes = (size_t)p.grid * KNN_TILE * KNN_SLOTS * sizeof(uint32_t);
    p.partial_bytes = p.n_splits > 1 ? (size_t)nq * p.n_splits * k * sizeof(uint32_t) : 0;
    return p;
}

void knn_hamming_launch(const KnnPlan& plan, const void* d_q, const void* d_pool, uint32_t* d_keys_out,
                        uint32_t* d_scratch, uint32_t* d_partial, const VoteArgs* vote, cudaStream_t stream,
                        int* launches) {
    if (plan.nq <= 0) return;
    KnnParams P;
    P.q = (const uint4*)d_q;
    P.pool = (const uint4*)d_pool;
    P.keys_out = d_keys_out;
    P.partial = d_partial;
    P.scratch = d_scratch;
    P.nq = plan.nq; P.nt = plan.nt; P.k = plan.k;
    P.n_tiles = plan.n_tiles; P.n_splits = plan.n_splits; P.split_len = plan.split_len;
    if (vote) P.vote = *vote;
    else P.vote = VoteArgs{nullptr, nullptr, nullptr, 0, 0.f};
    knn_hamming_kernel<<<plan.grid, KNN_THREADS, 0, stream>>>(P);
    SLIDEO_CUDA(cudaGetLastError());
    if (launches) ++*launches;
    if (plan.n_splits > 1) {
        knn_merge_kernel<<<cdiv(plan.nq, 4), 128, 0, stream>>>(d_partial, plan.nq, plan.n_splits, plan.k, d_keys_out, P.vote);
        SLIDEO_CUDA(cudaGetLastError());
        if (launches) ++*launches;
    }
}

void vote_argmax_launch(const int32_t* d_votes, int n_frames, int n_pages, const int32_t* d_frame_nkp,
                        int32_t* d_results, cudaStream_t stream) {
    if (n_frames <= 0) return;
    vote_argmax_kernel<<<cdiv(n_frames, 4), 128, 0, stream>>>(d_votes, n_frames, n_pages, d_frame_nkp, d_results);
    SLIDEO_CUDA(cudaGetLastError());
}

void keys_to_idx_dist_launch(const uint32_t* d_keys, size_t n, int32_t* d_idx, int32_t* d_dist, cudaStream_t stream) {
    if (!n) return;
    keys_to_idx_dist_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_keys, n, d_idx, d_dist);
    SLIDEO_CUDA(cudaGetLastError());
}

double microbench_run(int which, int num_sms, cudaStream_t stream) {
    const int blocks = which == 2 ? num_sms * KNN_CTAS_PER_SM : num_sms * 8;
    const int threads = which == 2 ? KNN_THREADS : 256;
    const int iters = which == 2 ? 32768 : 4096;
    uint32_t* d_out = nullptr;
    SLIDEO_CUDA(cudaMalloc(&d_out, (size_t)blocks * threads * 4));
    cudaEvent_t e0, e1;
    SLIDEO_CUDA(cudaEventCreate(&e0));
    SLIDEO_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        SLIDEO_CUDA(cudaEventRecord(e0, stream));
        if (which == 0) microbench_kernel<0><<<blocks, threads, 0, stream>>>(d_out, iters);
        else if (which == 1) microbench_kernel<1><<<blocks, threads, 0, stream>>>(d_out, iters);
        else microbench_mix_kernel<<<blocks, threads, 0, stream>>>(d_out, iters);
        SLIDEO_CUDA(cudaEventRecord(e1, stream));
        SLIDEO_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        SLIDEO_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    // which 0/1: thread-level ops per second; which 2: descriptor pairs per second
    const double ops = which == 2 ? (double)blocks * threads * (double)iters * KNN_QR : (double)blocks * threads * (double)iters * 64.0;
    return ops / (best * 1e-3);
}

}  // namespace slideo
