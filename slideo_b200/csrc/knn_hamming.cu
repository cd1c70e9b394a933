// knn_hamming.cu -- K8: exact brute-force Hamming k-NN (k <= 32) of 256-bit descriptors, K9: ratio vote.
//
// Replaces FlannMatcher::knn_match (crates/matching-opencv/src/flann.rs:73-89, called with k = 30 at
// crates/matching-opencv/src/lib.rs:266) with the exact semantics LSH approximates == cv2.BFMatcher(NORM_HAMMING)
// .knnMatch: the k smallest by (distance, pooled index), and the vote loop lib.rs:268-282.
//
// Design (integer-pipe bound, not HBM bound: the pool is <= 48 MB and lives in the 126 MB L2).  The ncu capture of the
// first version (profiles/k8_r1_summary.md) showed the ALU pipe 84 % busy at 24.8 ALU ops per pair; this version
// spends 13 LOP3 + 1 ISETP on the ALU pipe, 4 POPC on the XU pipe and 3 IMAD on the FMA pipe per pair:
//   * every thread owns QR query descriptors in registers for a whole pass over its share of the pool;
//   * the pool is pre-expanded once to 48 B rows {w0..w7, w0^w1^w2, w3^w4^w5, w0^..^w6, 0} and streamed through shared
//     memory in 12 KB chunks by 1-D TMA bulk copies (cp.async.bulk + mbarrier, 3 stages); all lanes read the same
//     row (LDS.128 broadcast);
//   * distance = popcount of the 8 xor-ed words through a carry-save adder tree whose sum bits come straight from the
//     pre-xor-ed words (s1 = q012^p012, s2 = q345^p345, s3 = q0..6^p0..6) and whose carries use the
//     "two inputs + sum" form, so x2, x5, x6 are never materialised;
//   * selection is exact: a pair survives iff dist < the distance of the thread's running k-th best (strict '<' is
//     exact because the pool is scanned in increasing index order, like the oracle's insertion); survivors are rare,
//     a warp vote per pooled row skips the append code entirely when no lane has one; appended keys
//     (dist << 23 | index) go to a 64-slot per-query buffer (L2-resident scratch) that a warp-cooperative 64-key
//     bitonic sort cuts back to the k best whenever it fills up; the final sort emits rows in oracle order;
//   * work is split stream-K style: the (query tile x pool chunk) units are divided evenly over the persistent grid,
//     so every CTA does the same amount of work whatever the shape; tiles covered by several CTAs go through
//     partial rows + a merge kernel;
//   * K9 (vote, lib.rs:270-282) is fused into the final sort: lane m holds neighbour m of the row.
#include <stdlib.h>

#include "common.cuh"

namespace slideo {

namespace {

constexpr int KNN_THREADS = 128;
constexpr int KNN_QR_MAX = 8;                    // queries per thread: 4 or 8 (template parameter of K8)
constexpr int KNN_SLOTS = 64;                    // candidate keys per query
constexpr int KNN_CHUNK = 256;                   // pooled descriptors per smem stage (12 KB of 48 B rows)
constexpr int KNN_ROW_U4 = 3;                    // uint4 per expanded pool row
constexpr int KNN_STAGES = 3;
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
// integer multiply-add pinned to the FMA pipe (IMAD), also for multiplier 1
__device__ __forceinline__ int imad(int a, int b, int c) {
    int r;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
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
    const uint4* pool;    // expanded rows, KNN_ROW_U4 uint4 each
    uint32_t* keys_out;   // [nq][k] -- may be null
    uint32_t* partial;    // [n_tiles][max_seg][tile][k]
    uint32_t* scratch;    // [grid][tile][KNN_SLOTS]
    int nq, nt, k, n_tiles, n_chunks, max_seg, tile;
    long long total_units;  // n_tiles * n_chunks
    VoteArgs vote;          // vote.votes == nullptr -> no fused vote
};

// stream-K bookkeeping: CTA b owns units [U*b/G, U*(b+1)/G); cta_of(u) = the CTA that owns unit u
__device__ __host__ __forceinline__ long long unit_begin(long long U, int G, int b) { return U * b / G; }
__device__ __host__ __forceinline__ int cta_of(long long U, int G, long long u) { return (int)(((u + 1) * G + U - 1) / U - 1); }

// Row emission shared by K8 (single-segment tiles) and the merge kernel: lane m < k holds neighbour m (sorted).
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

// pool rows 32 B -> 48 B {w0..w7, w0^w1^w2, w3^w4^w5, w0^..^w6, 0}
__global__ void __launch_bounds__(256) pool_expand_kernel(const uint4* __restrict__ src, int nt, uint4* __restrict__ dst) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nt) return;
    const uint4 a = __ldg(src + 2 * (size_t)j), b = __ldg(src + 2 * (size_t)j + 1);
    const uint32_t p012 = a.x ^ a.y ^ a.z, p345 = a.w ^ b.x ^ b.y;
    dst[3 * (size_t)j] = a;
    dst[3 * (size_t)j + 1] = b;
    dst[3 * (size_t)j + 2] = make_uint4(p012, p345, p012 ^ p345 ^ b.z, 0u);
}

template <int KNN_QR>
__global__ void __launch_bounds__(KNN_THREADS, KNN_CTAS_PER_SM) knn_hamming_kernel(const KnnParams P) {
    constexpr int KNN_TILE = KNN_THREADS * KNN_QR;
    extern __shared__ __align__(128) uint8_t s_raw[];
    uint4* s_pool = reinterpret_cast<uint4*>(s_raw);                                   // [STAGES][CHUNK * ROW_U4]
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_raw + (size_t)KNN_STAGES * KNN_CHUNK * KNN_ROW_U4 * 16);
    int* s_cnt = reinterpret_cast<int*>(s_full + KNN_STAGES);                          // [KNN_TILE] candidate counts (compaction, final sort)
    uint32_t* s_taud = reinterpret_cast<uint32_t*>(s_cnt + KNN_TILE);                  // [KNN_TILE] thresholds handed back by the compaction

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < KNN_STAGES; ++s) mbar_init(&s_full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    uint32_t* my_scratch = P.scratch + (size_t)blockIdx.x * KNN_TILE * KNN_SLOTS;
    uint32_t gchunk = 0;  // chunks consumed so far by this CTA (stage = gchunk % STAGES, parity from gchunk / STAGES)
    const int G = gridDim.x;
    long long u = unit_begin(P.total_units, G, blockIdx.x);
    const long long u_end = unit_begin(P.total_units, G, blockIdx.x + 1);

    while (u < u_end) {
        const int tile = (int)(u / P.n_chunks);
        const int c0 = (int)(u - (long long)tile * P.n_chunks);
        const int c1 = (int)min((long long)P.n_chunks, c0 + (u_end - u));
        const int n_chunks = c1 - c0;
        u += n_chunks;
        const int qbase = tile * KNN_TILE;
        const int t0 = c0 * KNN_CHUNK;
        const int t1 = min(P.nt, c1 * KNN_CHUNK);
        const long long tile_u0 = (long long)tile * P.n_chunks;
        const int first_cta = cta_of(P.total_units, G, tile_u0);
        const int n_seg = cta_of(P.total_units, G, tile_u0 + P.n_chunks - 1) - first_cta + 1;
        const int seg = blockIdx.x - first_cta;

        // producer prologue
        if (tid == 0) {
            for (int c = 0; c < min(n_chunks, KNN_STAGES); ++c) {
                const int s = (gchunk + c) % KNN_STAGES;
                const int n = min(KNN_CHUNK, t1 - (t0 + c * KNN_CHUNK));
                mbar_expect_tx(&s_full[s], (uint32_t)n * 48u);
                bulk_g2s(s_pool + (size_t)s * KNN_CHUNK * KNN_ROW_U4, P.pool + (size_t)(t0 + c * KNN_CHUNK) * KNN_ROW_U4, (uint32_t)n * 48u,
                         &s_full[s]);
            }
        }

        // this thread's queries: words 0,1,3,4,7 individually + the three pre-xor-ed sums
        uint32_t q0[KNN_QR], q1[KNN_QR], q3[KNN_QR], q4[KNN_QR], q7[KNN_QR], q012[KNN_QR], q345[KNN_QR], q06[KNN_QR];
        uint32_t taud[KNN_QR];   // distance of the running k-th best (strict '<' admits a pair)
        int ntaud2[KNN_QR];      // -2 * taud, the addend of the margin chain
        int cnt[KNN_QR];
#pragma unroll
        for (int i = 0; i < KNN_QR; ++i) {
            const int q = qbase + tid + i * KNN_THREADS;
            uint4 a = make_uint4(0, 0, 0, 0), b = a;
            if (q < P.nq) {
                a = __ldg(P.q + (size_t)q * 2);
                b = __ldg(P.q + (size_t)q * 2 + 1);
            }
            q0[i] = a.x; q1[i] = a.y; q3[i] = a.w; q4[i] = b.x; q7[i] = b.w;
            q012[i] = a.x ^ a.y ^ a.z;
            q345[i] = a.w ^ b.x ^ b.y;
            q06[i] = q012[i] ^ q345[i] ^ b.z;
            taud[i] = q < P.nq ? 512u : 0u;  // dummy queries never admit anything
            ntaud2[i] = -2 * (int)taud[i];
            cnt[i] = 0;
        }

        // margins of this thread's QR queries against one expanded pool row {a, b, e}: margin = 2 * (distance - taud), negative
        // iff the pair beats the running k-th distance.  The sums run as integer multiply-adds (FMA pipe): the ALU pipe is
        // the bottleneck and keeps only the 13 logic ops of the carry-save tree per pair.
        auto margins = [&](const uint4& a, const uint4& b, const uint4& e, int (&mg)[KNN_QR]) {
#pragma unroll
            for (int i = 0; i < KNN_QR; ++i) {
                const uint32_t x0 = q0[i] ^ a.x, x1 = q1[i] ^ a.y;
                const uint32_t s1 = q012[i] ^ e.x;                  // x0 ^ x1 ^ x2
                const uint32_t c1 = lop3<LUT_CARRY>(x0, x1, s1);    // maj(x0, x1, x2)
                const uint32_t x3 = q3[i] ^ a.w, x4 = q4[i] ^ b.x;
                const uint32_t s2 = q345[i] ^ e.y;                  // x3 ^ x4 ^ x5
                const uint32_t c2 = lop3<LUT_CARRY>(x3, x4, s2);    // maj(x3, x4, x5)
                const uint32_t s3 = q06[i] ^ e.z;                   // s1 ^ s2 ^ x6            (weight 1)
                const uint32_t c3 = lop3<LUT_CARRY>(s1, s2, s3);    // maj(s1, s2, x6)         (weight 2)
                const uint32_t x7 = q7[i] ^ b.w;                    //                         (weight 1)
                const uint32_t s5 = lop3<LUT_XOR3>(c1, c2, c3);     //                         (weight 2)
                const uint32_t c5 = lop3<LUT_MAJ>(c1, c2, c3);      //                         (weight 4)
                // 2 * (distance - taud): every multiplier differs from 1, so ptxas keeps the chain on the FMA pipe (IMAD)
                int m = imad(__popc(x7), 2, ntaud2[i]);
                m = imad(__popc(s3), 2, m);
                m = imad(__popc(s5), 4, m);
                mg[i] = imad(__popc(c5), 8, m);
            }
        };

        // warp-cooperative compaction of every candidate buffer of this warp that is (nearly) full: sort its keys,
        // keep the k best, tighten the owner's threshold.  State goes through shared memory so that ONE copy of the
        // sort serves all QR buffers of a lane.
        auto compact_all = [&]() {
#pragma unroll
            for (int i = 0; i < KNN_QR; ++i) s_cnt[tid + i * KNN_THREADS] = cnt[i];
            __syncwarp();
#pragma unroll 1
            for (int i = 0; i < KNN_QR; ++i) {
                const int base_row = warp * 32 + i * KNN_THREADS;
                unsigned need = __ballot_sync(FULL, s_cnt[base_row + lane] >= KNN_SLOTS - 1);
                while (need) {
                    const int L = __ffs(need) - 1;
                    need &= need - 1;
                    uint32_t* buf = my_scratch + (size_t)(base_row + L) * KNN_SLOTS;
                    const int n = s_cnt[base_row + L];
                    uint32_t k0 = lane < n ? __ldcg(buf + lane) : KEY_EMPTY;
                    uint32_t k1 = lane + 32 < n ? __ldcg(buf + lane + 32) : KEY_EMPTY;
                    warp_sort64(k0, k1, lane);
                    if (lane < P.k) buf[lane] = k0;
                    const uint32_t kth = __shfl_sync(FULL, k0, P.k - 1);
                    if (lane == L) {
                        s_taud[base_row + L] = kth == KEY_EMPTY ? 512u : (kth >> KEY_IDX_BITS);
                        s_cnt[base_row + L] = min(n, P.k);
                    }
                }
                __syncwarp();
            }
#pragma unroll
            for (int i = 0; i < KNN_QR; ++i) {
                if (cnt[i] >= KNN_SLOTS - 1) {
                    cnt[i] = s_cnt[tid + i * KNN_THREADS];
                    taud[i] = s_taud[tid + i * KNN_THREADS];
                    ntaud2[i] = -2 * (int)taud[i];
                }
            }
        };

        for (int c = 0; c < n_chunks; ++c, ++gchunk) {
            const int s = gchunk % KNN_STAGES;
            const int n = min(KNN_CHUNK, t1 - (t0 + c * KNN_CHUNK));
            mbar_wait(&s_full[s], (gchunk / KNN_STAGES) & 1);
            const uint4* sp = s_pool + (size_t)s * KNN_CHUNK * KNN_ROW_U4;
            const uint32_t gbase = (uint32_t)(t0 + c * KNN_CHUNK);

            for (int j0 = 0; j0 < n; j0 += 32) {
                const int nb = min(32, n - j0);
                // fast path: 32 rows, no selection work at all -- a lane only remembers WHICH rows held a survivor
                uint32_t mask = 0;
#pragma unroll 2
                for (int r = 0; r < nb; ++r) {
                    const uint4* row = sp + 3 * (j0 + r);
                    int mg[KNN_QR];
                    margins(row[0], row[1], row[2], mg);
                    int any = mg[0];
#pragma unroll
                    for (int i = 1; i < KNN_QR; ++i) any |= mg[i];   // sign bit set iff some margin is negative
                    if (any < 0) mask |= 1u << r;
                }
                // slow path (rare after the first few hundred rows of a segment): every lane revisits its own flagged
                // rows, lowest first, and appends the survivors; buffers are compacted cooperatively when one fills up
                while (__any_sync(FULL, mask != 0)) {
                    bool full = false;
                    if (mask != 0) {
                        const int r = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const uint4* row = sp + 3 * (j0 + r);
                        int mg[KNN_QR];
                        margins(row[0], row[1], row[2], mg);
                        const uint32_t gidx = gbase + (uint32_t)(j0 + r);
#pragma unroll
                        for (int i = 0; i < KNN_QR; ++i) {
                            if (mg[i] < 0) {
                                const uint32_t dist = (uint32_t)((mg[i] >> 1) + (int)taud[i]);
                                my_scratch[(size_t)(tid + i * KNN_THREADS) * KNN_SLOTS + cnt[i]] = (dist << KEY_IDX_BITS) | gidx;
                                ++cnt[i];
                                full |= cnt[i] >= KNN_SLOTS - 1;
                            }
                        }
                    }
                    if (__any_sync(FULL, full)) compact_all();
                }
            }

            __syncthreads();  // every warp is done with stage s
            if (tid == 0 && c + KNN_STAGES < n_chunks) {
                const int cn = c + KNN_STAGES;
                const int nn = min(KNN_CHUNK, t1 - (t0 + cn * KNN_CHUNK));
                mbar_expect_tx(&s_full[s], (uint32_t)nn * 48u);
                bulk_g2s(s_pool + (size_t)s * KNN_CHUNK * KNN_ROW_U4, P.pool + (size_t)(t0 + cn * KNN_CHUNK) * KNN_ROW_U4, (uint32_t)nn * 48u,
                         &s_full[s]);
            }
        }

        // final sort + emission: QR x 32 rows per warp, one row at a time, lane m <- neighbour m
#pragma unroll
        for (int i = 0; i < KNN_QR; ++i) s_cnt[tid + i * KNN_THREADS] = cnt[i];
        __syncwarp();
#pragma unroll 1
        for (int i = 0; i < KNN_QR; ++i) {
#pragma unroll 1
            for (int L = 0; L < 32; ++L) {
                const int row = warp * 32 + L + i * KNN_THREADS;
                const int q = qbase + row;
                if (q >= P.nq) break;  // warp-uniform
                const uint32_t* buf = my_scratch + (size_t)row * KNN_SLOTS;
                const int n = s_cnt[row];
                uint32_t k0 = lane < n ? __ldcg(buf + lane) : KEY_EMPTY;
                uint32_t k1 = lane + 32 < n ? __ldcg(buf + lane + 32) : KEY_EMPTY;
                warp_sort64(k0, k1, lane);
                if (n_seg == 1) {
                    emit_row(k0, lane, q, P.k, P.keys_out, P.vote);
                } else if (lane < P.k) {
                    P.partial[(((size_t)tile * P.max_seg + seg) * KNN_TILE + row) * P.k + lane] = k0;
                }
            }
        }
        __syncthreads();  // scratch + smem stages are reused by the next segment
    }
}

// merge of the per-segment partial rows of multi-segment tiles: one warp per query
__global__ void __launch_bounds__(128) knn_merge_kernel(const uint32_t* __restrict__ partial, int nq, int k, int n_chunks, int max_seg,
                                                        long long total_units, int G, int KNN_TILE, uint32_t* keys_out, const VoteArgs vote) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const int tile = q / KNN_TILE, row = q - tile * KNN_TILE;
    const long long tile_u0 = (long long)tile * n_chunks;
    const int n_seg = cta_of(total_units, G, tile_u0 + n_chunks - 1) - cta_of(total_units, G, tile_u0) + 1;
    if (n_seg == 1) return;  // emitted by K8 itself
    uint32_t k0 = KEY_EMPTY;
    for (int s = 0; s < n_seg; ++s) {
        uint32_t k1 = lane < k ? partial[(((size_t)tile * max_seg + s) * KNN_TILE + row) * k + lane] : KEY_EMPTY;
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
    // every chain is lane-dependent: warp-uniform values would be moved to the uniform datapath (UPOPC / ULOP3)
    const uint32_t t0 = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 1u;
    uint32_t a = t0, b = t0 * 40503u + 7u, c = a ^ 0x9E3779B9u, d = b + 0x7F4A7C15u;
    uint32_t e = a + 11u, f = b ^ 0x1234567u, g = c + 5u, h = d ^ 0xABCDEFu;
    if (WHICH == 0) {  // 8 independent LOP3 chains
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a = lop3<LUT_XOR3>(a, b, c); b = lop3<LUT_MAJ>(b, c, d); c = lop3<LUT_CARRY>(c, d, e); d = lop3<LUT_XOR3>(d, e, f);
                e = lop3<LUT_MAJ>(e, f, g); f = lop3<LUT_CARRY>(f, g, h); g = lop3<LUT_XOR3>(g, h, a); h = lop3<LUT_MAJ>(h, a, b);
            }
        }
    } else {  // 8 independent POPC chains (each popc feeds an IMAD on the FMA pipe, so only POPC loads the XU pipe)
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a = __popc(a) * 0x55555u; b = __popc(b) * 0x33333u; c = __popc(c) * 0x77777u; d = __popc(d) * 0x11111u;
                e = __popc(e) * 0x5a5a5u; f = __popc(f) * 0x3c3c3u; g = __popc(g) * 0x69696u; h = __popc(h) * 0x0f0f1u;
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

// which = 2: the K8 inner-loop instruction mix (13 LOP3 + 4 POPC + 3 IMAD + 1 ISETP per pair, QR queries per thread)
// with no memory traffic and no selection -- the achievable ceiling of this formulation in pairs/s.
__global__ void __launch_bounds__(KNN_THREADS, KNN_CTAS_PER_SM) microbench_mix_kernel(uint32_t* out, int iters) {
    constexpr int KNN_QR = KNN_QR_MAX;
    uint32_t q0[KNN_QR], q1[KNN_QR], q3[KNN_QR], q4[KNN_QR], q7[KNN_QR], q012[KNN_QR], q345[KNN_QR], q06[KNN_QR], best[KNN_QR];
    uint32_t seed = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 1u;
#pragma unroll
    for (int i = 0; i < KNN_QR; ++i) {
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { seed = seed * 1664525u + 1013904223u; w[j] = seed; }
        q0[i] = w[0]; q1[i] = w[1]; q3[i] = w[3]; q4[i] = w[4]; q7[i] = w[7];
        q012[i] = w[0] ^ w[1] ^ w[2];
        q345[i] = w[3] ^ w[4] ^ w[5];
        q06[i] = q012[i] ^ q345[i] ^ w[6];
        best[i] = (seed >> 27) + 1u;   // 1..32: unknown to the compiler, practically never beaten by a random 256-bit distance
    }
    uint4 a = make_uint4(seed, seed * 3u, seed * 5u, seed * 7u), b = make_uint4(seed * 11u, seed * 13u, seed * 17u, seed * 19u);
    uint4 e = make_uint4(seed * 23u, seed * 29u, seed * 31u, 0u);
    uint32_t hits = 0;
    for (int it = 0; it < iters; ++it) {
        int any = 0;
#pragma unroll
        for (int i = 0; i < KNN_QR; ++i) {
            const uint32_t x0 = q0[i] ^ a.x, x1 = q1[i] ^ a.y;
            const uint32_t s1 = q012[i] ^ e.x;
            const uint32_t c1 = lop3<LUT_CARRY>(x0, x1, s1);
            const uint32_t x3 = q3[i] ^ a.w, x4 = q4[i] ^ b.x;
            const uint32_t s2 = q345[i] ^ e.y;
            const uint32_t c2 = lop3<LUT_CARRY>(x3, x4, s2);
            const uint32_t s3 = q06[i] ^ e.z;
            const uint32_t c3 = lop3<LUT_CARRY>(s1, s2, s3);
            const uint32_t x7 = q7[i] ^ b.w;
            const uint32_t s5 = lop3<LUT_XOR3>(c1, c2, c3);
            const uint32_t c5 = lop3<LUT_MAJ>(c1, c2, c3);
            int m = imad(__popc(x7), 2, -2 * (int)best[i]);
            m = imad(__popc(s3), 2, m);
            m = imad(__popc(s5), 4, m);
            any |= imad(__popc(c5), 8, m);
        }
        const bool hit = any < 0;
        if (__any_sync(FULL, hit)) ++hits;   // data dependent, practically never taken after the first iterations
        if (hits > 1000000u) best[0] += 1u;
        // next "pooled row": a cheap dependent update (IMAD-class ops, amortised over QR pairs)
        a.x = a.x * 1664525u + hits;
        b.w = b.w * 22695477u + a.x;
        e.x = e.x * 69069u + b.w;
        uint32_t t = a.x; a.x = a.y; a.y = a.z; a.z = a.w; a.w = b.x; b.x = b.y; b.y = b.z; b.z = b.w; b.w = e.y; e.y = e.z; e.z = e.x; e.x = t;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = best[0] ^ best[1] ^ hits;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
KnnPlan knn_hamming_plan(int nq, int nt, int k, int num_sms, int ctas_per_sm_req) {
    KnnPlan p;
    p.nq = nq; p.nt = nt; p.k = k;
    // queries per thread: 8 amortises the pooled-row loads best, 4 gives twice as many tiles -> fewer, longer stream-K
    // segments per tile (every segment re-warms its selection thresholds from scratch)
    int ctas_per_sm = ctas_per_sm_req >= 1 && ctas_per_sm_req <= KNN_CTAS_PER_SM ? ctas_per_sm_req : KNN_CTAS_PER_SM;
    const long long target0 = (long long)num_sms * ctas_per_sm;
    p.qr = (long long)cdiv(nq, KNN_THREADS * 8) * 2 >= target0 * 3 ? 8 : 4;
    const int KNN_TILE = KNN_THREADS * p.qr;
    p.tile = KNN_TILE;
    p.n_tiles = cdiv(nq, KNN_TILE);
    p.n_chunks = cdiv(nt > 0 ? nt : 1, KNN_CHUNK);
    p.total_units = (long long)p.n_tiles * p.n_chunks;
    const long long target = target0;
    p.grid = (int)(p.total_units < target ? p.total_units : target);
    if (p.grid < 1) p.grid = 1;
    p.max_seg = 1;
    for (int t = 0; t < p.n_tiles; ++t) {
        const long long u0 = (long long)t * p.n_chunks;
        const int n_seg = cta_of(p.total_units, p.grid, u0 + p.n_chunks - 1) - cta_of(p.total_units, p.grid, u0) + 1;
        if (n_seg > p.max_seg) p.max_seg = n_seg;
    }
    p.scratch_bytes = (size_t)p.grid * KNN_TILE * KNN_SLOTS * sizeof(uint32_t);
    p.partial_bytes = p.max_seg > 1 ? (size_t)p.n_tiles * p.max_seg * KNN_TILE * k * sizeof(uint32_t) : 0;
    p.pool_bytes = (size_t)(nt > 0 ? nt : 1) * KNN_ROW_U4 * 16;
    return p;
}

size_t knn_pool_expanded_bytes(int nt) { return (size_t)(nt > 0 ? nt : 1) * KNN_ROW_U4 * 16; }

void knn_pool_expand_launch(const void* d_pool32, int nt, void* d_pool48, cudaStream_t stream) {
    if (nt <= 0) return;
    pool_expand_kernel<<<cdiv(nt, 256), 256, 0, stream>>>((const uint4*)d_pool32, nt, (uint4*)d_pool48);
    SLIDEO_CUDA(cudaGetLastError());
}

void knn_hamming_launch(const KnnPlan& plan, const void* d_q, const void* d_pool48, uint32_t* d_keys_out,
                        uint32_t* d_scratch, uint32_t* d_partial, const VoteArgs* vote, cudaStream_t stream,
                        int* launches) {
    if (plan.nq <= 0) return;
    KnnParams P;
    P.q = (const uint4*)d_q;
    P.pool = (const uint4*)d_pool48;
    P.keys_out = d_keys_out;
    P.partial = d_partial;
    P.scratch = d_scratch;
    P.nq = plan.nq; P.nt = plan.nt; P.k = plan.k;
    P.n_tiles = plan.n_tiles; P.n_chunks = plan.n_chunks; P.max_seg = plan.max_seg; P.total_units = plan.total_units;
    P.tile = plan.tile;
    if (vote) P.vote = *vote;
    else P.vote = VoteArgs{nullptr, nullptr, nullptr, 0, 0.f};
    const size_t smem = (size_t)KNN_STAGES * KNN_CHUNK * KNN_ROW_U4 * 16 + KNN_STAGES * sizeof(uint64_t) + 2 * plan.tile * sizeof(int);
    {   // function attributes are per device: set on every launch (cheap) so that ctxs on several GPUs of one process all get them
        const int max_smem = (int)((size_t)KNN_STAGES * KNN_CHUNK * KNN_ROW_U4 * 16 + KNN_STAGES * sizeof(uint64_t) + 2 * KNN_THREADS * KNN_QR_MAX * sizeof(int));
        SLIDEO_CUDA(cudaFuncSetAttribute(knn_hamming_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        SLIDEO_CUDA(cudaFuncSetAttribute(knn_hamming_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        SLIDEO_CUDA(cudaFuncSetAttribute(knn_hamming_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        SLIDEO_CUDA(cudaFuncSetAttribute(knn_hamming_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    }
    if (plan.qr == 8) knn_hamming_kernel<8><<<plan.grid, KNN_THREADS, smem, stream>>>(P);
    else knn_hamming_kernel<4><<<plan.grid, KNN_THREADS, smem, stream>>>(P);
    SLIDEO_CUDA(cudaGetLastError());
    if (launches) ++*launches;
    if (plan.max_seg > 1) {
        knn_merge_kernel<<<cdiv(plan.nq, 4), 128, 0, stream>>>(d_partial, plan.nq, plan.k, plan.n_chunks, plan.max_seg, plan.total_units,
                                                              plan.grid, plan.tile, d_keys_out, P.vote);
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
    const double ops = which == 2 ? (double)blocks * threads * (double)iters * KNN_QR_MAX : (double)blocks * threads * (double)iters * 64.0;
    return ops / (best * 1e-3);
}

}  // namespace slideo
