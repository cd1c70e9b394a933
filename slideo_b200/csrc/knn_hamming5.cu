// knn_hamming5.cu -- K8 v5: exact brute-force Hamming k-NN (k <= 32) of 256-bit descriptors in BIT-SLICED form, K9 vote fused.
//
// Same contract as knn_hamming.cu (K8 v4): replaces FlannMatcher::knn_match (crates/matching-opencv/src/flann.rs:73-89, k = 30 at
// crates/matching-opencv/src/lib.rs:266) with the exact semantics LSH approximates == cv2.BFMatcher(NORM_HAMMING).knnMatch:
// the k smallest by (distance, pooled index), and the vote loop lib.rs:268-282.
//
// Formulation.  v4 computes popcount(q ^ t) per pair: 13 LOP3 + 4 POPC, and the quarter-rate POPC pipe caps it at ~0.9 Tpair/s.
// v5 never popcounts.  The pool is transposed once into "slabs" of 4096 rows: word P[b][j] of a slab holds bit b of the 32 rows
// 32 j .. 32 j + 31.  A warp takes ONE query at a time (warp-uniform), walks the list of the query's set bits and adds the
// words P[b][.] of those bits with a Harley-Seal carry-save tree (2 LOP3 per added word):  c = |q & t| for 128 rows per lane, as
// bit planes.  With A = popc(q), pt = popc(t):  d = A + pt - 2c.  The planes are seeded with pt' = 256 - pt (precomputed per row,
// also bit-sliced), so the tree delivers s = 2c + pt' directly and  d = A + 256 - s.  Queries with more than 128 set bits walk
// the list of their CLEAR bits instead (c' = |~q & t| = pt - c,  d = A - 256 + s'), so a list never exceeds 128 entries.
// The test d < tau (tau = the query's running k-th distance, warp-uniform) is a 10-bit ripple carry of s + (1024 - T) on the
// planes: 10 LOP3 per 32 rows.  Total ~8.3 LOP3 per pair on the ALU pipe for a list of 128 entries, nothing on the XU pipe.
//
// The walk is as long as the list.  Descriptors of slides are far from balanced (the shorter list has ~90 entries on average), so
// the walk stops at the end of the list in whole blocks of 16 entries, and pool rows and queries are first XORed with the pool's
// majority vector (knn5_flip_kernel; distances are invariant, the lists shrink to ~82 entries).  A query then costs what its list
// is long: the queries of a tile are dealt to the warps by list length, and tile t takes rows ql * n_tiles + t of the launch (an
// even sample of the query range) so that the tiles of a wave cost the same.
//
// Structure.  One CTA of 16 warps per SM.  Shared memory: the current slab (136 KB, one 1-D TMA bulk copy), the set-bit lists of the
// CTA's 128 queries (64 KB), per-query compare masks + list length, the per-query sorted top-k keys (dist << 23 | index) and the
// dealing of the tile's queries to the warps.  Survivors of the
// test are rare after the first slabs; their exact distances are read off the planes and inserted into the sorted top-k by
// warp ballots.  A slab that yields more than 64 survivors for a query (the first one always does) is first cut down to the k best
// (+ ties) by a bit-sliced radix select.  Work items are (query tile x pool split); with more than one split per tile partial rows
// go through knn5_merge_kernel.  The query range can come from device memory (KnnDyn) so that the frame path never needs to
// know the keypoint counts on the host.
#include "knn_dev.cuh"

namespace slideo {

namespace {

constexpr int K5_THREADS = 512, K5_WARPS = 16;
constexpr int K5_QPW = KNN5_TILE / K5_WARPS;     // queries per warp
constexpr int K5_W = 4;                          // words per lane and bit row: 128 pooled rows per lane
constexpr int K5_ROWB = K5_W * 128;              // bytes per bit row
constexpr int K5_ROW_PT = 256;                   // rows 256..264: planes of pt' = 256 - popc(t)
constexpr int K5_ROW_VALID = 265;                // row 265: valid mask (pooled index < nt)
constexpr int K5_ROW_ZERO = 266;                 // row 266: zeros (list padding), never overwritten by the slab copy
constexpr int K5_SLAB_BYTES = 266 * K5_ROWB;     // what one bulk copy brings: P[256] + pt'[9] + valid
constexpr int K5_LIST = 128, K5_META = 16;
constexpr int K5_OFF_LIST = 267 * K5_ROWB;
constexpr int K5_OFF_META = K5_OFF_LIST + KNN5_TILE * K5_LIST * 4;
constexpr int K5_OFF_TOPK = K5_OFF_META + KNN5_TILE * K5_META * 4;
constexpr int K5_OFF_BAR = K5_OFF_TOPK + KNN5_TILE * 32 * 4;
constexpr int K5_OFF_PERM = K5_OFF_BAR + 16;
constexpr int K5_SMEM = K5_OFF_PERM + KNN5_TILE * 4;
constexpr int K5_META_NH = 14;                   // meta word 14: the query's list length in half blocks of 8 entries
// K5_GRAN: granularity at which the list walk stops at the end of a query's list: 0 = never (always 128 entries),
// 1 = whole blocks of 16 entries, 2 = half blocks of 8.  K5_FLIP: pool rows and queries are XORed with the pool's majority vector.
#ifndef K5_GRAN
#define K5_GRAN 1
#endif
#ifndef K5_FLIP
#define K5_FLIP 1
#endif
// K5_PREFETCH: the entries of the next block are requested before the adders of the current one (measured: +0.3 %, inside the
// run-to-run spread; off).
#ifndef K5_PREFETCH
#define K5_PREFETCH 0
#endif
constexpr int K5_FLIP_SAMPLE = 8192;             // rows of the pool the majority vector is counted on
constexpr int K5_RADIX_ABOVE = 64;               // survivors of one (query, slab) above which the radix select runs first
static_assert(K5_SLAB_BYTES == KNN5_SLAB_BYTES, "slab size");
static_assert(K5_W * 1024 == KNN5_SLAB_ROWS, "slab rows");
static_assert(K5_SMEM <= 232448, "shared memory budget of one SM");

struct V4 { uint32_t v[K5_W]; };

__device__ __forceinline__ V4 lds4(uint32_t addr) {
    V4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]) : "r"(addr));
    return r;
}
// full adder on bit planes: sum and carry of a + b + c (sum may alias a)
__device__ __forceinline__ void csa(V4& sum, V4& carry, const V4& a, const V4& b, const V4& c) {
#pragma unroll
    for (int w = 0; w < K5_W; ++w) {
        const uint32_t x = a.v[w], y = b.v[w], z = c.v[w];
        sum.v[w] = knn_lop3<KNN_LUT_XOR3>(x, y, z);
        carry.v[w] = knn_lop3<KNN_LUT_MAJ>(x, y, z);
    }
}
// integer multiply-add pinned to the FMA pipe (the ALU pipe is the bottleneck)
__device__ __forceinline__ int imad(int a, int b, int c) {
    int r;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(KNN_FULL, v, o);
    return v;
}

struct Knn5Params {
    const uint4* q;        // query descriptors (base of the range when dyn != nullptr)
    const uint8_t* slabs;  // bit-sliced pool
    uint32_t* keys_out;    // [nq][k], may be null
    uint32_t* partial;     // [item][KNN5_TILE][k], only written when splits > 1
    const KnnDyn* dyn;     // null: the static fields below describe the launch
    int nq, k, n_tiles, n_slabs, splits;
    VoteArgs vote;         // vote.votes == nullptr -> no fused vote
};

// compare masks of one query: the test "s >= T" as the carry out of s + C + cin over 10 bits, C + cin = 1024 - T.
//   normal list (inv == 0):  d = A + 256 - s < tau  <=>  s >= A + 257 - tau
//   complemented list     :  d = A - 256 + s < tau  <=>  !(s >= tau - A + 256)
__device__ __forceinline__ void k5_set_meta(uint32_t* meta, int lane, int A, bool inv, int tau) {
    int T = inv ? tau - A + 256 : A + 257 - tau;
    T = max(0, min(1024, T));
    int C = 1024 - T;
    const bool cin = C == 1024;
    if (cin) C = 1023;
    if (lane < 10) meta[lane] = (C >> lane) & 1 ? 0xFFFFFFFFu : 0u;
    else if (lane == 10) meta[10] = cin ? 0xFFFFFFFFu : 0u;
    else if (lane == 11) meta[11] = inv ? 0xFFFFFFFFu : 0u;
    else if (lane == 12) meta[12] = (uint32_t)A;
    else if (lane == 13) meta[13] = (uint32_t)tau;
}

struct K5Planes { V4 pl0, ones, twos, fours, eights, s16, s32, s64, s128, s256; };   // s = 2c + pt' as bit planes 0..9

// one query against the slab in shared memory: walks the query's list of bit rows (nh half blocks of 8 entries, padded with the
// zero row up to the next block boundary) and adds the words of those rows for this lane's 128 pooled rows.  Harley-Seal: 15 full
// adders per 16 inputs, a second level over the 8 weight-16 carries; the accumulators are seeded with pt' >> 1, so the total is
// s = 2c + pt'.  ORB descriptors of slides are far from balanced (mean list length 80-90 of 128 after the majority flip), so the walk
// stops at the end of the list: nh is warp-uniform.
__device__ __forceinline__ void k5_half(const uint4* lst, int lanebase, V4& ones, V4& twos, V4& fours, V4& e) {
    const uint4 e0 = lst[0], e1 = lst[1];
    const uint32_t ent[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
    V4 x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = lds4((uint32_t)imad((int)ent[i], K5_ROWB, lanebase));
    V4 ta, tb, fa, fb;
    csa(ones, ta, ones, x[0], x[1]);
    csa(ones, tb, ones, x[2], x[3]);
    csa(twos, fa, twos, ta, tb);
    csa(ones, ta, ones, x[4], x[5]);
    csa(ones, tb, ones, x[6], x[7]);
    csa(twos, fb, twos, ta, tb);
    csa(fours, e, fours, fa, fb);
}

__device__ __forceinline__ void k5_scan(const uint4* lst, int lanebase, int nh, K5Planes& P) {
    V4 ones = lds4(lanebase + (K5_ROW_PT + 1) * K5_ROWB), twos = lds4(lanebase + (K5_ROW_PT + 2) * K5_ROWB),
       fours = lds4(lanebase + (K5_ROW_PT + 3) * K5_ROWB), eights = lds4(lanebase + (K5_ROW_PT + 4) * K5_ROWB),
       s16 = lds4(lanebase + (K5_ROW_PT + 5) * K5_ROWB), s32 = lds4(lanebase + (K5_ROW_PT + 6) * K5_ROWB),
       s64 = lds4(lanebase + (K5_ROW_PT + 7) * K5_ROWB), s128 = lds4(lanebase + (K5_ROW_PT + 8) * K5_ROWB);
    V4 s256, o_prev, t32a, u64a;
#pragma unroll
    for (int w = 0; w < K5_W; ++w) s256.v[w] = o_prev.v[w] = t32a.v[w] = u64a.v[w] = 0;
#if K5_PREFETCH
    uint4 en0 = lst[0], en1 = lst[1], en2 = lst[2], en3 = lst[3];
#endif
#pragma unroll
    for (int blk = 0; blk < 8; ++blk) {
        V4 o;
        if (K5_GRAN == 0 || 2 * blk < nh) {
            if (K5_GRAN != 2 || 2 * blk + 1 < nh) {
                // whole block: all 16 loads are in flight before the first adder
#if K5_PREFETCH
                const uint4 e0 = en0, e1 = en1, e2 = en2, e3 = en3;
#else
                const uint4 e0 = lst[blk * 4], e1 = lst[blk * 4 + 1], e2 = lst[blk * 4 + 2], e3 = lst[blk * 4 + 3];
#endif
                const uint32_t ent[16] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w, e2.x, e2.y, e2.z, e2.w, e3.x, e3.y, e3.z, e3.w};
                V4 x[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = lds4((uint32_t)imad((int)ent[i], K5_ROWB, lanebase));
#if K5_PREFETCH
                // the entries of the next block (the list is padded to 128 entries, so they exist) are requested before this block's
                // adders: a block that starts behind a branch then waits for one shared-memory round trip instead of two
                if (blk < 7) { en0 = lst[blk * 4 + 4]; en1 = lst[blk * 4 + 5]; en2 = lst[blk * 4 + 6]; en3 = lst[blk * 4 + 7]; }
#endif
                V4 ta, tb, fa, fb, ea, eb;
                csa(ones, ta, ones, x[0], x[1]);
                csa(ones, tb, ones, x[2], x[3]);
                csa(twos, fa, twos, ta, tb);
                csa(ones, ta, ones, x[4], x[5]);
                csa(ones, tb, ones, x[6], x[7]);
                csa(twos, fb, twos, ta, tb);
                csa(fours, ea, fours, fa, fb);
                csa(ones, ta, ones, x[8], x[9]);
                csa(ones, tb, ones, x[10], x[11]);
                csa(twos, fa, twos, ta, tb);
                csa(ones, ta, ones, x[12], x[13]);
                csa(ones, tb, ones, x[14], x[15]);
                csa(twos, fb, twos, ta, tb);
                csa(fours, eb, fours, fa, fb);
                csa(eights, o, eights, ea, eb);
            } else {
                // the list ends in the first half of this block: 8 entries, the weight-8 carry meets no partner
                V4 ea;
                k5_half(lst + blk * 4, lanebase, ones, twos, fours, ea);
#pragma unroll
                for (int w = 0; w < K5_W; ++w) { o.v[w] = eights.v[w] & ea.v[w]; eights.v[w] ^= ea.v[w]; }
            }
        } else {
#pragma unroll
            for (int w = 0; w < K5_W; ++w) o.v[w] = 0;
        }
        if (blk & 1) {
            V4 t;
            csa(s16, t, s16, o_prev, o);
            if ((blk & 3) == 3) {
                V4 u;
                csa(s32, u, s32, t32a, t);
                if (blk == 7) {
                    V4 v;
                    csa(s64, v, s64, u64a, u);
#pragma unroll
                    for (int w = 0; w < K5_W; ++w) { s256.v[w] = s128.v[w] & v.v[w]; s128.v[w] ^= v.v[w]; }
                } else u64a = u;
            } else t32a = t;
        } else o_prev = o;
    }
    P.pl0 = lds4(lanebase + K5_ROW_PT * K5_ROWB);
    P.ones = ones; P.twos = twos; P.fours = fours; P.eights = eights; P.s16 = s16; P.s32 = s32; P.s64 = s64; P.s128 = s128; P.s256 = s256;
}

// survivors of the test d < tau: the carry out of s + C + cin over the 10 planes, inverted for a complemented list, masked by
// the slab's valid rows.  Returns the OR of the four result words.
__device__ __forceinline__ uint32_t k5_compare(const K5Planes& P, const uint32_t* meta, int lanebase, uint32_t (&res)[K5_W]) {
    const V4 valid = lds4(lanebase + K5_ROW_VALID * K5_ROWB);
    const uint4 m0 = *reinterpret_cast<const uint4*>(meta), m1 = *reinterpret_cast<const uint4*>(meta + 4),
                m2 = *reinterpret_cast<const uint4*>(meta + 8);
    uint32_t any = 0;
#pragma unroll
    for (int w = 0; w < K5_W; ++w) {
        uint32_t cy = knn_lop3<KNN_LUT_MAJ>(P.pl0.v[w], m0.x, m2.z);
        cy = knn_lop3<KNN_LUT_MAJ>(P.ones.v[w], m0.y, cy);
        cy = knn_lop3<KNN_LUT_MAJ>(P.twos.v[w], m0.z, cy);
        cy = knn_lop3<KNN_LUT_MAJ>(P.fours.v[w], m0.w, cy);
        cy = knn_lop3<KNN_LUT_MAJ>(P.eights.v[w], m1.x, cy);
        cy = knn_lop3<KNN_LUT_MAJ>(P.s16.v[w], m1.y, cy);
        cy = knn_lop3<KNN_LUT_MAJ>(P.s32.v[w], m1.z, cy);
        cy = knn_lop3<KNN_LUT_MAJ>(P.s64.v[w], m1.w, cy);
        cy = knn_lop3<KNN_LUT_MAJ>(P.s128.v[w], m2.x, cy);
        cy = knn_lop3<KNN_LUT_MAJ>(P.s256.v[w], m2.y, cy);
        res[w] = (cy ^ m2.w) & valid.v[w];
        any |= res[w];
    }
    return any;
}

__global__ void __launch_bounds__(K5_THREADS, 1) knn5_kernel(const Knn5Params P) {
    extern __shared__ __align__(128) uint8_t s_raw[];
    uint32_t* s_list = reinterpret_cast<uint32_t*>(s_raw + K5_OFF_LIST);   // [tile][128] bit-row indices
    uint32_t* s_meta = reinterpret_cast<uint32_t*>(s_raw + K5_OFF_META);   // [tile][16]
    uint32_t* s_topk = reinterpret_cast<uint32_t*>(s_raw + K5_OFF_TOPK);   // [tile][32] sorted keys
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_raw + K5_OFF_BAR);
    uint32_t* s_perm = reinterpret_cast<uint32_t*>(s_raw + K5_OFF_PERM);   // [warp][K5_QPW] queries of the tile, dealt by list length

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int q_off = 0, nq = P.nq, n_tiles = P.n_tiles, splits = P.splits;
    if (P.dyn != nullptr) {
        q_off = P.dyn->q0; nq = P.dyn->nq; n_tiles = P.dyn->n_tiles; splits = P.dyn->splits;
    }
    const int n_items = n_tiles * splits;
    if ((int)blockIdx.x >= n_items) return;

    if (tid == 0) {
        knn_mbar_init(s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < K5_ROWB / 4; i += K5_THREADS) reinterpret_cast<uint32_t*>(s_raw + K5_ROW_ZERO * K5_ROWB)[i] = 0;
    __syncthreads();

    const uint32_t sbase = knn_smem_u32(s_raw);
    const int lanebase = (int)sbase + lane * 16;
    uint32_t parity = 0;
#if K5_FLIP
    // the pool's majority vector (knn5_flip_kernel) sits behind the last slab; Hamming distances do not change when both sides are
    // XORed with the same vector, the set-bit lists get shorter
    const uint32_t flip = reinterpret_cast<const uint32_t*>(P.slabs + (size_t)P.n_slabs * K5_SLAB_BYTES)[lane & 7];
#else
    const uint32_t flip = 0;
#endif

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int tile = item / splits, split = item - tile * splits;
        const int slab0 = (int)((long long)P.n_slabs * split / splits), slab1 = (int)((long long)P.n_slabs * (split + 1) / splits);
        // query ql of tile t is row ql * n_tiles + t of the launch: every tile is an even sample of the whole query range, so that
        // tiles (and with them the CTAs of a wave) cost the same although the walk of a query costs what its list is long and
        // neighbouring keypoints have lists of similar length (tile-to-tile spread of contiguous tiles: 10-17 %)

        __syncthreads();   // the previous item is emitted: its lists, masks and top-k rows may be overwritten by any warp
        // ---- tile prologue: set-bit lists, compare masks, empty top-k of queries warp * 8 .. warp * 8 + 7 -------------------
#pragma unroll 1
        for (int qi = 0; qi < K5_QPW; ++qi) {
            const int ql = warp * K5_QPW + qi, q = ql * n_tiles + tile;
            uint32_t* lst = s_list + ql * K5_LIST;
            s_topk[ql * 32 + lane] = KEY_EMPTY;
            if (q >= nq) {   // warp-uniform
                if (lane == 0) s_meta[ql * K5_META + K5_META_NH] = 0;
                continue;
            }
            const uint32_t word = reinterpret_cast<const uint32_t*>(P.q + (size_t)(q_off + q) * 2)[lane & 7] ^ flip;
            int A = __popc(word);
            A += __shfl_xor_sync(KNN_FULL, A, 1);
            A += __shfl_xor_sync(KNN_FULL, A, 2);
            A += __shfl_xor_sync(KNN_FULL, A, 4);
            const bool inv = A > 128;
            // lane L owns byte L of the descriptor = bits 8L .. 8L+7
            const uint32_t wsrc = __shfl_sync(KNN_FULL, word, lane >> 2);
            uint32_t byte = (wsrc >> (8 * (lane & 3))) & 0xFFu;
            if (inv) byte ^= 0xFFu;
            const int cnt = __popc(byte);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(KNN_FULL, incl, o);
                if (lane >= o) incl += t;
            }
            int pos = incl - cnt;
            while (byte) {
                const int j = __ffs(byte) - 1;
                byte &= byte - 1;
                lst[pos++] = (uint32_t)(lane * 8 + j);
            }
            const int n = __shfl_sync(KNN_FULL, incl, 31);
            for (int i = n + lane; i < K5_LIST; i += 32) lst[i] = K5_ROW_ZERO;
            k5_set_meta(s_meta + ql * K5_META, lane, A, inv, 512);
            if (lane == 0) s_meta[ql * K5_META + K5_META_NH] = K5_GRAN == 2 ? (uint32_t)((n + 7) >> 3) : (uint32_t)(((n + 15) >> 4) << 1);
        }
        // ---- the tile's queries are dealt to the warps by list length (descending, boustrophedon), so that the warps of the CTA
        //      reach the per-slab barrier together although a query's walk costs what its list is long ----------------------------
        __syncthreads();
        if (tid < KNN5_TILE) {
            const uint32_t mine = s_meta[tid * K5_META + K5_META_NH];
            int rank = 0;
#pragma unroll 8
            for (int j = 0; j < KNN5_TILE; ++j) {
                const uint32_t o = s_meta[j * K5_META + K5_META_NH];
                rank += (o > mine || (o == mine && j < tid)) ? 1 : 0;
            }
            const int round = rank / K5_WARPS, j = rank - round * K5_WARPS;
            s_perm[((round & 1) ? K5_WARPS - 1 - j : j) * K5_QPW + round] = (uint32_t)tid;
        }

        if (slab0 >= slab1) __syncthreads();   // empty pool: the emission below still reads the dealing
        for (int slab = slab0; slab < slab1; ++slab) {
            __syncthreads();   // every warp is done with the previous slab; the lists and the dealing of this tile are written
            if (tid == 0) {
                knn_mbar_expect_tx(s_bar, K5_SLAB_BYTES);
                knn_bulk_g2s(s_raw, P.slabs + (size_t)slab * K5_SLAB_BYTES, K5_SLAB_BYTES, s_bar);
            }
            knn_mbar_wait(s_bar, parity);
            parity ^= 1;
            const uint32_t row0 = (uint32_t)slab * KNN5_SLAB_ROWS + (uint32_t)lane * 128u;

#pragma unroll 1
            for (int qi = 0; qi < K5_QPW; ++qi) {
                const int ql = (int)s_perm[warp * K5_QPW + qi];
                if (ql * n_tiles + tile >= nq) continue;   // warp-uniform
                uint32_t* meta = s_meta + ql * K5_META;
                K5Planes pl_;
                k5_scan(reinterpret_cast<const uint4*>(s_list + ql * K5_LIST), lanebase, (int)meta[K5_META_NH], pl_);
                uint32_t res[K5_W];
                const uint32_t any = k5_compare(pl_, meta, lanebase, res);
                const V4 &pl0 = pl_.pl0, &ones = pl_.ones, &twos = pl_.twos, &fours = pl_.fours, &eights = pl_.eights, &s16 = pl_.s16,
                         &s32 = pl_.s32, &s64 = pl_.s64, &s128 = pl_.s128, &s256 = pl_.s256;
                const uint4 m2 = *reinterpret_cast<const uint4*>(meta + 8);
                if (!__any_sync(KNN_FULL, any != 0)) continue;

                // ---- slow path: exact distances of the survivors, read off the planes and inserted into the query's sorted top-k -------
                const uint4 m3 = *reinterpret_cast<const uint4*>(meta + 12);
                const int A = (int)m3.x, tau = (int)m3.y;
                const uint32_t invm = m2.w;
                const V4* pl[10] = {&pl0, &ones, &twos, &fours, &eights, &s16, &s32, &s64, &s128, &s256};
                const int mine = __popc(res[0]) + __popc(res[1]) + __popc(res[2]) + __popc(res[3]);
                if (__any_sync(KNN_FULL, mine > 2) && warp_sum(mine) > K5_RADIX_ABOVE) {
                    // bit-sliced radix select: keep the P.k best survivors (largest s, or smallest s for a complemented list)
                    // plus everything tied with the k-th; whatever is dropped is beaten by >= k rows of this slab alone
                    uint32_t sure[K5_W] = {0, 0, 0, 0};
                    int need = P.k;
#pragma unroll
                    for (int p = 9; p >= 0; --p) {
                        uint32_t one[K5_W];
                        int cnt = 0;
#pragma unroll
                        for (int w = 0; w < K5_W; ++w) {
                            one[w] = res[w] & (pl[p]->v[w] ^ invm);
                            cnt += __popc(one[w]);
                        }
                        cnt = warp_sum(cnt);
                        if (cnt >= need) {
#pragma unroll
                            for (int w = 0; w < K5_W; ++w) res[w] = one[w];
                        } else {
                            need -= cnt;
#pragma unroll
                            for (int w = 0; w < K5_W; ++w) { sure[w] |= one[w]; res[w] &= ~one[w]; }
                        }
                    }
#pragma unroll
                    for (int w = 0; w < K5_W; ++w) res[w] |= sure[w];
                }
                // immediate insertion keeps tau fresh: a stale threshold admits far more survivors than the merge saves
                // (measured: pending buffers merged every 20 candidates ran 4 % slower than this)
                uint32_t cur = s_topk[ql * 32 + lane];
#pragma unroll
                for (int w = 0; w < K5_W; ++w) {
                    while (__any_sync(KNN_FULL, res[w] != 0)) {
                        uint32_t key = KEY_EMPTY;
                        if (res[w] != 0) {
                            const int r = __ffs(res[w]) - 1;
                            res[w] &= res[w] - 1;
                            // bit r of the 10 planes; AND on the ALU pipe, POPC on the (idle) XU pipe, shift-add on the FMA pipe
                            const uint32_t bit = 1u << r;
                            int s = 0;
#pragma unroll
                            for (int p = 0; p < 10; ++p) s = imad(__popc(pl[p]->v[w] & bit), 1 << p, s);
                            const int d = invm ? A - 256 + s : A + 256 - s;
                            key = ((uint32_t)d << KEY_IDX_BITS) | (row0 + (uint32_t)(w * 32 + r));
                        }
                        unsigned have = __ballot_sync(KNN_FULL, key != KEY_EMPTY);
                        while (have) {
                            const int src = __ffs(have) - 1;
                            have &= have - 1;
                            const uint32_t nk = __shfl_sync(KNN_FULL, key, src);
                            const int pos = __popc(__ballot_sync(KNN_FULL, cur < nk));
                            const uint32_t up = __shfl_up_sync(KNN_FULL, cur, 1);
                            if (lane == pos) cur = nk;
                            else if (lane > pos) cur = up;
                            if (lane >= P.k) cur = KEY_EMPTY;
                        }
                    }
                }
                s_topk[ql * 32 + lane] = cur;
                const uint32_t kth = __shfl_sync(KNN_FULL, cur, P.k - 1);
                const int tau_new = kth == KEY_EMPTY ? 512 : (int)(kth >> KEY_IDX_BITS);
                __syncwarp();   // every lane has read this query's masks / threshold before they are rewritten
                if (tau_new != tau) k5_set_meta(meta, lane, A, invm != 0, tau_new);
                __syncwarp();
            }
        }

        // ---- emission: lane m holds neighbour m of the row (fused K9 vote when the tile was not split) ----------------------
        __syncwarp();
#pragma unroll 1
        for (int qi = 0; qi < K5_QPW; ++qi) {
            const int ql = (int)s_perm[warp * K5_QPW + qi], q = ql * n_tiles + tile;   // the rows this warp owned during the scan
            if (q >= nq) continue;
            const uint32_t key = s_topk[ql * 32 + lane];
            if (splits == 1) {
                knn_emit_row(key, lane, q_off + q, P.k, P.keys_out, P.vote);
            } else if (lane < P.k) {
                P.partial[((size_t)item * KNN5_TILE + ql) * P.k + lane] = key;
            }
        }
        __syncwarp();
    }
}

// merge of the per-split partial rows: one warp per query
__global__ void __launch_bounds__(128) knn5_merge_kernel(const uint32_t* __restrict__ partial, const KnnDyn* __restrict__ dyn, int nq_s,
                                                         int splits_s, int k, uint32_t* keys_out, const VoteArgs vote) {
    int q_off = 0, nq = nq_s, splits = splits_s;
    if (dyn != nullptr) { q_off = dyn->q0; nq = dyn->nq; splits = dyn->splits; }
    const int n_tiles = (nq + KNN5_TILE - 1) / KNN5_TILE;
    if (splits <= 1) return;   // emitted by K8 itself
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const int row = q / n_tiles, tile = q - row * n_tiles;   // K8's query-to-tile map
    uint32_t k0 = KEY_EMPTY;
    for (int s = 0; s < splits; ++s) {
        uint32_t k1 = lane < k ? partial[(((size_t)tile * splits + s) * KNN5_TILE + row) * k + lane] : KEY_EMPTY;
        knn_warp_sort64(k0, k1, lane);
    }
    knn_emit_row(k0, lane, q_off + q, k, keys_out, vote);
}

// The pool's majority vector: bit b is set when more than half of (a fixed, evenly spaced sample of) the pooled rows have it set.
// Any vector keeps the distances exact; this one makes the set-bit lists short (descriptors of slides share their bias).
// One CTA of 32 warps; a warp reads 32 rows at a time and counts every bit column with a ballot.
__global__ void __launch_bounds__(1024) knn5_flip_kernel(const uint4* __restrict__ src, int nt, uint32_t* __restrict__ flip) {
    __shared__ int s_cnt[256];
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 256) s_cnt[tid] = 0;
    __syncthreads();
    const int n_sample = nt < K5_FLIP_SAMPLE ? nt : K5_FLIP_SAMPLE;
    int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // lane i: rows of the sample with bit 32 k + i set
    for (int i0 = (tid >> 5) * 32; i0 < n_sample; i0 += 1024) {   // warp-uniform
        const int i = i0 + lane;
        uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (i < n_sample) {
            const size_t row = (size_t)((long long)i * nt / n_sample);
            const uint4 a = __ldg(src + row * 2), b = __ldg(src + row * 2 + 1);
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                const int c = __popc(__ballot_sync(KNN_FULL, (w[k] >> b) & 1u));
                if (lane == b) cnt[k] += c;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&s_cnt[32 * k + lane], cnt[k]);
    __syncthreads();
    if (tid < 256) {
        const uint32_t word = __ballot_sync(KNN_FULL, 2 * s_cnt[tid] > n_sample);
        if (lane == 0) flip[tid >> 5] = word;
    }
}

// 32 B rows -> bit-sliced slabs.  One warp per 32 pooled rows (one word column of a slab).
__global__ void __launch_bounds__(256) knn5_bitslice_kernel(const uint4* __restrict__ src, int nt, uint8_t* __restrict__ slabs, int n_cols,
                                                            const uint32_t* __restrict__ flip) {
    const int lane = threadIdx.x & 31;
    const int col_g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // global word column: rows 32 col_g ..
    if (col_g >= n_cols) return;
    const int slab = col_g / (KNN5_SLAB_ROWS / 32), j = col_g - slab * (KNN5_SLAB_ROWS / 32);
    const int row = col_g * 32 + lane;
    uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (row < nt) {
        const uint4 a = __ldg(src + (size_t)row * 2), b = __ldg(src + (size_t)row * 2 + 1);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] ^= flip[k];
    }
    uint32_t* out = reinterpret_cast<uint32_t*>(slabs + (size_t)slab * K5_SLAB_BYTES) + j;
    int pc = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        pc += __popc(w[k]);
        uint32_t mine = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const uint32_t bal = __ballot_sync(KNN_FULL, (w[k] >> i) & 1u);
            if (lane == i) mine = bal;
        }
        out[(size_t)(32 * k + lane) * (K5_ROWB / 4)] = mine;   // bit row 32k + lane
    }
    const int ptp = 256 - pc;
    uint32_t mine = 0;
#pragma unroll
    for (int p = 0; p < 9; ++p) {
        const uint32_t bal = __ballot_sync(KNN_FULL, (ptp >> p) & 1);
        if (lane == p) mine = bal;
    }
    const uint32_t val = __ballot_sync(KNN_FULL, row < nt);
    if (lane < 9) out[(size_t)(K5_ROW_PT + lane) * (K5_ROWB / 4)] = mine;
    else if (lane == 9) out[(size_t)K5_ROW_VALID * (K5_ROWB / 4)] = val;
}

// picks the number of pool splits per query tile: minimise  rounds(S) * (slabs per item + fixed cost of an item)
__host__ __device__ inline int knn5_choose_splits(int n_tiles, int n_slabs, int grid_max) {
    if (n_tiles <= 0 || n_slabs <= 1) return 1;
    int s_max = n_slabs < 64 ? n_slabs : 64;
    const int lim = 4 * grid_max / n_tiles;            // keeps items <= max(n_tiles, 4 * grid)
    if (s_max > lim) s_max = lim < 1 ? 1 : lim;
    int best = 1;
    double best_cost = 0;
    for (int s = 1; s <= s_max; ++s) {
        const long long items = (long long)n_tiles * s;
        const long long rounds = (items + grid_max - 1) / grid_max;
        const double cost = (double)rounds * ((double)n_slabs / s + 1.5) * (s == 1 ? 1.0 : 1.03);
        if (s == 1 || cost < best_cost) { best = s; best_cost = cost; }
    }
    return best;
}

// fills *dyn for the next launch over the ready part of a query stream (see api.cu: the frame path).  Runs on the detection
// stream, so q_write / f_write cover exactly the batches whose descriptors are complete.
__global__ void knn5_plan_kernel(KnnStream* st, KnnDyn* dyn, int n_slabs, int nt, int grid_max, int flush) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int ready = st->q_write - st->q_matched;
    int take = ready;
    if (!flush) {   // steady state: whole waves of tiles only, the remainder rides with the next launch
        const int wave = grid_max * KNN5_TILE;
        take = ready / wave * wave;
    }
    dyn->q0 = st->q_matched;
    dyn->nq = take;
    dyn->n_tiles = (take + KNN5_TILE - 1) / KNN5_TILE;
    dyn->splits = knn5_choose_splits(dyn->n_tiles, n_slabs, grid_max);
    dyn->f_limit = st->f_write;
    st->q_matched += take;
    st->pairs += (unsigned long long)take * (unsigned long long)nt;
}

// number of frames of the stream that are complete once queries [0, q_end) are matched: frame f owns [frame_q0[f], frame_q0[f+1])
__device__ __forceinline__ int frames_complete(const int32_t* frame_q0, int f_lo, int f_hi, int q_end) {
    int lo = f_lo, hi = f_hi;   // invariant: frames < lo complete, frames >= hi not
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (frame_q0[mid + 1] <= q_end) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// K9 tail for the frames a K8 launch completed: argmax over the vote table (ties -> lowest page, lib.rs:284-295 head),
// written straight into the host-mapped result ring.  One warp per frame, grid-stride.
__global__ void __launch_bounds__(128) stream_finalize_kernel(const KnnStream* st, const KnnDyn* dyn, const int32_t* __restrict__ frame_q0,
                                                              const int32_t* __restrict__ votes, int n_pages,
                                                              const int32_t* __restrict__ frame_nkp, int32_t* h_ring, int ring_mask,
                                                              long long seq_base, int32_t* d_results) {
    const int lane = threadIdx.x & 31;
    const int f_lo = st->f_done;
    const int f_new = frames_complete(frame_q0, f_lo, dyn->f_limit, dyn->q0 + dyn->nq);
    const int n_warps = gridDim.x * (blockDim.x >> 5);
    for (int f = f_lo + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); f < f_new; f += n_warps) {
        int bv = 0, bp = -1;
        for (int p = lane; p < n_pages; p += 32) {
            const int v = votes[(size_t)f * n_pages + p];
            if (v > bv) { bv = v; bp = p; }   // increasing p per lane: first max wins
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int ov = __shfl_xor_sync(KNN_FULL, bv, o), op = __shfl_xor_sync(KNN_FULL, bp, o);
            if (ov > bv || (ov == bv && ov > 0 && op < bp)) { bv = ov; bp = op; }
        }
        if (lane == 0) {
            const int nk = frame_nkp[f];
            int32_t* r = h_ring + (size_t)((seq_base + f) & ring_mask) * 3;
            r[0] = bp; r[1] = bv; r[2] = nk;
            if (d_results) { d_results[3 * f] = bp; d_results[3 * f + 1] = bv; d_results[3 * f + 2] = nk; }
        }
    }
}

__global__ void stream_publish_kernel(KnnStream* st, const KnnDyn* dyn, const int32_t* __restrict__ frame_q0, long long seq_base,
                                      volatile long long* h_progress, volatile int* h_flags) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int f_new = frames_complete(frame_q0, st->f_done, dyn->f_limit, dyn->q0 + dyn->nq);
    st->f_done = f_new;
    if (st->flags) *h_flags = *h_flags | st->flags;
    __threadfence_system();   // the result rows of the finalize kernel (previous launch on this stream) and the flags first
    *h_progress = seq_base + f_new;
}

// The v5 inner loop (list walk + carry-save tree + compare) on synthetic shared-memory contents, no selection, no slab reloads:
// the ceiling of this formulation in descriptor pairs per second (slideo_b200_microbench which = 3).
__global__ void __launch_bounds__(K5_THREADS, 1) knn5_microbench_kernel(uint32_t* out, int iters) {
    extern __shared__ __align__(128) uint8_t s_raw[];
    uint32_t* s_words = reinterpret_cast<uint32_t*>(s_raw);
    uint32_t* s_list = reinterpret_cast<uint32_t*>(s_raw + K5_OFF_LIST);
    uint32_t* s_meta = reinterpret_cast<uint32_t*>(s_raw + K5_OFF_META);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t h = (blockIdx.x * K5_THREADS + tid) * 2654435761u + 12345u;
    for (int i = tid; i < 267 * K5_ROWB / 4; i += K5_THREADS) { h = h * 1664525u + 1013904223u; s_words[i] = h ^ (h >> 13); }
    for (int i = tid; i < KNN5_TILE * K5_LIST; i += K5_THREADS) { h = h * 1664525u + 1013904223u; s_list[i] = (h >> 8) & 255u; }
    __syncthreads();
    for (int ql = warp * K5_QPW; ql < (warp + 1) * K5_QPW; ++ql) k5_set_meta(s_meta + ql * K5_META, lane, 128, false, 40);
    __syncthreads();
    const int lanebase = (int)knn_smem_u32(s_raw) + lane * 16;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int qi = 0; qi < K5_QPW; ++qi) {
            const int ql = warp * K5_QPW + qi;
            K5Planes pl;
            k5_scan(reinterpret_cast<const uint4*>(s_list + ql * K5_LIST), lanebase, 16, pl);
            uint32_t res[K5_W];
            const uint32_t any = k5_compare(pl, s_meta + ql * K5_META, lanebase, res);
            if (__any_sync(KNN_FULL, any != 0)) acc += __popc(any);
        }
    }
    out[blockIdx.x * K5_THREADS + tid] = acc;
}

void k5_configure() {
    // per device and cheap: set on every launch (a process may drive several GPUs from several ctxs)
    SLIDEO_CUDA(cudaFuncSetAttribute(knn5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K5_SMEM));
    SLIDEO_CUDA(cudaFuncSetAttribute(knn5_microbench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K5_SMEM));
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
size_t knn5_pool_bytes(int nt) { return (size_t)knn5_slabs(nt) * K5_SLAB_BYTES; }

void knn5_pool_prepare_launch(const void* d_pool32, int nt, void* d_slabs, cudaStream_t stream) {
    const int n_cols = knn5_slabs(nt) * (KNN5_SLAB_ROWS / 32);
    // the majority vector lives behind the last slab (callers reserve knn5_pool_bytes(nt) + 64)
    uint32_t* d_flip = reinterpret_cast<uint32_t*>((uint8_t*)d_slabs + knn5_pool_bytes(nt));
    if (n_cols == 0 || !K5_FLIP) SLIDEO_CUDA(cudaMemsetAsync(d_flip, 0, 32, stream));
    if (n_cols == 0) return;   // empty pool: no slab, K8 emits empty rows
    if (K5_FLIP) knn5_flip_kernel<<<1, 1024, 0, stream>>>((const uint4*)d_pool32, nt, d_flip);
    knn5_bitslice_kernel<<<cdiv(n_cols, 8), 256, 0, stream>>>((const uint4*)d_pool32, nt, (uint8_t*)d_slabs, n_cols, d_flip);
    SLIDEO_CUDA(cudaGetLastError());
}

Knn5Plan knn5_plan(int nq, int nt, int k, int num_sms) {
    Knn5Plan p;
    p.nq = nq; p.nt = nt; p.k = k;
    p.n_tiles = cdiv(nq, KNN5_TILE);
    p.n_slabs = knn5_slabs(nt);
    p.splits = knn5_choose_splits(p.n_tiles, p.n_slabs, num_sms);
    const long long items = (long long)p.n_tiles * p.splits;
    p.grid = (int)(items < num_sms ? items : num_sms);
    if (p.grid < 1) p.grid = 1;
    p.partial_bytes = p.splits > 1 ? (size_t)items * KNN5_TILE * k * sizeof(uint32_t) : 0;
    return p;
}

size_t knn5_dyn_partial_bytes(int num_sms, int k) { return (size_t)4 * num_sms * KNN5_TILE * k * sizeof(uint32_t); }

void knn5_launch(const Knn5Plan& plan, const void* d_q, const void* d_slabs, uint32_t* d_keys_out, uint32_t* d_partial,
                 const VoteArgs* vote, cudaStream_t stream, int* launches) {
    if (plan.nq <= 0) return;
    Knn5Params P;
    P.q = (const uint4*)d_q;
    P.slabs = (const uint8_t*)d_slabs;
    P.keys_out = d_keys_out;
    P.partial = d_partial;
    P.dyn = nullptr;
    P.nq = plan.nq; P.k = plan.k; P.n_tiles = plan.n_tiles; P.n_slabs = plan.n_slabs; P.splits = plan.splits;
    P.vote = vote ? *vote : VoteArgs{nullptr, nullptr, nullptr, 0, 0.f};
    k5_configure();
    knn5_kernel<<<plan.grid, K5_THREADS, K5_SMEM, stream>>>(P);
    SLIDEO_CUDA(cudaGetLastError());
    if (launches) ++*launches;
    if (plan.splits > 1) {
        knn5_merge_kernel<<<cdiv(plan.nq, 4), 128, 0, stream>>>(d_partial, nullptr, plan.nq, plan.splits, plan.k, d_keys_out, P.vote);
        SLIDEO_CUDA(cudaGetLastError());
        if (launches) ++*launches;
    }
}

void knn5_plan_launch(KnnStream* d_state, KnnDyn* d_dyn, int nt, int num_sms, int flush, cudaStream_t stream) {
    knn5_plan_kernel<<<1, 32, 0, stream>>>(d_state, d_dyn, knn5_slabs(nt), nt, num_sms, flush);
    SLIDEO_CUDA(cudaGetLastError());
}

void stream_finalize_launch(KnnStream* d_state, const KnnDyn* d_dyn, const int32_t* d_frame_q0, const int32_t* d_votes, int n_pages,
                            const int32_t* d_frame_nkp, int32_t* h_ring, int ring_mask, long long seq_base, int32_t* d_results,
                            volatile long long* h_progress, volatile int* h_flags, cudaStream_t stream) {
    stream_finalize_kernel<<<32, 128, 0, stream>>>(d_state, d_dyn, d_frame_q0, d_votes, n_pages, d_frame_nkp, h_ring, ring_mask, seq_base,
                                                   d_results);
    stream_publish_kernel<<<1, 32, 0, stream>>>(d_state, d_dyn, d_frame_q0, seq_base, h_progress, h_flags);
    SLIDEO_CUDA(cudaGetLastError());
}

void knn5_launch_dyn(const KnnDyn* d_dyn, int nq_max, int nt, int k, int num_sms, const void* d_q_base, const void* d_slabs,
                     uint32_t* d_keys_base, uint32_t* d_partial, const VoteArgs* vote_base, cudaStream_t stream, int* launches) {
    Knn5Params P;
    P.q = (const uint4*)d_q_base;
    P.slabs = (const uint8_t*)d_slabs;
    P.keys_out = d_keys_base;
    P.partial = d_partial;
    P.dyn = d_dyn;
    P.nq = 0; P.k = k; P.n_tiles = 0; P.n_slabs = knn5_slabs(nt); P.splits = 1;
    P.vote = vote_base ? *vote_base : VoteArgs{nullptr, nullptr, nullptr, 0, 0.f};
    k5_configure();
    knn5_kernel<<<num_sms, K5_THREADS, K5_SMEM, stream>>>(P);
    SLIDEO_CUDA(cudaGetLastError());
    // partial rows only exist for launches the plan kernel decided to split (few tiles); the merge exits at once otherwise
    knn5_merge_kernel<<<cdiv(4 * num_sms * KNN5_TILE < nq_max ? 4 * num_sms * KNN5_TILE : nq_max, 4), 128, 0, stream>>>(
        d_partial, d_dyn, 0, 1, k, d_keys_base, P.vote);
    SLIDEO_CUDA(cudaGetLastError());
    if (launches) *launches += 2;
}

double knn5_microbench_run(int num_sms, cudaStream_t stream) {
    const int iters = 64;
    uint32_t* d_out = nullptr;
    SLIDEO_CUDA(cudaMalloc(&d_out, (size_t)num_sms * K5_THREADS * 4));
    cudaEvent_t e0, e1;
    SLIDEO_CUDA(cudaEventCreate(&e0));
    SLIDEO_CUDA(cudaEventCreate(&e1));
    k5_configure();
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        SLIDEO_CUDA(cudaEventRecord(e0, stream));
        knn5_microbench_kernel<<<num_sms, K5_THREADS, K5_SMEM, stream>>>(d_out, iters);
        SLIDEO_CUDA(cudaEventRecord(e1, stream));
        SLIDEO_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        SLIDEO_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return (double)num_sms * iters * KNN5_TILE * KNN5_SLAB_ROWS / (best * 1e-3);
}

}  // namespace slideo
