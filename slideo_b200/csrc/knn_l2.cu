// knn_l2.cu -- K10: exact brute-force L2 k-NN (k <= 32) of 128-d float descriptors on the 5th-gen tensor cores.
//
// The SIFT variant of the hot path (BASELINE.json north_star; not in the reference itself -- SURVEY.md D5/a10): the
// semantics are cv2.BFMatcher(NORM_L2).knnMatch = the k smallest by (sqrtf(sum (a-b)^2), pooled index), rows ascending.
//
// d^2 = |q|^2 + |t|^2 - 2 q.t is evaluated ENTIRELY inside one bf16 GEMM with K = 128 + 16:
//     A (queries, M side)  row = [ -2*q (128) | 1, 1, 1, qn_hi, qn_mid, qn_lo, 0 x 10 ]
//     B (pool,    N side)  row = [    t (128) | tn_hi, tn_mid, tn_lo, 1, 1, 1, 0 x 10 ]
// where the squared norms are split into three bf16 pieces (24 significant bits).  For integer-valued descriptors
// (cv2 SIFT emits integers 0..255: exactly representable in bf16, every product and partial sum an integer < 2^24)
// the fp32 accumulator in TMEM holds the exact squared distance, so sqrtf() of it is cv2's distance bit for bit.
//
// Kernel anatomy (persistent, one CTA per SM, 448 threads):
//     warp 0       TMA producer   cp.async.bulk.tensor.2d (SWIZZLE_128B main blocks, SWIZZLE_32B norm tail) + mbarriers
//     warp 1       MMA issuer     tcgen05.mma.cta_group::1.kind::f16, M = 128 queries x N = 256 pooled rows, 9 K-steps,
//                                 fp32 accumulators in TMEM (2 x 256 columns = all 512), tcgen05.commit -> mbarriers
//     warps 2-9    drain warps    tcgen05.ld.32x32b.x32: thread = query row, 32 pooled columns per load.  Fast path of a chunk: a
//                                 3-input-min tree and ONE fp32 compare against the row's threshold.  Slow path (a fifth of the
//                                 chunks at a 1 M pool): the hit 4-column group is picked by a select tree and its survivors are
//                                 appended as 64-bit keys (d2 bits << 32 | index) to the thread's region in L2 scratch.  Warps
//                                 2-5 drain columns 0..127 of EVERY tile, warps 6-9 columns 128..255 (TMEM buffer = tile
//                                 parity): an accumulator is back with the tensor pipe after half a drain.
//     warps 10-13  sorter warps   one per TMEM lane quarter: own ONE sorted list of the k best per query row, merge the regions
//                                 the drain threads hand over (streaming bitonic top-32), publish the row's k-th distance, and
//                                 emit the rows at the end of a work item.  The drain warps never sort and never wait for an
//                                 emission, so the tensor pipe does not either.
// A work item (query tile x pool range) of >= 4 tiles starts with a threshold pre-pass (l2_npre): a sixth of its tiles (4..32)
// run through the tensor pipe once only to bound every row's k-th distance by the maximum of k group minima.
#include <cuda.h>
#include <cuda_bf16.h>

#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <string>

#include "knn_l2.cuh"

namespace slideo {

namespace {

constexpr int L2_DIM = 128;
constexpr int L2_TAIL = 16;
constexpr int L2_BM = 128;              // queries per tile (TMEM lanes)
constexpr int L2_BN = 256;              // pooled descriptors per tile (TMEM columns)
constexpr int L2_STAGES = 2;            // B stages in shared memory
constexpr int L2_THREADS = 448;         // TMA warp, MMA warp, 8 drain warps, 4 sorter warps
constexpr int L2_BNH = L2_BN / 2;       // pooled columns of a tile drained by one epilogue group
constexpr int L2_TRIGGER = 32;          // upper bound of L2Params::trigger (new candidates of a list that schedule a cut-back)
// scratch per (column half, row): slots [0, 32) of the half-0 buffer hold the row's sorted k best as of the last cut-back (written
// by the row's sorter warp only; unused in the half-1 buffer); new candidates are appended by the drain thread of (half, row) to one
// of its two regions of L2_REGION slots.  When a region passes the trigger the thread hands it to the sorter warp and goes on
// appending to the other one; a region takes a whole half tile past the trigger.
constexpr int L2_REGION = L2_TRIGGER + L2_BNH;
constexpr int L2_SLOTS = 32 + 2 * L2_REGION;
constexpr int L2_RQ = 64;               // request ring of a drain warp (at most one request per list + the end marker in flight)
constexpr uint32_t L2_REQ_END = 0xFFFFFFFFu;
constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr uint64_t KEY64_EMPTY = 0xFFFFFFFFFFFFFFFFull;

constexpr uint32_t A_MAIN_BYTES = L2_BM * 128;   // one 64-wide K block of the query tile (SW128: 128 B per row)
constexpr uint32_t A_TAIL_BYTES = L2_BM * 32;
constexpr uint32_t B_MAIN_BYTES = L2_BN * 128;
constexpr uint32_t B_TAIL_BYTES = L2_BN * 32;
constexpr uint32_t A_BYTES = 2 * A_MAIN_BYTES + A_TAIL_BYTES;   // 36 KB
constexpr uint32_t B_BYTES = 2 * B_MAIN_BYTES + B_TAIL_BYTES;   // 72 KB
constexpr uint32_t SMEM_OPERANDS = A_BYTES + L2_STAGES * B_BYTES;   // 180 KB
constexpr uint32_t SMEM_CTRL = 128 + 3 * L2_BM * 4 + 8 * L2_RQ * 4 + 8 * 32 * 4 + 64;   // barriers/tmem ptr, tau / pre-pass exchange, request rings, done counters, ring tails + items done
constexpr uint32_t SMEM_TOTAL = SMEM_OPERANDS + SMEM_CTRL + 1024;   // + alignment slack

// ---- PTX wrappers -----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// the smallest float above x (x finite or +inf, never NaN; +inf stays): branch-free nextafterf(x, +inf)
__device__ __forceinline__ float next_up(float x) {
    const int b = __float_as_int(x);
    const int up = b >= 0 ? b + 1 : (b == (int)0x80000000 ? 1 : b - 1);
    return b == 0x7F800000 ? x : __int_as_float(up);
}
__device__ __forceinline__ float lds_volatile(uint32_t saddr) {
    float v;
    asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_volatile_u32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 |
// version 1 << 46 | layout << 61.  layout: 2 = SWIZZLE_128B (8 rows x 128 B atoms, SBO 1024), 6 = SWIZZLE_32B (8 x 32 B, SBO 256)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)layout << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32 (bit 4), A = B = bf16 (bits 7, 10), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24
// Measured (65 k x 1 M, profiles/r2_k10_steps.txt): without any selection the kernel runs at 1354 TFLOP/s on (2 * 128 + 3) FLOP per
// pair, i.e. the drain + min tree alone keeps up with the tensor pipe; everything below that is the slow path and the coupling it
// causes (a tile is released by ALL drain warps, so one warp's detour stalls the MMA of the tile after next).  Hence: cut-backs on
// their own warps, one divergent region per slow-path visit, a branch-free threshold set-up per tile, and a small instruction
// footprint (the sort network exists once, out of line: inlining it four more times cost 5 % on every shape).  Rejected after
// measuring: N = 128 MMAs with four TMEM buffers (tensor ceiling 1354 -> 1131: the A operand is re-read from shared memory twice as
// often), four epilogue groups on 64-column slices / x64 loads (round 1), three column parts per lane quarter with setmaxnreg (640
// threads), parking hit chunks in a shared-memory ring, a jump table over the union of hit groups, batching or software-pipelining
// the sorter's requests.
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(L2_BN >> 3) << 17) | ((uint32_t)(L2_BM >> 4) << 24);

// 32 keys, one per lane.  warp_sort32_desc: full bitonic sort, descending in lane order.  warp_merge32_asc: the last 5
// stages only -- sorts a BITONIC sequence ascending.  Together they give a streaming "32 smallest": with `top` ascending and a
// fresh chunk x descending, min(top, x) lane by lane holds the 32 smallest of both as a bitonic sequence.
__device__ __forceinline__ uint64_t warp_cmpx(uint64_t v, int stride, bool take_min) {
    const uint64_t o = __shfl_xor_sync(FULL, v, stride);
    const uint64_t mn = v < o ? v : o, mx = v < o ? o : v;
    return take_min ? mn : mx;
}
__device__ __forceinline__ uint64_t warp_sort32_desc(uint64_t v, int lane) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const bool lower = (lane & stride) == 0;
            const bool desc = size == 32 ? true : ((lane & size) == 0);   // final direction descending
            v = warp_cmpx(v, stride, lower != desc);
        }
    }
    return v;
}
__device__ __forceinline__ uint64_t warp_merge32_asc(uint64_t v, int lane) {
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) v = warp_cmpx(v, stride, (lane & stride) == 0);
    return v;
}

// one out-of-line copy for the sorter warps (the kernel's instruction footprint is shared with the drain warps' hot loop)
__device__ __noinline__ uint64_t warp_top32_merge(uint64_t top, uint64_t x, int lane) {
    return warp_merge32_asc(min(top, warp_sort32_desc(x, lane)), lane);
}

struct L2Params {
    int nq, nt, k;
    int n_mtiles, n_ntiles;
    int mt_a, ns_a, ns_b;  // query tiles [0, mt_a) are split ns_a ways over the pool, tiles [mt_a, n_mtiles) ns_b ways; partial rows have ns_max slots
    int ns_max;
    int trigger;           // new candidates of a list that schedule its cut-back (1..L2_TRIGGER)
    int pre_tiles;         // threshold pre-pass: pool tiles at the head of a work item whose accumulators are only reduced to group
    int pre_min, pre_floor; //   minima (see l2_npre); items with fewer than pre_min tiles have no pre-pass
    unsigned long long* prof;  // developer instrument (SLIDEO_L2_PROF): cycles of drain warps in {acc wait, drain, request posting, item tail}, of the MMA warp in {acc_empty wait, b_full wait}, of the sorter warps (busy, requests)
    int dbg;               // developer switch (SLIDEO_L2_DEBUG): 1 = epilogue skips the TMEM drain, 2 = drains but never selects
    uint64_t* scratch;     // [grid][2][L2_BM][L2_SLOTS]
    uint64_t* partial;     // [nq][ns_max][k]   (only rows of split tiles are used)
    int32_t* idx_out;      // [nq][k]
    float* dist_out;       // [nq][k]
};

// work item -> (query tile, pool split, splits of that tile, pool tile range)
struct L2Item { int mt, sp, ns, j0, j1; };
__device__ __forceinline__ L2Item l2_item(const L2Params& P, int item) {
    L2Item w;
    const int items_a = P.mt_a * P.ns_a;
    if (item < items_a) {
        w.ns = P.ns_a;
        w.mt = item / P.ns_a;
        w.sp = item - w.mt * P.ns_a;
    } else {
        const int j = item - items_a;
        w.ns = P.ns_b;
        w.mt = P.mt_a + j / P.ns_b;
        w.sp = j - (j / P.ns_b) * P.ns_b;
    }
    w.j0 = (int)((long long)P.n_ntiles * w.sp / w.ns);
    w.j1 = (int)((long long)P.n_ntiles * (w.sp + 1) / w.ns);
    return w;
}

// Threshold pre-pass.  A work item first runs its leading `pre_tiles` pool tiles through the tensor pipe WITHOUT selecting: every
// epilogue thread (= query row) only folds the accumulator columns it drains into ceil(k/2) group minima (a group = a fixed number
// of consecutive 32-column chunks).  The 2 * ceil(k/2) >= k minima of a row (both column halves) are k distinct pooled descriptors, so
// their maximum t0 bounds the row's k-th distance from above: the item then restarts at its first tile and selects with that
// threshold from the first column on, instead of appending whole tiles until its lists have filled and tightened (about half of
// all candidates a row ever appends fall into its first ~2000 columns).  Cost: pre_tiles extra tiles of pure MMA + min tree.
// Shorter items get a shorter pre-pass: a sixth of their tiles, at least pre_floor (every group minimum must cover a chunk);
// pre_min == pre_floor <= n, so the pre-pass never exceeds the item.
__device__ __forceinline__ int l2_npre(const L2Params& P, int j0, int j1) {
    const int n = j1 - j0;
    return n >= P.pre_min ? min(P.pre_tiles, max(P.pre_floor, n / 6)) : 0;
}

__device__ __forceinline__ void emit_l2_row(uint64_t key, int lane, int q, int k, int32_t* idx_out, float* dist_out) {
    if (lane < k) {
        const bool ok = key != KEY64_EMPTY;
        const float d2 = fmaxf(__uint_as_float((uint32_t)(key >> 32)), 0.f);
        idx_out[(size_t)q * k + lane] = ok ? (int32_t)(uint32_t)key : -1;
        dist_out[(size_t)q * k + lane] = ok ? sqrtf(d2) : -1.f;
    }
}

template <bool DEV>   // DEV: developer instruments (cycle counters, drain-only switches) compiled in
__global__ void __launch_bounds__(L2_THREADS, 1)
knn_l2_kernel(const __grid_constant__ CUtensorMap tm_q_main, const __grid_constant__ CUtensorMap tm_q_tail,
              const __grid_constant__ CUtensorMap tm_t_main, const __grid_constant__ CUtensorMap tm_t_tail, const L2Params P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                          // [main0 | main1 | tail]
    uint8_t* sB = smem + A_BYTES;                // [stage][main0 | main1 | tail]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_OPERANDS);
    uint64_t* a_full = bars + 0;
    uint64_t* a_empty = bars + 1;
    uint64_t* b_full = bars + 2;                 // [L2_STAGES]
    uint64_t* b_empty = bars + 4;                // [L2_STAGES]
    uint64_t* acc_full = bars + 6;               // [2]: TMEM buffer = tile parity, 256 columns each
    uint64_t* acc_empty = bars + 10;             // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 14);
    volatile float* s_tau = reinterpret_cast<volatile float*>(bars + 16);   // [L2_BM]: k-th distance of each row's list as of its last cut-back
    volatile float* s_pre = s_tau + L2_BM;                                  // [2][L2_BM]: pre-pass bound of each column half (l2_npre)
    volatile uint32_t* s_req = reinterpret_cast<volatile uint32_t*>(s_pre + 2 * L2_BM);   // [8][L2_RQ]: cut-back requests of a drain warp (row | region << 5 | candidates << 8)
    volatile uint32_t* s_done = s_req + 8 * L2_RQ;                          // [8][32]: requests the sorter has completed, per list
    volatile uint32_t* s_req_tail = s_done + 8 * 32;                        // [8]: requests a drain warp has posted
    volatile uint32_t* s_items_done = s_req_tail + 8;                       // [4]: work items a sorter warp has emitted

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int s = 0; s < L2_STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < L2_BM; i += L2_THREADS) s_tau[i] = __int_as_float(0x7F800000);
    for (int i = tid; i < 8 * 32 + 8 + 4; i += L2_THREADS) s_done[i] = 0;                       // s_done, s_req_tail, s_items_done
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int n_items = P.mt_a * P.ns_a + (P.n_mtiles - P.mt_a) * P.ns_b;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t it = 0, bcount = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const L2Item w = l2_item(P, item);
                const int mt = w.mt, j0 = w.j0, j1 = w.j1;
                mbar_wait(a_empty, (it & 1) ^ 1);      // MMAs of the previous item no longer read A
                mbar_expect_tx(a_full, A_BYTES);
                tma_load_2d(sA, &tm_q_main, 0, mt * L2_BM, a_full);
                tma_load_2d(sA + A_MAIN_BYTES, &tm_q_main, 64, mt * L2_BM, a_full);
                tma_load_2d(sA + 2 * A_MAIN_BYTES, &tm_q_tail, 0, mt * L2_BM, a_full);
                const int npre = l2_npre(P, j0, j1);
                for (int t = -npre; t < j1 - j0; ++t, ++bcount) {
                    const int j = j0 + (t < 0 ? t + npre : t);       // the pre-pass tiles, then the whole item from its first tile
                    const int s = bcount % L2_STAGES;
                    mbar_wait(&b_empty[s], ((bcount / L2_STAGES) & 1) ^ 1);
                    uint8_t* dst = sB + (size_t)s * B_BYTES;
                    mbar_expect_tx(&b_full[s], B_BYTES);
                    tma_load_2d(dst, &tm_t_main, 0, j * L2_BN, &b_full[s]);
                    tma_load_2d(dst + B_MAIN_BYTES, &tm_t_main, 64, j * L2_BN, &b_full[s]);
                    tma_load_2d(dst + 2 * B_MAIN_BYTES, &tm_t_tail, 0, j * L2_BN, &b_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t it = 0, bcount = 0, acc_use[2] = {0, 0};
            long long mw_acc = 0, mw_b = 0;
            const uint32_t a0 = smem_u32(sA), a1 = a0 + A_MAIN_BYTES, at = a0 + 2 * A_MAIN_BYTES;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const L2Item w = l2_item(P, item);
                const int j0 = w.j0, j1 = w.j1;
                mbar_wait(a_full, it & 1);
                const int n_seq = l2_npre(P, j0, j1) + j1 - j0;
                for (int t = 0; t < n_seq; ++t, ++bcount) {
                    const int b = t & 1;
                    const int s = bcount % L2_STAGES;
                    const long long m0 = DEV && P.prof ? clock64() : 0;
                    mbar_wait(&acc_empty[b], (acc_use[b] & 1) ^ 1);   // all eight drain warps are done with this TMEM buffer
                    ++acc_use[b];
                    const long long m1 = DEV && P.prof ? clock64() : 0;
                    mbar_wait(&b_full[s], (bcount / L2_STAGES) & 1);
                    if (DEV && P.prof) { mw_acc += m1 - m0; mw_b += clock64() - m1; }
                    tc_fence_after();
                    const uint32_t b0 = smem_u32(sB + (size_t)s * B_BYTES), b1 = b0 + B_MAIN_BYTES, bt = b0 + 2 * B_MAIN_BYTES;
                    const uint32_t d = tmem_base + (uint32_t)b * L2_BN;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        tc_mma_bf16(d, smem_desc(a0 + kk * 32, 1024, 2), smem_desc(b0 + kk * 32, 1024, 2), IDESC, kk > 0);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        tc_mma_bf16(d, smem_desc(a1 + kk * 32, 1024, 2), smem_desc(b1 + kk * 32, 1024, 2), IDESC, 1);
                    tc_mma_bf16(d, smem_desc(at, 256, 6), smem_desc(bt, 256, 6), IDESC, 1);
                    tc_commit(&b_empty[s]);        // smem stage reusable once these MMAs have read it
                    tc_commit(&acc_full[b]);       // accumulator ready for the epilogue
                }
                tc_commit(a_empty);
            }
            if (DEV && P.prof) { atomicAdd(P.prof + 4, (unsigned long long)mw_acc); atomicAdd(P.prof + 5, (unsigned long long)mw_b); }
        }
    } else if (warp >= 10) {
        // ===================== sorter warps =====================
        // Sorter s owns the lists of one TMEM lane quarter: ONE sorted list of the k best per query row (slots 0..31 of the row's
        // column-half-0 buffer), fed by the two drain warps of the quarter (s: columns 0..127 of every tile, s + 4: columns
        // 128..255).  A request names a row, the region holding new candidates and their number: the list is merged with the
        // candidates, 32 at a time -- chunk sorted descending, lane-wise min with the ascending list = the 32 smallest as a bitonic
        // sequence -- written back, and its k-th distance published for both drain threads of the row.  At the end of a work item
        // the sorter emits the rows.  The drain warps never sort and never wait for an emission: the tensor pipe does not either.
        const int s = warp - 10;
        const int row0 = ((s + 2) & 3) * 32;             // first row of the quarter
        uint32_t head[2] = {0, 0};
        uint32_t items_done = 0;
        long long ps_busy = 0;
        unsigned ps_req = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const L2Item w = l2_item(P, item);
            unsigned ended = 0, has_list = 0;            // bit r: row row0 + r has a sorted list in this item
            while (ended != 3u) {
                bool worked = false;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (ended >> i & 1u) continue;
                    const int e = s + 4 * i;                     // drain warp e: column half i of this quarter
                    const uint32_t tail = s_req_tail[e];
                    if (head[i] == tail) continue;
                    __threadfence_block();                       // the candidates of every posted request are visible from here on
                    const long long tb0 = DEV && P.prof ? clock64() : 0;
                    worked = true;
                    while (head[i] != tail) {
                        const uint32_t r = s_req[e * L2_RQ + (head[i] & (L2_RQ - 1))];
                        ++head[i];
                        if (r == L2_REQ_END) { ended |= 1u << i; break; }
                        const int rl = (int)(r & 31u), region = (int)(r >> 5 & 1u), n = (int)(r >> 8);
                        uint64_t* list = P.scratch + ((size_t)blockIdx.x * 2 * L2_BM + row0 + rl) * L2_SLOTS;
                        const uint64_t* cand = P.scratch + (((size_t)blockIdx.x * 2 + i) * L2_BM + row0 + rl) * L2_SLOTS + 32 + region * L2_REGION;
                        uint64_t top = has_list >> rl & 1u ? __ldcg(list + lane) : KEY64_EMPTY;   // all three loads in flight together
                        uint64_t x0 = lane < n ? __ldcg(cand + lane) : KEY64_EMPTY;
                        uint64_t x1 = 32 + lane < n ? __ldcg(cand + 32 + lane) : KEY64_EMPTY;
                        if (n > 0) top = warp_top32_merge(top, x0, lane);
                        if (n > 32) top = warp_top32_merge(top, x1, lane);
                        for (int c0 = 64; c0 < n; c0 += 32) {
                            const uint64_t x = c0 + lane < n ? __ldcg(cand + c0 + lane) : KEY64_EMPTY;
                            top = warp_top32_merge(top, x, lane);
                        }
                        if (lane >= P.k) top = KEY64_EMPTY;
                        list[lane] = top;
                        has_list |= 1u << rl;
                        const uint64_t kth = __shfl_sync(FULL, top, P.k - 1);
                        if (lane == 0) {
                            if (kth != KEY64_EMPTY) s_tau[row0 + rl] = __uint_as_float((uint32_t)(kth >> 32));
                            s_done[e * 32 + rl] = s_done[e * 32 + rl] + 1;   // the region may be appended to again
                        }
                        ++ps_req;
                    }
                    if (DEV && P.prof) ps_busy += clock64() - tb0;
                }
                if (!worked) __nanosleep(40);
            }
            // both drain warps have finished the item: emit the rows (oracle order: the list is sorted by (distance, index))
            for (int L = 0; L < 32; ++L) {
                const int qq = w.mt * L2_BM + row0 + L;
                if (qq >= P.nq) break;   // warp-uniform
                const uint64_t* list = P.scratch + ((size_t)blockIdx.x * 2 * L2_BM + row0 + L) * L2_SLOTS;
                const uint64_t m = has_list >> L & 1u ? __ldcg(list + lane) : KEY64_EMPTY;
                if (w.ns == 1) emit_l2_row(m, lane, qq, P.k, P.idx_out, P.dist_out);
                else if (lane < P.k) P.partial[((size_t)qq * P.ns_max + w.sp) * P.k + lane] = m;
            }
            s_tau[row0 + lane] = __int_as_float(0x7F800000);   // the next item starts unpruned
            __threadfence_block();
            __syncwarp();
            ++items_done;
            if (lane == 0) s_items_done[s] = items_done;       // the drain warps may read s_tau for the next item
        }
        if (DEV && P.prof && lane == 0) { atomicAdd(P.prof + 6, (unsigned long long)ps_busy); atomicAdd(P.prof + 7, (unsigned long long)ps_req); }
    } else {
        // ===================== drain warps (epilogue groups) =====================
        const int e = warp - 2;                      // drain warp
        const int g = e >> 2;                        // 0: columns 0..127 of every tile, 1: columns 128..255
        const int quarter = warp & 3;                // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;         // query row within the tile
        uint64_t* my_buf = P.scratch + (((size_t)blockIdx.x * 2 + g) * L2_BM + row) * L2_SLOTS;
        const uint32_t taddr_group = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)g * L2_BNH;
        uint32_t acc_phase = 0;                      // bit b: parity of the next completion of acc_full[b] this thread waits for
        const uint32_t sa_tau = smem_u32(const_cast<float*>(s_tau)) + (uint32_t)row * 4;
        const uint32_t sa_items_done = smem_u32(const_cast<uint32_t*>(s_items_done)) + (uint32_t)(e & 3) * 4;
        uint32_t items_begun = 0;                    // work items this warp has started before the current one
        const uint32_t sa_done = smem_u32(const_cast<uint32_t*>(s_done)) + (uint32_t)(e * 32 + lane) * 4;
        uint32_t n_posted = 0;                       // requests posted by this thread (s_done counts the completed ones)
        int region = 0;                              // region this thread appends to (survives work items: the other one may still be queued)
        uint32_t req_tail = 0;                       // requests posted by this warp (warp-uniform)
        long long pw = 0, pd = 0, pc = 0, pt = 0;

        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const L2Item w = l2_item(P, item);
            const int mt = w.mt, j0 = w.j0, j1 = w.j1;
            const int q = mt * L2_BM + row;
            float bound = q < P.nq ? __int_as_float(0x7F800000) : -1.f;   // pre-pass bound: +inf / never
            int cnt = 0;               // candidates in the region this thread appends to

            // hands the regions that passed the trigger (force: every region that holds anything, then the end-of-item marker) to
            // the sorter warp and switches the thread to its other region
            auto post = [&](bool force) {
                bool want = force ? cnt > 0 : cnt > P.trigger;
                if (want && lds_volatile_u32(sa_done) != n_posted) {
                    // the thread's previous request is still in the sorter's queue: wait only if the region cannot take another
                    // half tile (or at the end of the item), else try again after the next tile
                    if (force || cnt > L2_TRIGGER) { while (lds_volatile_u32(sa_done) != n_posted) {} }
                    else want = false;
                }
                const unsigned m = __ballot_sync(FULL, want);
                if (m == 0 && !force) return;
                if (want) {
                    s_req[e * L2_RQ + ((req_tail + __popc(m & ((1u << lane) - 1u))) & (L2_RQ - 1))] =
                        (uint32_t)lane | (uint32_t)region << 5 | (uint32_t)cnt << 8;
                    ++n_posted;
                    region ^= 1;
                    cnt = 0;
                }
                req_tail += __popc(m);
                if (force) {
                    if (lane == 0) s_req[e * L2_RQ + (req_tail & (L2_RQ - 1))] = L2_REQ_END;
                    ++req_tail;
                }
                __threadfence_block();     // this thread's appended candidates and its request before the tail moves
                __syncwarp();
                if (lane == 0) s_req_tail[e] = req_tail;
            };

            // ---- threshold pre-pass (l2_npre): group minima only, nothing is appended
            const int npre = l2_npre(P, j0, j1);
            if (npre) {
                const float inf = __int_as_float(0x7F800000);
                const int n_groups = (P.k + 1) / 2;
                const int per_group = npre * (L2_BNH / 32) / n_groups;   // 32-column chunks folded into one minimum (>= 1: host)
                float gmin = inf, t0 = -inf;
                int in_group = 0, groups = 0;
                for (int t = 0; t < npre; ++t) {
                    const int b = t & 1;
                    mbar_wait(&acc_full[b], (acc_phase >> b) & 1);
                    acc_phase ^= 1u << b;
                    tc_fence_after();
                    const uint32_t taddr_base = taddr_group + (uint32_t)b * L2_BN;
                    uint32_t va[32], vb[32];
                    tc_ld32(taddr_base, va);
#pragma unroll 1
                    for (int c = 0; c < L2_BNH / 32; c += 2) {
                        tc_ld_wait();
                        tc_ld32(taddr_base + (uint32_t)(c + 1) * 32, vb);
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            uint32_t (&v)[32] = half == 0 ? va : vb;
                            if (half == 1) {
                                tc_ld_wait();
                                if (c + 2 < L2_BNH / 32) tc_ld32(taddr_base + (uint32_t)(c + 2) * 32, va);
                            }
                            float m = inf;
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                m = fminf(m, fminf(fminf(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])),
                                                   fminf(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]))));
                            if (groups < n_groups && m < inf) {   // (a chunk of nothing but padding columns does not count)
                                gmin = fminf(gmin, m);
                                if (++in_group == per_group) { t0 = fmaxf(t0, gmin); gmin = inf; in_group = 0; ++groups; }
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[b]);
                }
                s_pre[g * L2_BM + row] = groups == n_groups ? t0 : inf;   // too few real columns: no bound
                asm volatile("bar.sync 1, 256;" ::: "memory");   // the next write of s_pre lies behind the two barriers of the item tail
                // k distinct pooled descriptors of this item lie at or below the larger half bound: candidates above it are never needed
                if (q < P.nq) bound = next_up(fmaxf(s_pre[row], s_pre[L2_BM + row]));
            }

            // s_tau belongs to the previous item until the sorter has emitted it (it is a pre-pass or a whole item behind at most)
            while (lds_volatile_u32(sa_items_done) != items_begun) {}
            ++items_begun;

            for (int j = j0; j < j1; ++j) {
                // both groups drain EVERY tile, half of its columns each: the accumulator goes back to the tensor pipe after
                // half a drain, and the MMA of tile j + 2 never queues behind a whole-tile epilogue
                const int b = (npre + j - j0) & 1;
                const long long t0 = DEV && P.prof ? clock64() : 0;
                mbar_wait(&acc_full[b], (acc_phase >> b) & 1);
                acc_phase ^= 1u << b;
                tc_fence_after();
                const long long t1 = DEV && P.prof ? clock64() : 0;
                const int col0 = j * L2_BN + g * L2_BNH;
                const uint32_t taddr_base = taddr_group + (uint32_t)b * L2_BN;
                // effective threshold: the pre-pass bound or the k-th distance of the row's list as of its last cut-back, whichever is
                // tighter -- non-strict (next_up): the list holds candidates of both column halves, an equal distance with a smaller
                // index could still displace its k-th entry; the sorter decides on the full (distance, index) key
                const float thr = fminf(bound, next_up(lds_volatile(sa_tau)));
                uint64_t* app = my_buf + 32 + region * L2_REGION;
                const int n_chunks = DEV && P.dbg == 1 ? 0 : L2_BNH / 32;
                uint32_t va[32], vb[32];
                if (n_chunks) tc_ld32(taddr_base, va);
#pragma unroll 1
                for (int c = 0; c < n_chunks; c += 2) {
                    // software pipeline: the next 32 columns are in flight while these are compared
                    tc_ld_wait();
                    tc_ld32(taddr_base + (uint32_t)(c + 1) * 32, vb);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t (&v)[32] = half == 0 ? va : vb;
                        if (half == 1) {
                            tc_ld_wait();
                            if (c + 2 < n_chunks) tc_ld32(taddr_base + (uint32_t)(c + 2) * 32, va);
                        }
                        // fast path: a 3-input-min tree per 4 columns, one compare per group
                        bool hit[8];
                        bool any = false;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float m = fminf(fminf(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])),
                                                  fminf(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])));
                            hit[i] = m < thr;
                            any |= hit[i];
                        }
                        if (DEV && P.dbg == 2) any = any && v[0] == 0x12345678u;
                        if (any) {
                            // slow path (a fifth of the chunks, two threads of the warp on average): ONE divergent region; the hit
                            // group's four values are picked by a select tree instead of a branch per group
                            unsigned hm = 0;
#pragma unroll
                            for (int i = 0; i < 8; ++i) hm |= hit[i] ? 1u << i : 0u;
                            const uint32_t cbase = (uint32_t)(col0 + (c + half) * 32);
                            do {
                                const int gi = __ffs(hm) - 1;
                                hm &= hm - 1;
                                const uint32_t col = cbase + 4u * (uint32_t)gi;
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    uint32_t x = v[28 + u];
#pragma unroll
                                    for (int i = 6; i >= 0; --i) x = gi == i ? v[4 * i + u] : x;
                                    const float val = __uint_as_float(x);
                                    asm volatile(
                                        "{\n\t.reg .pred p;\n\t"
                                        "setp.lt.f32 p, %0, %1;\n\t"
                                        "@p st.global.v2.u32 [%2], {%3, %4};\n\t}"
                                        ::"f"(val), "f"(thr), "l"(app + cnt), "r"(col + (uint32_t)u), "r"(x)
                                        : "memory");
                                    cnt += val < thr ? 1 : 0;
                                }
                            } while (hm);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[b]);
                const long long t2 = DEV && P.prof ? clock64() : 0;
                // the accumulator is back with the tensor pipe: hand the regions that passed the trigger to the sorter
                post(false);
                if (DEV && P.prof && lane == 0) {
                    const long long t3 = clock64();
                    pw += t1 - t0; pd += t2 - t1; pc += t3 - t2;
                }
            }

            // end of item: the rest of the candidates and the end marker go to the sorter, which emits the rows; the drain warps
            // go straight on to the next item
            const long long t4 = DEV && P.prof ? clock64() : 0;
            __syncwarp();
            post(true);
            if (DEV && P.prof && lane == 0) pt += clock64() - t4;
        }
        if (DEV && P.prof && lane == 0) {
            atomicAdd(P.prof + 0, (unsigned long long)pw); atomicAdd(P.prof + 1, (unsigned long long)pd);
            atomicAdd(P.prof + 2, (unsigned long long)pc); atomicAdd(P.prof + 3, (unsigned long long)pt);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// merge of per-split partial rows: one warp per query
__global__ void __launch_bounds__(128) l2_merge_kernel(const uint64_t* __restrict__ partial, int q_begin, int q_end, int n_splits, int ns_max,
                                                       int k, int32_t* idx_out, float* dist_out) {
    const int lane = threadIdx.x & 31;
    const int q = q_begin + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= q_end) return;
    uint64_t top = KEY64_EMPTY;   // ascending; every partial row is ascending too, so it is read back to front
    for (int s = 0; s < n_splits; ++s) {
        const uint64_t x = 31 - lane < k ? partial[((size_t)q * ns_max + s) * k + (31 - lane)] : KEY64_EMPTY;
        top = warp_merge32_asc(top < x ? top : x, lane);
    }
    emit_l2_row(top, lane, q, k, idx_out, dist_out);
}

// fp32 rows -> bf16 GEMM operands.  One warp per row, 4 elements per lane.
//   queries: main = bf16(-2 x), tail = [1, 1, 1, n_hi, n_mid, n_lo, 0...]
//   pool   : main = bf16(x),    tail = [n_hi, n_mid, n_lo, 1, 1, 1, 0...];  padding rows: main 0, n_hi = +inf
// n = sum of squares of the ROUNDED values, split into three bf16 pieces (exact to 24 bits).
__global__ void __launch_bounds__(256) l2_prepare_kernel(const float* __restrict__ src, int n, int n_pad, int is_query,
                                                         __nv_bfloat16* __restrict__ out_main, __nv_bfloat16* __restrict__ out_tail) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_pad) return;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < n) x = __ldg(reinterpret_cast<const float4*>(src + (size_t)row * L2_DIM) + lane);
    const float sc = is_query ? -2.f : 1.f;
    const __nv_bfloat16 b0 = __float2bfloat16_rn(sc * x.x), b1 = __float2bfloat16_rn(sc * x.y), b2 = __float2bfloat16_rn(sc * x.z),
                        b3 = __float2bfloat16_rn(sc * x.w);
    const float inv = is_query ? -0.5f : 1.f;
    const float r0 = __bfloat162float(b0) * inv, r1 = __bfloat162float(b1) * inv, r2 = __bfloat162float(b2) * inv, r3 = __bfloat162float(b3) * inv;
    float nn = __fmaf_rn(r3, r3, __fmaf_rn(r2, r2, __fmaf_rn(r1, r1, __fmul_rn(r0, r0))));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(FULL, nn, o);
    __nv_bfloat162* m2 = reinterpret_cast<__nv_bfloat162*>(out_main + (size_t)row * L2_DIM) + lane * 2;
    m2[0] = __nv_bfloat162(b0, b1);
    m2[1] = __nv_bfloat162(b2, b3);
    if (lane < L2_TAIL) {
        const __nv_bfloat16 hi = __float2bfloat16_rn(nn);
        const float e1 = nn - __bfloat162float(hi);
        const __nv_bfloat16 mid = __float2bfloat16_rn(e1);
        const float e2 = e1 - __bfloat162float(mid);
        const __nv_bfloat16 lo = __float2bfloat16_rn(e2);
        const __nv_bfloat16 one = __float2bfloat16_rn(1.f), zero = __float2bfloat16_rn(0.f);
        __nv_bfloat16 v = zero;
        if (is_query) {
            if (lane < 3) v = one;
            else if (lane == 3) v = hi;
            else if (lane == 4) v = mid;
            else if (lane == 5) v = lo;
        } else {
            if (row >= n) v = lane == 0 ? __ushort_as_bfloat16((unsigned short)0x7F80) : (lane >= 3 && lane < 6 ? one : zero);   // +inf norm
            else if (lane == 0) v = hi;
            else if (lane == 1) v = mid;
            else if (lane == 2) v = lo;
            else if (lane < 6) v = one;
        }
        out_tail[(size_t)row * L2_TAIL + lane] = v;
    }
}

// the reference vote (lib.rs:270-282) on float rows: one warp per query
__global__ void __launch_bounds__(128) l2_vote_kernel(const int32_t* __restrict__ idx, const float* __restrict__ dist, int nq, int k,
                                                      const int32_t* __restrict__ q_frame, const uint16_t* __restrict__ page_of,
                                                      int32_t* __restrict__ votes, int n_pages, float ratio) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const int i = lane < k ? idx[(size_t)q * k + lane] : -1;
    const float d = lane < k ? dist[(size_t)q * k + lane] : 0.f;
    const float best = __shfl_sync(FULL, d, 0);
    const int i0 = __shfl_sync(FULL, i, 0);
    if (i0 >= 0 && i >= 0 && d < __fmul_rn(best, ratio)) atomicAdd(&votes[(size_t)q_frame[q] * n_pages + page_of[i]], 1);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        SLIDEO_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !p) throw CudaError(cudaErrorNotSupported, "cuTensorMapEncodeTiled is not available in this driver");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

CUtensorMap make_map(const void* base, int inner, int rows, int box_inner, int box_rows, CUtensorMapSwizzle sw) {
    CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)inner * 2};
    const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaError(cudaErrorInvalidValue, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return tm;
}

void grow(void** p, size_t* cap, size_t bytes) {
    if (bytes <= *cap) return;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    SLIDEO_CUDA(cudaMalloc(p, want));
    *cap = want;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
L2Workspace::~L2Workspace() {
    if (d_q_main) cudaFree(d_q_main);
    if (d_q_tail) cudaFree(d_q_tail);
    if (d_scratch) cudaFree(d_scratch);
    if (d_part) cudaFree(d_part);
}

int l2_rows_padded(int n) { return (std::max(n, 1) + L2_BN - 1) / L2_BN * L2_BN; }
size_t l2_main_bytes(int n) { return (size_t)l2_rows_padded(n) * L2_DIM * 2; }
size_t l2_tail_bytes(int n) { return (size_t)l2_rows_padded(n) * L2_TAIL * 2; }

void l2_prepare_launch(const float* d_src, int n, bool is_query, void* d_main, void* d_tail, cudaStream_t stream) {
    const int n_pad = l2_rows_padded(n);
    l2_prepare_kernel<<<cdiv(n_pad, 8), 256, 0, stream>>>(d_src, n, n_pad, is_query ? 1 : 0, (__nv_bfloat16*)d_main, (__nv_bfloat16*)d_tail);
    SLIDEO_CUDA(cudaGetLastError());
}

namespace {
// developer knobs (profiling counters, split / trigger experiments): read from the environment ONCE per process, never on the
// launch path
struct L2Env { bool no_split, prof; int trigger, dbg, pre; };
const L2Env& l2_env() {
    static const L2Env e = [] {
        L2Env v;
        v.no_split = getenv("SLIDEO_L2_NO_SPLIT") != nullptr;
        v.prof = getenv("SLIDEO_L2_PROF") != nullptr;
        v.trigger = getenv("SLIDEO_L2_TRIGGER") ? std::min(L2_TRIGGER, std::max(1, atoi(getenv("SLIDEO_L2_TRIGGER")))) : 26;
        v.dbg = getenv("SLIDEO_L2_DEBUG") ? atoi(getenv("SLIDEO_L2_DEBUG")) : 0;
        v.pre = getenv("SLIDEO_L2_PRE") ? std::max(0, atoi(getenv("SLIDEO_L2_PRE"))) : 32;
        return v;
    }();
    return e;
}
}  // namespace

void l2_knn_launch(L2Workspace& ws, const float* d_q, int nq, const void* d_pool_main, const void* d_pool_tail, int nt, int k,
                   int32_t* d_idx, float* d_dist, int num_sms, cudaStream_t stream, int* launches) {
    if (nq <= 0) return;
    // queries -> bf16 operands
    grow(&ws.d_q_main, &ws.q_main_cap, l2_main_bytes(nq));
    grow(&ws.d_q_tail, &ws.q_tail_cap, l2_tail_bytes(nq));
    l2_prepare_launch(d_q, nq, true, ws.d_q_main, ws.d_q_tail, stream);
    if (launches) ++*launches;

    L2Params P;
    P.nq = nq; P.nt = nt; P.k = k;
    P.n_mtiles = cdiv(nq, L2_BM);
    P.n_ntiles = l2_rows_padded(nt) / L2_BN;
    // work items = query tiles (x pool splits) on a persistent grid of one CTA per SM.  Splitting a tile's pool restarts its
    // selection per part (measured: 921 -> 855 TFLOP/s when every tile is split in two), so tiles are split only
    //  * when there are too few of them to fill the machine (all tiles, ns_a ways), or
    //  * in the LAST wave: the n_mtiles % SMs tiles that would leave most SMs idle are split so that their parts fill one wave.
    // A part keeps at least pre_floor tiles where the pool allows it, so that it can run the threshold pre-pass: without a bound
    // the first tiles of an item append every column and the item is paced by its sorter warps.
    P.pre_floor = ((k + 1) / 2 + L2_BNH / 32 - 1) / (L2_BNH / 32);
    const int max_parts = std::max(1, P.n_ntiles / P.pre_floor);
    // parts per tile for `tiles` query tiles that share the machine: rounds x (pool / parts + the fixed cost of a part) is
    // minimised.  The fixed cost -- pre-pass, restarted selection, the sorters' end-of-item requests, emission -- is worth ~350
    // tiles (calibrated: 79 tiles x 3907 gain nothing from 9 parts, 98 tiles run 18 % faster as 294 thirds in two rounds than
    // as one round that leaves 50 SMs idle, 79 x 391 lose 70 % when cut in five)
    auto choose_parts = [&](int tiles, int limit) {
        int best = 1;
        double best_cost = 1e300;
        for (int c = 1; c <= std::min(limit, max_parts); ++c) {
            const double rounds = (double)cdiv(tiles * c, num_sms);
            const double cost = rounds * ((double)P.n_ntiles / c + std::min(32.0, std::max(4.0, P.n_ntiles / (6.0 * c))) + 350.0);
            if (cost < best_cost * 0.97) { best_cost = cost; best = c; }   // a finer split must pay for restarting the selection
        }
        return best;
    };
    P.mt_a = P.n_mtiles; P.ns_a = 1; P.ns_b = 1;
    if (P.n_mtiles < num_sms) {
        P.ns_a = choose_parts(P.n_mtiles, 64);
    } else if (!l2_env().no_split) {
        const int r = P.n_mtiles % num_sms;
        const int c = r > 0 ? choose_parts(r, 8) : 1;
        if (c >= 2) { P.mt_a = P.n_mtiles - r; P.ns_b = c; }
    }
    P.ns_max = std::max(P.ns_a, P.ns_b);
    const int n_items = P.mt_a * P.ns_a + (P.n_mtiles - P.mt_a) * P.ns_b;
    const int grid = std::min(n_items, num_sms);
    grow(&ws.d_scratch, &ws.scratch_cap, (size_t)grid * 2 * L2_BM * L2_SLOTS * 8);
    if (P.ns_max > 1) grow(&ws.d_part, &ws.part_cap, (size_t)nq * P.ns_max * k * 8);
    P.scratch = (uint64_t*)ws.d_scratch;
    P.partial = (uint64_t*)ws.d_part;
    P.idx_out = d_idx;
    P.dist_out = d_dist;
    P.trigger = l2_env().trigger;
    // pre-pass: a sixth of an item's tiles, pre_floor..pre_tiles of them, for items of at least pre_floor tiles (a 4-tile item runs
    // twice: the tensor pipe is not what bounds it); every group minimum must cover at least one 32-column chunk (a thread drains
    // L2_BNH / 32 chunks of a tile)
    P.pre_tiles = std::max(l2_env().pre, P.pre_floor);
    P.pre_min = l2_env().pre > 0 ? P.pre_floor : 0x7FFFFFFF;
    P.prof = nullptr;
    P.dbg = l2_env().dbg;

    const int q_pad = l2_rows_padded(nq), t_pad = l2_rows_padded(nt);
    const CUtensorMap tq_main = make_map(ws.d_q_main, L2_DIM, q_pad, 64, L2_BM, CU_TENSOR_MAP_SWIZZLE_128B);
    const CUtensorMap tq_tail = make_map(ws.d_q_tail, L2_TAIL, q_pad, L2_TAIL, L2_BM, CU_TENSOR_MAP_SWIZZLE_32B);
    const CUtensorMap tt_main = make_map(d_pool_main, L2_DIM, t_pad, 64, L2_BN, CU_TENSOR_MAP_SWIZZLE_128B);
    const CUtensorMap tt_tail = make_map(d_pool_tail, L2_TAIL, t_pad, L2_TAIL, L2_BN, CU_TENSOR_MAP_SWIZZLE_32B);

    // function attributes are per device: set on every launch (cheap) so that ctxs on several GPUs of one process all get them
    const bool dev = l2_env().prof || l2_env().dbg != 0;
    SLIDEO_CUDA(cudaFuncSetAttribute(dev ? knn_l2_kernel<true> : knn_l2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL));
    static unsigned long long* d_prof = nullptr;
    if (l2_env().prof) {
        if (!d_prof) SLIDEO_CUDA(cudaMalloc(&d_prof, 8 * sizeof(unsigned long long)));
        SLIDEO_CUDA(cudaMemsetAsync(d_prof, 0, 8 * sizeof(unsigned long long), stream));
        P.prof = d_prof;
    }
    if (dev) knn_l2_kernel<true><<<grid, L2_THREADS, SMEM_TOTAL, stream>>>(tq_main, tq_tail, tt_main, tt_tail, P);
    else knn_l2_kernel<false><<<grid, L2_THREADS, SMEM_TOTAL, stream>>>(tq_main, tq_tail, tt_main, tt_tail, P);
    SLIDEO_CUDA(cudaGetLastError());
    if (P.prof) {
        unsigned long long h[8];
        SLIDEO_CUDA(cudaStreamSynchronize(stream));
        SLIDEO_CUDA(cudaMemcpy(h, d_prof, sizeof h, cudaMemcpyDeviceToHost));
        const double ew = 8.0 * grid, mw = 1.0 * grid;   // epilogue warps, MMA warps
        fprintf(stderr, "[l2 prof] per drain warp (Mcycles): acc wait %.2f drain %.2f post %.2f tail %.2f | per MMA warp: acc_empty wait %.2f b_full wait %.2f | per sorter warp: busy %.2f, %.0f requests\n",
                h[0] / ew / 1e6, h[1] / ew / 1e6, h[2] / ew / 1e6, h[3] / ew / 1e6, h[4] / mw / 1e6, h[5] / mw / 1e6, h[6] / (4.0 * grid) / 1e6, h[7] / (4.0 * grid));
    }
    if (launches) ++*launches;
    // merge of the partial rows of the split tiles (one warp per query)
    if (P.ns_a > 1) {
        const int q1 = std::min(nq, P.mt_a * L2_BM);
        l2_merge_kernel<<<cdiv(q1, 4), 128, 0, stream>>>((const uint64_t*)ws.d_part, 0, q1, P.ns_a, P.ns_max, k, d_idx, d_dist);
        SLIDEO_CUDA(cudaGetLastError());
        if (launches) ++*launches;
    }
    if (P.ns_b > 1 && P.mt_a * L2_BM < nq) {
        const int q0 = P.mt_a * L2_BM;
        l2_merge_kernel<<<cdiv(nq - q0, 4), 128, 0, stream>>>((const uint64_t*)ws.d_part, q0, nq, P.ns_b, P.ns_max, k, d_idx, d_dist);
        SLIDEO_CUDA(cudaGetLastError());
        if (launches) ++*launches;
    }
}

void l2_vote_launch(const int32_t* d_idx, const float* d_dist, int nq, int k, const int32_t* d_q_frame, const uint16_t* d_page_of,
                    int32_t* d_votes, int n_pages, float ratio, cudaStream_t stream) {
    if (nq <= 0) return;
    l2_vote_kernel<<<cdiv(nq, 4), 128, 0, stream>>>(d_idx, d_dist, nq, k, d_q_frame, d_page_of, d_votes, n_pages, ratio);
    SLIDEO_CUDA(cudaGetLastError());
}

}  // namespace slideo
