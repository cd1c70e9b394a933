// knn_dev.cuh -- device helpers shared by the two Hamming k-NN kernels (knn_hamming.cu: K8 v4, knn_hamming5.cu: K8 v5):
// mbarrier / 1-D TMA bulk copy wrappers, LOP3 with an explicit truth table, the 64-key warp bitonic sort and the row
// emission with the fused K9 vote (crates/matching-opencv/src/lib.rs:270-282).
#pragma once
#include "common.cuh"

namespace slideo {

constexpr unsigned KNN_FULL = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t knn_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void knn_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(knn_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void knn_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(knn_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool knn_mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(knn_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void knn_mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!knn_mbar_try_wait(bar, parity)) {}
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void knn_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     knn_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(knn_smem_u32(bar))
                 : "memory");
}

template <int LUT>
__device__ __forceinline__ uint32_t knn_lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return r;
}
constexpr int KNN_LUT_XOR3 = 0x96;  // a ^ b ^ c
constexpr int KNN_LUT_MAJ = 0xE8;   // majority(a, b, c)
constexpr int KNN_LUT_CARRY = 0xD4; // majority(a, b, a ^ b ^ c): carry of a full adder given two inputs and the sum

// 64-key ascending bitonic sort across a warp: position p = lane (k0) and 32 + lane (k1).
__device__ __forceinline__ void knn_warp_sort64(uint32_t& k0, uint32_t& k1, int lane) {
#pragma unroll
    for (int size = 2; size <= 64; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride == 32) {
                uint32_t lo = min(k0, k1), hi = max(k0, k1);
                k0 = lo;
                k1 = hi;
            } else {
                uint32_t o0 = __shfl_xor_sync(KNN_FULL, k0, stride), o1 = __shfl_xor_sync(KNN_FULL, k1, stride);
                bool lower = (lane & stride) == 0;
                bool up0 = size == 64 ? true : (size == 32 ? true : (lane & size) == 0);
                bool up1 = size == 64 ? true : (size == 32 ? false : (lane & size) == 0);
                k0 = (lower == up0) ? min(k0, o0) : max(k0, o0);
                k1 = (lower == up1) ? min(k1, o1) : max(k1, o1);
            }
        }
    }
}

// Row emission shared by K8 (single-segment tiles) and the merge kernels: lane m < k holds neighbour m (sorted).
__device__ __forceinline__ void knn_emit_row(uint32_t key, int lane, int q, int k, uint32_t* keys_out, const VoteArgs& v) {
    if (keys_out != nullptr && lane < k) keys_out[(size_t)q * k + lane] = key;
    if (v.votes != nullptr) {
        uint32_t best = __shfl_sync(KNN_FULL, key, 0);
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

}  // namespace slideo
