// k8v5_probe.cu -- feasibility probe for a bit-sliced ("vertical") formulation of K8 (brute-force Hamming k-NN):
//   pool chunk bit-sliced in shared memory: word P[b][j] holds bit b of the 32 pooled rows 32*j .. 32*j+31;
//   a warp takes ONE query at a time (warp-uniform), walks the list of the query's set bits and adds the words P[b][.]
//   of those bits with a Harley-Seal carry-save tree: c = |q & t| for 32*W rows per lane, no POPC at all;
//   d = popc(q) + popc(t) - 2c is compared against the query's running k-th distance entirely in bit-sliced form.
// Prints pairs/s for W = 1, 2, 4 words per lane and checks masks + distances against a popcount reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/scratch/k8v5_probe.bin tools/scratch/k8v5_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s failed: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int NQ_CTA = 128;      // queries per CTA
constexpr int WARPS = 16;
constexpr int LIST = 128;        // list entries per query (set bits of q or of ~q, padded with the zero row 256)
constexpr int META = 16;         // per query: [0..9] carry masks of C = 1023 - K, [10] invert mask, [11] A, [12] tau

template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return r;
}
constexpr int XOR3 = 0x96, MAJ = 0xE8;

template <int W> struct Vec { uint32_t v[W]; };

template <int W>
__device__ __forceinline__ Vec<W> lds(uint32_t addr) {
    Vec<W> r;
    if constexpr (W == 1) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r.v[0]) : "r"(addr));
    else if constexpr (W == 2) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.v[0]), "=r"(r.v[1]) : "r"(addr));
    else asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]) : "r"(addr));
    return r;
}

// full adder on bit planes: (sum, carry) of a + b + c
template <int W>
__device__ __forceinline__ void csa(Vec<W>& sum, Vec<W>& carry, const Vec<W>& a, const Vec<W>& b, const Vec<W>& c) {
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const uint32_t x = a.v[w], y = b.v[w], z = c.v[w];
        sum.v[w] = lop3<XOR3>(x, y, z);
        carry.v[w] = lop3<MAJ>(x, y, z);
    }
}

__device__ __forceinline__ int imad(int a, int b, int c) {
    int r;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

template <int W>
__global__ void __launch_bounds__(WARPS * 32, 1) probe_kernel(const uint32_t* __restrict__ gP, const uint32_t* __restrict__ gpt,
                                                              const uint32_t* __restrict__ glist, const uint32_t* __restrict__ gmeta,
                                                              int iters, uint32_t* __restrict__ out_mask, uint32_t* __restrict__ out_planes,
                                                              unsigned long long* __restrict__ out_hits) {
    extern __shared__ __align__(128) uint32_t smem[];
    constexpr int ROWW = W * 32;                 // words per bit row
    uint32_t* sP = smem;                         // [257][ROWW]
    uint32_t* spt = sP + 257 * ROWW;             // [9][ROWW]
    uint32_t* slist = spt + 9 * ROWW;            // [NQ_CTA][LIST]
    uint32_t* smeta = slist + NQ_CTA * LIST;     // [NQ_CTA][META]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 257 * ROWW; i += blockDim.x) sP[i] = gP[i];
    for (int i = tid; i < 9 * ROWW; i += blockDim.x) spt[i] = gpt[i];
    for (int i = tid; i < NQ_CTA * LIST; i += blockDim.x) slist[i] = glist[i];
    for (int i = tid; i < NQ_CTA * META; i += blockDim.x) smeta[i] = gmeta[i];
    __syncthreads();

    const uint32_t sP_addr = (uint32_t)__cvta_generic_to_shared(sP);
    const int lanebase = (int)sP_addr + lane * W * 4;
    // pt' planes of this lane's rows: constant over the queries of a chunk
    Vec<W> ptp[9];
#pragma unroll
    for (int p = 0; p < 9; ++p)
#pragma unroll
        for (int w = 0; w < W; ++w) ptp[p].v[w] = spt[p * ROWW + lane * W + w];

    unsigned long long hits = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int qi = 0; qi < NQ_CTA / WARPS; ++qi) {
            const int q = warp * (NQ_CTA / WARPS) + qi;
            const uint4* lst = reinterpret_cast<const uint4*>(slist + q * LIST);
            // Harley-Seal accumulators seeded with pt' >> 1: total = 2c + pt'
            Vec<W> ones = ptp[1], twos = ptp[2], fours = ptp[3], eights = ptp[4], s16 = ptp[5], s32 = ptp[6], s64 = ptp[7], s128 = ptp[8];
            Vec<W> s256;
#pragma unroll
            for (int w = 0; w < W; ++w) s256.v[w] = 0;
            Vec<W> o_prev, t32a, u64a;
#pragma unroll
            for (int blk = 0; blk < 8; ++blk) {
                const uint4 e0 = lst[blk * 4], e1 = lst[blk * 4 + 1], e2 = lst[blk * 4 + 2], e3 = lst[blk * 4 + 3];
                const uint32_t ent[16] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w, e2.x, e2.y, e2.z, e2.w, e3.x, e3.y, e3.z, e3.w};
                Vec<W> x[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = lds<W>((uint32_t)imad((int)ent[i], W * 128, lanebase));
                Vec<W> ta, tb, fa, fb, ea, eb, o;
                csa<W>(ones, ta, ones, x[0], x[1]);
                csa<W>(ones, tb, ones, x[2], x[3]);
                csa<W>(twos, fa, twos, ta, tb);
                csa<W>(ones, ta, ones, x[4], x[5]);
                csa<W>(ones, tb, ones, x[6], x[7]);
                csa<W>(twos, fb, twos, ta, tb);
                csa<W>(fours, ea, fours, fa, fb);
                csa<W>(ones, ta, ones, x[8], x[9]);
                csa<W>(ones, tb, ones, x[10], x[11]);
                csa<W>(twos, fa, twos, ta, tb);
                csa<W>(ones, ta, ones, x[12], x[13]);
                csa<W>(ones, tb, ones, x[14], x[15]);
                csa<W>(twos, fb, twos, ta, tb);
                csa<W>(fours, eb, fours, fa, fb);
                csa<W>(eights, o, eights, ea, eb);
                // second level over the 8 weight-16 outputs
                if (blk & 1) {
                    Vec<W> t;
                    csa<W>(s16, t, s16, o_prev, o);
                    if ((blk & 3) == 3) {
                        Vec<W> u;
                        csa<W>(s32, u, s32, t32a, t);
                        if (blk == 7) {
                            Vec<W> v;
                            csa<W>(s64, v, s64, u64a, u);
#pragma unroll
                            for (int w = 0; w < W; ++w) { s256.v[w] = s128.v[w] & v.v[w]; s128.v[w] ^= v.v[w]; }
                        } else u64a = u;
                    } else t32a = t;
                } else o_prev = o;
            }
            // s = 2c + pt' as planes {ptp[0], ones, twos, fours, eights, s16, s32, s64, s128, s256}; pass iff s + C >= 1024
            const uint4* mt = reinterpret_cast<const uint4*>(smeta + q * META);
            const uint4 m0 = mt[0], m1 = mt[1], m2 = mt[2];
            uint32_t res[W];
#pragma unroll
            for (int w = 0; w < W; ++w) {
                uint32_t cy = ptp[0].v[w] & m0.x;                 // carry out of bit 0 (carry in = 0)
                cy = lop3<MAJ>(ones.v[w], m0.y, cy);
                cy = lop3<MAJ>(twos.v[w], m0.z, cy);
                cy = lop3<MAJ>(fours.v[w], m0.w, cy);
                cy = lop3<MAJ>(eights.v[w], m1.x, cy);
                cy = lop3<MAJ>(s16.v[w], m1.y, cy);
                cy = lop3<MAJ>(s32.v[w], m1.z, cy);
                cy = lop3<MAJ>(s64.v[w], m1.w, cy);
                cy = lop3<MAJ>(s128.v[w], m2.x, cy);
                cy = lop3<MAJ>(s256.v[w], m2.y, cy);
                res[w] = cy ^ m2.z;                               // invert in complement mode
            }
            uint32_t any = 0;
#pragma unroll
            for (int w = 0; w < W; ++w) any |= res[w];
            if (__any_sync(0xFFFFFFFFu, any != 0)) {
                hits += __popc(any);
                if (out_mask != nullptr && it == iters - 1) {
#pragma unroll
                    for (int w = 0; w < W; ++w) out_mask[((size_t)blockIdx.x * NQ_CTA + q) * ROWW + lane * W + w] = res[w];
                }
            } else if (out_mask != nullptr && it == iters - 1) {
#pragma unroll
                for (int w = 0; w < W; ++w) out_mask[((size_t)blockIdx.x * NQ_CTA + q) * ROWW + lane * W + w] = 0;
            }
            if (out_planes != nullptr && it == iters - 1 && blockIdx.x == 0) {
                const Vec<W>* pl[10] = {&ptp[0], &ones, &twos, &fours, &eights, &s16, &s32, &s64, &s128, &s256};
#pragma unroll
                for (int p = 0; p < 10; ++p)
#pragma unroll
                    for (int w = 0; w < W; ++w) out_planes[((size_t)q * 10 + p) * ROWW + lane * W + w] = pl[p]->v[w];
            }
        }
    }
    if (out_hits) atomicAdd(out_hits, hits);
}

static uint32_t rng_state = 12345u;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state ^ (rng_state >> 15); }

template <int W>
static int run(int tau, int iters, bool check) {
    constexpr int ROWW = W * 32, ROWS = ROWW * 32;
    // pool chunk: ROWS random 256-bit rows; queries: NQ_CTA random rows (a few near-duplicates of pool rows so that some distances are small)
    std::vector<uint32_t> pool((size_t)ROWS * 8), qs((size_t)NQ_CTA * 8);
    for (auto& v : pool) v = rnd();
    for (auto& v : qs) v = rnd();
    for (int q = 0; q < NQ_CTA; q += 5) {
        memcpy(&qs[(size_t)q * 8], &pool[(size_t)((q * 37) % ROWS) * 8], 32);
        for (int f = 0; f < q % 40; ++f) { int b = rnd() % 256; qs[(size_t)q * 8 + b / 32] ^= 1u << (b % 32); }
    }
    for (int q = 1; q < NQ_CTA; q += 7) for (int w = 0; w < 8; ++w) qs[(size_t)q * 8 + w] |= rnd();   // dense queries: complement mode
    std::vector<uint32_t> P((size_t)257 * ROWW, 0), pt((size_t)9 * ROWW, 0), list((size_t)NQ_CTA * LIST), meta((size_t)NQ_CTA * META, 0);
    std::vector<int> popt(ROWS);
    for (int r = 0; r < ROWS; ++r) {
        int pc = 0;
        for (int b = 0; b < 256; ++b)
            if (pool[(size_t)r * 8 + b / 32] >> (b % 32) & 1) { P[(size_t)b * ROWW + r / 32] |= 1u << (r % 32); ++pc; }
        popt[r] = pc;
        const int ptp = 256 - pc;
        for (int p = 0; p < 9; ++p) if (ptp >> p & 1) pt[(size_t)p * ROWW + r / 32] |= 1u << (r % 32);
    }
    std::vector<int> A(NQ_CTA), inv(NQ_CTA);
    for (int q = 0; q < NQ_CTA; ++q) {
        int a = 0;
        for (int b = 0; b < 256; ++b) a += qs[(size_t)q * 8 + b / 32] >> (b % 32) & 1;
        A[q] = a;
        inv[q] = a > 128;
        int n = 0;
        for (int b = 0; b < 256; ++b) {
            const int bit = qs[(size_t)q * 8 + b / 32] >> (b % 32) & 1;
            if (bit != inv[q]) list[(size_t)q * LIST + n++] = b;
        }
        while (n < LIST) list[(size_t)q * LIST + n++] = 256;
        // normal: pass iff s > K, K = A + 256 - tau;  complement: pass iff !(s > K2 - 1), K2 = tau - A + 256
        int K = inv[q] ? (tau - a + 256) - 1 : a + 256 - tau;
        uint32_t invm = inv[q] ? 0xFFFFFFFFu : 0u;
        if (K < 0) { K = 0; /* s > -1 always: emulate with C = 1023 and s+1023 >= 1024 iff s >= 1 ... handle exactly below */ }
        if (K > 1023) K = 1023;
        const int C = 1023 - K;
        for (int i = 0; i < 10; ++i) meta[(size_t)q * META + i] = (C >> i & 1) ? 0xFFFFFFFFu : 0u;
        meta[(size_t)q * META + 10] = invm;
        meta[(size_t)q * META + 11] = a;
        meta[(size_t)q * META + 12] = tau;
    }
    int nsm = 0;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    uint32_t *dP, *dpt, *dlist, *dmeta, *dmask, *dplanes;
    unsigned long long* dhits;
    CK(cudaMalloc(&dP, P.size() * 4)); CK(cudaMalloc(&dpt, pt.size() * 4)); CK(cudaMalloc(&dlist, list.size() * 4)); CK(cudaMalloc(&dmeta, meta.size() * 4));
    CK(cudaMalloc(&dmask, (size_t)nsm * NQ_CTA * ROWW * 4)); CK(cudaMalloc(&dplanes, (size_t)NQ_CTA * 10 * ROWW * 4)); CK(cudaMalloc(&dhits, 8));
    CK(cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dpt, pt.data(), pt.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dlist, list.data(), list.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dmeta, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dhits, 0, 8));
    const size_t smem = ((size_t)257 * ROWW + 9 * ROWW + NQ_CTA * LIST + NQ_CTA * META) * 4;
    CK(cudaFuncSetAttribute(probe_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    probe_kernel<W><<<nsm, WARPS * 32, smem>>>(dP, dpt, dlist, dmeta, 2, check ? dmask : nullptr, check ? dplanes : nullptr, dhits);
    CK(cudaDeviceSynchronize());
    int bad = 0;
    if (check) {
        std::vector<uint32_t> mask((size_t)NQ_CTA * ROWW), planes((size_t)NQ_CTA * 10 * ROWW);
        CK(cudaMemcpy(mask.data(), dmask, mask.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(planes.data(), dplanes, planes.size() * 4, cudaMemcpyDeviceToHost));
        long long npass = 0;
        for (int q = 0; q < NQ_CTA; ++q)
            for (int r = 0; r < ROWS; ++r) {
                int d = 0;
                for (int w = 0; w < 8; ++w) d += __builtin_popcount(qs[(size_t)q * 8 + w] ^ pool[(size_t)r * 8 + w]);
                int s = 0;
                for (int p = 0; p < 10; ++p) s |= (planes[((size_t)q * 10 + p) * ROWW + r / 32] >> (r % 32) & 1) << p;
                const int dd = inv[q] ? A[q] - 256 + s : A[q] + 256 - s;
                const int m = mask[(size_t)q * ROWW + r / 32] >> (r % 32) & 1;
                npass += m;
                if (dd != d || m != (d < tau)) {
                    if (bad < 10) printf("  MISMATCH W=%d q=%d r=%d: popcount d=%d, planes d=%d, mask=%d (tau %d, A %d, inv %d)\n", W, q, r, d, dd, m, tau, A[q], inv[q]);
                    ++bad;
                }
            }
        printf("W=%d check (tau=%d): %d mismatches of %d pairs, %lld pass\n", W, tau, bad, NQ_CTA * ROWS, npass);
    }
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        probe_kernel<W><<<nsm, WARPS * 32, smem>>>(dP, dpt, dlist, dmeta, iters, nullptr, nullptr, dhits);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const double pairs = (double)nsm * iters * NQ_CTA * ROWS;
    printf("W=%d tau=%d: %.3f ms for %.3e pairs -> %.1f Gpair/s (smem %zu B, %d SMs)\n", W, tau, best, pairs, pairs / best / 1e6, smem, nsm);
    cudaFree(dP); cudaFree(dpt); cudaFree(dlist); cudaFree(dmeta); cudaFree(dmask); cudaFree(dplanes); cudaFree(dhits);
    return bad;
}

int main() {
    int bad = 0;
    bad += run<1>(110, 200, true);
    bad += run<2>(110, 100, true);
    bad += run<4>(110, 50, true);
    run<1>(90, 200, false);
    run<2>(90, 100, false);
    run<4>(90, 50, false);
    printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
    return bad != 0;
}
