/*
 * bf_oracle.c -- CPU restatement (plain C) of the matcher + vote stage of the reference hot path.
 * TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline/reference arm).
 *
 * What it restates
 *   reference call sites: crates/matching-opencv/src/flann.rs:64-89 (pool = one Mat per page, added in
 *                         `images` order; knn_match(q, k) -> rows of {query_idx, train_idx, img_idx, distance})
 *                         crates/matching-opencv/src/lib.rs:266 (k = 30) and lib.rs:268-282 (1.05-ratio vote)
 *   arithmetic          : OpenCV DescriptorMatcher::knnMatch (third party, not under /root/reference).
 *                         The reference uses the *approximate, non-deterministic* FLANN-LSH index; the
 *                         contract is the exact brute-force k-NN that LSH
 *                         approximates == cv2.BFMatcher(NORM_HAMMING / NORM_L2).knnMatch:
 *                         the k smallest by (distance, global pooled index) ascending.
 *   parity pin          : tests/test_oracle_bf.py compares against cv2.BFMatcher on seeded pools with
 *                         planted duplicates (ties across pages) -- golden vectors in tests/golden/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int ham256(const uint8_t* a, const uint8_t* b) {
    uint64_t x[4], y[4];
    memcpy(x, a, 32);
    memcpy(y, b, 32);
    return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) + __builtin_popcountll(x[2] ^ y[2]) +
           __builtin_popcountll(x[3] ^ y[3]);
}

/* insertion into a sorted (dist, idx) list of length <= k; strict '<' keeps the earlier index on ties
 * because candidates are visited in increasing global index order. */
static inline void topk_push_i(int32_t* bd, int32_t* bi, int* len, int k, int d, int idx) {
    int n = *len;
    if (n == k && d >= bd[k - 1]) return;
    int p = n < k ? n : k - 1;
    while (p > 0 && bd[p - 1] > d) { bd[p] = bd[p - 1]; bi[p] = bi[p - 1]; --p; }
    bd[p] = d; bi[p] = idx;
    if (n < k) *len = n + 1;
}

/* Hamming brute-force k-NN.  q: nq x 32, t: nt x 32.  idx/dist: nq x k (rows padded with -1 when nt < k). */
void bf_knn_hamming(const uint8_t* q, int nq, const uint8_t* t, int nt, int k, int32_t* idx, int32_t* dist) {
    for (int i = 0; i < nq; ++i) {
        int32_t* bd = dist + (size_t)i * k;
        int32_t* bi = idx + (size_t)i * k;
        int len = 0;
        const uint8_t* qi = q + (size_t)i * 32;
        for (int j = 0; j < nt; ++j) topk_push_i(bd, bi, &len, k, ham256(qi, t + (size_t)j * 32), j);
        for (int p = len; p < k; ++p) { bd[p] = -1; bi[p] = -1; }
    }
}

static inline void topk_push_f(float* bd, int32_t* bi, int* len, int k, float d, int idx) {
    int n = *len;
    if (n == k && !(d < bd[k - 1])) return;
    int p = n < k ? n : k - 1;
    while (p > 0 && bd[p - 1] > d) { bd[p] = bd[p - 1]; bi[p] = bi[p - 1]; --p; }
    bd[p] = d; bi[p] = idx;
    if (n < k) *len = n + 1;
}

/* L2 brute-force k-NN on float descriptors (cv2.BFMatcher(NORM_L2)): dist = sqrtf(sum (a-b)^2) in fp32.
 * For the integer-valued descriptors cv2.SIFT emits, the fp32 sum is exact whatever the order. */
void bf_knn_l2(const float* q, int nq, const float* t, int nt, int dim, int k, int32_t* idx, float* dist) {
    for (int i = 0; i < nq; ++i) {
        float* bd = dist + (size_t)i * k;
        int32_t* bi = idx + (size_t)i * k;
        int len = 0;
        const float* qi = q + (size_t)i * dim;
        for (int j = 0; j < nt; ++j) {
            const float* tj = t + (size_t)j * dim;
            float s = 0.f;
            for (int c = 0; c < dim; ++c) { float d = qi[c] - tj[c]; s += d * d; }
            topk_push_f(bd, bi, &len, k, sqrtf(s), j);
        }
        for (int p = len; p < k; ++p) { bd[p] = -1.f; bi[p] = -1; }
    }
}

/* The vote of lib.rs:268-282 on k-NN rows of one frame.
 *   page_offsets: P+1 prefix offsets of the pooled descriptors (pool built in `images` order, lib.rs:259)
 *   votes       : P counters (zeroed here)
 * best = row[0]; every entry with (float)d < (float)best * 1.05f votes for its page.  best == 0 -> no votes.
 * Returns argmax page (ties -> lowest page index; the reference's HashMap order is arbitrary),
 * or -1 when nobody voted. */
int bf_vote(const int32_t* idx, const float* dist, int nq, int k, const int32_t* page_offsets, int npages,
            int32_t* votes) {
    memset(votes, 0, sizeof(int32_t) * (size_t)npages);
    for (int i = 0; i < nq; ++i) {
        const int32_t* bi = idx + (size_t)i * k;
        const float* bd = dist + (size_t)i * k;
        if (bi[0] < 0) continue;
        float lim = bd[0] * 1.05f;
        for (int p = 0; p < k && bi[p] >= 0; ++p) {
            if (bd[p] < lim) {
                int lo = 0, hi = npages;          /* page = last offset <= idx */
                while (hi - lo > 1) { int mid = (lo + hi) / 2; if (page_offsets[mid] <= bi[p]) lo = mid; else hi = mid; }
                votes[lo]++;
            }
        }
    }
    int best = -1, bv = 0;
    for (int p = 0; p < npages; ++p) if (votes[p] > bv) { bv = votes[p]; best = p; }
    return best;
}
