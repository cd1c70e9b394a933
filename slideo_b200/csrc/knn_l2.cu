// knn_l2.cu -- K10 (placeholder until the tcgen05 kernel lands): every entry point reports NOTIMPL.
#include "knn_l2.cuh"

namespace slideo {

L2Workspace::~L2Workspace() {
    if (d_qb) cudaFree(d_qb);
    if (d_qn) cudaFree(d_qn);
    if (d_part) cudaFree(d_part);
}
int l2_rows_padded(int n) { return (n + 255) / 256 * 256; }
void l2_prepare_launch(const float*, int, int, uint16_t*, float*, cudaStream_t) { throw NotImplError("SIFT128/L2 path not implemented yet"); }
void l2_knn_launch(L2Workspace&, const float*, int, const uint16_t*, const float*, int, int, int32_t*, float*, int, cudaStream_t, int*) {
    throw NotImplError("SIFT128/L2 path not implemented yet");
}
void l2_vote_launch(const int32_t*, const float*, int, int, const int32_t*, const uint16_t*, int32_t*, int, float, cudaStream_t) {
    throw NotImplError("SIFT128/L2 path not implemented yet");
}

}  // namespace slideo
