// probe: 3-D u8 TMA box with negative / out-of-range coordinates (what fast_kernel wants)
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../slideo_b200/csrc/tma.cuh"
using namespace slideo;
__global__ void k(const CUtensorMap* tm, int x, int y, int z, uint32_t* out) {
    __shared__ __align__(128) uint32_t s_px[40][20];
    __shared__ __align__(8) uint64_t s_bar;
    if (threadIdx.x == 0) {
        tma_mbar_init(&s_bar, 1);
        tma_mbar_expect_tx(&s_bar, 3200);
        tma_load_3d(&s_px[0][0], tm, x, y, z, &s_bar);
    }
    __syncthreads();
    tma_mbar_wait(&s_bar, 0);
    for (int i = threadIdx.x; i < 800; i += blockDim.x) out[i] = s_px[i / 20][i % 20];
}
int main() {
    const int pitch = 1920, h = 1080, n = 2;
    std::vector<uint8_t> img((size_t)pitch * h * n);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)(i * 7 + (i >> 11));
    uint8_t* d; cudaMalloc(&d, img.size()); cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
    CUtensorMap tm = tma_map_3d(CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, d, pitch, h, n, pitch, (uint64_t)pitch * h, 80, 40);
    CUtensorMap* dtm; cudaMalloc(&dtm, sizeof tm); cudaMemcpy(dtm, &tm, sizeof tm, cudaMemcpyHostToDevice);
    uint32_t* dout; cudaMalloc(&dout, 3200);
    const int cases[][3] = {{64, 32, 0}, {48, 28, 1}, {-16, -4, 0}, {1904, 1060, 1}};
    for (auto& c : cases) {
        k<<<1, 256>>>(dtm, c[0], c[1], c[2], dout);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<uint8_t> o(3200);
        cudaMemcpy(o.data(), dout, 3200, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r = 0; r < 40; ++r) for (int cc = 0; cc < 80; ++cc) {
            const int gx = c[0] + cc, gy = c[1] + r;
            const uint8_t want = (gx >= 0 && gx < pitch && gy >= 0 && gy < h) ? img[((size_t)c[2] * h + gy) * pitch + gx] : 0;
            bad += o[r * 80 + cc] != want;
        }
        printf("coords (%d, %d, %d): %s, mismatches %d\n", c[0], c[1], c[2], cudaGetErrorString(e), bad);
        if (e != cudaSuccess) break;
    }
    return 0;
}
