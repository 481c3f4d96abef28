// Unit harness (GPU box): device fast_score vs the oracle's host implementation on random 7x7 patches.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../pyorbslam_b200/csrc/kernels_image.cuh"
#include "../../oracle/cvprims.hpp"

__global__ void k_unit(const u8* patches, int n, int t, int* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = fast_score(patches + (size_t)i * 49 + 24, 7, t);
}

int main() {
    const int n = 1 << 18;
    std::vector<u8> h((size_t)n * 49);
    srand(3);
    for (auto& v : h) v = (rand() % 3 == 0) ? rand() % 256 : 100 + rand() % 30;
    u8* d; int* dout;
    cudaMalloc(&d, h.size()); cudaMalloc(&dout, n * 4);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    std::vector<int> got(n);
    for (int t : {7, 20, 0}) {
        k_unit<<<(n + 255) / 256, 256>>>(d, n, t, dout);
        cudaMemcpy(got.data(), dout, n * 4, cudaMemcpyDeviceToHost);
        int bad = 0, corners = 0;
        ptrdiff_t off[16];
        for (int k = 0; k < 16; ++k) off[k] = orbo::kFastDy[k] * 7 + orbo::kFastDx[k];
        for (int i = 0; i < n; ++i) {
            const u8* p = h.data() + (size_t)i * 49 + 24;
            int ref = orbo::fast_is_corner(p, off, t) ? orbo::fast_best(p, 7) - 1 : 0;
            corners += ref != 0;
            if (ref != got[i]) { if (bad < 5) printf("t=%d i=%d ref %d got %d\n", t, i, ref, got[i]); ++bad; }
        }
        printf("t=%d bad %d corners %d (%s)\n", t, bad, corners, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
