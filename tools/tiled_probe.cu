// Debug harness (not part of the product): runs the tiled syndrome launcher standalone.
// nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I quantum_css_codes_b200/csrc -o tools/tiled_probe tools/tiled_probe.cu
#include "../quantum_css_codes_b200/csrc/tiled_kernels.cu"
#include <cstdio>
#include <vector>
#include <random>
using namespace qcss;
int main(int argc, char** argv) {
    int n = 1600, m = 768, rw = 7;
    int64_t shots = argc > 1 ? atoll(argv[1]) : 64;
    std::mt19937 rng(1);
    std::vector<int32_t> ptr(m + 1);
    std::vector<uint16_t> cols;
    for (int i = 0; i < m; ++i) { ptr[i] = (int)cols.size(); for (int k = 0; k < rw; ++k) cols.push_back(rng() % n); }
    ptr[m] = (int)cols.size();
    int64_t stride32 = ((shots + 511) / 512) * 16;     // like planes.stride_words
    int64_t words = (shots + 31) / 32;
    std::vector<uint32_t> e((size_t)n * stride32);
    for (auto& v : e) v = rng();
    uint32_t *d_e, *d_s; int32_t* d_ptr; uint16_t* d_cols;
    cudaMalloc(&d_e, e.size() * 4); cudaMalloc(&d_s, (size_t)m * stride32 * 4);
    cudaMalloc(&d_ptr, ptr.size() * 4); cudaMalloc(&d_cols, cols.size() * 2);
    cudaMemcpy(d_e, e.data(), e.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_ptr, ptr.data(), ptr.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_cols, cols.data(), cols.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(d_s, 0, (size_t)m * stride32 * 4);
    SparseRows h{m, n, rw, d_ptr, d_cols};
    uint32_t tail = (shots & 31) ? ((1u << (shots & 31)) - 1u) : 0xFFFFFFFFu;
    cudaError_t err = launch_syndrome_tiled(h, d_e, stride32, d_s, stride32, words, tail, 0);
    printf("launch: %s\n", cudaGetErrorString(err));
    err = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(err));
    if (err == cudaSuccess) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        float best = 1e9f;
        for (int it = 0; it < 5; ++it) {
            cudaEventRecord(a);
            launch_syndrome_tiled(h, d_e, stride32, d_s, stride32, words, tail, 0);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        printf("best %.3f ms  -> %.1f GB/s algorithmic (%.1f%% of 6549)\n", best, (n + m) / 8.0 * shots / best / 1e6,
               (n + m) / 8.0 * shots / best / 1e6 / 65.491);
    }
    if (err != cudaSuccess) return 1;
    std::vector<uint32_t> s((size_t)m * stride32);
    cudaMemcpy(s.data(), d_s, s.size() * 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    const bool blocked = getenv("QCSS_TILED_BLOCKED") != nullptr;
    for (int i = 0; i < m; ++i)
        for (int64_t w = 0; w < words; ++w) {
            uint32_t want = 0;
            for (int k = ptr[i]; k < ptr[i + 1]; ++k)
                want ^= blocked ? e[((size_t)(w / 32) * n + cols[k]) * 32 + (w % 32)]
                                                      : e[(size_t)cols[k] * stride32 + w];
            if (w == words - 1) want &= tail;
            if (s[(size_t)i * stride32 + w] != want) ++bad;
        }
    printf("shots=%lld mismatches=%ld\n", (long long)shots, bad);
    return 0;
}
