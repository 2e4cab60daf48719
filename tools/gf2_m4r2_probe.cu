// Standalone timing of the third-generation batched GF(2) RREF kernel (csrc/gf2_m4r2.cu), for same-run A/B builds.
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 --expt-relaxed-constexpr -DQCSS_M4R4_PROF \
//        -I quantum_css_codes_b200/csrc -o /tmp/m4r4_probe tools/gf2_m4r4_probe.cu && /tmp/m4r4_probe [batch] [m] [n]
// Without -DQCSS_M4R4_PROF the kernel is the product kernel (timing only).
#ifdef PROBE_HEAD
#include "experiments/_gf2_m4r2_head.cu"
#else
#include "../quantum_css_codes_b200/csrc/gf2_m4r2.cu"
#endif

#include <cstdio>
#include <cstdlib>

__global__ void k_fill(uint64_t* p, size_t n, uint64_t seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t x = (i + 1) * 0x9E3779B97F4A7C15ull + seed;       // splitmix64
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
        x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
        p[i] = x ^ (x >> 31);
    }
}

int main(int argc, char** argv) {
    const int batch = argc > 1 ? atoi(argv[1]) : 592, m = argc > 2 ? atoi(argv[2]) : 1024, n = argc > 3 ? atoi(argv[3]) : 2048;
    const size_t words = (size_t)batch * m * ((n + 63) / 64);
    uint64_t *in, *out;
    int32_t* rank;
    cudaMalloc(&in, words * 8);
    cudaMalloc(&out, words * 8);
    cudaMalloc(&rank, batch * 4);
    k_fill<<<1024, 256>>>(in, words, 5);
    qcss::launch_gf2_m4r2(in, batch, m, n, out, rank, nullptr, 0);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError())); return 1; }
#ifdef QCSS_M4R4_PROF
    unsigned long long zero[64] = {};
    cudaMemcpyToSymbol(qcss::g_m4r4_prof, zero, sizeof zero);
#endif
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    qcss::launch_gf2_m4r2(in, batch, m, n, out, rank, nullptr, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int r0 = 0;
    cudaMemcpy(&r0, rank, 4, cudaMemcpyDeviceToHost);
    printf("{\"batch\": %d, \"m\": %d, \"n\": %d, \"ms\": %.4f, \"rank0\": %d", batch, m, n, ms, r0);
#ifdef QCSS_M4R4_PROF
    unsigned long long prof[64];
    cudaMemcpyFromSymbol(prof, qcss::g_m4r4_prof, sizeof prof);
    const char* names[16] = {"rp_publish", "rp_bar_pub", "rp_tabulate", "rp_bar_tab", "rp_reads", "word_boundary", "X_offer", "bar_offer", "Y_tabulate",
                             "Y_panel|bar_tab", "Y_reads", "bar_panel", "-", "Z_lookup_publish_track", "slab_load", "flush_write"};
    const int mats = (batch + 147) / 148;   // matrices CTA 0 processed (grid = 148)
    for (int w = 0; w < 2; ++w) {
        printf(", \"%s_kcycles_per_matrix\": {", w ? "warp1" : "warp0_panel");
        for (int i = 0; i < 16; ++i) printf("%s\"%s\": %.1f", i ? ", " : "", names[i], prof[w * 32 + i] / 1000.0 / mats);
        printf("}");
    }
#endif
    printf("}\n");
    return 0;
}
