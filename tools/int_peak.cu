// INT-pipe microbenchmark: measured denominators for the INT roofline.
// Dependent-free chains of LOP3 / IADD3 / IMAD / POPC per thread, enough warps to saturate issue.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/int_peak tools/int_peak.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 8u + i;
    uint32_t b = seed ^ 0x9E3779B9u, c = seed * 3u + 1u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
                if (OP == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
                if (OP == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                if (OP == 3) asm volatile("popc.b32 %0, %0;" : "+r"(a[i]));
                if (OP == 4) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b));
                if (OP == 5) asm volatile("prmt.b32 %0, %0, %1, 0x6240;" : "+r"(a[i]) : "r"(b));
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    if (s == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, int sms) {
    uint32_t* d;
    cudaMalloc(&d, 1 << 24);
    const int iters = 4096, blocks = sms * 8;
    k<OP><<<blocks, 256>>>(d, 16, 1);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<OP><<<blocks, 256>>>(d, iters, r + 2);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double ops = (double)blocks * 256 * iters * 32.0;
    printf("{\"op\": \"%s\", \"lane_ops_per_s\": %.4e, \"ms\": %.3f, \"lane_ops_per_clk_per_sm_at_1965MHz\": %.1f}\n", name,
           ops / (best * 1e-3), best, ops / (best * 1e-3) / sms / 1.965e9);
    cudaFree(d);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<0>("lop3", sms); run<1>("iadd", sms); run<2>("imad", sms); run<3>("popc", sms); run<4>("shf", sms); run<5>("prmt", sms);
    return 0;
}
