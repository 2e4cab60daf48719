// Tensor-pipe denominator for the dense syndrome kernel (K1'): back-to-back tcgen05.mma.kind::i8,
// M = 128, N = 256, K = 32, cta_group::1, operands fixed in shared memory (no data movement at all),
// two alternating TMEM accumulators, one issuing thread per CTA, one CTA per SM.  Prints one JSON line:
// int-ops/s (2 * M * N * K per MMA) for the whole GPU -- the `peak` of bench.py's c4_dense roofline.
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -o tools/i8_mma_peak tools/i8_mma_peak.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

constexpr int M = 128, N = 256, KC = 64;      // one "chunk" = 64 qubits = two K = 32 MMAs per accumulator

__global__ void __launch_bounds__(128, 1) k_peak(int chunks, uint32_t* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (M * KC + N * KC) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01000101u * (i & 1);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (unsigned)__cvta_generic_to_shared(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t a0 = (unsigned)__cvta_generic_to_shared(smem), b0 = a0 + M * KC;
        for (int c = 0; c < chunks; ++c) {
#pragma unroll
            for (int acc = 0; acc < 2; ++acc) {
#pragma unroll
                for (int ks = 0; ks < KC / 32; ++ks) {
                    const uint64_t da = make_desc(a0 + ks * 2 * (M / 8) * 128, (M / 8) * 128, 128);
                    const uint64_t db = make_desc(b0 + ks * 4 * (N / 16) * 128, (N / 16) * 128, 128);
                    const uint32_t accum = (c > 0 || ks > 0) ? 1u : 0u;
                    asm volatile(
                        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_base + acc * N),
                        "l"(da), "l"(db), "r"(idesc), "r"(accum)
                        : "memory");
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         (unsigned)__cvta_generic_to_shared(&bar))
                     : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nWL:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra WD;\nbra WL;\nWD:\n}\n" ::"r"(
            (unsigned)__cvta_generic_to_shared(&bar))
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tmem_base + ((uint32_t)(warp * 32) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (v == 0xDEADBEEFu) sink[blockIdx.x * 128 + tid] = v;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* sink;
    cudaMalloc(&sink, sms * 128 * 4);
    const size_t smem = (size_t)M * KC + (size_t)N * KC;
    cudaFuncSetAttribute(k_peak, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int chunks = 8192;
    k_peak<<<sms, 128, smem>>>(64, sink);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f, sum = 0.f;
    const int reps = 10;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        k_peak<<<sms, 128, smem>>>(chunks, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
        sum += ms;
    }
    const double ops = 2.0 * M * N * 32 * 4.0 * chunks * sms;        // 4 MMAs (K = 32) per chunk
    printf("{\"op\": \"tcgen05.mma.kind::i8 m128n256k32 cta_group::1\", \"sms\": %d, \"ms_best\": %.4f, \"ms_mean\": %.4f, "
           "\"int_ops_per_s_best\": %.4e, \"int_ops_per_s_mean\": %.4e, \"mma_cycles_at_1965MHz\": %.1f}\n",
           sms, best, sum / reps, ops / (best * 1e-3), ops / (sum / reps * 1e-3), best * 1e-3 * 1.965e9 / (4.0 * chunks));
    return 0;
}
