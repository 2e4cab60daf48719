// Back-to-back tcgen05.mma.cta_group::2.kind::i8 (M = 256 over a CTA pair, N = 256, K = 32), operands fixed in shared
// memory, B either MN-major no-swizzle (the dense syndrome kernel's layout) or K-major no-swizzle.  Diagnostic for
// tools/experiments/dense_kernels_cta_pair.cu.txt: does the pair reach the single-SM MMA rate with this operand layout?
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/i8_pair tools/experiments/i8_mma_pair_peak.cu
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
constexpr int KC = 64;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_pair(int chunks, int b_mn_major, uint32_t* sink, int commit_every, int wait_mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar, dummy, ready;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = tid; i < (128 * KC + 128 * KC) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01000101u * (i & 1);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"((unsigned)__cvta_generic_to_shared(&dummy)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&ready)));
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(&ready)) : "memory");   // phase 0 complete
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (unsigned)__cvta_generic_to_shared(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    if (rank == 0 && tid == 0) {
        const uint32_t idesc = (2u << 4) | ((uint32_t)(b_mn_major ? 1 : 0) << 16) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        const uint32_t a0 = (unsigned)__cvta_generic_to_shared(smem), b0 = a0 + 128 * KC;
        for (int c = 0; c < chunks; ++c) {
            if (wait_mode) {   // what the dense kernel's MMA lane does per chunk: wait (already complete here), then fence
                const unsigned addr = (unsigned)__cvta_generic_to_shared(&ready);
                if (wait_mode == 1)
                    asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], 0;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(addr) : "memory");
                else
                    asm volatile("{\n.reg .pred p;\nW2:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D2;\nbra W2;\nD2:\n}\n" ::"r"(addr) : "memory");
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
#pragma unroll
            for (int acc = 0; acc < 2; ++acc) {
#pragma unroll
                for (int ks = 0; ks < KC / 32; ++ks) {
                    const uint64_t da = make_desc(a0 + ks * 2 * 16 * 128, 16 * 128, 128);
                    const uint64_t db = b_mn_major ? make_desc(b0 + ks * 4 * (128 / 16) * 128, (128 / 16) * 128, 128)
                                                   : make_desc(b0 + ks * 2 * 16 * 128, 16 * 128, 128);
                    const uint32_t accum = (c > 0 || ks > 0) ? 1u : 0u;
                    asm volatile(
                        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_base + acc * 256),
                        "l"(da), "l"(db), "r"(idesc), "r"(accum)
                        : "memory");
                }
            }
            if (commit_every) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                         (unsigned)__cvta_generic_to_shared(&dummy)), "h"((uint16_t)3) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                         (unsigned)__cvta_generic_to_shared(&bar)), "h"((uint16_t)3)
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
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    sms &= ~1;
    uint32_t* sink;
    cudaMalloc(&sink, sms * 128 * 4);
    const size_t smem = (size_t)128 * KC * 2;
    cudaFuncSetAttribute(k_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int mn = 0; mn < 8; ++mn) {
        const int chunks = 8192, ce = 1, wm = mn >> 1;      // wait_mode 0 none, 1 acquire.cluster, 2 default (cta) scope, 3 = 2
        k_pair<<<sms, 128, smem>>>(64, mn & 1, sink, ce, wm);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e9f;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(e0);
            k_pair<<<sms, 128, smem>>>(chunks, mn & 1, sink, ce, wm);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        const double ops = 2.0 * 256 * 256 * 32 * 4.0 * chunks * (sms / 2);
        printf("{\"op\": \"tcgen05.mma.kind::i8 m256n256k32 cta_group::2\", \"b_layout\": \"%s\", \"ms\": %.4f, \"int_ops_per_s\": %.4e, "
               "\"commit_per_chunk\": %d, \"wait_per_chunk\": \"%s\", \"cycles_per_mma_at_1965MHz\": %.1f}\n", (mn & 1) ? "MN-major no-swizzle" : "K-major no-swizzle", best, ops / (best * 1e-3), ce,
               wm == 0 ? "none" : wm == 1 ? "try_wait.acquire.cluster + tcgen05.fence" : "try_wait (cta scope) + tcgen05.fence",
               best * 1e-3 * 1.965e9 / (4.0 * chunks));
    }
    return 0;
}
