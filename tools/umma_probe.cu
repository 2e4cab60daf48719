// Debug probe (not part of the product): minimal tcgen05.mma kind::i8 GEMM tile, checked on the host.
// D[128 x N] (s32, TMEM) = A[128 x K] (u8, K-major smem) * B[N x K]^T (u8; K-major or MN-major smem)
// nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

constexpr int M = 128;

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                     // version = 1 (Blackwell)
    return d;                                   // layout_type = 0 (no swizzle), base_offset = 0
}

// b_mn_major: B stored [K][N] style canonical MN-major; else K-major like A
__global__ void __launch_bounds__(128) k(const uint8_t* A, const uint8_t* B, int32_t* D, int N, int K, int b_mn_major) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* sA = smem;                          // per k-block (32 wide): [2 kcores][16 rowblocks][8 rows][16 B]
    uint8_t* sB = smem + (size_t)M * K;
    const int tid = threadIdx.x, warp = tid >> 5;
    // ---- stage operands into the canonical no-swizzle core-matrix layouts ----------------------
    // A, K-major: element (r, k): kc = k/16, rb = r/8 -> offset kc*(M/8)*128 + rb*128 + (r%8)*16 + k%16
    for (int idx = tid; idx < M * K; idx += 128) {
        const int r = idx / K, kk = idx % K;
        sA[(size_t)(kk / 16) * (M / 8) * 128 + (r / 8) * 128 + (r % 8) * 16 + (kk % 16)] = A[idx];
    }
    if (!b_mn_major) {
        for (int idx = tid; idx < N * K; idx += 128) {
            const int n = idx / K, kk = idx % K;
            sB[(size_t)(kk / 16) * (N / 8) * 128 + (n / 8) * 128 + (n % 8) * 16 + (kk % 16)] = B[idx];
        }
    } else {
        // B, MN-major: core matrix = 16 n (contiguous bytes) x 8 k (16-byte rows).
        // element (n, k): nb = n/16, kb = k/8 -> offset kb*(N/16)*128 + nb*128 + (k%8)*16 + n%16
        for (int idx = tid; idx < N * K; idx += 128) {
            const int n = idx / K, kk = idx % K;
            sB[(size_t)(kk / 8) * (N / 16) * 128 + (n / 16) * 128 + (kk % 8) * 16 + (n % 16)] = B[idx];
        }
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (unsigned)__cvta_generic_to_shared(&tmem_base_s)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    // ---- MMA: K/32 instructions accumulate into TMEM --------------------------------------------
    if (tid == 0) {
        // instruction descriptor: c=S32 (2), a=b=UINT8 (0), majors, N>>3 at bit 17, M>>4 at bit 24
        uint32_t idesc = (2u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | ((uint32_t)(b_mn_major ? 1 : 0) << 16) |
                         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t a0 = (unsigned)__cvta_generic_to_shared(sA), b0 = (unsigned)__cvta_generic_to_shared(sB);
        for (int ks = 0; ks < K / 32; ++ks) {
            // A: one MMA covers 2 K-cores: LBO = distance between K-cores, SBO = distance between row blocks
            const uint64_t da = make_desc(a0 + ks * 2 * (M / 8) * 128, (M / 8) * 128, 128);
            uint64_t db;
            if (!b_mn_major) db = make_desc(b0 + ks * 2 * (N / 8) * 128, (N / 8) * 128, 128);
            else             db = make_desc(b0 + ks * 4 * (N / 16) * 128, (N / 16) * 128, 128);   // 4 k-groups of 8
            const uint32_t acc = ks > 0 ? 1u : 0u;
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_base),
                "l"(da), "l"(db), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         (unsigned)__cvta_generic_to_shared(&bar))
                     : "memory");
    }
    // ---- wait, read back: warp w reads TMEM lanes 32w..32w+31 ------------------------------------
    asm volatile(
        "{\n.reg .pred p;\nWL:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra WD;\nbra WL;\nWD:\n}\n" ::"r"(
            (unsigned)__cvta_generic_to_shared(&bar))
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) D[(size_t)tid * N + c0 + j] = (int32_t)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
}

int main() {
    struct Cfg { int N, K, mn; };
    Cfg cfgs[] = {{64, 32, 0}, {64, 64, 0}, {128, 128, 0}, {256, 64, 0}, {64, 32, 1}, {64, 64, 1}, {256, 128, 1}};
    for (auto c : cfgs) {
        std::vector<uint8_t> A((size_t)M * c.K), B((size_t)c.N * c.K);
        srand(c.N * 7 + c.K);
        for (auto& v : A) v = rand() & 1;
        for (auto& v : B) v = rand() & 1;
        uint8_t *dA, *dB; int32_t* dD;
        cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dD, (size_t)M * c.N * 4);
        cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
        cudaMemset(dD, 0xFF, (size_t)M * c.N * 4);
        size_t smem = (size_t)M * c.K + (size_t)c.N * c.K;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<1, 128, smem>>>(dA, dB, dD, c.N, c.K, c.mn);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<int32_t> D((size_t)M * c.N);
        long bad = -1;
        if (e == cudaSuccess) {
            cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
            bad = 0;
            for (int r = 0; r < M; ++r)
                for (int n = 0; n < c.N; ++n) {
                    int32_t want = 0;
                    for (int kk = 0; kk < c.K; ++kk) want += (int32_t)A[(size_t)r * c.K + kk] * B[(size_t)n * c.K + kk];
                    if (D[(size_t)r * c.N + n] != want) ++bad;
                }
        }
        printf("N=%d K=%d b_mn=%d run=%s mismatches=%ld (D[0]=%d D[1]=%d)\n", c.N, c.K, c.mn, cudaGetErrorString(e), bad,
               D[0], D[1]);
        fflush(stdout);
        if (e != cudaSuccess) return 1;
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    return 0;
}
