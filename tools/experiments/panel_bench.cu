// Microbenchmark of the byte-space panel (gf2_m4r4.cu) in isolation: one panel warp per CTA, optionally with `load`
// other warps hammering shared memory with 128-bit loads (the table reads of the real kernel).  Reports cycles per
// panel (8 columns) for the REDUX and the SHFL broadcast.
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/panel_bench tools/experiments/panel_bench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

template <int MODE>   // 0 = REDUX, 1 = SHFL
__device__ __forceinline__ uint32_t panel(const uint16_t* rep, int lane, uint32_t* G) {
    const uint4 q = reinterpret_cast<const uint4*>(rep)[lane];
    const uint32_t prs0 = (__byte_perm(q.x, q.y, 0x7531u) >> 7) & 0x01010101u;
    const uint32_t prs1 = (__byte_perm(q.z, q.w, 0x7531u) >> 7) & 0x01010101u;
    uint32_t red0 = 0x03020100u + 0x08080808u * (uint32_t)lane, red1 = red0 + 0x04040404u;
    uint32_t y0 = 0u, y1 = 0u, myval = 0u, kmask = 0x01010101u;
    int k = 0;
#pragma unroll 1
    for (int col = 0; col < 8; ++col) {
        const uint32_t s0 = red0 >> col, s1 = red1 >> col;
        const uint32_t c = (s0 & prs0) + ((s1 & prs1) << 4);
        const unsigned vote = __ballot_sync(0xFFFFFFFFu, c != 0u);
        if (vote != 0u) {
            const int srcl = 31 - __clz((int)vote);
            const uint32_t pbit = 31u - (uint32_t)__clz((int)c);
            const uint32_t sel = 0x73625140u >> (pbit & 28u);
            const uint32_t mine = prmt(prmt(prmt(red0, red1, sel), prmt(y0, y1, sel), 0x0040u), sel, 0x4410u);
            uint32_t pack;
            if (MODE == 0) pack = __reduce_or_sync(0xFFFFFFFFu, lane == srcl ? mine : 0u);
            else pack = __shfl_sync(0xFFFFFFFFu, mine, srcl);
            const uint32_t v4 = prmt(pack, 0u, 0x0000u);
            const uint32_t yk4 = prmt(pack, 0u, 0x1111u) | kmask;
            const uint32_t M0 = (s0 & 0x01010101u) * 0xFFu, M1 = (s1 & 0x01010101u) * 0xFFu;
            red0 ^= M0 & v4;  red1 ^= M1 & v4;
            y0 ^= M0 & yk4;   y1 ^= M1 & yk4;
            if (lane == k) myval = (uint32_t)srcl * 8u + ((pack >> 16) & 7u);
            kmask <<= 1;
            ++k;
        }
    }
    reinterpret_cast<uint2*>(G)[lane] = make_uint2(y0, y1);
    return myval + k;
}

// branch-free columns: a column without a candidate multiplies its masks by zero instead of branching around the update
template <int MODE>   // 2 = SHFL, 3 = REDUX
__device__ __forceinline__ uint32_t panel_bf(const uint16_t* rep, int lane, uint32_t* G) {
    const uint4 q = reinterpret_cast<const uint4*>(rep)[lane];
    const uint32_t prs0 = (__byte_perm(q.x, q.y, 0x7531u) >> 7) & 0x01010101u;
    const uint32_t prs1 = (__byte_perm(q.z, q.w, 0x7531u) >> 7) & 0x01010101u;
    uint32_t red0 = 0x03020100u + 0x08080808u * (uint32_t)lane, red1 = red0 + 0x04040404u;
    uint32_t y0 = 0u, y1 = 0u, myval = 0u, kmask = 0x01010101u;
    uint32_t k = 0;
#pragma unroll
    for (int col = 0; col < 8; ++col) {
        const uint32_t s0 = red0 >> col, s1 = red1 >> col;
        const uint32_t c = (s0 & prs0) + ((s1 & prs1) << 4);
        const unsigned vote = __ballot_sync(0xFFFFFFFFu, c != 0u);
        const uint32_t found = vote != 0u ? 1u : 0u;
        const int srcl = 31 - __clz((int)vote);
        const uint32_t pbit = 31u - (uint32_t)__clz((int)c);
        const uint32_t sel = 0x73625140u >> (pbit & 28u);
        const uint32_t mine = prmt(prmt(prmt(red0, red1, sel), prmt(y0, y1, sel), 0x0040u), sel, 0x4410u);
        uint32_t pack;
        if (MODE == 3) pack = __reduce_or_sync(0xFFFFFFFFu, lane == srcl ? mine : 0u);
        else pack = __shfl_sync(0xFFFFFFFFu, mine, srcl & 31);
        const uint32_t v4 = prmt(pack, 0u, 0x0000u);
        const uint32_t yk4 = prmt(pack, 0u, 0x1111u) | kmask;
        const uint32_t fm = found * 0xFFu;
        const uint32_t M0 = (s0 & 0x01010101u) * fm, M1 = (s1 & 0x01010101u) * fm;
        red0 ^= M0 & v4;  red1 ^= M1 & v4;
        y0 ^= M0 & yk4;   y1 ^= M1 & yk4;
        if (found && lane == (int)k) myval = (uint32_t)srcl * 8u + ((pack >> 16) & 7u);
        kmask <<= found;
        k += found;
    }
    reinterpret_cast<uint2*>(G)[lane] = make_uint2(y0, y1);
    return myval + k;
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_bench(int iters, int load, unsigned long long* out, uint32_t* sink) {
    __shared__ __align__(16) uint16_t rep[256];
    __shared__ __align__(16) uint32_t G[64];
    __shared__ int stop;
    __shared__ __align__(16) uint4 table[2048];           // 32 KB
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 256; i += blockDim.x) rep[i] = (uint16_t)(((i * 2654435761u) >> 7 & 0x3FF) | ((i * 40503u >> 3 & 3) ? 0x8000 : 0));
    for (int i = tid; i < 2048; i += blockDim.x) table[i] = make_uint4(i, i * 3, i * 5, i * 7);
    if (tid == 0) stop = 0;
    __syncthreads();
    if (warp == 0) {
        uint32_t acc = 0;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            acc += MODE < 2 ? panel<MODE>(rep, lane, G) : panel_bf<MODE>(rep, lane, G);
            rep[(acc + it) & 255] ^= 0x8000;              // keep the compiler honest, vary the input
            __syncwarp();
        }
        const long long t1 = clock64();
        if (lane == 0) { out[blockIdx.x] = (unsigned long long)(t1 - t0); sink[blockIdx.x] = acc; }
        *reinterpret_cast<volatile int*>(&stop) = 1;
    } else if (warp <= load) {
        uint4 a = make_uint4(0, 0, 0, 0);
        uint32_t idx = tid;
        while (*reinterpret_cast<volatile int*>(&stop) != 1) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const uint4 v = table[(idx * 8 + (lane & 7)) & 2047];
                a.x ^= v.x; a.y ^= v.y; a.z ^= v.z; a.w ^= v.w;
                idx = idx * 1664525u + 1013904223u + a.x;
            }
        }
        if (a.x == 0xDEADBEEF) sink[1] = a.y;
    }
}

int main() {
    unsigned long long* out;
    uint32_t* sink;
    cudaMalloc(&out, 148 * 8);
    cudaMalloc(&sink, 148 * 4);
    const int iters = 2000;
    for (int mode = 0; mode < 4; ++mode)
        for (int load : {0, 7, 15, 31}) {
            cudaMemset(out, 0, 148 * 8);
            if (mode == 0) k_bench<0><<<148, 32 * (load + 1)>>>(iters, load, out, sink);
            else if (mode == 1) k_bench<1><<<148, 32 * (load + 1)>>>(iters, load, out, sink);
            else if (mode == 2) k_bench<2><<<148, 32 * (load + 1)>>>(iters, load, out, sink);
            else k_bench<3><<<148, 32 * (load + 1)>>>(iters, load, out, sink);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            unsigned long long h[148];
            cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
            printf("{\"broadcast\": \"%s\", \"other_warps_reading_smem\": %d, \"cycles_per_panel\": %.0f}\n", mode == 0 ? "REDUX" : mode == 1 ? "SHFL" : mode == 2 ? "SHFL branch-free" : "REDUX branch-free", load, (double)h[0] / iters);
        }
    return 0;
}
