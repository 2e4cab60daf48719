// Debug probe (not part of the product): which 2-D TMA box shapes load cleanly on this driver.
// nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void k(const __grid_constant__ CUtensorMap map, uint32_t* out, int box_w, int box_h, int boxes, int c0) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned bytes = (unsigned)(box_w * box_h * boxes * 4);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        for (int b = 0; b < boxes; ++b) {
            unsigned dst = (unsigned)__cvta_generic_to_shared(smem + (size_t)b * box_w * box_h * 4);
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                ::"r"(dst), "l"(&map), "r"(c0), "r"(b * box_h), "r"(bar_a) : "memory");
        }
    }
    asm volatile(
        "{\n.reg .pred p;\nWL:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra WD;\nbra WL;\nWD:\n}\n"
        ::"r"(bar_a) : "memory");
    const uint32_t* t = reinterpret_cast<const uint32_t*>(smem);
    for (int i = threadIdx.x; i < box_w * box_h * boxes; i += blockDim.x) out[i] = t[i];
}

int main() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    struct Cfg { int cols, rows, box_w, box_h, boxes; };
    Cfg cfgs[] = {{16, 1600, 16, 200, 8}, {64, 1600, 16, 200, 8}, {64, 1600, 32, 200, 4}, {64, 1600, 16, 128, 2},
                  {64, 1600, 16, 256, 2}, {64, 1600, 16, 100, 2}, {64, 1600, 16, 64, 3}, {64, 1600, 8, 200, 2},
                  {64, 256, 32, 64, 1}, {1024, 1600, 16, 200, 8}};
    for (auto c : cfgs) {
        size_t n = (size_t)c.cols * c.rows;
        std::vector<uint32_t> h(n);
        for (size_t i = 0; i < n; ++i) h[i] = (uint32_t)(i * 2654435761u);
        uint32_t *d, *o;
        cudaMalloc(&d, n * 4);
        size_t on = (size_t)c.box_w * c.box_h * c.boxes;
        cudaMalloc(&o, on * 4);
        cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
        CUtensorMap map;
        cuuint64_t dims[2] = {(cuuint64_t)c.cols, (cuuint64_t)c.rows};
        cuuint64_t strides[1] = {(cuuint64_t)c.cols * 4};
        cuuint32_t box[2] = {(cuuint32_t)c.box_w, (cuuint32_t)c.box_h};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        size_t smem = on * 4;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<1, 256, smem>>>(map, o, c.box_w, c.box_h, c.boxes, 0);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<uint32_t> g(on);
        int bad = -1;
        if (e == cudaSuccess) {
            cudaMemcpy(g.data(), o, on * 4, cudaMemcpyDeviceToHost);
            bad = 0;
            for (int b = 0; b < c.boxes; ++b)
                for (int y = 0; y < c.box_h; ++y)
                    for (int x = 0; x < c.box_w; ++x) {
                        size_t row = (size_t)b * c.box_h + y;
                        uint32_t want = row < (size_t)c.rows ? h[row * c.cols + x] : 0u;
                        if (g[((size_t)b * c.box_h + y) * c.box_w + x] != want) ++bad;
                    }
        }
        printf("cols=%d rows=%d box=%dx%d boxes=%d encode=%d run=%s mismatches=%d\n", c.cols, c.rows, c.box_w, c.box_h,
               c.boxes, (int)r, cudaGetErrorString(e), bad);
        fflush(stdout);
        if (e != cudaSuccess) { printf("sticky error; stopping\n"); return 1; }
        cudaFree(d); cudaFree(o);
    }
    return 0;
}
