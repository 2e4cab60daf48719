// K1 for codes of any size (hypergraph-product codes, n ~ 1600): sparse-row syndrome kernel.
//
// One CTA stages a tile of every error plane in shared memory -- TW consecutive 32-shot words
// (TW*4 bytes) of each of the n planes, fetched once from HBM with 16-byte cp.async -- and then
// forms every syndrome row as the XOR of the planes its CSR row names (css_code.py:728 with a
// sparse H).  Each error bit is read from HBM exactly once although column j of H feeds
// several rows; syndrome words go straight back to HBM with 16-byte stores.  Two CTAs fit per
// SM for n = 1600 (102 KB each), so one CTA's loads overlap the other's XORs.
#include <cuda_runtime.h>

#include "launch.h"

namespace qcss {

namespace {

constexpr int kTiledThreads = 512;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
}

// TW = words per plane per tile (multiple of 4).  smem layout: tile[n][TW] uint32.
template <int TW>
__global__ void __launch_bounds__(kTiledThreads)
k_syndrome_tiled(SparseRows h, const uint32_t* __restrict__ e, int64_t e_stride,
                 uint32_t* __restrict__ s, int64_t s_stride, int64_t words, uint32_t tail_mask) {
    extern __shared__ __align__(16) uint32_t tile[];
    constexpr int kQ = TW / 4;                       // 16-byte chunks per plane row
    const int64_t tiles = (words + TW - 1) / TW;
    const int64_t e_chunks = e_stride / 4;           // valid 16-byte chunks per plane

    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int64_t w0 = t * TW;
        // ---- stage: n * kQ chunks --------------------------------------------------------
        for (int idx = threadIdx.x; idx < h.n * kQ; idx += kTiledThreads) {
            const int j = idx / kQ, q = idx % kQ;
            const int64_t chunk = w0 / 4 + q;
            uint32_t* dst = tile + (size_t)j * TW + q * 4;
            if (chunk < e_chunks) cp_async16(dst, e + (int64_t)j * e_stride + chunk * 4);
            else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        // ---- rows: thread (slot, q) XORs 16 bytes of each named plane ---------------------
        const int q = threadIdx.x % kQ, slot = threadIdx.x / kQ;
        constexpr int kSlots = kTiledThreads / kQ;
        const int64_t wq = w0 + q * 4;
        for (int i = slot; i < h.m; i += kSlots) {
            const int beg = __ldg(h.row_ptr + i), end = __ldg(h.row_ptr + i + 1);
            uint4 acc = make_uint4(0u, 0u, 0u, 0u);
            for (int k = beg; k < end; ++k) {
                const int j = __ldg(h.cols + k);
                const uint4 v = *reinterpret_cast<const uint4*>(tile + (size_t)j * TW + q * 4);
                acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
            }
            if (wq < words) {
                // mask the tail word and anything past it
                uint32_t out[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const int64_t w = wq + v;
                    if (w >= words) out[v] = 0u;
                    else if (w == words - 1) out[v] &= tail_mask;
                }
                if (wq + 4 <= s_stride)
                    *reinterpret_cast<uint4*>(s + (int64_t)i * s_stride + wq) =
                        make_uint4(out[0], out[1], out[2], out[3]);
                else
                    for (int v = 0; v < 4 && wq + v < s_stride; ++v) s[(int64_t)i * s_stride + wq + v] = out[v];
            }
        }
        __syncthreads();
    }
}

template <int TW>
cudaError_t launch_tw(const SparseRows& h, const uint32_t* e, int64_t e_stride, uint32_t* s,
                      int64_t s_stride, int64_t words, uint32_t tail_mask, cudaStream_t stream) {
    const size_t smem = (size_t)h.n * TW * sizeof(uint32_t);
    cudaError_t err = cudaFuncSetAttribute(k_syndrome_tiled<TW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem);
    if (err != cudaSuccess) return err;
    int dev = 0, sms = 0, per_sm = 0;
    if ((err = cudaGetDevice(&dev)) != cudaSuccess) return err;
    if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return err;
    if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_syndrome_tiled<TW>, kTiledThreads,
                                                             smem)) != cudaSuccess) return err;
    if (per_sm < 1) per_sm = 1;
    const int64_t tiles = (words + TW - 1) / TW;
    int64_t grid = (int64_t)sms * per_sm;
    if (grid > tiles) grid = tiles;
    if (grid < 1) grid = 1;
    k_syndrome_tiled<TW><<<(unsigned)grid, kTiledThreads, smem, stream>>>(h, e, e_stride, s, s_stride, words,
                                                                        tail_mask);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_syndrome_tiled(const SparseRows& h, const uint32_t* e, int64_t e_stride, uint32_t* s,
                                  int64_t s_stride, int64_t words, uint32_t tail_mask,
                                  cudaStream_t stream) {
    const size_t budget = 100 * 1024;                 // two CTAs per SM
    if ((size_t)h.n * 16 * 4 <= budget) return launch_tw<16>(h, e, e_stride, s, s_stride, words, tail_mask, stream);
    if ((size_t)h.n * 8 * 4 <= budget) return launch_tw<8>(h, e, e_stride, s, s_stride, words, tail_mask, stream);
    if ((size_t)h.n * 4 * 4 <= 200 * 1024) return launch_tw<4>(h, e, e_stride, s, s_stride, words, tail_mask, stream);
    return cudaErrorInvalidValue;
}

}  // namespace qcss
