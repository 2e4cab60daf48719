// Per-syndrome histograms (the tallies that are all-reduced across GPUs include
// hist_x[2^m2], hist_z[2^m1]).  Input: syndrome planes as qcss_syndrome writes them (plane i = reference
// row i); output: hist[key] += number of shots whose big-endian key (bin_matrix.vec_to_int of the
// syndrome, bin_matrix.py:36-43) is `key`.
//
// One thread takes one 32-shot word of every plane, turns the m plane words into 32 keys with the
// register transpose of core.cuh (8 / 16 / 32 planes at a time) and counts them in a CTA-private
// shared-memory histogram (m <= 13: 32 KB of 32-bit bins, flushed with 64-bit global atomics), or with
// global atomics directly for 13 < m <= 24.
#include <cuda_runtime.h>

#include "core.cuh"
#include "launch.h"

namespace qcss {

namespace {

constexpr int kHistThreads = 256;
constexpr int kHistSmemM = 13;

template <int MB, bool SMEM>
__global__ void __launch_bounds__(kHistThreads)
k_syndrome_hist(const uint32_t* __restrict__ s, int64_t s_stride, int m, int64_t words, uint32_t tail_mask,
                unsigned long long* __restrict__ hist) {
    extern __shared__ uint32_t bins[];
    const uint32_t size = 1u << m;
    if (SMEM) {
        for (uint32_t i = threadIdx.x; i < size; i += kHistThreads) bins[i] = 0u;
        __syncthreads();
    }
    for (int64_t w = blockIdx.x * (int64_t)kHistThreads + threadIdx.x; w < words; w += (int64_t)gridDim.x * kHistThreads) {
        uint32_t v[MB];
#pragma unroll
        for (int t = 0; t < MB; ++t) {
            // v[t] = plane feeding key bit t = reference row m - 1 - t
            v[t] = (t < m) ? __ldg(s + (int64_t)(m - 1 - t) * s_stride + w) : 0u;
        }
        const uint32_t valid = (w == words - 1) ? tail_mask : 0xFFFFFFFFu;
        transpose_blocks<MB>(v);            // bits [b*MB, b*MB+MB) of v[j] = key of shot b*MB + j
        constexpr int kBlocks = 32 / MB;
        constexpr uint32_t kMask = MB == 32 ? 0xFFFFFFFFu : ((1u << MB) - 1u);
#pragma unroll
        for (int j = 0; j < MB; ++j) {
#pragma unroll
            for (int b = 0; b < kBlocks; ++b) {
                const int shot = b * MB + j;
                if ((valid >> shot) & 1u) {
                    const uint32_t key = (v[j] >> (b * MB)) & kMask;
                    if (SMEM) atomicAdd(bins + key, 1u);
                    else atomicAdd(hist + key, 1ull);
                }
            }
        }
    }
    if (SMEM) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < size; i += kHistThreads)
            if (bins[i] != 0u) atomicAdd(hist + i, (unsigned long long)bins[i]);
    }
}

template <int MB>
cudaError_t launch_hist_mb(const uint32_t* s, int64_t s_stride, int m, int64_t words, uint32_t tail_mask,
                           unsigned long long* hist, cudaStream_t stream) {
    int dev = 0, sms = 0;
    cudaError_t err;
    if ((err = cudaGetDevice(&dev)) != cudaSuccess) return err;
    if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return err;
    int64_t grid = (words + kHistThreads - 1) / kHistThreads;
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
    if (grid < 1) grid = 1;
    if (m <= kHistSmemM) {
        const size_t smem = ((size_t)1 << m) * sizeof(uint32_t);
        k_syndrome_hist<MB, true><<<(unsigned)grid, kHistThreads, smem, stream>>>(s, s_stride, m, words, tail_mask, hist);
    } else {
        k_syndrome_hist<MB, false><<<(unsigned)grid, kHistThreads, 0, stream>>>(s, s_stride, m, words, tail_mask, hist);
    }
    return cudaGetLastError();
}

}  // namespace

// s: syndrome planes as 32-bit words ([m][s_stride]); hist: uint64[2^m], accumulated.
cudaError_t launch_syndrome_hist(const uint32_t* s, int64_t s_stride, int m, int64_t words, uint32_t tail_mask,
                                 unsigned long long* hist, cudaStream_t stream) {
    if (m < 1 || m > 24) return cudaErrorInvalidValue;
    if (m <= 8) return launch_hist_mb<8>(s, s_stride, m, words, tail_mask, hist, stream);
    if (m <= 16) return launch_hist_mb<16>(s, s_stride, m, words, tail_mask, hist, stream);
    return launch_hist_mb<32>(s, s_stride, m, words, tail_mask, hist, stream);
}

}  // namespace qcss
