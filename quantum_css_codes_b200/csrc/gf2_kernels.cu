// K4: batched GF(2) Gauss-Jordan elimination (reduced row echelon form, rank, pivot columns)
// behind bin_matrix.reduced_row_echelon_form (reference bin_matrix.py:8-34).
//
// One CTA owns one matrix.  Rows are packed 64 columns per word.  The CTA copies the matrix to
// its output slot and reduces it in place; when the packed matrix fits in shared memory the
// elimination runs there and only the result is written back.  Column by column:
//   1. every warp ballots "row has a 1 in column c" over its rows at or below the pivot counter,
//      the lowest such row index wins (block min);
//   2. that row is exchanged with row `lead` (the reference adds rows instead of swapping,
//      bin_matrix.py:23-24 -- same row space, and the RREF is canonical so the bits agree);
//   3. every other row with a 1 in column c gets the pivot row XORed in, word-parallel, starting
//      at word c/64 (the pivot row is zero left of c).
// This is the general-shape kernel (any number of rows); matrices with <= 1024 rows take the
// register-resident blocked kernels in gf2_m4r2.cu / gf2_m4r.cu.
#include <cuda_runtime.h>

#include "launch.h"
#include "options.h"
#ifdef QCSS_EXPERIMENTS
#include <cstdlib>
namespace qcss {
bool gf2_fast_supported(int m, int n);
cudaError_t launch_gf2_fast(const uint64_t* in, int batch, int m, int n, uint64_t* out, int32_t* rank, int32_t* pivots,
                            cudaStream_t stream);
}
#endif

namespace qcss {

namespace {

constexpr int kGf2Threads = 256;

template <bool IN_SMEM>
__global__ void __launch_bounds__(kGf2Threads)
k_gf2_rref(const uint64_t* __restrict__ in, int batch, int m, int n, uint64_t* __restrict__ out,
           int32_t* __restrict__ rank_out, int32_t* __restrict__ piv_out) {
    extern __shared__ __align__(16) uint64_t smat[];
    __shared__ int s_pivot_row;
    const int W = (n + 63) >> 6;
    const int npiv = m < n ? m : n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kWarps = kGf2Threads / 32;

    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        const uint64_t* src = in + (size_t)b * m * W;
        uint64_t* dst = out + (size_t)b * m * W;
        uint64_t* a = IN_SMEM ? smat : dst;
        // mask of valid columns in the last word, so padding never becomes a pivot
        const uint64_t last_mask = (n & 63) ? ((1ull << (n & 63)) - 1ull) : ~0ull;
        for (int idx = threadIdx.x; idx < m * W; idx += kGf2Threads) {
            uint64_t v = src[idx];
            if ((idx % W) == W - 1) v &= last_mask;
            a[idx] = v;
        }
        __syncthreads();

        int lead = 0;
        for (int c = 0; c < n && lead < m; ++c) {
            const int cw = c >> 6;
            const uint64_t cbit = 1ull << (c & 63);
            // 1. pivot search: lowest row >= lead with the bit set
            if (threadIdx.x == 0) s_pivot_row = m;
            __syncthreads();
            int found = m;
            for (int base = lead + warp * 32; base < m && found == m; base += kWarps * 32) {
                const int row = base + lane;
                const bool hit = row < m && (a[(size_t)row * W + cw] & cbit);
                const unsigned vote = __ballot_sync(0xFFFFFFFFu, hit);
                if (vote) found = base + __ffs(vote) - 1;
            }
            if (lane == 0 && found < m) atomicMin(&s_pivot_row, found);
            __syncthreads();
            const int prow = s_pivot_row;
            if (prow == m) { __syncthreads(); continue; }
            // 2. bring the pivot row to `lead`
            if (prow != lead) {
                for (int w = cw + threadIdx.x; w < W; w += kGf2Threads) {
                    const uint64_t t = a[(size_t)prow * W + w];
                    a[(size_t)prow * W + w] = a[(size_t)lead * W + w];
                    a[(size_t)lead * W + w] = t;
                }
                // words left of cw are zero in both rows at or below `lead`
            }
            __syncthreads();
            // 3. clear column c everywhere else: one warp per row, lanes over words
            for (int row = warp; row < m; row += kWarps) {
                if (row == lead) continue;
                // the predicate word is also lane 0's first XOR target: read it once per warp, before any write
                const bool hit = __shfl_sync(0xFFFFFFFFu, (int)((a[(size_t)row * W + cw] & cbit) != 0), 0) != 0;
                if (hit) {
                    for (int w = cw + lane; w < W; w += 32)
                        a[(size_t)row * W + w] ^= a[(size_t)lead * W + w];
                }
                __syncwarp();
            }
            if (threadIdx.x == 0 && piv_out != nullptr) piv_out[(size_t)b * npiv + lead] = c;
            ++lead;
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            if (rank_out != nullptr) rank_out[b] = lead;
            if (piv_out != nullptr)
                for (int k = lead; k < npiv; ++k) piv_out[(size_t)b * npiv + k] = -1;
        }
        if (IN_SMEM) {
            for (int idx = threadIdx.x; idx < m * W; idx += kGf2Threads) dst[idx] = a[idx];
        }
        __syncthreads();
    }
}

}  // namespace

cudaError_t launch_gf2_rref(const uint64_t* in, int batch, int m, int n, uint64_t* out, int32_t* rank,
                            int32_t* pivots, cudaStream_t stream) {
    if (batch <= 0 || m <= 0 || n <= 0) return cudaSuccess;
    // option "gf2_kernel" (options.h) forces one implementation where its shape limits allow; all of them
    // return the same canonical RREF, rank and pivots (tests/test_gpu_gf2.py runs every choice)
#ifdef QCSS_EXPERIMENTS
    if (getenv("QCSS_GF2_V1") != nullptr && gf2_fast_supported(m, n))       // first generation (tools/experiments)
        return launch_gf2_fast(in, batch, m, n, out, rank, pivots, stream);
#endif
    const int pick = options().gf2_kernel;
    if (pick != 1) {
        // By shape (measured, tools/gf2_shapes.py): the fourth generation replays a 1024-column slab faster (2500 against
        // 3250 cycles per block and 1024 columns) but discovers pivots slower (4400 against 3960 cycles per strip), and a
        // half-empty last slab costs it a full replay: it wins when the matrix is wider than one slab and its width
        // fills the 1024-column slabs to more than half (1024 x 2048: 3.59 against 3.76 ms per 1184 matrices, x 4096:
        // 6.14 against 7.13; x 1536: 3.58 against 2.92, x 2560: 4.88 against 4.60).
        const bool wide = n > 1024 && (((n + 511) / 512) & 1) == 0;
        if ((pick == 4 || (pick == 0 && gf2_m4r2_supported(m, n) && wide)) && gf2_m4r4_supported(m, n))
            return launch_gf2_m4r4(in, batch, m, n, out, rank, pivots, stream);
        if ((pick == 3 && m <= 1024) || (pick == 0 && gf2_m4r2_supported(m, n)))
            return launch_gf2_m4r2(in, batch, m, n, out, rank, pivots, stream);
        if (gf2_m4r_supported(m, n)) return launch_gf2_m4r(in, batch, m, n, out, rank, pivots, stream);
    }
    const int W = (n + 63) >> 6;
    const size_t bytes = (size_t)m * W * sizeof(uint64_t);
    int dev = 0, sms = 0;
    cudaError_t err;
    if ((err = cudaGetDevice(&dev)) != cudaSuccess) return err;
    if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return err;
    const bool in_smem = bytes <= 200 * 1024;
    if (in_smem) {
        if ((err = cudaFuncSetAttribute(k_gf2_rref<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)bytes)) != cudaSuccess) return err;
        int per_sm = 1;
        if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gf2_rref<true>, kGf2Threads,
                                                                 bytes)) != cudaSuccess) return err;
        if (per_sm < 1) per_sm = 1;
        int grid = sms * per_sm;
        if (grid > batch) grid = batch;
        k_gf2_rref<true><<<grid, kGf2Threads, bytes, stream>>>(in, batch, m, n, out, rank, pivots);
    } else {
        int grid = sms * 4;
        if (grid > batch) grid = batch;
        k_gf2_rref<false><<<grid, kGf2Threads, 0, stream>>>(in, batch, m, n, out, rank, pivots);
    }
    return cudaGetLastError();
}

}  // namespace qcss
