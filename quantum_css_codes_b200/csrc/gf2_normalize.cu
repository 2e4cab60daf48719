// CSS construction numerics on the device: the standard form of css_code.py:809-836
// with its COLUMN-SWAP pivot rule, and the CSS condition H1.H2^T = 0 of css_code.py:47-49.
//
// normalize_parity_check is not an RREF: the pivot of row i is forced to column i + offset.  When no row
// j >= i has a 1 there, the reference swaps two QUBITS (columns) -- column i + offset with the first column
// at or right of it where row i has a 1 -- and the swap is replayed on the partner matrix, because it
// relabels qubits of the whole code.  Which column is chosen depends on the partially reduced row, so the
// elimination is sequential in i; the parallelism is inside a step.  One CTA owns one matrix (rows of
// 64-bit words in shared memory when they fit in 200 KB, in place in global memory otherwise):
//
//   A  every thread tests rows i.. for a 1 in the pivot column, shared atomicMin picks the FIRST such row
//      (css_code.py:817);
//   B  found: row i ^= that row unless row i already has the 1 (:819-821);
//      not found: first set bit of row i from the pivot column on (:824), none => "rows are not independent"
//      (:825-826); otherwise swap the two columns in every row of this matrix and of the partner (:828-829)
//      and log the pair;
//   C  one warp per row: rows j != i with a 1 in the pivot column ^= row i (:832-834).
//
// Results are the reference's np.mod(h, 2) and its qubit_swaps list, bit for bit and in order.
#include <cuda_runtime.h>

#include "launch.h"

namespace qcss {

namespace {

constexpr int kNormThreads = 1024;
constexpr size_t kNormSmemMax = 200 * 1024;
constexpr int kNone = 0x7FFFFFFF;

__device__ __forceinline__ void swap_bits(uint64_t* row, int c0, int c1) {
    const int w0 = c0 >> 6, b0 = c0 & 63, w1 = c1 >> 6, b1 = c1 & 63;
    const uint64_t v0 = (row[w0] >> b0) & 1ull, v1 = (row[w1] >> b1) & 1ull;
    if (v0 != v1) {
        row[w0] ^= 1ull << b0;
        row[w1] ^= 1ull << b1;          // w0 == w1 is fine: two different bits of the same word
    }
}

// mats: [batch][m][W] in place.  partner (may be NULL): [batch][mp][W], receives the column swaps only.
// gate (may be NULL): per-matrix status written by an earlier stage; a non-zero entry skips the matrix.
__global__ void __launch_bounds__(kNormThreads)
k_normalize(uint64_t* __restrict__ mats, int m, int n, int W, int offset, uint64_t* __restrict__ partner, int mp,
            int32_t* __restrict__ swaps, int32_t* __restrict__ n_swaps, int32_t* __restrict__ status, int fail_code,
            int use_smem) {
    extern __shared__ uint64_t s_mat[];
    __shared__ int s_row, s_col, s_nswaps;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (status[b] != 0) return;                              // an earlier stage already failed this matrix
    uint64_t* G = mats + (size_t)b * m * W;
    uint64_t* P = partner != nullptr ? partner + (size_t)b * mp * W : nullptr;
    uint64_t* R = use_smem ? s_mat : G;
    const uint64_t tail = (n & 63) ? ((1ull << (n & 63)) - 1ull) : ~0ull;
    for (int idx = tid; idx < m * W; idx += kNormThreads) {
        uint64_t v = G[idx];
        if (idx % W == W - 1) v &= tail;                     // padding bits never take part
        R[idx] = v;
    }
    if (tid == 0) s_nswaps = n_swaps != nullptr ? n_swaps[b] : 0;
    __syncthreads();

    for (int i = 0; i < m; ++i) {
        const int col = i + offset, cw = col >> 6, cb = col & 63;
        if (tid == 0) { s_row = kNone; s_col = kNone; }
        __syncthreads();
        for (int j = i + tid; j < m; j += kNormThreads)
            if ((R[(size_t)j * W + cw] >> cb) & 1ull) atomicMin(&s_row, j);
        __syncthreads();
        const int row = s_row;
        if (row == kNone) {
            for (int w = cw + tid; w < W; w += kNormThreads) {
                uint64_t v = R[(size_t)i * W + w];
                if (w == cw) v &= ~0ull << cb;
                if (v != 0ull) atomicMin(&s_col, w * 64 + __ffsll((long long)v) - 1);
            }
            __syncthreads();
            const int c2 = s_col;
            if (c2 == kNone) {                               // uniform: every thread reads the same s_col
                if (tid == 0) status[b] = fail_code;
                return;
            }
            for (int j = tid; j < m; j += kNormThreads) swap_bits(R + (size_t)j * W, col, c2);
            if (P != nullptr)
                for (int j = tid; j < mp; j += kNormThreads) swap_bits(P + (size_t)j * W, col, c2);
            if (tid == 0) {
                if (swaps != nullptr) {
                    swaps[((size_t)b * n + s_nswaps) * 2 + 0] = col;
                    swaps[((size_t)b * n + s_nswaps) * 2 + 1] = c2;
                }
                s_nswaps++;
            }
            __syncthreads();
        } else if (row != i) {
            for (int w = tid; w < W; w += kNormThreads) R[(size_t)i * W + w] ^= R[(size_t)row * W + w];
            __syncthreads();
        }
        const uint64_t* pivot = R + (size_t)i * W;
        for (int j = warp; j < m; j += kNormThreads / 32) {
            if (j == i) continue;
            uint64_t* target = R + (size_t)j * W;
            const int hit = __shfl_sync(0xFFFFFFFFu, lane == 0 ? (int)((target[cw] >> cb) & 1ull) : 0, 0);
            if (hit)
                for (int w = lane; w < W; w += 32) target[w] ^= pivot[w];
        }
        __syncthreads();
    }
    if (use_smem)
        for (int idx = tid; idx < m * W; idx += kNormThreads) G[idx] = R[idx];
    if (tid == 0 && n_swaps != nullptr) n_swaps[b] = s_nswaps;
}

// status[0] = fail_code when some row of H1 and some row of H2 overlap on an odd number of columns.
__global__ void k_css_condition(const uint64_t* __restrict__ h1, int r1, const uint64_t* __restrict__ h2, int r2,
                                int n, int W, int32_t* __restrict__ status, int fail_code) {
    const uint64_t tail = (n & 63) ? ((1ull << (n & 63)) - 1ull) : ~0ull;
    const size_t pairs = (size_t)r1 * r2;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < pairs; idx += (size_t)gridDim.x * blockDim.x) {
        const uint64_t* a = h1 + (idx / r2) * W;
        const uint64_t* c = h2 + (idx % r2) * W;
        unsigned parity = 0;
        for (int w = 0; w < W; ++w) {
            uint64_t v = a[w] & c[w];
            if (w == W - 1) v &= tail;
            parity ^= (unsigned)__popcll(v);
        }
        if (parity & 1u) atomicCAS(status, 0, fail_code);
    }
}

}  // namespace

// d_mats is normalised in place; d_swaps [batch][n][2] / d_nswaps [batch] (either may be NULL) receive the qubit
// swaps (d_nswaps is read first, so two stages can append to one log); d_status [batch] must be zeroed by the
// first stage's caller and is set to fail_code where the rows turn out to be dependent.
cudaError_t launch_gf2_normalize(uint64_t* d_mats, int batch, int m, int n, int offset, uint64_t* d_partner, int mp,
                                 int32_t* d_swaps, int32_t* d_nswaps, int32_t* d_status, int fail_code,
                                 cudaStream_t stream) {
    if (batch <= 0 || m <= 0 || n <= 0) return cudaSuccess;
    const int W = (n + 63) / 64;
    const size_t bytes = (size_t)m * W * 8;
    const int use_smem = bytes <= kNormSmemMax;
    cudaError_t err = cudaSuccess;
    if (use_smem && bytes > 48 * 1024) {
        err = cudaFuncSetAttribute(k_normalize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNormSmemMax);
        if (err != cudaSuccess) return err;
    }
    k_normalize<<<batch, kNormThreads, use_smem ? bytes : 0, stream>>>(d_mats, m, n, W, offset, d_partner, mp, d_swaps,
                                                                        d_nswaps, d_status, fail_code, use_smem);
    return cudaGetLastError();
}

cudaError_t launch_css_condition(const uint64_t* d_h1, int r1, const uint64_t* d_h2, int r2, int n, int32_t* d_status,
                                 int fail_code, cudaStream_t stream) {
    if (r1 <= 0 || r2 <= 0 || n <= 0) return cudaSuccess;
    const size_t pairs = (size_t)r1 * r2;
    const unsigned blocks = (unsigned)((pairs + 255) / 256 < 148 * 8 ? (pairs + 255) / 256 : 148 * 8);
    k_css_condition<<<blocks, 256, 0, stream>>>(d_h1, r1, d_h2, r2, n, (n + 63) / 64, d_status, fail_code);
    return cudaGetLastError();
}

}  // namespace qcss
