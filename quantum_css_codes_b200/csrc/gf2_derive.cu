// K4 derived products: batched null space and linear solve on top of the batched RREF.
//
// The reference has neither (bin_matrix.py stops at reduced_row_echelon_form); BASELINE
// config 5 asks for "row reduction / rank / null space" of 4096 matrices, so both are built on the
// device from the RREF R (rank r, pivot columns p_0 < ... < p_{r-1}) without a round trip to the host:
//
//   null space   one basis vector per free column f, in increasing order of f:
//                    x[f] = 1,  x[p_i] = R[i][f],  every other free variable 0
//                i.e. the free columns of R transposed and scattered to the pivot positions.  A CTA
//                takes one matrix and one 64-column word of free columns; a warp takes 32 rows, and
//                one ballot per free column turns "bit f of 32 rows" into 32 bits of the basis vector,
//                which land with one or two 32-bit atomic ORs when the 32 pivots are consecutive
//                columns (the usual case) and bit by bit otherwise.
//   solve        RREF of the augmented matrix [A | b]; inconsistent iff column n holds a pivot;
//                otherwise x[p_i] = R[i][n] with every free variable 0.
#include <cuda_runtime.h>

#include "launch.h"

namespace qcss {

namespace {

constexpr int kNsThreads = 256;

__global__ void __launch_bounds__(kNsThreads)
k_nullspace(const uint64_t* __restrict__ rref, const int32_t* __restrict__ rank, const int32_t* __restrict__ piv,
            int batch, int m, int n, int max_rows, uint32_t* __restrict__ basis, int32_t* __restrict__ overflow) {
    const int W = (n + 63) >> 6;
    const int npiv = m < n ? m : n;
    const int b = blockIdx.x / W, fw = blockIdx.x % W;
    if (b >= batch) return;
    __shared__ unsigned long long s_pivmask;
    __shared__ int s_before;
    if (threadIdx.x == 0) { s_pivmask = 0ull; s_before = 0; }
    __syncthreads();
    const int r = rank[b];
    const int32_t* p = piv + (size_t)b * npiv;
    int before = 0;
    for (int i = threadIdx.x; i < r; i += kNsThreads) {
        const int c = p[i];
        if ((c >> 6) == fw) atomicOr(&s_pivmask, 1ull << (c & 63));
        before += (c < 64 * fw);
    }
    if (before) atomicAdd(&s_before, before);
    __syncthreads();
    const int cols_here = (n - 64 * fw) < 64 ? (n - 64 * fw) : 64;
    const uint64_t valid = cols_here == 64 ? ~0ull : ((1ull << cols_here) - 1ull);
    const uint64_t freemask = ~s_pivmask & valid;
    if (freemask == 0ull) return;
    const int tbase = 64 * fw - s_before;                 // basis row of the first free column of this word
    const int nfree = __popcll(freemask);
    if (tbase + nfree > max_rows) {
        if (threadIdx.x == 0) atomicMax(overflow, tbase + nfree);
    }
    const size_t W32 = (size_t)W * 2;
    uint32_t* out = basis + (size_t)b * max_rows * W32;
    const uint64_t* R = rref + (size_t)b * m * W;
    // x[f] = 1
    for (int k = threadIdx.x; k < 64; k += kNsThreads) {
        if ((freemask >> k) & 1ull) {
            const int t = tbase + __popcll(freemask & ((1ull << k) - 1ull));
            const int f = 64 * fw + k;
            if (t < max_rows) atomicOr(out + (size_t)t * W32 + (f >> 5), 1u << (f & 31));
        }
    }
    // x[p_i] = R[i][f]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i0 = warp * 32; i0 < r; i0 += (kNsThreads / 32) * 32) {
        const int i = i0 + lane;
        const uint64_t v = (i < r) ? (R[(size_t)i * W + fw] & freemask) : 0ull;
        const int pi = (i < r) ? p[i] : -1;
        const int p0 = __shfl_sync(0xFFFFFFFFu, pi, 0);
        const bool run = __all_sync(0xFFFFFFFFu, pi == p0 + lane);      // 32 consecutive pivot columns
        uint64_t todo = freemask;
        while (todo != 0ull) {
            const int k = __ffsll((long long)todo) - 1;
            todo &= todo - 1ull;
            const unsigned rows = __ballot_sync(0xFFFFFFFFu, (v >> k) & 1ull);
            if (rows == 0u) continue;
            const int t = tbase + __popcll(freemask & ((1ull << k) - 1ull));
            if (t >= max_rows) continue;
            uint32_t* row = out + (size_t)t * W32;
            if (run) {
                if (lane == 0) {
                    const int sh = p0 & 31;
                    atomicOr(row + (p0 >> 5), rows << sh);
                    if (sh != 0 && (rows >> (32 - sh)) != 0u) atomicOr(row + (p0 >> 5) + 1, rows >> (32 - sh));
                }
            } else if ((rows >> lane) & 1u) {
                atomicOr(row + (pi >> 5), 1u << (pi & 31));
            }
        }
    }
}

// [A | b]: rows of ceil((n + 1) / 64) words, rhs bit i of matrix b at column n
__global__ void k_augment(const uint64_t* __restrict__ mats, const uint64_t* __restrict__ rhs, int batch, int m,
                          int n, uint64_t* __restrict__ aug) {
    const int W = (n + 63) >> 6, Wa = (n + 64) >> 6, Wr = (m + 63) >> 6;
    const size_t total = (size_t)batch * m * Wa;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int w = (int)(idx % Wa);
        const size_t row = idx / Wa;
        const int i = (int)(row % m);
        const size_t b = row / m;
        uint64_t v = 0ull;
        if (w < W) {
            v = mats[row * W + w];
            if (w == W - 1 && (n & 63)) v &= (1ull << (n & 63)) - 1ull;
        }
        if (w == (n >> 6)) v |= ((rhs[b * Wr + (i >> 6)] >> (i & 63)) & 1ull) << (n & 63);
        aug[idx] = v;
    }
}

__global__ void k_solve_extract(const uint64_t* __restrict__ rref, const int32_t* __restrict__ rank,
                                const int32_t* __restrict__ piv, int batch, int m, int n,
                                uint32_t* __restrict__ x, int32_t* __restrict__ consistent) {
    const int Wa = (n + 64) >> 6, W = (n + 63) >> 6;
    const int npiv = m < (n + 1) ? m : (n + 1);
    const int b = blockIdx.x;
    if (b >= batch) return;
    const int r = rank[b];
    const int32_t* p = piv + (size_t)b * npiv;
    const bool ok = (r == 0) || p[r - 1] != n;            // a pivot in the rhs column <=> 0 = 1
    if (threadIdx.x == 0 && consistent != nullptr) consistent[b] = ok ? 1 : 0;
    if (!ok) return;
    for (int i = threadIdx.x; i < r; i += blockDim.x) {
        const uint64_t w = rref[((size_t)b * m + i) * Wa + (n >> 6)];
        if ((w >> (n & 63)) & 1ull) atomicOr(x + (size_t)b * W * 2 + (p[i] >> 5), 1u << (p[i] & 31));
    }
}

struct Scratch {
    void* p = nullptr;
    cudaStream_t s;
    explicit Scratch(cudaStream_t st) : s(st) {}
    cudaError_t get(size_t bytes) { return cudaMallocAsync(&p, bytes, s); }
    ~Scratch() { if (p) cudaFreeAsync(p, s); }
};

}  // namespace

// overflow (device int, may be NULL): set to the largest number of basis rows any matrix needed when
// that exceeds max_rows (rows past the capacity are dropped, nothing is written out of bounds).
cudaError_t launch_gf2_nullspace(const uint64_t* d_mats, int batch, int m, int n, int max_rows, uint64_t* d_basis,
                                 int32_t* d_rank, int32_t* d_overflow, cudaStream_t stream) {
    if (batch <= 0 || m <= 0 || n <= 0) return cudaSuccess;
    const size_t W = (size_t)(n + 63) / 64, npiv = (size_t)(m < n ? m : n);
    const size_t mat_bytes = (size_t)batch * m * W * 8, piv_bytes = (size_t)batch * npiv * 4;
    Scratch sc(stream);
    cudaError_t err = sc.get(mat_bytes + piv_bytes + (size_t)batch * 4 + 16);
    if (err != cudaSuccess) return err;
    uint64_t* d_rref = (uint64_t*)sc.p;
    int32_t* d_piv = (int32_t*)((uint8_t*)sc.p + mat_bytes);
    int32_t* rank = d_rank != nullptr ? d_rank : (int32_t*)((uint8_t*)sc.p + mat_bytes + piv_bytes);
    int32_t* ovf = d_overflow != nullptr ? d_overflow : (int32_t*)((uint8_t*)sc.p + mat_bytes + piv_bytes + (size_t)batch * 4);
    if ((err = launch_gf2_rref(d_mats, batch, m, n, d_rref, rank, d_piv, stream)) != cudaSuccess) return err;
    if ((err = cudaMemsetAsync(d_basis, 0, (size_t)batch * max_rows * W * 8, stream)) != cudaSuccess) return err;
    if ((err = cudaMemsetAsync(ovf, 0, 4, stream)) != cudaSuccess) return err;
    if (max_rows > 0) {
        k_nullspace<<<(unsigned)((size_t)batch * W), kNsThreads, 0, stream>>>(d_rref, rank, d_piv, batch, m, n, max_rows,
                                                                             (uint32_t*)d_basis, ovf);
        if ((err = cudaGetLastError()) != cudaSuccess) return err;
    }
    return cudaSuccess;
}

cudaError_t launch_gf2_solve(const uint64_t* d_mats, const uint64_t* d_rhs, int batch, int m, int n, uint64_t* d_x,
                             int32_t* d_consistent, cudaStream_t stream) {
    if (batch <= 0 || m <= 0 || n <= 0) return cudaSuccess;
    const size_t W = (size_t)(n + 63) / 64, Wa = (size_t)(n + 64) / 64;
    const size_t npiv = (size_t)(m < n + 1 ? m : n + 1);
    const size_t aug_bytes = (size_t)batch * m * Wa * 8, piv_bytes = (size_t)batch * npiv * 4;
    Scratch sc(stream);
    cudaError_t err = sc.get(2 * aug_bytes + piv_bytes + (size_t)batch * 4);
    if (err != cudaSuccess) return err;
    uint64_t* d_aug = (uint64_t*)sc.p;
    uint64_t* d_rref = (uint64_t*)((uint8_t*)sc.p + aug_bytes);
    int32_t* d_piv = (int32_t*)((uint8_t*)sc.p + 2 * aug_bytes);
    int32_t* d_rank = (int32_t*)((uint8_t*)sc.p + 2 * aug_bytes + piv_bytes);
    const size_t total = (size_t)batch * m * Wa;
    const unsigned blocks = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    k_augment<<<blocks, 256, 0, stream>>>(d_mats, d_rhs, batch, m, n, d_aug);
    if ((err = cudaGetLastError()) != cudaSuccess) return err;
    if ((err = launch_gf2_rref(d_aug, batch, m, n + 1, d_rref, d_rank, d_piv, stream)) != cudaSuccess) return err;
    if ((err = cudaMemsetAsync(d_x, 0, (size_t)batch * W * 8, stream)) != cudaSuccess) return err;
    k_solve_extract<<<batch, 256, 0, stream>>>(d_rref, d_rank, d_piv, batch, m, n, (uint32_t*)d_x, d_consistent);
    return cudaGetLastError();
}

}  // namespace qcss
