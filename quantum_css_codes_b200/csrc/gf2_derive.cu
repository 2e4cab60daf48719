// K4 derived products: batched null space and linear solve on top of the batched RREF.
//
// The reference has neither (bin_matrix.py stops at reduced_row_echelon_form); BASELINE
// config 5 asks for "row reduction / rank / null space" of 4096 matrices, so both are built on the
// device from the RREF R (rank r, pivot columns p_0 < ... < p_{r-1}) without a round trip to the host:
//
//   null space   one basis vector per free column f, in increasing order of f:
//                    x[f] = 1,  x[p_i] = R[i][f],  every other free variable 0
//                i.e. the free columns of R transposed and scattered to the pivot positions.  A CTA
//                takes one matrix and 256 columns; a warp takes 32 rows, and a 32 x 32 shuffle transpose
//                turns "32 rows x 32 columns" into one word per free column -- 32 bits of that basis
//                vector, stored whole when the 32 pivots are consecutive aligned columns (the usual
//                case), with two atomic ORs when consecutive but unaligned, bit by bit otherwise.
//   solve        RREF of the augmented matrix [A | b]; inconsistent iff column n holds a pivot;
//                otherwise x[p_i] = R[i][n] with every free variable 0.
#include <cuda_runtime.h>

#include "launch.h"

namespace qcss {

namespace {

constexpr int kNsThreads = 256;
constexpr int kNsWords = 4;                       // 64-bit words (256 columns) of free-column candidates per CTA
static_assert(kNsThreads == 64 * kNsWords, "one thread per column of the group");

// 32 x 32 bit transpose across a warp: lane l enters with row l, leaves with column l (bit j = row j).
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, int lane) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const uint32_t m = d == 16 ? 0x0000FFFFu : d == 8 ? 0x00FF00FFu : d == 4 ? 0x0F0F0F0Fu : d == 2 ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xFFFFFFFFu, x, d);
        x = (lane & d) ? ((x & ~m) | ((y & ~m) >> d)) : ((x & m) | ((y & m) << d));
    }
    return x;
}

// One CTA per (matrix, group of kNsWords words of columns).  A warp takes 32 rows; each lane loads its row's
// words of the group, and a shuffle transpose turns "32 rows x 32 columns" into one 32-bit word per free
// column -- 32 bits of that column's basis vector at the rows' pivot positions.  When the 32 pivots are
// consecutive columns (the usual case) the word lands with one plain store (aligned) or two atomic ORs;
// otherwise it is cut at the skipped columns into sub-runs, two atomic ORs each.  (First version: one ballot
// per free column and lane 0 doing every atomic -- 10.4 ms for C5.  Second: this transpose with a bit-by-bit
// fallback -- 6.0 ms, 56 % of the instructions in the fallback although only the last 32 rows of a random
// matrix, where pivot columns start to skip, take it: profiles/r01_gf2_nullspace_ncu_lines.txt.)
__global__ void __launch_bounds__(kNsThreads)
k_nullspace(const uint64_t* __restrict__ rref, const int32_t* __restrict__ rank, const int32_t* __restrict__ piv,
            int batch, int m, int n, int max_rows, uint32_t* __restrict__ basis, int32_t* __restrict__ overflow) {
    const int W = (n + 63) >> 6, G = (W + kNsWords - 1) / kNsWords;
    const int npiv = m < n ? m : n;
    const int b = blockIdx.x / G, fw0 = (blockIdx.x % G) * kNsWords;
    if (b >= batch) return;
    // pivot columns are sorted: thread c looks its column 64 * fw0 + c up by binary search, a ballot per warp
    // assembles the masks (no shared-memory atomics: 64-bit ones are CAS loops and all of a CTA's would collide)
    __shared__ uint32_t s_pivmask[2 * kNsWords];
    __shared__ int s_before;
    const int r = rank[b];
    const int32_t* p = piv + (size_t)b * npiv;
    {
        const int col = 64 * fw0 + threadIdx.x;
        int lo = 0, hi = r;                                  // first index with p[idx] >= col
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (p[mid] < col) lo = mid + 1; else hi = mid;
        }
        const bool is_pivot = lo < r && p[lo] == col;
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, is_pivot);
        if ((threadIdx.x & 31) == 0) s_pivmask[threadIdx.x >> 5] = mask;
        if (threadIdx.x == 0) s_before = lo;                 // pivots left of the group
    }
    __syncthreads();
    uint64_t freemask[kNsWords];
    int tbase[kNsWords];
    int run_t = 64 * fw0 - s_before, any_free = 0;     // basis row of the first free column of the group
#pragma unroll
    for (int k = 0; k < kNsWords; ++k) {
        const int cols_here = n - 64 * (fw0 + k);
        const uint64_t valid = cols_here >= 64 ? ~0ull : (cols_here > 0 ? ((1ull << cols_here) - 1ull) : 0ull);
        freemask[k] = ~((uint64_t)s_pivmask[2 * k] | ((uint64_t)s_pivmask[2 * k + 1] << 32)) & valid;
        tbase[k] = run_t;
        run_t += __popcll(freemask[k]);
        any_free |= freemask[k] != 0ull;
    }
    if (!any_free) return;
    if (run_t > max_rows && threadIdx.x == 0) atomicMax(overflow, run_t);
    const size_t W32 = (size_t)W * 2;
    uint32_t* out = basis + (size_t)b * max_rows * W32;
    const uint64_t* R = rref + (size_t)b * m * W;
    // x[f] = 1: thread c of the CTA owns column 64 * fw0 + c
    {
        const int k = threadIdx.x >> 6, c = threadIdx.x & 63;
        uint64_t fm = 0ull;
        int tb = 0;
#pragma unroll
        for (int u = 0; u < kNsWords; ++u)
            if (u == k) { fm = freemask[u]; tb = tbase[u]; }
        if (k < kNsWords && ((fm >> c) & 1ull)) {
            const int t = tb + __popcll(fm & ((1ull << c) - 1ull));
            const int f = 64 * (fw0 + k) + c;
            if (t < max_rows) atomicOr(out + (size_t)t * W32 + (f >> 5), 1u << (f & 31));
        }
    }
    // x[p_i] = R[i][f]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i0 = warp * 32; i0 < r; i0 += (kNsThreads / 32) * 32) {
        const int i = i0 + lane;
        const int pi = (i < r) ? p[i] : -1;
        const int p0 = __shfl_sync(0xFFFFFFFFu, pi, 0);
        const bool run = __all_sync(0xFFFFFFFFu, pi == p0 + lane);      // 32 consecutive pivot columns
        // otherwise: maximal sub-runs of consecutive pivot columns (a skipped column starts a new one)
        const int prev = __shfl_up_sync(0xFFFFFFFFu, pi, 1);
        const uint32_t starts = __ballot_sync(0xFFFFFFFFu, lane == 0 || pi != prev + 1);
#pragma unroll
        for (int k = 0; k < kNsWords; ++k) {
            if (freemask[k] == 0ull) continue;                           // uniform over the CTA
            const uint64_t v = (i < r && fw0 + k < W) ? (R[(size_t)i * W + fw0 + k] & freemask[k]) : 0ull;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t fm_half = (uint32_t)(freemask[k] >> (32 * h));
                if (fm_half == 0u) continue;
                const uint32_t rows = warp_transpose32((uint32_t)(v >> (32 * h)), lane);   // column 32h + lane
                if (rows == 0u || !((fm_half >> lane) & 1u)) continue;
                const int t = tbase[k] + __popcll(freemask[k] & ((1ull << (32 * h + lane)) - 1ull));
                if (t >= max_rows) continue;
                uint32_t* row = out + (size_t)t * W32;
                if (run) {
                    const int sh = p0 & 31;
                    if (sh == 0) {
                        row[p0 >> 5] = rows;          // the 32 pivots own this word: no other writer, no free column
                    } else {
                        atomicOr(row + (p0 >> 5), rows << sh);
                        if ((rows >> (32 - sh)) != 0u) atomicOr(row + (p0 >> 5) + 1, rows >> (32 - sh));
                    }
                } else {
                    uint32_t todo = starts;
                    while (todo != 0u) {
                        const int j0 = __ffs((int)todo) - 1;
                        todo &= todo - 1u;
                        const int len = (todo != 0u ? __ffs((int)todo) - 1 : 32) - j0;
                        const uint32_t seg = (rows >> j0) & (len == 32 ? 0xFFFFFFFFu : ((1u << len) - 1u));
                        if (seg == 0u) continue;
                        const int pj = p[i0 + j0], sh = pj & 31;
                        atomicOr(row + (pj >> 5), seg << sh);
                        if (sh != 0 && (seg >> (32 - sh)) != 0u) atomicOr(row + (pj >> 5) + 1, seg >> (32 - sh));
                    }
                }
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
        const size_t groups = (W + kNsWords - 1) / kNsWords;
        k_nullspace<<<(unsigned)((size_t)batch * groups), kNsThreads, 0, stream>>>(d_rref, rank, d_piv, batch, m, n, max_rows,
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
