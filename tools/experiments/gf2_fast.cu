// K4, fast path: batched GF(2) Gauss-Jordan for matrices with up to 1024 rows (any width).
// Replaces the per-column Python loop of bin_matrix.reduced_row_echelon_form (bin_matrix.py:8-34)
// with a blocked "four Russians" elimination that keeps the working set in registers.
//
// One CTA per matrix, one thread per row (blockDim = rows rounded up to 32).  The matrix is walked
// in SLABS of 1024 columns (32 words).  Inside a slab, warp w holds rows 32w..32w+31 in registers:
// lane l keeps word l of each of the 32 rows (r[i] = word l of row 32w+i), so a row operation is
// one conflict-free 128-byte shared-memory read per row.
//
//   discovery (pivot slab): for each strip of 8 columns, every thread keeps the reduced strip byte
//     of "its" row (thread t <-> row t); per column one ballot + one block barrier picks an unused
//     row with a 1 there (any row will do: the RREF is canonical, the pivot order is free), and
//     each row records whether it must receive that pivot (elimination bit z).
//   apply: for a block of <= 8 pivots, row_i ^= sum_u y_i[u] * P_u with P_u the pivot rows as they
//     are at block start and y_i = z_i * L (L unwinds "pivot u had pivot v added before it was
//     chosen").  All 256 combinations of the P_u are tabulated in shared memory (TP), so a row
//     update is ONE table read per row per 8 pivots instead of up to 8 row XORs.
//   replay (later slabs): the elimination bits Z (1 bit per row per pivot, <= 128 KB) are kept in
//     shared memory, so columns right of the pivot slab get exactly the same row operations, 8
//     pivots at a time, without ever looking at the left part again.
//
// Rows leave the CTA in pivot order (row holding pivot k -> output row k), zero rows last: the
// canonical RREF the reference returns.
#include <cuda_runtime.h>

#include "launch.h"

namespace qcss {

namespace {

constexpr int kSlabWords = 32;

struct FastLayout {
    int mpad;         // threads per CTA = rows rounded up to a multiple of 32
    int kmax;         // max pivots = min(m, n)
    int zrows;        // bytes of Z per row (+1 spill byte)
    size_t off_tp, off_p, off_s, off_pivrow, off_rowpiv, off_pivcol, off_cand, off_lrow, total;
};

__host__ __device__ inline FastLayout make_layout(int m, int n) {
    FastLayout L;
    L.mpad = (m + 31) & ~31;
    L.kmax = m < n ? m : n;
    L.zrows = (L.kmax + 7) / 8 + 1;
    size_t off = (size_t)L.zrows * L.mpad;
    off = (off + 127) & ~(size_t)127;
    L.off_tp = off;      off += 256 * kSlabWords * sizeof(uint32_t);
    L.off_p = off;       off += 8 * kSlabWords * sizeof(uint32_t);
    L.off_s = off;       off += L.mpad;
    L.off_pivrow = off;  off += (size_t)L.kmax * sizeof(int16_t);
    off = (off + 3) & ~(size_t)3;
    L.off_rowpiv = off;  off += (size_t)L.mpad * sizeof(int16_t);
    off = (off + 3) & ~(size_t)3;
    L.off_pivcol = off;  off += (size_t)L.kmax * sizeof(int32_t);
    off = (off + 7) & ~(size_t)7;
    L.off_cand = off;    off += 2 * 32 * sizeof(uint16_t);
    L.off_lrow = off;    off += 8;                      // 8-byte aligned: read as one uint2
    L.total = (off + 15) & ~(size_t)15;
    return L;
}

__device__ __forceinline__ uint32_t select_reg(const uint32_t (&r)[32], int idx) {
    uint32_t v = 0u;
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i == idx) v = r[i];
    return v;
}

__global__ void __launch_bounds__(1024, 1)
k_gf2_fast(const uint32_t* __restrict__ in, int batch, int m, int n, uint32_t* __restrict__ out,
           int32_t* __restrict__ rank_out, int32_t* __restrict__ piv_out) {
    extern __shared__ __align__(128) uint8_t smem[];
    const FastLayout L = make_layout(m, n);
    uint8_t* Zb = smem;                                                   // [zrows][mpad]
    uint32_t* TP = reinterpret_cast<uint32_t*>(smem + L.off_tp);          // [256][32]
    uint32_t* P = reinterpret_cast<uint32_t*>(smem + L.off_p);            // [8][32]
    uint8_t* S = smem + L.off_s;                                          // [mpad] strip bytes
    int16_t* pivrow = reinterpret_cast<int16_t*>(smem + L.off_pivrow);    // [kmax]
    int16_t* rowpiv = reinterpret_cast<int16_t*>(smem + L.off_rowpiv);    // [mpad]
    int32_t* pivcol = reinterpret_cast<int32_t*>(smem + L.off_pivcol);    // [kmax]
    uint16_t* cand = reinterpret_cast<uint16_t*>(smem + L.off_cand);      // [2][32]
    uint8_t* Lrow = smem + L.off_lrow;                                    // [8]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nwarps = blockDim.x >> 5;
    const int mpad = L.mpad;
    const int W32 = ((n + 63) >> 6) * 2;                 // 32-bit words per packed row
    const int nslabs = (W32 + kSlabWords - 1) / kSlabWords;
    const int npiv = L.kmax;

    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        const uint32_t* src = in + (size_t)b * m * W32;
        uint32_t* dst = out + (size_t)b * m * W32;
        // ---- per-matrix state ----------------------------------------------------------------
        {
            uint4* z4 = reinterpret_cast<uint4*>(Zb);
            const int n16 = (int)(L.off_tp / 16);
            for (int i = tid; i < n16; i += blockDim.x) z4[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        rowpiv[tid] = -1;
        bool used = tid >= m;                            // padding rows never become pivots
        int K = 0;                                       // pivots found so far (uniform)
        __syncthreads();

        // Applies pivots [t0, t0+k) to the slab held in r[].
        uint32_t r[32];
        auto apply = [&](int t0, int k) {
            // A1: owners publish the pivot rows' words of this slab.  Lane l looks at row 32*warp+l;
            // a ballot tells the warp which of its rows (usually none or one) are pivots of the block.
            {
                const int pk = (int)rowpiv[tid] - t0;
                unsigned mine = __ballot_sync(0xFFFFFFFFu, (unsigned)pk < (unsigned)k);
                while (mine != 0u) {
                    const int i = __ffs(mine) - 1;
                    mine &= mine - 1u;
                    const int u = __shfl_sync(0xFFFFFFFFu, pk, i);
                    P[u * kSlabWords + lane] = select_reg(r, i);
                }
            }
            // A2: warp 0 unwinds the order dependence between the block's pivots
            const int byte0 = t0 >> 3, off = t0 & 7;
            if (warp == 0) {
                uint32_t zrow = 0u;
                if (lane < k) {
                    const int p = pivrow[t0 + lane];
                    zrow = (((uint32_t)Zb[(size_t)byte0 * mpad + p] |
                             ((uint32_t)Zb[(size_t)(byte0 + 1) * mpad + p] << 8)) >> off);
                }
                uint32_t lr = 1u << lane;
#pragma unroll
                for (int v = 0; v < 7; ++v) {
                    const uint32_t lv = __shfl_sync(0xFFFFFFFFu, lr, v);
                    if (lane > v && lane < k && ((zrow >> v) & 1u)) lr ^= lv;
                }
                if (lane < 8) Lrow[lane] = (lane < k) ? (uint8_t)lr : (uint8_t)0;
            }
            __syncthreads();
            // A3: combination byte of my own row: y = (z with my own pivot bit dropped) * L
            uint32_t y = 0u;
            {
                uint32_t z = (((uint32_t)Zb[(size_t)byte0 * mpad + tid] |
                               ((uint32_t)Zb[(size_t)(byte0 + 1) * mpad + tid] << 8)) >> off) & ((1u << k) - 1u);
                const int self = (int)rowpiv[tid] - t0;
                if (self >= 0 && self < k) z &= ~(1u << self);
                const uint2 lr = *reinterpret_cast<const uint2*>(Lrow);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t lu = ((u < 4 ? lr.x : lr.y) >> (8 * (u & 3))) & 0xFFu;
                    if ((z >> u) & 1u) y ^= lu;
                }
            }
            // A3': table of all combinations of the k pivot rows (entries beyond 2^k are unused)
            {
                uint32_t pu[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) pu[u] = (u < k) ? P[u * kSlabWords + lane] : 0u;
                const int entries = 1 << k;
                for (int e0 = warp * 8; e0 < entries; e0 += nwarps * 8) {
                    uint32_t base = 0u;
#pragma unroll
                    for (int u = 3; u < 8; ++u)
                        if ((e0 >> u) & 1) base ^= pu[u];
                    const uint32_t c1 = base ^ pu[0], c2 = base ^ pu[1], c3 = c1 ^ pu[1];
                    const uint32_t c4 = base ^ pu[2], c5 = c1 ^ pu[2], c6 = c2 ^ pu[2], c7 = c3 ^ pu[2];
                    uint32_t* t = TP + (size_t)e0 * kSlabWords + lane;
                    t[0 * kSlabWords] = base; t[1 * kSlabWords] = c1; t[2 * kSlabWords] = c2;
                    t[3 * kSlabWords] = c3;   t[4 * kSlabWords] = c4; t[5 * kSlabWords] = c5;
                    t[6 * kSlabWords] = c6;   t[7 * kSlabWords] = c7;
                }
            }
            __syncthreads();
            // A4: one table read per row
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const uint32_t yi = __shfl_sync(0xFFFFFFFFu, y, i);
                r[i] ^= TP[yi * kSlabWords + lane];
            }
        };

        for (int slab = 0; slab < nslabs; ++slab) {
            const int wi = slab * kSlabWords + lane;
            // ---- load the slab into registers (columns >= n masked off) --------------------------
            uint32_t colmask = 0u;
            if (wi < W32) {
                const int c_lo = wi * 32;
                colmask = (c_lo + 32 <= n) ? 0xFFFFFFFFu : (c_lo < n ? ((1u << (n - c_lo)) - 1u) : 0u);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int row = warp * 32 + i;
                r[i] = (row < m && colmask != 0u) ? (__ldg(src + (size_t)row * W32 + wi) & colmask) : 0u;
            }
            // ---- replay every pivot found in earlier slabs ----------------------------------------
            for (int t0 = 0; t0 < K; t0 += 8) apply(t0, (K - t0) < 8 ? (K - t0) : 8);
            // ---- discovery: strips of 8 columns of this slab --------------------------------------
            const int slab_words = (W32 - slab * kSlabWords) < kSlabWords ? (W32 - slab * kSlabWords) : kSlabWords;
            for (int cw = 0; cw < slab_words && K < m; ++cw) {
                for (int sb = 0; sb < 4 && K < m; ++sb) {
                    const int c0 = (slab * kSlabWords + cw) * 32 + sb * 8;
                    if (c0 >= n) break;
                    // (a) current strip byte of every row -> S (lane cw of each warp owns the word)
                    {
                        const uint32_t sel = (uint32_t)sb | ((uint32_t)(4 + sb) << 4);
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const uint32_t t1 = __byte_perm(r[i], r[i + 1], sel);
                            const uint32_t t2 = __byte_perm(r[i + 2], r[i + 3], sel);
                            const uint32_t w4 = __byte_perm(t1, t2, 0x5410u);
                            if (lane == cw) *reinterpret_cast<uint32_t*>(S + warp * 32 + i) = w4;
                        }
                    }
                    __syncthreads();
                    uint32_t br = S[tid];                // reduced strip byte of my row
                    uint32_t z = 0u;
                    int k = 0;
#pragma unroll
                    for (int col = 0; col < 8; ++col) {
                        if (c0 + col < n && K + k < m) {
                            const bool hit = !used && ((br >> col) & 1u);
                            const unsigned vote = __ballot_sync(0xFFFFFFFFu, hit);
                            const int srcl = vote ? (__ffs(vote) - 1) : 0;
                            const uint32_t vb = __shfl_sync(0xFFFFFFFFu, br, srcl);
                            uint16_t* cbuf = cand + (col & 1) * 32;
                            if (lane == 0) cbuf[warp] = vote ? (uint16_t)(srcl | (vb << 8)) : (uint16_t)0xFFFFu;
                            __syncthreads();
                            const uint32_t cv = (lane < nwarps) ? cbuf[lane] : 0xFFFFu;
                            const unsigned wv = __ballot_sync(0xFFFFFFFFu, (cv & 0xFFu) != 0xFFu);
                            if (wv != 0u) {
                                const int wmin = __ffs(wv) - 1;
                                const uint32_t e = __shfl_sync(0xFFFFFFFFu, cv, wmin);
                                const int p = wmin * 32 + (int)(e & 0xFFu);
                                const uint32_t v = e >> 8;
                                const uint32_t zb = (br >> col) & 1u;
                                z |= zb << k;
                                if (tid == p) {
                                    used = true;
                                    rowpiv[tid] = (int16_t)(K + k);
                                    pivrow[K + k] = (int16_t)p;
                                    pivcol[K + k] = c0 + col;
                                } else if (zb) {
                                    br ^= v;
                                }
                                ++k;
                            }
                        }
                    }
                    if (k > 0) {
                        // (c) append my row's k elimination bits to Z at bit offset K
                        const int byte0 = K >> 3, off = K & 7;
                        const uint32_t zz = z << off;
                        Zb[(size_t)byte0 * mpad + tid] |= (uint8_t)(zz & 0xFFu);
                        if (off + k > 8) Zb[(size_t)(byte0 + 1) * mpad + tid] |= (uint8_t)(zz >> 8);
                    }
                    __syncthreads();
                    if (k > 0) {
                        apply(K, k);
                        K += k;
                    }
                }
            }
            // ---- write the slab out in pivot order; rows without a pivot so far are zero here -----
            __syncthreads();
            if (wi < W32) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int pk = rowpiv[warp * 32 + i];
                    if (pk >= 0) dst[(size_t)pk * W32 + wi] = r[i];
                }
                for (int row = K + warp; row < m; row += nwarps) dst[(size_t)row * W32 + wi] = 0u;
            }
        }
        // ---- rank and pivot columns -------------------------------------------------------------
        __syncthreads();
        if (tid == 0 && rank_out != nullptr) rank_out[b] = K;
        if (piv_out != nullptr)
            for (int t = tid; t < npiv; t += blockDim.x) piv_out[(size_t)b * npiv + t] = (t < K) ? pivcol[t] : -1;
        __syncthreads();
    }
}

}  // namespace

bool gf2_fast_supported(int m, int n) {
    if (m < 1 || m > 1024 || n < 1) return false;
    return make_layout(m, n).total <= 224 * 1024;
}

cudaError_t launch_gf2_fast(const uint64_t* in, int batch, int m, int n, uint64_t* out, int32_t* rank,
                            int32_t* pivots, cudaStream_t stream) {
    const FastLayout L = make_layout(m, n);
    cudaError_t err = cudaFuncSetAttribute(k_gf2_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
    if (err != cudaSuccess) return err;
    int dev = 0, sms = 0, per_sm = 1;
    if ((err = cudaGetDevice(&dev)) != cudaSuccess) return err;
    if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return err;
    if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gf2_fast, L.mpad, L.total)) != cudaSuccess)
        return err;
    if (per_sm < 1) per_sm = 1;
    int grid = sms * per_sm;
    if (grid > batch) grid = batch;
    k_gf2_fast<<<grid, L.mpad, L.total, stream>>>(reinterpret_cast<const uint32_t*>(in), batch, m, n,
                                                 reinterpret_cast<uint32_t*>(out), rank, pivots);
    return cudaGetLastError();
}

}  // namespace qcss
