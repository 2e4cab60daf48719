// K4, second generation: batched GF(2) Gauss-Jordan for matrices with up to 1024 rows (any width).
// Replaces the per-column Python loop of bin_matrix.reduced_row_echelon_form (bin_matrix.py:8-34).
//
// Same data placement as gf2_fast.cu -- one CTA per matrix, the matrix walked in SLABS of 1024
// columns held in registers (warp w keeps rows 32w..32w+31, lane l word l of each, r[i] = word l of
// row 32w+i), row updates through a 256-entry "four Russians" table in shared memory -- but the
// bookkeeping around the table reads is cut to a fraction (gf2_fast.cu spent 80 % of its issue slots
// there, profiles/r01_gf2_fast_ncu_summary.txt):
//
//   panel     Eliminating an 8-column strip acts on a row through its strip BYTE only, so the whole
//             1024 x 8 panel is factored in "byte space" by one warp: lane l owns the byte values
//             8l..8l+7 (reduced value and combination byte, four values per register, SIMD within
//             the register) plus one representative unused row per value (REP, written by the rows
//             themselves).  A pivot is one ballot; eliminating all 256 values is two LOP3 per
//             register.  The result is a 256-byte table G: strip byte -> combination y of the
//             strip-start pivot rows that a row with that byte receives (the <= 8 pivot rows
//             themselves are tracked separately, PY).  Every row then gets its y with ONE lookup.
//   apply     pivot rows are published from registers, all 2^k combinations are tabulated (TP), and
//             a row update is ONE table read:  r[i] ^= TP[y_i].  The combination bytes of a warp's
//             32 rows travel as 8 packed words through shared memory, the address is one shift +
//             one LOP3.  In the pivot slab the table reads of block s are deferred until the panel of
//             strip s+1 is being factored (look-ahead: the panel only needs each row's pivot word,
//             which is tracked per thread through TW, the pivot-word column of TP).
//   replay    y is kept bit-sliced per pivot (Ys, 128 bytes per pivot, <= 129 KB), so the slabs to
//             the right repeat the same blocks with no recomputation: gather y, tabulate, apply.
//
// Rows leave the CTA in pivot order (row holding pivot k -> output row k), zero rows last: the
// canonical RREF the reference returns.  Any unused row with a 1 may serve as pivot because the
// RREF is unique.
#include <cuda_runtime.h>

#include <cstdlib>

#include "launch.h"

namespace qcss {

namespace {

constexpr int kSlabWords = 32;

// Fixed shared-memory map (bytes): every offset is a compile-time constant so no address lives in
// a register.  Ys (the only array whose size depends on the matrix) comes last.
constexpr int kOffTP = 0;                          // uint32 [256][32]  combination table
constexpr int kOffTW = kOffTP + 256 * 32 * 4;      // uint32 [256]      its pivot-word column
constexpr int kOffP = kOffTW + 256 * 4;            // uint32 [8][32]    published pivot rows
constexpr int kOffG = kOffP + 8 * 32 * 4;          // uint8  [256]      strip byte -> y
constexpr int kOffPY = kOffG + 256;                // uint8  [8] (+8)   y of the strip's pivot rows
constexpr int kOffMisc = kOffPY + 16;              // int    [4]
constexpr int kOffRep = kOffMisc + 16;             // uint16 [2][256]   representative row | 0x8000
constexpr int kOffS32 = kOffRep + 2 * 256 * 2;     // uint32 [1024]     word transpose
constexpr int kOffYb = kOffS32 + 1024 * 4;         // uint8  [1024]     combination bytes of a block
constexpr int kOffRowpiv = kOffYb + 1024;          // int16  [1024]
constexpr int kOffPivrow = kOffRowpiv + 1024 * 2;  // int16  [1024]
constexpr int kOffPivcol = kOffPivrow + 1024 * 2;  // int32  [1024]
constexpr int kOffBlk = kOffPivcol + 1024 * 4;     // uint32 [1024]     (K | k << 16) per block
constexpr int kOffYs = (kOffBlk + 1024 * 4 + 127) & ~127;   // uint32 [kmax + 8][nw]

__host__ __device__ inline size_t m4r_smem_bytes(int m, int n) {
    const int nw = (m + 31) >> 5;
    const int kmax = m < n ? m : n;
    return (size_t)kOffYs + (size_t)(kmax + 8) * nw * sizeof(uint32_t);
}

__device__ __forceinline__ uint32_t pick_reg(const uint32_t (&r)[32], int idx) {
    uint32_t v = 0u;
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i == idx) v = r[i];
    return v;
}

// NWC: warps per CTA when known at compile time (32 for the 1024-row case), 0 = runtime.
template <int NWC, bool LOOKAHEAD>
__global__ void __launch_bounds__(1024, 1)
k_gf2_m4r(const uint32_t* __restrict__ in, int batch, int m, int n, uint32_t* __restrict__ out,
          int32_t* __restrict__ rank_out, int32_t* __restrict__ piv_out) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* const TP = reinterpret_cast<uint32_t*>(smem + kOffTP);
    uint32_t* const TW = reinterpret_cast<uint32_t*>(smem + kOffTW);
    uint32_t* const P = reinterpret_cast<uint32_t*>(smem + kOffP);
    uint8_t* const G = smem + kOffG;
    uint8_t* const PY = smem + kOffPY;
    volatile int* const misc = reinterpret_cast<volatile int*>(smem + kOffMisc);
    uint16_t* const REP = reinterpret_cast<uint16_t*>(smem + kOffRep);
    uint32_t* const S32 = reinterpret_cast<uint32_t*>(smem + kOffS32);
    uint8_t* const Yb = smem + kOffYb;
    int16_t* const rowpiv = reinterpret_cast<int16_t*>(smem + kOffRowpiv);
    int16_t* const pivrow = reinterpret_cast<int16_t*>(smem + kOffPivrow);
    int32_t* const pivcol = reinterpret_cast<int32_t*>(smem + kOffPivcol);
    uint32_t* const blk = reinterpret_cast<uint32_t*>(smem + kOffBlk);
    uint32_t* const Ys = reinterpret_cast<uint32_t*>(smem + kOffYs);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nw = NWC ? NWC : (int)(blockDim.x >> 5);
    const int W32 = ((n + 63) >> 6) * 2;                 // 32-bit words per packed row
    const int nslabs = (W32 + kSlabWords - 1) / kSlabWords;
    const int npiv = m < n ? m : n;
    const uint32_t lane4 = (uint32_t)lane * 4u;

    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        const uint32_t* src = in + (size_t)b * m * W32;
        uint32_t* dst = out + (size_t)b * m * W32;
        rowpiv[tid] = -1;
        for (int i = tid; i < 256; i += blockDim.x) reinterpret_cast<uint32_t*>(REP)[i] = 0u;   // both buffers
        bool used = tid >= m;                            // padding rows never become pivots
        int K = 0, nblk = 0, strip_no = 0;
        __syncthreads();

        uint32_t r[32];
        uint32_t cur = 0u;

        // ---- pieces of one block [Kb, Kb + k) -------------------------------------------------------
        // publish: rows that are pivots of the block write their slab words to P
        auto publish = [&](int Kb, int k, int pk_abs) {
            const int pk = pk_abs - Kb;
            unsigned mine = __ballot_sync(0xFFFFFFFFu, (unsigned)pk < (unsigned)k);
            while (mine != 0u) {
                const int i = __ffs(mine) - 1;
                mine &= mine - 1u;
                const int u = __shfl_sync(0xFFFFFFFFu, pk, i);
                P[u * kSlabWords + lane] = pick_reg(r, i);
            }
        };
        // tabulate: all combinations of the k published rows (entries beyond 2^k are never read)
        auto tabulate = [&](int k, bool piv, int cw) {
            uint32_t pu[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) pu[u] = (u < k) ? P[u * kSlabWords + lane] : 0u;
            const int entries = 1 << k;
            for (int e0 = warp * 8; e0 < entries; e0 += nw * 8) {
                uint32_t base = 0u;
#pragma unroll
                for (int u = 3; u < 8; ++u)
                    if ((e0 >> u) & 1) base ^= pu[u];
                const uint32_t c1 = base ^ pu[0], c2 = base ^ pu[1], c3 = c1 ^ pu[1];
                const uint32_t c4 = base ^ pu[2], c5 = c1 ^ pu[2], c6 = c2 ^ pu[2], c7 = c3 ^ pu[2];
                uint32_t* t = TP + e0 * kSlabWords + lane;
                t[0 * kSlabWords] = base; t[1 * kSlabWords] = c1; t[2 * kSlabWords] = c2;
                t[3 * kSlabWords] = c3;   t[4 * kSlabWords] = c4; t[5 * kSlabWords] = c5;
                t[6 * kSlabWords] = c6;   t[7 * kSlabWords] = c7;
                if (piv && lane == cw) {
                    uint4* w4 = reinterpret_cast<uint4*>(TW + e0);
                    w4[0] = make_uint4(base, c1, c2, c3);
                    w4[1] = make_uint4(c4, c5, c6, c7);
                }
            }
        };
        // table_reads: one read per row; the 32 combination bytes of this warp arrive as 8 words
        auto table_reads = [&]() {
            const uint4* yw = reinterpret_cast<const uint4*>(Yb + warp * 32);
            const uint4 ya = yw[0], yb = yw[1];
            const uint32_t yy[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
            const uint8_t* tp = smem + kOffTP;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint32_t w = yy[q];
                const uint32_t a0 = ((w << 7) & 0x7F80u) | lane4;
                const uint32_t a1 = ((w >> 1) & 0x7F80u) | lane4;
                const uint32_t a2 = ((w >> 9) & 0x7F80u) | lane4;
                const uint32_t a3 = ((w >> 17) & 0x7F80u) | lane4;
                r[4 * q + 0] ^= *reinterpret_cast<const uint32_t*>(tp + a0);
                r[4 * q + 1] ^= *reinterpret_cast<const uint32_t*>(tp + a1);
                r[4 * q + 2] ^= *reinterpret_cast<const uint32_t*>(tp + a2);
                r[4 * q + 3] ^= *reinterpret_cast<const uint32_t*>(tp + a3);
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
            // ---- replay every block found in earlier slabs ----------------------------------------
            if (nblk > 0) {
                const int mypiv = rowpiv[tid];
                for (int bi = 0; bi < nblk; ++bi) {
                    const uint32_t e = blk[bi];
                    const int Kb = (int)(e & 0xFFFFu), k = (int)(e >> 16);
                    // gather y from the bit-sliced store (rows past the block hold junk: masked off)
                    uint32_t y = 0u;
                    {
                        const uint32_t* ys = Ys + Kb * nw + warp;
#pragma unroll
                        for (int j = 7; j >= 0; --j) y = y * 2u + ((ys[j * nw] >> lane) & 1u);
                        y &= (1u << k) - 1u;
                    }
                    __syncwarp();                        // lanes still reading the previous block's bytes
                    Yb[tid] = (uint8_t)y;
                    publish(Kb, k, mypiv);
                    __syncthreads();
                    tabulate(k, false, 0);
                    __syncthreads();
                    table_reads();
                }
            }
            // ---- discovery: strips of 8 columns of this slab --------------------------------------
            const int slab_words = (W32 - slab * kSlabWords) < kSlabWords ? (W32 - slab * kSlabWords) : kSlabWords;
            for (int cw = 0; cw < slab_words && K < m; ++cw) {
                if (slab * kSlabWords * 32 + cw * 32 >= n) break;
                // my row's current word cw, out of lane cw's registers (r[] is up to date here)
                if (lane == cw) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) S32[warp * 32 + i] = r[i];
                }
                __syncwarp();
                cur = S32[tid];
                bool pending = false;                    // table reads of the last block not done yet
                for (int sb = 0; sb < 4 && K < m; ++sb) {
                    const int c0 = (slab * kSlabWords + cw) * 32 + sb * 8;
                    if (c0 >= n) break;
                    // (a) every unused row offers itself as the representative of its strip byte
                    const uint32_t byte = (cur >> (8 * sb)) & 0xFFu;
                    uint16_t* rep = REP + (strip_no & 1) * 256;
                    for (int i = tid; i < 128; i += blockDim.x)
                        reinterpret_cast<uint32_t*>(REP + ((strip_no + 1) & 1) * 256)[i] = 0u;
                    if (!used) rep[byte] = (uint16_t)(tid | 0x8000);
                    ++strip_no;
                    __syncthreads();
                    // (b) warp 0 factors the panel in byte space; the others do the deferred table reads
                    if (warp == 0) {
                        const uint4 q = reinterpret_cast<const uint4*>(rep)[lane];   // values 8l..8l+7
                        const uint32_t prs0 = (__byte_perm(q.x, q.y, 0x7531u) >> 7) & 0x01010101u;
                        const uint32_t prs1 = (__byte_perm(q.z, q.w, 0x7531u) >> 7) & 0x01010101u;
                        uint32_t red0 = 0x03020100u + 0x08080808u * (uint32_t)lane, red1 = red0 + 0x04040404u;
                        uint32_t y0 = 0u, y1 = 0u;
                        uint32_t pred = 0u, py = 0u;     // lane u < k: strip byte and y of pivot row u
                        int k = 0;
#pragma unroll
                        for (int col = 0; col < 8; ++col) {
                            const uint32_t m0 = (red0 >> col) & 0x01010101u, m1 = (red1 >> col) & 0x01010101u;
                            const uint32_t cand0 = m0 & prs0, cand1 = m1 & prs1;
                            const unsigned vote = __ballot_sync(0xFFFFFFFFu, (cand0 | cand1) != 0u);
                            if (vote != 0u) {
                                const int srcl = __ffs(vote) - 1;
                                // my first candidate value e (0..7) and its reduced byte / y / row
                                const uint32_t e = cand0 ? (uint32_t)(__ffs(cand0) - 1) >> 3
                                                         : 4u + ((uint32_t)(__ffs(cand1) - 1) >> 3);
                                const uint32_t v_me = __byte_perm(red0, red1, e) & 0xFFu;
                                const uint32_t y_me = __byte_perm(y0, y1, e) & 0xFFu;
                                const uint32_t ra = (e & 4u) ? q.z : q.x, rb = (e & 4u) ? q.w : q.y;
                                const uint32_t e2 = (e & 3u) * 2u;
                                const uint32_t row_me = __byte_perm(ra, rb, e2 | ((e2 + 1u) << 4)) & 0x3FFu;
                                uint32_t pack = v_me | (y_me << 8) | (row_me << 16);
                                pack = __shfl_sync(0xFFFFFFFFu, pack, srcl);
                                const uint32_t v = pack & 0xFFu, yp = (pack >> 8) & 0xFFu, prow = pack >> 16;
                                const uint32_t yk = yp | (1u << k);
                                const uint32_t v4 = v * 0x01010101u, yk4 = yk * 0x01010101u;
                                const uint32_t M0 = m0 * 0xFFu, M1 = m1 * 0xFFu;
                                red0 ^= M0 & v4;  red1 ^= M1 & v4;
                                y0 ^= M0 & yk4;   y1 ^= M1 & yk4;
                                if (lane < k && ((pred >> col) & 1u)) { pred ^= v; py ^= yk; }
                                if (lane == k) { pred = v; py = yp; }
                                if (lane == 0) {
                                    pivrow[K + k] = (int16_t)prow;
                                    pivcol[K + k] = c0 + col;
                                    rowpiv[prow] = (int16_t)(K + k);
                                }
                                ++k;
                            }
                        }
                        reinterpret_cast<uint2*>(G)[lane] = make_uint2(y0, y1);
                        if (lane < 8) PY[lane] = (uint8_t)py;
                        if (lane == 0) {
                            misc[0] = k;
                            if (k > 0) blk[nblk] = (uint32_t)K | ((uint32_t)k << 16);
                        }
                    }
                    if (LOOKAHEAD && pending) {
                        table_reads();
                        pending = false;
                    }
                    __syncthreads();
                    const int k = misc[0];
                    if (k > 0) {
                        if (!LOOKAHEAD && pending) table_reads();
                        // (c) my combination byte: by strip byte, or my own entry if I am one of the pivots
                        uint32_t y = G[byte];
                        const int mypiv = rowpiv[tid];
                        if (mypiv >= K) y = PY[mypiv - K];
                        used = mypiv >= 0;
                        Yb[tid] = (uint8_t)y;
                        {   // bit-sliced copy for the replays
                            uint32_t mine = 0u;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const uint32_t bj = __ballot_sync(0xFFFFFFFFu, (y >> j) & 1u);
                                if (lane == j) mine = bj;
                            }
                            if (lane < k) Ys[(K + lane) * nw + warp] = mine;
                        }
                        publish(K, k, mypiv);
                        __syncthreads();
                        tabulate(k, true, cw);
                        __syncthreads();
                        cur ^= TW[y];
                        pending = true;
                        K += k;
                        ++nblk;
                    }
                }
                if (pending) table_reads();              // r[] current again before the next word / write-out
            }
            // ---- write the slab out in pivot order; rows without a pivot so far are zero here -----
            __syncthreads();
            if (wi < W32) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int pk = rowpiv[warp * 32 + i];
                    if (pk >= 0) dst[(size_t)pk * W32 + wi] = r[i];
                }
                for (int row = K + warp; row < m; row += nw) dst[(size_t)row * W32 + wi] = 0u;
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

bool gf2_m4r_supported(int m, int n) {
    if (m < 1 || m > 1024 || n < 1) return false;
    return m4r_smem_bytes(m, n) <= 224 * 1024;
}

cudaError_t launch_gf2_m4r(const uint64_t* in, int batch, int m, int n, uint64_t* out, int32_t* rank,
                           int32_t* pivots, cudaStream_t stream) {
    const size_t smem = m4r_smem_bytes(m, n);
    const int mpad = (m + 31) & ~31;
    constexpr bool lookahead = true;
    using Kern = void (*)(const uint32_t*, int, int, int, uint32_t*, int32_t*, int32_t*);
    Kern kern = mpad == 1024 ? (lookahead ? k_gf2_m4r<32, true> : k_gf2_m4r<32, false>)
                             : (lookahead ? k_gf2_m4r<0, true> : k_gf2_m4r<0, false>);
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    int dev = 0, sms = 0, per_sm = 1;
    if ((err = cudaGetDevice(&dev)) != cudaSuccess) return err;
    if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return err;
    if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, mpad, smem)) != cudaSuccess) return err;
    if (per_sm < 1) per_sm = 1;
    int grid = sms * per_sm;
    if (grid > batch) grid = batch;
    kern<<<grid, mpad, smem, stream>>>(reinterpret_cast<const uint32_t*>(in), batch, m, n,
                                       reinterpret_cast<uint32_t*>(out), rank, pivots);
    return cudaGetLastError();
}

}  // namespace qcss
