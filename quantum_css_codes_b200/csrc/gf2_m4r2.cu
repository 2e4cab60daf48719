// K4, third generation: batched GF(2) Gauss-Jordan for matrices with up to 1024 rows (any width),
// TWO matrices resident per SM.  Replaces the per-column Python loop of
// bin_matrix.reduced_row_echelon_form (bin_matrix.py:8-34).
//
// gf2_m4r.cu keeps a 1024 x 1024-bit slab in the registers of one 1024-thread CTA; every strip then
// waits for the one warp that factors the panel, and nothing else can run on the SM meanwhile (ncu:
// 24 % of the stall samples sit behind that barrier).  Here a CTA has 512 threads and walks its matrix
// in slabs of 512 columns (16 words): lane l keeps word l & 15 of rows 64w + 32(l >> 4) + i in r[i],
// so a warp instruction still updates 32 words (two rows).  With 64 registers per thread two CTAs --
// two independent matrices -- share an SM, and the panel / barrier latency of one is filled by the
// table reads of the other.
//
//   panel     byte space, one warp, as in gf2_m4r.cu: a row enters an 8-column strip only through its
//             strip byte, so the 1024 x 8 panel reduces to a 256-entry problem (lane l owns the byte
//             values 8l..8l+7, four per register), seeded by one representative unused row per value
//             (REP).  Output: G[byte] = combination y of the strip-start pivot rows for a row with
//             that strip byte, and PY[u] for the pivot rows themselves.
//   apply     pivot rows are published from registers, all 2^k combinations are tabulated (TP, 32
//             words per entry = the 16 slab words twice, so the two half-warps read different banks),
//             and a row update is ONE table read: r[i] ^= TP[y_i]; address = one shift + one LOP3.
//   replay    the combination bytes of every block (1 byte per row) go to an L2-resident scratch
//             buffer, and the slabs to the right stream them back with cp.async (three-deep ring),
//             one block ahead of the table reads.
//
// Rows leave the CTA in pivot order (row holding pivot k -> output row k), zero rows last: the
// canonical RREF the reference returns.  Any unused row with a 1 may serve as pivot because the RREF
// is unique.
#include <cuda_runtime.h>

#include <cstdlib>

#include "launch.h"

namespace qcss {

namespace {

constexpr int kSW = 16;                             // slab width in 32-bit words

// Fixed shared-memory map (bytes); every offset is a compile-time constant.
constexpr int oTP = 0;                              // uint32 [256][32]  combination table
constexpr int oTW = oTP + 256 * 32 * 4;             // uint32 [256]      its pivot-word column
constexpr int oP = oTW + 256 * 4;                   // uint32 [8][16]    published pivot rows
constexpr int oG = oP + 8 * kSW * 4;                // uint8  [256]      strip byte -> y
constexpr int oPY = oG + 256;                       // uint8  [16]       y of the strip's pivot rows
constexpr int oMisc = oPY + 16;                     // int    [4]
constexpr int oRep = oMisc + 16;                    // uint16 [2][256]   representative row | 0x8000
constexpr int oS32 = oRep + 2 * 256 * 2;            // uint32 [1024]     word transpose
constexpr int oYb = oS32 + 1024 * 4;                // uint8  [3][1024]  combination bytes (ring)
constexpr int oRowpiv = oYb + 3 * 1024;             // int16  [1024]
constexpr int oPivrow = oRowpiv + 1024 * 2;         // int16  [1024]
constexpr int oPivcol = oPivrow + 1024 * 2;         // int32  [1024]
constexpr int oBlk = oPivcol + 1024 * 4;            // uint32 [1024]     (K | k << 16) per block
constexpr int kSmemBytes = oBlk + 1024 * 4;

__device__ __forceinline__ uint32_t pick_reg(const uint32_t (&r)[32], int idx) {
    uint32_t v;
    switch (idx) {                                  // idx is warp-uniform: one indirect branch
#define QCSS_PICK(i) case i: v = r[i]; break;
        QCSS_PICK(0) QCSS_PICK(1) QCSS_PICK(2) QCSS_PICK(3) QCSS_PICK(4) QCSS_PICK(5) QCSS_PICK(6) QCSS_PICK(7)
        QCSS_PICK(8) QCSS_PICK(9) QCSS_PICK(10) QCSS_PICK(11) QCSS_PICK(12) QCSS_PICK(13) QCSS_PICK(14)
        QCSS_PICK(15) QCSS_PICK(16) QCSS_PICK(17) QCSS_PICK(18) QCSS_PICK(19) QCSS_PICK(20) QCSS_PICK(21)
        QCSS_PICK(22) QCSS_PICK(23) QCSS_PICK(24) QCSS_PICK(25) QCSS_PICK(26) QCSS_PICK(27) QCSS_PICK(28)
        QCSS_PICK(29) QCSS_PICK(30)
#undef QCSS_PICK
        default: v = r[31]; break;
    }
    return v;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(512, 2)
k_gf2_m4r2(const uint32_t* __restrict__ in, int batch, int m, int n, uint32_t* __restrict__ out,
           int32_t* __restrict__ rank_out, int32_t* __restrict__ piv_out, uint8_t* __restrict__ yscratch,
           int cap_blocks) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* const TP = reinterpret_cast<uint32_t*>(smem + oTP);
    uint32_t* const TW = reinterpret_cast<uint32_t*>(smem + oTW);
    uint32_t* const P = reinterpret_cast<uint32_t*>(smem + oP);
    uint8_t* const G = smem + oG;
    uint8_t* const PY = smem + oPY;
    volatile int* const misc = reinterpret_cast<volatile int*>(smem + oMisc);
    uint16_t* const REP = reinterpret_cast<uint16_t*>(smem + oRep);
    uint32_t* const S32 = reinterpret_cast<uint32_t*>(smem + oS32);
    uint8_t* const Yb = smem + oYb;
    int16_t* const rowpiv = reinterpret_cast<int16_t*>(smem + oRowpiv);
    int16_t* const pivrow = reinterpret_cast<int16_t*>(smem + oPivrow);
    int32_t* const pivcol = reinterpret_cast<int32_t*>(smem + oPivcol);
    uint32_t* const blk = reinterpret_cast<uint32_t*>(smem + oBlk);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wl = lane & 15, h = lane >> 4;
    const int nw = (int)(blockDim.x >> 5);
    const int mrows = nw * 64;                           // rows incl. padding
    const int W32 = ((n + 63) >> 6) * 2;                 // 32-bit words per packed row
    const int nslabs = (W32 + kSW - 1) / kSW;
    const int npiv = m < n ? m : n;
    const uint32_t lane4 = (uint32_t)lane * 4u;
    const int ra = warp * 64 + lane, rb = ra + 32;       // the two rows whose scalars this thread tracks
    const int row0 = warp * 64 + h * 32;                 // r[i] belongs to row row0 + i
    uint8_t* const Yg = yscratch + (size_t)blockIdx.x * cap_blocks * mrows;

    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        const uint32_t* src = in + (size_t)b * m * W32;
        uint32_t* dst = out + (size_t)b * m * W32;
        rowpiv[ra] = -1;
        rowpiv[rb] = -1;
        for (int i = tid; i < 256; i += blockDim.x) reinterpret_cast<uint32_t*>(REP)[i] = 0u;   // both buffers
        bool used_a = ra >= m, used_b = rb >= m;         // padding rows never become pivots
        int K = 0, nblk = 0, strip_no = 0;
        __syncthreads();

        uint32_t r[32];

        // publish: the owners of the block's pivot rows write their slab words to P
        auto publish = [&](int Kb, int k) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (u < k) {
                    const int p = pivrow[Kb + u];
                    if ((p >> 6) == warp) {
                        const uint32_t v = pick_reg(r, p & 31);
                        if (((p >> 5) & 1) == h) P[u * kSW + wl] = v;
                    }
                }
            }
        };
        // tabulate: all combinations of the k published rows (entries beyond 2^k are never read)
        auto tabulate = [&](int k, bool piv, int cw) {
            uint32_t pu[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) pu[u] = (u < k) ? P[u * kSW + wl] : 0u;
            const int entries = 1 << k;
            for (int e0 = warp * 8; e0 < entries; e0 += nw * 8) {
                uint32_t base = 0u;
#pragma unroll
                for (int u = 3; u < 8; ++u)
                    if ((e0 >> u) & 1) base ^= pu[u];
                const uint32_t c1 = base ^ pu[0], c2 = base ^ pu[1], c3 = c1 ^ pu[1];
                const uint32_t c4 = base ^ pu[2], c5 = c1 ^ pu[2], c6 = c2 ^ pu[2], c7 = c3 ^ pu[2];
                uint32_t* t = TP + e0 * 32 + lane;
                t[0 * 32] = base; t[1 * 32] = c1; t[2 * 32] = c2; t[3 * 32] = c3;
                t[4 * 32] = c4;   t[5 * 32] = c5; t[6 * 32] = c6; t[7 * 32] = c7;
                if (piv && lane == cw) {
                    uint4* w4 = reinterpret_cast<uint4*>(TW + e0);
                    w4[0] = make_uint4(base, c1, c2, c3);
                    w4[1] = make_uint4(c4, c5, c6, c7);
                }
            }
        };
        // table_reads: one read per row; the 32 combination bytes of my half-warp's rows are 8 words
        auto table_reads = [&](const uint8_t* ybase) {
            const uint4* yw = reinterpret_cast<const uint4*>(ybase + row0);
            const uint4 ya = yw[0], yb = yw[1];
            const uint32_t yy[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
            const uint8_t* tp = smem + oTP;
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
        // prefetch: combination bytes of recorded block bi -> ring slot bi % 3
        auto prefetch = [&](int bi) {
            if (tid * 16 < mrows) cp_async16(Yb + (bi % 3) * 1024 + tid * 16, Yg + (size_t)bi * mrows + tid * 16);
        };

        for (int slab = 0; slab < nslabs; ++slab) {
            const int wi = slab * kSW + wl;
            // ---- load the slab into registers (columns >= n masked off) --------------------------
            uint32_t colmask = 0u;
            if (wi < W32) {
                const int c_lo = wi * 32;
                colmask = (c_lo + 32 <= n) ? 0xFFFFFFFFu : (c_lo < n ? ((1u << (n - c_lo)) - 1u) : 0u);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int row = row0 + i;
                r[i] = (row < m && colmask != 0u) ? (__ldg(src + (size_t)row * W32 + wi) & colmask) : 0u;
            }
            // ---- replay every block found in earlier slabs ----------------------------------------
            if (nblk > 0) {
                prefetch(0);
                cp_async_commit();
                for (int bi = 0; bi < nblk; ++bi) {
                    if (bi + 1 < nblk) prefetch(bi + 1);
                    cp_async_commit();
                    const uint32_t e = blk[bi];
                    const int k = (int)(e >> 16);
                    publish((int)(e & 0xFFFFu), k);
                    cp_async_wait<1>();                  // block bi has landed (bi + 1 may be in flight)
                    __syncthreads();
                    tabulate(k, false, 0);
                    __syncthreads();
                    table_reads(Yb + (bi % 3) * 1024);
                }
                cp_async_wait<0>();
            }
            // ---- discovery: strips of 8 columns of this slab --------------------------------------
            const int slab_words = (W32 - slab * kSW) < kSW ? (W32 - slab * kSW) : kSW;
            for (int cw = 0; cw < slab_words && K < m; ++cw) {
                if ((slab * kSW + cw) * 32 >= n) break;
                // my rows' current word cw, out of the registers of the two lanes that hold it
                if (wl == cw) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) S32[row0 + i] = r[i];
                }
                __syncwarp();
                uint32_t cur_a = S32[ra], cur_b = S32[rb];
                for (int sb = 0; sb < 4 && K < m; ++sb) {
                    const int c0 = (slab * kSW + cw) * 32 + sb * 8;
                    if (c0 >= n) break;
                    // (a) every unused row offers itself as the representative of its strip byte
                    const uint32_t byte_a = (cur_a >> (8 * sb)) & 0xFFu, byte_b = (cur_b >> (8 * sb)) & 0xFFu;
                    uint16_t* rep = REP + (strip_no & 1) * 256;
                    for (int i = tid; i < 128; i += blockDim.x)
                        reinterpret_cast<uint32_t*>(REP + ((strip_no + 1) & 1) * 256)[i] = 0u;
                    if (!used_a) rep[byte_a] = (uint16_t)(ra | 0x8000);
                    if (!used_b) rep[byte_b] = (uint16_t)(rb | 0x8000);
                    ++strip_no;
                    __syncthreads();
                    // (b) warp 0 factors the panel in byte space
                    if (warp == 0) {
                        const uint4 q = reinterpret_cast<const uint4*>(rep)[lane];   // values 8l..8l+7
                        const uint32_t prs0 = (__byte_perm(q.x, q.y, 0x7531u) >> 7) & 0x01010101u;
                        const uint32_t prs1 = (__byte_perm(q.z, q.w, 0x7531u) >> 7) & 0x01010101u;
                        uint32_t red0 = 0x03020100u + 0x08080808u * (uint32_t)lane, red1 = red0 + 0x04040404u;
                        uint32_t y0 = 0u, y1 = 0u;
                        uint32_t pred = 0u, py = 0u, myval = 0u, mycol = 0u;   // lane u < k: pivot u
                        int k = 0;
#pragma unroll
                        for (int col = 0; col < 8; ++col) {
                            const uint32_t s0 = red0 >> col, s1 = red1 >> col;
                            const uint32_t cand0 = s0 & prs0, cand1 = s1 & prs1;
                            const unsigned vote = __ballot_sync(0xFFFFFFFFu, (cand0 | cand1) != 0u);
                            if (vote != 0u) {
                                const int srcl = __ffs(vote) - 1;
                                // my first candidate value e (0..7), its reduced byte and y
                                const uint32_t e = cand0 ? (uint32_t)(__ffs(cand0) - 1) >> 3
                                                         : 4u + ((uint32_t)(__ffs(cand1) - 1) >> 3);
                                uint32_t pack = (__byte_perm(red0, red1, e) & 0xFFu) |
                                                ((__byte_perm(y0, y1, e) & 0xFFu) << 8) | (e << 16);
                                pack = __shfl_sync(0xFFFFFFFFu, pack, srcl);
                                const uint32_t v = pack & 0xFFu, yp = (pack >> 8) & 0xFFu;
                                const uint32_t yk = yp | (1u << k);
                                const uint32_t v4 = v * 0x01010101u, yk4 = yk * 0x01010101u;
                                const uint32_t M0 = (s0 & 0x01010101u) * 0xFFu, M1 = (s1 & 0x01010101u) * 0xFFu;
                                red0 ^= M0 & v4;  red1 ^= M1 & v4;
                                y0 ^= M0 & yk4;   y1 ^= M1 & yk4;
                                if (lane < k && ((pred >> col) & 1u)) { pred ^= v; py ^= yk; }
                                if (lane == k) {
                                    pred = v; py = yp;
                                    myval = (uint32_t)srcl * 8u + (pack >> 16);
                                    mycol = (uint32_t)col;
                                }
                                ++k;
                            }
                        }
                        reinterpret_cast<uint2*>(G)[lane] = make_uint2(y0, y1);
                        if (lane < k) {
                            const int prow = rep[myval] & 0x3FF;
                            pivrow[K + lane] = (int16_t)prow;
                            pivcol[K + lane] = c0 + (int)mycol;
                            rowpiv[prow] = (int16_t)(K + lane);
                            PY[lane] = (uint8_t)py;
                        }
                        if (lane == 0) {
                            misc[0] = k;
                            if (k > 0) blk[nblk] = (uint32_t)K | ((uint32_t)k << 16);
                        }
                    }
                    __syncthreads();
                    const int k = misc[0];
                    if (k > 0) {
                        // (c) my rows' combination bytes: by strip byte, or the pivot's own entry
                        uint32_t ya = G[byte_a], yb = G[byte_b];
                        const int pa = rowpiv[ra], pb = rowpiv[rb];
                        if (pa >= K) ya = PY[pa - K];
                        if (pb >= K) yb = PY[pb - K];
                        used_a = pa >= 0 || ra >= m;
                        used_b = pb >= 0 || rb >= m;
                        Yb[ra] = (uint8_t)ya;
                        Yb[rb] = (uint8_t)yb;
                        Yg[(size_t)nblk * mrows + ra] = (uint8_t)ya;     // for the replays
                        Yg[(size_t)nblk * mrows + rb] = (uint8_t)yb;
                        publish(K, k);
                        __syncthreads();
                        tabulate(k, true, cw);
                        __syncthreads();
                        table_reads(Yb);
                        cur_a ^= TW[ya];
                        cur_b ^= TW[yb];
                        K += k;
                        ++nblk;
                    }
                }
            }
            // ---- write the slab out in pivot order; rows without a pivot so far are zero here -----
            __syncthreads();
            if (wi < W32) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int pk = rowpiv[row0 + i];
                    if (pk >= 0) dst[(size_t)pk * W32 + wi] = r[i];
                }
                for (int row = K + warp * 2 + h; row < m; row += nw * 2) dst[(size_t)row * W32 + wi] = 0u;
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

bool gf2_m4r2_supported(int m, int n) { return m >= 1 && m <= 1024 && n >= 1; }

cudaError_t launch_gf2_m4r2(const uint64_t* in, int batch, int m, int n, uint64_t* out, int32_t* rank,
                            int32_t* pivots, cudaStream_t stream) {
    const int nw = (m + 63) / 64;
    const int threads = nw * 32;
    cudaError_t err = cudaFuncSetAttribute(k_gf2_m4r2, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (err != cudaSuccess) return err;
    int dev = 0, sms = 0, per_sm = 1;
    if ((err = cudaGetDevice(&dev)) != cudaSuccess) return err;
    if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return err;
    if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gf2_m4r2, threads, kSmemBytes)) != cudaSuccess)
        return err;
    if (per_sm < 1) per_sm = 1;
    int grid = sms * per_sm;
    if (grid > batch) grid = batch;
    // one combination byte per row per block, at most one block per pivot and per 8-column strip
    const int kmax = m < n ? m : n;
    const int strips = (n + 7) / 8;
    const int cap_blocks = kmax < strips ? kmax : strips;
    const size_t scratch = (size_t)grid * cap_blocks * (size_t)(nw * 64);
    uint8_t* d_scratch = nullptr;
    if ((err = cudaMallocAsync(reinterpret_cast<void**>(&d_scratch), scratch, stream)) != cudaSuccess) return err;
    k_gf2_m4r2<<<grid, threads, kSmemBytes, stream>>>(reinterpret_cast<const uint32_t*>(in), batch, m, n,
                                                     reinterpret_cast<uint32_t*>(out), rank, pivots, d_scratch,
                                                     cap_blocks);
    err = cudaGetLastError();
    const cudaError_t ferr = cudaFreeAsync(d_scratch, stream);
    return err != cudaSuccess ? err : ferr;
}

}  // namespace qcss
