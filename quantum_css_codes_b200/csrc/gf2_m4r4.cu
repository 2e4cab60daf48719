// K4, fourth generation: batched GF(2) Gauss-Jordan for matrices with up to 1024 rows (any width), one
// matrix per SM.  Replaces the per-column Python loop of bin_matrix.reduced_row_echelon_form
// (bin_matrix.py:8-34).
//
// gf2_m4r2.cu (two 512-thread CTAs per SM, 512-column slabs) spends ~1350 shared-memory wavefronts per block
// application of 512 columns: 64-byte table entries collide pairwise in the banks unless the table is stored twice
// (256 extra store wavefronts), the 128-bit pivot-row loads of the tabulation cost four wavefronts each (~450), and a
// 2048-column matrix needs 448 block applications.  Here the slab is 1024 columns:
//
//   layout    1024 threads; lane l = (g = l >> 3, j = l & 7) keeps words 4j .. 4j+3 of rows 32 warp + 8 g + i in
//             r[i] (uint4, i < 8).  A table entry is 128 bytes = all 32 banks, so the eight lanes of one
//             quarter-warp phase of an LDS.128 read ONE entry: conflict-free by construction, no second copy.
//   tabulate  lane = word: pivot rows are read with 32-bit loads (one wavefront per pivot row per warp) and every
//             warp writes eight entries with 128-byte-per-warp stores: ~430 wavefronts per block.
//   apply     r[i] ^= TP[y_i]: 1024 wavefronts per block of 1024 columns; 256 block applications for 1024 x 2048.
//   replay    the combination bytes of every block (1 byte per row) STAY IN SHARED MEMORY (one CTA per SM leaves
//             ~175 KB: 175 blocks of 1024 rows); only blocks beyond that spill to an L2-resident scratch.
//   panel     byte space, one warp, as in gf2_m4r.cu / gf2_m4r2.cu (a row enters an 8-column strip only through its
//             strip byte; lane l owns byte values 8l .. 8l+7).  The block found in strip s is applied lazily under
//             the panel of strip s + 1, and the tracked word of every row is advanced from the pivot rows' TRACKED
//             words (PW, gathered by the panel warp out of the rows' mirror S32) instead of the published rows:
//             TWO block barriers per strip (offer | panel + apply | look up, publish, track), three at a word
//             boundary, and no flush between words.
//
// Rows leave the CTA in pivot order (row holding pivot k -> output row k), zero rows last: the canonical RREF the
// reference returns.  Any unused row with a 1 may serve as pivot because the RREF is unique.
#include <cuda_runtime.h>

#include <cstdlib>

#include "launch.h"

namespace qcss {

namespace {

constexpr int kSW = 32;                             // slab width in 32-bit words

// Phase profile (tools/gf2_m4r4_probe.cu compiles this file with -DQCSS_M4R4_PROF): cycles of warps 0 and 1 of CTA 0.
#ifdef QCSS_M4R4_PROF
__device__ unsigned long long g_m4r4_prof[64];
#define PROF_DECL long long prof_last = clock64(); const bool prof_on = blockIdx.x == 0 && lane == 0 && warp <= 1;
#define PROF(i)                                                                                         \
    do {                                                                                                \
        if (prof_on) {                                                                                  \
            const long long prof_now = clock64();                                                       \
            atomicAdd(&g_m4r4_prof[warp * 32 + (i)], (unsigned long long)(prof_now - prof_last));       \
            prof_last = prof_now;                                                                       \
        }                                                                                               \
    } while (0)
#else
#define PROF_DECL
#define PROF(i) do {} while (0)
#endif

// Fixed shared-memory map (bytes); the replay bytes (Ys) come last and take what is left.
constexpr int oTP = 0;                              // uint32 [256][32]  combination table
constexpr int oP = oTP + 256 * kSW * 4;             // uint32 [8][32]    published pivot rows
constexpr int oG = oP + 8 * kSW * 4;                // uint8  [256]      strip byte -> y
constexpr int oPY = oG + 256;                       // uint8  [16]       y of the strip's pivot rows
constexpr int oPW = oPY + 16;                       // uint32 [8]        tracked word of the strip's pivot rows
constexpr int oMisc = oPW + 32;                     // int    [4]
constexpr int oRep = oMisc + 16;                    // uint16 [2][256]   representative row | 0x8000
constexpr int oS32 = oRep + 2 * 256 * 2;            // uint32 [1024]     tracked word of every row
constexpr int oRowpiv = oS32 + 1024 * 4;            // int16  [1024]
constexpr int oBlk = oRowpiv + 1024 * 2;            // uint32 [1024]     (K | k << 16) per block
constexpr int oYs = (oBlk + 1024 * 4 + 127) & ~127; // uint8  [cap_smem][mrows]
constexpr int kMaxDynSmem = 227 * 1024;

__device__ __forceinline__ uint4 pick4(const uint4 (&r)[8], int idx) {
    switch (idx) {                                  // idx is warp-uniform
        case 0: return r[0];
        case 1: return r[1];
        case 2: return r[2];
        case 3: return r[3];
        case 4: return r[4];
        case 5: return r[5];
        case 6: return r[6];
        default: return r[7];
    }
}
__device__ __forceinline__ uint32_t comp4(const uint4& v, int c) {
    return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w));
}
// PTX prmt: only selector nibbles whose output byte is used need to be valid
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// integer multiplies as PTX so that they stay on the FMA pipe (a shift would be free to move to the ALU pipe)
__device__ __forceinline__ uint32_t mul_lo(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t mul_hi(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t mad_hi(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ void xor4(uint4& a, const uint4& b) { a.x ^= b.x; a.y ^= b.y; a.z ^= b.z; a.w ^= b.w; }

// XOR of the words pw[u] selected by the bits of y
__device__ __forceinline__ uint32_t comb8(uint32_t y, const uint4& lo, const uint4& hi) {
    uint32_t t = lo.x & (0u - (y & 1u));
    t ^= lo.y & (0u - ((y >> 1) & 1u));
    t ^= lo.z & (0u - ((y >> 2) & 1u));
    t ^= lo.w & (0u - ((y >> 3) & 1u));
    t ^= hi.x & (0u - ((y >> 4) & 1u));
    t ^= hi.y & (0u - ((y >> 5) & 1u));
    t ^= hi.z & (0u - ((y >> 6) & 1u));
    t ^= hi.w & (0u - ((y >> 7) & 1u));
    return t;
}

__global__ void __launch_bounds__(1024, 1)
k_gf2_m4r4(const uint32_t* __restrict__ in, int batch, int m, int n, uint32_t* __restrict__ out,
           int32_t* __restrict__ rank_out, int32_t* __restrict__ piv_out, uint8_t* __restrict__ yscratch,
           int cap_smem, int cap_spill) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* const TP = reinterpret_cast<uint32_t*>(smem + oTP);
    uint32_t* const P = reinterpret_cast<uint32_t*>(smem + oP);
    uint8_t* const G = smem + oG;
    uint8_t* const PY = smem + oPY;
    uint32_t* const PW = reinterpret_cast<uint32_t*>(smem + oPW);
    volatile int* const misc = reinterpret_cast<volatile int*>(smem + oMisc);
    uint16_t* const REP = reinterpret_cast<uint16_t*>(smem + oRep);
    uint32_t* const S32 = reinterpret_cast<uint32_t*>(smem + oS32);
    int16_t* const rowpiv = reinterpret_cast<int16_t*>(smem + oRowpiv);
    uint32_t* const blk = reinterpret_cast<uint32_t*>(smem + oBlk);
    uint8_t* const Ys = smem + oYs;

    int tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 3, j = lane & 7;               // row group within the warp, word quad in the slab
    const int nw = (int)(blockDim.x >> 5);
    const int nthreads = nw * 32;
    const int mrows = nthreads;                          // rows incl. padding; thread tid tracks row tid
    const int W32 = ((n + 63) >> 6) * 2;                 // 32-bit words per packed row
    const int nslabs = (W32 + kSW - 1) / kSW;
    const int npiv = m < n ? m : n;
    const int row0 = (warp << 5) + (g << 3);             // r[i] belongs to row row0 + i
    const uint32_t laneoff = (uint32_t)j << 4;           // byte offset of my quad inside a table entry
    PROF_DECL

    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        rowpiv[tid] = -1;
        for (int i = tid; i < 256; i += nthreads) reinterpret_cast<uint32_t*>(REP)[i] = 0u;   // both buffers
        bool used = tid >= m;                            // padding rows never become pivots
        int mypiv = -1;                                  // pivot index of row tid
        int K = 0, nblk = 0, strip_no = 0;
        __syncthreads();

        uint4 r[8];

        // combination bytes of block bi for row `row`: shared memory, or the spill area behind it
        auto spill_ptr = [&](int bi) -> uint8_t* {
            return yscratch + ((size_t)blockIdx.x * cap_spill + (size_t)(bi - cap_smem)) * mrows;
        };
        auto store_y = [&](int bi, uint32_t y) {
            if (bi < cap_smem) Ys[bi * mrows + tid] = (uint8_t)y;
            else spill_ptr(bi)[tid] = (uint8_t)y;
        };
        auto load_y8 = [&](int bi) -> uint2 {
            if (bi < cap_smem) return *reinterpret_cast<const uint2*>(Ys + bi * mrows + row0);
            return __ldcg(reinterpret_cast<const uint2*>(spill_ptr(bi) + row0));
        };
        // publish: the rows that are pivots [Kb, Kb + k) write their slab words to P
        auto publish = [&](int Kb, int k) {
            unsigned mine = __ballot_sync(0xFFFFFFFFu, (unsigned)(mypiv - Kb) < (unsigned)k);
            while (mine != 0u) {
                const int i = __ffs(mine) - 1;
                mine &= mine - 1u;
                const int u = __shfl_sync(0xFFFFFFFFu, mypiv, i) - Kb;
                const uint4 v = pick4(r, i & 7);
                if (g == (i >> 3)) *reinterpret_cast<uint4*>(P + u * kSW + j * 4) = v;
            }
        };
        // tabulate: all combinations of the k published rows (entries beyond 2^k are never read, so stale rows
        // u >= k of P are harmless).  lane = word; a warp pass writes eight entries.
        // The warps w0 = 0 .. wn-1 of the tabulating set share the entries.
        auto tabulate = [&](int k, int w0, int wn) {
            const int entries = 1 << k;
            for (int E = w0 * 8; E < entries; E += wn * 8) {
                uint32_t base = 0u;
#pragma unroll
                for (int u = 3; u < 8; ++u)
                    if ((E >> u) & 1) base ^= P[u * kSW + lane];     // warp-uniform
                const uint32_t p0 = P[lane], p1 = P[kSW + lane], p2 = P[2 * kSW + lane];
                const uint32_t c1 = base ^ p0, c2 = base ^ p1, c3 = c1 ^ p1;
                uint32_t* t = TP + E * kSW + lane;
                t[0 * kSW] = base; t[1 * kSW] = c1; t[2 * kSW] = c2; t[3 * kSW] = c3;
                t[4 * kSW] = base ^ p2; t[5 * kSW] = c1 ^ p2; t[6 * kSW] = c2 ^ p2; t[7 * kSW] = c3 ^ p2;
            }
        };
        // table_reads: one 128-bit read per row; the combination bytes of my group's 8 rows are 2 words
        auto table_reads = [&](int bi) {
            const uint2 yv = load_y8(bi);
            const uint8_t* tp = smem + oTP;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const uint32_t w = q ? yv.y : yv.x;
#ifdef QCSS_M4R4_ALU_ADDR
                const uint32_t a0 = ((w << 7) & 0x7F80u) | laneoff;
                const uint32_t a1 = ((w >> 1) & 0x7F80u) | laneoff;
                const uint32_t a2 = ((w >> 9) & 0x7F80u) | laneoff;
                const uint32_t a3 = ((w >> 17) & 0x7F80u) | laneoff;
#else
                // entry address = 128 byte + 16 j, computed on the FMA pipe (integer multiply-add, high and low halves):
                // the ALU pipe (one warp instruction per two cycles per scheduler) keeps the 32 XORs, and the panel warp's
                // ALU-only chain waits less behind its scheduler-mates
                // (ptxas turns the last multiply-add into LEA.HI, ALU pipe: one ALU instruction per row instead of two; hiding
                //  the constant from it keeps everything on the FMA pipe but measured 4 % slower)
                const uint32_t a0 = mad_hi(mul_lo(w, 1u << 24), 1u << 15, laneoff);
                const uint32_t a1 = mad_hi(mul_lo(mul_hi(w, 1u << 24), 1u << 24), 1u << 15, laneoff);
                const uint32_t a2 = mad_hi(mul_lo(mul_hi(w, 1u << 16), 1u << 24), 1u << 15, laneoff);
                const uint32_t a3 = mad_hi(mul_lo(mul_hi(w, 1u << 8), 1u << 24), 1u << 15, laneoff);
#endif
                xor4(r[4 * q + 0], *reinterpret_cast<const uint4*>(tp + a0));
                xor4(r[4 * q + 1], *reinterpret_cast<const uint4*>(tp + a1));
                xor4(r[4 * q + 2], *reinterpret_cast<const uint4*>(tp + a2));
                xor4(r[4 * q + 3], *reinterpret_cast<const uint4*>(tp + a3));
            }
        };

        for (int slab = 0; slab < nslabs; ++slab) {
            const int wi = slab * kSW + j * 4;           // first of my four words
            // ---- load the slab into registers (columns >= n masked off) --------------------------
            uint32_t cm[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int c_lo = (wi + c) * 32;
                cm[c] = (wi + c >= W32) ? 0u : ((c_lo + 32 <= n) ? 0xFFFFFFFFu : (c_lo < n ? ((1u << (n - c_lo)) - 1u) : 0u));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = row0 + i;
                uint2 lo = make_uint2(0u, 0u), hi = make_uint2(0u, 0u);
                if (row < m) {
                    const uint32_t* ptr = in + ((size_t)b * m + row) * W32 + wi;
                    if (wi < W32) lo = __ldg(reinterpret_cast<const uint2*>(ptr));
                    if (wi + 2 < W32) hi = __ldg(reinterpret_cast<const uint2*>(ptr + 2));
                }
                r[i] = make_uint4(lo.x & cm[0], lo.y & cm[1], hi.x & cm[2], hi.y & cm[3]);
            }
            PROF(14);
            // ---- replay every block found in earlier slabs ----------------------------------------
            for (int bi = 0; bi < nblk; ++bi) {
                const uint32_t e = blk[bi];
                const int k = (int)(e >> 16);
                publish((int)(e & 0xFFFFu), k);
                PROF(0);
                __syncthreads();                         // P complete; every warp is done with the previous table
                PROF(1);
                tabulate(k, warp, nw);
                PROF(2);
                __syncthreads();
                PROF(3);
                table_reads(bi);
                PROF(4);
            }
            // ---- discovery: strips of 8 columns of this slab --------------------------------------
            // pending: block nblk - 1 is published (P) and recorded (combination bytes) but not yet applied to r[];
            // it is tabulated and read while warp 0 factors the next panel.  Uniform over the CTA.
            const int slab_words = (W32 - slab * kSW) < kSW ? (W32 - slab * kSW) : kSW;
            bool pending = false;
            int pend_k = 0;
            uint32_t y_last = 0u;                        // my row's combination byte in the pending block
            for (int cw = 0; cw < slab_words && K < m; ++cw) {
                if ((slab * kSW + cw) * 32 >= n) break;
                // my row's word cw: out of the registers of the lanes that hold it (current but for the pending
                // block), then the pending block's contribution from its published pivot rows
                if (j == (cw >> 2)) {
                    const int c = cw & 3;
#pragma unroll
                    for (int i = 0; i < 8; ++i) S32[row0 + i] = comp4(r[i], c);
                }
                if (pending) __syncthreads(); else __syncwarp();       // P was published after the last barrier
                uint32_t cur = S32[tid];
                if (pending) {
                    uint32_t t = 0u;
#pragma unroll
                    for (int u = 0; u < 8; ++u) t ^= P[u * kSW + cw] & (0u - ((y_last >> u) & 1u));
                    cur ^= t;
                }
                PROF(5);
                for (int sb = 0; sb < 4 && K < m; ++sb) {
                    const int c0 = (slab * kSW + cw) * 32 + sb * 8;
                    if (c0 >= n) break;
                    // (X) every unused row offers itself as the representative of its strip byte and mirrors its
                    //     tracked word
                    const uint32_t byte = (cur >> (8 * sb)) & 0xFFu;
                    uint16_t* rep = REP + (strip_no & 1) * 256;
                    for (int i = tid; i < 128; i += nthreads)
                        reinterpret_cast<uint32_t*>(REP + ((strip_no + 1) & 1) * 256)[i] = 0u;
                    if (!used) rep[byte] = (uint16_t)(tid | 0x8000);
                    S32[tid] = cur;
                    ++strip_no;
                    PROF(6);
                    __syncthreads();
                    PROF(7);
                    // (Y) warp 0 factors the panel in byte space; the other warps apply the pending block meanwhile
                    if (warp == 0) {
                        // Byte space: lane l owns the byte values 8l .. 8l+7 (value 8l+e in byte e & 3 of red{e >> 2}), their
                        // reduced value (red) and combination byte (y).  A column is branch-free: every lane keys its
                        // last candidate (a present value whose reduced byte has the column's bit) as
                        // [reduced byte, y, row | 0x8000] and ONE warp max-reduction (REDUX, uniform datapath -- a
                        // shuffle would queue behind the other warps' table reads in the load/store pipe) elects the
                        // pivot and broadcasts all of it; a column without candidates reduces to 0 and multiplies its
                        // update masks by zero.  Dependent chain per column: shift, and, shift-add, find-leading-one,
                        // shift, byte-permute x3, select, reduce, byte-permute, xor (tools/experiments/panel_bench.cu:
                        // 1100 cycles per panel alone, 1500 under 31 warps of table reads; the branching ballot + shuffle
                        // form took 1450 / 2300).
                        const uint4 q = reinterpret_cast<const uint4*>(rep)[lane];   // rows of the values 8l..8l+7 | 0x8000
                        const uint32_t ql0 = prmt(q.x, q.y, 0x6420u), ql1 = prmt(q.z, q.w, 0x6420u);   // low bytes of the rows
                        const uint32_t qh0 = prmt(q.x, q.y, 0x7531u), qh1 = prmt(q.z, q.w, 0x7531u);   // high bytes | 0x80
                        const uint32_t prs0 = (qh0 >> 7) & 0x01010101u, prs1 = (qh1 >> 7) & 0x01010101u;
                        uint32_t red0 = 0x03020100u + 0x08080808u * (uint32_t)lane, red1 = red0 + 0x04040404u;
                        uint32_t y0 = 0u, y1 = 0u;
                        uint32_t myrow = 0u, mycol = 0u;       // lane u < k: pivot u
                        uint32_t kmask = 0x01010101u;          // bit k of every byte lane
                        uint32_t k = 0u;
#pragma unroll
                        for (int col = 0; col < 8; ++col) {
                            const uint32_t s0 = red0 >> col, s1 = red1 >> col;
                            // candidates: bit 8i = value i, bit 8i + 4 = value 4 + i
                            const uint32_t c = (s0 & prs0) + ((s1 & prs1) << 4);
                            const uint32_t pbit = 31u - (uint32_t)__clz((int)c);
                            const uint32_t sel = 0x73625140u >> (pbit & 28u);       // nibble 0: byte index i + 4h of bit 8i + 4h
                            const uint32_t key = prmt(prmt(prmt(ql0, ql1, sel), prmt(qh0, qh1, sel), 0x0040u),
                                                      prmt(prmt(y0, y1, sel), prmt(red0, red1, sel), 0x0040u), 0x5410u);
                            const uint32_t pack = __reduce_max_sync(0xFFFFFFFFu, c != 0u ? key : 0u);
                            const uint32_t found = pack != 0u ? 1u : 0u;
                            const uint32_t v4 = prmt(pack, 0u, 0x3333u);            // reduced byte of the pivot in every byte lane
                            const uint32_t yk4 = prmt(pack, 0u, 0x2222u) | kmask;
                            const uint32_t fm = found * 0xFFu;
                            const uint32_t M0 = (s0 & 0x01010101u) * fm, M1 = (s1 & 0x01010101u) * fm;
                            red0 ^= M0 & v4;  red1 ^= M1 & v4;
                            y0 ^= M0 & yk4;   y1 ^= M1 & yk4;
                            if (found != 0u && lane == (int)k) {
                                myrow = pack & 0x3FFu;
                                mycol = (uint32_t)col;
                            }
                            kmask <<= found;
                            k += found;
                        }
                        reinterpret_cast<uint2*>(G)[lane] = make_uint2(y0, y1);
                        __syncwarp();
                        if (lane < 8) {
                            uint32_t pw = 0u;
                            if (lane < (int)k) {
                                // A pivot row must keep a single 1 in its own column: its combination byte is that of
                                // (its strip byte ^ the unit byte of its column) = {itself} ^ G[unit], by linearity.
                                const uint32_t py = G[1u << mycol] ^ (1u << lane);
                                rowpiv[myrow] = (int16_t)(K + lane);
                                PY[lane] = (uint8_t)py;
                                pw = S32[myrow];
                                if (piv_out != nullptr) piv_out[(size_t)b * npiv + K + lane] = c0 + (int)mycol;
                            }
                            PW[lane] = pw;
                        }
                        if (lane == 0) {
                            misc[0] = (int)k;
                            if (k > 0u) blk[nblk] = (uint32_t)K | (k << 16);
                        }
                        PROF(9);
                        if (pending) {
                            // my own rows: the other warps have long since tabulated (they arrived on barrier 2 on the way)
                            asm volatile("bar.sync 2, %0;" ::"r"(nthreads) : "memory");
                            table_reads(nblk - 1);
                        }
                        PROF(10);
                    } else if (pending) {
                        // The warps that share warp 0's scheduler (warp & 3 == 0) sit the tabulation out: the panel's
                        // instructions wait for the ALU pipe behind theirs (ncu source page: math-pipe throttle and
                        // not-selected on every panel instruction; with the other warps idle the panel takes 1450 cycles
                        // in this kernel, under their tabulation and table reads 2500).  The other three quarters cover
                        // the 256 entries in at most two passes each, as 31 warps did.  -1 %.
                        if (warp & 3) tabulate(pend_k, warp - (warp >> 2) - 1, nw - ((nw + 3) >> 2));
                        PROF(8);
                        asm volatile("bar.arrive 2, %0;" ::"r"(nthreads) : "memory");
                        asm volatile("bar.sync 1, %0;" ::"r"(nthreads - 32) : "memory");   // warps 1 .. nw-1 have tabulated
                        PROF(9);
                        table_reads(nblk - 1);
                        PROF(10);
                    }
                    __syncthreads();
                    PROF(11);
                    // (Z) look up my combination byte, record it, publish the new pivot rows, advance the tracked word
                    PROF(12);
                    pending = false;
                    const int k = misc[0];
                    if (k > 0) {
                        uint32_t y = G[byte];
                        const int pa = rowpiv[tid];
                        if (pa >= K) {
                            y = PY[pa - K];
                            mypiv = pa;
                            used = true;
                        }
                        store_y(nblk, y);
                        publish(K, k);
                        const uint4 lo = *reinterpret_cast<const uint4*>(PW), hi = *reinterpret_cast<const uint4*>(PW + 4);
                        cur ^= comb8(y, lo, hi);
                        y_last = y;
                        pending = true;
                        pend_k = k;
                        K += k;
                        ++nblk;
                    }
                    PROF(13);
                }
            }
            // ---- apply the last block, write the slab out in pivot order; rows without a pivot so far are zero here
            __syncthreads();                             // warp 0's deferred table reads are done; bytes and P visible
            if (pending) {
                tabulate(pend_k, warp, nw);
                __syncthreads();
                table_reads(nblk - 1);
            }
            if (wi < W32) {
                const bool hi_ok = wi + 2 < W32;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int pk = rowpiv[row0 + i];
                    if (pk >= 0) {
                        uint32_t* ptr = out + ((size_t)b * m + pk) * W32 + wi;
                        *reinterpret_cast<uint2*>(ptr) = make_uint2(r[i].x, r[i].y);
                        if (hi_ok) *reinterpret_cast<uint2*>(ptr + 2) = make_uint2(r[i].z, r[i].w);
                    }
                }
                for (int row = K + warp * 4 + g; row < m; row += nw * 4) {
                    uint32_t* ptr = out + ((size_t)b * m + row) * W32 + wi;
                    *reinterpret_cast<uint2*>(ptr) = make_uint2(0u, 0u);
                    if (hi_ok) *reinterpret_cast<uint2*>(ptr + 2) = make_uint2(0u, 0u);
                }
            }
        }
        PROF(15);
        // ---- rank and the unused pivot slots ------------------------------------------------------
        if (tid == 0 && rank_out != nullptr) rank_out[b] = K;
        if (piv_out != nullptr)
            for (int t = K + tid; t < npiv; t += nthreads) piv_out[(size_t)b * npiv + t] = -1;
        __syncthreads();
    }
}

}  // namespace

// Warp 0 factors panels while the other warps tabulate: at least two warps.
bool gf2_m4r4_supported(int m, int n) { return m > 32 && m <= 1024 && n >= 1; }

cudaError_t launch_gf2_m4r4(const uint64_t* in, int batch, int m, int n, uint64_t* out, int32_t* rank,
                            int32_t* pivots, cudaStream_t stream) {
    const int nw = (m + 31) / 32;
    const int threads = nw * 32;
    // one combination byte per row per block, at most one block per pivot and per 8-column strip
    const int kmax = m < n ? m : n;
    const int strips = (n + 7) / 8;
    const int cap_blocks = kmax < strips ? kmax : strips;
    int cap_smem = (kMaxDynSmem - oYs) / threads;
    if (cap_smem > cap_blocks) cap_smem = cap_blocks;
    const int cap_spill = cap_blocks - cap_smem;
    const int smem_bytes = oYs + cap_smem * threads;
    cudaError_t err = cudaFuncSetAttribute(k_gf2_m4r4, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (err != cudaSuccess) return err;
    int dev = 0, sms = 0;
    if ((err = cudaGetDevice(&dev)) != cudaSuccess) return err;
    if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return err;
    int grid = sms;
    if (grid > batch) grid = batch;
    uint8_t* d_scratch = nullptr;
    const size_t scratch = (size_t)grid * cap_spill * (size_t)threads;
    if (scratch > 0) {
        static bool pool_set[64] = {};                   // keep freed scratch in the stream-ordered pool
        if (dev >= 0 && dev < 64 && !pool_set[dev]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                unsigned long long keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            pool_set[dev] = true;
        }
        if ((err = cudaMallocAsync(reinterpret_cast<void**>(&d_scratch), scratch, stream)) != cudaSuccess) return err;
    }
    k_gf2_m4r4<<<grid, threads, smem_bytes, stream>>>(reinterpret_cast<const uint32_t*>(in), batch, m, n,
                                                     reinterpret_cast<uint32_t*>(out), rank, pivots, d_scratch,
                                                     cap_smem, cap_spill);
    err = cudaGetLastError();
    if (d_scratch != nullptr) {
        const cudaError_t ferr = cudaFreeAsync(d_scratch, stream);
        if (err == cudaSuccess) err = ferr;
    }
    return err;
}

}  // namespace qcss
