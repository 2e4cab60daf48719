// K4, third generation: batched GF(2) Gauss-Jordan for matrices with up to 1024 rows (any width),
// TWO matrices resident per SM.  Replaces the per-column Python loop of
// bin_matrix.reduced_row_echelon_form (bin_matrix.py:8-34).
//
// gf2_m4r.cu keeps a 1024 x 1024-bit slab in the registers of one 1024-thread CTA; every strip then
// waits for the one warp that factors the panel, and nothing else can run on the SM meanwhile (ncu:
// 24 % of the stall samples sit behind that barrier).  Here a CTA has 512 threads and walks its matrix
// in slabs of 512 columns (16 words): lane l keeps words 4(l & 3) .. 4(l & 3) + 3 of rows
// 64w + 8(l >> 2) + i in r[i] (uint4, i < 8), so ONE 128-bit shared-memory read updates four words and a
// warp instruction updates eight rows.  With 64 registers per thread two CTAs -- two independent
// matrices -- share an SM, and the panel / barrier latency of one is filled by the table reads of the
// other.
//
//   panel     byte space, one warp, as in gf2_m4r.cu: a row enters an 8-column strip only through its
//             strip byte, so the 1024 x 8 panel reduces to a 256-entry problem (lane l owns the byte
//             values 8l..8l+7, four per register), seeded by one representative unused row per value
//             (REP).  Output: G[byte] = combination y of the strip-start pivot rows for a row with
//             that strip byte, and PY[u] for the pivot rows themselves.
//   apply     pivot rows are published from registers, all 2^k combinations are tabulated (TP, 32
//             words per entry = the 16 slab words twice, so that the two rows served by one
//             quarter-warp phase of an LDS.128 hit different banks), and a row update is ONE table
//             read: r[i] ^= TP[y_i]; address = one shift + one LOP3.
//   replay    the combination bytes of every block (1 byte per row) go to an L2-resident scratch
//             buffer, and the slabs to the right stream them back with cp.async (three-deep ring),
//             one block ahead of the table reads.
//   lazy      (round 2) in the discovery loop the block found in strip s is applied while warp 0 already
//             factors the panel of strip s + 1: the next strip's byte of a row needs only the tracked
//             word, updated from the eight published pivot rows.  Three block barriers per strip
//             instead of four; 4.15 -> 3.85 us per strip.
//   measured  clock64 phase profile of one CTA (two per SM), cycles per replayed block: loop top 534,
//             publish 347, barrier 141, tabulate 1343, barrier 150, table reads 856 (= 3371); per discovery
//             strip ~7000.  Tried on top and dropped (both bit-exact): both table copies written by one
//             thread (half the pivot-row loads, 8 of 16 warps tabulating): 3 % slower; a software-pipelined
//             replay (next block tabulated into a second table in the same phase as the current block's
//             reads, its pivot rows served first; tools/experiments/gf2_m4r2_pipelined_replay.cu.txt): 26 %
//             slower per replayed block -- with 32 warps per SM and ~1350 shared-memory wavefronts per block
//             the phases are bound by the two CTAs' combined shared-memory and issue traffic, not by the
//             barriers between them.
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
constexpr int oOwn = oBlk + 1024 * 4;               // uint32 [1024]     owning warp of each pivot of a block
constexpr int kSmemBytes = oOwn + 1024 * 4;

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
__device__ __forceinline__ void xor4(uint4& a, const uint4& b) { a.x ^= b.x; a.y ^= b.y; a.z ^= b.z; a.w ^= b.w; }
__device__ __forceinline__ void xor4_if(uint4& a, const uint4& b, bool c) {
    const uint32_t mk = c ? 0xFFFFFFFFu : 0u;
    a.x ^= b.x & mk; a.y ^= b.y & mk; a.z ^= b.z & mk; a.w ^= b.w & mk;
}
// PTX prmt: only selector nibbles whose output byte is used need to be valid
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
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
    uint32_t* const own = reinterpret_cast<uint32_t*>(smem + oOwn);

    // tid is read once through volatile asm: the compiler otherwise re-reads the special register (a
    // scoreboard stall each time) whenever register pressure makes it drop a derived index.
    int tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, j = lane & 3;               // row group within the warp, word quad in the slab
    const int nw = (int)(blockDim.x >> 5);
    const int mrows = nw * 64;                           // rows incl. padding
    const int W32 = ((n + 63) >> 6) * 2;                 // 32-bit words per packed row
    const int nslabs = (W32 + kSW - 1) / kSW;
    const int npiv = m < n ? m : n;
    const int ra = warp * 64 + lane, rb = ra + 32;       // the two rows whose scalars this thread tracks
    const int row0 = (tid & ~3) << 1;                    // = 64 warp + 8 g: r[i] belongs to row row0 + i
    // byte offset of my quad inside a table entry: copy (g & 1) of the 16 slab words, quad j
    const uint32_t laneoff = (uint32_t)(tid & 7) << 4;   // = 64 (g & 1) + 16 j
    const uint32_t warp_nib = (uint32_t)warp * 0x11111111u;
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

        uint4 r[8];

        // publish: the owners of the block's pivot rows write their slab words to P.  own[] holds the
        // owning warp of each of the block's pivots, one nibble per pivot.
        auto publish = [&](int bi, int Kb, int k) {
            const uint32_t x = own[bi] ^ warp_nib;
            uint32_t z = ~(x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x11111111u;   // nibbles equal to my warp
            while (z != 0u) {
                const int u = (__ffs(z) - 1) >> 2;
                z &= z - 1u;
                if (u < k) {
                    const int p = pivrow[Kb + u];
                    const uint4 v = pick4(r, p & 7);
                    if (((p >> 3) & 7) == g) *reinterpret_cast<uint4*>(P + u * kSW + j * 4) = v;
                }
            }
        };
#ifndef QCSS_M4R2_TAB128
        // tabulate: all combinations of the k published rows (entries beyond 2^k are never read; rows u >= k of P hold
        // stale data that only reaches those).  lane = word of the 128-byte entry (the 16 slab words twice): the pivot
        // rows are read with 32-bit loads (one wavefront per row per warp -- a 128-bit load costs four however many
        // lanes share an address) and a warp pass writes eight entries, 128 contiguous bytes per store: 432 wavefronts
        // per block instead of 640.  Same-run A/B (tools/gf2_m4r2_probe.cu, -DQCSS_M4R2_TAB128 = the round-1 form
        // below): 2.06 / 3.76 / 7.14 ms against 2.13 / 3.90 / 7.41 ms per 1184 matrices at n = 1024 / 2048 / 4096.
        // skip0: warp 0 sits out (discovery: it factors the next panel meanwhile) and warps 1 .. nw-1 cover the entries
        auto tabulate = [&](int k, bool skip0) {
            const uint32_t* pw = P + (lane & (kSW - 1));
            const int entries = 1 << k;
            const int w0 = skip0 ? warp - 1 : warp, wn = skip0 ? nw - 1 : nw;
            if (w0 < 0) return;
            for (int E = w0 * 8; E < entries; E += wn * 8) {
                uint32_t base = 0u;
#pragma unroll
                for (int u = 3; u < 8; ++u)
                    if ((E >> u) & 1) base ^= pw[u * kSW];       // warp-uniform
                const uint32_t p0 = pw[0], p1 = pw[kSW], p2 = pw[2 * kSW];
                const uint32_t c1 = base ^ p0, c2 = base ^ p1, c3 = c1 ^ p1;
                uint32_t* t = reinterpret_cast<uint32_t*>(smem + oTP) + E * 32 + lane;
                t[0 * 32] = base; t[1 * 32] = c1; t[2 * 32] = c2; t[3 * 32] = c3;
                t[4 * 32] = base ^ p2; t[5 * 32] = c1 ^ p2; t[6 * 32] = c2 ^ p2; t[7 * 32] = c3 ^ p2;
            }
        };
#else
        // tabulate: all combinations of the k published rows (entries beyond 2^k are never read).  One
        // warp pass writes 16 entries: lane group g covers entries E + 4(g >> 1) + {0..3}, copy g & 1.
        // skip0: warp 0 sits out (discovery: it factors the next panel meanwhile) and warps 1 .. nw-1 cover the entries
        auto tabulate = [&](int k, bool skip0) {
            // Rows u >= k of P hold stale data; they only reach entries >= 2^k, which are never read.
            const uint4* pq = reinterpret_cast<const uint4*>(P) + j;      // quad j of pivot row u: pq[4 * u]
            const int entries = 1 << k;
            const int es = g >> 1;
            const int w0 = skip0 ? warp - 1 : warp, wn = skip0 ? nw - 1 : nw;
            if (w0 < 0) return;
            for (int E = w0 * 16; E < entries; E += wn * 16) {
                uint4 c0 = make_uint4(0u, 0u, 0u, 0u);
                xor4_if(c0, pq[4 * 2], es & 1);
                xor4_if(c0, pq[4 * 3], es & 2);
#pragma unroll
                for (int u = 4; u < 8; ++u)
                    if ((E >> u) & 1) xor4(c0, pq[4 * u]);       // warp-uniform
                const int e = E + 4 * es;
                uint8_t* t = smem + oTP + e * 128 + laneoff;
                *reinterpret_cast<uint4*>(t + 0 * 128) = c0;
                const uint4 p0 = pq[0], p1 = pq[4];
                uint4 c1 = c0;
                xor4(c1, p0);
                *reinterpret_cast<uint4*>(t + 1 * 128) = c1;
                xor4(c0, p1);
                *reinterpret_cast<uint4*>(t + 2 * 128) = c0;
                xor4(c1, p1);
                *reinterpret_cast<uint4*>(t + 3 * 128) = c1;
            }
        };
#endif
        // table_reads: one 128-bit read per row; the combination bytes of my group's 8 rows are 2 words
        auto table_reads = [&](const uint8_t* ybase) {
            const uint2 yv = *reinterpret_cast<const uint2*>(ybase + row0);
            const uint8_t* tp = smem + oTP;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const uint32_t w = q ? yv.y : yv.x;
                const uint32_t a0 = ((w << 7) & 0x7F80u) | laneoff;
                const uint32_t a1 = ((w >> 1) & 0x7F80u) | laneoff;
                const uint32_t a2 = ((w >> 9) & 0x7F80u) | laneoff;
                const uint32_t a3 = ((w >> 17) & 0x7F80u) | laneoff;
                xor4(r[4 * q + 0], *reinterpret_cast<const uint4*>(tp + a0));
                xor4(r[4 * q + 1], *reinterpret_cast<const uint4*>(tp + a1));
                xor4(r[4 * q + 2], *reinterpret_cast<const uint4*>(tp + a2));
                xor4(r[4 * q + 3], *reinterpret_cast<const uint4*>(tp + a3));
            }
        };
        // prefetch: combination bytes of recorded block bi -> ring slot bi % 3
        auto prefetch = [&](int bi) {
            if (tid * 16 < mrows) cp_async16(Yb + (bi % 3) * 1024 + tid * 16, Yg + (size_t)bi * mrows + tid * 16);
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
                    const uint32_t* ptr = src + (size_t)row * W32 + wi;
                    if (wi < W32) lo = __ldg(reinterpret_cast<const uint2*>(ptr));
                    if (wi + 2 < W32) hi = __ldg(reinterpret_cast<const uint2*>(ptr + 2));
                }
                r[i] = make_uint4(lo.x & cm[0], lo.y & cm[1], hi.x & cm[2], hi.y & cm[3]);
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
                    publish(bi, (int)(e & 0xFFFFu), k);
                    cp_async_wait<1>();                  // block bi has landed (bi + 1 may be in flight)
                    __syncthreads();
                    tabulate(k, false);
                    __syncthreads();
                    table_reads(Yb + (bi % 3) * 1024);
                }
                cp_async_wait<0>();
            }
            // ---- discovery: strips of 8 columns of this slab --------------------------------------
            // The block found in strip s is applied LAZILY: its tabulation and table reads run in warps 1 .. nw-1 while
            // warp 0 already factors the panel of strip s + 1.  That is possible because the next strip's byte of a row
            // needs only the tracked word: cur ^= XOR of the published pivot rows' word cw selected by the row's
            // combination byte -- the same value the table entry holds, without the table.  (Round-1 form: panel, then
            // publish / tabulate / read with four block barriers per strip; measured 4.15 us per strip against 1.69 us
            // for a replayed block, tools/gf2_shapes.py.)  `pending` is uniform over the CTA.
            const int slab_words = (W32 - slab * kSW) < kSW ? (W32 - slab * kSW) : kSW;
            bool pending = false;                        // block nblk - 1 is published and recorded but not yet applied to r[]
            int pend_k = 0;
            // apply the pending block with every warp (word boundaries, end of the slab)
            auto flush = [&]() {
                if (pending) {
                    tabulate(pend_k, false);
                    __syncthreads();
                    table_reads(Yb);
                    pending = false;
                }
            };
            for (int cw = 0; cw < slab_words && K < m; ++cw) {
                if ((slab * kSW + cw) * 32 >= n) break;
                flush();                                 // r[] must be current before its word cw is read out
                // my rows' current word cw, out of the registers of the lanes that hold it
                if (j == (cw >> 2)) {
                    const int c = cw & 3;
#pragma unroll
                    for (int i = 0; i < 8; ++i) S32[row0 + i] = comp4(r[i], c);
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
                    // (b) warp 0 factors the panel in byte space; the other warps apply the pending block meanwhile
                    if (warp == 0) {
                        // Lane l owns the byte values 8l .. 8l+7 (value 8l+e in byte e & 3 of red{e >> 2}), their reduced
                        // value (red) and combination byte (y).  A column is branch-free: every lane keys its last
                        // candidate (a present value whose reduced byte has the column's bit) as [reduced byte, y,
                        // row | 0x8000] and ONE warp max-reduction (REDUX, uniform datapath; a shuffle would queue behind
                        // the table reads in the load/store pipe) elects the pivot and broadcasts all of it; a column
                        // without candidates reduces to 0 and multiplies its update masks by zero
                        // (tools/experiments/panel_bench.cu: 1100 cycles alone, 1500 under 31 warps of table reads; the
                        // round-1 ballot + shuffle form with its branches: 1450 / 2300).
                        const uint4 q = reinterpret_cast<const uint4*>(rep)[lane];   // rows of the values 8l..8l+7 | 0x8000
                        const uint32_t ql0 = prmt(q.x, q.y, 0x6420u), ql1 = prmt(q.z, q.w, 0x6420u);   // low bytes of the rows
                        const uint32_t qh0 = prmt(q.x, q.y, 0x7531u), qh1 = prmt(q.z, q.w, 0x7531u);   // high bytes | 0x80
                        const uint32_t prs0 = (qh0 >> 7) & 0x01010101u, prs1 = (qh1 >> 7) & 0x01010101u;
                        uint32_t red0 = 0x03020100u + 0x08080808u * (uint32_t)lane, red1 = red0 + 0x04040404u;
                        uint32_t y0 = 0u, y1 = 0u;
                        uint32_t myrow = 0u, mycol = 0u;       // lane u < k: pivot u
                        uint32_t kmask = 0x01010101u;          // bit k of every byte lane
                        uint32_t ku = 0u;
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
                            if (found != 0u && lane == (int)ku) {
                                myrow = pack & 0x3FFu;
                                mycol = (uint32_t)col;
                            }
                            kmask <<= found;
                            ku += found;
                        }
                        const int k = (int)ku;
                        reinterpret_cast<uint2*>(G)[lane] = make_uint2(y0, y1);
                        __syncwarp();
                        uint32_t nib = 0u;
                        if (lane < k) {
                            // A pivot row must keep a single 1 in its own column: its combination byte is that of
                            // (its strip byte ^ the unit byte of its column) = {itself} ^ G[unit], by linearity.
                            const uint32_t py = G[1u << mycol] ^ (1u << lane);
                            const int prow = (int)myrow;
                            pivrow[K + lane] = (int16_t)prow;
                            pivcol[K + lane] = c0 + (int)mycol;
                            rowpiv[prow] = (int16_t)(K + lane);
                            PY[lane] = (uint8_t)py;
                            nib = (uint32_t)(prow >> 6) << (4 * lane);
                        }
                        nib = __reduce_or_sync(0xFFFFFFFFu, nib);
                        if (lane == 0) {
                            misc[0] = k;
                            if (k > 0) {
                                blk[nblk] = (uint32_t)K | ((uint32_t)k << 16);
                                own[nblk] = nib;
                            }
                        }
                    } else if (pending) {
                        tabulate(pend_k, true);
                        asm volatile("bar.sync 1, %0;" ::"r"((nw - 1) * 32) : "memory");   // warps 1 .. nw-1 only
                        table_reads(Yb);
                    }
                    __syncthreads();
                    if (pending && warp == 0) table_reads(Yb);     // my rows: the table stays intact until the next tabulation
                    pending = false;
                    const int k = misc[0];
                    if (k > 0) {
                        // (c) my rows' combination bytes: by strip byte, or the pivot's own entry
                        uint32_t ya = G[byte_a], yb = G[byte_b];
                        const int pa = rowpiv[ra], pb = rowpiv[rb];
                        if (pa >= K) ya = PY[pa - K];
                        if (pb >= K) yb = PY[pb - K];
                        used_a = pa >= 0 || ra >= m;
                        used_b = pb >= 0 || rb >= m;
                        __syncwarp();                                // warp 0: its table_reads above read Yb
                        Yb[ra] = (uint8_t)ya;
                        Yb[rb] = (uint8_t)yb;
                        Yg[(size_t)nblk * mrows + ra] = (uint8_t)ya;     // for the replays
                        Yg[(size_t)nblk * mrows + rb] = (uint8_t)yb;
                        publish(nblk, K, k);
                        __syncthreads();
                        // the tracked word after this block, from the published pivot rows alone
                        uint32_t ta = 0u, tb = 0u;
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const uint32_t pw = P[u * kSW + cw];
                            ta ^= pw & (0u - ((ya >> u) & 1u));
                            tb ^= pw & (0u - ((yb >> u) & 1u));
                        }
                        cur_a ^= ta;
                        cur_b ^= tb;
                        pending = true;
                        pend_k = k;
                        K += k;
                        ++nblk;
                        if (nw == 1) flush();                        // a one-warp CTA has nobody to overlap with
                    }
                }
            }
            flush();
            // ---- write the slab out in pivot order; rows without a pivot so far are zero here -----
            __syncthreads();
            if (wi < W32) {
                const bool hi_ok = wi + 2 < W32;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int pk = rowpiv[row0 + i];
                    if (pk >= 0) {
                        uint32_t* ptr = dst + (size_t)pk * W32 + wi;
                        *reinterpret_cast<uint2*>(ptr) = make_uint2(r[i].x, r[i].y);
                        if (hi_ok) *reinterpret_cast<uint2*>(ptr + 2) = make_uint2(r[i].z, r[i].w);
                    }
                }
                for (int row = K + warp * 8 + g; row < m; row += nw * 8) {
                    uint32_t* ptr = dst + (size_t)row * W32 + wi;
                    *reinterpret_cast<uint2*>(ptr) = make_uint2(0u, 0u);
                    if (hi_ok) *reinterpret_cast<uint2*>(ptr + 2) = make_uint2(0u, 0u);
                }
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

// Two 512-thread CTAs per SM pay off once a matrix fills most of a CTA; smaller matrices keep more
// warps busy in gf2_m4r.cu (measured: 256 x 512 1.4 ms vs 3.5 ms per 4096, 768 x 1600 6.0 vs 7.2 ms per 2048).
bool gf2_m4r2_supported(int m, int n) { return m >= 640 && m <= 1024 && n >= 1; }

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
    {   // keep freed scratch in the stream-ordered pool instead of returning it to the OS at every sync
        static bool pool_set[64] = {};
        if (dev >= 0 && dev < 64 && !pool_set[dev]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                unsigned long long keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            pool_set[dev] = true;
        }
    }
    if ((err = cudaMallocAsync(reinterpret_cast<void**>(&d_scratch), scratch, stream)) != cudaSuccess) return err;
    k_gf2_m4r2<<<grid, threads, kSmemBytes, stream>>>(reinterpret_cast<const uint32_t*>(in), batch, m, n,
                                                     reinterpret_cast<uint32_t*>(out), rank, pivots, d_scratch,
                                                     cap_blocks);
    err = cudaGetLastError();
    const cudaError_t ferr = cudaFreeAsync(d_scratch, stream);
    return err != cudaSuccess ? err : ferr;
}

}  // namespace qcss
