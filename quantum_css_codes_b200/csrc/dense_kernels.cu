// K1': syndrome of a large DENSE parity check on the 5th-generation tensor cores.
//
// S = H.E mod 2 (css_code.py:728) is a real dense contraction only when H is dense: then it is
// computed as an int8 GEMM  D[row][shot] = sum_j H[row][j] * E[j][shot]  with 0/1 operands and
// int32 accumulators in TMEM, and the epilogue keeps bit 0 of each accumulator.
//
//   A = H     : u8, K-major canonical no-swizzle core-matrix layout, prepared once on the host
//               (qcss_code_create) in exactly the order the kernel stages it, so staging a
//               256-row x 64-qubit block is ONE contiguous 16 KB cp.async.bulk.
//   B = E     : error bit planes are expanded to u8 in shared memory by the CTA itself
//               (4 bits -> 4 bytes with one IMAD + one LOP3), in the MN-major canonical layout:
//               16 consecutive shots of one qubit are one 16-byte row of a core matrix.
//   D         : two 128 x 256 accumulators (two 128-row tiles of H share the expanded E tile) fill the
//               512 TMEM columns; tcgen05.mma.cta_group::1.kind::i8, M = 128, N = 256, K = 32.
//   CTA       : PERSISTENT, one per SM, looping over (shot tile, row group) work items; TMEM, barriers
//               and the pipeline state live across items.  Five roles:
//                 warps 0-7   expanders, four groups of two warps; group G owns the chunks g = G mod 4, so four
//                             per-chunk chains (load -> stage free -> stores -> proxy fence -> arrive) overlap.
//                             A chunk is 64 qubits x 256 shots; lane (plane = lane & 7 of an octet, quarter =
//                             lane >> 3) holds 64 shots of a plane in two registers, loaded one group-chunk
//                             ahead with a coalesced 8-byte __ldg (the four lanes of a plane read one 32-byte
//                             sector) and writes four 16-byte rows; a quarter-warp store covers the eight
//                             planes of one 128-byte core matrix, so the stores are bank-conflict free;
//                 warp 8      one lane issues every tcgen05.mma and commits to the stage's `free` barrier;
//                 warp 9      one lane runs ahead with the 16 KB bulk copies of the H blocks;
//                 warps 12-19 epilogue, one warpgroup per accumulator: tcgen05.ld 64 columns per wait, bit 0 of
//                             32 accumulators -> one syndrome word, while the expanders already fill the ring
//                             for the next item.
//   history   : round 1 staged the raw error bits with one 8-byte cp.async per thread, every lane on a
//               different plane: ncu showed the shared-memory pipe 68 % busy with 63 % of its wavefronts
//               flagged as conflicts (each returning sector wrote its own wavefront), and the kernel was
//               bound there -- ~680 LSU wavefronts + ~380 wavefronts of operand reads by the tensor core
//               per chunk against 512 cycles of MMA work (profiles/r01_dense_tcgen05_v2_ncu_summary.txt);
//               32768 one-shot CTAs each paid TMEM allocation, barrier set-up and an exposed epilogue.
#include <cuda_runtime.h>

#include "launch.h"

namespace qcss {

namespace {

constexpr int kExpanders = 256;              // warps 0-7
constexpr int kMmaWarp = 8, kTmaWarp = 9, kEpiWarp0 = 12;
constexpr int kDenseThreads = 640;         // warps 10-11 idle; warps 12-19 are two aligned warpgroups for the epilogue
constexpr int kMT = 2;                       // 128-row tiles of H per work item
constexpr int kNT = 256;                     // shots per work item
constexpr int kKC = 64;                      // qubits per stage
constexpr int kStages = 6;
constexpr int kGroups = 4;                   // expander groups (two warps each); group G owns chunks g = G mod 4
constexpr int kABlock = 128 * kKC;           // bytes of one 128-row x 64-qubit block of H
constexpr int kABytes = kMT * kABlock;       // 16 KB
constexpr int kBBytes = kKC * kNT;           // 16 KB
constexpr int kStageBytes = kABytes + kBBytes;

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;                  // descriptor version 1 (sm_100); no swizzle, base offset 0
    return d;
}

__device__ __forceinline__ uint32_t spread4(uint32_t nibble) {      // 4 bits -> 4 bytes of 0/1
    return (nibble * 0x00204081u) & 0x01010101u;
}

__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MMA_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MMA_DONE;\n"
        "bra MMA_WAIT;\n"
        "MMA_DONE:\n"
        "}\n" ::"r"(addr),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

__global__ void __launch_bounds__(kDenseThreads, 1)
k_syndrome_mma(const uint8_t* __restrict__ hq, int m, int kchunks, int mgroups, const uint32_t* __restrict__ e,
               int n, int64_t e_stride, uint32_t* __restrict__ s, int64_t s_stride, int64_t words,
               uint32_t tail_mask, int64_t items) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[kStages];
    __shared__ __align__(8) uint64_t free_bar[kStages];
    __shared__ __align__(8) uint64_t acc_full, acc_free;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&full_bar[i])),
                         "r"(2 + 1));                          // the two expander warps of the chunk's group + the H bulk copy
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&free_bar[i])));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&acc_full)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"((unsigned)__cvta_generic_to_shared(&acc_free)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (unsigned)__cvta_generic_to_shared(&tmem_base_s)),
                     "r"(kMT * kNT));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    // instruction descriptor: D = S32, A = B = unsigned 8-bit, A K-major, B MN-major, N = 256, M = 128
    const uint32_t idesc = (2u << 4) | (1u << 16) | ((uint32_t)(kNT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int64_t my_items = (items > blockIdx.x) ? (items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t total_chunks = my_items * kchunks;          // chunks flow through the ring across work items

    if (warp == kMmaWarp) {
        // ---- MMA issuer: one lane, no block barriers ------------------------------------------------------
        if (lane == 0) {
            int64_t g = 0;
            for (int64_t it = 0; it < my_items; ++it) {
                if (it > 0) {                                  // the epilogue has drained the accumulators of item it - 1
                    mbar_wait_parity(&acc_free, (unsigned)((it - 1) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                for (int kc = 0; kc < kchunks; ++kc, ++g) {
                    const int st = (int)(g % kStages);
                    mbar_wait_parity(&full_bar[st], (unsigned)((g / kStages) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a0 = (unsigned)__cvta_generic_to_shared(smem + (size_t)st * kStageBytes), b0 = a0 + kABytes;
#pragma unroll
                    for (int mt = 0; mt < kMT; ++mt) {
#pragma unroll
                        for (int ks = 0; ks < kKC / 32; ++ks) {
                            const uint64_t da = umma_desc(a0 + mt * kABlock + ks * 2 * 16 * 128, 16 * 128, 128);
                            const uint64_t db = umma_desc(b0 + ks * 4 * (kNT / 16) * 128, (kNT / 16) * 128, 128);
                            const uint32_t acc = (kc > 0 || ks > 0) ? 1u : 0u;
                            asm volatile(
                                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_base + mt * kNT),
                                "l"(da), "l"(db), "r"(idesc), "r"(acc)
                                : "memory");
                        }
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     (unsigned)__cvta_generic_to_shared(&free_bar[st]))
                                 : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                 (unsigned)__cvta_generic_to_shared(&acc_full))
                             : "memory");
            }
        }
    } else if (warp == kTmaWarp) {
        // ---- H loader: the pre-laid-out 16 KB block of every chunk, ONE bulk copy each, as far ahead as the ring allows
        if (lane == 0) {
            int64_t g = 0;
            for (int64_t it = 0; it < my_items; ++it) {
                const int mg = (int)((blockIdx.x + it * gridDim.x) % mgroups);
                for (int kc = 0; kc < kchunks; ++kc, ++g) {
                    const int st = (int)(g % kStages);
                    const int64_t use = g / kStages;
                    if (use > 0) mbar_wait_parity(&free_bar[st], (unsigned)((use - 1) & 1));
                    const uint8_t* asrc = hq + ((size_t)mg * kchunks + kc) * kABytes;
                    const unsigned bar = (unsigned)__cvta_generic_to_shared(&full_bar[st]);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kABytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     (unsigned)__cvta_generic_to_shared(smem + (size_t)st * kStageBytes)),
                                 "l"(asrc), "r"(kABytes), "r"(bar)
                                 : "memory");
                }
            }
        }
    } else if (warp < kExpanders / 32) {
        // ---- expanders: four groups of two warps, group G owns the chunks g = G (mod 4), so four per-chunk
        //      chains (load wait -> stage wait -> stores -> proxy fence -> arrive) run concurrently.  Within a chunk a
        //      warp covers 32 planes as four octets; lane = (plane of the octet, 64-shot quarter of the tile). ---------
        const int grp = warp >> 1, half = warp & 1;
        const int pl = lane & 7, q = lane >> 3;
        const uint32_t dst_off = kABytes + (uint32_t)(half * 4) * ((kNT / 16) * 128) + (uint32_t)pl * 16 + (uint32_t)q * (4 * 128);
        // fetch cursor: (tile, chunk) of this group's next load, advanced without divisions
        const int step_tile = (int)(gridDim.x / (unsigned)mgroups), step_mg = (int)(gridDim.x % (unsigned)mgroups);
        int64_t f_tile = blockIdx.x / mgroups, f_left = (total_chunks > grp) ? (total_chunks - grp + kGroups - 1) / kGroups : 0;
        int f_mg = (int)(blockIdx.x % mgroups), f_kc = grp;
        while (f_kc >= kchunks && f_left > 0) {                // fewer than four chunks per item
            f_kc -= kchunks;
            f_tile += step_tile; f_mg += step_mg;
            if (f_mg >= mgroups) { f_mg -= mgroups; ++f_tile; }
        }
        auto fetch = [&](uint2 (&v)[4]) {                      // my words of the group's next chunk (zeros past the batch / past n)
#pragma unroll
            for (int o = 0; o < 4; ++o) v[o] = make_uint2(0u, 0u);
            if (f_left > 0) {
                const int64_t w = f_tile * (kNT / 32) + 2 * q;
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const int j = f_kc * kKC + (half * 4 + o) * 8 + pl;
                    if (j < n && w < e_stride) v[o] = __ldg(reinterpret_cast<const uint2*>(e + (int64_t)j * e_stride + w));
                }
                --f_left;
                f_kc += kGroups;
                while (f_kc >= kchunks) {
                    f_kc -= kchunks;
                    f_tile += step_tile; f_mg += step_mg;
                    if (f_mg >= mgroups) { f_mg -= mgroups; ++f_tile; }
                }
            }
        };
        uint2 cur[4], nxt[4];
        fetch(cur);
        for (int64_t g = grp; g < total_chunks; g += kGroups) {
            fetch(nxt);                                        // chunk g + 4: in flight while chunk g expands
            const int st = (int)(g % kStages);
            const int64_t use = g / kStages;
            if (use > 0) mbar_wait_parity(&free_bar[st], (unsigned)((use - 1) & 1));
            uint8_t* dst = smem + (size_t)st * kStageBytes + dst_off;
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const uint32_t wv[2] = {cur[o].x, cur[o].y};
#pragma unroll
                for (int h = 0; h < 4; ++h) {                  // 4 groups of 16 shots: core matrices 4q .. 4q + 3
                    const uint32_t v = wv[h >> 1] >> (16 * (h & 1));
                    *reinterpret_cast<uint4*>(dst + o * ((kNT / 16) * 128) + h * 128) =
                        make_uint4(spread4(v & 0xFu), spread4((v >> 4) & 0xFu), spread4((v >> 8) & 0xFu), spread4((v >> 12) & 0xFu));
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic writes -> tensor-core reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[st]);
#pragma unroll
            for (int o = 0; o < 4; ++o) cur[o] = nxt[o];
        }
    } else if (warp >= kEpiWarp0) {
        // ---- epilogue: bit 0 of the accumulators -> syndrome words.  Two warpgroups, one per 128-row accumulator;
        //      a warp owns the TMEM lane quadrant warp % 4 and reads 64 columns per wait. -----------------------------
        const int lane_base = (warp & 3) * 32, mt = (warp - kEpiWarp0) >> 2;
        for (int64_t it = 0; it < my_items; ++it) {
            const int64_t item = blockIdx.x + it * gridDim.x;
            const int mg = (int)(item % mgroups);
            const int64_t w0 = (item / mgroups) * (kNT / 32);
            mbar_wait_parity(&acc_full, (unsigned)(it & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = (mg * kMT + mt) * 128 + lane_base + lane;
            uint32_t* dst = s + (int64_t)row * s_stride + w0;
#pragma unroll 1
            for (int c = 0; c < kNT / 32; c += 2) {
                uint32_t v[64];
                const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(mt * kNT + c * 32);
#pragma unroll
                for (int hlf = 0; hlf < 2; ++hlf) {
                    uint32_t* u = v + 32 * hlf;
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                          "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
                          "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
                          "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
                        : "r"(taddr + 32 * hlf));
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int hlf = 0; hlf < 2; ++hlf) {
                    uint32_t word = 0u;
#pragma unroll
                    for (int b = 0; b < 32; ++b) word |= (v[32 * hlf + b] & 1u) << b;
                    const int64_t w = w0 + c + hlf;
                    if (row < m && w < words) dst[c + hlf] = (w == words - 1) ? (word & tail_mask) : word;
                }
            }
            // every accumulator column has been read: hand TMEM back to the MMA issuer
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_free);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kMT * kNT));
}

}  // namespace

// Bytes of the pre-laid-out H for an m x n check matrix.
size_t dense_h_bytes(int m, int n) {
    const int mgroups = (m + kMT * 128 - 1) / (kMT * 128), kchunks = (n + kKC - 1) / kKC;
    return (size_t)mgroups * kchunks * kABytes;
}

// H (row-major 0/1 bytes) -> [mgroup][kchunk][row tile][K-major canonical core-matrix block]
void dense_h_layout(int m, int n, const uint8_t* H, uint8_t* out) {
    const int mgroups = (m + kMT * 128 - 1) / (kMT * 128), kchunks = (n + kKC - 1) / kKC;
    for (size_t i = 0; i < dense_h_bytes(m, n); ++i) out[i] = 0;
    for (int r = 0; r < m; ++r) {
        const int mg = r / (kMT * 128), mt = (r / 128) % kMT, rr = r % 128;
        for (int j = 0; j < n; ++j) {
            if (!(H[(size_t)r * n + j] & 1)) continue;
            const int kc = j / kKC, kk = j % kKC;
            const size_t off = ((size_t)mg * kchunks + kc) * kABytes + (size_t)mt * kABlock +
                               (size_t)(kk / 16) * 16 * 128 + (rr / 8) * 128 + (rr % 8) * 16 + (kk % 16);
            out[off] = 1;
        }
    }
    (void)mgroups;
}

cudaError_t launch_syndrome_mma(const uint8_t* hq, int m, int n, const uint32_t* e, int64_t e_stride, uint32_t* s,
                                int64_t s_stride, int64_t words, uint32_t tail_mask, cudaStream_t stream) {
    const int mgroups = (m + kMT * 128 - 1) / (kMT * 128), kchunks = (n + kKC - 1) / kKC;
    const int64_t tiles = (words + kNT / 32 - 1) / (kNT / 32);
    const size_t smem = (size_t)kStages * kStageBytes;
    cudaError_t err = cudaFuncSetAttribute(k_syndrome_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    const int64_t items = tiles * mgroups;                    // (shot tile, row group) work items, row groups adjacent
    if (items <= 0) return cudaErrorInvalidValue;
    int dev = 0, sms = 0;
    if ((err = cudaGetDevice(&dev)) != cudaSuccess) return err;
    if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return err;
    const int64_t grid = items < sms ? items : sms;           // persistent: one CTA per SM
    k_syndrome_mma<<<(unsigned)grid, kDenseThreads, smem, stream>>>(hq, m, kchunks, mgroups, e, n, e_stride, s, s_stride,
                                                                    words, tail_mask, items);
    return cudaGetLastError();
}

}  // namespace qcss
