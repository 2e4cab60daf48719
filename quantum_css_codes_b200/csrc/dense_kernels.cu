// K1': syndrome of a large DENSE parity check on the 5th-generation tensor cores.
//
// S = H.E mod 2 (css_code.py:728) is a real dense contraction only when H is dense: then it is
// computed as an int8 GEMM  D[row][shot] = sum_j H[row][j] * E[j][shot]  with 0/1 operands and
// int32 accumulators in TMEM, and the epilogue keeps bit 0 of each accumulator.
//
//   A = H     : u8, K-major canonical no-swizzle core-matrix layout, prepared once on the host
//               (qcss_code_create) in exactly the order the kernel stages it, so staging a
//               128-row x 64-qubit block is a contiguous 8 KB cp.async copy.
//   B = E     : error bit planes are expanded to u8 in shared memory by the CTA itself
//               (4 bits -> 4 bytes with one IMAD + one LOP3), in the MN-major canonical layout:
//               16 consecutive shots of one qubit are one 16-byte row of a core matrix.
//   D         : two 128 x 256 accumulators (two 128-row tiles of H share the expanded E tile) fill the
//               512 TMEM columns; tcgen05.mma.cta_group::1.kind::i8, M = 128, N = 256, K = 32.
//   pipeline  : warp-specialised.  Warps 0-7 (producers) stage chunks of 64 qubits into a ring of 6
//               shared-memory stages (H block and raw error bits by cp.async three chunks ahead,
//               then the bit expansion) and arrive on the stage's `full` mbarrier; one lane of warp 8
//               waits for `full`, issues the chunk's four MMAs and commits them to the stage's `free`
//               mbarrier, which the producers wait on before refilling it.  The MMA lane never joins
//               a block barrier: with the first version (every thread staged, one __syncthreads per
//               chunk, thread 0 issuing) the tensor pipe idled 70 % of the time although neither the
//               expansion nor the copies were the limit (timing knobs: 6.59 ms as is, 5.75 ms with
//               all data movement removed, 2.68 ms with the MMAs issued back to back).
//   epilogue  : tcgen05.ld 32 columns at a time, bit 0 of 32 accumulators -> one syndrome word.
#include <cuda_runtime.h>

#include "launch.h"

namespace qcss {

namespace {

constexpr int kProducers = 256;              // warps 0-7: stage H, expand E; also the epilogue warps
constexpr int kMmaThreads = kProducers + 32; // warp 8: one lane issues every tcgen05.mma
constexpr int kMT = 2;                       // 128-row tiles of H per CTA
constexpr int kNT = 256;                     // shots per CTA
constexpr int kKC = 64;                      // qubits per stage
constexpr int kStages = 6;
constexpr int kABlock = 128 * kKC;           // bytes of one 128-row x 64-qubit block of H
constexpr int kABytes = kMT * kABlock;       // 16 KB
constexpr int kBBytes = kKC * kNT;           // 16 KB
constexpr int kRBytes = kKC * (kNT / 8);      // 2 KB: the chunk's raw error bits (64 qubits x 256 shots)
constexpr int kStageBytes = kABytes + kBBytes + kRBytes;

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

__global__ void __launch_bounds__(kMmaThreads, 1)
k_syndrome_mma(const uint8_t* __restrict__ hq, int m, int kchunks, int mgroups, const uint32_t* __restrict__ e,
               int n, int64_t e_stride, uint32_t* __restrict__ s, int64_t s_stride, int64_t words,
               uint32_t tail_mask) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[kStages];
    __shared__ __align__(8) uint64_t free_bar[kStages];
    __shared__ __align__(8) uint64_t done_bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mg = blockIdx.x % mgroups;
    const int64_t tile = blockIdx.x / mgroups;
    const int64_t w0 = tile * (kNT / 32);                     // first shot word of this tile

    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&full_bar[i])),
                         "r"(kProducers / 32 + 1));            // one arrival per producer warp + the H bulk copy
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&free_bar[i])));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&done_bar)));
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
    constexpr int kAhead = 4;                                 // H blocks and error bits are fetched four chunks ahead

    if (warp == kProducers / 32) {
        // ---- MMA warp: one lane, no block barriers ------------------------------------------------------
        if (lane == 0) {
            for (int kc = 0; kc < kchunks; ++kc) {
                const int st = kc % kStages, use = kc / kStages;
                mbar_wait_parity(&full_bar[st], (unsigned)(use & 1));
                const uint8_t* sA = smem + (size_t)st * kStageBytes;
                const uint32_t a0 = (unsigned)__cvta_generic_to_shared(sA), b0 = a0 + kABytes;
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
                             (unsigned)__cvta_generic_to_shared(&done_bar))
                         : "memory");
        }
    } else {
        // ---- producers: expansion role of this thread: qubit kq of the chunk, shot words 2*wp, 2*wp+1 ----
        const int kq = tid % kKC, wp = tid / kKC;             // 64 x 4
        auto issue_copies = [&](int c) {                      // 16 KB block of the pre-laid-out H + 2 KB of error bits
            const int st = c % kStages, use = c / kStages;
            if (use > 0) mbar_wait_parity(&free_bar[st], (unsigned)((use - 1) & 1));
            uint8_t* sA = smem + (size_t)st * kStageBytes;
            if (tid == 0) {
                // the pre-laid-out 16 KB block of H: ONE bulk copy (TMA), completing on the stage's full barrier
                const uint8_t* asrc = hq + ((size_t)mg * kchunks + c) * kABytes;
                const unsigned bar = (unsigned)__cvta_generic_to_shared(&full_bar[st]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kABytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 (unsigned)__cvta_generic_to_shared(sA)),
                             "l"(asrc), "r"(kABytes), "r"(bar)
                             : "memory");
            }
            {   // the 8 bytes of error bits THIS thread expands (qubit c*64 + kq, shot words w0 + 2wp, +1): no other
                // producer reads them, so no barrier is needed between the copy and the expansion
                const int j = c * kKC + kq;
                const int64_t w = w0 + 2 * wp;
                uint8_t* dstp = sA + kABytes + kBBytes + wp * (kKC * 8) + kq * 8;      // [wp][kq]: lanes contiguous
                if (j < n && w < e_stride) {
                    const unsigned dst = (unsigned)__cvta_generic_to_shared(dstp);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(e + (int64_t)j * e_stride + w)
                                 : "memory");
                } else {
                    *reinterpret_cast<uint2*>(dstp) = make_uint2(0u, 0u);
                }
            }
        };
        // Two chunks per iteration: their expansions interleave (the iteration is a chain of dependent
        // instructions, so one chunk at a time left the tensor pipe waiting ~1000 cycles per chunk), and the
        // refill of the stages four and five chunks ahead comes AFTER the expansion, off the critical path.
        auto expand = [&](int kc) {
            const int st = kc % kStages;
            uint8_t* sB = smem + (size_t)st * kStageBytes + kABytes;
            const uint2 bits = *reinterpret_cast<const uint2*>(sB + kBBytes + wp * (kKC * 8) + kq * 8);
            const uint32_t wv[2] = {bits.x, bits.y};
#pragma unroll
            for (int h = 0; h < 4; ++h) {                     // 4 groups of 16 shots
                const uint32_t v = wv[h >> 1] >> (16 * (h & 1));
                const uint4 bytes = make_uint4(spread4(v & 0xFu), spread4((v >> 4) & 0xFu),
                                               spread4((v >> 8) & 0xFu), spread4((v >> 12) & 0xFu));
                const int nb = (2 * wp) * 2 + h;              // 16-shot group index within the tile
                *reinterpret_cast<uint4*>(sB + (size_t)(kq / 8) * (kNT / 16) * 128 + nb * 128 + (kq % 8) * 16) = bytes;
            }
        };
        auto arrive = [&](int kc) {
            if (lane == 0)
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(
                                 (unsigned)__cvta_generic_to_shared(&full_bar[kc % kStages]))
                             : "memory");
        };
        for (int c = 0; c < kAhead && c < kchunks; c += 2) {          // chunks 0..3 as two groups
            issue_copies(c);
            if (c + 1 < kchunks) issue_copies(c + 1);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (int kc = 0; kc < kchunks; kc += 2) {
            const bool two = kc + 1 < kchunks;
            asm volatile("cp.async.wait_group 1;" ::: "memory");      // my bits of chunks kc, kc+1 have landed
            expand(kc);
            if (two) expand(kc + 1);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");            // generic writes -> tensor-core reads
            __syncwarp();
            arrive(kc);
            if (two) arrive(kc + 1);
            if (kc + kAhead < kchunks) issue_copies(kc + kAhead);
            if (kc + kAhead + 1 < kchunks) issue_copies(kc + kAhead + 1);
            asm volatile("cp.async.commit_group;" ::: "memory");                  // one group per iteration (possibly empty)
        }
    }
    // ---- epilogue: bit 0 of the accumulators -> syndrome words ------------------------------------
    mbar_wait_parity(&done_bar, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < kProducers / 32) {
        const int mt = warp >> 2, lane_base = (warp & 3) * 32;
        const int row = (mg * kMT + mt) * 128 + lane_base + lane;
        uint32_t out[kNT / 32];
#pragma unroll
        for (int c = 0; c < kNT / 32; ++c) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(mt * kNT + c * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            uint32_t word = 0u;
#pragma unroll
            for (int b = 0; b < 32; ++b) word |= (v[b] & 1u) << b;
            out[c] = word;
        }
        if (row < m) {
#pragma unroll
            for (int c = 0; c < kNT / 32; ++c) {
                const int64_t w = w0 + c;
                if (w < words) s[(int64_t)row * s_stride + w] = (w == words - 1) ? (out[c] & tail_mask) : out[c];
            }
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
    const int64_t grid = tiles * mgroups;
    if (grid <= 0 || grid > 0x7FFFFFFF) return cudaErrorInvalidValue;
    k_syndrome_mma<<<(unsigned)grid, kMmaThreads, smem, stream>>>(hq, m, kchunks, mgroups, e, n, e_stride, s, s_stride,
                                                                  words, tail_mask);
    return cudaGetLastError();
}

}  // namespace qcss
