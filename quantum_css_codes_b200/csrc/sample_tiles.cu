// K3 for codes of any size: depolarising Philox sampler fused into the sparse syndrome kernel, so that the
// 2 n error bits per shot of a large code (hypergraph product n = 1600: 400 B/shot) never exist in HBM --
// only the 2 m syndrome bits leave the SM (tile-major, the layout qcss_syndrome_tiles uses).
//
// One CTA works on a sub-tile of 256 shots (8 words) at a time:
//   sample   thread t draws site-words (qubit j, word w), j * 8 + w = t, t + 1024, ... with the K3 sampler
//            (core.cuh: gap sampler below p = 1/128, bit-serial above; Philox counter = (global word, site = j,
//            block), key = seed -- the stream of qcss_mc_sample with n > 32 sites) into two n x 32 B arrays (X
//            and Z words) in shared memory; on the gap path in two phases -- first blocks for everyone, a queue
//            for the few site-words that hold an error, one queued item per lane (see the kernel);
//   XOR      thread (row slot, 16-byte chunk q) folds the rows of parity_check_c2 over the X array and of
//            parity_check_c1 over the Z array (supports in shared memory, rows padded to groups of four entries
//            with an all-zero row as the pad: one 8-byte index load and four LDS.128 per group, no remainder loop)
//            and writes 16 bytes of the syndrome tile per row; optionally the sampled errors are written too.
// INT-bound by construction: one first-look Philox block per 32 shots per EIGHT qubits (core.cuh) against 7 loads +
// XORs per 128 shots per check.  Reading a resident batch instead costs (n + m) / 8 bytes per shot per type.
#include <cuda_runtime.h>

#include "launch.h"

namespace qcss {

namespace {

constexpr int kSampleThreads = 1024;
constexpr int kSubWords = 8;                       // words per sub-tile: 256 shots, 32-byte rows
constexpr int kTileWords = 32;                     // tile-major layout: 1024 shots per tile
constexpr int kSlots = kSampleThreads / 2;         // 512 row slots x 2 chunks of 16 bytes

struct SampleArgs {
    SparseRows hx;            // parity_check_c2: acts on X errors (which = 2)
    SparseRows hz;            // parity_check_c1: acts on Z errors (which = 1)
    uint32_t* sx;             // [tiles][hx.m][32] or null
    uint32_t* sz;             // [tiles][hz.m][32] or null
    uint32_t* ex;             // [tiles][n][32] or null
    uint32_t* ez;
    int64_t words;            // ceil(shots / 32)
    uint32_t tail_mask;
    uint64_t seed, first_word;
    uint32_t thr, use_gap;
    GapTable gap;
};

// supports in the padded form (launch.h: rows in groups of four entries, pad = the all-zero row n)
__device__ __forceinline__ void stage_csr(const SparseRows& h, uint16_t* ptr4, uint16_t* cols4) {
    for (int i = threadIdx.x; i <= h.m; i += kSampleThreads) ptr4[i] = (uint16_t)__ldg(h.row_ptr4 + i);
    for (int k = threadIdx.x; k < 4 * h.groups; k += kSampleThreads) cols4[k] = __ldg(h.cols4 + k);
}

__device__ __forceinline__ void xor_rows(const SparseRows& h, const uint16_t* ptr4, const uint16_t* cols4,
                                         const uint32_t* planes, uint32_t* out, int64_t tile, int sub, int64_t words,
                                         uint32_t tail_mask) {
    const int q = threadIdx.x & 1, slot = threadIdx.x >> 1;
    const int64_t w0 = tile * kTileWords + sub * kSubWords + q * 4;          // first of this thread's 4 words
    const uint32_t* const mine = planes + q * 4;
    for (int i = slot; i < h.m; i += kSlots) {
        uint4 acc = make_uint4(0u, 0u, 0u, 0u);
        const int g0 = ptr4[i], g1 = ptr4[i + 1];
        for (int g = g0; g < g1; ++g) {                       // four entries per step: one 8-byte index load, four LDS.128
            const uint2 cc = *reinterpret_cast<const uint2*>(cols4 + 4 * g);
            const uint4 v0 = *reinterpret_cast<const uint4*>(mine + (cc.x & 0xFFFFu) * kSubWords);
            const uint4 v1 = *reinterpret_cast<const uint4*>(mine + (cc.x >> 16) * kSubWords);
            const uint4 v2 = *reinterpret_cast<const uint4*>(mine + (cc.y & 0xFFFFu) * kSubWords);
            const uint4 v3 = *reinterpret_cast<const uint4*>(mine + (cc.y >> 16) * kSubWords);
            acc.x ^= v0.x ^ v1.x; acc.y ^= v0.y ^ v1.y; acc.z ^= v0.z ^ v1.z; acc.w ^= v0.w ^ v1.w;
            acc.x ^= v2.x ^ v3.x; acc.y ^= v2.y ^ v3.y; acc.z ^= v2.z ^ v3.z; acc.w ^= v2.w ^ v3.w;
        }
        uint32_t o[4] = {acc.x, acc.y, acc.z, acc.w};
        if (w0 + 4 > words - 1) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const int64_t w = w0 + v;
                if (w >= words) o[v] = 0u;
                else if (w == words - 1) o[v] &= tail_mask;
            }
        }
        *reinterpret_cast<uint4*>(out + ((size_t)tile * h.m + i) * kTileWords + sub * kSubWords + q * 4) =
            make_uint4(o[0], o[1], o[2], o[3]);
    }
}

__global__ void __launch_bounds__(kSampleThreads, 1)
k_sample_syndrome_tiles(const __grid_constant__ SampleArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int n = a.hx.n;
    uint32_t* const px = reinterpret_cast<uint32_t*>(smem);                  // [n + 1][8]; row n stays zero (CSR pad)
    uint32_t* const pz = px + (size_t)(n + 1) * kSubWords;
    uint16_t* const cols_x = reinterpret_cast<uint16_t*>(pz + (size_t)(n + 1) * kSubWords);   // 8-byte aligned
    uint16_t* const cols_z = cols_x + 4 * a.hx.groups;
    uint16_t* const ptr_x = cols_z + 4 * a.hz.groups;
    uint16_t* const ptr_z = ptr_x + ((a.hx.m + 2) & ~1);
    uint16_t* const queue = ptr_z + ((a.hz.m + 2) & ~1);                     // [n * 8] site-words with an error
    __shared__ GapTable s_gap;
    __shared__ int q_count;
    if (threadIdx.x == 0) q_count = 0;
    if (threadIdx.x < 32) s_gap.cdf[threadIdx.x] = a.gap.cdf[threadIdx.x];
    if (threadIdx.x == 32) s_gap.inv = a.gap.inv;
    if (a.sx != nullptr) stage_csr(a.hx, ptr_x, cols_x);
    if (a.sz != nullptr) stage_csr(a.hz, ptr_z, cols_z);
    if (threadIdx.x < kSubWords) px[n * kSubWords + threadIdx.x] = pz[n * kSubWords + threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t cdf31 = s_gap.cdf[31], look_hi = gap_look16(cdf31) << 16;
    const int64_t tiles = (a.words + kTileWords - 1) / kTileWords;
    const int64_t subs = tiles * (kTileWords / kSubWords);
    for (int64_t st = blockIdx.x; st < subs; st += gridDim.x) {
        const int64_t tile = st / (kTileWords / kSubWords);
        const int sub = (int)(st % (kTileWords / kSubWords));
        const int64_t wbase = tile * kTileWords + sub * kSubWords;
        auto put = [&](int idx, uint32_t x, uint32_t z, int64_t gw) {
            if (gw == a.words - 1) { x &= a.tail_mask; z &= a.tail_mask; }
            px[idx] = x;
            pz[idx] = z;
            const int j = idx / kSubWords, w = idx % kSubWords;
            if (a.ex != nullptr) a.ex[((size_t)tile * n + j) * kTileWords + sub * kSubWords + w] = x;
            if (a.ez != nullptr) a.ez[((size_t)tile * n + j) * kTileWords + sub * kSubWords + w] = z;
        };
        const int w = threadIdx.x % kSubWords;              // 1024 % 8 == 0: a thread keeps its word column
        const int64_t gw = wbase + w;
        const int total = n * kSubWords;
        if (a.use_gap) {
            // gap path: everything starts as "no error" (16-byte stores; px and pz are contiguous)
            uint4* const z4 = reinterpret_cast<uint4*>(px);
            for (int i = threadIdx.x; i < 2 * (total + kSubWords) / 4; i += kSampleThreads) z4[i] = make_uint4(0u, 0u, 0u, 0u);
            if (a.ex != nullptr || a.ez != nullptr) {
                for (int i = threadIdx.x; i < total / 4; i += kSampleThreads) {
                    const int j = i / (kSubWords / 4), c = i % (kSubWords / 4);
                    const size_t off = ((size_t)tile * n + j) * kTileWords + sub * kSubWords + c * 4;
                    if (a.ex != nullptr) *reinterpret_cast<uint4*>(a.ex + off) = make_uint4(0u, 0u, 0u, 0u);
                    if (a.ez != nullptr) *reinterpret_cast<uint4*>(a.ez + off) = make_uint4(0u, 0u, 0u, 0u);
                }
            }
            // (phase 2, which overwrites single words of these, comes after the barrier that ends phase 1)
        }
        if (gw >= a.words) {
            if (!a.use_gap)
                for (int idx = threadIdx.x; idx < total; idx += kSampleThreads) put(idx, 0u, 0u, gw);
        } else if (a.use_gap) {
            // Phase 1: first-look Philox block of every eight site-words (two per iteration: independent 10-round chains); 97 %
            // of them (p = 1e-3) hold no error and are done after one compare.  The others are QUEUED instead of
            // being finished in place: with ~1 erring lane per warp-instruction, finishing in place makes every
            // warp pay the gap logic for one useful lane (65 % of the warps at p = 1e-3; ncu: the logic was
            // most of the 161 instructions per site-word).  Phase 2 hands the queue out one item per lane.
            Philox ph;
            ph.k0 = (uint32_t)a.seed;
            ph.k1 = (uint32_t)(a.seed >> 32);
            const uint64_t g = a.first_word + (uint64_t)gw;
            const uint32_t g_lo = (uint32_t)g, g_hi = (uint32_t)(g >> 32);
            // (the arrays were zeroed with 16-byte stores before this phase; a clean site-word costs nothing more.)
            // Eight qubits share their first-look block (core.cuh): group index gi = (qubit >> 3) * kSubWords + word.
            // The hits of a group collect in a mask, branch-free; one shared atomic per thread with hits.
            auto first_look = [&](int gi, const uint32_t (&hb)[4]) {
                const int j0 = 8 * (gi / kSubWords);
                uint32_t m = 0u;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (gap_look(hb, c, look_hi)) m |= 1u << c;
                if (j0 + 8 > n) m &= (1u << (n - j0)) - 1u;
                if (m != 0u) {
                    int at = atomicAdd(&q_count, (int)popc32(m));
                    for (; m != 0u; m &= m - 1u) queue[at++] = (uint16_t)((j0 + (int)ctz32(m)) * kSubWords + w);
                }
            };
            const int groups = ((n + 7) / 8) * kSubWords;
            int gi = threadIdx.x;
            for (; gi + kSampleThreads < groups; gi += 2 * kSampleThreads) {
                uint32_t b0[4], b1[4];
                gap_first8(ph, g_lo, g_hi, (uint32_t)(gi / kSubWords), b0);
                gap_first8(ph, g_lo, g_hi, (uint32_t)((gi + kSampleThreads) / kSubWords), b1);
                first_look(gi, b0);
                first_look(gi + kSampleThreads, b1);
            }
            if (gi < groups) {
                uint32_t b0[4];
                gap_first8(ph, g_lo, g_hi, (uint32_t)(gi / kSubWords), b0);
                first_look(gi, b0);
            }
        } else {
            for (int idx = threadIdx.x; idx < total; idx += kSampleThreads) {
                uint32_t x, z;
                sample_site_word(a.seed, a.first_word + (uint64_t)gw, (uint32_t)(idx / kSubWords), a.thr, x, z);
                put(idx, x, z, gw);
            }
        }
        __syncthreads();
        if (a.use_gap) {
            const int count = q_count;
            for (int k = threadIdx.x; k < count; k += kSampleThreads) {
                const int idx = queue[k];
                const int64_t qw = wbase + idx % kSubWords;
                uint32_t x, z;
                sample_site_word_gap(a.seed, a.first_word + (uint64_t)qw, (uint32_t)(idx / kSubWords), s_gap, cdf31, x, z);
                put(idx, x, z, qw);
            }
            __syncthreads();
            if (threadIdx.x == 0) q_count = 0;
        }
        if (a.sx != nullptr) xor_rows(a.hx, ptr_x, cols_x, px, a.sx, tile, sub, a.words, a.tail_mask);
        if (a.sz != nullptr) xor_rows(a.hz, ptr_z, cols_z, pz, a.sz, tile, sub, a.words, a.tail_mask);
        __syncthreads();
    }
}

}  // namespace

// cudaErrorInvalidValue when the two error arrays and the supports do not fit shared memory (n > ~3000).
cudaError_t launch_sample_syndrome_tiles(const SparseRows& hx, const SparseRows& hz, uint32_t* sx, uint32_t* sz,
                                         uint32_t* ex, uint32_t* ez, int64_t words, uint32_t tail_mask, uint64_t seed,
                                         uint64_t first_word, uint32_t thr, uint32_t use_gap, const GapTable& gap,
                                         cudaStream_t stream) {
    if (hx.n != hz.n || hx.n * kSubWords > 65535 || hx.groups > 65535 || hz.groups > 65535 || hx.m > 65534 || hz.m > 65534) return cudaErrorInvalidValue;
    const size_t smem = (size_t)2 * (hx.n + 1) * kSubWords * 4 + 8 * (size_t)(hx.groups + hz.groups) +
                        2 * (size_t)(((hx.m + 2) & ~1) + ((hz.m + 2) & ~1)) +
                        2 * (size_t)hx.n * kSubWords;                              // + the queue of erring site-words
    if (smem > 226 * 1024) return cudaErrorInvalidValue;
    cudaError_t err = cudaFuncSetAttribute(k_sample_syndrome_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    int dev = 0, sms = 0;
    if ((err = cudaGetDevice(&dev)) != cudaSuccess) return err;
    if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return err;
    int per_sm = 1;
    if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sample_syndrome_tiles, kSampleThreads, smem)) !=
        cudaSuccess) return err;
    if (per_sm < 1) per_sm = 1;
    const int64_t subs = ((words + kTileWords - 1) / kTileWords) * (kTileWords / kSubWords);
    int64_t grid = (int64_t)sms * per_sm;
    if (grid > subs) grid = subs;
    if (grid < 1) grid = 1;
    SampleArgs a;
    a.hx = hx; a.hz = hz; a.sx = sx; a.sz = sz; a.ex = ex; a.ez = ez;
    a.words = words; a.tail_mask = tail_mask; a.seed = seed; a.first_word = first_word;
    a.thr = thr; a.use_gap = use_gap; a.gap = gap;
    k_sample_syndrome_tiles<<<(unsigned)grid, kSampleThreads, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace qcss
