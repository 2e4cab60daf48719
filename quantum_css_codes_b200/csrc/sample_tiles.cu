// K3 for codes of any size: depolarising Philox sampler fused into the sparse syndrome kernel, so that the
// 2 n error bits per shot of a large code (hypergraph product n = 1600: 400 B/shot) never exist in HBM --
// only the 2 m syndrome bits leave the SM (tile-major, the layout qcss_syndrome_tiles uses).
//
// One CTA works on a sub-tile of 256 shots (8 words) at a time:
//   sample   thread t draws site-words (qubit j, word w), j * 8 + w = t, t + 1024, ... with the K3 sampler
//            (core.cuh: gap sampler below p = 1/64, bit-serial above; Philox counter = (global word, site = j,
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

// p >= 1/64 (bit-serial sampler): most site-words hold errors, so the error words are materialised in shared memory
// and the checks gather them.
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
    if (a.sx != nullptr) stage_csr(a.hx, ptr_x, cols_x);
    if (a.sz != nullptr) stage_csr(a.hz, ptr_z, cols_z);
    if (threadIdx.x < kSubWords) px[n * kSubWords + threadIdx.x] = pz[n * kSubWords + threadIdx.x] = 0u;
    __syncthreads();
    const int64_t tiles = (a.words + kTileWords - 1) / kTileWords;
    const int64_t subs = tiles * (kTileWords / kSubWords);
    for (int64_t st = blockIdx.x; st < subs; st += gridDim.x) {
        const int64_t tile = st / (kTileWords / kSubWords);
        const int sub = (int)(st % (kTileWords / kSubWords));
        const int64_t wbase = tile * kTileWords + sub * kSubWords;
        const int w = threadIdx.x % kSubWords;              // 1024 % 8 == 0: a thread keeps its word column
        const int64_t gw = wbase + w;
        const int total = n * kSubWords;
        for (int idx = threadIdx.x; idx < total; idx += kSampleThreads) {
            uint32_t x = 0u, z = 0u;
            if (gw < a.words) sample_site_word(a.seed, a.first_word + (uint64_t)gw, (uint32_t)(idx / kSubWords), a.thr, x, z);
            if (gw == a.words - 1) { x &= a.tail_mask; z &= a.tail_mask; }
            px[idx] = x;
            pz[idx] = z;
            const int j = idx / kSubWords;
            if (a.ex != nullptr) a.ex[((size_t)tile * n + j) * kTileWords + sub * kSubWords + w] = x;
            if (a.ez != nullptr) a.ez[((size_t)tile * n + j) * kTileWords + sub * kSubWords + w] = z;
        }
        __syncthreads();
        if (a.sx != nullptr) xor_rows(a.hx, ptr_x, cols_x, px, a.sx, tile, sub, a.words, a.tail_mask);
        if (a.sz != nullptr) xor_rows(a.hz, ptr_z, cols_z, pz, a.sz, tile, sub, a.words, a.tail_mask);
        __syncthreads();
    }
}

// p < 1/64 (gap sampler): 97 % of the site-words (p = 1e-3) hold no error, so no error word is ever stored.  Per
// sub-tile of 256 shots:
//   1  first-look Philox block of every eight site-words (core.cuh; two per iteration: independent 10-round chains);
//      the few sites that may hold an error are QUEUED instead of being finished in place -- with ~1 erring lane per
//      warp-instruction, finishing in place makes every warp pay the gap logic for one useful lane;
//   2  the queue is handed out one item per lane: finish the draw (same streams, so the same bits) and SCATTER the error
//      word into the syndrome accumulators of the checks its qubit takes part in (transposed supports in shared memory,
//      ~3.4 checks per qubit and type for the hypergraph-product code: ~2700 shared atomics per sub-tile against the
//      10 752 LDS.128 of a gather over mostly-zero words);
//   3  the accumulators (m x 32 B per type) are written out tile-major and cleared.
// Two block barriers per sub-tile; T threads per CTA, as many CTAs per SM as fit (phases of different CTAs overlap).
template <int T>
__global__ void __launch_bounds__(T, 1024 / T)
k_sample_scatter_tiles(const __grid_constant__ SampleArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int n = a.hx.n, mx = a.hx.m, mz = a.hz.m;
    uint32_t* const acc_x = reinterpret_cast<uint32_t*>(smem);               // [mx][8]
    uint32_t* const acc_z = acc_x + (size_t)mx * kSubWords;                  // [mz][8]
    uint16_t* const cptr_x = reinterpret_cast<uint16_t*>(acc_z + (size_t)mz * kSubWords);
    uint16_t* const cptr_z = cptr_x + ((n + 2) & ~1);
    uint16_t* const rows_x = cptr_z + ((n + 2) & ~1);
    uint16_t* const rows_z = rows_x + ((a.hx.nnz + 1) & ~1);
    uint16_t* const queue = rows_z + ((a.hz.nnz + 1) & ~1);                  // [n * 8] site-words that may hold an error
    __shared__ GapTable s_gap;
    __shared__ int q_count[2];
    const int tid = threadIdx.x;
    if (tid < 2) q_count[tid] = 0;
    if (tid < 32) s_gap.cdf[tid] = a.gap.cdf[tid];
    if (tid == 32) s_gap.inv = a.gap.inv;
    for (int i = tid; i < (mx + mz) * kSubWords; i += T) acc_x[i] = 0u;
    for (int j = tid; j <= n; j += T) {
        cptr_x[j] = (uint16_t)__ldg(a.hx.col_ptr + j);
        cptr_z[j] = (uint16_t)__ldg(a.hz.col_ptr + j);
    }
    for (int k = tid; k < a.hx.nnz; k += T) rows_x[k] = __ldg(a.hx.rows + k);
    for (int k = tid; k < a.hz.nnz; k += T) rows_z[k] = __ldg(a.hz.rows + k);
    __syncthreads();
    const uint32_t cdf31 = s_gap.cdf[31], look_hi = gap_look16(cdf31) << 16;
    Philox ph;
    ph.k0 = (uint32_t)a.seed;
    ph.k1 = (uint32_t)(a.seed >> 32);
    const int64_t tiles = (a.words + kTileWords - 1) / kTileWords;
    const int64_t subs = tiles * (kTileWords / kSubWords);
    const int w = tid % kSubWords;                          // T % 8 == 0: a thread keeps its word column
    const int groups = ((n + 7) / 8) * kSubWords;           // group index gi = (qubit >> 3) * kSubWords + word
    int parity = 0;
    for (int64_t st = blockIdx.x; st < subs; st += gridDim.x, parity ^= 1) {
        const int64_t tile = st / (kTileWords / kSubWords);
        const int sub = (int)(st % (kTileWords / kSubWords));
        const int64_t wbase = tile * kTileWords + sub * kSubWords;
        const int64_t gw = wbase + w;
        int* const qc = &q_count[parity];
        if (a.ex != nullptr || a.ez != nullptr) {            // requested error tiles start as "no error" too
            for (int i = tid; i < n * kSubWords / 4; i += T) {
                const int j = i / (kSubWords / 4), c = i % (kSubWords / 4);
                const size_t off = ((size_t)tile * n + j) * kTileWords + sub * kSubWords + c * 4;
                if (a.ex != nullptr) *reinterpret_cast<uint4*>(a.ex + off) = make_uint4(0u, 0u, 0u, 0u);
                if (a.ez != nullptr) *reinterpret_cast<uint4*>(a.ez + off) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        // ---- 1 ----
        if (gw < a.words) {
            const uint64_t g = a.first_word + (uint64_t)gw;
            const uint32_t g_lo = (uint32_t)g, g_hi = (uint32_t)(g >> 32);
            // the hits of a group collect in a mask, branch-free; one shared atomic per thread with hits
            auto first_look = [&](int gi, const uint32_t (&hb)[4]) {
                const int j0 = 8 * (gi / kSubWords);
                uint32_t m = 0u;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (gap_look(hb, c, look_hi)) m |= 1u << c;
                if (j0 + 8 > n) m &= (1u << (n - j0)) - 1u;
                if (m != 0u) {
                    int at = atomicAdd(qc, (int)popc32(m));
                    for (; m != 0u; m &= m - 1u) queue[at++] = (uint16_t)((j0 + (int)ctz32(m)) * kSubWords + w);
                }
            };
            int gi = tid;
            for (; gi + T < groups; gi += 2 * T) {
                uint32_t b0[4], b1[4];
                gap_first8(ph, g_lo, g_hi, (uint32_t)(gi / kSubWords), b0);
                gap_first8(ph, g_lo, g_hi, (uint32_t)((gi + T) / kSubWords), b1);
                first_look(gi, b0);
                first_look(gi + T, b1);
            }
            if (gi < groups) {
                uint32_t b0[4];
                gap_first8(ph, g_lo, g_hi, (uint32_t)(gi / kSubWords), b0);
                first_look(gi, b0);
            }
        }
        __syncthreads();
        // ---- 2 ----
        const int count = *qc;
        if (tid == 0) q_count[parity ^ 1] = 0;               // the other counter: untouched until the next phase 1
        for (int k = tid; k < count; k += T) {
            const int idx = queue[k], j = idx / kSubWords, qw_ = idx % kSubWords;
            const int64_t qw = wbase + qw_;
            uint32_t x, z;
            sample_site_word_gap(a.seed, a.first_word + (uint64_t)qw, (uint32_t)j, s_gap, cdf31, x, z);
            if (qw == a.words - 1) { x &= a.tail_mask; z &= a.tail_mask; }
            if (a.ex != nullptr) a.ex[((size_t)tile * n + j) * kTileWords + sub * kSubWords + qw_] = x;
            if (a.ez != nullptr) a.ez[((size_t)tile * n + j) * kTileWords + sub * kSubWords + qw_] = z;
            if (x != 0u && a.sx != nullptr)
                for (int r = cptr_x[j]; r < cptr_x[j + 1]; ++r) atomicXor(acc_x + rows_x[r] * kSubWords + qw_, x);
            if (z != 0u && a.sz != nullptr)
                for (int r = cptr_z[j]; r < cptr_z[j + 1]; ++r) atomicXor(acc_z + rows_z[r] * kSubWords + qw_, z);
        }
        __syncthreads();
        // ---- 3 (words past the end of the batch were never sampled and the last word was masked above) ----
        auto flush = [&](uint32_t* acc, int m, uint32_t* out) {
            for (int i = tid; i < m * (kSubWords / 4); i += T) {
                const int row = i / (kSubWords / 4), c = i % (kSubWords / 4);
                uint4* const src = reinterpret_cast<uint4*>(acc + row * kSubWords + c * 4);
                const uint4 v = *src;
                *src = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(out + ((size_t)tile * m + row) * kTileWords + sub * kSubWords + c * 4) = v;
            }
        };
        if (a.sx != nullptr) flush(acc_x, mx, a.sx);
        if (a.sz != nullptr) flush(acc_z, mz, a.sz);
        // (the next scatter comes after the barrier that ends the next phase 1)
    }
}

template <class Kernel>
cudaError_t launch_sampler(Kernel kernel, int threads, size_t smem, const SampleArgs& a, cudaStream_t stream) {
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    int dev = 0, sms = 0;
    if ((err = cudaGetDevice(&dev)) != cudaSuccess) return err;
    if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return err;
    int per_sm = 1;
    if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem)) != cudaSuccess) return err;
    if (per_sm < 1) per_sm = 1;
    const int64_t subs = ((a.words + kTileWords - 1) / kTileWords) * (kTileWords / kSubWords);
    int64_t grid = (int64_t)sms * per_sm;
    if (grid > subs) grid = subs;
    if (grid < 1) grid = 1;
    kernel<<<(unsigned)grid, threads, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace

// cudaErrorInvalidValue when the shared-memory arrays of the chosen form do not fit (gather: the two error arrays,
// n > ~3000; scatter: the two syndrome accumulators).
cudaError_t launch_sample_syndrome_tiles(const SparseRows& hx, const SparseRows& hz, uint32_t* sx, uint32_t* sz,
                                         uint32_t* ex, uint32_t* ez, int64_t words, uint32_t tail_mask, uint64_t seed,
                                         uint64_t first_word, uint32_t thr, uint32_t use_gap, const GapTable& gap,
                                         cudaStream_t stream) {
    if (hx.n != hz.n || hx.n * kSubWords > 65535 || hx.groups > 65535 || hz.groups > 65535 || hx.m > 65534 || hz.m > 65534 ||
        hx.nnz > 65535 || hz.nnz > 65535) return cudaErrorInvalidValue;
    SampleArgs a;
    a.hx = hx; a.hz = hz; a.sx = sx; a.sz = sz; a.ex = ex; a.ez = ez;
    a.words = words; a.tail_mask = tail_mask; a.seed = seed; a.first_word = first_word;
    a.thr = thr; a.use_gap = use_gap; a.gap = gap;
    if (use_gap) {
        const size_t smem = (size_t)(hx.m + hz.m) * kSubWords * 4 + 2 * (size_t)(2 * ((hx.n + 2) & ~1)) +
                            2 * (size_t)(((hx.nnz + 1) & ~1) + ((hz.nnz + 1) & ~1)) + 2 * (size_t)hx.n * kSubWords;
        if (smem > 226 * 1024) return cudaErrorInvalidValue;
        if (2 * smem + 4096 <= 226 * 1024) return launch_sampler(k_sample_scatter_tiles<512>, 512, smem, a, stream);
        return launch_sampler(k_sample_scatter_tiles<1024>, 1024, smem, a, stream);
    }
    const size_t smem = (size_t)2 * (hx.n + 1) * kSubWords * 4 + 8 * (size_t)(hx.groups + hz.groups) +
                        2 * (size_t)(((hx.m + 2) & ~1) + ((hz.m + 2) & ~1));
    if (smem > 226 * 1024) return cudaErrorInvalidValue;
    return launch_sampler(k_sample_syndrome_tiles, kSampleThreads, smem, a, stream);
}

}  // namespace qcss
