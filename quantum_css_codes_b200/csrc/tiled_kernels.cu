// K1 for codes of any size (hypergraph-product codes, n ~ 1600): sparse-row syndrome kernels.
//
// A CTA stages a tile of every error plane in shared memory -- TW consecutive 32-shot words
// (TW*4 bytes) of each of the n planes -- and forms every syndrome row as the XOR of the planes
// its CSR row names (css_code.py:728 with a sparse H).  Each error bit leaves HBM exactly once
// although column j of H feeds several rows; syndrome words go back with 16-byte stores.
//
//  k_syndrome_tma   (main path): persistent CTA per SM, two-stage ring.  The planes are a 2-D
//                   tensor (words x planes); one thread issues cp.async.bulk.tensor (TMA) box
//                   loads of [rows_per_box planes][TW words] that complete on an mbarrier while
//                   the whole CTA XORs the previous tile.  Out-of-range words / planes are
//                   zero-filled by the TMA unit, so ragged tails need no special case on input.
//  k_syndrome_tiled (large n): single-stage tile filled with 16-byte cp.async by all threads;
//                   used when two TMA stages do not fit in shared memory.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>

#include <cstdlib>

#include "launch.h"

namespace qcss {

namespace {

constexpr int kTiledThreads = 512;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
}

// XOR of the planes of every CSR row for one staged tile; thread (slot, q) owns 16 bytes of a row.
template <int TW>
__device__ __forceinline__ void xor_rows(const SparseRows& h, const uint32_t* tile, uint32_t* __restrict__ s,
                                         int64_t s_stride, int64_t w0, int64_t words, uint32_t tail_mask) {
    constexpr int kQ = TW / 4;
    constexpr int kSlots = kTiledThreads / kQ;
    const int q = threadIdx.x % kQ, slot = threadIdx.x / kQ;
    const int64_t wq = w0 + q * 4;
    if (wq >= words) return;
    const bool ragged = wq + 4 > words - 1;          // touches the tail word or runs past it
    for (int i = slot; i < h.m; i += kSlots) {
        const int beg = __ldg(h.row_ptr + i), end = __ldg(h.row_ptr + i + 1);
        uint4 acc = make_uint4(0u, 0u, 0u, 0u);
        for (int k = beg; k < end; ++k) {
            const int j = __ldg(h.cols + k);
            const uint4 v = *reinterpret_cast<const uint4*>(tile + (size_t)j * TW + q * 4);
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
        uint32_t* dst = s + (int64_t)i * s_stride + wq;
        if (!ragged) {
            *reinterpret_cast<uint4*>(dst) = acc;
        } else {
            uint32_t out[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const int64_t w = wq + v;
                if (w >= words) out[v] = 0u;
                else if (w == words - 1) out[v] &= tail_mask;
            }
            if (wq + 4 <= s_stride)
                *reinterpret_cast<uint4*>(dst) = make_uint4(out[0], out[1], out[2], out[3]);
            else
                for (int v = 0; v < 4 && wq + v < s_stride; ++v) dst[v] = out[v];
        }
    }
}

// Same, with the row supports in shared memory in ELL form: ell[i][0..WP) are the plane indices of
// row i (0xFFFF = padding), WP a multiple of 8 so one 16-byte read fetches 8 of them.  All the
// plane reads of a row are independent, so the loop is shared-memory-bandwidth bound.
template <int TW, int WP, int THREADS = kTiledThreads>
__device__ __forceinline__ void xor_rows_ell(int m, const uint16_t* ell, const uint32_t* tile,
                                             uint32_t* __restrict__ s, int64_t s_stride, int64_t w0, int64_t words,
                                             uint32_t tail_mask) {
    constexpr int kQ = TW / 4;
    constexpr int kSlots = THREADS / kQ;
    const int q = threadIdx.x % kQ, slot = threadIdx.x / kQ;
    const int64_t wq = w0 + q * 4;
    if (wq >= words) return;
    const bool ragged = wq + 4 > words - 1;
    for (int i = slot; i < m; i += kSlots) {
        uint4 acc = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int g = 0; g < WP / 8; ++g) {
            const uint4 idx = *reinterpret_cast<const uint4*>(ell + (size_t)i * WP + g * 8);
            const uint32_t iw[4] = {idx.x, idx.y, idx.z, idx.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t j = (iw[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
                if (j != 0xFFFFu) {
                    const uint4 v = *reinterpret_cast<const uint4*>(tile + (size_t)j * TW + q * 4);
                    acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
                }
            }
        }
        uint32_t* dst = s + (int64_t)i * s_stride + wq;
        if (!ragged) {
            *reinterpret_cast<uint4*>(dst) = acc;
        } else {
            uint32_t out[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const int64_t w = wq + v;
                if (w >= words) out[v] = 0u;
                else if (w == words - 1) out[v] &= tail_mask;
            }
            if (wq + 4 <= s_stride)
                *reinterpret_cast<uint4*>(dst) = make_uint4(out[0], out[1], out[2], out[3]);
            else
                for (int v = 0; v < 4 && wq + v < s_stride; ++v) dst[v] = out[v];
        }
    }
}

// ---- TMA two-stage pipeline ---------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(addr),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            (unsigned)__cvta_generic_to_shared(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"((unsigned)__cvta_generic_to_shared(bar))
        : "memory");
}

struct TmaShape {
    int box_rows;     // planes per TMA box (<= 256)
    int boxes;        // boxes per tile; boxes * box_rows >= n
};

// Boxes of equal height covering n planes.  The height is a multiple of 8 so that every box starts
// on a 128-byte boundary of the stage buffer (TMA destination alignment) for any tile width >= 4.
inline TmaShape tma_shape(int n) {
    TmaShape sh;
    sh.boxes = (n + 255) / 256;
    sh.box_rows = (((n + sh.boxes - 1) / sh.boxes) + 7) & ~7;
    return sh;
}

// WP = 0: row supports read from the CSR arrays in global memory; WP = 8 / 16: ELL copy in smem.
template <int TW, int WP>
__global__ void __launch_bounds__(kTiledThreads, 1)
k_syndrome_tma(const __grid_constant__ CUtensorMap map, const CUtensorMap* __restrict__ gmap, int dbg,
               SparseRows h, TmaShape shape, uint32_t* __restrict__ s, int64_t s_stride, int64_t words,
               uint32_t tail_mask) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[2];
    const unsigned stage_words = (unsigned)(shape.boxes * shape.box_rows * TW);
    const unsigned stage_bytes = stage_words * (unsigned)sizeof(uint32_t);
    uint32_t* const stage0 = reinterpret_cast<uint32_t*>(smem_raw);
    uint16_t* const ell = reinterpret_cast<uint16_t*>(smem_raw + 2 * (size_t)stage_bytes);
    const int64_t tiles = (words + TW - 1) / TW;
    const CUtensorMap* mp = (dbg & 2) ? gmap : &map;
    if constexpr (WP > 0) {
        for (int idx = threadIdx.x; idx < h.m * WP; idx += kTiledThreads) {
            const int i = idx / WP, k = idx % WP;
            const int beg = __ldg(h.row_ptr + i), end = __ldg(h.row_ptr + i + 1);
            ell[idx] = (beg + k < end) ? __ldg(h.cols + beg + k) : (uint16_t)0xFFFFu;
        }
    }

    if (threadIdx.x == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (!(dbg & 1)) asm volatile("prefetch.tensormap [%0];" ::"l"(mp) : "memory");
    }
    __syncthreads();

    auto issue = [&](int64_t tile, int st) {
        mbar_expect_tx(&full_bar[st], stage_bytes);
        for (int b = 0; b < shape.boxes; ++b)
            tma_load_2d(stage0 + (size_t)st * stage_words + (size_t)b * shape.box_rows * TW, mp,
                        (int)(tile * TW), b * shape.box_rows, &full_bar[st]);
    };

    // tile -> CTA mapping: interleaved (CTA b takes b, b+grid, ...) or, with dbg bit 16, a contiguous
    // range per CTA (consecutive tiles of one CTA then share 256-byte L2 lines)
    const bool contig = (dbg & 16) != 0;
    const int64_t per_cta = (tiles + gridDim.x - 1) / gridDim.x;
    int64_t t = contig ? (int64_t)blockIdx.x * per_cta : (int64_t)blockIdx.x;
    const int64_t t_end = contig ? (t + per_cta < tiles ? t + per_cta : tiles) : tiles;
    const int64_t t_step = contig ? 1 : (int64_t)gridDim.x;
    int st = 0;
    unsigned parity[2] = {0u, 0u};
    if (threadIdx.x == 0 && t < t_end) issue(t, 0);
    for (; t < t_end; t += t_step) {
        const int64_t next = t + t_step;
        // stage st^1 was drained by everyone before the __syncthreads that ended the last pass
        if (threadIdx.x == 0 && next < t_end) issue(next, st ^ 1);
        mbar_wait(&full_bar[st], parity[st]);
        parity[st] ^= 1u;
        if constexpr (WP > 0)
            xor_rows_ell<TW, WP>(h.m, ell, stage0 + (size_t)st * stage_words, s, s_stride, t * TW, words, tail_mask);
        else
            xor_rows<TW>(h, stage0 + (size_t)st * stage_words, s, s_stride, t * TW, words, tail_mask);
        __syncthreads();
        st ^= 1;
    }
}

// ---- single-stage cp.async variant ---------------------------------------------------------
template <int TW>
__global__ void __launch_bounds__(kTiledThreads)
k_syndrome_tiled(SparseRows h, const uint32_t* __restrict__ e, int64_t e_stride,
                 uint32_t* __restrict__ s, int64_t s_stride, int64_t words, uint32_t tail_mask) {
    extern __shared__ __align__(16) uint32_t tile[];
    constexpr int kQ = TW / 4;                       // 16-byte chunks per plane row
    const int64_t tiles = (words + TW - 1) / TW;
    const int64_t e_chunks = e_stride / 4;           // valid 16-byte chunks per plane

    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int64_t w0 = t * TW;
        for (int idx = threadIdx.x; idx < h.n * kQ; idx += kTiledThreads) {
            const int j = idx / kQ, q = idx % kQ;
            const int64_t chunk = w0 / 4 + q;
            uint32_t* dst = tile + (size_t)j * TW + q * 4;
            if (chunk < e_chunks) cp_async16(dst, e + (int64_t)j * e_stride + chunk * 4);
            else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        xor_rows<TW>(h, tile, s, s_stride, w0, words, tail_mask);
        __syncthreads();
    }
}

// ---- wide single-stage tiles with L2 prefetch (default path) ---------------------------------------
// The widest tile that fits one CTA per SM (TW = 32 words = one 128-byte line per plane for n <= 1640)
// is staged with 16-byte cp.async by all 1024 threads.  While the CTA XORs tile t it has already asked
// L2 to fetch tile t+grid (one prefetch per 128-byte line), so the next staging pass is served from
// L2 and the DRAM latency hides behind the XOR phase.  Measured on B200 (HGP-1600, 5e7 shots): 128-byte
// rows reach 1.5x the bandwidth of the 64-byte-row TMA ring, whose per-row request rate is the limit.
constexpr int kWideThreads = 1024;

template <int TW, int WP>
__global__ void __launch_bounds__(kWideThreads, 1)
k_syndrome_wide(SparseRows h, const uint32_t* __restrict__ e, int64_t e_stride, uint32_t* __restrict__ s,
                int64_t s_stride, int64_t words, uint32_t tail_mask) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint32_t* const tile = reinterpret_cast<uint32_t*>(smem_raw);
    uint16_t* const ell = reinterpret_cast<uint16_t*>(smem_raw + (size_t)h.n * TW * sizeof(uint32_t));
    constexpr int kQ = TW / 4;                       // 16-byte chunks per plane row
    const int64_t tiles = (words + TW - 1) / TW;
    const int64_t e_chunks = e_stride / 4;
    for (int idx = threadIdx.x; idx < h.m * WP; idx += kWideThreads) {
        const int i = idx / WP, k = idx % WP;
        const int beg = __ldg(h.row_ptr + i), end = __ldg(h.row_ptr + i + 1);
        ell[idx] = (beg + k < end) ? __ldg(h.cols + beg + k) : (uint16_t)0xFFFFu;
    }
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int64_t w0 = t * TW;
        for (int idx = threadIdx.x; idx < h.n * kQ; idx += kWideThreads) {
            const int j = idx / kQ, q = idx % kQ;
            const int64_t chunk = w0 / 4 + q;
            uint32_t* dst = tile + (size_t)j * TW + q * 4;
            if (chunk < e_chunks) cp_async16(dst, e + (int64_t)j * e_stride + chunk * 4);
            else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const int64_t next = t + gridDim.x;
        if (next < tiles) {
            constexpr int kLines = (TW * 4 + 127) / 128;         // 128-byte lines per plane row
            const int64_t nw0 = next * TW;
            for (int idx = threadIdx.x; idx < h.n * kLines; idx += kWideThreads) {
                const int j = idx / kLines, l = idx % kLines;
                const int64_t w = nw0 + l * 32;
                if (w < e_stride)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(e + (int64_t)j * e_stride + w));
            }
        }
        xor_rows_ell<TW, WP, kWideThreads>(h.m, ell, tile, s, s_stride, w0, words, tail_mask);
        __syncthreads();
    }
}

template <int TW, int WP>
cudaError_t launch_wide(const SparseRows& h, const uint32_t* e, int64_t e_stride, uint32_t* s, int64_t s_stride,
                        int64_t words, uint32_t tail_mask, cudaStream_t stream);

// ---- plane-split accumulate ring (default path for n x 128 B > one stage) --------------------------
// 128-byte plane rows are what the memory system likes (64-byte rows run at about half the L2 request
// rate), but n = 1600 planes x 128 B is a whole SM's shared memory.  So the planes are split into P
// parts of <= pp planes; a tile is processed as P "part-tiles" that flow through an NST-deep cp.async
// ring (NST - 1 part-tiles in flight while one is XORed, ONE block barrier per part-tile), and every
// thread keeps the partial syndromes of its RPT row chunks in registers across the parts of a tile
// (row i, 16-byte chunk q <-> thread (i % 128) * 8 + q).
//
// The XOR loop touches exactly the support: at kernel start the CSR rows are re-bucketed per part
// into shared memory (poff[part][row] -> pent[], entries = local plane index * 8 = the plane row's
// byte offset >> 4), so a part-tile costs one 128-byte shared-memory wavefront per support entry.
// History (HGP-1600, B200): testing every entry against the part bounds in every part (two-stage
// ring) 59 % of the HBM copy peak with 77 % of the tile time spent issuing instructions; sentinel-
// padded per-part ELL, same ring: 61 % -- fewer instructions but 1.7x the shared-memory wavefronts.
// With the XOR loop switched off (QCSS_RING_DBG=1: loads + stores only) the kernel runs at 4.29 TB/s,
// with the loads switched off (QCSS_RING_DBG=2) at 6.24 TB/s-equivalent: the limit is the memory side
// of n = 1600 separate 128-byte streams per SM, not the XOR work (profiles/r01_hgp_ring_ncu_summary.txt).
constexpr int kSplitTW = 32;                  // words per plane row per tile: one 128-byte line
constexpr int kSplitSlots = kWideThreads / 8; // 128 row slots x 8 chunks

struct SplitShape {
    int parts;        // P
    int pp;           // planes per part
    int nnz;          // support entries (capacity of pent)
    int dbg;          // experiments: 1 = skip the XOR loop, 2 = skip the loads
};

inline size_t split_csr_bytes(int m, SplitShape shape) {
    return (((size_t)shape.parts * m + 2) * sizeof(uint16_t) + (size_t)(shape.nnz + 2) * sizeof(uint16_t) + 15) & ~(size_t)15;
}
inline size_t split_smem_bytes(int m, int nst, SplitShape shape) {
    return (size_t)nst * shape.pp * kSplitTW * sizeof(uint32_t) + split_csr_bytes(m, shape);
}

// Supports bucketed by part (count, scan, fill): poff[part * m + row] .. poff[part * m + row + 1] index pent[],
// whose entries are local plane index * (row bytes / 16).  Shared by the plane-major and tile-major rings.
__device__ __forceinline__ void ring_bucket_supports(const SparseRows& h, int parts, int pp, uint16_t* poff,
                                                     uint16_t* pent, int* wsum) {
    constexpr int TW = kSplitTW;
    const int m = h.m, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) poff[0] = 0;
    for (int i = tid; i < m; i += kWideThreads) {
        int cnt[16];
#pragma unroll
        for (int p = 0; p < 16; ++p) cnt[p] = 0;
        const int beg = __ldg(h.row_ptr + i), end = __ldg(h.row_ptr + i + 1);
        for (int t = beg; t < end; ++t) {
            const int p = __ldg(h.cols + t) / pp;
#pragma unroll
            for (int u = 0; u < 16; ++u) cnt[u] += (u == p);
        }
#pragma unroll
        for (int p = 0; p < 16; ++p)
            if (p < parts) poff[1 + p * m + i] = (uint16_t)cnt[p];
    }
    __syncthreads();
    {
        const int L = parts * m, per = (L + kWideThreads - 1) / kWideThreads;
        const int beg = tid * per < L ? tid * per : L, end = beg + per < L ? beg + per : L;
        int sum = 0;
        for (int x = beg; x < end; ++x) sum += poff[1 + x];
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = wsum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xFFFFFFFFu, w, d);
                if (lane >= d) w += t;
            }
            wsum[lane] = w;
        }
        __syncthreads();
        int run = incl - sum + (warp ? wsum[warp - 1] : 0);
        for (int x = beg; x < end; ++x) {
            run += poff[1 + x];
            poff[1 + x] = (uint16_t)run;
        }
    }
    __syncthreads();
    for (int i = tid; i < m; i += kWideThreads) {
        int fill[16];
#pragma unroll
        for (int p = 0; p < 16; ++p) fill[p] = 0;
        const int beg = __ldg(h.row_ptr + i), end = __ldg(h.row_ptr + i + 1);
        for (int t = beg; t < end; ++t) {
            const int c = __ldg(h.cols + t);
            const int p = c / pp;
            int pos = 0;
#pragma unroll
            for (int u = 0; u < 16; ++u)
                if (u == p) pos = fill[u]++;
            pent[poff[p * m + i] + pos] = (uint16_t)((c - p * pp) * (TW * 4 / 16));
        }
    }

}

template <int RPT, int NST>
__global__ void __launch_bounds__(kWideThreads, 1)
k_syndrome_ring(SparseRows h, SplitShape shape, const uint32_t* __restrict__ e, int64_t e_stride,
                uint32_t* __restrict__ s, int64_t s_stride, int64_t words, uint32_t tail_mask) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ int wsum[32];
    constexpr int TW = kSplitTW, kQ = TW / 4;
    const int pp = shape.pp, parts = shape.parts, m = h.m;
    const uint32_t stage_bytes = (uint32_t)pp * TW * sizeof(uint32_t);
    uint16_t* const poff = reinterpret_cast<uint16_t*>(smem_raw + (size_t)NST * stage_bytes);   // [parts * m + 1]
    uint16_t* const pent = poff + (size_t)parts * m + 2;                                         // [nnz]
    const int64_t tiles = (words + TW - 1) / TW;
    const int64_t e_chunks = e_stride / 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    ring_bucket_supports(h, parts, pp, poff, pent, wsum);

    const int q = tid % kQ, slot = tid / kQ;
    const int64_t my_tiles = (tiles > blockIdx.x) ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t steps = my_tiles * parts;
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);

    // issue: thread (slot, q) copies chunk q of planes slot, slot + 128, ... of part-tile `step`
    auto issue = [&](int64_t step) {
        if (step < steps && shape.dbg != 2) {
            const int64_t t = blockIdx.x + (step / parts) * gridDim.x;
            const int part = (int)(step % parts);
            const int lo = part * pp;
            const int cnt = (h.n - lo) < pp ? (h.n - lo) : pp;
            const uint32_t off0 = (uint32_t)(step % NST) * stage_bytes + (uint32_t)slot * (TW * 4) + q * 16;
            const int64_t chunk = t * (TW / 4) + q;
            if (chunk < e_chunks) {
                const uint32_t* src = e + (int64_t)(lo + slot) * e_stride + chunk * 4;
                const int64_t src_step = (int64_t)kSplitSlots * e_stride;
                uint32_t dst = smem_base + off0;
                for (int j = slot; j < cnt; j += kSplitSlots) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                    src += src_step;
                    dst += kSplitSlots * TW * 4;
                }
            } else {
                for (int j = slot; j < cnt; j += kSplitSlots)
                    *reinterpret_cast<uint4*>(smem_raw + off0 + (size_t)(j - slot) * TW * 4) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    uint4 acc[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) acc[r] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int st = 0; st < NST - 1; ++st) issue(st);
    for (int64_t step = 0; step < steps; ++step) {
        asm volatile("cp.async.wait_group %0;" ::"n"(NST - 2) : "memory");   // part-tile `step` has landed
        __syncthreads();                       // ... for every thread; part-tile step - 1 is consumed (and pent is ready)
        issue(step + NST - 1);                 // refill the slot consumed in the previous step
        const int part = (int)(step % parts);
        const uint8_t* buf = smem_raw + (size_t)(step % NST) * stage_bytes + q * 16;
        const uint16_t* po = poff + (size_t)part * m;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int i = slot + r * kSplitSlots;
            if (i < m && shape.dbg != 1) {
                const int o0 = po[i], o1 = po[i + 1];
                for (int k = o0; k < o1; ++k) {
                    const uint4 v = *reinterpret_cast<const uint4*>(buf + ((uint32_t)pent[k] << 4));
                    acc[r].x ^= v.x; acc[r].y ^= v.y; acc[r].z ^= v.z; acc[r].w ^= v.w;
                }
            }
        }
        if (part == parts - 1) {
            const int64_t t = blockIdx.x + (step / parts) * gridDim.x;
            const int64_t wq = t * TW + q * 4;
            const bool ragged = wq + 4 > words - 1;
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const int i = slot + r * kSplitSlots;
                if (i < m && wq < words) {
                    uint32_t* dst = s + (int64_t)i * s_stride + wq;
                    if (!ragged) {
                        *reinterpret_cast<uint4*>(dst) = acc[r];
                    } else {
                        uint32_t out[4] = {acc[r].x, acc[r].y, acc[r].z, acc[r].w};
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            const int64_t w = wq + v;
                            if (w >= words) out[v] = 0u;
                            else if (w == words - 1) out[v] &= tail_mask;
                        }
                        if (wq + 4 <= s_stride)
                            *reinterpret_cast<uint4*>(dst) = make_uint4(out[0], out[1], out[2], out[3]);
                        else
                            for (int v = 0; v < 4 && wq + v < s_stride; ++v) dst[v] = out[v];
                    }
                }
                acc[r] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int RPT, int NST>
cudaError_t launch_ring(const SparseRows& h, SplitShape shape, const uint32_t* e, int64_t e_stride, uint32_t* s,
                        int64_t s_stride, int64_t words, uint32_t tail_mask, cudaStream_t stream);

// ---- tile-major ring: the same accumulate ring fed by bulk copies ---------------------------------------
// Layout "tiles": a batch is stored as [tile][plane][32 words] -- the n plane rows (128 B each) of one tile
// of 1024 shots are CONTIGUOUS (n x 128 B = 200 KB for n = 1600), padded to whole tiles with zero bits.
// A part-tile is then one contiguous run of pp x 128 B, which a single thread moves with ONE
// cp.async.bulk (the TMA unit's 1-D path) completing on the stage's mbarrier: no per-thread cp.async
// address arithmetic, no 1600 separate 128-byte streams per SM spread over as many DRAM pages -- the two
// things the plane-major ring is bounded by (QCSS_RING_DBG measurements above).  Syndromes are written in
// the same layout, [tile][row][32 words]: the 1024 threads of a CTA store 16 KB contiguous per row group.
// The XOR loop, the per-part support buckets and the register partial sums are the plane-major ring's.
__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

template <int RPT, int NST>
__global__ void __launch_bounds__(kWideThreads, 1)
k_syndrome_tiles(SparseRows h, SplitShape shape, const uint32_t* __restrict__ e, uint32_t* __restrict__ s, int64_t tiles,
                 int64_t words, uint32_t tail_mask) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ int wsum[32];
    __shared__ __align__(8) uint64_t full_bar[NST];
    constexpr int TW = kSplitTW, kQ = TW / 4;
    const int pp = shape.pp, parts = shape.parts, m = h.m, n = h.n;
    const uint32_t stage_bytes = (uint32_t)pp * TW * sizeof(uint32_t);
    uint16_t* const poff = reinterpret_cast<uint16_t*>(smem_raw + (size_t)NST * stage_bytes);
    uint16_t* const pent = poff + (size_t)parts * m + 2;
    const int tid = threadIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int st = 0; st < NST; ++st) mbar_init(&full_bar[st], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    ring_bucket_supports(h, parts, pp, poff, pent, wsum);
    __syncthreads();

    const int q = tid % kQ, slot = tid / kQ;
    const int64_t my_tiles = (tiles > blockIdx.x) ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t steps = my_tiles * parts;
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);

    // one thread, one bulk copy per part-tile
    auto issue = [&](int64_t step) {
        if (tid == 0 && step < steps) {
            const int64_t t = blockIdx.x + (step / parts) * gridDim.x;
            const int part = (int)(step % parts);
            const int lo = part * pp;
            const int cnt = (n - lo) < pp ? (n - lo) : pp;
            const uint32_t bytes = (uint32_t)cnt * TW * sizeof(uint32_t);
            const int st = (int)(step % NST);
            mbar_expect_tx(&full_bar[st], bytes);
            bulk_load_1d(smem_base + (uint32_t)st * stage_bytes, e + ((size_t)t * n + lo) * TW, bytes, &full_bar[st]);
        }
    };

    uint4 acc[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) acc[r] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int st = 0; st < NST - 1; ++st) issue(st);
    for (int64_t step = 0; step < steps; ++step) {
        mbar_wait(&full_bar[step % NST], (unsigned)((step / NST) & 1));      // part-tile `step` has landed
        __syncthreads();                       // part-tile step - 1 is consumed by every thread
        issue(step + NST - 1);                 // refill the slot consumed in the previous step
        const int part = (int)(step % parts);
        const uint8_t* buf = smem_raw + (size_t)(step % NST) * stage_bytes + q * 16;
        const uint16_t* po = poff + (size_t)part * m;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int i = slot + r * kSplitSlots;
            if (i < m && shape.dbg != 1) {
                const int o0 = po[i], o1 = po[i + 1];
                for (int k = o0; k < o1; ++k) {
                    const uint4 v = *reinterpret_cast<const uint4*>(buf + ((uint32_t)pent[k] << 4));
                    acc[r].x ^= v.x; acc[r].y ^= v.y; acc[r].z ^= v.z; acc[r].w ^= v.w;
                }
            }
        }
        if (part == parts - 1) {
            const int64_t t = blockIdx.x + (step / parts) * gridDim.x;
            const int64_t wq = t * TW + q * 4;
            const bool ragged = wq + 4 > words - 1;
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const int i = slot + r * kSplitSlots;
                if (i < m) {
                    uint32_t out[4] = {acc[r].x, acc[r].y, acc[r].z, acc[r].w};
                    if (ragged) {
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            const int64_t w = wq + v;
                            if (w >= words) out[v] = 0u;
                            else if (w == words - 1) out[v] &= tail_mask;
                        }
                    }
                    *reinterpret_cast<uint4*>(s + ((size_t)t * m + i) * TW + q * 4) = make_uint4(out[0], out[1], out[2], out[3]);
                }
                acc[r] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
    }
}

// ---- L1-resident gather (experiment, QCSS_TILED_L1=1) ----------------------------------------------
// No staging at all: thread (row slot, 16-byte chunk q) gathers the chunk of every plane in its rows'
// supports straight from global memory with cached loads and relies on the SM's L1 (configured to its
// maximum, ~200+ KB) to serve the 3-4 rows that share a plane; no barriers, no cp.async bookkeeping.
template <int RPT, int WP>
__global__ void __launch_bounds__(kWideThreads, 1)
k_syndrome_l1(SparseRows h, const uint32_t* __restrict__ e, int64_t e_stride, uint32_t* __restrict__ s,
              int64_t s_stride, int64_t words, uint32_t tail_mask) {
    extern __shared__ __align__(128) uint8_t smem_l1[];
    uint16_t* const ell = reinterpret_cast<uint16_t*>(smem_l1);          // [m][WP]
    constexpr int TW = kSplitTW, kQ = TW / 4;
    for (int idx = threadIdx.x; idx < h.m * WP; idx += kWideThreads) {
        const int i = idx / WP, k = idx % WP;
        const int beg = __ldg(h.row_ptr + i), end = __ldg(h.row_ptr + i + 1);
        ell[idx] = (beg + k < end) ? __ldg(h.cols + beg + k) : (uint16_t)0xFFFFu;
    }
    __syncthreads();
    const int q = threadIdx.x % kQ, slot = threadIdx.x / kQ;
    const int64_t tiles = (words + TW - 1) / TW;
    const int64_t e_chunks = e_stride / 4;
    const uint4* const e4 = reinterpret_cast<const uint4*>(e);
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int64_t chunk = t * kQ + q;
        const bool in = chunk < e_chunks;
        const int64_t wq = t * TW + q * 4;
        const bool ragged = wq + 4 > words - 1;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int i = slot + r * kSplitSlots;
            if (i < h.m) {
                uint4 acc = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                for (int g = 0; g < WP / 8; ++g) {
                    const uint4 idx = *reinterpret_cast<const uint4*>(ell + (size_t)i * WP + g * 8);
                    const uint32_t iw[4] = {idx.x, idx.y, idx.z, idx.w};
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t c = (iw[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
                        if (c != 0xFFFFu && in) {
                            const uint4 v = __ldg(e4 + (int64_t)c * e_chunks + chunk);
                            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
                        }
                    }
                }
                if (wq < words) {
                    uint32_t* dst = s + (int64_t)i * s_stride + wq;
                    if (!ragged) {
                        __stcs(reinterpret_cast<uint4*>(dst), acc);
                    } else {
                        uint32_t out[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            const int64_t w = wq + v;
                            if (w >= words) out[v] = 0u;
                            else if (w == words - 1) out[v] &= tail_mask;
                        }
                        if (wq + 4 <= s_stride)
                            *reinterpret_cast<uint4*>(dst) = make_uint4(out[0], out[1], out[2], out[3]);
                        else
                            for (int v = 0; v < 4 && wq + v < s_stride; ++v) dst[v] = out[v];
                    }
                }
            }
        }
    }
}

template <int RPT, int WP>
cudaError_t launch_l1(const SparseRows& h, const uint32_t* e, int64_t e_stride, uint32_t* s, int64_t s_stride,
                      int64_t words, uint32_t tail_mask, cudaStream_t stream);

// ---- host side --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

cudaError_t device_info(int* sms) {
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    return cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev);
}

template <int TW, int WP>
cudaError_t launch_tma(const SparseRows& h, const uint32_t* e, int64_t e_stride, uint32_t* s, int64_t s_stride,
                       int64_t words, uint32_t tail_mask, cudaStream_t stream) {
    EncodeTiledFn encode = encode_tiled_fn();
    if (encode == nullptr) return cudaErrorNotSupported;
    const TmaShape shape = tma_shape(h.n);
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)e_stride, (cuuint64_t)h.n};
    const cuuint64_t strides[1] = {(cuuint64_t)e_stride * sizeof(uint32_t)};
    const cuuint32_t box[2] = {(cuuint32_t)TW, (cuuint32_t)shape.box_rows};
    const cuuint32_t elem[2] = {1, 1};
    static int promo = -1;
    if (promo < 0) {
        const char* v = getenv("QCSS_TMA_PROMO");
        promo = v ? atoi(v) : 128;
    }
    const CUtensorMapL2promotion l2p = promo >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                      : (promo >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                      : (promo >= 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                                                     : CU_TENSOR_MAP_L2_PROMOTION_NONE));
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint32_t*>(e), dims, strides, box, elem,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2p,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    const size_t smem = 2 * (size_t)shape.boxes * shape.box_rows * TW * sizeof(uint32_t) +
                        (size_t)h.m * WP * sizeof(uint16_t);
    cudaError_t err = cudaFuncSetAttribute(k_syndrome_tma<TW, WP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    int sms = 0;
    if ((err = device_info(&sms)) != cudaSuccess) return err;
    const int64_t tiles = (words + TW - 1) / TW;
    int64_t grid = sms < tiles ? sms : tiles;
    if (grid < 1) grid = 1;
    // debugging knob QCSS_TMA_DBG: bit0 = no descriptor prefetch, bit1 = descriptor read from global memory
    static int dbg = -1;
    static CUtensorMap* d_map = nullptr;
    if (dbg < 0) {
        const char* v = getenv("QCSS_TMA_DBG");
        dbg = v ? atoi(v) : 0;
    }
    if (dbg & 2) {
        if (d_map == nullptr && (err = cudaMalloc((void**)&d_map, sizeof(CUtensorMap))) != cudaSuccess) return err;
        if ((err = cudaMemcpyAsync(d_map, &map, sizeof(CUtensorMap), cudaMemcpyHostToDevice, stream)) != cudaSuccess)
            return err;
    }
    k_syndrome_tma<TW, WP><<<(unsigned)grid, kTiledThreads, smem, stream>>>(map, d_map, dbg, h, shape, s, s_stride,
                                                                          words, tail_mask);
    return cudaGetLastError();
}

template <int TW>
cudaError_t launch_tw(const SparseRows& h, const uint32_t* e, int64_t e_stride, uint32_t* s,
                      int64_t s_stride, int64_t words, uint32_t tail_mask, cudaStream_t stream) {
    const size_t smem = (size_t)h.n * TW * sizeof(uint32_t);
    cudaError_t err = cudaFuncSetAttribute(k_syndrome_tiled<TW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem);
    if (err != cudaSuccess) return err;
    int sms = 0, per_sm = 0;
    if ((err = device_info(&sms)) != cudaSuccess) return err;
    if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_syndrome_tiled<TW>, kTiledThreads,
                                                             smem)) != cudaSuccess) return err;
    if (per_sm < 1) per_sm = 1;
    const int64_t tiles = (words + TW - 1) / TW;
    int64_t grid = (int64_t)sms * per_sm;
    if (grid > tiles) grid = tiles;
    if (grid < 1) grid = 1;
    k_syndrome_tiled<TW><<<(unsigned)grid, kTiledThreads, smem, stream>>>(h, e, e_stride, s, s_stride, words,
                                                                        tail_mask);
    return cudaGetLastError();
}

template <int TW, int WP>
cudaError_t launch_wide(const SparseRows& h, const uint32_t* e, int64_t e_stride, uint32_t* s, int64_t s_stride,
                        int64_t words, uint32_t tail_mask, cudaStream_t stream) {
    const size_t smem = (size_t)h.n * TW * sizeof(uint32_t) + (size_t)h.m * WP * sizeof(uint16_t);
    cudaError_t err = cudaFuncSetAttribute(k_syndrome_wide<TW, WP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem);
    if (err != cudaSuccess) return err;
    int sms = 0;
    if ((err = device_info(&sms)) != cudaSuccess) return err;
    const int64_t tiles = (words + TW - 1) / TW;
    int64_t grid = sms < tiles ? sms : tiles;
    if (grid < 1) grid = 1;
    k_syndrome_wide<TW, WP><<<(unsigned)grid, kWideThreads, smem, stream>>>(h, e, e_stride, s, s_stride, words,
                                                                          tail_mask);
    return cudaGetLastError();
}

template <int RPT, int NST>
cudaError_t launch_ring(const SparseRows& h, SplitShape shape, const uint32_t* e, int64_t e_stride, uint32_t* s,
                        int64_t s_stride, int64_t words, uint32_t tail_mask, cudaStream_t stream) {
    const size_t smem = split_smem_bytes(h.m, NST, shape);
    cudaError_t err = cudaFuncSetAttribute(k_syndrome_ring<RPT, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem);
    if (err != cudaSuccess) return err;
    int sms = 0;
    if ((err = device_info(&sms)) != cudaSuccess) return err;
    const int64_t tiles = (words + kSplitTW - 1) / kSplitTW;
    int64_t grid = sms < tiles ? sms : tiles;
    if (grid < 1) grid = 1;
    k_syndrome_ring<RPT, NST><<<(unsigned)grid, kWideThreads, smem, stream>>>(h, shape, e, e_stride, s, s_stride,
                                                                           words, tail_mask);
    return cudaGetLastError();
}

template <int RPT, int NST>
cudaError_t launch_tiles(const SparseRows& h, SplitShape shape, const uint32_t* e, uint32_t* s, int64_t words,
                         uint32_t tail_mask, cudaStream_t stream) {
    const size_t smem = split_smem_bytes(h.m, NST, shape);
    cudaError_t err = cudaFuncSetAttribute(k_syndrome_tiles<RPT, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem);
    if (err != cudaSuccess) return err;
    int sms = 0;
    if ((err = device_info(&sms)) != cudaSuccess) return err;
    const int64_t tiles = (words + kSplitTW - 1) / kSplitTW;
    int64_t grid = sms < tiles ? sms : tiles;
    if (grid < 1) grid = 1;
    k_syndrome_tiles<RPT, NST><<<(unsigned)grid, kWideThreads, smem, stream>>>(h, shape, e, s, tiles, words, tail_mask);
    return cudaGetLastError();
}

template <int NST>
cudaError_t dispatch_tiles(const SparseRows& h, SplitShape shape, const uint32_t* e, uint32_t* s, int64_t words,
                           uint32_t tail_mask, cudaStream_t stream) {
    const int rpt = (h.m + kSplitSlots - 1) / kSplitSlots;
    if (rpt <= 2) return launch_tiles<2, NST>(h, shape, e, s, words, tail_mask, stream);
    if (rpt <= 4) return launch_tiles<4, NST>(h, shape, e, s, words, tail_mask, stream);
    if (rpt <= 6) return launch_tiles<6, NST>(h, shape, e, s, words, tail_mask, stream);
    return launch_tiles<8, NST>(h, shape, e, s, words, tail_mask, stream);
}

template <int NST>
cudaError_t dispatch_ring(const SparseRows& h, SplitShape shape, const uint32_t* e, int64_t e_stride, uint32_t* s,
                          int64_t s_stride, int64_t words, uint32_t tail_mask, cudaStream_t stream) {
    const int rpt = (h.m + kSplitSlots - 1) / kSplitSlots;
    if (rpt <= 2) return launch_ring<2, NST>(h, shape, e, e_stride, s, s_stride, words, tail_mask, stream);
    if (rpt <= 4) return launch_ring<4, NST>(h, shape, e, e_stride, s, s_stride, words, tail_mask, stream);
    if (rpt <= 6) return launch_ring<6, NST>(h, shape, e, e_stride, s, s_stride, words, tail_mask, stream);
    return launch_ring<8, NST>(h, shape, e, e_stride, s, s_stride, words, tail_mask, stream);
}

template <int RPT, int WP>
cudaError_t launch_l1(const SparseRows& h, const uint32_t* e, int64_t e_stride, uint32_t* s, int64_t s_stride,
                      int64_t words, uint32_t tail_mask, cudaStream_t stream) {
    const size_t smem = (size_t)h.m * WP * sizeof(uint16_t);
    cudaError_t err = cudaFuncSetAttribute(k_syndrome_l1<RPT, WP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(k_syndrome_l1<RPT, WP>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                    cudaSharedmemCarveoutMaxL1)) != cudaSuccess) return err;
    int sms = 0;
    if ((err = device_info(&sms)) != cudaSuccess) return err;
    const int64_t tiles = (words + kSplitTW - 1) / kSplitTW;
    int64_t grid = sms < tiles ? sms : tiles;
    if (grid < 1) grid = 1;
    k_syndrome_l1<RPT, WP><<<(unsigned)grid, kWideThreads, smem, stream>>>(h, e, e_stride, s, s_stride, words, tail_mask);
    return cudaGetLastError();
}

size_t tma_stage_bytes(int n, int tw) {
    const TmaShape sh = tma_shape(n);
    return (size_t)sh.boxes * sh.box_rows * tw * sizeof(uint32_t);
}

}  // namespace

// Tile-major layout (see k_syndrome_tiles): e = [tiles][n][32 words], s = [tiles][m][32 words].
// cudaErrorInvalidValue when the shape does not fit the ring (m > 1024 rows, > 65000 support entries, or no
// (stages, parts) split fits shared memory).
cudaError_t launch_syndrome_tiles(const SparseRows& h, const uint32_t* e, uint32_t* s, int64_t words,
                                  uint32_t tail_mask, cudaStream_t stream) {
    if (h.m > 8 * kSplitSlots || h.nnz > 65000) return cudaErrorInvalidValue;
    const size_t cap = 226 * 1024;
    const char* knob = getenv("QCSS_TILES");               // "NST,parts" (experiments)
    int want_nst = 0, want_parts = 0;
    if (knob != nullptr) sscanf(knob, "%d,%d", &want_nst, &want_parts);
    const int nsts[3] = {2, 3, 4};
    for (int a = 0; a < 3; ++a) {
        const int nst = nsts[a];
        if (want_nst != 0 && nst != want_nst) continue;
        for (int parts = (want_parts ? want_parts : (nst == 2 ? 1 : nst)); parts <= 16; ++parts) {
            const int pp = (h.n + parts - 1) / parts;
            const SplitShape shape{parts, pp, h.nnz, getenv("QCSS_RING_DBG") ? atoi(getenv("QCSS_RING_DBG")) : 0};
            if (pp * 8 <= 0xFFFF && split_smem_bytes(h.m, nst, shape) <= cap) {
                if (nst == 4) return dispatch_tiles<4>(h, shape, e, s, words, tail_mask, stream);
                if (nst == 3) return dispatch_tiles<3>(h, shape, e, s, words, tail_mask, stream);
                return dispatch_tiles<2>(h, shape, e, s, words, tail_mask, stream);
            }
            if (want_parts) break;
        }
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_syndrome_tiled(const SparseRows& h, const uint32_t* e, int64_t e_stride, uint32_t* s,
                                  int64_t s_stride, int64_t words, uint32_t tail_mask,
                                  cudaStream_t stream) {
    if (getenv("QCSS_TILED_L1") != nullptr && h.max_row_weight <= 8 && h.m <= 8 * kSplitSlots) {
        const int rpt = (h.m + kSplitSlots - 1) / kSplitSlots;
        if (rpt <= 2) return launch_l1<2, 8>(h, e, e_stride, s, s_stride, words, tail_mask, stream);
        if (rpt <= 4) return launch_l1<4, 8>(h, e, e_stride, s, s_stride, words, tail_mask, stream);
        if (rpt <= 6) return launch_l1<6, 8>(h, e, e_stride, s, s_stride, words, tail_mask, stream);
        return launch_l1<8, 8>(h, e, e_stride, s, s_stride, words, tail_mask, stream);
    }
    // default: plane-split accumulate ring (128-byte rows, NST-deep cp.async ring, partial sums in registers)
    if (getenv("QCSS_TILED_TMA") == nullptr && getenv("QCSS_TILED_NO_TMA") == nullptr &&
        getenv("QCSS_TILED_WIDE") == nullptr && h.m <= 8 * kSplitSlots && h.nnz <= 65000) {
        const size_t cap = 226 * 1024;
        const char* knob = getenv("QCSS_RING");            // "NST,parts" (experiments)
        int want_nst = 0, want_parts = 0;
        if (knob != nullptr) sscanf(knob, "%d,%d", &want_nst, &want_parts);
        // Fewest, largest part-tiles win: measured on HGP-1600 (5e7 shots) NST,parts = 2,2: 4.03 TB/s,
        // 3,3: 3.71, 4,4: 3.50, 4,8: 2.49 -- the per-part-tile barrier and row bookkeeping cost more than
        // the extra part-tiles in flight buy.  Bulk L2 prefetch of the next round in per-plane runs
        // (cp.async.bulk.prefetch.L2) was tried and lost 15 %.
        const int nsts[3] = {2, 3, 4};
        for (int a = 0; a < 3; ++a) {
            const int nst = nsts[a];
            if (want_nst != 0 && nst != want_nst) continue;
            for (int parts = (want_parts ? want_parts : nst); parts <= 16; ++parts) {
                const int pp = (h.n + parts - 1) / parts;
                const SplitShape shape{parts, pp, h.nnz, getenv("QCSS_RING_DBG") ? atoi(getenv("QCSS_RING_DBG")) : 0};
                if (pp * 8 <= 0xFFFF && split_smem_bytes(h.m, nst, shape) <= cap) {
                    if (nst == 4) return dispatch_ring<4>(h, shape, e, e_stride, s, s_stride, words, tail_mask, stream);
                    if (nst == 3) return dispatch_ring<3>(h, shape, e, e_stride, s, s_stride, words, tail_mask, stream);
                    return dispatch_ring<2>(h, shape, e, e_stride, s, s_stride, words, tail_mask, stream);
                }
                if (want_parts) break;
            }
        }
    }
    // widest single-stage tile with the ELL supports next to it in shared memory
    if (getenv("QCSS_TILED_TMA") == nullptr && getenv("QCSS_TILED_NO_TMA") == nullptr) {
        const size_t cap = 226 * 1024;
        const int wp = h.max_row_weight <= 8 ? 8 : (h.max_row_weight <= 16 ? 16 : 0);
        if (wp != 0) {
            const size_t ell_bytes = (size_t)h.m * wp * 2;
#define QCSS_WIDE_CASE(TWV)                                                                                   \
    if ((size_t)h.n * TWV * 4 + ell_bytes <= cap) {                                                           \
        if (wp == 8) return launch_wide<TWV, 8>(h, e, e_stride, s, s_stride, words, tail_mask, stream);        \
        return launch_wide<TWV, 16>(h, e, e_stride, s, s_stride, words, tail_mask, stream);                    \
    }
            QCSS_WIDE_CASE(32)
            QCSS_WIDE_CASE(16)
            QCSS_WIDE_CASE(8)
            QCSS_WIDE_CASE(4)
#undef QCSS_WIDE_CASE
        }
    }
    const size_t budget = 220 * 1024;                 // dynamic shared memory for the two TMA stages
    if (getenv("QCSS_TILED_NO_TMA") != nullptr) {
        if (atoi(getenv("QCSS_TILED_NO_TMA")) == 32 && (size_t)h.n * 32 * 4 <= 220 * 1024)
            return launch_tw<32>(h, e, e_stride, s, s_stride, words, tail_mask, stream);
        if ((size_t)h.n * 16 * 4 <= 100 * 1024) return launch_tw<16>(h, e, e_stride, s, s_stride, words, tail_mask, stream);
        return launch_tw<4>(h, e, e_stride, s, s_stride, words, tail_mask, stream);
    }
    // ELL copy of the row supports in shared memory when it fits next to the two stages
    const size_t total = 226 * 1024;
    auto ell_fits = [&](int tw, int wp) {
        return h.max_row_weight <= wp && 2 * tma_stage_bytes(h.n, tw) + (size_t)h.m * wp * 2 <= total;
    };
#define QCSS_TMA_CASE(TWV)                                                                                     \
    if (2 * tma_stage_bytes(h.n, TWV) <= budget) {                                                             \
        if (ell_fits(TWV, 8)) return launch_tma<TWV, 8>(h, e, e_stride, s, s_stride, words, tail_mask, stream);   \
        if (ell_fits(TWV, 16)) return launch_tma<TWV, 16>(h, e, e_stride, s, s_stride, words, tail_mask, stream); \
        return launch_tma<TWV, 0>(h, e, e_stride, s, s_stride, words, tail_mask, stream);                       \
    }
    QCSS_TMA_CASE(16)
    QCSS_TMA_CASE(8)
    QCSS_TMA_CASE(4)
#undef QCSS_TMA_CASE
    if ((size_t)h.n * 4 * 4 <= 200 * 1024) return launch_tw<4>(h, e, e_stride, s, s_stride, words, tail_mask, stream);
    return cudaErrorInvalidValue;
}

}  // namespace qcss
