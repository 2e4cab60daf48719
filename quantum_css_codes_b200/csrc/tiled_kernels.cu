// K1 for codes of any size (hypergraph-product codes, n ~ 1600): sparse-row syndrome kernels.
//
// A CTA stages a tile of every error plane in shared memory and forms every syndrome row as the XOR of
// the planes its CSR row names (css_code.py:728 with a sparse H).  Each error bit leaves HBM exactly
// once although column j of H feeds several rows; syndrome words go back with 16-byte stores.
//
//  k_syndrome_tiles  tile-major batches [tile of 1024 shots][plane][128 B]: one cp.async.bulk per part-tile,
//                    NST-deep mbarrier ring, partial syndromes in registers (90-100 % of the HBM copy peak).
//  k_syndrome_ring   plane-major batches: the same accumulate ring fed by per-thread 16-byte cp.async.
//  k_syndrome_tiled  any shape the rings do not cover (m > 1024 rows, > 65000 support entries): single-stage
//                    cp.async tile, CSR rows read from global memory.
// The superseded generations (TMA box ring, wide single stage, L1 gather) and every timing knob live in
// tools/experiments/tiled_variants.inc, compiled only with -DQCSS_EXPERIMENTS; this library reads no
// environment variable.
#include <cuda_runtime.h>

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

// ---- mbarrier helpers (tile-major ring) ----------------------------------------------------------
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

// ---- single-stage cp.async tile (general fallback) -------------------------------------------------
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

constexpr int kWideThreads = 1024;

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
// With the XOR loop compiled out (loads + stores only) the kernel ran at 4.29 TB/s, with the loads compiled
// out at 6.24 TB/s-equivalent (round-1 timing builds): the limit is the memory side of n = 1600 separate
// 128-byte streams per SM, not the XOR work (profiles/r01_hgp_ring_ncu_summary.txt).
constexpr int kSplitTW = 32;                  // words per plane row per tile: one 128-byte line
constexpr int kSplitSlots = kWideThreads / 8; // 128 row slots x 8 chunks

struct SplitShape {
    int parts;        // P
    int pp;           // planes per part
    int nnz;          // support entries (capacity of pent)
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
        if (step < steps) {
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
            if (i < m) {
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
// things the plane-major ring is bounded by (measurements above).  Syndromes are written in
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
            if (i < m) {
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

// ---- host side --------------------------------------------------------------------------------
cudaError_t device_info(int* sms) {
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    return cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev);
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

#ifdef QCSS_EXPERIMENTS
#include "../../tools/experiments/tiled_variants.inc"
#endif

}  // namespace

// Tile-major layout (see k_syndrome_tiles): e = [tiles][n][32 words], s = [tiles][m][32 words].
// cudaErrorInvalidValue when the shape does not fit the ring (m > 1024 rows, > 65000 support entries, or no
// (stages, parts) split fits shared memory).  Fewest, largest part-tiles win (HGP-1600: NST,parts = 2,2 runs at
// 6.2 TB/s, 3,3 at 5.1-5.4, 4,4 at 4.4-4.8): the first shape that fits is taken in that order.
cudaError_t launch_syndrome_tiles(const SparseRows& h, const uint32_t* e, uint32_t* s, int64_t words,
                                  uint32_t tail_mask, cudaStream_t stream) {
    if (h.m > 8 * kSplitSlots || h.nnz > 65000) return cudaErrorInvalidValue;
    const size_t cap = 226 * 1024;
    const int nsts[3] = {2, 3, 4};
    for (int a = 0; a < 3; ++a) {
        const int nst = nsts[a];
        for (int parts = (nst == 2 ? 1 : nst); parts <= 16; ++parts) {
            const int pp = (h.n + parts - 1) / parts;
            const SplitShape shape{parts, pp, h.nnz};
            if (pp * 8 <= 0xFFFF && split_smem_bytes(h.m, nst, shape) <= cap) {
                if (nst == 4) return dispatch_tiles<4>(h, shape, e, s, words, tail_mask, stream);
                if (nst == 3) return dispatch_tiles<3>(h, shape, e, s, words, tail_mask, stream);
                return dispatch_tiles<2>(h, shape, e, s, words, tail_mask, stream);
            }
        }
    }
    return cudaErrorInvalidValue;
}

bool syndrome_tiles_supported(const SparseRows& h) {
    if (h.m > 8 * kSplitSlots || h.nnz > 65000) return false;
    for (int nst = 2; nst <= 4; ++nst)
        for (int parts = (nst == 2 ? 1 : nst); parts <= 16; ++parts) {
            const int pp = (h.n + parts - 1) / parts;
            if (pp * 8 <= 0xFFFF && split_smem_bytes(h.m, nst, SplitShape{parts, pp, h.nnz}) <= 226 * 1024) return true;
        }
    return false;
}

cudaError_t launch_syndrome_tiled(const SparseRows& h, const uint32_t* e, int64_t e_stride, uint32_t* s,
                                  int64_t s_stride, int64_t words, uint32_t tail_mask,
                                  cudaStream_t stream) {
#ifdef QCSS_EXPERIMENTS
    {
        cudaError_t err = cudaSuccess;
        if (experiment_dispatch(h, e, e_stride, s, s_stride, words, tail_mask, stream, &err)) return err;
    }
#endif
    // plane-split accumulate ring (128-byte rows, NST-deep cp.async ring, partial sums in registers).  Fewest,
    // largest part-tiles win: measured on HGP-1600 (5e7 shots) NST,parts = 2,2: 4.03 TB/s, 3,3: 3.71, 4,4: 3.50,
    // 4,8: 2.49 -- the per-part-tile barrier and row bookkeeping cost more than the extra part-tiles in flight buy.
    if (h.m <= 8 * kSplitSlots && h.nnz <= 65000) {
        const size_t cap = 226 * 1024;
        const int nsts[3] = {2, 3, 4};
        for (int a = 0; a < 3; ++a) {
            const int nst = nsts[a];
            for (int parts = nst; parts <= 16; ++parts) {
                const int pp = (h.n + parts - 1) / parts;
                const SplitShape shape{parts, pp, h.nnz};
                if (pp * 8 <= 0xFFFF && split_smem_bytes(h.m, nst, shape) <= cap) {
                    if (nst == 4) return dispatch_ring<4>(h, shape, e, e_stride, s, s_stride, words, tail_mask, stream);
                    if (nst == 3) return dispatch_ring<3>(h, shape, e, e_stride, s, s_stride, words, tail_mask, stream);
                    return dispatch_ring<2>(h, shape, e, e_stride, s, s_stride, words, tail_mask, stream);
                }
            }
        }
    }
    // any other shape: single-stage tile, CSR rows from global memory
    if ((size_t)h.n * 32 * 4 <= 220 * 1024) return launch_tw<32>(h, e, e_stride, s, s_stride, words, tail_mask, stream);
    if ((size_t)h.n * 16 * 4 <= 220 * 1024) return launch_tw<16>(h, e, e_stride, s, s_stride, words, tail_mask, stream);
    if ((size_t)h.n * 4 * 4 <= 220 * 1024) return launch_tw<4>(h, e, e_stride, s, s_stride, words, tail_mask, stream);
    return cudaErrorInvalidValue;
}

}  // namespace qcss
