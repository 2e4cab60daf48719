// Host data formats either side of the hot path (the reference passes numpy arrays by reference).
//
//  k_pack_shots     the reference's own layout -- a C-contiguous (shots, n) array of 0/1 elements, uint8 or the
//                   int64 of numpy dtype='int' (css_code.py:39-40) -- to bit planes, ON THE DEVICE: a CTA stages
//                   1024 shots x JC qubits in shared memory with coalesced loads and forms every 32-shot plane
//                   word with one ballot.  Replaces the numpy transpose + packbits the Python layer used to run
//                   on one host core before every call.
//  k_unpack_planes  the inverse for results: planes -> (shots, m) bytes (syndromes, corrections, flags).
//  k_decode_events  SPARSE batches: a list of (shot, qubit, Pauli) events sorted by shot.  At p = 1e-3 a Steane
//                   batch is 0.056 bytes per shot in this form against 1.75 as bit planes, and error-free shots
//                   need no work at all (zero syndrome -> table entry of key 0), so the kernel is event driven:
//                   the first event of each shot XORs the big-endian column keys of that shot's events
//                   (bin_matrix.py:36-43 key order), reads the same flip / miss table byte the dense kernels use
//                   and tallies.  Exactly the tallies of the plane path on the same batch.
//  k_events_*       bit planes -> the sorted event list (count, scan, fill): how a resident or freshly sampled
//                   batch is handed to host code in sparse form.
#include <cuda_runtime.h>

#include "launch.h"

namespace qcss {

namespace {

constexpr int kFmtThreads = 256;
constexpr int kBlockShots = 1024;          // shots per CTA iteration: one 128-byte row per plane
constexpr int kFlatMaxN = 128;             // n <= 128: the 1024-shot block is one contiguous run of the input
constexpr int kSegCols = 64;               // wider arrays: 64 columns per pass, rows padded to 68 bytes in smem
constexpr int kSegPitch = kSegCols + 4;

// ---- (shots, n) elements -> planes -------------------------------------------------------------------
// grid-stride over (block of 1024 shots, column chunk).  Word g of column j of block b goes to
// planes[j * pstride + b * bstride + g]: plane-major (pstride = words per plane, bstride = 32, limit = words per
// plane) or tile-major [block][column][32 words] (pstride = 32, bstride = 32 n, limit = 32).
template <int EB>
__global__ void __launch_bounds__(kFmtThreads)
k_pack_shots(const uint8_t* __restrict__ src, int n, int64_t shots, uint32_t* __restrict__ planes, int64_t pstride,
             int64_t bstride, int64_t limit) {
    extern __shared__ __align__(16) uint8_t fsm[];
    const bool flat = n <= kFlatMaxN;
    const int jc = flat ? n : kSegCols, pitch = flat ? n : kSegPitch;
    const int chunks = flat ? 1 : (n + kSegCols - 1) / kSegCols;
    uint8_t* const bits = fsm;                                                     // [1024][pitch] 0/1 bytes
    uint32_t* const outw = reinterpret_cast<uint32_t*>(fsm + (((size_t)kBlockShots * pitch + 15) & ~(size_t)15));   // [jc][32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t blocks = (shots + kBlockShots - 1) / kBlockShots;
    for (int64_t item = blockIdx.x; item < blocks * chunks; item += gridDim.x) {
        const int64_t blk = item / chunks;
        const int j0 = (int)(item % chunks) * kSegCols;
        const int cols = (n - j0) < jc ? (n - j0) : jc;
        const int64_t s0 = blk * kBlockShots;
        const int rows = (int)((shots - s0) < kBlockShots ? (shots - s0) : kBlockShots);
        if (flat) {
            const int64_t cnt = (int64_t)rows * n;
            const uint8_t* base = src + (size_t)s0 * n * EB;
            if constexpr (EB == 1) {
                const int64_t bulk = cnt & ~(int64_t)15;
                for (int64_t i = (int64_t)tid * 16; i < bulk; i += kFmtThreads * 16) {
                    uint4 v = __ldcs(reinterpret_cast<const uint4*>(base + i));
                    v.x &= 0x01010101u; v.y &= 0x01010101u; v.z &= 0x01010101u; v.w &= 0x01010101u;
                    *reinterpret_cast<uint4*>(bits + i) = v;
                }
                for (int64_t i = bulk + tid; i < cnt; i += kFmtThreads) bits[i] = base[i] & 1;
            } else {
                const unsigned long long* b64 = reinterpret_cast<const unsigned long long*>(base);
                for (int64_t i = tid; i < cnt; i += kFmtThreads) bits[i] = (uint8_t)(__ldcs(b64 + i) & 1ull);
            }
        } else {
            for (int r = warp; r < rows; r += kFmtThreads / 32) {
                const uint8_t* row = src + ((size_t)(s0 + r) * n + j0) * EB;
                for (int c = lane; c < cols; c += 32) {
                    uint8_t b;
                    if constexpr (EB == 1) b = row[c] & 1;
                    else b = (uint8_t)(reinterpret_cast<const unsigned long long*>(row)[c] & 1ull);
                    bits[(size_t)r * pitch + c] = b;
                }
            }
        }
        __syncthreads();
        // warp w forms the words of shot groups w, w + 8, ...: lane = shot, one ballot per (group, column)
        for (int g = warp; g < kBlockShots / 32; g += kFmtThreads / 32) {
            const int r = g * 32 + lane;
            const bool live = r < rows;
            const uint8_t* rowp = bits + (size_t)r * pitch;
            for (int c = 0; c < cols; ++c) {
                const unsigned word = __ballot_sync(0xFFFFFFFFu, live && rowp[c]);
                if (lane == 0) outw[c * 32 + g] = word;
            }
        }
        __syncthreads();
        const int64_t w0 = blk * bstride, wl = (limit == 32 && pstride == 32) ? 0 : blk * 32;
        for (int i = tid; i < cols * 32; i += kFmtThreads) {
            const int c = i >> 5, g = i & 31;
            if (wl + g < limit) planes[(int64_t)(j0 + c) * pstride + w0 + g] = outw[i];
        }
        __syncthreads();
    }
}

// ---- planes -> (shots, m) bytes ------------------------------------------------------------------------
__global__ void __launch_bounds__(kFmtThreads)
k_unpack_planes(const uint32_t* __restrict__ planes, int64_t pstride, int64_t bstride, int64_t limit, int m, int64_t shots,
                uint8_t* __restrict__ dst) {
    extern __shared__ __align__(16) uint32_t usm[];                                 // [jc][33]
    const bool flat = m <= kFlatMaxN;
    const int jc = flat ? m : kSegCols;
    const int chunks = flat ? 1 : (m + kSegCols - 1) / kSegCols;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t blocks = (shots + kBlockShots - 1) / kBlockShots;
    for (int64_t item = blockIdx.x; item < blocks * chunks; item += gridDim.x) {
        const int64_t blk = item / chunks;
        const int j0 = (int)(item % chunks) * kSegCols;
        const int cols = (m - j0) < jc ? (m - j0) : jc;
        const int64_t s0 = blk * kBlockShots, w0 = blk * bstride, wl = (limit == 32 && pstride == 32) ? 0 : blk * 32;
        const int rows = (int)((shots - s0) < kBlockShots ? (shots - s0) : kBlockShots);
        for (int i = tid; i < cols * 32; i += kFmtThreads) {
            const int c = i >> 5, g = i & 31;
            usm[c * 33 + g] = (wl + g < limit) ? __ldcs(planes + (int64_t)(j0 + c) * pstride + w0 + g) : 0u;
        }
        __syncthreads();
        if (flat) {
            // the block's bytes are one contiguous run of rows * m; 4 bytes per thread per step
            uint8_t* base = dst + (size_t)s0 * m;
            const int64_t cnt = (int64_t)rows * m, bulk = cnt & ~(int64_t)3;
            for (int64_t i = (int64_t)tid * 4; i < bulk; i += kFmtThreads * 4) {
                uint32_t v = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int r = (int)((i + k) / m), c = (int)((i + k) - (int64_t)r * m);
                    v |= ((usm[c * 33 + (r >> 5)] >> (r & 31)) & 1u) << (8 * k);
                }
                *reinterpret_cast<uint32_t*>(base + i) = v;
            }
            for (int64_t i = bulk + tid; i < cnt; i += kFmtThreads) {
                const int r = (int)(i / m), c = (int)(i - (int64_t)r * m);
                base[i] = (uint8_t)((usm[c * 33 + (r >> 5)] >> (r & 31)) & 1u);
            }
        } else {
            for (int r = warp; r < rows; r += kFmtThreads / 32) {
                uint8_t* row = dst + (size_t)(s0 + r) * m + j0;
                for (int c = lane; c < cols; c += 32) row[c] = (uint8_t)((usm[c * 33 + (r >> 5)] >> (r & 31)) & 1u);
            }
        }
        __syncthreads();
    }
}

// ---- sparse events ---------------------------------------------------------------------------------
// event = shot << 18 | qubit << 2 | pauli, pauli bit 0 = X component, bit 1 = Z component (1 X, 2 Z, 3 Y)
struct EventTables {
    uint32_t colkey_x[kMaxN];      // big-endian key contribution of column j of parity_check_c2 (X errors)
    uint32_t colkey_z[kMaxN];      // ... of parity_check_c1 (Z errors)
    uint32_t lmask_x, lmask_z;     // Lz / Lx rows as bit masks
    const uint8_t* fm_x;           // [2^m2] bit0 = L.correction, bit1 = miss
    const uint8_t* fm_z;
    int n;
};

// aux[0] = shots that own at least one event, aux[1] = error flag (1 bad field, 2 not sorted)
__global__ void __launch_bounds__(kFmtThreads)
k_decode_events(EventTables t, const unsigned long long* __restrict__ ev, int64_t count, int64_t shots,
                unsigned long long* __restrict__ tally, unsigned long long* __restrict__ aux) {
    uint32_t c[6] = {0u, 0u, 0u, 0u, 0u, 0u};          // fail_x, fail_z, fail_any, miss_x, miss_z, event shots
    uint32_t bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * kFmtThreads + threadIdx.x; i < count; i += (int64_t)gridDim.x * kFmtThreads) {
        const unsigned long long e = ev[i];
        const unsigned long long shot = e >> 18;
        if (i > 0) {
            const unsigned long long prev = ev[i - 1] >> 18;
            if (prev == shot) continue;                 // not the first event of its shot
            if (prev > shot) bad |= 2u;
        }
        if (shot >= (unsigned long long)shots) bad |= 1u;
        uint32_t kx = 0, kz = 0, mx = 0, mz = 0;
        for (int64_t k = i; k < count; ++k) {
            const unsigned long long f = ev[k];
            if ((f >> 18) != shot) break;
            const uint32_t q = (uint32_t)(f >> 2) & 0xFFFFu, p = (uint32_t)f & 3u;
            if (q >= (uint32_t)t.n || p == 0u) { bad |= 1u; continue; }
            if (p & 1u) { kx ^= t.colkey_x[q]; mx ^= 1u << q; }
            if (p & 2u) { kz ^= t.colkey_z[q]; mz ^= 1u << q; }
        }
        const uint32_t fx = t.fm_x[kx], fz = t.fm_z[kz];
        const uint32_t flip_x = (__popc(mx & t.lmask_x) ^ fx) & 1u, flip_z = (__popc(mz & t.lmask_z) ^ fz) & 1u;
        c[0] += flip_x; c[1] += flip_z; c[2] += flip_x | flip_z;
        c[3] += (fx >> 1) & 1u; c[4] += (fz >> 1) & 1u; c[5] += 1u;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) c[k] = __reduce_add_sync(0xFFFFFFFFu, c[k]);
    bad = __reduce_or_sync(0xFFFFFFFFu, bad);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 5; ++k)
            if (c[k]) atomicAdd(tally + 1 + k, (unsigned long long)c[k]);
        if (c[5]) atomicAdd(aux, (unsigned long long)c[5]);
        if (bad) atomicOr(aux + 1, (unsigned long long)bad);
    }
}

// planes -> events.  Pass 1: events per CTA range; pass 2 (one CTA): exclusive scan of the CTA totals; pass 3: fill.
// One thread owns one 32-shot word of all 2 n planes, so the events leave sorted by shot, then qubit.
__device__ __forceinline__ uint32_t word_any(const uint32_t* __restrict__ ex, const uint32_t* __restrict__ ez, int n,
                                             int64_t stride32, int64_t w, uint32_t mask, uint32_t* cnt) {
    uint32_t any = 0, total = 0;
    for (int j = 0; j < n; ++j) {
        const uint32_t v = (__ldg(ex + (int64_t)j * stride32 + w) | __ldg(ez + (int64_t)j * stride32 + w)) & mask;
        any |= v;
        total += __popc(v);
    }
    *cnt = total;
    return any;
}

__global__ void __launch_bounds__(kFmtThreads)
k_events_count(const uint32_t* __restrict__ ex, const uint32_t* __restrict__ ez, int n, int64_t stride32, int64_t words,
               uint32_t tail_mask, int64_t per_cta, unsigned long long* __restrict__ cta_total) {
    __shared__ unsigned long long part[kFmtThreads / 32];
    const int64_t lo = (int64_t)blockIdx.x * per_cta, hi = (lo + per_cta < words) ? lo + per_cta : words;
    unsigned long long sum = 0;
    for (int64_t w = lo + threadIdx.x; w < hi; w += kFmtThreads) {
        uint32_t cnt;
        word_any(ex, ez, n, stride32, w, w == words - 1 ? tail_mask : 0xFFFFFFFFu, &cnt);
        sum += cnt;
    }
    for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
        for (int k = 0; k < kFmtThreads / 32; ++k) s += part[k];
        cta_total[blockIdx.x] = s;
    }
}

__global__ void k_events_scan(unsigned long long* cta_total, int ctas, unsigned long long* total_out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long run = 0;
        for (int b = 0; b < ctas; ++b) {
            const unsigned long long v = cta_total[b];
            cta_total[b] = run;
            run += v;
        }
        *total_out = run;
    }
}

__global__ void __launch_bounds__(kFmtThreads)
k_events_fill(const uint32_t* __restrict__ ex, const uint32_t* __restrict__ ez, int n, int64_t stride32, int64_t words,
              uint32_t tail_mask, int64_t per_cta, const unsigned long long* __restrict__ cta_base, int64_t first_shot,
              unsigned long long* __restrict__ out, int64_t capacity) {
    __shared__ uint32_t wsum[kFmtThreads / 32];
    __shared__ unsigned long long run_base;
    const int64_t lo = (int64_t)blockIdx.x * per_cta, hi = (lo + per_cta < words) ? lo + per_cta : words;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) run_base = cta_base[blockIdx.x];
    __syncthreads();
    for (int64_t w0 = lo; w0 < hi; w0 += kFmtThreads) {
        const int64_t w = w0 + threadIdx.x;
        uint32_t cnt = 0, any = 0;
        const uint32_t mask = (w == words - 1) ? tail_mask : 0xFFFFFFFFu;
        if (w < hi) any = word_any(ex, ez, n, stride32, w, mask, &cnt);
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (int k = 0; k < kFmtThreads / 32; ++k) {
            if (k < warp) before += wsum[k];
            total += wsum[k];
        }
        unsigned long long pos = run_base + before + incl - cnt;
        while (any) {
            const int b = __ffs(any) - 1;
            any &= any - 1;
            const unsigned long long shot = (unsigned long long)(first_shot + w * 32 + b);
            for (int j = 0; j < n; ++j) {
                const uint32_t x = (__ldg(ex + (int64_t)j * stride32 + w) >> b) & 1u;
                const uint32_t z = (__ldg(ez + (int64_t)j * stride32 + w) >> b) & 1u;
                if (x | z) {
                    if ((int64_t)pos < capacity) out[pos] = (shot << 18) | ((unsigned long long)j << 2) | (x | (z << 1));
                    ++pos;
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) run_base += total;
        __syncthreads();
    }
}

int fmt_grid(int64_t items) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t g = (int64_t)sms * 4;
    if (g > items) g = items;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace

// stride32 > 0: plane-major with that many 32-bit words per plane; stride32 == 0: tile-major [tile][column][32 words]
cudaError_t launch_pack_shots(const void* d_src, int elem_bytes, int n, int64_t shots, uint32_t* d_planes, int64_t stride32,
                              cudaStream_t stream) {
    if (shots <= 0 || n <= 0) return cudaSuccess;
    const int64_t pstride = stride32 ? stride32 : 32, bstride = stride32 ? 32 : (int64_t)32 * n, limit = stride32 ? stride32 : 32;
    const bool flat = n <= kFlatMaxN;
    const int pitch = flat ? n : kSegPitch, jc = flat ? n : kSegCols;
    const size_t smem = (((size_t)kBlockShots * pitch + 15) & ~(size_t)15) + (size_t)jc * 32 * 4;
    const int64_t items = ((shots + kBlockShots - 1) / kBlockShots) * (flat ? 1 : (n + kSegCols - 1) / kSegCols);
    cudaError_t err;
    if (elem_bytes == 1) {
        if ((err = cudaFuncSetAttribute(k_pack_shots<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
        k_pack_shots<1><<<fmt_grid(items), kFmtThreads, smem, stream>>>((const uint8_t*)d_src, n, shots, d_planes, pstride, bstride, limit);
    } else if (elem_bytes == 8) {
        if ((err = cudaFuncSetAttribute(k_pack_shots<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
        k_pack_shots<8><<<fmt_grid(items), kFmtThreads, smem, stream>>>((const uint8_t*)d_src, n, shots, d_planes, pstride, bstride, limit);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_unpack_planes(const uint32_t* d_planes, int64_t stride32, int m, int64_t shots, uint8_t* d_dst,
                                 cudaStream_t stream) {
    if (shots <= 0 || m <= 0) return cudaSuccess;
    const int64_t pstride = stride32 ? stride32 : 32, bstride = stride32 ? 32 : (int64_t)32 * m, limit = stride32 ? stride32 : 32;
    const bool flat = m <= kFlatMaxN;
    const size_t smem = (size_t)(flat ? m : kSegCols) * 33 * 4;
    const int64_t items = ((shots + kBlockShots - 1) / kBlockShots) * (flat ? 1 : (m + kSegCols - 1) / kSegCols);
    k_unpack_planes<<<fmt_grid(items), kFmtThreads, smem, stream>>>(d_planes, pstride, bstride, limit, m, shots, d_dst);
    return cudaGetLastError();
}

cudaError_t launch_decode_events(const GenericSide& x, const uint32_t* rows_x, uint32_t lmask_x, const GenericSide& z,
                                 const uint32_t* rows_z, uint32_t lmask_z, const unsigned long long* d_events, int64_t count,
                                 int64_t shots, unsigned long long* d_tally, unsigned long long* d_aux, cudaStream_t stream) {
    if (count <= 0) return cudaSuccess;
    EventTables t;
    for (int j = 0; j < kMaxN; ++j) {
        uint32_t kx = 0, kz = 0;
        for (int b = 0; b < x.m; ++b) kx |= ((rows_x[b] >> j) & 1u) << b;
        for (int b = 0; b < z.m; ++b) kz |= ((rows_z[b] >> j) & 1u) << b;
        t.colkey_x[j] = kx;
        t.colkey_z[j] = kz;
    }
    t.lmask_x = lmask_x;
    t.lmask_z = lmask_z;
    t.fm_x = x.lut_fm;
    t.fm_z = z.lut_fm;
    t.n = x.n;
    k_decode_events<<<fmt_grid((count + kFmtThreads - 1) / kFmtThreads), kFmtThreads, 0, stream>>>(t, d_events, count, shots, d_tally,
                                                                                              d_aux);
    return cudaGetLastError();
}

// d_work: >= (ctas + 1) * 8 bytes of scratch; *d_count receives the number of events (also when it exceeds capacity)
cudaError_t launch_events_from_planes(const uint32_t* d_ex, const uint32_t* d_ez, int n, int64_t stride32, int64_t words,
                                      uint32_t tail_mask, int64_t first_shot, unsigned long long* d_events, int64_t capacity,
                                      unsigned long long* d_count, unsigned long long* d_work, int ctas, cudaStream_t stream) {
    if (words <= 0) return cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), stream);
    const int64_t per_cta = (((words + ctas - 1) / ctas) + kFmtThreads - 1) / kFmtThreads * kFmtThreads;
    k_events_count<<<ctas, kFmtThreads, 0, stream>>>(d_ex, d_ez, n, stride32, words, tail_mask, per_cta, d_work);
    k_events_scan<<<1, 32, 0, stream>>>(d_work, ctas, d_count);
    k_events_fill<<<ctas, kFmtThreads, 0, stream>>>(d_ex, d_ez, n, stride32, words, tail_mask, per_cta, d_work, first_shot,
                                                    d_events, capacity);
    return cudaGetLastError();
}

// ---- compacted host planes -> dense planes (host_compact.h) --------------------------------------------------------
// One warp per block of 2048 plane words: lane l holds bitmap word l (64 words of the block) and the exclusive prefix of
// its population count; in round r the warp writes words 32 r .. 32 r + 31 -- one coalesced 256-byte store -- each lane
// fetching its value, if its bit is set, at rank = prefix of the bitmap word + bits below its own.
namespace {

__global__ void __launch_bounds__(kFmtThreads)
k_zs_expand(const unsigned long long* __restrict__ bm, const uint32_t* __restrict__ off,
            const unsigned long long* __restrict__ vals, unsigned long long* __restrict__ out_x,
            unsigned long long* __restrict__ out_z, int n, int blocks_per_row, int tasks, int64_t slot_stride, int64_t cw) {
    const int lane = threadIdx.x & 31;
    const int task = (int)(((int64_t)blockIdx.x * kFmtThreads + threadIdx.x) >> 5);
    if (task >= tasks) return;                               // whole warps leave together
    const int row = task / blocks_per_row, blk = task % blocks_per_row;
    unsigned long long* const dst = (row < n ? out_x + (int64_t)row * slot_stride : out_z + (int64_t)(row - n) * slot_stride) +
                                    (int64_t)blk * 2048;
    const int64_t left = cw - (int64_t)blk * 2048;           // words of this block inside the chunk
    const unsigned long long mine = bm[(size_t)task * 32 + lane];
    int incl = __popcll(mine);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += up;
    }
    const int excl = incl - __popcll(mine);
    const unsigned long long* const v = vals + off[task];
#pragma unroll 4
    for (int r = 0; r < 64; ++r) {
        const unsigned long long word = __shfl_sync(0xFFFFFFFFu, mine, r >> 1);
        const int before = __shfl_sync(0xFFFFFFFFu, excl, r >> 1);
        const int bit = ((r & 1) << 5) + lane;
        unsigned long long value = 0ull;
        if ((word >> bit) & 1ull) value = v[before + __popcll(word & ((1ull << bit) - 1ull))];
        if (32 * r + lane < left) dst[32 * r + lane] = value;
    }
}

}  // namespace

cudaError_t launch_zs_expand(const uint64_t* d_bm, const uint32_t* d_off, const uint64_t* d_vals, uint64_t* d_x, uint64_t* d_z,
                             int n, int blocks_per_row, int64_t slot_stride, int64_t cw, cudaStream_t stream) {
    const int tasks = 2 * n * blocks_per_row;
    if (tasks <= 0) return cudaSuccess;
    const int64_t threads = (int64_t)tasks * 32;
    k_zs_expand<<<(unsigned)((threads + kFmtThreads - 1) / kFmtThreads), kFmtThreads, 0, stream>>>(
        reinterpret_cast<const unsigned long long*>(d_bm), d_off, reinterpret_cast<const unsigned long long*>(d_vals),
        reinterpret_cast<unsigned long long*>(d_x), reinterpret_cast<unsigned long long*>(d_z), n, blocks_per_row, tasks,
        slot_stride, cw);
    return cudaGetLastError();
}

}  // namespace qcss
