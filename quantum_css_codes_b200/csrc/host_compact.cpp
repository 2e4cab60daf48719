// Host side of the compacting host->device path (host_compact.h).  Plain C++: compiled by the host compiler only.
#include "host_compact.h"

#include <immintrin.h>

namespace qcss {

namespace {

// branch-free: every word is stored, the cursor advances only past the non-zero ones (dst needs one word of slack)
size_t compact_scalar(const uint64_t* src, int words, uint64_t* bm, uint64_t* dst, int bm_words = kZsBlockWords / 64) {
    uint64_t* d = dst;
    int w = 0;
    for (; w + 64 <= words; w += 64) {
        uint64_t bits = 0;
        for (int i = 0; i < 64; ++i) {
            const uint64_t v = src[w + i];
            *d = v;
            d += (v != 0);
            bits |= (uint64_t)(v != 0) << i;
        }
        bm[w >> 6] = bits;
    }
    if (w < words) {
        uint64_t bits = 0;
        for (int i = 0; w + i < words; ++i) {
            const uint64_t v = src[w + i];
            *d = v;
            d += (v != 0);
            bits |= (uint64_t)(v != 0) << i;
        }
        bm[w >> 6] = bits;
        w += 64;
    }
    for (; (w >> 6) < bm_words; w += 64) bm[w >> 6] = 0;
    return (size_t)(d - dst);
}

// eight words per step: test mask, compress, one full-width store (dst needs eight words of slack)
__attribute__((target("avx512f,popcnt"))) size_t compact_avx512(const uint64_t* src, int words, uint64_t* bm, uint64_t* dst) {
    uint64_t* d = dst;
    int w = 0;
    for (; w + 64 <= words; w += 64) {
        uint64_t bits = 0;
        for (int i = 0; i < 64; i += 8) {
            const __m512i v = _mm512_loadu_si512(reinterpret_cast<const void*>(src + w + i));
            const __mmask8 k = _mm512_test_epi64_mask(v, v);
            _mm512_storeu_si512(reinterpret_cast<void*>(d), _mm512_maskz_compress_epi64(k, v));
            d += _mm_popcnt_u32((unsigned)k);
            bits |= (uint64_t)k << i;
        }
        bm[w >> 6] = bits;
    }
    if (w < words)                                   // ragged end of a chunk: the scalar form finishes the bitmap too
        return (size_t)(d - dst) + compact_scalar(src + w, words - w, bm + (w >> 6), d, kZsBlockWords / 64 - (w >> 6));
    for (; w < kZsBlockWords; w += 64) bm[w >> 6] = 0;
    return (size_t)(d - dst);
}

bool has_avx512() {
    static const bool yes = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("popcnt");
    return yes;
}

}  // namespace

size_t zs_compact_range(const uint64_t* ex, const uint64_t* ez, int64_t e_stride, int n, int64_t w0, int64_t cw,
                        int blocks_per_row, int task0, int task1, uint64_t* bm, uint32_t* off, uint64_t* vals,
                        size_t region_base, size_t region_cap) {
    const bool wide = has_avx512();
    size_t used = 0;
    for (int task = task0; task < task1; ++task) {
        const int row = task / blocks_per_row, blk = task % blocks_per_row;
        const int64_t first = (int64_t)blk * kZsBlockWords;
        const int words = (int)((cw - first) < kZsBlockWords ? (cw - first) : kZsBlockWords);
        if (used + (size_t)kZsBlockWords + 8 > region_cap) return SIZE_MAX;
        const uint64_t* src = (row < n ? ex + (int64_t)row * e_stride : ez + (int64_t)(row - n) * e_stride) + w0 + first;
        uint64_t* dst = vals + region_base + used;
        off[task] = (uint32_t)(region_base + used);
        used += wide ? compact_avx512(src, words, bm + (size_t)task * 32, dst)
                     : compact_scalar(src, words, bm + (size_t)task * 32, dst);
    }
    return used;
}

double zs_density(const uint64_t* p, int64_t words) {
    if (words <= 0) return 0.0;
    int64_t nz = 0;
    for (int64_t i = 0; i < words; ++i) nz += p[i] != 0;
    return (double)nz / (double)words;
}

}  // namespace qcss
