// How fast can the host's cores compact sparse bit planes (zero-word suppression)?  Decides whether a compacting
// host->device path can beat the plain PCIe copy (55 GB/s on this pool), and which inner loop to ship.
//   nvcc -O3 -std=c++17 -o host_scan_bw host_scan_bw.cu -Xcompiler -pthread
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <immintrin.h>

// per word: store, advance past non-zero, set bit (branch-free)
static size_t compact_perword(const uint64_t* src, size_t words, uint64_t* bm, uint64_t* dst) {
    uint64_t* d = dst;
    for (size_t w = 0; w < words; w += 64) {
        uint64_t bits = 0;
        for (int i = 0; i < 64; ++i) {
            const uint64_t v = src[w + i];
            *d = v;
            d += (v != 0);
            bits |= (uint64_t)(v != 0) << i;
        }
        bm[w / 64] = bits;
    }
    return (size_t)(d - dst);
}
// mask of 64 words first (AVX2 compares), then walk its set bits
__attribute__((target("avx2"))) static size_t compact_avx2(const uint64_t* src, size_t words, uint64_t* bm, uint64_t* dst) {
    uint64_t* d = dst;
    const __m256i zero = _mm256_setzero_si256();
    for (size_t w = 0; w < words; w += 64) {
        uint64_t zmask = 0;
        for (int i = 0; i < 64; i += 4) {
            const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + w + i));
            zmask |= (uint64_t)(unsigned)_mm256_movemask_pd(_mm256_castsi256_pd(_mm256_cmpeq_epi64(v, zero))) << i;
        }
        uint64_t bits = ~zmask;
        bm[w / 64] = bits;
        while (bits) { *d++ = src[w + __builtin_ctzll(bits)]; bits &= bits - 1; }
    }
    return (size_t)(d - dst);
}
__attribute__((target("avx512f,popcnt"))) static size_t compact_avx512(const uint64_t* src, size_t words, uint64_t* bm, uint64_t* dst) {
    uint64_t* d = dst;
    for (size_t w = 0; w < words; w += 64) {
        uint64_t bits = 0;
        for (int i = 0; i < 64; i += 8) {
            const __m512i v = _mm512_loadu_si512(reinterpret_cast<const void*>(src + w + i));
            const __mmask8 k = _mm512_test_epi64_mask(v, v);
            _mm512_storeu_si512(reinterpret_cast<void*>(d), _mm512_maskz_compress_epi64(k, v));
            d += _mm_popcnt_u32((unsigned)k);
            bits |= (uint64_t)k << i;
        }
        bm[w / 64] = bits;
    }
    return (size_t)(d - dst);
}
static size_t read_only(const uint64_t* src, size_t words, uint64_t* bm, uint64_t*) {
    uint64_t s = 0;
    for (size_t w = 0; w < words; ++w) s += src[w];
    bm[0] = s;
    return 0;
}

int main() {
    const size_t bytes = (size_t)4 << 30, words = bytes / 8;
    const int hw = (int)std::thread::hardware_concurrency();
    printf("{\"avx2\": %d, \"avx512f\": %d, \"hw\": %d}\n", __builtin_cpu_supports("avx2") ? 1 : 0, __builtin_cpu_supports("avx512f") ? 1 : 0, hw);
    uint64_t* buf = (uint64_t*)malloc(bytes);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < hw; ++t) th.emplace_back([=] {
            uint64_t s = 88172645463325252ull + t;
            for (size_t i = words * t / hw; i < words * (t + 1) / hw; ++i) {
                s ^= s << 13; s ^= s >> 7; s ^= s << 17;
                buf[i] = (s % 100 < 6) ? (1ull << (s >> 58)) : 0ull;
            }
        });
        for (auto& x : th) x.join();
    }
    typedef size_t (*fn_t)(const uint64_t*, size_t, uint64_t*, uint64_t*);
    struct { const char* name; fn_t fn; bool ok; } forms[] = {
        {"read_only", read_only, true}, {"perword", compact_perword, true},
        {"avx2_mask_walk", compact_avx2, __builtin_cpu_supports("avx2") != 0},
        {"avx512_compress", compact_avx512, __builtin_cpu_supports("avx512f") != 0}};
    for (auto& f : forms) {
        if (!f.ok) continue;
        for (int threads : {1, 8, 16}) {
            if (threads > hw) continue;
            std::vector<std::vector<uint64_t>> out(threads), bms(threads);
            for (int t = 0; t < threads; ++t) { out[t].resize(2048 + 64); bms[t].resize(32); }
            std::atomic<size_t> kept{0};
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            for (int t = 0; t < threads; ++t) th.emplace_back([&, t] {
                size_t k = 0;
                for (size_t b = words * t / threads / 2048 * 2048; b + 2048 <= words * (t + 1) / threads; b += 2048)
                    k += f.fn(buf + b, 2048, bms[t].data(), out[t].data());
                kept += k;
            });
            for (auto& x : th) x.join();
            const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("{\"probe\": \"host_scan_bw\", \"form\": \"%s\", \"threads\": %d, \"GBps\": %.1f, \"nonzero_frac\": %.4f}\n", f.name, threads,
                   bytes / s / 1e9, (double)kept.load() / words);
        }
    }
    free(buf);
    return 0;
}
