// How fast can the host's cores compact sparse bit planes (zero-word suppression)?  Decides whether a compacting
// host->device path could beat the plain PCIe copy (55 GB/s on this pool).  nvcc -O3 -o host_scan_bw host_scan_bw.cu
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

static size_t compact_block(const uint64_t* src, size_t words, uint64_t* bm, uint64_t* dst) {
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

int main(int argc, char** argv) {
    const size_t bytes = (size_t)4 << 30, words = bytes / 8;
    const int hw = (int)std::thread::hardware_concurrency();
    for (int pinned = 0; pinned < 2; ++pinned) {
        uint64_t* buf = nullptr;
        if (pinned) { if (cudaHostAlloc((void**)&buf, bytes, cudaHostAllocDefault) != cudaSuccess) { printf("{\"error\": \"cudaHostAlloc\"}\n"); return 1; } }
        else buf = (uint64_t*)malloc(bytes);
        // ~6 % non-zero words (p = 1e-3: 1 - 0.999^64), filled in parallel
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
        for (int threads : {1, 4, 8, 16, 32}) {
            if (threads > 2 * hw) continue;
            std::vector<std::vector<uint64_t>> out(threads), bms(threads);
            for (int t = 0; t < threads; ++t) { out[t].resize(2048 + 64); bms[t].resize(32); }
            std::atomic<size_t> kept{0};
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            for (int t = 0; t < threads; ++t) th.emplace_back([&, t] {
                size_t k = 0;
                for (size_t b = words * t / threads / 2048 * 2048; b + 2048 <= words * (t + 1) / threads; b += 2048)
                    k += compact_block(buf + b, 2048, bms[t].data(), out[t].data());
                kept += k;
            });
            for (auto& x : th) x.join();
            const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("{\"probe\": \"host_scan_bw\", \"pinned\": %d, \"threads\": %d, \"hw\": %d, \"GBps\": %.1f, \"nonzero_frac\": %.4f}\n",
                   pinned, threads, hw, bytes / s / 1e9, (double)kept.load() / words);
        }
        if (pinned) cudaFreeHost(buf); else free(buf);
    }
    return 0;
}
