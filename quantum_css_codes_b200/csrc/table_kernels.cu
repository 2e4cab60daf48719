// GPU-assisted syndrome table: the weight layers of css_code.syndrome_table
// (reference css_code.py:715-735) enumerated and checked for collisions on the device.
//
// The reference walks w = 0, 1, ... and, inside a layer, every weight-w vector in
// bin_matrix.weight_w_vectors order (supports in lexicographic order, bin_matrix.py:57-72); the first
// syndrome that repeats -- against a lower layer or inside the layer -- ends the search with t = w - 1
// and the layer is discarded.  Here one thread takes one rank r of the layer, unranks it to its
// support (combinatorial number system, lexicographic), XORs the big-endian column keys
// (bin_matrix.vec_to_int of the columns of H, bin_matrix.py:36-43) and inserts the key into an
// open-addressing hash set shared by all layers; finding the key already present raises the collision
// flag.  Keys and supports are written at index r, so the host reads the table back in exactly the
// reference's insertion order.
#include <cuda_runtime.h>

#include <vector>

#include "launch.h"

namespace qcss {

namespace {

__device__ __forceinline__ uint64_t hash64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

// returns true when the key was already present
__device__ __forceinline__ bool set_insert(unsigned long long* set, uint64_t mask, uint64_t key) {
    const unsigned long long tag = key + 1ull;            // 0 = empty slot
    uint64_t slot = hash64(key) & mask;
    while (true) {
        const unsigned long long prev = atomicCAS(set + slot, 0ull, tag);
        if (prev == 0ull) return false;
        if (prev == tag) return true;
        slot = (slot + 1) & mask;
    }
}

__global__ void k_table_layer(const uint64_t* __restrict__ colkeys, int n, int w, int64_t total,
                              const uint64_t* __restrict__ binom, unsigned long long* __restrict__ set,
                              uint64_t mask, int64_t* __restrict__ keys_out, uint64_t* __restrict__ supp_out,
                              int* __restrict__ collision) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < total; r += (int64_t)gridDim.x * blockDim.x) {
        uint64_t x = (uint64_t)r, supp = 0ull, key = 0ull;
        int c = 0;
        for (int i = 0; i < w; ++i) {
            // supports starting with element c (w - 1 - i more to pick from the n - 1 - c above it)
            while (true) {
                const uint64_t cnt = binom[(n - 1 - c) * 65 + (w - 1 - i)];
                if (x < cnt) break;
                x -= cnt;
                ++c;
            }
            supp |= 1ull << c;
            key ^= colkeys[c];
            ++c;
        }
        keys_out[r] = (int64_t)key;
        supp_out[r] = supp;
        if (set_insert(set, mask, key)) *collision = 1;
    }
}

__global__ void k_table_reinsert(const int64_t* __restrict__ keys, int64_t count, unsigned long long* __restrict__ set,
                                 uint64_t mask) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < count; r += (int64_t)gridDim.x * blockDim.x)
        set_insert(set, mask, (uint64_t)keys[r]);
}

}  // namespace

struct TableBuild {
    int n = 0, t = 0;
    int64_t count = 0;
    int64_t* d_keys = nullptr;
    uint64_t* d_supp = nullptr;
};

void table_free(TableBuild* tb) {
    if (tb == nullptr) return;
    cudaFree(tb->d_keys);
    cudaFree(tb->d_supp);
    delete tb;
}

// H: row-major 0/1 bytes (m x n), n <= 64, m <= 62.  max_entries bounds the device arrays.
cudaError_t table_build(int n, int m, const uint8_t* H, int64_t max_entries, TableBuild** out, const char** why) {
    *why = "";
    std::vector<uint64_t> binom(65 * 65, 0);
    for (int a = 0; a < 65; ++a) {
        binom[a * 65] = 1;
        for (int b = 1; b <= a; ++b) {
            const unsigned __int128 v = (unsigned __int128)binom[(a - 1) * 65 + b - 1] + binom[(a - 1) * 65 + b];
            binom[a * 65 + b] = v > 0x7FFFFFFFFFFFFFFFull ? 0x7FFFFFFFFFFFFFFFull : (uint64_t)v;
        }
    }
    std::vector<uint64_t> colkeys(n, 0);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j)
            if (H[(size_t)i * n + j] & 1) colkeys[j] |= 1ull << (m - 1 - i);

    TableBuild* tb = new TableBuild;
    tb->n = n;
    uint64_t *d_col = nullptr, *d_binom = nullptr;
    unsigned long long* d_set = nullptr;
    int* d_flag = nullptr;
    int64_t cap_entries = 0;          // capacity of d_keys / d_supp
    uint64_t set_size = 0;
    cudaError_t err = cudaMalloc((void**)&d_col, (size_t)n * 8 + 8);
    if (err == cudaSuccess) err = cudaMalloc((void**)&d_binom, binom.size() * 8);
    if (err == cudaSuccess) err = cudaMalloc((void**)&d_flag, sizeof(int));
    if (err == cudaSuccess) err = cudaMemcpy(d_col, colkeys.data(), (size_t)n * 8, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMemcpy(d_binom, binom.data(), binom.size() * 8, cudaMemcpyHostToDevice);
    int t = n;
    for (int w = 0; w <= n && err == cudaSuccess; ++w) {
        const uint64_t layer = binom[n * 65 + w];
        if (layer > (uint64_t)max_entries || tb->count + (int64_t)layer > max_entries) {
            *why = "syndrome table exceeds max_entries";
            err = cudaErrorMemoryAllocation;
            break;
        }
        const int64_t need = tb->count + (int64_t)layer;
        if (need > cap_entries) {                       // grow the key / support arrays
            int64_t ncap = cap_entries ? cap_entries : 1024;
            while (ncap < need) ncap *= 2;
            int64_t* nk = nullptr;
            uint64_t* ns = nullptr;
            err = cudaMalloc((void**)&nk, (size_t)ncap * 8);
            if (err == cudaSuccess) err = cudaMalloc((void**)&ns, (size_t)ncap * 8);
            if (err == cudaSuccess && tb->count) err = cudaMemcpy(nk, tb->d_keys, (size_t)tb->count * 8, cudaMemcpyDeviceToDevice);
            if (err == cudaSuccess && tb->count) err = cudaMemcpy(ns, tb->d_supp, (size_t)tb->count * 8, cudaMemcpyDeviceToDevice);
            if (err != cudaSuccess) { cudaFree(nk); cudaFree(ns); break; }
            cudaFree(tb->d_keys);
            cudaFree(tb->d_supp);
            tb->d_keys = nk;
            tb->d_supp = ns;
            cap_entries = ncap;
        }
        if ((uint64_t)need * 2 > set_size) {            // grow the hash set and re-insert the accepted layers
            uint64_t nsz = set_size ? set_size : 4096;
            while (nsz < (uint64_t)need * 2) nsz *= 2;
            cudaFree(d_set);
            d_set = nullptr;
            if ((err = cudaMalloc((void**)&d_set, nsz * 8)) != cudaSuccess) break;
            if ((err = cudaMemset(d_set, 0, nsz * 8)) != cudaSuccess) break;
            set_size = nsz;
            if (tb->count) {
                const int blocks = (int)((tb->count + 255) / 256 < 148 * 8 ? (tb->count + 255) / 256 : 148 * 8);
                k_table_reinsert<<<blocks, 256>>>(tb->d_keys, tb->count, d_set, set_size - 1);
            }
        }
        if ((err = cudaMemset(d_flag, 0, sizeof(int))) != cudaSuccess) break;
        const int blocks = (int)((layer + 255) / 256 < 148 * 8 ? (layer + 255) / 256 : 148 * 8);
        k_table_layer<<<blocks, 256>>>(d_col, n, w, (int64_t)layer, d_binom, d_set, set_size - 1,
                                       tb->d_keys + tb->count, tb->d_supp + tb->count, d_flag);
        if ((err = cudaGetLastError()) != cudaSuccess) break;
        int flag = 0;
        if ((err = cudaMemcpy(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        if (flag) {                                     // reference: return w - 1, table (layer discarded)
            t = w - 1;
            break;
        }
        tb->count = need;
    }
    tb->t = t;
    cudaFree(d_col);
    cudaFree(d_binom);
    cudaFree(d_flag);
    cudaFree(d_set);
    if (err != cudaSuccess) {
        table_free(tb);
        tb = nullptr;
    }
    *out = tb;
    return err;
}

int table_t(const TableBuild* tb) { return tb->t; }
int64_t table_count(const TableBuild* tb) { return tb->count; }
cudaError_t table_read(const TableBuild* tb, int64_t* keys, uint64_t* supports) {
    if (tb->count == 0) return cudaSuccess;
    cudaError_t err = cudaMemcpy(keys, tb->d_keys, (size_t)tb->count * 8, cudaMemcpyDeviceToHost);
    if (err == cudaSuccess) err = cudaMemcpy(supports, tb->d_supp, (size_t)tb->count * 8, cudaMemcpyDeviceToHost);
    return err;
}

}  // namespace qcss
