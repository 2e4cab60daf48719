// Zero-word suppression of host bit planes (host side of the compacting host->device path of qcss_decode_xz).
//
// At p = 1e-3 a 64-bit plane word holds an error with probability 1 - 0.999^64 = 6 %: the planes a caller hands to
// qcss_decode_xz are 94 % zero words, and the PCIe link (55 GB/s measured on this pool) is what bounds the call.  The
// host's cores read memory faster than that, so they compact each chunk -- per block of kZsBlockWords words of one
// plane: a bitmap of its non-zero words + those words -- the link carries ~8 % of the bytes, and a small kernel
// (format_kernels.cu::k_zs_expand) rebuilds the dense chunk in HBM for the unchanged decode kernel.
#pragma once
#include <cstddef>
#include <cstdint>

namespace qcss {

constexpr int kZsBlockWords = 2048;            // 64-bit words per block: 16 KB of a plane, 32 bitmap words

// Tasks of one chunk are numbered row-major: task = row * blocks_per_row + block, rows 0 .. n-1 = X planes, n .. 2n-1 = Z
// planes.  Compacts tasks [task0, task1): bitmap to bm[task * 32 ..], value offset (in words, relative to the start of
// the value area) to off[task], the non-zero words to vals[region_base + ...].  Returns the number of value words
// written, or SIZE_MAX when they would not fit region_cap (the chunk is then sent uncompacted).
size_t zs_compact_range(const uint64_t* ex, const uint64_t* ez, int64_t e_stride, int n, int64_t w0, int64_t cw,
                        int blocks_per_row, int task0, int task1, uint64_t* bm, uint32_t* off, uint64_t* vals,
                        size_t region_base, size_t region_cap);

// fraction of non-zero words among the first `words` words at p (a quick look before choosing the path)
double zs_density(const uint64_t* p, int64_t words);

}  // namespace qcss
