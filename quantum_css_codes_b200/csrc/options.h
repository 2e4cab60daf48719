// Process-wide kernel SELECTION options, set only through qcss_set_option (include/qcss.h).
//
// Every option chooses between implementations that produce bit-identical results (the parity
// tests run both sides of each and compare).  Nothing here is read from the environment: an
// inherited variable must never change what the library computes or which kernel it runs.
#pragma once

namespace qcss {

struct Options {
    int gapq = 1;         // Monte-Carlo below p = 1/64: CTA-wide two-phase gap sampler (1) or in-place form (0)
    int dense = -1;       // large check matrices: -1 = by size and density, 0 = sparse kernels, 1 = tensor cores
    int named = 1;        // 1 = use a built-in static descriptor when the code matches one, 0 = generic kernels
    int gf2_kernel = 0;   // batched RREF: 0 = by shape, 1 = column-by-column, 2 = m4r (one-warp panel), 3 = m4r2, 4 = m4r4
    int host_compact = 1; // qcss_decode_xz on host planes: 1 = zero-word suppression by host threads when the planes are sparse
                          // (host_compact.h), 0 = always the plain chunked copy
    int host_threads = 0; // worker threads of that path: 0 = min(16, hardware threads)
};

Options& options();

}  // namespace qcss
