// Internal launch interface between the C ABI (api.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "decode.cuh"
#include "ec_rounds.cuh"

namespace qcss {

// ---- small codes (small_kernels.cu) -----------------------------------------------------------
struct SmallLaunch {
    const GenericSide* x;      // which = 2 side (X errors)
    const GenericSide* z;      // which = 1 side (Z errors)
    DecodeIO io;
    int named_id;              // index into named::kNamed, or -1 for the generic kernels
    bool sample;               // fused Philox sampler instead of loading error planes
    bool gapq = true;          // sampler form below p = 1/64 (options.h): two-phase queue or in place
};
cudaError_t launch_small(const SmallLaunch& l, cudaStream_t stream);
int match_named(const GenericSide& x, const uint32_t* rows_x, uint32_t lx, const GenericSide& z,
                const uint32_t* rows_z, uint32_t lz);
const char* named_name(int id);
int small_bucket_m(int mx, int mz);

// ---- repeated Steane EC, Pauli-frame Monte Carlo (ec_kernels.cu) -------------------------------
struct EcLaunch {
    const GenericSide* x;
    const GenericSide* z;
    EcParams ec;
    int named_id;
    bool gapq = true;
};
cudaError_t launch_ec_rounds(const EcLaunch& l, cudaStream_t stream);

// ---- any-size sparse syndrome (tiled_kernels.cu) --------------------------------------------
struct SparseRows {            // CSR of one parity-check matrix, device pointers
    int m, n, max_row_weight, nnz;
    const int32_t* row_ptr;    // [m + 1]
    const uint16_t* cols;      // [nnz]
    // the same supports with every row padded to a multiple of four entries (pad = column index n, which the fused
    // sampler keeps as an all-zero row): four column indices per 8-byte load, no remainder loop (sample_tiles.cu)
    int groups;                // total number of 4-entry groups
    const int32_t* row_ptr4;   // [m + 1], in groups
    const uint16_t* cols4;     // [4 * groups], 8-byte aligned
    // the transpose (CSC): the checks each column takes part in -- the fused sampler's gap path scatters the few
    // sampled errors into syndrome accumulators instead of gathering mostly-zero error words
    const int32_t* col_ptr;    // [n + 1]
    const uint16_t* rows;      // [nnz]
};
cudaError_t launch_syndrome_tiled(const SparseRows& h, const uint32_t* e_planes, int64_t e_stride,
                                  uint32_t* s_planes, int64_t s_stride, int64_t words,
                                  uint32_t tail_mask, cudaStream_t stream);

// tile-major layout: e = [tiles][n][32 words], s = [tiles][m][32 words], tile = 1024 shots
cudaError_t launch_syndrome_tiles(const SparseRows& h, const uint32_t* e_tiles, uint32_t* s_tiles, int64_t words,
                                  uint32_t tail_mask, cudaStream_t stream);
bool syndrome_tiles_supported(const SparseRows& h);

// fused Philox sampler + sparse syndromes, tile-major outputs (sample_tiles.cu); hx acts on X errors, hz on Z errors
cudaError_t launch_sample_syndrome_tiles(const SparseRows& hx, const SparseRows& hz, uint32_t* sx, uint32_t* sz,
                                         uint32_t* ex, uint32_t* ez, int64_t words, uint32_t tail_mask, uint64_t seed,
                                         uint64_t first_word, uint32_t thr, uint32_t use_gap, const GapTable& gap,
                                         cudaStream_t stream);

// ---- dense syndrome on tensor cores (dense_kernels.cu) ---------------------------------------
size_t dense_h_bytes(int m, int n);
void dense_h_layout(int m, int n, const uint8_t* H, uint8_t* out);      // host-side operand layout
cudaError_t launch_syndrome_mma(const uint8_t* hq, int m, int n, const uint32_t* e_planes, int64_t e_stride,
                                uint32_t* s_planes, int64_t s_stride, int64_t words, uint32_t tail_mask,
                                cudaStream_t stream);

// ---- batched GF(2) Gauss-Jordan (gf2_kernels.cu) --------------------------------------------
cudaError_t launch_gf2_rref(const uint64_t* in, int batch, int m, int n, uint64_t* out,
                            int32_t* rank, int32_t* pivots, cudaStream_t stream);
// second generation (gf2_m4r.cu): one-warp bit-sliced panel, packed combination bytes
bool gf2_m4r_supported(int m, int n);
cudaError_t launch_gf2_m4r(const uint64_t* in, int batch, int m, int n, uint64_t* out,
                           int32_t* rank, int32_t* pivots, cudaStream_t stream);

// fourth generation (gf2_m4r4.cu): 1024-column slabs, conflict-free 128-byte table entries, replay bytes in shared memory
bool gf2_m4r4_supported(int m, int n);
cudaError_t launch_gf2_m4r4(const uint64_t* in, int batch, int m, int n, uint64_t* out,
                            int32_t* rank, int32_t* pivots, cudaStream_t stream);
// third generation (gf2_m4r2.cu): 512-column slabs, two matrices per SM, replay bytes through L2
bool gf2_m4r2_supported(int m, int n);
cudaError_t launch_gf2_m4r2(const uint64_t* in, int batch, int m, int n, uint64_t* out,
                            int32_t* rank, int32_t* pivots, cudaStream_t stream);

// null space / solve on top of the RREF (gf2_derive.cu)
cudaError_t launch_gf2_nullspace(const uint64_t* d_mats, int batch, int m, int n, int max_rows, uint64_t* d_basis,
                                 int32_t* d_rank, int32_t* d_overflow, cudaStream_t stream);
cudaError_t launch_gf2_solve(const uint64_t* d_mats, const uint64_t* d_rhs, int batch, int m, int n, uint64_t* d_x,
                             int32_t* d_consistent, cudaStream_t stream);

// standard form with the reference's column-swap pivot rule, CSS condition (gf2_normalize.cu)
cudaError_t launch_gf2_normalize(uint64_t* d_mats, int batch, int m, int n, int offset, uint64_t* d_partner, int mp,
                                 int32_t* d_swaps, int32_t* d_nswaps, int32_t* d_status, int fail_code,
                                 cudaStream_t stream);
cudaError_t launch_css_condition(const uint64_t* d_h1, int r1, const uint64_t* d_h2, int r2, int n, int32_t* d_status,
                                 int fail_code, cudaStream_t stream);

// GPU-assisted syndrome table (table_kernels.cu)
struct TableBuild;
cudaError_t table_build(int n, int m, const uint8_t* H, int64_t max_entries, TableBuild** out, const char** why);
int table_t(const TableBuild* tb);
int64_t table_count(const TableBuild* tb);
cudaError_t table_read(const TableBuild* tb, int64_t* keys, uint64_t* supports);
void table_free(TableBuild* tb);

// host data formats (format_kernels.cu): the reference's (shots, n) arrays <-> bit planes, sparse event lists
cudaError_t launch_pack_shots(const void* d_src, int elem_bytes, int n, int64_t shots, uint32_t* d_planes, int64_t stride32,
                              cudaStream_t stream);
cudaError_t launch_unpack_planes(const uint32_t* d_planes, int64_t stride32, int m, int64_t shots, uint8_t* d_dst,
                                 cudaStream_t stream);
cudaError_t launch_decode_events(const GenericSide& x, const uint32_t* rows_x, uint32_t lmask_x, const GenericSide& z,
                                 const uint32_t* rows_z, uint32_t lmask_z, const unsigned long long* d_events, int64_t count,
                                 int64_t shots, unsigned long long* d_tally, unsigned long long* d_aux, cudaStream_t stream);
// compacted host planes (host_compact.h) -> the dense planes of one chunk: d_x / d_z rows of slot_stride words, cw valid
cudaError_t launch_zs_expand(const uint64_t* d_bm, const uint32_t* d_off, const uint64_t* d_vals, uint64_t* d_x, uint64_t* d_z,
                             int n, int blocks_per_row, int64_t slot_stride, int64_t cw, cudaStream_t stream);
cudaError_t launch_events_from_planes(const uint32_t* d_ex, const uint32_t* d_ez, int n, int64_t stride32, int64_t words,
                                      uint32_t tail_mask, int64_t first_shot, unsigned long long* d_events, int64_t capacity,
                                      unsigned long long* d_count, unsigned long long* d_work, int ctas, cudaStream_t stream);

// per-syndrome histogram over syndrome planes (hist_kernels.cu)
cudaError_t launch_syndrome_hist(const uint32_t* s, int64_t s_stride, int m, int64_t words, uint32_t tail_mask,
                                 unsigned long long* hist, cudaStream_t stream);

}  // namespace qcss
