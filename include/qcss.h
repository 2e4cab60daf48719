/*
 * qcss.h -- C ABI of libqcss.so, the B200 (sm_100a) implementation of the quantum-css-codes hot
 * path: Pauli-error sampling -> syndrome H.e mod 2 -> lookup decode -> logical-failure tally,
 * plus the batched GF(2) row-reduction toolkit.
 *
 * The reference (jimpo/quantum-css-codes) is pure Python and has no FFI; its boundary for this
 * path is the module surface of bin_matrix.py / css_code.py.  Every entry point below names the
 * reference lines it replaces.  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - Every function returns 0 (QCSS_OK) or a negative QCSS_ERR_* code; qcss_last_error() gives
 *     a thread-local message for the last failure.
 *   - The caller owns every host buffer; the library never keeps a caller pointer after return.
 *     Device memory lives behind opaque handles.  No callbacks.  A handle is not thread-safe.
 *   - One process drives one GPU (qcss_set_device); multi-GPU runs are one process per GPU.
 *   - Batches are BIT PLANES: a batch of `shots` binary vectors of length n is n planes of
 *     uint64 words; shot s of plane j is bit (s % 64) of word planes[j * stride + s / 64].
 *     `stride` (in uint64 words) must be a multiple of 2 and >= ceil(shots / 128) * 2; padding
 *     bits are ignored on input and written as zero on output.
 *   - `which` follows the reference's naming (css_code.py:28-30, 461-470):
 *         which = 2 : X errors  -> parity_check_c2, _c2_syndromes, logical Z row (Lz)
 *         which = 1 : Z errors  -> parity_check_c1, _c1_syndromes, logical X row (Lx)
 *   - Table keys are big-endian: key = sum_i s[i] << (m-1-i)   (bin_matrix.py:36-43).
 *   - There is no CPU fallback anywhere in this library.
 */
#ifndef QCSS_H
#define QCSS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define QCSS_API __attribute__((visibility("default")))
#else
#define QCSS_API
#endif

#define QCSS_OK               0
#define QCSS_ERR_INVALID     -1   /* bad argument (Python side raises ValueError)            */
#define QCSS_ERR_CUDA        -2   /* CUDA runtime / driver failure                            */
#define QCSS_ERR_UNSUPPORTED -3   /* shape outside what the kernels cover (no fallback)       */
#define QCSS_ERR_NOMEM       -4

typedef struct qcss_code qcss_code;
typedef struct qcss_table qcss_table;

/* Tallies of one run: fail_x counts Lz.(e_x ^ c_x) = 1, fail_z counts
 * Lx.(e_z ^ c_z) = 1, fail_any their union, miss_* syndromes absent from the table. */
typedef struct qcss_tally {
    uint64_t shots, fail_x, fail_z, fail_any, miss_x, miss_z;
} qcss_tally;

/* Optional inputs/outputs of one decode pass, all DEVICE pointers, NULL = not wanted.
 * ex/ez: error planes [n][e_stride].  synd_*: syndrome planes [m][s_stride] in reference row
 * order.  corr_*: correction planes [n][c_stride].  flip_* / miss_*: one plane each
 * (length e_stride).  tally: uint64[6] in qcss_tally order, ACCUMULATED atomically. */
typedef struct qcss_decode_io {
    const uint64_t* ex;
    const uint64_t* ez;
    int64_t e_stride;
    uint64_t* synd_x;
    uint64_t* synd_z;
    int64_t s_stride;
    uint64_t* corr_x;
    uint64_t* corr_z;
    int64_t c_stride;
    uint64_t* flip_x;
    uint64_t* flip_z;
    uint64_t* miss_x;
    uint64_t* miss_z;
    uint64_t* tally;
} qcss_decode_io;

/* ---- library / device ------------------------------------------------------------------ */
QCSS_API int qcss_version(void);
QCSS_API const char* qcss_last_error(void);
QCSS_API int qcss_device_count(int* count);
QCSS_API int qcss_set_device(int device);
QCSS_API int qcss_host_alloc(void** ptr, size_t bytes);          /* pinned host memory for the e2e path */
QCSS_API int qcss_host_free(void* ptr);
/* Kernel SELECTION options (process-wide).  Each one chooses between implementations whose results are
 * bit-identical -- the parity tests run both sides -- and none is ever read from the environment:
 *   "gapq"       1 (default) two-phase (queue) gap sampler below p = 1/64, 0 the in-place form
 *   "dense"      -1 (default) large check matrices go to the tensor-core kernel by size and density,
 *                0 never, 1 always; taken at qcss_code_create
 *   "named"      1 (default) use a built-in static descriptor when the code matches one, 0 generic kernels;
 *                taken at qcss_code_create
 *   "gf2_kernel" 0 (default) batched RREF kernel by shape, 1 column-by-column, 2 m4r, 3 m4r2 (rows <= 1024),
 *                4 m4r4 (32 < rows <= 1024); a forced kernel falls back to the next one its shape limits allow
 *   "host_compact" 1 (default) qcss_decode_xz zero-word-suppresses sparse host planes with a team of host threads before
 *                the copy (csrc/host_compact.h), 0 always the plain chunked copy
 *   "host_threads" size of that team: 0 (default) = min(16, hardware threads); fewer than 4 disables the path
 * QCSS_ERR_INVALID for an unknown name or a value out of range. */
QCSS_API int qcss_set_option(const char* name, int value);
QCSS_API int qcss_get_option(const char* name, int* value);

/* ---- code object ----------------------------------------------------------------------- */
/* Uploads what CSSCode.__init__ computed (css_code.py:32-75): the NORMALISED parity checks
 * H1 (m1 x n) and H2 (m2 x n) as row-major 0/1 bytes, the logical rows Lx, Lz (n bytes each,
 * css_code.py:124-161; NULL for syndrome-only codes), and the flattened syndrome tables
 * _c1_syndromes / _c2_syndromes (css_code.py:715-735): n_k keys and n_k x n correction bytes
 * (NULL / 0 when the code has no table).  Decoding needs n <= 32 and m <= 16; syndromes work
 * for any n, m. */
QCSS_API int qcss_code_create(int n, int m1, const uint8_t* H1, int m2, const uint8_t* H2,
                     const uint8_t* Lx, const uint8_t* Lz,
                     int64_t n1, const int64_t* keys1, const uint8_t* corr1,
                     int64_t n2, const int64_t* keys2, const uint8_t* corr2,
                     qcss_code** out);
/* Opt-in extension (the reference raises "currently only supports CSS codes for a single logical
 * qubit", css_code.py:74-75): k logical qubits.  Lx / Lz are k x n row-major 0/1 bytes (x_operator_matrix /
 * z_operator_matrix, css_code.py:124-161, which already return k rows).  A shot counts as a logical failure when ANY
 * logical operator flips: flip outputs are the union over the k rows, fail_x / fail_z / fail_any count it; syndromes,
 * corrections and misses are unchanged.  k > 1 runs one decode pass per logical row plus a union kernel (no fused
 * sampler tally, no sparse / EC / specialised paths: QCSS_ERR_UNSUPPORTED). */
QCSS_API int qcss_code_create_multi(int n, int m1, const uint8_t* H1, int m2, const uint8_t* H2, int k, const uint8_t* Lx,
                           const uint8_t* Lz, int64_t n1, const int64_t* keys1, const uint8_t* corr1, int64_t n2,
                           const int64_t* keys2, const uint8_t* corr2, qcss_code** out);
QCSS_API int qcss_code_destroy(qcss_code* code);
/* Human-readable name of the kernel family the code dispatches to (tests). */
QCSS_API int qcss_code_kernel_name(const qcss_code* code, char* buf, int buflen);

/* Kernel specialisation for one code: the static kernel family (H, L and the m <= 5 decode truth tables
 * as compile-time constants -- what runs at the HBM roofline) for ANY decodable code, not only the
 * descriptors built into the library.  qcss_code_spec_source writes the CUDA translation unit (*needed
 * = bytes incl. the terminator; call with buf = NULL to size it); compile it for sm_100a with
 *     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr
 *          -Xcompiler -fPIC -shared -I <quantum_css_codes_b200/csrc> spec.cu -o spec.so
 * and hand the shared object to qcss_code_load_specialized; the code's decode / Monte-Carlo launches
 * then use it (qcss_code_kernel_name reports "small-static(jit:<tag>)").  Results are bit-identical to
 * the generic kernels. */
QCSS_API int qcss_code_spec_source(const qcss_code* code, char* buf, int64_t cap, int64_t* needed);
QCSS_API int qcss_code_load_specialized(qcss_code* code, const char* so_path, const char* tag);
/* The same specialisation IN PROCESS, with no toolkit and no subprocess on the box: the translation unit is compiled
 * by NVRTC (libnvrtc.so.12, opened with dlopen: `nvrtc_path`, or the loader's search path and /usr/local/cuda/lib64
 * when NULL) against the kernel headers embedded in this library, the cubin is loaded with cudaLibraryLoadData and
 * cached in `cache_dir` (NULL = no disk cache) under the hash of source and headers.  ~10 s per distinct code the
 * first time.  qcss_code_kernel_name then reports "small-static(nvrtc:<tag>)".  QCSS_ERR_UNSUPPORTED when the code is
 * not decodable or libnvrtc cannot be found. */
QCSS_API int qcss_code_specialize(qcss_code* code, const char* nvrtc_path, const char* cache_dir);

/* ---- K1 syndrome: replaces np.mod(np.matmul(parity_check, e), 2), css_code.py:728 ------- */
QCSS_API int qcss_syndrome(qcss_code* code, int which, const uint64_t* e_planes, int64_t e_stride,
                  int64_t shots, uint64_t* s_planes, int64_t s_stride);
QCSS_API int qcss_syndrome_dev(qcss_code* code, int which, const uint64_t* d_e_planes, int64_t e_stride,
                      int64_t shots, uint64_t* d_s_planes, int64_t s_stride, void* stream);

/* ---- The same syndrome idiom (css_code.py:728) on TILE-MAJOR batches, for the sparse any-size path
 *      (codes beyond the lookup decoder's n <= 32 / m <= 16, e.g. hypergraph products):
 *      e_tiles[tile][plane j][16 uint64] -- the n plane rows of one tile of 1024 shots are contiguous
 *      (shot s of the batch = bit s % 64 of word (s % 1024) / 64 of row j of tile s / 1024); the last
 *      tile is padded with zero bits.  s_tiles[tile][row i][16 uint64] likewise, row i = syndrome bit i
 *      in the reference's row order.  ceil(shots / 1024) tiles each; 128-byte aligned.  One bulk copy
 *      per part-tile instead of n separate 128-byte streams: see csrc/tiled_kernels.cu.
 *      QCSS_ERR_UNSUPPORTED for codes served by the small or dense kernels. ------------------------- */
QCSS_API int qcss_syndrome_tiles(qcss_code* code, int which, const uint64_t* e_tiles, int64_t shots, uint64_t* s_tiles);
/*      K3 for codes of any size: the depolarising Philox sampler of qcss_mc_sample (same streams: counter =
 *      (global 32-shot word, qubit, block), key = seed) fused into the sparse syndrome kernel; the sampled
 *      errors never reach HBM unless asked for.  sx_tiles = syndromes of the X errors under parity_check_c2 (which = 2),
 *      sz_tiles = syndromes of the Z errors under parity_check_c1 (which = 1), ex_tiles / ez_tiles = the
 *      sampled errors themselves; any may be NULL; all tile-major as above.  first_shot: multiple of 1024. */
QCSS_API int qcss_sample_syndrome_tiles(qcss_code* code, double p, int64_t shots, uint64_t seed, int64_t first_shot,
                               uint64_t* sx_tiles, uint64_t* sz_tiles, uint64_t* ex_tiles, uint64_t* ez_tiles);
QCSS_API int qcss_sample_syndrome_tiles_dev(qcss_code* code, double p, int64_t shots, uint64_t seed, int64_t first_shot,
                                   uint64_t* d_sx_tiles, uint64_t* d_sz_tiles, uint64_t* d_ex_tiles,
                                   uint64_t* d_ez_tiles, void* stream);
QCSS_API int qcss_syndrome_tiles_dev(qcss_code* code, int which, const uint64_t* d_e_tiles, int64_t shots,
                            uint64_t* d_s_tiles, void* stream);

/* Per-syndrome histogram (the tallies a multi-GPU run all-reduces): hist[key] += number of
 * shots whose syndrome has big-endian key `key` (bin_matrix.vec_to_int, bin_matrix.py:36-43);
 * hist has 2^m uint64 entries, m <= 24.  The device form accumulates into d_hist, the host form
 * overwrites hist. */
QCSS_API int qcss_syndrome_hist(qcss_code* code, int which, const uint64_t* e_planes, int64_t e_stride, int64_t shots,
                       uint64_t* hist);
QCSS_API int qcss_syndrome_hist_dev(qcss_code* code, int which, const uint64_t* d_e_planes, int64_t e_stride,
                           int64_t shots, uint64_t* d_hist, void* stream);

/* ---- K1+K2 lookup decode + logical check: replaces the table scan of
 *      quil_classical_correct (css_code.py:649-685) and the Lz/Lx readout (css_code.py:641-646).
 *      corr/flip/miss planes may be NULL; tally may be NULL. ------------------------------ */
QCSS_API int qcss_decode(qcss_code* code, int which, const uint64_t* e_planes, int64_t e_stride,
                int64_t shots, uint64_t* corr_planes, uint64_t* flip_plane, uint64_t* miss_plane,
                qcss_tally* tally);
/* Both Pauli types of the same shots, tallies only; streams the host buffers through the GPU
 * in chunks (copies overlap the kernels). */
QCSS_API int qcss_decode_xz(qcss_code* code, const uint64_t* ex_planes, const uint64_t* ez_planes,
                   int64_t e_stride, int64_t shots, qcss_tally* tally);
/* How the last qcss_decode_xz call on this code moved its planes: bytes sent host -> device and the size of the host
 * team that compacted them (0 = plain chunked copies).  Sparse planes (a 64-bit plane word is non-zero with probability
 * 6 % at p = 1e-3) are zero-word-suppressed by host threads before they cross the link -- option "host_compact" (1 / 0),
 * "host_threads" (0 = min(16, hardware threads)); csrc/host_compact.h.  Same tallies either way. */
QCSS_API int qcss_code_last_transfer(const qcss_code* code, int64_t* h2d_bytes, int* host_threads);
/* General device-pointer form, asynchronous on `stream` (a cudaStream_t, NULL = default). */
QCSS_API int qcss_decode_dev(qcss_code* code, const qcss_decode_io* io, int64_t shots, void* stream);

/* ---- The same calls on the reference's OWN data layout (numpy arrays passed by reference): a C-contiguous
 *      (shots, n) array of 0/1 elements, elem_bytes = 1 (uint8 / bool) or 8 (int64, the reference's dtype='int',
 *      css_code.py:39-40; only bit 0 of an element is used, which is np.mod(x, 2) for two's-complement values).
 *      The transposition to bit planes happens ON THE DEVICE (csrc/format_kernels.cu: coalesced loads of 1024-shot
 *      blocks, one ballot per plane word), chunks overlap their host-to-device copies; results come back in the
 *      reference's layout too: s_out (shots, m) bytes = np.mod(E @ H.T, 2), corr_out (shots, n) bytes, flip_out /
 *      miss_out (shots,) bytes.  Outputs may be NULL.  qcss_pack_shots_dev / qcss_unpack_planes_dev are the
 *      device-pointer transposers themselves (source 16-byte aligned, destination 4-byte aligned). ---------------- */
QCSS_API int qcss_syndrome_shots(qcss_code* code, int which, const void* errors, int elem_bytes, int64_t shots, uint8_t* s_out);
QCSS_API int qcss_decode_shots(qcss_code* code, int which, const void* errors, int elem_bytes, int64_t shots, uint8_t* corr_out,
                      uint8_t* flip_out, uint8_t* miss_out, qcss_tally* tally);
QCSS_API int qcss_decode_xz_shots(qcss_code* code, const void* x_errors, const void* z_errors, int elem_bytes, int64_t shots,
                         qcss_tally* tally);
QCSS_API int qcss_pack_shots_dev(const void* d_src, int elem_bytes, int n, int64_t shots, uint64_t* d_planes, int64_t stride,
                        void* stream);
QCSS_API int qcss_unpack_planes_dev(const uint64_t* d_planes, int64_t stride, int m, int64_t shots, uint8_t* d_dst, void* stream);

/* ---- SPARSE batches: the errors of a batch as a list of events sorted by shot (non-decreasing),
 *          event = shot << 18 | qubit << 2 | pauli,   pauli: 1 = X, 2 = Z, 3 = Y (bit 0 = X component, bit 1 = Z component);
 *      repeated (shot, qubit) events compose by XOR.  At p = 1e-3 a Steane batch is 0.056 bytes per shot in this form
 *      (1.75 as bit planes), and the decode is event driven: shots without events have the zero syndrome and take the
 *      table entry of key 0; every other shot XORs the big-endian column keys (bin_matrix.py:36-43) of its events and
 *      reads the table as quil_classical_correct does (css_code.py:649-685).  Tallies equal those of qcss_decode_xz on
 *      the same batch.  QCSS_ERR_INVALID for unsorted events, qubit >= n, shot >= shots or pauli = 0 (the _dev form
 *      reports those in *d_status: bit 0 bad field, bit 1 unsorted; d_tally = uint64[6], accumulated).
 *      qcss_events_from_planes_dev writes the event list of a batch of bit planes (sorted by shot, then qubit; shot
 *      numbers start at first_shot) and its length to *d_count -- also when that exceeds `capacity`, in which case
 *      only the first `capacity` events are stored. ----------------------------------------------------------------- */
QCSS_API int qcss_decode_xz_sparse(qcss_code* code, const uint64_t* events, int64_t n_events, int64_t shots, qcss_tally* tally);
QCSS_API int qcss_decode_xz_sparse_dev(qcss_code* code, const uint64_t* d_events, int64_t n_events, int64_t shots,
                              uint64_t* d_tally, int32_t* d_status, void* stream);
QCSS_API int qcss_events_from_planes_dev(qcss_code* code, const uint64_t* d_ex, const uint64_t* d_ez, int64_t e_stride,
                                int64_t shots, int64_t first_shot, uint64_t* d_events, int64_t capacity, uint64_t* d_count,
                                void* stream);

/* ---- K3 fused Philox sampler + K1 + K2 (no reference counterpart).
 *      Depolarising noise: each qubit of each shot gets X, Y or Z with probability p/3 each.  For
 *      p >= 1/64 the per-shot error probability is exactly floor(p * 2^32) / 2^32; below that the
 *      gaps between errors are drawn by inverse CDF from a table quantised to 2^-32.  Streams are keyed by
 *      (seed, global shot word, qubit) so results do not depend on how shots are split over calls or GPUs;
 *      first_shot must be a multiple of 128. ------------------------------------------------ */
QCSS_API int qcss_mc_run(qcss_code* code, double p, int64_t shots, uint64_t seed, int64_t first_shot,
                qcss_tally* tally);
QCSS_API int qcss_mc_run_dev(qcss_code* code, double p, int64_t shots, uint64_t seed, int64_t first_shot,
                    uint64_t* d_tally, void* stream);
/* The same sampler writing its error planes out (parity tests against the oracle sampler). */
QCSS_API int qcss_mc_sample(qcss_code* code, double p, int64_t shots, uint64_t seed, int64_t first_shot,
                   uint64_t* ex_planes, uint64_t* ez_planes, int64_t e_stride);
QCSS_API int qcss_mc_sample_dev(qcss_code* code, double p, int64_t shots, uint64_t seed, int64_t first_shot,
                       uint64_t* d_ex_planes, uint64_t* d_ez_planes, int64_t e_stride, void* stream);

/* ---- Pauli-frame Monte Carlo of repeated Steane error correction: the gadget emitted by
 *      CSSCode.error_correct (css_code.py:436-470; the reference only runs it on a QVM,
 *      test/test_fidelity.py), tracked as Pauli errors.  Per round: depolarising(p_data) on the data;
 *      a |+>_L ancilla with depolarising(p_ancilla) takes the data's X errors through a transversal
 *      CNOT (its Z errors flow back), is measured, and the X frame is updated exactly as
 *      quil_classical_correct does with parity_check_c2 / _c2_syndromes (css_code.py:649-685); the same
 *      with a |0>_L ancilla, CNOT ancilla -> data and parity_check_c1 / _c1_syndromes for Z.  After
 *      `rounds` rounds the residual is decoded ideally and tallied like qcss_mc_run.  rounds = 1 and
 *      p_ancilla = 0 reproduce qcss_mc_run(p_data) bit for bit; shot-sharding and first_shot as there.
 *      The model is spelled out in csrc/ec_rounds.cuh and restated in oracle/ec_rounds.py. -------------- */
QCSS_API int qcss_ec_run(qcss_code* code, double p_data, double p_ancilla, int rounds, int64_t shots, uint64_t seed,
                int64_t first_shot, qcss_tally* tally);
QCSS_API int qcss_ec_run_dev(qcss_code* code, double p_data, double p_ancilla, int rounds, int64_t shots, uint64_t seed,
                    int64_t first_shot, uint64_t* d_tally /* [6], accumulated */, void* stream);

/* ---- K4 batched GF(2) Gauss-Jordan: replaces bin_matrix.reduced_row_echelon_form
 *      (bin_matrix.py:8-34; the reference has no batch API and no rank/pivot outputs).
 *      mats: batch matrices of m rows, each row ceil(n/64) uint64 words, column 64w+j is bit j
 *      of word w.  out: same layout, the canonical RREF (zero rows last).  rank: [batch].
 *      pivots: [batch][min(m,n)] pivot column of row i, -1 past the rank; may be NULL.
 *      in and out must not overlap. ------------------------------------------------------- */
QCSS_API int qcss_gf2_rref(const uint64_t* mats, int batch, int m, int n, uint64_t* out, int32_t* rank,
                  int32_t* pivots);
QCSS_API int qcss_gf2_rref_dev(const uint64_t* d_mats, int batch, int m, int n, uint64_t* d_out,
                      int32_t* d_rank, int32_t* d_pivots, void* stream);

/* ---- K4 derived: batched null space and solve (no reference counterpart: bin_matrix.py stops at
 *      the RREF; BASELINE config 5 names the null space).
 *      nullspace: basis[batch][max_basis_rows][ceil(n/64)]; matrix b gets n - rank[b] basis vectors,
 *      one per free column f in increasing order (x[f] = 1, x[pivot_i] = RREF[i][f], other free
 *      variables 0), zero rows after them.  The host form fails with QCSS_ERR_INVALID when a matrix
 *      needs more than max_basis_rows; the device form drops the extra rows and writes the needed
 *      count to *d_overflow (device int, may be NULL; 0 = everything fitted).
 *      solve: rhs[batch][ceil(m/64)] packed right-hand sides, x[batch][ceil(n/64)]; consistent[b] =
 *      1 and x = the solution with every free variable 0, or consistent[b] = 0 and x = 0. -------- */
QCSS_API int qcss_gf2_nullspace(const uint64_t* mats, int batch, int m, int n, int max_basis_rows, uint64_t* basis,
                       int32_t* rank);
QCSS_API int qcss_gf2_nullspace_dev(const uint64_t* d_mats, int batch, int m, int n, int max_basis_rows,
                           uint64_t* d_basis, int32_t* d_rank, int32_t* d_overflow, void* stream);
QCSS_API int qcss_gf2_solve(const uint64_t* mats, const uint64_t* rhs, int batch, int m, int n, uint64_t* x,
                   int32_t* consistent);
QCSS_API int qcss_gf2_solve_dev(const uint64_t* d_mats, const uint64_t* d_rhs, int batch, int m, int n, uint64_t* d_x,
                       int32_t* d_consistent, void* stream);

/* ---- CSS construction numerics on the device.
 *      qcss_gf2_normalize replaces css_code.normalize_parity_check (css_code.py:809-836): every matrix
 *      (layout as for qcss_gf2_rref) is brought to [.. I ..] with the identity block at columns
 *      offset .. offset+m-1 using the reference's pivot rule -- first row at or below the diagonal with a
 *      1 in the pivot column, else a QUBIT swap of the pivot column with the first later column where
 *      row i has a 1.  out = np.mod(h, 2) of the reference; swaps[b][k] = (i + offset, col) pairs in
 *      order, n_swaps[b] of them, capacity n pairs per matrix (unused entries -1); status[b] = 0 or
 *      QCSS_FORM_DEPENDENT_ROWS ("rows are not independent", css_code.py:825-826; out is then
 *      unspecified).  n < offset + m fails with QCSS_ERR_INVALID "not enough columns" (:811-812).
 *      The _dev form works in place on device matrices.
 *      qcss_css_standard_form replaces the numeric core of CSSCode.__init__ (css_code.py:47-61): the CSS
 *      condition H1.H2^T = 0 mod 2, H1 normalised at offset 0 with its swaps replayed on H2, then H2
 *      normalised at offset r1 with its swaps replayed on H1.  *status = 0, QCSS_FORM_NOT_CSS
 *      ("C_2 dual code must be a subspace of C_1", :48-49), QCSS_FORM_DEPENDENT_ROWS_C1 / _C2 or
 *      QCSS_FORM_FEW_COLUMNS_C1 / _C2 -- whichever the reference would raise first.
 *      swaps (capacity n pairs, may be NULL) logs both stages in order; *n_swaps their count. ------- */
#define QCSS_FORM_OK                0
#define QCSS_FORM_NOT_CSS           1
#define QCSS_FORM_DEPENDENT_ROWS    2
#define QCSS_FORM_DEPENDENT_ROWS_C1 2
#define QCSS_FORM_DEPENDENT_ROWS_C2 3
#define QCSS_FORM_FEW_COLUMNS_C1    4   /* "not enough columns": n < r1, or n < r1 + r2 (css_code.py:811-812), */
#define QCSS_FORM_FEW_COLUMNS_C2    5   /* reported in the reference's order of checks                          */
QCSS_API int qcss_gf2_normalize(const uint64_t* mats, int batch, int m, int n, int offset, uint64_t* out, int32_t* swaps,
                       int32_t* n_swaps, int32_t* status);
QCSS_API int qcss_gf2_normalize_dev(uint64_t* d_mats, int batch, int m, int n, int offset, int32_t* d_swaps,
                           int32_t* d_n_swaps, int32_t* d_status, void* stream);
QCSS_API int qcss_css_standard_form(const uint64_t* H1, int r1, const uint64_t* H2, int r2, int n, uint64_t* out1,
                           uint64_t* out2, int32_t* swaps, int32_t* n_swaps, int32_t* status);

/* ---- GPU-assisted syndrome table: replaces the weight-layer search of css_code.syndrome_table
 *      (css_code.py:715-735; bin_matrix.weight_w_vectors order, bin_matrix.py:57-72).
 *      H: row-major 0/1 bytes of an m x n parity check, n <= 64, m <= 62.  Layers w = 0, 1, ... are
 *      enumerated on the device; the first layer containing a repeated syndrome stops the search:
 *      *t = w - 1 and the layer is discarded, exactly as the reference does.  *n_entries = table
 *      size (fails with QCSS_ERR_NOMEM when it would exceed max_entries).  qcss_table_read copies
 *      the entries out in the reference's insertion order (weight, then lexicographic support):
 *      keys[i] = bin_matrix.vec_to_int(H.e mod 2) (big-endian), supports[i] = bit j set <=> e[j] = 1. */
QCSS_API int qcss_table_build(int n, int m, const uint8_t* H, int64_t max_entries, qcss_table** out, int* t,
                     int64_t* n_entries);
QCSS_API int qcss_table_read(const qcss_table* table, int64_t* keys, uint64_t* supports);
QCSS_API int qcss_table_destroy(qcss_table* table);

#ifdef __cplusplus
}
#endif
#endif /* QCSS_H */
