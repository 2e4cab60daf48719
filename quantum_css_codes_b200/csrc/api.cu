// C ABI of libqcss.so (include/qcss.h): code objects, argument checking, launches, and the
// host-buffer entry points that stream caller memory through the GPU.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/qcss.h"
#include "host_compact.h"
#include "launch.h"
#include "options.h"
#include "small_common.cuh"      // NamedArgs, launch-shape helpers (no kernel is instantiated in this file)

using namespace qcss;

namespace qcss {
Options& options() {
    static Options o;
    return o;
}
}  // namespace qcss

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define QCSS_CUDA(expr)                                                                       \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return fail(_e == cudaErrorMemoryAllocation ? QCSS_ERR_NOMEM : QCSS_ERR_CUDA,     \
                        "%s failed: %s", #expr, cudaGetErrorString(_e));                      \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

constexpr int kSlots = 3;
// Largest Monte-Carlo launch: a one-wave grid has >= 148 * 8 warps on a B200, so a warp sees at most
// 2^38 / 1184 = 2.3e8 shots (events) per launch -- a factor 18 below the 32-bit reduce in block_tally.
constexpr int64_t kMaxShotsPerLaunch = (int64_t)1 << 38;

}  // namespace

struct SpecKernels;
void spec_rtc_free(SpecKernels* k);

// Logical rows beyond the first (allow_multi_logical, k > 1): the same H and table with another L row, i.e. another
// flip bit per table entry.
struct LogicalExtra {
    GenericSide side_x{}, side_z{};
    uint32_t rows_x[kMaxM] = {0}, rows_z[kMaxM] = {0};
    uint32_t lmask_x = 0, lmask_z = 0;
    DevBuf fm_x, fm_z, co_x, co_z, e32_x, e32_z;
};

struct qcss_code {
    int n = 0, m1 = 0, m2 = 0;
    int k = 1;                          // logical qubits; k > 1 only through qcss_code_create_multi
    std::vector<LogicalExtra*> extra;   // logical rows 1 .. k-1
    bool small = false;                 // register-resident kernels apply (n <= 32, m <= 16)
    GenericSide side_x{}, side_z{};     // x: which = 2 (H2, _c2_syndromes, Lz); z: which = 1
    uint32_t rows_x[kMaxM] = {0}, rows_z[kMaxM] = {0};
    uint32_t lmask_x = 0, lmask_z = 0;
    uint32_t fm0_x = 0, fm0_z = 0;      // table byte of the zero syndrome (bit0 L.correction, bit1 miss)
    int named_id = -1;
    // kernels compiled for THIS code (qcss_code_spec_source -> nvcc -> qcss_code_load_specialized)
    struct SpecKernels* rtc = nullptr;  // kernels compiled in-process with NVRTC (qcss_code_specialize)
    void* spec_dl = nullptr;
    int (*spec_launch)(const void* small_launch, void* stream) = nullptr;
    char spec_tag[32] = "";
    DevBuf fm_x, fm_z, co_x, co_z, e32_x, e32_z;      // lookup tables
    SparseRows sp1{}, sp2{};            // CSR of H1 / H2 for the tiled kernel
    DevBuf sp1_ptr, sp1_cols, sp2_ptr, sp2_cols;
    DevBuf hq1, hq2;                    // dense H1 / H2 in tensor-core operand layout (dense codes only)
    bool dense1 = false, dense2 = false;
    // host-buffer paths
    cudaStream_t stream = nullptr;
    cudaStream_t slot_stream[kSlots] = {nullptr, nullptr, nullptr};
    DevBuf slot_x[kSlots], slot_z[kSlots];
    DevBuf slot_rx[kSlots], slot_rz[kSlots];          // raw (shots, n) rows / event lists of the format entry points
    DevBuf buf_a, buf_b, buf_c, buf_d, buf_e;
    DevBuf tally;
    // compacting host->device path of qcss_decode_xz (host_compact.h): pinned staging, its device mirror, copy-done events
    void* zs_host[kSlots] = {nullptr, nullptr, nullptr};
    size_t zs_host_cap[kSlots] = {0, 0, 0};
    DevBuf zs_dev[kSlots];
    cudaEvent_t zs_ev[kSlots] = {nullptr, nullptr, nullptr};
    int64_t last_h2d_bytes = 0;         // bytes the last host-buffer decode call sent over the link
    int last_host_threads = 0;          // size of its compacting team (0: plain copies)
};

namespace {

cudaError_t launch_spec_rtc(const SpecKernels& k, const SmallLaunch& l, cudaStream_t stream);

// static kernels of a matching descriptor, kernels compiled for this code, or the generic ones
cudaError_t launch_small_any(const qcss_code* c, const SmallLaunch& l, cudaStream_t stream) {
    if (c->rtc != nullptr && l.named_id < 0) return launch_spec_rtc(*c->rtc, l, stream);
    if (c->spec_launch != nullptr && l.named_id < 0) return (cudaError_t)c->spec_launch(&l, stream);
    return launch_small(l, stream);
}

uint32_t tail_mask_for(int64_t shots) {
    const int r = (int)(shots & 31);
    return r ? ((1u << r) - 1u) : 0xFFFFFFFFu;
}

int check_planes(const void* p, int64_t stride, int64_t shots, const char* what) {
    if (p == nullptr) return fail(QCSS_ERR_INVALID, "%s is NULL", what);
    if (((uintptr_t)p & 15u) != 0) return fail(QCSS_ERR_INVALID, "%s must be 16-byte aligned", what);
    if (stride < 2 || (stride & 1) || stride * 64 < ((shots + 127) / 128) * 128)
        return fail(QCSS_ERR_INVALID, "%s: stride (%lld words) must be even and cover %lld shots",
                    what, (long long)stride, (long long)shots);
    return QCSS_OK;
}

// Host-buffer entry points validate shots and strides BEFORE sizing any allocation or copy from them.
int check_host_stride(int64_t stride, int64_t shots, const char* what) {
    if (shots < 0) return fail(QCSS_ERR_INVALID, "shots must be >= 0");
    if (stride < 2 || (stride & 1) || stride > ((int64_t)1 << 40) || stride * 64 < ((shots + 127) / 128) * 128)
        return fail(QCSS_ERR_INVALID, "%s: stride (%lld words) must be even and cover %lld shots",
                    what, (long long)stride, (long long)shots);
    return QCSS_OK;
}

// Build one side: masks, logical row, dense tables.
int build_side(qcss_code* c, GenericSide& s, uint32_t* rows, uint32_t& lmask, int m, const uint8_t* H,
               const uint8_t* L, int64_t nk, const int64_t* keys, const uint8_t* corr, DevBuf& d_fm,
               DevBuf& d_co, DevBuf& d_e32) {
    const int n = c->n;
    memset(&s, 0, sizeof(s));
    s.n = n;
    s.m = m;
    s.mode = kModeNone;
    lmask = 0;
    for (int t = 0; t < m; ++t) {
        uint32_t r = 0;
        for (int j = 0; j < n; ++j)
            if (H[(size_t)(m - 1 - t) * n + j] & 1) { r |= 1u << j; s.mask[t][j] = 0xFFFFFFFFu; }
        rows[t] = r;
    }
    if (L != nullptr)
        for (int j = 0; j < n; ++j)
            if (L[j] & 1) { lmask |= 1u << j; s.lexp[j] = 0xFFFFFFFFu; }
    if (keys == nullptr || nk <= 0) return QCSS_OK;
    if (L == nullptr) return fail(QCSS_ERR_INVALID, "a syndrome table needs the logical operator row");
    const size_t size = (size_t)1 << m;
    std::vector<uint8_t> fm(size, 2);           // default: miss
    std::vector<uint32_t> co(size, 0);
    for (int64_t k = 0; k < nk; ++k) {
        if (keys[k] < 0 || (uint64_t)keys[k] >= size)
            return fail(QCSS_ERR_INVALID, "table key %lld out of range for m = %d", (long long)keys[k], m);
        uint32_t cm = 0;
        for (int j = 0; j < n; ++j)
            if (corr[(size_t)k * n + j] & 1) cm |= 1u << j;
        co[keys[k]] = cm;
        fm[keys[k]] = (uint8_t)(__builtin_popcount(cm & lmask) & 1);
    }
    s.has_miss = 0;
    for (size_t k = 0; k < size; ++k)
        if (fm[k] & 2) s.has_miss = 1;
    if (&s == &c->side_x) c->fm0_x = fm[0];
    else if (&s == &c->side_z) c->fm0_z = fm[0];
    s.mode = (m <= kSlicedM) ? kModeSliced : kModeLut;
    if (m <= kSlicedM) {
        for (size_t k = 0; k < size; ++k) {
            s.tt_flip |= (uint32_t)(fm[k] & 1) << k;
            s.tt_miss |= (uint32_t)((fm[k] >> 1) & 1) << k;
            for (int j = 0; j < n; ++j) s.tt_corr[j] |= ((co[k] >> j) & 1u) << k;
        }
    }
    QCSS_CUDA(d_fm.reserve(size));
    QCSS_CUDA(d_co.reserve(size * sizeof(uint32_t)));
    QCSS_CUDA(cudaMemcpy(d_fm.p, fm.data(), size, cudaMemcpyHostToDevice));
    QCSS_CUDA(cudaMemcpy(d_co.p, co.data(), size * sizeof(uint32_t), cudaMemcpyHostToDevice));
    s.lut_fm = (const uint8_t*)d_fm.p;
    s.lut_corr = (const uint32_t*)d_co.p;
    if (m > kSlicedM && m <= kMaxE32M) {
        std::vector<uint32_t> e32(size);
        for (size_t k = 0; k < size; ++k)
            e32[k] = ((fm[k] & 1) ? 0x0000FFFFu : 0u) | ((fm[k] & 2) ? 0xFFFF0000u : 0u);
        QCSS_CUDA(d_e32.reserve(size * sizeof(uint32_t)));
        QCSS_CUDA(cudaMemcpy(d_e32.p, e32.data(), size * sizeof(uint32_t), cudaMemcpyHostToDevice));
        s.lut_e32 = (const uint32_t*)d_e32.p;
    }
    return QCSS_OK;
}

int build_sparse(int m, int n, const uint8_t* H, SparseRows& sp, DevBuf& d_ptr, DevBuf& d_cols) {
    std::vector<int32_t> ptr(m + 1, 0);
    std::vector<uint16_t> cols;
    int maxw = 0;
    for (int i = 0; i < m; ++i) {
        int w = 0;
        for (int j = 0; j < n; ++j)
            if (H[(size_t)i * n + j] & 1) { cols.push_back((uint16_t)j); ++w; }
        ptr[i + 1] = (int32_t)cols.size();
        if (w > maxw) maxw = w;
    }
    if (cols.empty()) cols.push_back(0);
    // padded copy: rows in groups of four entries, pad = n (launch.h)
    std::vector<int32_t> ptr4(m + 1, 0);
    std::vector<uint16_t> cols4;
    for (int i = 0; i < m; ++i) {
        for (int k = ptr[i]; k < ptr[i + 1]; ++k) cols4.push_back(cols[k]);
        while (cols4.size() & 3) cols4.push_back((uint16_t)n);
        ptr4[i + 1] = (int32_t)(cols4.size() / 4);
    }
    // transpose (CSC), rows of a column in increasing order
    std::vector<int32_t> cptr(n + 1, 0);
    std::vector<uint16_t> rows((size_t)ptr[m] ? (size_t)ptr[m] : 1, 0);
    for (int k = 0; k < ptr[m]; ++k) ++cptr[cols[k] + 1];
    for (int j = 0; j < n; ++j) cptr[j + 1] += cptr[j];
    {
        std::vector<int32_t> at(cptr.begin(), cptr.end() - 1);
        for (int i = 0; i < m; ++i)
            for (int k = ptr[i]; k < ptr[i + 1]; ++k) rows[at[cols[k]]++] = (uint16_t)i;
    }
    const size_t cols_pad = (cols.size() + 3) & ~(size_t)3;            // cols4 starts 8-byte aligned
    QCSS_CUDA(d_ptr.reserve((ptr.size() + ptr4.size() + cptr.size()) * sizeof(int32_t)));
    QCSS_CUDA(d_cols.reserve((cols_pad + cols4.size() + 4 + rows.size()) * sizeof(uint16_t)));
    QCSS_CUDA(cudaMemcpy(d_ptr.p, ptr.data(), ptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    QCSS_CUDA(cudaMemcpy((int32_t*)d_ptr.p + ptr.size(), ptr4.data(), ptr4.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    QCSS_CUDA(cudaMemcpy((int32_t*)d_ptr.p + ptr.size() + ptr4.size(), cptr.data(), cptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    QCSS_CUDA(cudaMemcpy(d_cols.p, cols.data(), cols.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    if (!cols4.empty())
        QCSS_CUDA(cudaMemcpy((uint16_t*)d_cols.p + cols_pad, cols4.data(), cols4.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    QCSS_CUDA(cudaMemcpy((uint16_t*)d_cols.p + cols_pad + cols4.size() + 4, rows.data(), rows.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    sp.col_ptr = (const int32_t*)d_ptr.p + ptr.size() + ptr4.size();
    sp.rows = (const uint16_t*)d_cols.p + cols_pad + cols4.size() + 4;
    sp.m = m;
    sp.n = n;
    sp.max_row_weight = maxw;
    sp.nnz = ptr[m];
    sp.row_ptr = (const int32_t*)d_ptr.p;
    sp.cols = (const uint16_t*)d_cols.p;
    sp.groups = ptr4[m];
    sp.row_ptr4 = (const int32_t*)d_ptr.p + ptr.size();
    sp.cols4 = (const uint16_t*)d_cols.p + cols_pad;
    return QCSS_OK;
}

// A check matrix goes to the tensor-core kernel when it is big and dense enough for the contraction
// to be a real GEMM (option "dense" = 1 / 0 forces the choice, for the parity tests and the comparison runs).
bool wants_dense(int m, int n, const uint8_t* H) {
    if (options().dense >= 0) return options().dense != 0;
    if ((size_t)m * n < (size_t)256 * 512) return false;
    size_t nnz = 0;
    for (size_t i = 0; i < (size_t)m * n; ++i) nnz += H[i] & 1;
    return nnz * 8 >= (size_t)m * n;                       // density >= 1/8
}

int build_dense(int m, int n, const uint8_t* H, DevBuf& d_hq) {
    const size_t bytes = dense_h_bytes(m, n);
    std::vector<uint8_t> hq(bytes);
    dense_h_layout(m, n, H, hq.data());
    QCSS_CUDA(d_hq.reserve(bytes));
    QCSS_CUDA(cudaMemcpy(d_hq.p, hq.data(), bytes, cudaMemcpyHostToDevice));
    return QCSS_OK;
}

int ensure_streams(qcss_code* c) {
    if (c->stream == nullptr) QCSS_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int i = 0; i < kSlots; ++i)
        if (c->slot_stream[i] == nullptr)
            QCSS_CUDA(cudaStreamCreateWithFlags(&c->slot_stream[i], cudaStreamNonBlocking));
    return QCSS_OK;
}

DecodeIO make_io(const qcss_decode_io* io, int64_t shots) {
    DecodeIO d;
    memset(&d, 0, sizeof(d));
    d.ex = (const uint32_t*)io->ex;
    d.ez = (const uint32_t*)io->ez;
    d.e_stride = io->e_stride * 2;
    d.synd_x = (uint32_t*)io->synd_x;
    d.synd_z = (uint32_t*)io->synd_z;
    d.s_stride = io->s_stride * 2;
    d.corr_x = (uint32_t*)io->corr_x;
    d.corr_z = (uint32_t*)io->corr_z;
    d.c_stride = io->c_stride * 2;
    d.flip_x = (uint32_t*)io->flip_x;
    d.flip_z = (uint32_t*)io->flip_z;
    d.miss_x = (uint32_t*)io->miss_x;
    d.miss_z = (uint32_t*)io->miss_z;
    d.tally = (unsigned long long*)io->tally;
    d.words = (shots + 31) / 32;
    d.tail_mask = tail_mask_for(shots);
    d.sides = (io->ex ? 1 : 0) | (io->ez ? 2 : 0);
    return d;
}

int threshold_from_p(double p, uint32_t* thr) {
    if (!(p >= 0.0) || p > 1.0) return fail(QCSS_ERR_INVALID, "p must be in [0, 1]");
    double t = std::floor(p * 4294967296.0);
    if (t > 4294967295.0) t = 4294967295.0;
    *thr = (uint32_t)t;
    return QCSS_OK;
}

// Gap-sampler table (core.cuh GapTable): cdf[k] = floor((1 - (1-p)^(k+1)) * 2^32), (1-p)^(k+1) by repeated
// multiplication in double precision -- oracle/philox.py::gap_table performs the same IEEE operations.
void gap_table_from_p(double p, GapTable* t) {
    const double q = 1.0 - p;
    double acc = 1.0;
    for (int k = 0; k < 32; ++k) {
        acc = acc * q;
        double v = std::floor((1.0 - acc) * 4294967296.0);
        if (v > 4294967295.0) v = 4294967295.0;
        if (v < 0.0) v = 0.0;
        t->cdf[k] = (uint32_t)v;
    }
    t->inv = t->cdf[0] ? (uint32_t)(4294967295u / t->cdf[0]) : 0xFFFFFFFFu;
}

int launch_decode_multi(qcss_code* c, const qcss_decode_io* io, int64_t shots, cudaStream_t stream);

int launch_decode(qcss_code* c, const qcss_decode_io* io, int64_t shots, cudaStream_t stream) {
    if (shots < 0) return fail(QCSS_ERR_INVALID, "shots must be >= 0");
    if (io->ex == nullptr && io->ez == nullptr) return fail(QCSS_ERR_INVALID, "no error planes given");
    int rc;
    if (io->ex && (rc = check_planes(io->ex, io->e_stride, shots, "ex planes"))) return rc;
    if (io->ez && (rc = check_planes(io->ez, io->e_stride, shots, "ez planes"))) return rc;
    if ((io->synd_x || io->synd_z) && io->s_stride * 64 < shots)
        return fail(QCSS_ERR_INVALID, "syndrome stride too small");
    if ((io->corr_x || io->corr_z) && io->c_stride * 64 < shots)
        return fail(QCSS_ERR_INVALID, "correction stride too small");
    if (!c->small)
        return fail(QCSS_ERR_UNSUPPORTED,
                    "lookup decode covers n <= %d and m <= %d (this code: n = %d, m1 = %d, m2 = %d); "
                    "use qcss_syndrome for syndrome extraction", kMaxN, kMaxM, c->n, c->m1, c->m2);
    const bool wants_decode = io->corr_x || io->corr_z || io->flip_x || io->flip_z || io->miss_x ||
                              io->miss_z || io->tally;
    if (wants_decode) {
        if (io->ex && c->side_x.mode == kModeNone)
            return fail(QCSS_ERR_INVALID, "code has no _c2_syndromes table: cannot decode X errors");
        if (io->ez && c->side_z.mode == kModeNone)
            return fail(QCSS_ERR_INVALID, "code has no _c1_syndromes table: cannot decode Z errors");
    }
    if (shots == 0) return QCSS_OK;
    if (c->k > 1 && wants_decode) return launch_decode_multi(c, io, shots, stream);
    SmallLaunch l;
    l.x = &c->side_x;
    l.z = &c->side_z;
    l.io = make_io(io, shots);
    l.named_id = c->named_id;
    l.sample = false;
    QCSS_CUDA(launch_small_any(c, l, stream));
    return QCSS_OK;
}

// k > 1 (allow_multi_logical; the reference raises at css_code.py:74-75): a shot fails when ANY logical operator
// flips.  One pass of the FULL kernels per logical row writes that row's flip planes, then one kernel ORs them,
// counts and (on request) stores the union as the flip output.  Syndromes, corrections and misses do not depend on
// the logical row and come from the first pass.
__global__ void k_union_flips(const uint64_t* __restrict__ fx, const uint64_t* __restrict__ fz, int k, int64_t stride,
                              int64_t words64, uint64_t tail_mask64, uint64_t* __restrict__ out_x, uint64_t* __restrict__ out_z,
                              unsigned long long* __restrict__ tally, const unsigned long long* __restrict__ first_pass) {
    unsigned long long cx = 0, cz = 0, ca = 0;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < words64; w += (int64_t)gridDim.x * blockDim.x) {
        uint64_t ux = 0, uz = 0;
        for (int i = 0; i < k; ++i) {
            if (fx) ux |= fx[(int64_t)i * stride + w];
            if (fz) uz |= fz[(int64_t)i * stride + w];
        }
        if (w == words64 - 1) { ux &= tail_mask64; uz &= tail_mask64; }
        if (out_x) out_x[w] = ux;
        if (out_z) out_z[w] = uz;
        cx += __popcll(ux); cz += __popcll(uz); ca += __popcll(ux | uz);
    }
    for (int d = 16; d >= 1; d >>= 1) {
        cx += __shfl_xor_sync(0xFFFFFFFFu, cx, d);
        cz += __shfl_xor_sync(0xFFFFFFFFu, cz, d);
        ca += __shfl_xor_sync(0xFFFFFFFFu, ca, d);
    }
    if (tally != nullptr && (threadIdx.x & 31) == 0) {
        if (cx) atomicAdd(tally + 1, cx);
        if (cz) atomicAdd(tally + 2, cz);
        if (ca) atomicAdd(tally + 3, ca);
    }
    if (tally != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(tally + 4, first_pass[4]);
        atomicAdd(tally + 5, first_pass[5]);
    }
}

int launch_decode_multi(qcss_code* c, const qcss_decode_io* io, int64_t shots, cudaStream_t stream) {
    const int k = c->k;
    const int64_t stride = io->e_stride, words64 = (shots + 63) / 64;
    const size_t plane = (size_t)stride * 8;
    uint8_t* scratch = nullptr;
    QCSS_CUDA(cudaMallocAsync((void**)&scratch, 2 * (size_t)k * plane + 64, stream));
    cudaError_t e = cudaMemsetAsync(scratch, 0, 2 * (size_t)k * plane + 64, stream);
    uint64_t* fx = reinterpret_cast<uint64_t*>(scratch);
    uint64_t* fz = fx + (size_t)k * stride;
    unsigned long long* first = reinterpret_cast<unsigned long long*>(scratch + 2 * (size_t)k * plane);
    for (int i = 0; i < k && e == cudaSuccess; ++i) {
        qcss_decode_io pass;
        if (i == 0) {
            pass = *io;                                   // syndromes, corrections, misses: once
            pass.tally = io->tally ? reinterpret_cast<uint64_t*>(first) : nullptr;
        } else {
            memset(&pass, 0, sizeof(pass));
            pass.ex = io->ex;
            pass.ez = io->ez;
            pass.e_stride = io->e_stride;
        }
        pass.flip_x = io->ex ? fx + (size_t)i * stride : nullptr;
        pass.flip_z = io->ez ? fz + (size_t)i * stride : nullptr;
        SmallLaunch l;
        l.x = i == 0 ? &c->side_x : &c->extra[i - 1]->side_x;
        l.z = i == 0 ? &c->side_z : &c->extra[i - 1]->side_z;
        l.io = make_io(&pass, shots);
        l.named_id = -1;
        l.sample = false;
        e = launch_small(l, stream);
    }
    if (e == cudaSuccess) {
        const int r = (int)(shots & 63);
        const uint64_t tail = r ? ((1ull << r) - 1ull) : ~0ull;
        const int64_t blocks = (words64 + 255) / 256;
        k_union_flips<<<(unsigned)(blocks < 1184 ? blocks : 1184), 256, 0, stream>>>(
            io->ex ? fx : nullptr, io->ez ? fz : nullptr, k, stride, words64, tail, io->flip_x, io->flip_z,
            reinterpret_cast<unsigned long long*>(io->tally), first);
        e = cudaGetLastError();
    }
    cudaFreeAsync(scratch, stream);
    QCSS_CUDA(e);
    return QCSS_OK;
}

int launch_mc(qcss_code* c, double p, int64_t shots, uint64_t seed, int64_t first_shot,
              uint64_t* d_tally, uint64_t* d_ex, uint64_t* d_ez, int64_t e_stride, cudaStream_t stream) {
    if (shots < 0) return fail(QCSS_ERR_INVALID, "shots must be >= 0");
    if (first_shot < 0 || (first_shot & 127)) return fail(QCSS_ERR_INVALID, "first_shot must be a multiple of 128");
    if (!c->small) return fail(QCSS_ERR_UNSUPPORTED, "Monte-Carlo run covers n <= %d and m <= %d", kMaxN, kMaxM);
    uint32_t thr = 0;
    int rc = threshold_from_p(p, &thr);
    if (rc) return rc;
    if (d_tally && (c->side_x.mode == kModeNone || c->side_z.mode == kModeNone))
        return fail(QCSS_ERR_INVALID, "Monte-Carlo tallies need both syndrome tables");
    if (d_ex || d_ez) {
        if (d_ex && (rc = check_planes(d_ex, e_stride, shots, "ex planes"))) return rc;
        if (d_ez && (rc = check_planes(d_ez, e_stride, shots, "ez planes"))) return rc;
    }
    if (shots == 0) return QCSS_OK;
    if (c->k > 1 && d_tally) {
        // k > 1: no fused tally kernel; the same Philox streams are written out as planes (chunks of 2^26 shots)
        // and decoded by launch_decode_multi
        const int64_t chunk = (int64_t)1 << 26, cwords = chunk / 64;
        uint64_t* planes = nullptr;
        QCSS_CUDA(cudaMallocAsync((void**)&planes, 2 * (size_t)c->n * cwords * 8, stream));
        for (int64_t done = 0; done < shots && rc == QCSS_OK; done += chunk) {
            const int64_t part = shots - done < chunk ? shots - done : chunk;
            rc = launch_mc(c, p, part, seed, first_shot + done, nullptr, planes, planes + (size_t)c->n * cwords, cwords, stream);
            if (rc == QCSS_OK && (d_ex || d_ez)) rc = fail(QCSS_ERR_UNSUPPORTED, "k > 1: tallies and sampled planes are separate calls");
            if (rc == QCSS_OK) {
                qcss_decode_io io;
                memset(&io, 0, sizeof(io));
                io.ex = planes;
                io.ez = planes + (size_t)c->n * cwords;
                io.e_stride = cwords;
                io.tally = d_tally;
                rc = launch_decode(c, &io, part, stream);
            }
        }
        cudaFreeAsync(planes, stream);
        return rc;
    }
    if (shots > kMaxShotsPerLaunch) {
        // the kernels count events in 32-bit per-thread registers and reduce them per warp in 32 bits before
        // widening: bound the shots of one launch so a warp's share stays far below 2^32 (tallies accumulate)
        for (int64_t done = 0; done < shots; done += kMaxShotsPerLaunch) {
            const int64_t part = shots - done < kMaxShotsPerLaunch ? shots - done : kMaxShotsPerLaunch;
            const int64_t word_off = done / 64;
            rc = launch_mc(c, p, part, seed, first_shot + done, d_tally, d_ex ? d_ex + word_off : nullptr,
                           d_ez ? d_ez + word_off : nullptr, e_stride, stream);
            if (rc) return rc;
        }
        return QCSS_OK;
    }
    SmallLaunch l;
    l.x = &c->side_x;
    l.z = &c->side_z;
    memset(&l.io, 0, sizeof(l.io));
    l.io.words = (shots + 31) / 32;
    l.io.tail_mask = tail_mask_for(shots);
    l.io.sides = 3;
    l.io.tally = (unsigned long long*)d_tally;
    l.io.ex_out = (uint32_t*)d_ex;
    l.io.ez_out = (uint32_t*)d_ez;
    l.io.e_stride = e_stride * 2;
    l.io.seed = seed;
    l.io.first_word = (uint64_t)(first_shot / 32);
    l.io.thr = thr;
    // error rates below 1/64 take the gap sampler (core.cuh::kGapThreshold: measured crossover p ~ 0.02)
    l.io.use_gap = (thr < kGapThreshold) ? 1u : 0u;
    gap_table_from_p(p, &l.io.gap);
    l.named_id = c->named_id;
    l.sample = true;
    l.gapq = options().gapq != 0;
    QCSS_CUDA(launch_small_any(c, l, stream));
    return QCSS_OK;
}

int launch_ec(qcss_code* c, double p_data, double p_anc, int rounds, int64_t shots, uint64_t seed, int64_t first_shot,
              uint64_t* d_tally, cudaStream_t stream) {
    if (shots < 0) return fail(QCSS_ERR_INVALID, "shots must be >= 0");
    if (rounds < 0 || rounds > (1 << 20)) return fail(QCSS_ERR_INVALID, "rounds must be in [0, 2^20]");
    if (first_shot < 0 || (first_shot & 127)) return fail(QCSS_ERR_INVALID, "first_shot must be a multiple of 128");
    if (!c->small) return fail(QCSS_ERR_UNSUPPORTED, "error-correction Monte Carlo covers n <= %d and m <= %d", kMaxN, kMaxM);
    if (!d_tally) return fail(QCSS_ERR_INVALID, "tally is NULL");
    if (c->side_x.mode == kModeNone || c->side_z.mode == kModeNone)
        return fail(QCSS_ERR_INVALID, "error-correction Monte Carlo needs both syndrome tables");
    if (c->k > 1) return fail(QCSS_ERR_UNSUPPORTED, "error-correction Monte Carlo covers k = 1");
    EcLaunch l;
    memset(&l.ec, 0, sizeof(l.ec));
    int rc = threshold_from_p(p_data, &l.ec.thr_p);
    if (rc) return rc;
    if ((rc = threshold_from_p(p_anc, &l.ec.thr_q))) return rc;
    if (shots == 0) return QCSS_OK;
    if (shots > kMaxShotsPerLaunch) {                       // 32-bit per-warp event counters: see launch_mc
        for (int64_t done = 0; done < shots; done += kMaxShotsPerLaunch) {
            const int64_t part = shots - done < kMaxShotsPerLaunch ? shots - done : kMaxShotsPerLaunch;
            if ((rc = launch_ec(c, p_data, p_anc, rounds, part, seed, first_shot + done, d_tally, stream))) return rc;
        }
        return QCSS_OK;
    }
    l.x = &c->side_x;
    l.z = &c->side_z;
    l.named_id = c->named_id;
    l.ec.tally = (unsigned long long*)d_tally;
    l.ec.words = (shots + 31) / 32;
    l.ec.tail_mask = tail_mask_for(shots);
    l.ec.rounds = rounds;
    l.ec.seed = seed;
    l.ec.first_word = (uint64_t)(first_shot / 32);
    l.ec.gap_p = (l.ec.thr_p < kGapThreshold) ? 1u : 0u;                    // same sampler rule as qcss_mc_run
    l.ec.gap_q = (l.ec.thr_q < kGapThreshold) ? 1u : 0u;
    l.gapq = options().gapq != 0;
    gap_table_from_p(p_data, &l.ec.tab_p);
    gap_table_from_p(p_anc, &l.ec.tab_q);
    QCSS_CUDA(launch_ec_rounds(l, stream));
    return QCSS_OK;
}

int launch_syndrome(qcss_code* c, int which, const uint64_t* d_e, int64_t e_stride, int64_t shots,
                    uint64_t* d_s, int64_t s_stride, cudaStream_t stream) {
    if (which != 1 && which != 2) return fail(QCSS_ERR_INVALID, "which must be 1 or 2");
    if (shots < 0) return fail(QCSS_ERR_INVALID, "shots must be >= 0");
    int rc;
    if ((rc = check_planes(d_e, e_stride, shots, "error planes"))) return rc;
    if ((rc = check_planes(d_s, s_stride, shots, "syndrome planes"))) return rc;
    if (shots == 0) return QCSS_OK;
    if (c->small) {
        qcss_decode_io io;
        memset(&io, 0, sizeof(io));
        io.e_stride = e_stride;
        io.s_stride = s_stride;
        if (which == 2) { io.ex = d_e; io.synd_x = d_s; }
        else            { io.ez = d_e; io.synd_z = d_s; }
        SmallLaunch l;
        l.x = &c->side_x;
        l.z = &c->side_z;
        l.io = make_io(&io, shots);
        l.named_id = c->named_id;
        l.sample = false;
        QCSS_CUDA(launch_small_any(c, l, stream));
        return QCSS_OK;
    }
    if ((which == 1) ? c->dense1 : c->dense2) {
        const DevBuf& hq = (which == 1) ? c->hq1 : c->hq2;
        QCSS_CUDA(launch_syndrome_mma((const uint8_t*)hq.p, (which == 1) ? c->m1 : c->m2, c->n, (const uint32_t*)d_e,
                                      e_stride * 2, (uint32_t*)d_s, s_stride * 2, (shots + 31) / 32,
                                      tail_mask_for(shots), stream));
        return QCSS_OK;
    }
    const SparseRows& sp = (which == 1) ? c->sp1 : c->sp2;
    cudaError_t e = launch_syndrome_tiled(sp, (const uint32_t*)d_e, e_stride * 2, (uint32_t*)d_s,
                                          s_stride * 2, (shots + 31) / 32, tail_mask_for(shots), stream);
    if (e == cudaErrorInvalidValue && (size_t)c->n * 16 > 200 * 1024)
        return fail(QCSS_ERR_UNSUPPORTED, "n = %d is too large for the shared-memory tile", c->n);
    QCSS_CUDA(e);
    return QCSS_OK;
}

void tally_from(const uint64_t* h, int64_t shots, qcss_tally* t) {
    t->shots = (uint64_t)shots;
    t->fail_x = h[1];
    t->fail_z = h[2];
    t->fail_any = h[3];
    t->miss_x = h[4];
    t->miss_z = h[5];
}

}  // namespace

extern "C" {

QCSS_API int qcss_version(void) { return 200; }

QCSS_API int qcss_set_option(const char* name, int value) {
    if (!name) return fail(QCSS_ERR_INVALID, "option name is NULL");
    Options& o = options();
    const std::string key(name);
    if (key == "gapq" && (value == 0 || value == 1)) o.gapq = value;
    else if (key == "dense" && value >= -1 && value <= 1) o.dense = value;
    else if (key == "named" && (value == 0 || value == 1)) o.named = value;
    else if (key == "gf2_kernel" && value >= 0 && value <= 4) o.gf2_kernel = value;
    else if (key == "host_compact" && (value == 0 || value == 1)) o.host_compact = value;
    else if (key == "host_threads" && value >= 0 && value <= 256) o.host_threads = value;
    else return fail(QCSS_ERR_INVALID, "unknown option or value out of range: %s = %d", name, value);
    return QCSS_OK;
}

QCSS_API int qcss_get_option(const char* name, int* value) {
    if (!name || !value) return fail(QCSS_ERR_INVALID, "NULL argument");
    const Options& o = options();
    const std::string key(name);
    if (key == "gapq") *value = o.gapq;
    else if (key == "dense") *value = o.dense;
    else if (key == "named") *value = o.named;
    else if (key == "gf2_kernel") *value = o.gf2_kernel;
    else if (key == "host_compact") *value = o.host_compact;
    else if (key == "host_threads") *value = o.host_threads;
    else return fail(QCSS_ERR_INVALID, "unknown option: %s", name);
    return QCSS_OK;
}

QCSS_API const char* qcss_last_error(void) { return g_err; }

QCSS_API int qcss_device_count(int* count) {
    if (!count) return fail(QCSS_ERR_INVALID, "count is NULL");
    QCSS_CUDA(cudaGetDeviceCount(count));
    return QCSS_OK;
}

QCSS_API int qcss_set_device(int device) {
    QCSS_CUDA(cudaSetDevice(device));
    return QCSS_OK;
}

QCSS_API int qcss_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return fail(QCSS_ERR_INVALID, "ptr is NULL");
    QCSS_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
    return QCSS_OK;
}

QCSS_API int qcss_host_free(void* ptr) {
    if (ptr) QCSS_CUDA(cudaFreeHost(ptr));
    return QCSS_OK;
}

QCSS_API int qcss_code_create(int n, int m1, const uint8_t* H1, int m2, const uint8_t* H2, const uint8_t* Lx,
                     const uint8_t* Lz, int64_t n1, const int64_t* keys1, const uint8_t* corr1,
                     int64_t n2, const int64_t* keys2, const uint8_t* corr2, qcss_code** out) {
    if (!out) return fail(QCSS_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (n <= 0 || m1 <= 0 || m2 <= 0 || !H1 || !H2) return fail(QCSS_ERR_INVALID, "bad code dimensions");
    if (n > 65535) return fail(QCSS_ERR_UNSUPPORTED, "n = %d exceeds 65535", n);
    if ((keys1 && !corr1) || (keys2 && !corr2)) return fail(QCSS_ERR_INVALID, "table keys without corrections");
    qcss_code* c = new (std::nothrow) qcss_code();
    if (!c) return fail(QCSS_ERR_NOMEM, "out of host memory");
    c->n = n;
    c->m1 = m1;
    c->m2 = m2;
    c->small = (n <= kMaxN && m1 <= kMaxM && m2 <= kMaxM);
    int rc = QCSS_OK;
    if (c->small) {
        rc = build_side(c, c->side_x, c->rows_x, c->lmask_x, m2, H2, Lz, n2, keys2, corr2, c->fm_x, c->co_x, c->e32_x);
        if (!rc) rc = build_side(c, c->side_z, c->rows_z, c->lmask_z, m1, H1, Lx, n1, keys1, corr1, c->fm_z, c->co_z, c->e32_z);
        if (!rc) c->named_id = match_named(c->side_x, c->rows_x, c->lmask_x, c->side_z, c->rows_z, c->lmask_z);
        if (!options().named) c->named_id = -1;         // option "named" = 0: the generic (runtime-H) kernels
    } else if ((keys1 && n1 > 0) || (keys2 && n2 > 0)) {
        // tables are accepted but unusable: decode entry points will report UNSUPPORTED
    }
    if (!rc) rc = build_sparse(m1, n, H1, c->sp1, c->sp1_ptr, c->sp1_cols);
    if (!rc) rc = build_sparse(m2, n, H2, c->sp2, c->sp2_ptr, c->sp2_cols);
    if (!rc && !c->small) {
        c->dense1 = wants_dense(m1, n, H1);
        c->dense2 = wants_dense(m2, n, H2);
        if (c->dense1) rc = build_dense(m1, n, H1, c->hq1);
        if (!rc && c->dense2) rc = build_dense(m2, n, H2, c->hq2);
    }
    if (rc) {
        qcss_code_destroy(c);
        return rc;
    }
    *out = c;
    return QCSS_OK;
}

QCSS_API int qcss_code_create_multi(int n, int m1, const uint8_t* H1, int m2, const uint8_t* H2, int k, const uint8_t* Lx,
                           const uint8_t* Lz, int64_t n1, const int64_t* keys1, const uint8_t* corr1, int64_t n2,
                           const int64_t* keys2, const uint8_t* corr2, qcss_code** out) {
    if (k < 1 || k > 64) return fail(QCSS_ERR_INVALID, "k must be in [1, 64]");
    if (k > 1 && (!Lx || !Lz)) return fail(QCSS_ERR_INVALID, "k > 1 needs the logical operator rows");
    int rc = qcss_code_create(n, m1, H1, m2, H2, Lx, Lz, n1, keys1, corr1, n2, keys2, corr2, out);
    if (rc || k == 1) return rc;
    qcss_code* c = *out;
    if (!c->small || c->side_x.mode == kModeNone || c->side_z.mode == kModeNone) {
        qcss_code_destroy(c);
        *out = nullptr;
        return fail(QCSS_ERR_UNSUPPORTED, "k > 1 needs a decodable code (n <= %d, m <= %d, both tables)", kMaxN, kMaxM);
    }
    c->k = k;
    c->named_id = -1;                                   // runtime-L kernels
    for (int i = 1; i < k && !rc; ++i) {
        LogicalExtra* x = new (std::nothrow) LogicalExtra();
        if (!x) { rc = fail(QCSS_ERR_NOMEM, "out of host memory"); break; }
        c->extra.push_back(x);
        rc = build_side(c, x->side_x, x->rows_x, x->lmask_x, m2, H2, Lz + (size_t)i * n, n2, keys2, corr2, x->fm_x, x->co_x, x->e32_x);
        if (!rc) rc = build_side(c, x->side_z, x->rows_z, x->lmask_z, m1, H1, Lx + (size_t)i * n, n1, keys1, corr1, x->fm_z, x->co_z, x->e32_z);
    }
    if (rc) {
        qcss_code_destroy(c);
        *out = nullptr;
    }
    return rc;
}

QCSS_API int qcss_code_destroy(qcss_code* c) {
    if (!c) return QCSS_OK;
    c->fm_x.release(); c->fm_z.release(); c->co_x.release(); c->co_z.release();
    c->e32_x.release(); c->e32_z.release();
    c->sp1_ptr.release(); c->sp1_cols.release(); c->sp2_ptr.release(); c->sp2_cols.release();
    c->hq1.release(); c->hq2.release();
    for (int i = 0; i < kSlots; ++i) {
        c->slot_x[i].release();
        c->slot_z[i].release();
        c->slot_rx[i].release();
        c->slot_rz[i].release();
        if (c->slot_stream[i]) cudaStreamDestroy(c->slot_stream[i]);
        if (c->zs_host[i]) cudaFreeHost(c->zs_host[i]);
        c->zs_dev[i].release();
        if (c->zs_ev[i]) cudaEventDestroy(c->zs_ev[i]);
    }
    c->buf_a.release(); c->buf_b.release(); c->buf_c.release(); c->buf_d.release(); c->buf_e.release();
    c->tally.release();
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->spec_dl) dlclose(c->spec_dl);
    spec_rtc_free(c->rtc);
    for (LogicalExtra* x : c->extra) {
        x->fm_x.release(); x->fm_z.release(); x->co_x.release(); x->co_z.release(); x->e32_x.release(); x->e32_z.release();
        delete x;
    }
    delete c;
    return QCSS_OK;
}

QCSS_API int qcss_code_kernel_name(const qcss_code* c, char* buf, int buflen) {
    if (!c || !buf || buflen <= 0) return fail(QCSS_ERR_INVALID, "bad arguments");
    if (!c->small && (c->dense1 || c->dense2))
        snprintf(buf, buflen, "dense-tcgen05(n=%d%s%s)", c->n, c->dense1 ? ",c1" : "", c->dense2 ? ",c2" : "");
    else if (!c->small)
        snprintf(buf, buflen, "tiled-sparse(n=%d)", c->n);
    else if (c->named_id >= 0)
        snprintf(buf, buflen, "small-static(%s)", named_name(c->named_id));
    else if (c->rtc != nullptr)
        snprintf(buf, buflen, "small-static(nvrtc:%s)", c->spec_tag);
    else if (c->spec_launch != nullptr)
        snprintf(buf, buflen, "small-static(jit:%s)", c->spec_tag);
    else
        snprintf(buf, buflen, "small-generic(nb=%d,mb=%d)", c->n <= 16 ? 16 : 32,
                 small_bucket_m(c->side_x.m, c->side_z.m));
    return QCSS_OK;
}

// ---- device-pointer entry points ------------------------------------------------------------

QCSS_API int qcss_syndrome_dev(qcss_code* c, int which, const uint64_t* d_e, int64_t e_stride, int64_t shots,
                      uint64_t* d_s, int64_t s_stride, void* stream) {
    if (!c) return fail(QCSS_ERR_INVALID, "code is NULL");
    return launch_syndrome(c, which, d_e, e_stride, shots, d_s, s_stride, (cudaStream_t)stream);
}

// tile-major batches (sparse any-size path only): [tile][plane][16 uint64], tile = 1024 shots
static int launch_syndrome_tiles_checked(qcss_code* c, int which, const uint64_t* d_e, int64_t shots, uint64_t* d_s,
                                         cudaStream_t stream) {
    if (which != 1 && which != 2) return fail(QCSS_ERR_INVALID, "which must be 1 or 2");
    if (shots < 0) return fail(QCSS_ERR_INVALID, "shots must be >= 0");
    if (!d_e || !d_s) return fail(QCSS_ERR_INVALID, "NULL tiles");
    if ((((uintptr_t)d_e) | ((uintptr_t)d_s)) & 127u) return fail(QCSS_ERR_INVALID, "tiles must be 128-byte aligned");
    if (c->small || ((which == 1) ? c->dense1 : c->dense2))
        return fail(QCSS_ERR_UNSUPPORTED, "the tile-major layout serves the sparse any-size syndrome path "
                                          "(n > %d or m > %d, sparse rows); use qcss_syndrome_dev", kMaxN, kMaxM);
    if (shots == 0) return QCSS_OK;
    const SparseRows& sp = (which == 1) ? c->sp1 : c->sp2;
    cudaError_t e = launch_syndrome_tiles(sp, (const uint32_t*)d_e, (uint32_t*)d_s, (shots + 31) / 32,
                                          tail_mask_for(shots), stream);
    if (e == cudaErrorInvalidValue)
        return fail(QCSS_ERR_UNSUPPORTED, "n = %d, m = %d does not fit the tile-major ring", c->n, sp.m);
    QCSS_CUDA(e);
    return QCSS_OK;
}

static int launch_sample_tiles_checked(qcss_code* c, double p, int64_t shots, uint64_t seed, int64_t first_shot,
                                       uint64_t* d_sx, uint64_t* d_sz, uint64_t* d_ex, uint64_t* d_ez, cudaStream_t stream) {
    if (shots < 0) return fail(QCSS_ERR_INVALID, "shots must be >= 0");
    if (first_shot < 0 || (first_shot & 1023)) return fail(QCSS_ERR_INVALID, "first_shot must be a multiple of 1024 (one tile)");
    if (!d_sx && !d_sz && !d_ex && !d_ez) return fail(QCSS_ERR_INVALID, "no output requested");
    if ((((uintptr_t)d_sx) | ((uintptr_t)d_sz) | ((uintptr_t)d_ex) | ((uintptr_t)d_ez)) & 15u)
        return fail(QCSS_ERR_INVALID, "tiles must be 16-byte aligned");
    uint32_t thr = 0;
    int rc = threshold_from_p(p, &thr);
    if (rc) return rc;
    if (c->small) return fail(QCSS_ERR_UNSUPPORTED, "codes with n <= %d and m <= %d sample through qcss_mc_run / qcss_mc_sample", kMaxN, kMaxM);
    if (shots == 0) return QCSS_OK;
    GapTable gap;
    gap_table_from_p(p, &gap);
    const uint32_t use_gap = (thr < kGapThreshold) ? 1u : 0u;
    cudaError_t e = launch_sample_syndrome_tiles(c->sp2, c->sp1, (uint32_t*)d_sx, (uint32_t*)d_sz, (uint32_t*)d_ex,
                                                 (uint32_t*)d_ez, (shots + 31) / 32, tail_mask_for(shots), seed,
                                                 (uint64_t)(first_shot / 32), thr, use_gap, gap, stream);
    if (e == cudaErrorInvalidValue)
        return fail(QCSS_ERR_UNSUPPORTED, "n = %d does not fit the fused sampler's shared-memory arrays", c->n);
    QCSS_CUDA(e);
    return QCSS_OK;
}

QCSS_API int qcss_sample_syndrome_tiles_dev(qcss_code* c, double p, int64_t shots, uint64_t seed, int64_t first_shot,
                                   uint64_t* d_sx_tiles, uint64_t* d_sz_tiles, uint64_t* d_ex_tiles,
                                   uint64_t* d_ez_tiles, void* stream) {
    if (!c) return fail(QCSS_ERR_INVALID, "code is NULL");
    return launch_sample_tiles_checked(c, p, shots, seed, first_shot, d_sx_tiles, d_sz_tiles, d_ex_tiles, d_ez_tiles,
                                       (cudaStream_t)stream);
}

QCSS_API int qcss_sample_syndrome_tiles(qcss_code* c, double p, int64_t shots, uint64_t seed, int64_t first_shot,
                               uint64_t* sx_tiles, uint64_t* sz_tiles, uint64_t* ex_tiles, uint64_t* ez_tiles) {
    if (!c) return fail(QCSS_ERR_INVALID, "code is NULL");
    if (shots < 0) return fail(QCSS_ERR_INVALID, "shots must be >= 0");
    int rc = ensure_streams(c);
    if (rc) return rc;
    const size_t tiles = (size_t)((shots + 1023) / 1024);
    if (tiles == 0) return QCSS_OK;
    const size_t sxb = tiles * c->m2 * 128, szb = tiles * c->m1 * 128, eb = tiles * c->n * 128;
    if (sx_tiles) QCSS_CUDA(c->buf_a.reserve(sxb));
    if (sz_tiles) QCSS_CUDA(c->buf_b.reserve(szb));
    if (ex_tiles) QCSS_CUDA(c->buf_c.reserve(eb));
    if (ez_tiles) QCSS_CUDA(c->buf_d.reserve(eb));
    rc = launch_sample_tiles_checked(c, p, shots, seed, first_shot, sx_tiles ? (uint64_t*)c->buf_a.p : nullptr,
                                     sz_tiles ? (uint64_t*)c->buf_b.p : nullptr, ex_tiles ? (uint64_t*)c->buf_c.p : nullptr,
                                     ez_tiles ? (uint64_t*)c->buf_d.p : nullptr, c->stream);
    if (rc) { cudaStreamSynchronize(c->stream); return rc; }   // pending copies must not outlive the call
    if (sx_tiles) QCSS_CUDA(cudaMemcpyAsync(sx_tiles, c->buf_a.p, sxb, cudaMemcpyDeviceToHost, c->stream));
    if (sz_tiles) QCSS_CUDA(cudaMemcpyAsync(sz_tiles, c->buf_b.p, szb, cudaMemcpyDeviceToHost, c->stream));
    if (ex_tiles) QCSS_CUDA(cudaMemcpyAsync(ex_tiles, c->buf_c.p, eb, cudaMemcpyDeviceToHost, c->stream));
    if (ez_tiles) QCSS_CUDA(cudaMemcpyAsync(ez_tiles, c->buf_d.p, eb, cudaMemcpyDeviceToHost, c->stream));
    QCSS_CUDA(cudaStreamSynchronize(c->stream));
    return QCSS_OK;
}

QCSS_API int qcss_syndrome_tiles_dev(qcss_code* c, int which, const uint64_t* d_e_tiles, int64_t shots, uint64_t* d_s_tiles,
                            void* stream) {
    if (!c) return fail(QCSS_ERR_INVALID, "code is NULL");
    return launch_syndrome_tiles_checked(c, which, d_e_tiles, shots, d_s_tiles, (cudaStream_t)stream);
}

QCSS_API int qcss_syndrome_tiles(qcss_code* c, int which, const uint64_t* e_tiles, int64_t shots, uint64_t* s_tiles) {
    if (!c) return fail(QCSS_ERR_INVALID, "code is NULL");
    if (which != 1 && which != 2) return fail(QCSS_ERR_INVALID, "which must be 1 or 2");
    if (!e_tiles || !s_tiles) return fail(QCSS_ERR_INVALID, "NULL tiles");
    if (shots < 0) return fail(QCSS_ERR_INVALID, "shots must be >= 0");
    int rc = ensure_streams(c);
    if (rc) return rc;
    const int m = (which == 1) ? c->m1 : c->m2;
    const size_t tiles = (size_t)((shots + 1023) / 1024);
    const size_t eb = tiles * c->n * 128, sb = tiles * m * 128;
    if (tiles == 0) return QCSS_OK;
    QCSS_CUDA(c->buf_a.reserve(eb));
    QCSS_CUDA(c->buf_b.reserve(sb));
    QCSS_CUDA(cudaMemcpyAsync(c->buf_a.p, e_tiles, eb, cudaMemcpyHostToDevice, c->stream));
    rc = launch_syndrome_tiles_checked(c, which, (const uint64_t*)c->buf_a.p, shots, (uint64_t*)c->buf_b.p, c->stream);
    if (rc) { cudaStreamSynchronize(c->stream); return rc; }   // pending copies must not outlive the call
    QCSS_CUDA(cudaMemcpyAsync(s_tiles, c->buf_b.p, sb, cudaMemcpyDeviceToHost, c->stream));
    QCSS_CUDA(cudaStreamSynchronize(c->stream));
    return QCSS_OK;
}

QCSS_API int qcss_decode_dev(qcss_code* c, const qcss_decode_io* io, int64_t shots, void* stream) {
    if (!c || !io) return fail(QCSS_ERR_INVALID, "code or io is NULL");
    return launch_decode(c, io, shots, (cudaStream_t)stream);
}

QCSS_API int qcss_mc_run_dev(qcss_code* c, double p, int64_t shots, uint64_t seed, int64_t first_shot,
                    uint64_t* d_tally, void* stream) {
    if (!c || !d_tally) return fail(QCSS_ERR_INVALID, "code or tally is NULL");
    return launch_mc(c, p, shots, seed, first_shot, d_tally, nullptr, nullptr, 0, (cudaStream_t)stream);
}

QCSS_API int qcss_mc_sample_dev(qcss_code* c, double p, int64_t shots, uint64_t seed, int64_t first_shot,
                       uint64_t* d_ex, uint64_t* d_ez, int64_t e_stride, void* stream) {
    if (!c) return fail(QCSS_ERR_INVALID, "code is NULL");
    if (!d_ex && !d_ez) return fail(QCSS_ERR_INVALID, "no output planes given");
    return launch_mc(c, p, shots, seed, first_shot, nullptr, d_ex, d_ez, e_stride, (cudaStream_t)stream);
}

// ---- host-buffer entry points ---------------------------------------------------------------

QCSS_API int qcss_syndrome(qcss_code* c, int which, const uint64_t* e_planes, int64_t e_stride, int64_t shots,
                  uint64_t* s_planes, int64_t s_stride) {
    if (!c) return fail(QCSS_ERR_INVALID, "code is NULL");
    if (which != 1 && which != 2) return fail(QCSS_ERR_INVALID, "which must be 1 or 2");
    if (!e_planes || !s_planes) return fail(QCSS_ERR_INVALID, "NULL planes");
    int rc;
    if ((rc = check_host_stride(e_stride, shots, "error planes"))) return rc;
    if ((rc = check_host_stride(s_stride, shots, "syndrome planes"))) return rc;
    if ((rc = ensure_streams(c))) return rc;
    const int m = (which == 1) ? c->m1 : c->m2;
    const size_t eb = (size_t)c->n * e_stride * 8, sb = (size_t)m * s_stride * 8;
    QCSS_CUDA(c->buf_a.reserve(eb));
    QCSS_CUDA(c->buf_b.reserve(sb));
    QCSS_CUDA(cudaMemcpyAsync(c->buf_a.p, e_planes, eb, cudaMemcpyHostToDevice, c->stream));
    QCSS_CUDA(cudaMemsetAsync(c->buf_b.p, 0, sb, c->stream));
    rc = launch_syndrome(c, which, (const uint64_t*)c->buf_a.p, e_stride, shots, (uint64_t*)c->buf_b.p,
                         s_stride, c->stream);
    if (rc) { cudaStreamSynchronize(c->stream); return rc; }   // pending copies must not outlive the call
    QCSS_CUDA(cudaMemcpyAsync(s_planes, c->buf_b.p, sb, cudaMemcpyDeviceToHost, c->stream));
    QCSS_CUDA(cudaStreamSynchronize(c->stream));
    return QCSS_OK;
}

QCSS_API int qcss_decode(qcss_code* c, int which, const uint64_t* e_planes, int64_t e_stride, int64_t shots,
                uint64_t* corr_planes, uint64_t* flip_plane, uint64_t* miss_plane, qcss_tally* tally) {
    if (!c) return fail(QCSS_ERR_INVALID, "code is NULL");
    if (which != 1 && which != 2) return fail(QCSS_ERR_INVALID, "which must be 1 or 2");
    if (!e_planes) return fail(QCSS_ERR_INVALID, "NULL planes");
    int rc;
    if ((rc = check_host_stride(e_stride, shots, "error planes"))) return rc;
    if ((rc = ensure_streams(c))) return rc;
    const size_t eb = (size_t)c->n * e_stride * 8, pb = (size_t)e_stride * 8;
    QCSS_CUDA(c->buf_a.reserve(eb));
    QCSS_CUDA(c->buf_b.reserve(eb));
    QCSS_CUDA(c->buf_c.reserve(pb));
    QCSS_CUDA(c->buf_d.reserve(pb));
    QCSS_CUDA(c->tally.reserve(6 * sizeof(uint64_t)));
    QCSS_CUDA(cudaMemcpyAsync(c->buf_a.p, e_planes, eb, cudaMemcpyHostToDevice, c->stream));
    QCSS_CUDA(cudaMemsetAsync(c->buf_b.p, 0, eb, c->stream));
    QCSS_CUDA(cudaMemsetAsync(c->buf_c.p, 0, pb, c->stream));
    QCSS_CUDA(cudaMemsetAsync(c->buf_d.p, 0, pb, c->stream));
    QCSS_CUDA(cudaMemsetAsync(c->tally.p, 0, 6 * sizeof(uint64_t), c->stream));
    qcss_decode_io io;
    memset(&io, 0, sizeof(io));
    io.e_stride = io.c_stride = e_stride;
    io.tally = (uint64_t*)c->tally.p;
    if (which == 2) {
        io.ex = (const uint64_t*)c->buf_a.p;
        io.corr_x = corr_planes ? (uint64_t*)c->buf_b.p : nullptr;
        io.flip_x = (uint64_t*)c->buf_c.p;
        io.miss_x = (uint64_t*)c->buf_d.p;
    } else {
        io.ez = (const uint64_t*)c->buf_a.p;
        io.corr_z = corr_planes ? (uint64_t*)c->buf_b.p : nullptr;
        io.flip_z = (uint64_t*)c->buf_c.p;
        io.miss_z = (uint64_t*)c->buf_d.p;
    }
    rc = launch_decode(c, &io, shots, c->stream);
    if (rc) { cudaStreamSynchronize(c->stream); return rc; }   // pending copies must not outlive the call
    uint64_t h[6] = {0, 0, 0, 0, 0, 0};
    if (corr_planes) QCSS_CUDA(cudaMemcpyAsync(corr_planes, c->buf_b.p, eb, cudaMemcpyDeviceToHost, c->stream));
    if (flip_plane) QCSS_CUDA(cudaMemcpyAsync(flip_plane, c->buf_c.p, pb, cudaMemcpyDeviceToHost, c->stream));
    if (miss_plane) QCSS_CUDA(cudaMemcpyAsync(miss_plane, c->buf_d.p, pb, cudaMemcpyDeviceToHost, c->stream));
    QCSS_CUDA(cudaMemcpyAsync(h, c->tally.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    QCSS_CUDA(cudaStreamSynchronize(c->stream));
    if (tally) tally_from(h, shots, tally);
    return QCSS_OK;
}

}  // extern "C"

namespace {

// The compacting host->device pipeline (host_compact.h).  Two sets of `n_rows` rows of `total_words` 64-bit words each
// (bit planes: n rows, `src_stride` words apart; the reference's (shots, n) arrays: one contiguous row per Pauli type) are
// cut into chunks of `chunk_words` words per row (a multiple of kZsBlockWords); host threads compact chunk after chunk
// (bitmap of non-zero words + the words) into a ring of pinned staging buffers while the main thread copies the finished
// ones, k_zs_expand rebuilds them in dst_x / dst_z (rows `dst_stride` words apart) and `after(chunk, slot, w0, cw,
// stream)` launches whatever consumes a chunk.  A chunk whose non-zero words do not fit half its size goes over with
// plain copies.  At p = 1e-3 the link carries ~8 % of the plane bytes (1 % of the uint8 rows); the call is then bound by
// how fast the host cores read memory (113 GB/s with 16 threads on this pool's boxes against 55 GB/s of PCIe).
template <class After>
int zs_pipeline(qcss_code* c, const uint64_t* src_x, const uint64_t* src_z, int64_t src_stride, int n_rows, int64_t total_words,
                int64_t chunk_words, DevBuf (&dst_x)[kSlots], DevBuf (&dst_z)[kSlots], int64_t dst_stride, int threads,
                After after) {
    const int n = n_rows, rows = 2 * n, T = threads;
    const int64_t nchunks = (total_words + chunk_words - 1) / chunk_words;
    const int bpr_max = (int)(chunk_words / kZsBlockWords);
    const size_t tasks_max = (size_t)rows * bpr_max;
    const size_t bm_bytes = tasks_max * 32 * 8, off_bytes = (tasks_max * 4 + 7) & ~(size_t)7;
    const size_t region_cap = ((size_t)rows * chunk_words / 2) / T + kZsBlockWords + 8;       // words per worker
    const size_t stage_bytes = bm_bytes + off_bytes + (size_t)T * region_cap * 8;
    if ((size_t)T * region_cap >= ((size_t)1 << 32)) return fail(QCSS_ERR_INVALID, "chunk too large for 32-bit value offsets");
    for (int i = 0; i < kSlots; ++i) {
        QCSS_CUDA(c->zs_dev[i].reserve(stage_bytes));
        if (c->zs_host_cap[i] < stage_bytes) {
            if (c->zs_host[i]) cudaFreeHost(c->zs_host[i]);
            c->zs_host[i] = nullptr;
            c->zs_host_cap[i] = 0;
            QCSS_CUDA(cudaHostAlloc(&c->zs_host[i], stage_bytes, cudaHostAllocDefault));
            c->zs_host_cap[i] = stage_bytes;
        }
        if (c->zs_ev[i] == nullptr) QCSS_CUDA(cudaEventCreateWithFlags(&c->zs_ev[i], cudaEventDisableTiming));
    }
    // chunk ci may be compacted into staging slot ci % kSlots once `allowed` >= ci (the copy of chunk ci - kSlots is done)
    std::atomic<int64_t> allowed{kSlots - 1};
    std::atomic<int> stop{0};
    std::vector<std::atomic<int>> done((size_t)nchunks);
    for (auto& d : done) d.store(0, std::memory_order_relaxed);
    std::vector<size_t> used((size_t)nchunks * T, 0);
    auto geometry = [&](int64_t ci, int64_t& w0, int64_t& cw, int& bpr) {
        w0 = ci * chunk_words;
        cw = (total_words - w0 < chunk_words) ? (total_words - w0) : chunk_words;
        bpr = (int)((cw + kZsBlockWords - 1) / kZsBlockWords);
    };
    auto worker = [&](int t) {
        for (int64_t ci = 0; ci < nchunks; ++ci) {
            while (allowed.load(std::memory_order_acquire) < ci) {
                if (stop.load(std::memory_order_relaxed)) return;
                std::this_thread::yield();
            }
            if (stop.load(std::memory_order_relaxed)) return;
            int64_t w0, cw;
            int bpr;
            geometry(ci, w0, cw, bpr);
            const int tasks = rows * bpr;
            uint8_t* const base = static_cast<uint8_t*>(c->zs_host[ci % kSlots]);
            used[(size_t)ci * T + t] = zs_compact_range(src_x, src_z, src_stride, n, w0, cw, bpr, (int)((int64_t)tasks * t / T),
                                                        (int)((int64_t)tasks * (t + 1) / T), reinterpret_cast<uint64_t*>(base),
                                                        reinterpret_cast<uint32_t*>(base + bm_bytes),
                                                        reinterpret_cast<uint64_t*>(base + bm_bytes + off_bytes),
                                                        (size_t)t * region_cap, region_cap);
            done[(size_t)ci].fetch_add(1, std::memory_order_release);
        }
    };
    std::vector<std::thread> team;
    team.reserve((size_t)T);
    try {
        for (int t = 0; t < T; ++t) team.emplace_back(worker, t);
    } catch (...) {                                          // no more threads to be had: nothing has been enqueued yet
        stop.store(1, std::memory_order_relaxed);
        allowed.store(nchunks + kSlots, std::memory_order_release);
        for (auto& th : team) th.join();
        return fail(QCSS_ERR_NOMEM, "could not start %d host threads for the compacting copy (option host_compact = 0 disables it)", T);
    }
    int rc = QCSS_OK;
    cudaError_t err = cudaSuccess;
    int64_t sent = 0;
    for (int64_t ci = 0; ci < nchunks && rc == QCSS_OK && err == cudaSuccess; ++ci) {
        const int slot = (int)(ci % kSlots);
        cudaStream_t st = c->slot_stream[slot];
        if (ci >= 1) {                                       // staging of chunk ci - 1 is free once its copies have landed
            err = cudaEventSynchronize(c->zs_ev[(ci - 1) % kSlots]);
            allowed.store(ci - 1 + kSlots, std::memory_order_release);
            if (err != cudaSuccess) break;
        }
        while (done[(size_t)ci].load(std::memory_order_acquire) < T) std::this_thread::yield();
        int64_t w0, cw;
        int bpr;
        geometry(ci, w0, cw, bpr);
        const int tasks = rows * bpr;
        bool fits = true;
        for (int t = 0; t < T; ++t) fits = fits && used[(size_t)ci * T + t] != SIZE_MAX;
        uint8_t* const h = static_cast<uint8_t*>(c->zs_host[slot]);
        uint8_t* const d = static_cast<uint8_t*>(c->zs_dev[slot].p);
        if (fits) {
            err = cudaMemcpyAsync(d, h, (size_t)tasks * 32 * 8, cudaMemcpyHostToDevice, st);
            if (err == cudaSuccess) err = cudaMemcpyAsync(d + bm_bytes, h + bm_bytes, (size_t)tasks * 4, cudaMemcpyHostToDevice, st);
            sent += (int64_t)tasks * (32 * 8 + 4);
            for (int t = 0; t < T && err == cudaSuccess; ++t) {
                const size_t words = used[(size_t)ci * T + t];
                if (words == 0) continue;
                const size_t at = bm_bytes + off_bytes + (size_t)t * region_cap * 8;
                err = cudaMemcpyAsync(d + at, h + at, words * 8, cudaMemcpyHostToDevice, st);
                sent += (int64_t)words * 8;
            }
            if (err == cudaSuccess) err = cudaEventRecord(c->zs_ev[slot], st);
            if (err == cudaSuccess)
                err = launch_zs_expand(reinterpret_cast<const uint64_t*>(d), reinterpret_cast<const uint32_t*>(d + bm_bytes),
                                       reinterpret_cast<const uint64_t*>(d + bm_bytes + off_bytes),
                                       static_cast<uint64_t*>(dst_x[slot].p), static_cast<uint64_t*>(dst_z[slot].p), n, bpr,
                                       dst_stride, cw, st);
        } else {                                             // dense chunk: the plain strided copy
            err = cudaMemcpy2DAsync(dst_x[slot].p, dst_stride * 8, src_x + w0, src_stride * 8, cw * 8, n, cudaMemcpyHostToDevice, st);
            if (err == cudaSuccess)
                err = cudaMemcpy2DAsync(dst_z[slot].p, dst_stride * 8, src_z + w0, src_stride * 8, cw * 8, n, cudaMemcpyHostToDevice, st);
            if (err == cudaSuccess) err = cudaEventRecord(c->zs_ev[slot], st);
            sent += 2 * cw * 8 * n;
        }
        if (err != cudaSuccess) break;
        rc = after(ci, slot, w0, cw, st);
    }
    stop.store(1, std::memory_order_relaxed);
    allowed.store(nchunks + kSlots, std::memory_order_release);
    for (auto& th : team) th.join();
    for (int i = 0; i < kSlots; ++i) {
        const cudaError_t e2 = cudaStreamSynchronize(c->slot_stream[i]);
        if (err == cudaSuccess) err = e2;
    }
    c->last_h2d_bytes = sent;
    c->last_host_threads = T;
    if (rc != QCSS_OK) return rc;
    QCSS_CUDA(err);
    return QCSS_OK;
}

// qcss_decode_xz on sparse host planes through the pipeline above; the result is the plain path's, bit for bit (same
// kernel, same planes).
int decode_xz_compacted(qcss_code* c, const uint64_t* ex, const uint64_t* ez, int64_t e_stride, int64_t shots, int threads) {
    const int n = c->n;
    const int64_t total_words = (shots + 63) / 64;
    int64_t chunk_words = ((int64_t)(32u << 20) / ((int64_t)n * 8)) / kZsBlockWords * kZsBlockWords;
    if (chunk_words < kZsBlockWords) chunk_words = kZsBlockWords;
    const size_t slot_bytes = (size_t)n * chunk_words * 8;
    for (int i = 0; i < kSlots; ++i) {
        QCSS_CUDA(c->slot_x[i].reserve(slot_bytes));
        QCSS_CUDA(c->slot_z[i].reserve(slot_bytes));
    }
    return zs_pipeline(c, ex, ez, e_stride, n, total_words, chunk_words, c->slot_x, c->slot_z, chunk_words, threads,
                       [&](int64_t, int slot, int64_t w0, int64_t cw, cudaStream_t st) {
                           const int64_t cshots = (w0 + cw == total_words) ? (shots - w0 * 64) : cw * 64;
                           qcss_decode_io io;
                           memset(&io, 0, sizeof(io));
                           io.ex = (const uint64_t*)c->slot_x[slot].p;
                           io.ez = (const uint64_t*)c->slot_z[slot].p;
                           io.e_stride = chunk_words;
                           io.tally = (uint64_t*)c->tally.p;
                           return launch_decode(c, &io, cshots, st);
                       });
}

// the compacting path pays when the data is sparse, long enough to amortise a thread team, and cores are there
int compacting_threads(const uint64_t* x, const uint64_t* z, int64_t words_per_set) {
    if (options().host_compact == 0) return 0;
    if (words_per_set * 16 < ((int64_t)64 << 20)) return 0;                        // < 64 MB in all: one plain copy
    if ((((uintptr_t)x) | ((uintptr_t)z)) & 7u) return 0;                          // word loads want 8-byte alignment
    // a core streams 7-13 GB/s (profiles/r02_host_scan_bw.jsonl): left to itself the path needs eight of them to beat the
    // link (measured on an 8-GPU box with 4 cores per rank: 18 GB/s per GPU against 23 GB/s of plain copies)
    int threads = options().host_threads;
    if (threads <= 0) {
        threads = (int)std::thread::hardware_concurrency();
        if (threads > 16) threads = 16;
        if (threads < 8) return 0;
    }
    if (threads < 4) return 0;
    const int64_t probe = words_per_set < 65536 ? words_per_set : 65536;
    if (zs_density(x, probe) > 0.25 || zs_density(z, probe) > 0.25) return 0;      // dense data: nothing to gain
    return threads;
}

}  // namespace

extern "C" {

QCSS_API int qcss_decode_xz(qcss_code* c, const uint64_t* ex, const uint64_t* ez, int64_t e_stride, int64_t shots,
                   qcss_tally* tally) {
    if (!c || !tally) return fail(QCSS_ERR_INVALID, "code or tally is NULL");
    if (!ex || !ez) return fail(QCSS_ERR_INVALID, "NULL planes");
    if (shots < 0) return fail(QCSS_ERR_INVALID, "shots must be >= 0");
    if (e_stride < 2 || (e_stride & 1) || e_stride * 64 < shots)
        return fail(QCSS_ERR_INVALID, "stride must be even and cover all shots");
    int rc = ensure_streams(c);
    if (rc) return rc;
    QCSS_CUDA(c->tally.reserve(6 * sizeof(uint64_t)));
    QCSS_CUDA(cudaMemsetAsync(c->tally.p, 0, 6 * sizeof(uint64_t), c->stream));
    QCSS_CUDA(cudaStreamSynchronize(c->stream));
    if (const int threads = compacting_threads(ex, ez, (int64_t)c->n * ((shots + 63) / 64))) {
        rc = decode_xz_compacted(c, ex, ez, e_stride, shots, threads);
        if (rc) return rc;
        uint64_t hz[6];
        QCSS_CUDA(cudaMemcpy(hz, c->tally.p, sizeof(hz), cudaMemcpyDeviceToHost));
        tally_from(hz, shots, tally);
        return QCSS_OK;
    }
    // chunk = a slice of every plane, sized so one slot holds ~32 MB per Pauli type
    const int64_t total_words = (shots + 63) / 64;
    int64_t chunk_words = ((int64_t)(32u << 20) / ((int64_t)c->n * 8)) & ~(int64_t)1;
    if (chunk_words < 2) chunk_words = 2;
    if (chunk_words > ((total_words + 1) & ~(int64_t)1)) chunk_words = (total_words + 1) & ~(int64_t)1;
    const size_t slot_bytes = (size_t)c->n * chunk_words * 8;
    for (int i = 0; i < kSlots; ++i) {
        QCSS_CUDA(c->slot_x[i].reserve(slot_bytes));
        QCSS_CUDA(c->slot_z[i].reserve(slot_bytes));
    }
    int slot = 0;
    for (int64_t w0 = 0; w0 < total_words; w0 += chunk_words, slot = (slot + 1) % kSlots) {
        const int64_t cw = (total_words - w0 < chunk_words) ? (total_words - w0) : chunk_words;
        const int64_t cshots = (w0 + cw == total_words) ? (shots - w0 * 64) : cw * 64;
        cudaStream_t st = c->slot_stream[slot];
        // a partial last chunk leaves stale words behind cw inside the slot: they sit past
        // `words` for this launch and are masked out by the kernel
        QCSS_CUDA(cudaMemcpy2DAsync(c->slot_x[slot].p, chunk_words * 8, ex + w0, e_stride * 8, cw * 8, c->n,
                                    cudaMemcpyHostToDevice, st));
        QCSS_CUDA(cudaMemcpy2DAsync(c->slot_z[slot].p, chunk_words * 8, ez + w0, e_stride * 8, cw * 8, c->n,
                                    cudaMemcpyHostToDevice, st));
        qcss_decode_io io;
        memset(&io, 0, sizeof(io));
        io.ex = (const uint64_t*)c->slot_x[slot].p;
        io.ez = (const uint64_t*)c->slot_z[slot].p;
        io.e_stride = chunk_words;
        io.tally = (uint64_t*)c->tally.p;
        rc = launch_decode(c, &io, cshots, st);
        if (rc) {
            for (int i = 0; i < kSlots; ++i) cudaStreamSynchronize(c->slot_stream[i]);
            return rc;
        }
    }
    for (int i = 0; i < kSlots; ++i) QCSS_CUDA(cudaStreamSynchronize(c->slot_stream[i]));
    c->last_h2d_bytes = 2 * total_words * 8 * (int64_t)c->n;
    c->last_host_threads = 0;
    uint64_t h[6];
    QCSS_CUDA(cudaMemcpy(h, c->tally.p, sizeof(h), cudaMemcpyDeviceToHost));
    tally_from(h, shots, tally);
    return QCSS_OK;
}

QCSS_API int qcss_code_last_transfer(const qcss_code* c, int64_t* h2d_bytes, int* host_threads) {
    if (!c) return fail(QCSS_ERR_INVALID, "code is NULL");
    if (h2d_bytes) *h2d_bytes = c->last_h2d_bytes;
    if (host_threads) *host_threads = c->last_host_threads;
    return QCSS_OK;
}

QCSS_API int qcss_mc_run(qcss_code* c, double p, int64_t shots, uint64_t seed, int64_t first_shot, qcss_tally* tally) {
    if (!c || !tally) return fail(QCSS_ERR_INVALID, "code or tally is NULL");
    int rc = ensure_streams(c);
    if (rc) return rc;
    QCSS_CUDA(c->tally.reserve(6 * sizeof(uint64_t)));
    QCSS_CUDA(cudaMemsetAsync(c->tally.p, 0, 6 * sizeof(uint64_t), c->stream));
    rc = launch_mc(c, p, shots, seed, first_shot, (uint64_t*)c->tally.p, nullptr, nullptr, 0, c->stream);
    if (rc) { cudaStreamSynchronize(c->stream); return rc; }   // pending copies must not outlive the call
    uint64_t h[6];
    QCSS_CUDA(cudaMemcpyAsync(h, c->tally.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    QCSS_CUDA(cudaStreamSynchronize(c->stream));
    tally_from(h, shots, tally);
    return QCSS_OK;
}

QCSS_API int qcss_ec_run_dev(qcss_code* c, double p_data, double p_ancilla, int rounds, int64_t shots, uint64_t seed,
                    int64_t first_shot, uint64_t* d_tally, void* stream) {
    if (!c) return fail(QCSS_ERR_INVALID, "code is NULL");
    return launch_ec(c, p_data, p_ancilla, rounds, shots, seed, first_shot, d_tally, (cudaStream_t)stream);
}

QCSS_API int qcss_ec_run(qcss_code* c, double p_data, double p_ancilla, int rounds, int64_t shots, uint64_t seed,
                int64_t first_shot, qcss_tally* tally) {
    if (!c || !tally) return fail(QCSS_ERR_INVALID, "code or tally is NULL");
    int rc = ensure_streams(c);
    if (rc) return rc;
    QCSS_CUDA(c->tally.reserve(6 * sizeof(uint64_t)));
    QCSS_CUDA(cudaMemsetAsync(c->tally.p, 0, 6 * sizeof(uint64_t), c->stream));
    rc = launch_ec(c, p_data, p_ancilla, rounds, shots, seed, first_shot, (uint64_t*)c->tally.p, c->stream);
    if (rc) { cudaStreamSynchronize(c->stream); return rc; }   // pending copies must not outlive the call
    uint64_t h[6];
    QCSS_CUDA(cudaMemcpyAsync(h, c->tally.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    QCSS_CUDA(cudaStreamSynchronize(c->stream));
    tally_from(h, shots, tally);
    return QCSS_OK;
}

QCSS_API int qcss_mc_sample(qcss_code* c, double p, int64_t shots, uint64_t seed, int64_t first_shot, uint64_t* ex,
                   uint64_t* ez, int64_t e_stride) {
    if (!c) return fail(QCSS_ERR_INVALID, "code is NULL");
    if (!ex || !ez) return fail(QCSS_ERR_INVALID, "NULL planes");
    int rc;
    if ((rc = check_host_stride(e_stride, shots, "error planes"))) return rc;
    if (first_shot < 0 || (first_shot & 127)) return fail(QCSS_ERR_INVALID, "first_shot must be a multiple of 128");
    if ((rc = ensure_streams(c))) return rc;
    const size_t eb = (size_t)c->n * e_stride * 8;
    QCSS_CUDA(c->buf_a.reserve(eb));
    QCSS_CUDA(c->buf_b.reserve(eb));
    QCSS_CUDA(cudaMemsetAsync(c->buf_a.p, 0, eb, c->stream));
    QCSS_CUDA(cudaMemsetAsync(c->buf_b.p, 0, eb, c->stream));
    rc = launch_mc(c, p, shots, seed, first_shot, nullptr, (uint64_t*)c->buf_a.p, (uint64_t*)c->buf_b.p,
                   e_stride, c->stream);
    if (rc) { cudaStreamSynchronize(c->stream); return rc; }   // pending copies must not outlive the call
    QCSS_CUDA(cudaMemcpyAsync(ex, c->buf_a.p, eb, cudaMemcpyDeviceToHost, c->stream));
    QCSS_CUDA(cudaMemcpyAsync(ez, c->buf_b.p, eb, cudaMemcpyDeviceToHost, c->stream));
    QCSS_CUDA(cudaStreamSynchronize(c->stream));
    return QCSS_OK;
}

// ---- K4 -------------------------------------------------------------------------------------

QCSS_API int qcss_gf2_rref_dev(const uint64_t* d_mats, int batch, int m, int n, uint64_t* d_out, int32_t* d_rank,
                      int32_t* d_pivots, void* stream) {
    if (batch < 0 || m < 0 || n < 0) return fail(QCSS_ERR_INVALID, "negative dimensions");
    if (batch == 0 || m == 0 || n == 0) return QCSS_OK;
    if (!d_mats || !d_out) return fail(QCSS_ERR_INVALID, "NULL matrices");
    if (d_mats == d_out) return fail(QCSS_ERR_INVALID, "in-place reduction is not supported");
    QCSS_CUDA(launch_gf2_rref(d_mats, batch, m, n, d_out, d_rank, d_pivots, (cudaStream_t)stream));
    return QCSS_OK;
}

QCSS_API int qcss_gf2_rref(const uint64_t* mats, int batch, int m, int n, uint64_t* out, int32_t* rank,
                  int32_t* pivots) {
    if (batch < 0 || m < 0 || n < 0) return fail(QCSS_ERR_INVALID, "negative dimensions");
    if (batch == 0 || m == 0 || n == 0) return QCSS_OK;
    if (!mats || !out) return fail(QCSS_ERR_INVALID, "NULL matrices");
    const size_t W = (size_t)(n + 63) / 64, bytes = (size_t)batch * m * W * 8;
    const size_t npiv = (size_t)(m < n ? m : n);
    void *d_in = nullptr, *d_out = nullptr;
    int32_t *d_rank = nullptr, *d_piv = nullptr;
    int rc = QCSS_OK;
    cudaError_t e = cudaMalloc(&d_in, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_rank, (size_t)batch * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_piv, (size_t)batch * npiv * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemcpy(d_in, mats, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_gf2_rref((const uint64_t*)d_in, batch, m, n, (uint64_t*)d_out, d_rank, d_piv, 0);
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, bytes, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && rank) e = cudaMemcpy(rank, d_rank, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && pivots) e = cudaMemcpy(pivots, d_piv, (size_t)batch * npiv * sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess)
        rc = fail(e == cudaErrorMemoryAllocation ? QCSS_ERR_NOMEM : QCSS_ERR_CUDA, "gf2_rref: %s", cudaGetErrorString(e));
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_rank); cudaFree(d_piv);
    return rc;
}

QCSS_API int qcss_gf2_nullspace_dev(const uint64_t* d_mats, int batch, int m, int n, int max_basis_rows,
                           uint64_t* d_basis, int32_t* d_rank, int32_t* d_overflow, void* stream) {
    if (batch < 0 || m < 0 || n < 0 || max_basis_rows < 0) return fail(QCSS_ERR_INVALID, "negative dimensions");
    if (batch == 0 || m == 0 || n == 0) return QCSS_OK;
    if (!d_mats || (!d_basis && max_basis_rows > 0)) return fail(QCSS_ERR_INVALID, "NULL matrices");
    QCSS_CUDA(launch_gf2_nullspace(d_mats, batch, m, n, max_basis_rows, d_basis, d_rank, d_overflow,
                                   (cudaStream_t)stream));
    return QCSS_OK;
}

QCSS_API int qcss_gf2_nullspace(const uint64_t* mats, int batch, int m, int n, int max_basis_rows, uint64_t* basis,
                       int32_t* rank) {
    if (batch < 0 || m < 0 || n < 0 || max_basis_rows < 0) return fail(QCSS_ERR_INVALID, "negative dimensions");
    if (batch == 0 || m == 0 || n == 0) return QCSS_OK;
    if (!mats || (!basis && max_basis_rows > 0)) return fail(QCSS_ERR_INVALID, "NULL matrices");
    const size_t W = (size_t)(n + 63) / 64, in_bytes = (size_t)batch * m * W * 8;
    const size_t out_bytes = (size_t)batch * max_basis_rows * W * 8;
    void *d_in = nullptr, *d_out = nullptr;
    int32_t* d_rank = nullptr;
    int32_t need = 0;
    int rc = QCSS_OK;
    cudaError_t e = cudaMalloc(&d_in, in_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, out_bytes + 16);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_rank, ((size_t)batch + 1) * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemcpy(d_in, mats, in_bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = launch_gf2_nullspace((const uint64_t*)d_in, batch, m, n, max_basis_rows, (uint64_t*)d_out, d_rank,
                                 d_rank + batch, 0);
    if (e == cudaSuccess) e = cudaMemcpy(&need, d_rank + batch, sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && out_bytes) e = cudaMemcpy(basis, d_out, out_bytes, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && rank) e = cudaMemcpy(rank, d_rank, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess)
        rc = fail(e == cudaErrorMemoryAllocation ? QCSS_ERR_NOMEM : QCSS_ERR_CUDA, "gf2_nullspace: %s", cudaGetErrorString(e));
    else if (need > max_basis_rows)
        rc = fail(QCSS_ERR_INVALID, "gf2_nullspace: a matrix needs %d basis rows, capacity is %d", need, max_basis_rows);
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_rank);
    return rc;
}

QCSS_API int qcss_gf2_solve_dev(const uint64_t* d_mats, const uint64_t* d_rhs, int batch, int m, int n, uint64_t* d_x,
                       int32_t* d_consistent, void* stream) {
    if (batch < 0 || m < 0 || n < 0) return fail(QCSS_ERR_INVALID, "negative dimensions");
    if (batch == 0 || m == 0 || n == 0) return QCSS_OK;
    if (!d_mats || !d_rhs || !d_x) return fail(QCSS_ERR_INVALID, "NULL operands");
    QCSS_CUDA(launch_gf2_solve(d_mats, d_rhs, batch, m, n, d_x, d_consistent, (cudaStream_t)stream));
    return QCSS_OK;
}

QCSS_API int qcss_gf2_solve(const uint64_t* mats, const uint64_t* rhs, int batch, int m, int n, uint64_t* x,
                   int32_t* consistent) {
    if (batch < 0 || m < 0 || n < 0) return fail(QCSS_ERR_INVALID, "negative dimensions");
    if (batch == 0 || m == 0 || n == 0) return QCSS_OK;
    if (!mats || !rhs || !x) return fail(QCSS_ERR_INVALID, "NULL operands");
    const size_t W = (size_t)(n + 63) / 64, Wr = (size_t)(m + 63) / 64;
    const size_t in_bytes = (size_t)batch * m * W * 8, rhs_bytes = (size_t)batch * Wr * 8, x_bytes = (size_t)batch * W * 8;
    void *d_in = nullptr, *d_rhs = nullptr, *d_x = nullptr;
    int32_t* d_ok = nullptr;
    int rc = QCSS_OK;
    cudaError_t e = cudaMalloc(&d_in, in_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&d_rhs, rhs_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&d_x, x_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_ok, (size_t)batch * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemcpy(d_in, mats, in_bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_rhs, rhs, rhs_bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = launch_gf2_solve((const uint64_t*)d_in, (const uint64_t*)d_rhs, batch, m, n, (uint64_t*)d_x, d_ok, 0);
    if (e == cudaSuccess) e = cudaMemcpy(x, d_x, x_bytes, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && consistent) e = cudaMemcpy(consistent, d_ok, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess)
        rc = fail(e == cudaErrorMemoryAllocation ? QCSS_ERR_NOMEM : QCSS_ERR_CUDA, "gf2_solve: %s", cudaGetErrorString(e));
    cudaFree(d_in); cudaFree(d_rhs); cudaFree(d_x); cudaFree(d_ok);
    return rc;
}

// ---- CSS construction numerics ----------------------------------------------------

QCSS_API int qcss_gf2_normalize_dev(uint64_t* d_mats, int batch, int m, int n, int offset, int32_t* d_swaps,
                           int32_t* d_n_swaps, int32_t* d_status, void* stream) {
    if (batch < 0 || m < 0 || n < 0 || offset < 0) return fail(QCSS_ERR_INVALID, "negative dimensions");
    if (n < offset + m) return fail(QCSS_ERR_INVALID, "not enough columns");
    if (batch == 0 || m == 0) return QCSS_OK;
    if (!d_mats || !d_status) return fail(QCSS_ERR_INVALID, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    QCSS_CUDA(cudaMemsetAsync(d_status, 0, (size_t)batch * sizeof(int32_t), st));
    if (d_n_swaps) QCSS_CUDA(cudaMemsetAsync(d_n_swaps, 0, (size_t)batch * sizeof(int32_t), st));
    QCSS_CUDA(launch_gf2_normalize(d_mats, batch, m, n, offset, nullptr, 0, d_swaps, d_n_swaps, d_status,
                                   QCSS_FORM_DEPENDENT_ROWS, st));
    return QCSS_OK;
}

QCSS_API int qcss_gf2_normalize(const uint64_t* mats, int batch, int m, int n, int offset, uint64_t* out, int32_t* swaps,
                       int32_t* n_swaps, int32_t* status) {
    if (batch < 0 || m < 0 || n < 0 || offset < 0) return fail(QCSS_ERR_INVALID, "negative dimensions");
    if (n < offset + m) return fail(QCSS_ERR_INVALID, "not enough columns");
    if (batch == 0 || m == 0) return QCSS_OK;
    if (!mats || !out || !status) return fail(QCSS_ERR_INVALID, "NULL argument");
    const size_t W = (size_t)(n + 63) / 64, bytes = (size_t)batch * m * W * 8;
    const size_t swap_bytes = (size_t)batch * n * 2 * sizeof(int32_t), cnt_bytes = (size_t)batch * sizeof(int32_t);
    void* d_mat = nullptr;
    int32_t *d_swaps = nullptr, *d_cnt = nullptr, *d_status = nullptr;
    int rc = QCSS_OK;
    cudaError_t e = cudaMalloc(&d_mat, bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_swaps, swap_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_cnt, cnt_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_status, cnt_bytes);
    if (e == cudaSuccess) e = cudaMemcpy(d_mat, mats, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(d_swaps, 0xFF, swap_bytes);
    if (e == cudaSuccess) {
        rc = qcss_gf2_normalize_dev((uint64_t*)d_mat, batch, m, n, offset, d_swaps, d_cnt, d_status, nullptr);
        if (rc == QCSS_OK) e = cudaDeviceSynchronize();
    }
    if (rc == QCSS_OK) {
        if (e == cudaSuccess) e = cudaMemcpy(out, d_mat, bytes, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && swaps) e = cudaMemcpy(swaps, d_swaps, swap_bytes, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && n_swaps) e = cudaMemcpy(n_swaps, d_cnt, cnt_bytes, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(status, d_status, cnt_bytes, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess)
            rc = fail(e == cudaErrorMemoryAllocation ? QCSS_ERR_NOMEM : QCSS_ERR_CUDA, "gf2_normalize: %s",
                      cudaGetErrorString(e));
    }
    cudaFree(d_mat); cudaFree(d_swaps); cudaFree(d_cnt); cudaFree(d_status);
    return rc;
}

QCSS_API int qcss_css_standard_form(const uint64_t* H1, int r1, const uint64_t* H2, int r2, int n, uint64_t* out1,
                           uint64_t* out2, int32_t* swaps, int32_t* n_swaps, int32_t* status) {
    if (r1 < 1 || r2 < 1 || n < 1) return fail(QCSS_ERR_INVALID, "empty parity check");
    if (!H1 || !H2 || !out1 || !out2 || !status) return fail(QCSS_ERR_INVALID, "NULL argument");
    const size_t W = (size_t)(n + 63) / 64, b1 = (size_t)r1 * W * 8, b2 = (size_t)r2 * W * 8;
    const size_t swap_bytes = (size_t)n * 2 * sizeof(int32_t);
    void *d1 = nullptr, *d2 = nullptr;
    int32_t *d_swaps = nullptr, *d_meta = nullptr;          // meta[0] = status, meta[1] = swap count
    int rc = QCSS_OK;
    cudaError_t e = cudaMalloc(&d1, b1);
    if (e == cudaSuccess) e = cudaMalloc(&d2, b2);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_swaps, swap_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_meta, 2 * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemcpy(d1, H1, b1, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d2, H2, b2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(d_swaps, 0xFF, swap_bytes);
    if (e == cudaSuccess) e = cudaMemset(d_meta, 0, 2 * sizeof(int32_t));
    // css_code.py:47-49, then :55-61: H1 at offset 0 with its swaps replayed on H2, H2 at offset r1 with its
    // swaps replayed on H1.  Each stage is skipped once the status word is non-zero; the reference's
    // "not enough columns" checks (:811-812) sit between the stages, in its order.
    int few_columns = 0;
    if (e == cudaSuccess)
        e = launch_css_condition((const uint64_t*)d1, r1, (const uint64_t*)d2, r2, n, d_meta, QCSS_FORM_NOT_CSS, 0);
    if (n < r1) few_columns = QCSS_FORM_FEW_COLUMNS_C1;
    if (e == cudaSuccess && !few_columns)
        e = launch_gf2_normalize((uint64_t*)d1, 1, r1, n, 0, (uint64_t*)d2, r2, d_swaps, d_meta + 1, d_meta,
                                 QCSS_FORM_DEPENDENT_ROWS_C1, 0);
    if (!few_columns && n < r1 + r2) few_columns = QCSS_FORM_FEW_COLUMNS_C2;
    if (e == cudaSuccess && !few_columns)
        e = launch_gf2_normalize((uint64_t*)d2, 1, r2, n, r1, (uint64_t*)d1, r1, d_swaps, d_meta + 1, d_meta,
                                 QCSS_FORM_DEPENDENT_ROWS_C2, 0);
    int32_t meta[2] = {0, 0};
    if (e == cudaSuccess) e = cudaMemcpy(meta, d_meta, sizeof(meta), cudaMemcpyDeviceToHost);
    if (meta[0] == 0) meta[0] = few_columns;
    if (e == cudaSuccess) e = cudaMemcpy(out1, d1, b1, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(out2, d2, b2, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && swaps) e = cudaMemcpy(swaps, d_swaps, swap_bytes, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess)
        rc = fail(e == cudaErrorMemoryAllocation ? QCSS_ERR_NOMEM : QCSS_ERR_CUDA, "css_standard_form: %s",
                  cudaGetErrorString(e));
    *status = meta[0];
    if (n_swaps) *n_swaps = meta[1];
    cudaFree(d1); cudaFree(d2); cudaFree(d_swaps); cudaFree(d_meta);
    return rc;
}

// ---- GPU-assisted syndrome table -----------------------------------------------

QCSS_API int qcss_table_build(int n, int m, const uint8_t* H, int64_t max_entries, qcss_table** out, int* t,
                     int64_t* n_entries) {
    if (!H || !out || !t || !n_entries) return fail(QCSS_ERR_INVALID, "NULL argument");
    if (n < 1 || n > 64) return fail(QCSS_ERR_UNSUPPORTED, "syndrome table on the device needs 1 <= n <= 64 (got %d)", n);
    if (m < 1 || m > 62) return fail(QCSS_ERR_UNSUPPORTED, "syndrome table on the device needs 1 <= m <= 62 (got %d)", m);
    if (max_entries < 1) return fail(QCSS_ERR_INVALID, "max_entries must be positive");
    TableBuild* tb = nullptr;
    const char* why = "";
    const cudaError_t e = table_build(n, m, H, max_entries, &tb, &why);
    if (e != cudaSuccess)
        return fail(e == cudaErrorMemoryAllocation ? QCSS_ERR_NOMEM : QCSS_ERR_CUDA, "table_build: %s %s",
                    cudaGetErrorString(e), why);
    *out = reinterpret_cast<qcss_table*>(tb);
    *t = table_t(tb);
    *n_entries = table_count(tb);
    return QCSS_OK;
}

QCSS_API int qcss_table_read(const qcss_table* table, int64_t* keys, uint64_t* supports) {
    if (!table || !keys || !supports) return fail(QCSS_ERR_INVALID, "NULL argument");
    QCSS_CUDA(table_read(reinterpret_cast<const TableBuild*>(table), keys, supports));
    return QCSS_OK;
}

QCSS_API int qcss_table_destroy(qcss_table* table) {
    table_free(reinterpret_cast<TableBuild*>(table));
    return QCSS_OK;
}

}  // extern "C"

// ---- per-code kernel specialisation -------------------------------------------------------------
// The static kernel family (H, L and the m <= 5 truth tables as compile-time constants) is not limited
// to the three descriptors built into the library: qcss_code_spec_source writes the translation unit
// for THIS code (the same descriptor layout tools/gen_named_codes.py emits), the host compiles it with
// nvcc for sm_100a into a shared object, and qcss_code_load_specialized routes the code's launches to it.

namespace {

void emit_side(std::string& out, const char* name, const GenericSide& s, const uint32_t* rows, uint32_t lmask) {
    const bool sliced = s.m <= kSlicedM;
    const int mb = sliced ? s.m : (s.m <= 8 ? 8 : 16);
    char line[256];
    out += "struct "; out += name; out += " {\n";
    snprintf(line, sizeof(line), "    static constexpr int N = %d, M = %d, MB = %d;\n", s.n, s.m, mb); out += line;
    snprintf(line, sizeof(line), "    static constexpr bool kSliced = %s, kHasMiss = %s;\n", sliced ? "true" : "false",
             s.has_miss ? "true" : "false"); out += line;
    snprintf(line, sizeof(line), "    static constexpr uint32_t kL = 0x%xu, kTtFlip = 0x%xu, kTtMiss = 0x%xu;\n", lmask,
             sliced ? s.tt_flip : 0u, sliced ? s.tt_miss : 0u); out += line;
    out += "    QCSS_HD static constexpr uint32_t row(int t) {\n        constexpr uint32_t r[MB] = {";
    for (int t = 0; t < mb; ++t) { snprintf(line, sizeof(line), "%s0x%xu", t ? ", " : "", t < s.m ? rows[t] : 0u); out += line; }
    out += "};\n        return r[t];\n    }\n";
    out += "    QCSS_HD static constexpr uint32_t tt_corr(int j) {\n        constexpr uint32_t r[N] = {";
    for (int j = 0; j < s.n; ++j) { snprintf(line, sizeof(line), "%s0x%xu", j ? ", " : "", sliced ? s.tt_corr[j] : 0u); out += line; }
    out += "};\n        return r[j];\n    }\n};\n\n";
}

}  // namespace

extern "C" {

QCSS_API int qcss_code_spec_source(const qcss_code* c, char* buf, int64_t cap, int64_t* needed) {
    if (!c || !needed) return fail(QCSS_ERR_INVALID, "bad arguments");
    if (c->k > 1) return fail(QCSS_ERR_UNSUPPORTED, "specialisation covers k = 1");
    if (!c->small || c->side_x.mode == kModeNone || c->side_z.mode == kModeNone)
        return fail(QCSS_ERR_UNSUPPORTED, "specialisation needs a decodable code (n <= %d, m <= %d, both tables)", kMaxN, kMaxM);
    std::string out = "// GENERATED by qcss_code_spec_source -- kernels specialised for one code.\n"
                      "#include \"small_common.cuh\"\n\nnamespace qcss {\nnamespace spec {\nnamespace {   // internal linkage: "
                      "several specialised objects can live in one process\n\n";
    emit_side(out, "Spec_X", c->side_x, c->rows_x, c->lmask_x);
    emit_side(out, "Spec_Z", c->side_z, c->rows_z, c->lmask_z);
    out += "}  // namespace\n}  // namespace spec\n}  // namespace qcss\n\n"
           "extern \"C\" __attribute__((visibility(\"default\"))) int qcss_spec_abi(void) {\n"
           "    return (int)sizeof(qcss::SmallLaunch) * 1000 + (int)(sizeof(qcss::GenericSide) % 1000);\n}\n"
           "extern \"C\" __attribute__((visibility(\"default\"))) int qcss_spec_launch(const void* l, void* stream) {\n"
           "    return (int)qcss::small::launch_named<qcss::spec::Spec_X, qcss::spec::Spec_Z>(\n"
           "        *static_cast<const qcss::SmallLaunch*>(l), static_cast<cudaStream_t>(stream));\n}\n";
    *needed = (int64_t)out.size() + 1;
    if (buf != nullptr && cap >= *needed) memcpy(buf, out.c_str(), out.size() + 1);
    return QCSS_OK;
}

QCSS_API int qcss_code_load_specialized(qcss_code* c, const char* so_path, const char* tag) {
    if (!c || !so_path) return fail(QCSS_ERR_INVALID, "bad arguments");
    void* dl = dlopen(so_path, RTLD_NOW | RTLD_LOCAL);
    if (!dl) return fail(QCSS_ERR_INVALID, "cannot load %s: %s", so_path, dlerror());
    auto abi = reinterpret_cast<int (*)(void)>(dlsym(dl, "qcss_spec_abi"));
    auto fn = reinterpret_cast<int (*)(const void*, void*)>(dlsym(dl, "qcss_spec_launch"));
    const int want = (int)sizeof(SmallLaunch) * 1000 + (int)(sizeof(GenericSide) % 1000);
    if (!abi || !fn || abi() != want) {
        dlclose(dl);
        return fail(QCSS_ERR_INVALID, "%s was not built against this library's headers", so_path);
    }
    if (c->spec_dl) dlclose(c->spec_dl);
    c->spec_dl = dl;
    c->spec_launch = fn;
    snprintf(c->spec_tag, sizeof(c->spec_tag), "%s", tag ? tag : "");
    return QCSS_OK;
}

}  // extern "C"

// ---- per-syndrome histograms ------------------------------------------------

extern "C" {

QCSS_API int qcss_syndrome_hist_dev(qcss_code* c, int which, const uint64_t* d_e, int64_t e_stride, int64_t shots,
                           uint64_t* d_hist, void* stream) {
    if (!c) return fail(QCSS_ERR_INVALID, "code is NULL");
    if (which != 1 && which != 2) return fail(QCSS_ERR_INVALID, "which must be 1 or 2");
    if (!d_hist) return fail(QCSS_ERR_INVALID, "NULL histogram");
    const int m = (which == 1) ? c->m1 : c->m2;
    if (m < 1 || m > 24) return fail(QCSS_ERR_UNSUPPORTED, "syndrome histograms cover 1 <= m <= 24 (got %d)", m);
    if (shots <= 0) return shots == 0 ? QCSS_OK : fail(QCSS_ERR_INVALID, "shots must be >= 0");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t s_stride = ((shots + 127) / 128) * 2;                  // uint64 words per syndrome plane
    void* d_s = nullptr;
    QCSS_CUDA(cudaMallocAsync(&d_s, (size_t)m * s_stride * 8, st));
    int rc = launch_syndrome(c, which, d_e, e_stride, shots, (uint64_t*)d_s, s_stride, st);
    cudaError_t e = cudaSuccess;
    if (!rc)
        e = launch_syndrome_hist((const uint32_t*)d_s, s_stride * 2, m, (shots + 31) / 32, tail_mask_for(shots),
                                 (unsigned long long*)d_hist, st);
    cudaFreeAsync(d_s, st);
    if (rc) return rc;
    QCSS_CUDA(e);
    return QCSS_OK;
}

QCSS_API int qcss_syndrome_hist(qcss_code* c, int which, const uint64_t* e_planes, int64_t e_stride, int64_t shots,
                       uint64_t* hist) {
    if (!c) return fail(QCSS_ERR_INVALID, "code is NULL");
    if (which != 1 && which != 2) return fail(QCSS_ERR_INVALID, "which must be 1 or 2");
    if (!e_planes || !hist) return fail(QCSS_ERR_INVALID, "NULL argument");
    const int m = (which == 1) ? c->m1 : c->m2;
    if (m < 1 || m > 24) return fail(QCSS_ERR_UNSUPPORTED, "syndrome histograms cover 1 <= m <= 24 (got %d)", m);
    int rc;
    if ((rc = check_host_stride(e_stride, shots, "error planes"))) return rc;
    if ((rc = ensure_streams(c))) return rc;
    const size_t eb = (size_t)c->n * e_stride * 8, hb = ((size_t)1 << m) * 8;
    QCSS_CUDA(c->buf_a.reserve(eb));
    QCSS_CUDA(c->buf_c.reserve(hb));
    QCSS_CUDA(cudaMemcpyAsync(c->buf_a.p, e_planes, eb, cudaMemcpyHostToDevice, c->stream));
    QCSS_CUDA(cudaMemsetAsync(c->buf_c.p, 0, hb, c->stream));
    rc = qcss_syndrome_hist_dev(c, which, (const uint64_t*)c->buf_a.p, e_stride, shots, (uint64_t*)c->buf_c.p, c->stream);
    if (rc) { cudaStreamSynchronize(c->stream); return rc; }   // pending copies must not outlive the call
    QCSS_CUDA(cudaMemcpyAsync(hist, c->buf_c.p, hb, cudaMemcpyDeviceToHost, c->stream));
    QCSS_CUDA(cudaStreamSynchronize(c->stream));
    return QCSS_OK;
}

}  // extern "C"

// ---- host data formats (the reference passes numpy arrays) ------------------------

namespace {

int check_elem(int elem_bytes) {
    if (elem_bytes != 1 && elem_bytes != 8)
        return fail(QCSS_ERR_INVALID, "elem_bytes must be 1 (uint8) or 8 (int64, numpy dtype='int'), got %d", elem_bytes);
    return QCSS_OK;
}

// shots per chunk of the streaming entry points: ~32 MB of raw rows per Pauli type, whole 1024-shot blocks
int64_t chunk_shots_for(int n, int elem_bytes, int64_t shots) {
    int64_t cs = ((int64_t)(32u << 20) / ((int64_t)n * elem_bytes)) & ~(int64_t)1023;
    if (cs < 1024) cs = 1024;
    const int64_t all = (shots + 1023) & ~(int64_t)1023;
    return cs < all ? cs : all;
}

__global__ void k_events_finish(unsigned long long* tally, const unsigned long long* aux, unsigned long long shots,
                                uint32_t f0x, uint32_t f0z, int32_t* status) {
    // shots without any event have the zero syndrome: they take the table entry of key 0 (normally "no flip, hit")
    const unsigned long long quiet = shots - aux[0];
    if (quiet != 0 && (f0x | f0z) != 0) {
        tally[1] += quiet * (f0x & 1u);
        tally[2] += quiet * (f0z & 1u);
        tally[3] += quiet * ((f0x | f0z) & 1u);
        tally[4] += quiet * ((f0x >> 1) & 1u);
        tally[5] += quiet * ((f0z >> 1) & 1u);
    }
    if (status != nullptr) *status = (int32_t)aux[1];
}

int sparse_ready(qcss_code* c, uint32_t* f0x, uint32_t* f0z) {
    if (!c->small) return fail(QCSS_ERR_UNSUPPORTED, "lookup decode covers n <= %d and m <= %d", kMaxN, kMaxM);
    if (c->side_x.mode == kModeNone || c->side_z.mode == kModeNone)
        return fail(QCSS_ERR_INVALID, "sparse decode needs both syndrome tables");
    if (c->k > 1) return fail(QCSS_ERR_UNSUPPORTED, "sparse decode covers k = 1");
    *f0x = c->fm0_x;
    *f0z = c->fm0_z;
    return QCSS_OK;
}

}  // namespace

extern "C" {

QCSS_API int qcss_pack_shots_dev(const void* d_src, int elem_bytes, int n, int64_t shots, uint64_t* d_planes, int64_t stride,
                        void* stream) {
    int rc;
    if ((rc = check_elem(elem_bytes))) return rc;
    if (n < 1 || shots < 0) return fail(QCSS_ERR_INVALID, "bad dimensions");
    if (!d_src || !d_planes) return fail(QCSS_ERR_INVALID, "NULL argument");
    if (((uintptr_t)d_src & 15u) != 0) return fail(QCSS_ERR_INVALID, "source must be 16-byte aligned");
    if ((rc = check_planes(d_planes, stride, shots, "planes"))) return rc;
    QCSS_CUDA(launch_pack_shots(d_src, elem_bytes, n, shots, (uint32_t*)d_planes, stride * 2, (cudaStream_t)stream));
    return QCSS_OK;
}

QCSS_API int qcss_unpack_planes_dev(const uint64_t* d_planes, int64_t stride, int m, int64_t shots, uint8_t* d_dst, void* stream) {
    if (m < 1 || shots < 0) return fail(QCSS_ERR_INVALID, "bad dimensions");
    if (!d_planes || !d_dst) return fail(QCSS_ERR_INVALID, "NULL argument");
    if (((uintptr_t)d_dst & 3u) != 0) return fail(QCSS_ERR_INVALID, "destination must be 4-byte aligned");
    int rc;
    if ((rc = check_planes(d_planes, stride, shots, "planes"))) return rc;
    QCSS_CUDA(launch_unpack_planes((const uint32_t*)d_planes, stride * 2, m, shots, d_dst, (cudaStream_t)stream));
    return QCSS_OK;
}

// Both Pauli types of the same shots in the reference's own layout; tallies only.  Chunks of whole 1024-shot blocks
// flow host -> device on three streams; each chunk is transposed to planes on the device and decoded there.
QCSS_API int qcss_decode_xz_shots(qcss_code* c, const void* ex, const void* ez, int elem_bytes, int64_t shots, qcss_tally* tally) {
    if (!c || !tally) return fail(QCSS_ERR_INVALID, "code or tally is NULL");
    if (!ex || !ez) return fail(QCSS_ERR_INVALID, "NULL errors");
    if (shots < 0) return fail(QCSS_ERR_INVALID, "shots must be >= 0");
    int rc;
    if ((rc = check_elem(elem_bytes))) return rc;
    if (!c->small) return fail(QCSS_ERR_UNSUPPORTED, "lookup decode covers n <= %d and m <= %d", kMaxN, kMaxM);
    if ((rc = ensure_streams(c))) return rc;
    QCSS_CUDA(c->tally.reserve(8 * sizeof(uint64_t)));
    QCSS_CUDA(cudaMemsetAsync(c->tally.p, 0, 8 * sizeof(uint64_t), c->stream));
    QCSS_CUDA(cudaStreamSynchronize(c->stream));
    int64_t cs = chunk_shots_for(c->n, elem_bytes, shots);
    const size_t row_bytes = (size_t)c->n * elem_bytes;
    // 0/1 rows of a low-rate batch are almost all zero words (uint8, n = 7, p = 1e-3: 99 %): whole groups of 16384 shots
    // (a whole number of 2048-word blocks for any row size) go through the compacting pipeline, the rest as plain copies
    constexpr int64_t kGroup = 16384;
    const int64_t main_shots = shots / kGroup * kGroup;
    int threads = 0;
    if (main_shots > 0 && cs >= kGroup)
        threads = compacting_threads((const uint64_t*)ex, (const uint64_t*)ez, (int64_t)((size_t)main_shots * row_bytes / 8));
    if (threads) cs = cs / kGroup * kGroup;
    const int64_t cwords = cs / 64;                                   // uint64 words per plane per chunk (even)
    for (int i = 0; i < kSlots && shots > 0; ++i) {
        QCSS_CUDA(c->slot_rx[i].reserve((size_t)cs * row_bytes));
        QCSS_CUDA(c->slot_rz[i].reserve((size_t)cs * row_bytes));
        QCSS_CUDA(c->slot_x[i].reserve((size_t)c->n * cwords * 8));
        QCSS_CUDA(c->slot_z[i].reserve((size_t)c->n * cwords * 8));
    }
    auto consume = [&](int slot, int64_t part, cudaStream_t st) -> int {       // raw rows of a slot -> planes -> tallies
        QCSS_CUDA(launch_pack_shots(c->slot_rx[slot].p, elem_bytes, c->n, part, (uint32_t*)c->slot_x[slot].p, cwords * 2, st));
        QCSS_CUDA(launch_pack_shots(c->slot_rz[slot].p, elem_bytes, c->n, part, (uint32_t*)c->slot_z[slot].p, cwords * 2, st));
        qcss_decode_io io;
        memset(&io, 0, sizeof(io));
        io.ex = (const uint64_t*)c->slot_x[slot].p;
        io.ez = (const uint64_t*)c->slot_z[slot].p;
        io.e_stride = cwords;
        io.tally = (uint64_t*)c->tally.p;
        return launch_decode(c, &io, part, st);
    };
    int64_t first_plain = 0;
    c->last_h2d_bytes = 0;
    c->last_host_threads = 0;
    if (threads) {
        const int64_t chunk_words = (int64_t)((size_t)cs * row_bytes / 8), total_words = (int64_t)((size_t)main_shots * row_bytes / 8);
        rc = zs_pipeline(c, (const uint64_t*)ex, (const uint64_t*)ez, total_words, 1, total_words, chunk_words, c->slot_rx, c->slot_rz,
                         chunk_words, threads, [&](int64_t ci, int slot, int64_t, int64_t, cudaStream_t st) {
                             const int64_t s0 = ci * cs;
                             return consume(slot, main_shots - s0 < cs ? main_shots - s0 : cs, st);
                         });
        if (rc) return rc;
        first_plain = main_shots;
    }
    int slot = 0;
    for (int64_t s0 = first_plain; s0 < shots; s0 += cs, slot = (slot + 1) % kSlots) {
        const int64_t part = shots - s0 < cs ? shots - s0 : cs;
        cudaStream_t st = c->slot_stream[slot];
        QCSS_CUDA(cudaMemcpyAsync(c->slot_rx[slot].p, (const uint8_t*)ex + (size_t)s0 * row_bytes, (size_t)part * row_bytes,
                                  cudaMemcpyHostToDevice, st));
        QCSS_CUDA(cudaMemcpyAsync(c->slot_rz[slot].p, (const uint8_t*)ez + (size_t)s0 * row_bytes, (size_t)part * row_bytes,
                                  cudaMemcpyHostToDevice, st));
        c->last_h2d_bytes += 2 * part * (int64_t)row_bytes;
        rc = consume(slot, part, st);
        if (rc) {
            for (int i = 0; i < kSlots; ++i) cudaStreamSynchronize(c->slot_stream[i]);
            return rc;
        }
    }
    for (int i = 0; i < kSlots; ++i) QCSS_CUDA(cudaStreamSynchronize(c->slot_stream[i]));
    uint64_t h[6];
    QCSS_CUDA(cudaMemcpy(h, c->tally.p, sizeof(h), cudaMemcpyDeviceToHost));
    tally_from(h, shots, tally);
    return QCSS_OK;
}

// One Pauli type in the reference's layout: s_out / corr_out are (shots, m) / (shots, n) bytes, flip_out / miss_out
// (shots,) bytes; any of them may be NULL.  decode = false: syndromes only (works for codes of any size).
static int run_shots(qcss_code* c, int which, const void* e, int elem_bytes, int64_t shots, uint8_t* s_out, uint8_t* corr_out,
                     uint8_t* flip_out, uint8_t* miss_out, qcss_tally* tally, bool decode) {
    if (!c) return fail(QCSS_ERR_INVALID, "code is NULL");
    if (which != 1 && which != 2) return fail(QCSS_ERR_INVALID, "which must be 1 or 2");
    if (!e) return fail(QCSS_ERR_INVALID, "NULL errors");
    if (shots < 0) return fail(QCSS_ERR_INVALID, "shots must be >= 0");
    int rc;
    if ((rc = check_elem(elem_bytes))) return rc;
    if ((rc = ensure_streams(c))) return rc;
    const int m = (which == 1) ? c->m1 : c->m2;
    if (decode && !c->small) return fail(QCSS_ERR_UNSUPPORTED, "lookup decode covers n <= %d and m <= %d", kMaxN, kMaxM);
    QCSS_CUDA(c->tally.reserve(8 * sizeof(uint64_t)));
    QCSS_CUDA(cudaMemsetAsync(c->tally.p, 0, 8 * sizeof(uint64_t), c->stream));
    const int64_t cs = chunk_shots_for(c->n > m ? c->n : m, elem_bytes, shots);
    const int64_t cwords = cs / 64;
    const size_t row_bytes = (size_t)c->n * elem_bytes;
    cudaStream_t st = c->stream;
    const bool tiles = !c->small && !((which == 1) ? c->dense1 : c->dense2) &&
                       syndrome_tiles_supported((which == 1) ? c->sp1 : c->sp2);
    if (shots > 0) {
        QCSS_CUDA(c->buf_a.reserve((size_t)cs * row_bytes));                 // raw rows
        QCSS_CUDA(c->buf_b.reserve((size_t)c->n * cwords * 8));              // error planes
        QCSS_CUDA(c->buf_c.reserve((size_t)(m > c->n ? m : c->n) * cwords * 8));   // syndrome or correction planes
        QCSS_CUDA(c->buf_d.reserve((size_t)2 * cwords * 8));                 // flip, miss planes
        QCSS_CUDA(c->buf_e.reserve((size_t)cs * (size_t)(m > c->n ? m : c->n)));   // bytes going back
    }
    for (int64_t s0 = 0; s0 < shots; s0 += cs) {
        const int64_t part = shots - s0 < cs ? shots - s0 : cs;
        QCSS_CUDA(cudaMemcpyAsync(c->buf_a.p, (const uint8_t*)e + (size_t)s0 * row_bytes, (size_t)part * row_bytes,
                                  cudaMemcpyHostToDevice, st));
        if (!decode && tiles) {
            // sparse any-size codes: the rows are transposed straight into the tile-major layout the bulk-copy ring
            // streams at 90-100 % of the HBM copy rate, and the syndromes come back out of it
            QCSS_CUDA(launch_pack_shots(c->buf_a.p, elem_bytes, c->n, part, (uint32_t*)c->buf_b.p, 0, st));
            rc = launch_syndrome_tiles_checked(c, which, (const uint64_t*)c->buf_b.p, part, (uint64_t*)c->buf_c.p, st);
            if (rc) { cudaStreamSynchronize(st); return rc; }
            QCSS_CUDA(launch_unpack_planes((const uint32_t*)c->buf_c.p, 0, m, part, (uint8_t*)c->buf_e.p, st));
            QCSS_CUDA(cudaMemcpyAsync(s_out + (size_t)s0 * m, c->buf_e.p, (size_t)part * m, cudaMemcpyDeviceToHost, st));
            continue;
        }
        QCSS_CUDA(launch_pack_shots(c->buf_a.p, elem_bytes, c->n, part, (uint32_t*)c->buf_b.p, cwords * 2, st));
        if (!decode) {
            rc = launch_syndrome(c, which, (const uint64_t*)c->buf_b.p, cwords, part, (uint64_t*)c->buf_c.p, cwords, st);
            if (rc) { cudaStreamSynchronize(st); return rc; }
            QCSS_CUDA(launch_unpack_planes((const uint32_t*)c->buf_c.p, cwords * 2, m, part, (uint8_t*)c->buf_e.p, st));
            QCSS_CUDA(cudaMemcpyAsync(s_out + (size_t)s0 * m, c->buf_e.p, (size_t)part * m, cudaMemcpyDeviceToHost, st));
            continue;
        }
        qcss_decode_io io;
        memset(&io, 0, sizeof(io));
        io.e_stride = io.c_stride = cwords;
        io.tally = (uint64_t*)c->tally.p;
        uint64_t* flip = (uint64_t*)c->buf_d.p;
        uint64_t* miss = flip + cwords;
        QCSS_CUDA(cudaMemsetAsync(c->buf_d.p, 0, (size_t)2 * cwords * 8, st));
        if (corr_out) QCSS_CUDA(cudaMemsetAsync(c->buf_c.p, 0, (size_t)c->n * cwords * 8, st));
        if (which == 2) { io.ex = (const uint64_t*)c->buf_b.p; io.corr_x = corr_out ? (uint64_t*)c->buf_c.p : nullptr; io.flip_x = flip; io.miss_x = miss; }
        else            { io.ez = (const uint64_t*)c->buf_b.p; io.corr_z = corr_out ? (uint64_t*)c->buf_c.p : nullptr; io.flip_z = flip; io.miss_z = miss; }
        rc = launch_decode(c, &io, part, st);
        if (rc) { cudaStreamSynchronize(st); return rc; }
        if (corr_out) {
            QCSS_CUDA(launch_unpack_planes((const uint32_t*)c->buf_c.p, cwords * 2, c->n, part, (uint8_t*)c->buf_e.p, st));
            QCSS_CUDA(cudaMemcpyAsync(corr_out + (size_t)s0 * c->n, c->buf_e.p, (size_t)part * c->n, cudaMemcpyDeviceToHost, st));
        }
        if (flip_out) {
            QCSS_CUDA(launch_unpack_planes((const uint32_t*)flip, cwords * 2, 1, part, (uint8_t*)c->buf_e.p, st));
            QCSS_CUDA(cudaMemcpyAsync(flip_out + s0, c->buf_e.p, (size_t)part, cudaMemcpyDeviceToHost, st));
        }
        if (miss_out) {
            QCSS_CUDA(launch_unpack_planes((const uint32_t*)miss, cwords * 2, 1, part, (uint8_t*)c->buf_e.p, st));
            QCSS_CUDA(cudaMemcpyAsync(miss_out + s0, c->buf_e.p, (size_t)part, cudaMemcpyDeviceToHost, st));
        }
    }
    uint64_t h[6] = {0, 0, 0, 0, 0, 0};
    QCSS_CUDA(cudaMemcpyAsync(h, c->tally.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    QCSS_CUDA(cudaStreamSynchronize(st));
    if (tally) tally_from(h, shots, tally);
    return QCSS_OK;
}

QCSS_API int qcss_syndrome_shots(qcss_code* c, int which, const void* e, int elem_bytes, int64_t shots, uint8_t* s_out) {
    if (!s_out) return fail(QCSS_ERR_INVALID, "NULL output");
    return run_shots(c, which, e, elem_bytes, shots, s_out, nullptr, nullptr, nullptr, nullptr, false);
}

QCSS_API int qcss_decode_shots(qcss_code* c, int which, const void* e, int elem_bytes, int64_t shots, uint8_t* corr_out,
                      uint8_t* flip_out, uint8_t* miss_out, qcss_tally* tally) {
    return run_shots(c, which, e, elem_bytes, shots, nullptr, corr_out, flip_out, miss_out, tally, true);
}

// ---- sparse batches ----------------------------------------------------------------------------------------------

QCSS_API int qcss_decode_xz_sparse_dev(qcss_code* c, const uint64_t* d_events, int64_t n_events, int64_t shots, uint64_t* d_tally,
                              int32_t* d_status, void* stream) {
    if (!c || !d_tally) return fail(QCSS_ERR_INVALID, "code or tally is NULL");
    if (n_events < 0 || shots < 0) return fail(QCSS_ERR_INVALID, "negative count");
    if (n_events > 0 && !d_events) return fail(QCSS_ERR_INVALID, "NULL events");
    uint32_t f0x, f0z;
    int rc = sparse_ready(c, &f0x, &f0z);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* aux = nullptr;
    QCSS_CUDA(cudaMallocAsync((void**)&aux, 2 * sizeof(unsigned long long), st));
    cudaError_t e = cudaMemsetAsync(aux, 0, 2 * sizeof(unsigned long long), st);
    if (e == cudaSuccess)
        e = launch_decode_events(c->side_x, c->rows_x, c->lmask_x, c->side_z, c->rows_z, c->lmask_z,
                                 (const unsigned long long*)d_events, n_events, shots, (unsigned long long*)d_tally, aux, st);
    if (e == cudaSuccess) {
        k_events_finish<<<1, 1, 0, st>>>((unsigned long long*)d_tally, aux, (unsigned long long)shots, f0x, f0z, d_status);
        e = cudaGetLastError();
    }
    cudaFreeAsync(aux, st);
    QCSS_CUDA(e);
    return QCSS_OK;
}

QCSS_API int qcss_decode_xz_sparse(qcss_code* c, const uint64_t* events, int64_t n_events, int64_t shots, qcss_tally* tally) {
    if (!c || !tally) return fail(QCSS_ERR_INVALID, "code or tally is NULL");
    if (n_events < 0 || shots < 0) return fail(QCSS_ERR_INVALID, "negative count");
    if (n_events > 0 && !events) return fail(QCSS_ERR_INVALID, "NULL events");
    uint32_t f0x, f0z;
    int rc = sparse_ready(c, &f0x, &f0z);
    if (rc) return rc;
    if ((rc = ensure_streams(c))) return rc;
    QCSS_CUDA(c->tally.reserve(8 * sizeof(uint64_t)));                   // 6 tallies, event-shot count, status bits
    QCSS_CUDA(cudaMemsetAsync(c->tally.p, 0, 8 * sizeof(uint64_t), c->stream));
    QCSS_CUDA(cudaStreamSynchronize(c->stream));
    const int64_t chunk = (int64_t)4 << 20;                             // events per chunk (32 MB)
    const int64_t first = n_events < chunk ? n_events : chunk;
    for (int i = 0; i < kSlots && n_events > 0; ++i) QCSS_CUDA(c->slot_rx[i].reserve((size_t)first * 8));
    unsigned long long* d_tally = (unsigned long long*)c->tally.p;
    int slot = 0;
    for (int64_t pos = 0; pos < n_events; slot = (slot + 1) % kSlots) {
        int64_t end = pos + chunk < n_events ? pos + chunk : n_events;
        // a shot's events stay in one chunk: move the cut back to the start of the shot it would split
        while (end < n_events && end > pos && (events[end] >> 18) == (events[end - 1] >> 18)) --end;
        if (end == pos) {
            for (int i = 0; i < kSlots; ++i) cudaStreamSynchronize(c->slot_stream[i]);
            return fail(QCSS_ERR_INVALID, "shot %llu has more than %lld events", (unsigned long long)(events[pos] >> 18), (long long)chunk);
        }
        if (end < n_events && (events[end - 1] >> 18) > (events[end] >> 18)) {
            for (int i = 0; i < kSlots; ++i) cudaStreamSynchronize(c->slot_stream[i]);
            return fail(QCSS_ERR_INVALID, "events must be sorted by shot (event %lld)", (long long)end);
        }
        cudaStream_t st = c->slot_stream[slot];
        QCSS_CUDA(cudaMemcpyAsync(c->slot_rx[slot].p, events + pos, (size_t)(end - pos) * 8, cudaMemcpyHostToDevice, st));
        QCSS_CUDA(launch_decode_events(c->side_x, c->rows_x, c->lmask_x, c->side_z, c->rows_z, c->lmask_z,
                                       (const unsigned long long*)c->slot_rx[slot].p, end - pos, shots, d_tally, d_tally + 6, st));
        pos = end;
    }
    for (int i = 0; i < kSlots; ++i) QCSS_CUDA(cudaStreamSynchronize(c->slot_stream[i]));
    k_events_finish<<<1, 1, 0, c->stream>>>(d_tally, d_tally + 6, (unsigned long long)shots, f0x, f0z, nullptr);
    QCSS_CUDA(cudaGetLastError());
    uint64_t h[8];
    QCSS_CUDA(cudaMemcpyAsync(h, c->tally.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    QCSS_CUDA(cudaStreamSynchronize(c->stream));
    if (h[7] & 1) return fail(QCSS_ERR_INVALID, "an event names a qubit >= n, a shot >= shots or Pauli type 0");
    if (h[7] & 2) return fail(QCSS_ERR_INVALID, "events must be sorted by shot");
    tally_from(h, shots, tally);
    return QCSS_OK;
}

QCSS_API int qcss_events_from_planes_dev(qcss_code* c, const uint64_t* d_ex, const uint64_t* d_ez, int64_t e_stride, int64_t shots,
                                int64_t first_shot, uint64_t* d_events, int64_t capacity, uint64_t* d_count, void* stream) {
    if (!c || !d_count) return fail(QCSS_ERR_INVALID, "NULL argument");
    if (capacity < 0 || first_shot < 0 || (capacity > 0 && !d_events)) return fail(QCSS_ERR_INVALID, "bad event buffer");
    if (c->n > 65535 || (uint64_t)(first_shot + shots) >> 46) return fail(QCSS_ERR_UNSUPPORTED, "event fields overflow");
    int rc;
    if ((rc = check_planes(d_ex, e_stride, shots, "ex planes"))) return rc;
    if ((rc = check_planes(d_ez, e_stride, shots, "ez planes"))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int ctas = 1184;
    unsigned long long* work = nullptr;
    QCSS_CUDA(cudaMallocAsync((void**)&work, (size_t)(ctas + 1) * sizeof(unsigned long long), st));
    const cudaError_t e = launch_events_from_planes((const uint32_t*)d_ex, (const uint32_t*)d_ez, c->n, e_stride * 2,
                                                    (shots + 31) / 32, tail_mask_for(shots), first_shot,
                                                    (unsigned long long*)d_events, capacity, (unsigned long long*)d_count, work,
                                                    ctas, st);
    cudaFreeAsync(work, st);
    QCSS_CUDA(e);
    return QCSS_OK;
}

}  // extern "C"

// ---- in-process specialisation with NVRTC -----------------------------------------------------
// The static kernel family for ANY decodable code without a toolkit on the box: the descriptor translation unit is
// compiled by libnvrtc inside this process from headers EMBEDDED in the library (csrc/build/embedded_headers.inc,
// generated by build.py from core.cuh / decode.cuh / small_common.cuh, which compile under __CUDACC_RTC__ without any
// system header), loaded with cudaLibraryLoadData and launched through cudaKernel_t handles with the same launch
// shapes as small::launch_named.  The cubin is cached on disk by content hash.

#include <nvrtc.h>

#include "build/embedded_headers.inc"

struct SpecKernels {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t fast = nullptr, full = nullptr, s_fast = nullptr, s_full = nullptr, gapq = nullptr;
    bool x_sliced = false, z_sliced = false;
    int gapq_w = 1;
    size_t gapq_smem = 0;
};

void spec_rtc_free(SpecKernels* k) {
    if (!k) return;
    if (k->lib) cudaLibraryUnload(k->lib);
    delete k;
}

namespace {

struct Nvrtc {
    void* dl = nullptr;
    nvrtcResult (*create)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*compile)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*log_size)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*log)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*cubin_size)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*cubin)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*destroy)(nvrtcProgram*) = nullptr;
};

int nvrtc_open(const char* path, Nvrtc* nv) {
    const char* tries[] = {path, "libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
    for (const char* t : tries) {
        if (t == nullptr || *t == 0) continue;
        nv->dl = dlopen(t, RTLD_NOW | RTLD_LOCAL);
        if (nv->dl) break;
    }
    if (!nv->dl) return fail(QCSS_ERR_UNSUPPORTED, "libnvrtc not found (pass its path): %s", dlerror());
#define QCSS_NVRTC_SYM(field, name)                                                   \
    nv->field = reinterpret_cast<decltype(nv->field)>(dlsym(nv->dl, name));            \
    if (!nv->field) return fail(QCSS_ERR_UNSUPPORTED, "libnvrtc lacks %s", name);
    QCSS_NVRTC_SYM(create, "nvrtcCreateProgram")
    QCSS_NVRTC_SYM(compile, "nvrtcCompileProgram")
    QCSS_NVRTC_SYM(log_size, "nvrtcGetProgramLogSize")
    QCSS_NVRTC_SYM(log, "nvrtcGetProgramLog")
    QCSS_NVRTC_SYM(cubin_size, "nvrtcGetCUBINSize")
    QCSS_NVRTC_SYM(cubin, "nvrtcGetCUBIN")
    QCSS_NVRTC_SYM(destroy, "nvrtcDestroyProgram")
#undef QCSS_NVRTC_SYM
    return QCSS_OK;
}

uint64_t fnv1a(const std::string& s, uint64_t h = 1469598103934665603ull) {
    for (unsigned char ch : s) { h ^= ch; h *= 1099511628211ull; }
    return h;
}

// source -> cubin for sm_100a (QCSS_ERR_INVALID with the compiler log on failure)
int nvrtc_compile(const Nvrtc& nv, const std::string& src, std::vector<char>* cubin) {
    nvrtcProgram prog = nullptr;
    if (nv.create(&prog, src.c_str(), "qcss_spec.cu", kEmbeddedHeaderCount, kEmbeddedHeaderText, kEmbeddedHeaderName) != NVRTC_SUCCESS)
        return fail(QCSS_ERR_CUDA, "nvrtcCreateProgram failed");
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-default-device"};
    const nvrtcResult r = nv.compile(prog, 3, opts);
    if (r != NVRTC_SUCCESS) {
        size_t n = 0;
        nv.log_size(prog, &n);
        std::vector<char> log(n + 1, 0);
        nv.log(prog, log.data());
        nv.destroy(&prog);
        return fail(QCSS_ERR_INVALID, "NVRTC compilation failed: %.400s", log.data());
    }
    size_t n = 0;
    nv.cubin_size(prog, &n);
    cubin->assign(n, 0);
    nv.cubin(prog, cubin->data());
    nv.destroy(&prog);
    return QCSS_OK;
}

cudaError_t launch_one_rt(cudaKernel_t kernel, const small::NamedArgs& args, int64_t units, size_t smem, cudaStream_t stream) {
    const void* fn = reinterpret_cast<const void*>(kernel);
    int sms = 0;
    cudaError_t e = small::sm_count(&sms);
    if (e != cudaSuccess) return e;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, small::kThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const int64_t want = (units + small::kThreads - 1) / small::kThreads, wave = (int64_t)sms * per_sm;
    int64_t grid = want < wave ? want : wave;
    if (grid < 1) grid = 1;
    void* params[1] = {const_cast<small::NamedArgs*>(&args)};
    return cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(small::kThreads), params, smem, stream);
}

// small::launch_split with runtime kernel handles: whole units through FAST, the ragged rest through FULL
cudaError_t launch_split_rt(int vec, cudaKernel_t kfast, cudaKernel_t kfull, small::NamedArgs a, const SmallLaunch& l,
                            size_t smem_fast, size_t smem, cudaStream_t stream) {
    const DecodeIO io = l.io;
    int64_t fast_units = 0;
    if (small::fast_eligible(l)) {
        const int64_t whole_words = (io.tail_mask == 0xFFFFFFFFu) ? io.words : io.words - 1;
        fast_units = whole_words / vec;
    }
    if (fast_units > 0) {
        a.io = io;
        a.io.words = fast_units * vec;
        a.io.tail_mask = 0xFFFFFFFFu;
        const cudaError_t e = launch_one_rt(kfast, a, fast_units, smem_fast, stream);
        if (e != cudaSuccess) return e;
    }
    const int64_t done = fast_units * vec;
    if (done < io.words) {
        a.io = io;
        a.io.words = io.words - done;
        a.io.first_word = io.first_word + (uint64_t)done;
        if (a.io.ex) a.io.ex += done;
        if (a.io.ez) a.io.ez += done;
        if (a.io.synd_x) a.io.synd_x += done;
        if (a.io.synd_z) a.io.synd_z += done;
        if (a.io.corr_x) a.io.corr_x += done;
        if (a.io.corr_z) a.io.corr_z += done;
        if (a.io.flip_x) a.io.flip_x += done;
        if (a.io.flip_z) a.io.flip_z += done;
        if (a.io.miss_x) a.io.miss_x += done;
        if (a.io.miss_z) a.io.miss_z += done;
        if (a.io.ex_out) a.io.ex_out += done;
        if (a.io.ez_out) a.io.ez_out += done;
        return launch_one_rt(kfull, a, (a.io.words + vec - 1) / vec, smem, stream);
    }
    return cudaSuccess;
}

// small::launch_named for the in-process kernels
cudaError_t launch_spec_rtc(const SpecKernels& k, const SmallLaunch& l, cudaStream_t stream) {
    small::NamedArgs a;
    a.fm_x = l.x->lut_fm;
    a.co_x = l.x->lut_corr;
    a.e32_x = l.x->lut_e32;
    a.fm_z = l.z->lut_fm;
    a.co_z = l.z->lut_corr;
    a.e32_z = l.z->lut_e32;
    a.io = l.io;
    const size_t smem = small::lut_smem(*l.x, *l.z, !k.x_sliced, !k.z_sliced, false);
    const size_t smem_fast = small::lut_smem(*l.x, *l.z, !k.x_sliced, !k.z_sliced, true);
    if (l.sample && l.io.use_gap && l.gapq)
        return launch_split_rt(k.gapq_w, k.gapq, k.s_full, a, l, smem_fast + k.gapq_smem, smem, stream);
    if (l.sample) return launch_split_rt(1, k.s_fast, k.s_full, a, l, smem_fast, smem, stream);
    return launch_split_rt(4, k.fast, k.full, a, l, smem_fast, smem, stream);
}

}  // namespace

extern "C" {

QCSS_API int qcss_code_specialize(qcss_code* c, const char* nvrtc_path, const char* cache_dir) {
    if (!c) return fail(QCSS_ERR_INVALID, "code is NULL");
    if (c->k > 1) return fail(QCSS_ERR_UNSUPPORTED, "specialisation covers k = 1");
    if (!c->small || c->side_x.mode == kModeNone || c->side_z.mode == kModeNone)
        return fail(QCSS_ERR_UNSUPPORTED, "specialisation needs a decodable code (n <= %d, m <= %d, both tables)", kMaxN, kMaxM);
    std::string src = "#include \"small_common.cuh\"\nnamespace qcss {\nnamespace spec {\n";
    emit_side(src, "Spec_X", c->side_x, c->rows_x, c->lmask_x);
    emit_side(src, "Spec_Z", c->side_z, c->rows_z, c->lmask_z);
    src += "}\n}\nusing namespace qcss;\n";
    const char* decl = "extern \"C\" __global__ void __launch_bounds__(256, %d) %s(const __grid_constant__ small::NamedArgs a) { %s(a); }\n";
    char line[512];
    snprintf(line, sizeof(line), decl, 2, "spec_fast", "small::named_body<spec::Spec_X, spec::Spec_Z, 4, false, true>"); src += line;
    snprintf(line, sizeof(line), decl, 1, "spec_full", "small::named_body<spec::Spec_X, spec::Spec_Z, 4, false, false>"); src += line;
    snprintf(line, sizeof(line), decl, 3, "spec_sample_fast", "small::named_body<spec::Spec_X, spec::Spec_Z, 1, true, true>"); src += line;
    snprintf(line, sizeof(line), decl, 1, "spec_sample_full", "small::named_body<spec::Spec_X, spec::Spec_Z, 1, true, false>"); src += line;
    snprintf(line, sizeof(line), decl, 3, "spec_gapq", "small::named_gapq_body<spec::Spec_X, spec::Spec_Z>"); src += line;

    uint64_t h = fnv1a(src);
    for (int i = 0; i < kEmbeddedHeaderCount; ++i) h = fnv1a(kEmbeddedHeaderText[i], h);
    char tag[32];
    snprintf(tag, sizeof(tag), "%016llx", (unsigned long long)h);
    std::vector<char> cubin;
    std::string cache_file;
    if (cache_dir != nullptr && *cache_dir) {
        cache_file = std::string(cache_dir) + "/qcss_rtc_" + tag + ".cubin";
        if (FILE* f = fopen(cache_file.c_str(), "rb")) {
            fseek(f, 0, SEEK_END);
            const long n = ftell(f);
            fseek(f, 0, SEEK_SET);
            if (n > 0) {
                cubin.resize((size_t)n);
                if (fread(cubin.data(), 1, (size_t)n, f) != (size_t)n) cubin.clear();
            }
            fclose(f);
        }
    }
    if (cubin.empty()) {
        Nvrtc nv;
        int rc = nvrtc_open(nvrtc_path, &nv);
        if (rc) return rc;
        rc = nvrtc_compile(nv, src, &cubin);
        dlclose(nv.dl);
        if (rc) return rc;
        if (!cache_file.empty()) {
            const std::string tmp = cache_file + ".tmp" + std::to_string((long)getpid());
            if (FILE* f = fopen(tmp.c_str(), "wb")) {
                const bool ok = fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
                fclose(f);
                if (ok) rename(tmp.c_str(), cache_file.c_str());      // atomic: concurrent builders race benignly
                else remove(tmp.c_str());
            }
        }
    }
    SpecKernels* k = new (std::nothrow) SpecKernels();
    if (!k) return fail(QCSS_ERR_NOMEM, "out of host memory");
    cudaError_t e = cudaLibraryLoadData(&k->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&k->fast, k->lib, "spec_fast");
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&k->full, k->lib, "spec_full");
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&k->s_fast, k->lib, "spec_sample_fast");
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&k->s_full, k->lib, "spec_sample_full");
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&k->gapq, k->lib, "spec_gapq");
    if (e != cudaSuccess) {
        spec_rtc_free(k);
        return fail(QCSS_ERR_CUDA, "loading the specialised kernels failed: %s", cudaGetErrorString(e));
    }
    // launch shapes of small::launch_named / GapqShape for the static policies of this code
    const auto mb = [](const GenericSide& s) { return s.m <= kSlicedM ? s.m : (s.m <= 8 ? 8 : 16); };
    k->x_sliced = c->side_x.m <= kSlicedM;
    k->z_sliced = c->side_z.m <= kSlicedM;
    const int rows = mb(c->side_x) + mb(c->side_z) + 2, T = small::kThreads;
    k->gapq_w = (rows * 4 * T * 4 <= 36 * 1024) ? 4 : ((rows * 2 * T * 4 <= 48 * 1024) ? 2 : 1);
    k->gapq_smem = (size_t)rows * k->gapq_w * T * 4 + (size_t)T * k->gapq_w * c->n * 2;
    spec_rtc_free(c->rtc);
    c->rtc = k;
    snprintf(c->spec_tag, sizeof(c->spec_tag), "%.12s", tag);
    return QCSS_OK;
}

}  // extern "C"
