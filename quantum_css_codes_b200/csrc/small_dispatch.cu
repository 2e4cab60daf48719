// Host-side dispatch for the small-code kernels: descriptor matching and family selection.
#include "named_codes.inc"
#include "small_common.cuh"

namespace qcss {

namespace {

bool side_matches(const GenericSide& s, const uint32_t* rowmask, uint32_t lmask, const named::SideInfo& d) {
    if (s.n != d.n || s.m != d.m || s.mode == kModeNone) return false;
    if ((s.has_miss != 0) != d.has_miss || lmask != d.l) return false;
    for (int t = 0; t < d.m; ++t)
        if (rowmask[t] != d.rows[t]) return false;
    if (d.sliced) {
        if (s.tt_flip != d.tt_flip || s.tt_miss != d.tt_miss) return false;
        for (int j = 0; j < d.n; ++j)
            if (s.tt_corr[j] != d.tt_corr[j]) return false;
    }
    return true;
}

}  // namespace

int small_bucket_m(int mx, int mz) {
    int m = mx > mz ? mx : mz;
    if (m <= kSlicedM) return kSlicedM;
    if (m <= 8) return 8;
    return 16;
}

int match_named(const GenericSide& x, const uint32_t* rows_x, uint32_t lx, const GenericSide& z,
                const uint32_t* rows_z, uint32_t lz) {
    for (int i = 0; i < named::kNumNamed; ++i)
        if (side_matches(x, rows_x, lx, named::kNamed[i].x) && side_matches(z, rows_z, lz, named::kNamed[i].z))
            return i;
    return -1;
}

const char* named_name(int id) {
    return (id >= 0 && id < named::kNumNamed) ? named::kNamed[id].name : "generic";
}

cudaError_t launch_small(const SmallLaunch& l, cudaStream_t stream) {
    switch (l.named_id) {               // order of tools/gen_named_codes.py
        case 0: return launch_small_steane(l, stream);
        case 1: return launch_small_qrm15(l, stream);
        case 2: return launch_small_golay23(l, stream);
        default: break;
    }
    const int mb = small_bucket_m(l.x->m, l.z->m);
    if (l.x->n <= 16) return launch_small_generic16(l, mb, stream);
    return launch_small_generic32(l, mb, stream);
}

}  // namespace qcss
