"""GPU parity tests proper: the CUDA path, called through the C ABI (libqcss.so via ctypes),
against the oracle on the same inputs.  Bit-exact everywhere (integer work)."""

import os

import numpy as np
import pytest

from oracle import css as ocss, montecarlo as omc, philox as ophilox
import css_code
from quantum_css_codes_b200 import CSSCode, SyndromeCode, codes, planes, _native

pytestmark = pytest.mark.gpu

NAMES = ["steane", "qrm15", "golay23"]
_cache = {}


def pair(name):
    if name not in _cache:
        h1, h2 = getattr(codes, name)()
        _cache[name] = (CSSCode(np.array(h1), np.array(h2)), ocss.build_css(np.array(h1), np.array(h2)))
    return _cache[name]


def test_library_loaded_and_device_present():
    import ctypes
    lib = _native.load()
    count = ctypes.c_int()
    _native.check(lib.qcss_device_count(ctypes.byref(count)))
    assert count.value >= 1
    assert lib.qcss_version() >= 100


@pytest.mark.parametrize("name", NAMES)
def test_named_codes_use_static_kernels(name):
    code, _ = pair(name)
    assert code.device.kernel_name() == f"small-static({name})"


@pytest.mark.parametrize("name", NAMES)
def test_weight_le_1_errors(name):
    """BASELINE config 1 on the GPU: syndrome + decode of every weight <= 1 X / Z error."""
    code, ref = pair(name)
    errs = np.vstack([np.zeros((1, code.n), dtype=np.uint8), np.eye(code.n, dtype=np.uint8)])
    for which in (1, 2):
        h, table, lop = ocss.pauli_side(ref, which)
        assert np.array_equal(code.syndromes(errs, which), omc.syndromes_batch(h, errs))
        out = code.decode(errs, which)
        assert np.array_equal(out["correction"], errs)
        assert not out["flip"].any() and not out["miss"].any()
        keys = omc.keys_batch(omc.syndromes_batch(h, errs))
        assert all(int(k) in table for k in keys)


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("shots,p", [(1, 0.3), (127, 0.1), (128, 0.5), (129, 0.02), (100003, 0.05)])
def test_random_batches_bit_exact(name, shots, p):
    code, ref = pair(name)
    rng = np.random.default_rng(shots)
    ex, ez = omc.sample_depolarizing(rng, shots, code.n, p)
    for which, errs in ((2, ex), (1, ez)):
        h, table, lop = ocss.pauli_side(ref, which)
        want = omc.decode_batch(h, table, lop, errs)
        assert np.array_equal(code.syndromes(errs, which), want["synd"])
        out = code.decode(errs, which)
        assert np.array_equal(out["correction"], want["corr"])
        assert np.array_equal(out["flip"], want["flip"])
        assert np.array_equal(out["miss"], want["miss"])
    assert code.decode_xz(ex, ez) == omc.tally_xz(ref, ex, ez)


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("which", [1, 2])
def test_golden_reference_decodes(golden, name, which):
    """Outputs of the unmodified reference (tests/golden) reproduced by the CUDA path."""
    code, _ = pair(name)
    pre = f"{name}_w{which}"
    errs = golden[pre + "_errs"].astype(np.uint8)
    assert np.array_equal(code.syndromes(errs, which), golden[pre + "_synd"])
    out = code.decode(errs, which)
    assert np.array_equal(out["correction"], golden[pre + "_corr"])
    assert np.array_equal(out["flip"], golden[pre + "_flip"])
    assert np.array_equal(out["miss"], golden[pre + "_miss"])


def test_steane_all_pauli_patterns():
    """All 4^7 Pauli patterns (joint X/Z): tallies equal the oracle's, incl. fail_any."""
    code, ref = pair("steane")
    pats = omc.all_patterns(14)
    ex, ez = pats[:, :7].copy(), pats[:, 7:].copy()
    got = code.decode_xz(ex, ez)
    assert got == omc.tally_xz(ref, ex, ez)
    assert got["fail_x"] == 64 * 128 and got["fail_z"] == 64 * 128


A4 = {("qrm15", 1): [0, 0, 105, 35, 1260, 168, 4725, 435, 6000, 280, 2835, 105, 420, 0, 15, 1],
      ("qrm15", 2): [0, 0, 0, 0, 965, 1211, 3625, 2055, 4380, 1380, 1792, 400, 455, 105, 15, 1],
      ("golay23", 1): [0, 0, 0, 0, 8855, 5313, 86779, 28589, 429088, 101200, 1005928, 171304, 1180774, 138138,
                       715990, 61226, 216568, 14168, 28336, 0, 1771, 253, 23, 1]}
A4["golay23", 2] = A4["golay23", 1]
A4_MISS_QRM_X = [0, 0, 0, 0, 840, 1848, 1960, 2520, 2520, 1960, 1848, 840, 0, 0, 0, 0]


@pytest.mark.parametrize("name,which", list(A4.keys()))
def test_exhaustive_failure_weight_enumerators(name, which):
    """Every one of the 2^n patterns of one Pauli type: flips (and misses) by weight = SURVEY A.4."""
    code, _ = pair(name)
    n = code.n
    flips = np.zeros(n + 1, dtype=np.int64)
    misses = np.zeros(n + 1, dtype=np.int64)
    chunk = 1 << 21
    for lo in range(0, 1 << n, chunk):
        pats = omc.all_patterns(n, lo, min(1 << n, lo + chunk))
        out = code.decode(pats, which)
        wt = pats.sum(axis=1)
        flips += np.bincount(wt[out["flip"] == 1], minlength=n + 1)
        misses += np.bincount(wt[out["miss"] == 1], minlength=n + 1)
    assert flips.tolist() == A4[(name, which)]
    assert misses.tolist() == (A4_MISS_QRM_X if (name, which) == ("qrm15", 2) else [0] * (n + 1))


@pytest.mark.parametrize("n,m1,m2", [(5, 2, 2), (12, 5, 4), (16, 8, 7), (20, 3, 12), (32, 16, 9), (9, 6, 1),
                                     (32, 5, 5), (17, 8, 8)])
def test_generic_kernels_partial_tables(n, m1, m2):
    """Codes that match no compiled-in descriptor run the generic (runtime-H) kernels."""
    rng = np.random.default_rng(n * 100 + m1)
    sides = {}
    for which, m in ((1, m1), (2, m2)):
        h = rng.integers(0, 2, size=(m, n))
        lrow = rng.integers(0, 2, size=n)
        keys = rng.permutation(1 << m)[: max(1, (1 << m) * 2 // 3)]
        sides[which] = (h, {int(k): rng.integers(0, 2, size=n) for k in keys}, lrow[None, :])
    dev = _native.DeviceCode(n, sides[1][0], sides[2][0], sides[1][2][0], sides[2][2][0], sides[1][1], sides[2][1])
    assert dev.kernel_name().startswith("small-generic")
    shots = 5000
    ex = rng.integers(0, 2, size=(shots, n), dtype=np.uint8)
    ez = rng.integers(0, 2, size=(shots, n), dtype=np.uint8)
    tall = {}
    for which, errs in ((2, ex), (1, ez)):
        want = omc.decode_batch(*sides[which], errs)
        e_planes = planes.pack_planes(errs)
        s = planes.unpack_planes(dev.syndrome_planes(e_planes, shots, which), shots)
        assert np.array_equal(s, want["synd"])
        corr, flip, miss, tally = dev.decode_planes(e_planes, shots, which)
        assert np.array_equal(planes.unpack_planes(corr, shots), want["corr"])
        assert np.array_equal(planes.unpack_plane(flip, shots), want["flip"])
        assert np.array_equal(planes.unpack_plane(miss, shots), want["miss"])
        tall[which] = want
    got = dev.decode_xz_planes(planes.pack_planes(ex), planes.pack_planes(ez), shots)
    assert got["fail_any"] == int((tall[2]["flip"] | tall[1]["flip"]).sum())
    assert got["miss_x"] == int(tall[2]["miss"].sum()) and got["miss_z"] == int(tall[1]["miss"].sum())


@pytest.mark.parametrize("compiler", ["nvrtc", "nvcc"])
@pytest.mark.parametrize("n,m1,m2", [(5, 2, 2), (12, 5, 4), (20, 3, 12), (32, 16, 9), (17, 8, 8)])
def test_specialized_kernels_match_generic_and_oracle(n, m1, m2, compiler):
    """Kernels compiled for ONE code -- in process by NVRTC (qcss_code_specialize: embedded headers -> cubin ->
    cudaLibraryLoadData, no subprocess) or by an nvcc subprocess (qcss_code_spec_source -> qcss_code_load_specialized)
    -- must give the oracle's bits on every output (shared inputs) and the generic kernels' tallies on the fused
    Philox run (which has no oracle for random tables beyond the sampler itself)."""
    from quantum_css_codes_b200 import specialize
    rng = np.random.default_rng(n * 100 + m1)
    sides = {}
    for which, m in ((1, m1), (2, m2)):
        h = rng.integers(0, 2, size=(m, n))
        lrow = rng.integers(0, 2, size=n)
        keys = rng.permutation(1 << m)[: max(1, (1 << m) * 2 // 3)]
        sides[which] = (h, {int(k): rng.integers(0, 2, size=n) for k in keys}, lrow[None, :])
    make = lambda: _native.DeviceCode(n, sides[1][0], sides[2][0], sides[1][2][0], sides[2][2][0], sides[1][1], sides[2][1])
    generic, dev = make(), make()
    assert specialize.specialize(dev, compiler).startswith("small-static(nvrtc:" if compiler == "nvrtc" else "small-static(jit:")
    assert generic.kernel_name().startswith("small-generic")
    for shots in (1, 127, 5000, 128 * 300 + 5):
        ex = rng.integers(0, 2, size=(shots, n), dtype=np.uint8)
        ez = (rng.random((shots, n)) < 0.05).astype(np.uint8)
        tall = {}
        for which, errs in ((2, ex), (1, ez)):
            want = omc.decode_batch(*sides[which], errs)
            e_planes = planes.pack_planes(errs)
            corr, flip, miss, tally = dev.decode_planes(e_planes, shots, which)
            assert np.array_equal(planes.unpack_planes(corr, shots), want["corr"])
            assert np.array_equal(planes.unpack_plane(flip, shots), want["flip"])
            assert np.array_equal(planes.unpack_plane(miss, shots), want["miss"])
            tall[which] = want
        got = dev.decode_xz_planes(planes.pack_planes(ex), planes.pack_planes(ez), shots)
        assert got == generic.decode_xz_planes(planes.pack_planes(ex), planes.pack_planes(ez), shots)
        assert got["fail_x"] == int(tall[2]["flip"].sum()) and got["fail_z"] == int(tall[1]["flip"].sum())
        assert got["fail_any"] == int((tall[2]["flip"] | tall[1]["flip"]).sum())
    for p_err in (1e-3, 0.2):
        assert dev.mc_run(p_err, (1 << 20) + 77, seed=5, first_shot=256) == generic.mc_run(p_err, (1 << 20) + 77, seed=5, first_shot=256)
        sx, sz = dev.mc_sample(p_err, 3000, seed=9)
        gx, gz = generic.mc_sample(p_err, 3000, seed=9)
        assert np.array_equal(sx, gx) and np.array_equal(sz, gz)


def test_specialize_through_csscode_and_cache():
    """CSSCode.specialize() on codes the library has no descriptor for; the second build of the same
    code reuses the cached shared object."""
    from quantum_css_codes_b200 import specialize
    hx, hz = codes.shor9()
    code, again = CSSCode(hx, hz), CSSCode(hx, hz)
    ref = ocss.build_css(hx, hz)
    assert code.device.kernel_name().startswith("small-generic")
    generic_low_p = code.monte_carlo(2e-3, 5_000_011, seed=6)            # generic kernels: in-place gap sampler
    name = code.specialize("nvcc")
    assert name.startswith("small-static(jit:") and again.specialize("nvcc") == name
    so = [f for f in os.listdir(specialize.JIT_DIR) if name[len("small-static(jit:"):-1] in f]
    assert len(so) == 1
    third = CSSCode(hx, hz)
    rtc = third.specialize()                                             # default: NVRTC, in process
    assert rtc.startswith("small-static(nvrtc:") and CSSCode(hx, hz).specialize() == rtc
    assert any(f.startswith("qcss_rtc_") and f.endswith(".cubin") for f in os.listdir(specialize.JIT_DIR))
    assert third.monte_carlo(2e-3, 5_000_011, seed=6) == generic_low_p
    rng = np.random.default_rng(9)
    ex, ez = omc.sample_depolarizing(rng, 20000, code.n, 0.1)
    assert code.decode_xz(ex, ez) == omc.tally_xz(ref, ex, ez)
    got = code.monte_carlo(0.05, 1 << 16, seed=3)
    sx, sz = ophilox.sample_bits(3, 0, 1 << 16, code.n, 0.05)
    assert got == omc.tally_xz(ref, sx, sz)
    # the specialised kernels sample low error rates with the CTA-wide two-phase gap sampler: same streams
    assert code.monte_carlo(2e-3, 5_000_011, seed=6) == generic_low_p
    got = code.monte_carlo(2e-3, 40_000, seed=7)
    sx, sz = ophilox.sample_bits(7, 0, 40_000, code.n, 2e-3)
    assert got == omc.tally_xz(ref, sx, sz)


# ---- fused Philox sampler -----------------------------------------------------------------------

@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("p", [1e-3, 0.012, 0.0156, 0.0157, 0.05, 0.5, 0.0, 1.0])
def test_fused_sampler_bit_exact_vs_oracle(name, p):
    code, ref = pair(name)
    shots, seed, first = 6000, 0xC0FFEE1234, 128 * 11
    ex, ez = code.sample_errors(p, shots, seed, first)
    ox, oz = ophilox.sample_bits(seed, first, shots, code.n, p)
    assert np.array_equal(ex, ox) and np.array_equal(ez, oz)
    assert code.monte_carlo(p, shots, seed, first) == omc.tally_xz(ref, ox, oz)


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("p,shots", [(1e-3, 300_000_077), (5e-3, 50_000_000), (0.0078, 20_000_033), (0.0156, 20_000_033),
                                     (1e-5, 100_000_000)])
def test_gap_sampler_queue_kernel_equals_in_place_kernel(name, p, shots):
    """The CTA-wide two-phase gap sampler (k_small_named_gapq, default below p = 1/64) and the in-place kernel
    (option "gapq" = 0) draw from the same Philox streams: identical tallies, at sizes that give every CTA many
    iterations, queue loads from almost empty (p = 1e-5) to ~40 % of the site-words (p just below 1/64), and a
    ragged tail; a non-aligned first_shot shard as well."""
    code, _ = pair(name)
    with _native.option("gapq", 0):
        want = code.monte_carlo(p, shots, seed=0xA11CE)
        want_shard = code.monte_carlo(p, 1_000_001, seed=0xA11CE, first_shot=128 * 12345)
    assert code.monte_carlo(p, shots, seed=0xA11CE) == want
    assert code.monte_carlo(p, 1_000_001, seed=0xA11CE, first_shot=128 * 12345) == want_shard
    assert want["fail_any"] > 0 or p < 1e-4


@pytest.mark.parametrize("make", ["shor9", "surface3", "surface5"])
def test_gap_sampler_queue_kernel_generic_codes(make):
    """The same A/B for codes without a static descriptor (generic kernels: runtime H in the parameter block,
    buckets n <= 16 / 32 and m <= 5 / 8 / 16), and against the oracle sampler on a small run."""
    hx, hz = {"shor9": codes.shor9, "surface3": lambda: codes.rotated_surface(3),
              "surface5": lambda: codes.rotated_surface(5)}[make]()
    code = CSSCode(np.array(hx), np.array(hz))
    ref = ocss.build_css(np.array(hx), np.array(hz))
    assert code.device.kernel_name().startswith("small-generic")
    for p, shots in ((1e-3, 40_000_077), (6e-3, 8_000_000)):
        with _native.option("gapq", 0):
            want = code.monte_carlo(p, shots, seed=0xBEEF)
        assert code.monte_carlo(p, shots, seed=0xBEEF) == want
    got = code.monte_carlo(3e-3, 30_000, seed=12, first_shot=256)
    sx, sz = ophilox.sample_bits(12, 256, 30_000, code.n, 3e-3)
    assert got == omc.tally_xz(ref, sx, sz)


def test_monte_carlo_independent_of_sharding():
    code, _ = pair("steane")
    whole = code.monte_carlo(0.05, 1 << 20, seed=11)
    parts = [code.monte_carlo(0.05, 1 << 18, seed=11, first_shot=i << 18) for i in range(4)]
    for key in whole:
        assert whole[key] == sum(p[key] for p in parts)


@pytest.mark.parametrize("name,p,shots", [("steane", 1e-2, 1 << 30), ("steane", 5e-2, 1 << 28),
                                          ("qrm15", 5e-2, 1 << 28), ("golay23", 5e-2, 1 << 28)])
def test_monte_carlo_rates_within_binomial_ci(name, p, shots):
    """GPU-sampled logical error rates vs the exact enumerator rates, 5 sigma."""
    code, ref = pair(name)
    got = code.monte_carlo(p, shots, seed=2026)
    q = 2 * p / 3
    for key, which in (("fail_x", 2), ("fail_z", 1)):
        h, table, lop = ocss.pauli_side(ref, which)
        flips, misses = (A4[(name, which)], None) if (name, which) in A4 else omc.failure_enumerator(h, table, lop)
        rate = omc.exact_rate(flips, q)
        sigma = np.sqrt(rate * (1 - rate) / shots)
        assert abs(got[key] / shots - rate) < 5 * sigma, (key, got[key] / shots, rate)
    assert got["shots"] == shots


def test_large_host_buffer_decode_crosses_chunks():
    """qcss_decode_xz streams host planes in chunks; a batch spanning several chunks must tally
    exactly like the fused run that sampled the same planes on the device."""
    code, ref = pair("steane")
    shots = 90_000_001
    ex, ez = code.device.mc_sample(0.02, shots, seed=5)
    got = code.device.decode_xz_planes(ex, ez, shots)
    assert got == code.monte_carlo(0.02, shots, seed=5)
    sub = 200_000
    ox, oz = planes.unpack_planes(ex[:, : sub // 64], sub), planes.unpack_planes(ez[:, : sub // 64], sub)
    sx, sz = ophilox.sample_bits(5, 0, sub, code.n, 0.02)
    assert np.array_equal(ox, sx) and np.array_equal(oz, sz)


# ---- hypergraph-product code: syndrome only ---------------------------------------------------

def test_hgp1600_syndromes_golden_and_random(golden):
    hx, hz = codes.hgp1600()
    code = SyndromeCode(hx, hz)
    assert code.device.kernel_name().startswith("tiled-sparse")
    errs = np.unpackbits(golden["hgp_errs"], axis=1, bitorder="little")[:, :1600]
    for which, h, key in ((2, hz, "hgp_synd_hz"), (1, hx, "hgp_synd_hx")):
        want = np.unpackbits(golden[key], axis=1, bitorder="little")[:, :768]
        assert np.array_equal(code.syndromes(errs, which), want)
    rng = np.random.default_rng(16)
    for shots, p in ((1, 0.5), (513, 0.5), (20000, 1e-3)):
        errs = (rng.random((shots, 1600)) < p).astype(np.uint8)
        for which, h in ((2, hz), (1, hx)):
            assert np.array_equal(code.syndromes(errs, which), omc.syndromes_batch(h, errs))


def test_hgp1600_tile_major_layout(golden):
    """The tile-major entry point (qcss_syndrome_tiles: one bulk copy per part-tile) gives the same syndromes:
    golden reference outputs, ragged and multi-tile batches, zero padding in the last tile."""
    hx, hz = codes.hgp1600()
    code = SyndromeCode(hx, hz)
    errs = np.unpackbits(golden["hgp_errs"], axis=1, bitorder="little")[:, :1600]
    for which, key in ((2, "hgp_synd_hz"), (1, "hgp_synd_hx")):
        want = np.unpackbits(golden[key], axis=1, bitorder="little")[:, :768]
        assert np.array_equal(code.syndromes_tiled(errs, which), want)
    rng = np.random.default_rng(17)
    for shots, p in ((1, 0.5), (1023, 0.5), (1024, 0.2), (1025, 0.5), (20 * 1024 + 77, 1e-3)):
        errs = (rng.random((shots, 1600)) < p).astype(np.uint8)
        for which, h in ((2, hz), (1, hx)):
            tiles = code.device.syndrome_tiles(planes.pack_tiles(errs), shots, which)
            assert np.array_equal(planes.unpack_tiles(tiles, shots), omc.syndromes_batch(h, errs))
            pad = np.unpackbits(tiles[-1].view(np.uint8), axis=1, bitorder="little")[:, shots - (len(tiles) - 1) * 1024:]
            assert not pad.any()
    small, _ = pair("steane")
    with pytest.raises(_native.NativeLibraryError, match="tile-major layout serves the sparse"):
        small.device.syndrome_tiles(planes.pack_tiles(np.zeros((5, 7), dtype=np.uint8)), 5, 1)


@pytest.mark.parametrize("p", [1e-3, 0.05])
def test_hgp1600_fused_sampler(p):
    """K3 for large codes (qcss_sample_syndrome_tiles): the errors drawn inside the syndrome kernel are the
    oracle sampler's (same Philox streams, n = 1600 sites), the syndromes are H.e of exactly those errors, and a
    sharded run reproduces the whole run."""
    hx, hz = codes.hgp1600()
    code = SyndromeCode(hx, hz)
    shots, seed, first = 3000, 0xC0FFEE, 2048
    s_x, s_z, e_x, e_z = code.sample_syndromes(p, shots, seed=seed, first_shot=first, return_errors=True)
    want_x, want_z = ophilox.sample_bits(seed, first, shots, 1600, p)
    assert np.array_equal(e_x, want_x) and np.array_equal(e_z, want_z)
    assert np.array_equal(s_x, omc.syndromes_batch(hz, want_x))          # which = 2: parity_check_c2 on X errors
    assert np.array_equal(s_z, omc.syndromes_batch(hx, want_z))          # which = 1: parity_check_c1 on Z errors
    only_x, only_z = code.sample_syndromes(p, shots, seed=seed, first_shot=first)
    assert np.array_equal(only_x, s_x) and np.array_equal(only_z, s_z)
    part_x, part_z = code.sample_syndromes(p, 1024 + 7, seed=seed, first_shot=first + 1024)
    assert np.array_equal(part_x, s_x[1024:2048 + 7]) and np.array_equal(part_z, s_z[1024:2048 + 7])
    with pytest.raises(ValueError, match="multiple of 1024"):
        code.sample_syndromes(p, 10, first_shot=512)
    small, _ = pair("steane")
    with pytest.raises(_native.NativeLibraryError, match="sample through qcss_mc_run"):
        small.device.sample_syndrome_tiles(p, 10)


def test_tile_major_empty_batches():
    hx, hz = codes.hgp1600()
    code = SyndromeCode(hx, hz)
    assert code.syndromes_tiled(np.zeros((0, 1600), dtype=np.uint8), 1).shape == (0, 768)
    s_x, s_z = code.sample_syndromes(1e-3, 0)
    assert s_x.shape == (0, 768) and s_z.shape == (0, 768)
    assert planes.pack_tiles(np.zeros((0, 5), dtype=np.uint8)).shape == (0, 5, 16)


def test_fused_sampler_midsize_code():
    rng = np.random.default_rng(301)
    mats = []
    for m in (130, 77):
        h = np.zeros((m, 300), dtype=np.int64)
        for i in range(m):
            h[i, rng.choice(300, size=9, replace=False)] = 1
        mats.append(h)
    code = SyndromeCode(mats[0], mats[1])
    for p, shots in ((0.004, 5000), (0.3, 1500)):
        s_x, s_z, e_x, e_z = code.sample_syndromes(p, shots, seed=5, return_errors=True)
        want_x, want_z = ophilox.sample_bits(5, 0, shots, 300, p)
        assert np.array_equal(e_x, want_x) and np.array_equal(e_z, want_z)
        assert np.array_equal(s_x, omc.syndromes_batch(mats[1], want_x))
        assert np.array_equal(s_z, omc.syndromes_batch(mats[0], want_z))


@pytest.mark.parametrize("n,m1,m2", [(41, 9, 23), (70, 100, 3), (127, 64, 64), (1000, 17, 700)])
@pytest.mark.parametrize("p", [2e-3, 0.0156, 0.0157, 0.2])
def test_fused_sampler_ragged_shapes(n, m1, m2, p):
    """Both fused kernels (scatter below p = 1/64, gather above) on shapes that stress their index arithmetic: n not
    a multiple of 8 (partial first-look groups), more checks than qubits, rows of weight 0 .. 9 (padded 4-entry
    groups), columns no check touches (empty transposed supports), a ragged last tile."""
    rng = np.random.default_rng(n * 1000 + m1)
    mats = []
    for m in (m1, m2):
        h = np.zeros((m, n), dtype=np.int64)
        for i in range(m):
            h[i, rng.choice(n - 3, size=int(rng.integers(0, 10)), replace=False)] = 1     # the last 3 columns stay empty
        mats.append(h)
    code = SyndromeCode(mats[0], mats[1])
    shots = 2 * 1024 + 333
    s_x, s_z, e_x, e_z = code.sample_syndromes(p, shots, seed=77, first_shot=1024, return_errors=True)
    want_x, want_z = ophilox.sample_bits(77, 1024, shots, n, p)
    assert np.array_equal(e_x, want_x) and np.array_equal(e_z, want_z)
    assert np.array_equal(s_x, omc.syndromes_batch(mats[1], want_x))
    assert np.array_equal(s_z, omc.syndromes_batch(mats[0], want_z))
    only_x, only_z = code.sample_syndromes(p, shots, seed=77, first_shot=1024)
    assert np.array_equal(only_x, s_x) and np.array_equal(only_z, s_z)


@pytest.mark.parametrize("n,m,row_w,shots", [(40, 20, 5, 3000), (300, 130, 9, 1025), (700, 1024, 3, 5000),
                                             (3000, 900, 6, 2500)])
def test_tile_major_random_sparse(n, m, row_w, shots):
    """Shapes that take 1 part (whole tile per stage), several parts, and the full 1024 rows."""
    rng = np.random.default_rng(n * 3 + m)
    mats = []
    for _ in range(2):
        h = np.zeros((m, n), dtype=np.int64)
        for i in range(m):
            h[i, rng.choice(n, size=row_w, replace=False)] = 1
        mats.append(h)
    code = SyndromeCode(mats[0], mats[1])
    errs = (rng.random((shots, n)) < 0.3).astype(np.uint8)
    for which in (1, 2):
        assert np.array_equal(code.syndromes_tiled(errs, which), omc.syndromes_batch(mats[which - 1], errs))


@pytest.mark.parametrize("n,m,row_w,shots", [(40, 20, 5, 3000), (300, 130, 9, 1025), (3000, 900, 6, 700),
                                             (6000, 64, 30, 130)])
def test_tiled_syndromes_random_sparse(n, m, row_w, shots):
    """Any-size sparse syndrome kernels (TMA two-stage for n <= 7040, cp.async tile beyond)."""
    rng = np.random.default_rng(n + m)
    mats = []
    for _ in range(2):
        h = np.zeros((m, n), dtype=np.int64)
        for i in range(m):
            h[i, rng.choice(n, size=row_w, replace=False)] = 1
        mats.append(h)
    code = SyndromeCode(mats[0], mats[1])
    errs = (rng.random((shots, n)) < 0.3).astype(np.uint8)
    for which in (1, 2):
        assert np.array_equal(code.syndromes(errs, which), omc.syndromes_batch(mats[which - 1], errs))


@pytest.mark.parametrize("m,n,shots", [(300, 700, 1000), (1024, 2048, 4096), (130, 4100, 257), (256, 512, 33)])
def test_dense_syndromes_tensor_core_and_lop3_paths(m, n, shots):
    """Dense random H: the tcgen05 int8 MMA kernel (K1') and the bit-sliced kernel give the reference's
    np.mod(np.matmul(H, e), 2) bit for bit."""
    rng = np.random.default_rng(m * 7 + n)
    h1 = rng.integers(0, 2, size=(m, n))
    h2 = rng.integers(0, 2, size=(m // 2 + 1, n))
    errs = rng.integers(0, 2, size=(shots, n), dtype=np.uint8)
    want = {1: omc.syndromes_batch(h1, errs), 2: omc.syndromes_batch(h2, errs)}
    for force, prefix in ((1, "dense-tcgen05"), (0, "tiled-sparse")):
        with _native.option("dense", force):
            code = SyndromeCode(h1, h2)
            assert code.device.kernel_name().startswith(prefix)
        for which in (1, 2):
            assert np.array_equal(code.syndromes(errs, which), want[which]), (force, which)
    auto = SyndromeCode(h1, h2)
    assert auto.device.kernel_name().startswith("dense-tcgen05" if m * n >= 256 * 512 else "tiled-sparse")


def test_error_paths():
    code, _ = pair("steane")
    with pytest.raises(ValueError):
        code.syndromes(np.zeros((4, 7), dtype=np.uint8), 3)
    with pytest.raises(ValueError):
        code.monte_carlo(1.5, 100)
    with pytest.raises(ValueError):
        code.monte_carlo(0.1, 128, first_shot=64)
    hx, hz = codes.hgp1600()
    with pytest.raises(_native.NativeLibraryError, match="lookup decode covers"):
        SyndromeCode(hx, hz).device.decode_planes(planes.pack_planes(np.zeros((4, 1600), dtype=np.uint8)), 4, 1)


# ---- GPU-assisted syndrome table -------------------------------------------------

def _same_table(a, b):
    (ta, da), (tb, db) = a, b
    assert ta == tb and len(da) == len(db)
    for (ka, va), (kb, vb) in zip(da.items(), db.items()):           # insertion order matters
        assert type(ka) is type(kb) is np.int64 and ka == kb
        assert va.dtype == vb.dtype and np.array_equal(va, vb)


@pytest.mark.parametrize("name", ["steane", "qrm15", "golay23"])
def test_syndrome_table_gpu_named_codes(name, golden):
    """Device weight-layer search == host search == the reference's captured tables
    (css_code.py:715-735), including dict order and key type."""
    host = css_code.CSSCode(*[np.array(h) for h in getattr(codes, name)()])
    dev = css_code.CSSCode(*[np.array(h) for h in getattr(codes, name)()], table_builder="gpu")
    assert dev.t == host.t == int(golden[f"{name}_nkt"][2])
    for which, tag in (("_c1_syndromes", "c1"), ("_c2_syndromes", "c2")):
        _same_table((host.t, getattr(host, which)), (dev.t, getattr(dev, which)))
        table = getattr(dev, which)
        assert np.array_equal(np.array(list(table.keys())), golden[f"{name}_{tag}_keys"])
        assert np.array_equal(np.array(list(table.values())), golden[f"{name}_{tag}_vals"])
    for h in (host.parity_check_c1, host.parity_check_c2):
        _same_table(css_code.syndrome_table(h), css_code.syndrome_table_gpu(h))


def test_syndrome_table_gpu_larger_codes():
    rng = np.random.default_rng(2026)
    cases = [rng.integers(0, 2, size=(18, 40)), rng.integers(0, 2, size=(30, 52)),
             np.concatenate([np.eye(24, dtype=np.int64), rng.integers(0, 2, size=(24, 12))], axis=1),
             np.ones((1, 5), dtype=np.int64), np.eye(6, dtype=np.int64)]
    for h in cases:
        _same_table(css_code.syndrome_table(h), css_code.syndrome_table_gpu(h))
    assert css_code.syndrome_table_gpu(np.eye(6, dtype=np.int64))[0] == 6        # no collision at all: t = n
    with pytest.raises(MemoryError):
        css_code.syndrome_table_gpu(np.eye(20, dtype=np.int64), max_entries=1000)


# ---- per-syndrome histograms ------------------------------------------------

@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("shots", [1, 95, 4096, 100003])
def test_syndrome_histograms(name, shots):
    """hist[key] = number of shots with that big-endian syndrome key, both Pauli types; sums to shots."""
    code, ref = pair(name)
    rng = np.random.default_rng(shots)
    errs = (rng.random((shots, code.n)) < 0.2).astype(np.uint8)
    for which in (1, 2):
        h, table, lop = ocss.pauli_side(ref, which)
        keys = omc.keys_batch(omc.syndromes_batch(h, errs))
        want = np.bincount(keys, minlength=1 << h.shape[0]).astype(np.uint64)
        got = code.syndrome_histogram(errs, which)
        assert got.dtype == np.uint64 and np.array_equal(got, want)
        assert int(got.sum()) == shots


def test_syndrome_histogram_device_accumulates():
    import torch
    code, _ = pair("golay23")
    dev = code.device
    shots = 1 << 22
    stride = ((shots + 127) // 128) * 2
    ex = torch.empty((code.n, stride), dtype=torch.int64, device="cuda")
    ez = torch.empty((code.n, stride), dtype=torch.int64, device="cuda")
    dev.mc_sample_dev(0.02, shots, 4, 0, ex.data_ptr(), ez.data_ptr(), stride, 0)
    hist = torch.zeros(1 << 11, dtype=torch.int64, device="cuda")
    dev.syndrome_hist_dev(2, ex.data_ptr(), stride, shots, hist.data_ptr(), 0)
    dev.syndrome_hist_dev(2, ex.data_ptr(), stride, shots, hist.data_ptr(), 0)
    torch.cuda.synchronize()
    assert int(hist.sum()) == 2 * shots
    once = code.syndrome_histogram(planes.unpack_planes(ex.cpu().numpy().view(np.uint64), 4096), 2)
    sub = torch.zeros(1 << 11, dtype=torch.int64, device="cuda")
    dev.syndrome_hist_dev(2, ex.data_ptr(), stride, 4096, sub.data_ptr(), 0)
    torch.cuda.synchronize()
    assert np.array_equal(sub.cpu().numpy().astype(np.uint64), once)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_attach_to_reference_shaped_object_matches_oracle(name):
    """VERDICT r1 missing #8: the device path bound to an EXISTING code object (duck-typed like the reference's
    CSSCode) gives the oracle's tallies on shared inputs and the same Monte-Carlo stream as our own CSSCode."""
    from test_host_api import _ReferenceShapedCode
    from quantum_css_codes_b200 import attach
    obj = _ReferenceShapedCode(name)
    bound = attach(obj)
    code, ref = pair(name)
    rng = np.random.default_rng(5)
    ex, ez = omc.sample_depolarizing(rng, 20_000, code.n, 0.03)
    assert bound.decode_xz(ex, ez) == omc.tally_xz(ref, ex, ez)
    assert np.array_equal(bound.syndromes(ex, 2), omc.syndromes_batch(ref.parity_check_c2, ex))
    assert bound.monte_carlo(2e-3, 3_000_000, seed=3) == code.monte_carlo(2e-3, 3_000_000, seed=3)
    assert bound.device.kernel_name() == code.device.kernel_name()
