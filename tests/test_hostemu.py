"""CPU checks of the CUDA kernels' per-thread code through the host emulation (tests/hostemu):
the same __host__ __device__ functions the sm_100a kernels call, run sequentially on the host and
compared bit-for-bit with the oracle.  The GPU parity tests proper are in test_gpu_*.py."""

import ctypes

import numpy as np
import pytest

import emu
from oracle import css as ocss, montecarlo as omc, philox as ophilox
from quantum_css_codes_b200 import codes, planes

NAMED = {"steane": 0, "qrm15": 1, "golay23": 2}


def build(name):
    code = ocss.build_css(*[np.array(h) for h in getattr(codes, name)()])
    sx = emu.Side(code.parity_check_c2, code.lz[0], code.c2_syndromes)
    sz = emu.Side(code.parity_check_c1, code.lx[0], code.c1_syndromes)
    return code, sx, sz


def test_transpose32():
    rng = np.random.default_rng(1)
    w = rng.integers(0, 1 << 32, size=32, dtype=np.uint64).astype(np.uint32)
    bits = ((w[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(np.uint8)   # [i][k]
    out = w.copy()
    emu.lib().emu_transpose32(out.ctypes.data_as(ctypes.c_void_p))
    obits = ((out[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(np.uint8)
    assert np.array_equal(obits, bits.T)


def test_philox_kat():
    """Random123 known-answer vectors for Philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        c = np.array(ctr, dtype=np.uint32); k = np.array(key, dtype=np.uint32); o = np.zeros(4, dtype=np.uint32)
        emu.lib().emu_philox(c.ctypes.data_as(ctypes.c_void_p), k.ctypes.data_as(ctypes.c_void_p),
                             o.ctypes.data_as(ctypes.c_void_p))
        assert tuple(int(v) for v in o) == want
        assert tuple(int(v) for v in ophilox.philox4x32_10(c, k)) == want


def test_named_descriptors_match_reference_matrices(golden):
    """The compiled-in descriptors are the reference's normalised matrices (golden fixtures)."""
    for name, idx in NAMED.items():
        for which_x, hkey, lkey in ((1, "h2", "lz"), (0, "h1", "lx")):
            n, m, l = ctypes.c_int(), ctypes.c_int(), ctypes.c_uint32()
            rows = (ctypes.c_uint32 * 16)()
            assert emu.lib().emu_named_side(idx, which_x, ctypes.byref(n), ctypes.byref(m), rows, ctypes.byref(l)) == 0
            h = golden[f"{name}_{hkey}"]
            assert (n.value, m.value) == (h.shape[1], h.shape[0])
            for t in range(m.value):
                assert rows[t] == sum(int(b) << j for j, b in enumerate(h[m.value - 1 - t]))
            assert l.value == sum(int(b) << j for j, b in enumerate(golden[f"{name}_{lkey}"][0]))


def check_against_oracle(code, sx, sz, ex, ez, named_id):
    shots = ex.shape[0]
    got = emu.decode(sx, sz, planes.pack_planes(ex), planes.pack_planes(ez), shots, named_id,
                     want=("synd", "corr", "flip", "miss"))
    for tag, which, errs in (("x", 2, ex), ("z", 1, ez)):
        h, table, lop = ocss.pauli_side(code, which)
        want = omc.decode_batch(h, table, lop, errs)
        assert np.array_equal(planes.unpack_planes(got["synd_" + tag], shots), want["synd"]), tag
        assert np.array_equal(planes.unpack_planes(got["corr_" + tag], shots), want["corr"]), tag
        assert np.array_equal(planes.unpack_plane(got["flip_" + tag], shots), want["flip"]), tag
        assert np.array_equal(planes.unpack_plane(got["miss_" + tag], shots), want["miss"]), tag
    assert got["tally"] == omc.tally_xz(code, ex, ez)
    # padding bits of every output plane stay zero
    for key, val in got.items():
        if key != "tally":
            bits = np.unpackbits(np.atleast_2d(val).view(np.uint8), axis=1, bitorder="little")
            assert not bits[:, shots:].any(), key


@pytest.mark.parametrize("name", list(NAMED))
@pytest.mark.parametrize("static", [True, False])
@pytest.mark.parametrize("shots", [1, 31, 128, 1000])
def test_decode_random_batches(name, static, shots):
    code, sx, sz = build(name)
    rng = np.random.default_rng(shots * 7 + len(name))
    ex = (rng.random((shots, code.n)) < 0.15).astype(np.uint8)
    ez = (rng.random((shots, code.n)) < 0.15).astype(np.uint8)
    ex[0] = 0
    ez[-1] = 1
    check_against_oracle(code, sx, sz, ex, ez, NAMED[name] if static else -1)


@pytest.mark.parametrize("name", list(NAMED))
def test_weight_le_1_errors_decode_to_zero_residual(name):
    """BASELINE config 1: every weight <= 1 X/Z error is corrected with no logical flip."""
    code, sx, sz = build(name)
    errs = np.vstack([np.zeros((1, code.n), dtype=np.uint8), np.eye(code.n, dtype=np.uint8)])
    got = emu.decode(sx, sz, planes.pack_planes(errs), planes.pack_planes(errs), len(errs), NAMED[name],
                     want=("corr", "flip", "miss"))
    assert np.array_equal(planes.unpack_planes(got["corr_x"], len(errs)), errs)
    assert np.array_equal(planes.unpack_planes(got["corr_z"], len(errs)), errs)
    assert got["tally"] == dict(shots=len(errs), fail_x=0, fail_z=0, fail_any=0, miss_x=0, miss_z=0)


@pytest.mark.parametrize("name,static", [("steane", True), ("steane", False), ("qrm15", True), ("qrm15", False)])
def test_exhaustive_enumerators(name, static):
    """All 2^n patterns of each Pauli type: failure counts equal SURVEY A.4."""
    code, sx, sz = build(name)
    pats = omc.all_patterns(code.n)
    got = emu.decode(sx, sz, planes.pack_planes(pats), planes.pack_planes(pats), len(pats),
                     NAMED[name] if static else -1)
    want = omc.tally_xz(code, pats, pats)
    assert got["tally"] == want
    if name == "steane":
        assert want["fail_x"] == 64 and want["fail_z"] == 64
    else:
        assert want["fail_x"] == 16384 and want["miss_x"] == 14336 and want["fail_z"] == 16384


def test_golay_sample_of_patterns():
    code, sx, sz = build("golay23")
    rng = np.random.default_rng(23)
    idx = rng.integers(0, 1 << 23, size=20000)
    pats = ((idx[:, None] >> np.arange(23)[None, :]) & 1).astype(np.uint8)
    check_against_oracle(code, sx, sz, pats, pats[::-1].copy(), NAMED["golay23"])
    check_against_oracle(code, sx, sz, pats, pats[::-1].copy(), -1)


@pytest.mark.parametrize("n,m1,m2", [(5, 2, 2), (12, 5, 4), (16, 8, 7), (20, 3, 12), (32, 16, 9), (9, 6, 1)])
def test_generic_random_codes_with_partial_tables(n, m1, m2):
    """Synthetic (H, L, table) triples, tables covering a random subset of the keys (misses)."""
    rng = np.random.default_rng(n * 100 + m1)
    sides = {}
    for which, m in ((1, m1), (2, m2)):
        h = rng.integers(0, 2, size=(m, n))
        lrow = rng.integers(0, 2, size=n)
        keys = rng.permutation(1 << m)[: max(1, (1 << m) * 2 // 3)]
        table = {int(k): rng.integers(0, 2, size=n) for k in keys}
        sides[which] = (h, table, lrow[None, :])
    sx = emu.Side(sides[2][0], sides[2][2][0], sides[2][1])
    sz = emu.Side(sides[1][0], sides[1][2][0], sides[1][1])
    shots = 777
    ex = rng.integers(0, 2, size=(shots, n), dtype=np.uint8)
    ez = rng.integers(0, 2, size=(shots, n), dtype=np.uint8)
    got = emu.decode(sx, sz, planes.pack_planes(ex), planes.pack_planes(ez), shots, -1,
                     want=("synd", "corr", "flip", "miss"))
    for tag, which, errs in (("x", 2, ex), ("z", 1, ez)):
        want = omc.decode_batch(*sides[which], errs)
        assert np.array_equal(planes.unpack_planes(got["synd_" + tag], shots), want["synd"])
        assert np.array_equal(planes.unpack_planes(got["corr_" + tag], shots), want["corr"])
        assert np.array_equal(planes.unpack_plane(got["flip_" + tag], shots), want["flip"])
        assert np.array_equal(planes.unpack_plane(got["miss_" + tag], shots), want["miss"])


@pytest.mark.parametrize("name", ["steane", "golay23"])
@pytest.mark.parametrize("p", [1e-3, 0.012, 0.05, 0.5, 0.0, 1.0, 1e-9])
def test_fused_sampler_bit_exact(name, p):
    """The fused sampler draws exactly the oracle sampler's bits and tallies them like the oracle."""
    code, sx, sz = build(name)
    shots, seed, first = 4000, 0x5EED1234ABCD, 128 * 7
    thr = ophilox.threshold(p)
    got = emu.decode(sx, sz, shots=shots, named_id=NAMED[name], sample=dict(seed=seed, first_shot=first, thr=thr, p=p))
    ex, ez = ophilox.sample_bits(seed, first, shots, code.n, p)
    assert np.array_equal(planes.unpack_planes(got["ex"], shots), ex)
    assert np.array_equal(planes.unpack_planes(got["ez"], shots), ez)
    assert got["tally"] == omc.tally_xz(code, ex, ez)
    if p == 0.0:
        assert not ex.any() and not ez.any()


@pytest.mark.parametrize("p", [0.3, 0.01])
def test_sampler_statistics(p):
    """Per-qubit X/Y/Z frequencies of the oracle sampler (which the kernel matches bit-for-bit): the
    bit-serial sampler (p = 0.3) and the gap sampler (p = 0.01 < 1/64), plus lane uniformity of the latter."""
    shots = 400000 if p > 0.1 else 3200000
    ex, ez = ophilox.sample_bits(99, 0, shots, 3, p)
    for freq in (np.mean(ex & ~ez & 1), np.mean(ez & ~ex & 1), np.mean(ex & ez)):
        assert abs(freq - p / 3) < 5 * np.sqrt(p / 3 * (1 - p / 3) / (3 * shots))
    err = (ex | ez)[:, 0].reshape(-1, 32).sum(axis=0)                 # errors per lane position
    expect = shots / 32 * p
    assert np.all(np.abs(err - expect) < 5 * np.sqrt(expect))


@pytest.mark.parametrize("name", list(NAMED))
@pytest.mark.parametrize("static", [True, False])
@pytest.mark.parametrize("p", [0.003, 0.03, 0.2])
def test_fast_instantiation_tallies(name, static, p):
    """The tally-only FAST instantiation (the Monte-Carlo hot path) counts exactly like the oracle,
    from loaded planes and from the fused sampler."""
    code, sx, sz = build(name)
    shots = 128 * 40
    rng = np.random.default_rng(5)
    ex = (rng.random((shots, code.n)) < p).astype(np.uint8)
    ez = (rng.random((shots, code.n)) < p).astype(np.uint8)
    nid = NAMED[name] if static else -1
    got = emu.decode(sx, sz, planes.pack_planes(ex), planes.pack_planes(ez), shots, nid, fast=True)
    assert got["tally"] == omc.tally_xz(code, ex, ez)
    got = emu.decode(sx, sz, shots=shots, named_id=nid, fast=True,
                     sample=dict(seed=77, first_shot=256, thr=ophilox.threshold(p), p=p))
    ox, oz = ophilox.sample_bits(77, 256, shots, code.n, p)
    assert got["tally"] == omc.tally_xz(code, ox, oz)


@pytest.mark.parametrize("n,m1,m2", [(12, 7, 8), (20, 13, 6), (32, 9, 11), (16, 14, 10)])
@pytest.mark.parametrize("p", [0.004, 0.3])
def test_fast_generic_partial_tables(n, m1, m2, p):
    """FAST tallies of synthetic codes whose tables miss keys (incl. possibly key 0)."""
    rng = np.random.default_rng(n + m1 * 31)
    sides = {}
    for which, m in ((1, m1), (2, m2)):
        h = rng.integers(0, 2, size=(m, n))
        lrow = rng.integers(0, 2, size=n)
        keys = rng.permutation(1 << m)[: max(1, (1 << m) * 2 // 3)]
        sides[which] = (h, {int(k): rng.integers(0, 2, size=n) for k in keys}, lrow[None, :])
    sx = emu.Side(sides[2][0], sides[2][2][0], sides[2][1])
    sz = emu.Side(sides[1][0], sides[1][2][0], sides[1][1])
    shots = 128 * 30
    ex = (rng.random((shots, n)) < p).astype(np.uint8)
    ez = (rng.random((shots, n)) < p).astype(np.uint8)
    got = emu.decode(sx, sz, planes.pack_planes(ex), planes.pack_planes(ez), shots, -1, fast=True)["tally"]
    dx = omc.decode_batch(*sides[2], ex)
    dz = omc.decode_batch(*sides[1], ez)
    assert got["fail_x"] == int(dx["flip"].sum()) and got["fail_z"] == int(dz["flip"].sum())
    assert got["fail_any"] == int((dx["flip"] | dz["flip"]).sum())
    assert got["miss_x"] == int(dx["miss"].sum()) and got["miss_z"] == int(dz["miss"].sum())


@pytest.mark.parametrize("name", list(NAMED))
@pytest.mark.parametrize("static", [True, False])
@pytest.mark.parametrize("p", [1e-3, 7e-3])
def test_two_phase_gap_sampler_replay(name, static, p):
    """The CTA-wide two-phase gap sampler (queue of erring site-words -> shared accumulators -> decode), replayed
    on the host over the kernels' own primitives, tallies exactly what the oracle sampler + oracle decode give:
    static descriptors and the generic row masks (rowbit / lbit with a runtime qubit index), W = 4 / 2 / 1 words
    per thread, several CTA iterations."""
    code, sx, sz = build(name)
    shots, seed, first = 128 * 300, 0xFACE, 128 * 5
    got = emu.mc_gapq(sx, sz, p, shots, seed, first, NAMED[name] if static else -1)
    ox, oz = ophilox.sample_bits(seed, first, shots, code.n, p)
    assert got == omc.tally_xz(code, ox, oz)


@pytest.mark.parametrize("n,cw,w0,threads,density", [(7, 5000, 3, 4, 0.06), (1, 1, 0, 1, 1.0), (3, 2048, 0, 16, 0.0),
                                                     (5, 2049, 7, 5, 0.3), (23, 4096 + 63, 1, 7, 0.02), (2, 64, 5, 3, 0.5)])
def test_host_plane_compaction_round_trip(n, cw, w0, threads, density):
    """csrc/host_compact.cpp (zero-word suppression of host planes before the copy, the product's own code) followed by
    k_zs_expand's index arithmetic restated on the host gives back the chunk, word for word: ragged block ends, empty
    and full blocks, uneven worker ranges, an offset into a padded stride."""
    rng = np.random.default_rng(n * 100 + cw)
    stride = w0 + cw + 9
    def planes_():
        p = rng.integers(1, 1 << 63, size=(n, stride), dtype=np.uint64)
        return np.where(rng.random((n, stride)) < density, p, np.uint64(0))
    ex, ez = planes_(), planes_()
    got = emu.zs_roundtrip(ex, ez, w0, cw, threads)
    if got is None:                                           # more than half the words are non-zero somewhere
        assert density >= 0.3
        return
    assert np.array_equal(got[0], ex[:, w0:w0 + cw]) and np.array_equal(got[1], ez[:, w0:w0 + cw])
