"""GPU parity for the host data formats either side of the hot path: the reference's
own (shots, n) arrays transposed on the device, and sparse event lists -- every result against the oracle
(numpy restatement of css_code.py:728, bin_matrix.py:36-43, css_code.py:649-685) on the same inputs."""

import numpy as np
import pytest
import torch

from oracle import css as ocss, montecarlo as omc
from quantum_css_codes_b200 import CSSCode, SyndromeCode, codes, planes, _native

pytestmark = pytest.mark.gpu

NAMES = ["steane", "qrm15", "golay23"]
_cache = {}


def pair(name):
    if name not in _cache:
        h1, h2 = getattr(codes, name)()
        _cache[name] = (CSSCode(np.array(h1), np.array(h2)), ocss.build_css(np.array(h1), np.array(h2)))
    return _cache[name]


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("dtype", [np.uint8, np.int64, np.bool_, np.int32])
@pytest.mark.parametrize("shots", [1, 31, 1023, 1025, 70_001])
def test_shot_major_syndromes_and_decode_match_oracle(name, dtype, shots):
    """(shots, n) arrays of every dtype the reference might hold (dtype='int' is int64) go to the device as they are
    (uint8 / int64; others reduced mod 2) and are transposed there: syndromes, corrections, flips, misses, tallies."""
    code, ref = pair(name)
    rng = np.random.default_rng(shots * 7 + len(name))
    ex, ez = omc.sample_depolarizing(rng, shots, code.n, 0.08)
    ex, ez = ex.astype(dtype), ez.astype(dtype)
    for which, errs in ((2, ex), (1, ez)):
        h, table, lop = ocss.pauli_side(ref, which)
        assert np.array_equal(code.syndromes(errs, which), omc.syndromes_batch(h, errs.astype(np.int64)))
        got = code.decode(errs, which)
        want = omc.decode_batch(h, table, lop, errs.astype(np.int64))
        assert np.array_equal(got["correction"], want["corr"])
        assert np.array_equal(got["flip"], want["flip"]) and np.array_equal(got["miss"], want["miss"])
        assert got["tally"]["shots"] == shots
    assert code.decode_xz(ex, ez) == omc.tally_xz(ref, ex.astype(np.int64), ez.astype(np.int64))


def test_shot_major_negative_and_large_int64_values_use_bit_zero():
    """np.mod(x, 2) of the reference (css_code.py:39-40) equals bit 0 of a two's-complement int64."""
    code, ref = pair("steane")
    rng = np.random.default_rng(3)
    raw = rng.integers(-2**40, 2**40, size=(5000, 7), dtype=np.int64)
    bits = np.mod(raw, 2)
    assert np.array_equal(code.device.syndrome_shots(raw, 2), omc.syndromes_batch(ref.parity_check_c2, bits))


@pytest.mark.parametrize("n,m,shots", [(129, 40, 3000), (300, 130, 2500), (1600, 768, 1100), (64, 200, 4097)])
@pytest.mark.parametrize("dtype", [np.uint8, np.int64])
def test_shot_major_wide_codes_segmented_transposer(n, m, shots, dtype):
    """n or m above 128 take the segmented (64 columns per pass) form of the device transposers."""
    rng = np.random.default_rng(n + m)
    mats = []
    for rows in (m, max(1, m // 2)):
        h = np.zeros((rows, n), dtype=np.int64)
        for i in range(rows):
            h[i, rng.choice(n, size=min(7, n), replace=False)] = 1
        mats.append(h)
    code = SyndromeCode(mats[0], mats[1])
    errs = (rng.random((shots, n)) < 0.3).astype(dtype)
    for which in (1, 2):
        assert np.array_equal(code.syndromes(errs, which), omc.syndromes_batch(mats[which - 1], errs.astype(np.int64)))


def test_pack_unpack_device_round_trip():
    lib = _native.load()
    rng = np.random.default_rng(11)
    for n, shots in ((7, 5000), (23, 1024), (128, 2049), (200, 1500)):
        bits = rng.integers(0, 2, size=(shots, n), dtype=np.uint8)
        stride = planes.stride_words(shots)
        for arr in (bits, bits.astype(np.int64) * 3 - 2 * (bits.astype(np.int64) == 0)):   # odd <-> 1, even <-> 0
            src = torch.from_numpy(arr).cuda()
            pl = torch.zeros((n, stride), dtype=torch.int64, device="cuda")
            back = torch.zeros((shots, n), dtype=torch.uint8, device="cuda")
            _native.check(lib.qcss_pack_shots_dev(src.data_ptr(), arr.dtype.itemsize, n, shots, pl.data_ptr(), stride, 0))
            _native.check(lib.qcss_unpack_planes_dev(pl.data_ptr(), stride, n, shots, back.data_ptr(), 0))
            torch.cuda.synchronize()
            assert np.array_equal(pl.cpu().numpy().view(np.uint64), planes.pack_planes(bits))
            assert np.array_equal(back.cpu().numpy(), bits)


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("p,shots", [(1e-3, 400_000), (0.3, 3_000), (0.0, 1000)])
def test_sparse_events_match_oracle_tally(name, p, shots):
    code, ref = pair(name)
    rng = np.random.default_rng(int(p * 1000) + shots)
    ex, ez = omc.sample_depolarizing(rng, shots, code.n, p)
    events = planes.events_from_arrays(ex, ez)
    assert code.decode_xz_sparse(events, shots) == omc.tally_xz(ref, ex, ez)
    bx, bz = planes.arrays_from_events(events, shots, code.n)
    assert np.array_equal(bx, ex) and np.array_equal(bz, ez)


def test_sparse_events_many_chunks_and_xor_composition():
    """More events than one 4 Mi-event chunk (cuts fall between shots), and repeated (shot, qubit) events cancel."""
    code, ref = pair("golay23")
    rng = np.random.default_rng(99)
    shots = 1_200_000
    ex, ez = omc.sample_depolarizing(rng, shots, code.n, 0.3)
    events = planes.events_from_arrays(ex, ez)
    assert events.size > (4 << 20) + 1000
    want = omc.tally_xz(ref, ex, ez)
    assert code.decode_xz_sparse(events, shots) == want
    dup = np.sort(np.concatenate([events, events[:1000], events[:1000]]), kind="stable")     # each doubled pair cancels... twice = no-op
    assert code.decode_xz_sparse(dup, shots) == want
    once = np.sort(np.concatenate([events, events[:1000]]), kind="stable")                   # first 1000 events cancelled
    cx, cz = planes.arrays_from_events(once, shots, code.n)
    assert code.decode_xz_sparse(once, shots) == omc.tally_xz(ref, cx, cz)


def test_sparse_events_argument_errors():
    code, _ = pair("steane")
    ev = planes.events_from_arrays(np.eye(7, dtype=np.uint8), np.zeros((7, 7), dtype=np.uint8))
    assert code.decode_xz_sparse(ev, 7)["shots"] == 7
    with pytest.raises(ValueError, match="sorted"):
        code.decode_xz_sparse(ev[::-1].copy(), 7)
    with pytest.raises(ValueError, match="qubit"):
        code.decode_xz_sparse(np.array([(3 << 18) | (9 << 2) | 1], dtype=np.uint64), 7)          # qubit 9 >= n
    with pytest.raises(ValueError, match="qubit"):
        code.decode_xz_sparse(np.array([(3 << 18) | (1 << 2) | 0], dtype=np.uint64), 7)          # pauli 0
    with pytest.raises(ValueError, match="qubit"):
        code.decode_xz_sparse(np.array([(7 << 18) | (1 << 2) | 1], dtype=np.uint64), 7)          # shot >= shots
    assert code.decode_xz_sparse(np.zeros(0, dtype=np.uint64), 1000) == dict(shots=1000, fail_x=0, fail_z=0, fail_any=0,
                                                                           miss_x=0, miss_z=0)
    big = SyndromeCode(np.eye(40, 60, dtype=int), np.eye(40, 60, dtype=int))
    with pytest.raises(_native.NativeLibraryError):
        big.device.decode_xz_sparse(ev, 7)


@pytest.mark.parametrize("name", NAMES)
def test_events_from_planes_on_device_and_sparse_dev_decode(name):
    """Device-sampled planes -> event list on the device (sorted by shot, then qubit) == numpy's list of the same
    batch; the event-driven decode of that list == the plane kernel's tally of the batch."""
    code, ref = pair(name)
    dev = code.device
    shots, first = 3_000_017, 128 * 77
    stride = planes.stride_words(shots)
    ex = torch.zeros((code.n, stride), dtype=torch.int64, device="cuda")
    ez = torch.zeros((code.n, stride), dtype=torch.int64, device="cuda")
    dev.mc_sample_dev(4e-3, shots, 5, first, ex.data_ptr(), ez.data_ptr(), stride, 0)
    cap = 1 << 20
    events = torch.zeros(cap, dtype=torch.int64, device="cuda")
    count = torch.zeros(1, dtype=torch.int64, device="cuda")
    dev.events_from_planes_dev(ex.data_ptr(), ez.data_ptr(), stride, shots, 0, events.data_ptr(), cap, count.data_ptr(), 0)
    torch.cuda.synchronize()
    k = int(count.item())
    hx = planes.unpack_planes(ex.cpu().numpy().view(np.uint64), shots)
    hz = planes.unpack_planes(ez.cpu().numpy().view(np.uint64), shots)
    want_events = planes.events_from_arrays(hx, hz)
    assert k == want_events.size and k < cap
    assert np.array_equal(events[:k].cpu().numpy().view(np.uint64), want_events)
    tally = torch.zeros(6, dtype=torch.int64, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    dev.decode_xz_sparse_dev(events.data_ptr(), k, shots, tally.data_ptr(), status.data_ptr(), 0)
    dense = torch.zeros(6, dtype=torch.int64, device="cuda")
    dev.decode_dev(shots, 0, ex=ex.data_ptr(), ez=ez.data_ptr(), e_stride=stride, tally=dense.data_ptr())
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    assert tally[1:].tolist() == dense[1:].tolist() == [omc.tally_xz(ref, hx, hz)[f] for f in _native.TALLY_FIELDS[1:]]
    # capacity smaller than the list: the count is still the full length
    dev.events_from_planes_dev(ex.data_ptr(), ez.data_ptr(), stride, shots, 0, events.data_ptr(), 100, count.data_ptr(), 0)
    torch.cuda.synchronize()
    assert int(count.item()) == k


@pytest.mark.parametrize("make", ["four_two_two", "hamming_k2", "golay_k3"])
def test_allow_multi_logical_decode_matches_oracle(make):
    """k > 1 (opt-in extension): a shot fails when ANY of the k logical operators flips.  Corrections,
    flips (union over the rows), misses and tallies against the oracle on shared inputs; the Monte-Carlo tally against
    the oracle on the sampler's own error bits; unsupported paths say so."""
    from oracle import philox as ophilox
    if make == "four_two_two":
        h1 = h2 = np.ones((1, 4), dtype=int)
    elif make == "hamming_k2":
        h1 = np.array(codes.hamming_7_4()); h2 = h1[:2]
    else:
        h1 = np.array(codes.golay23()[0]); h2 = h1[:9]
    code = CSSCode(h1.copy(), h2.copy(), allow_multi_logical=True)
    ref = ocss.build_css(h1.copy(), h2.copy(), allow_k_not_1=True)
    assert code.k == ref.k > 1 and code.device.k == ref.k
    rng = np.random.default_rng(code.n)
    shots = 20_011
    ex, ez = omc.sample_depolarizing(rng, shots, code.n, 0.06)
    for which, errs in ((2, ex), (1, ez)):
        h, table, lop = ocss.pauli_side(ref, which)
        want = omc.decode_batch(h, table, lop, errs)
        got = code.decode(errs.astype(np.int64), which)
        assert np.array_equal(got["correction"], want["corr"])
        assert np.array_equal(got["flip"], want["flip"]) and np.array_equal(got["miss"], want["miss"])
        assert np.array_equal(code.syndromes(errs, which), want["synd"])
    assert code.decode_xz(ex, ez) == omc.tally_xz(ref, ex, ez)
    assert code.device.decode_xz_planes(planes.pack_planes(ex), planes.pack_planes(ez), shots) == omc.tally_xz(ref, ex, ez)
    got = code.monte_carlo(0.03, 70_001, seed=21, first_shot=128 * 5)
    sx, sz = ophilox.sample_bits(21, 128 * 5, 70_001, code.n, 0.03)
    assert got == omc.tally_xz(ref, sx, sz)
    single = [ocss.build_css(h1.copy(), h2.copy(), allow_k_not_1=True) for _ in range(1)][0]
    single.lz, single.lx = ref.lz[:1], ref.lx[:1]
    assert omc.tally_xz(single, sx, sz)["fail_any"] <= got["fail_any"]          # more logical rows, more ways to fail
    for call in (lambda: code.specialize(), lambda: code.decode_xz_sparse(planes.events_from_arrays(ex, ez), shots),
                 lambda: code.error_correct_monte_carlo(1e-3, 1e-3, 2, 1000)):
        with pytest.raises(_native.NativeLibraryError):
            call()


def _sparse_planes(rng, n, words, stride, density_of_chunk):
    """(n, stride) uint64 planes; ``density_of_chunk(word index array) -> probability`` that a word is non-zero."""
    out = np.zeros((n, stride), dtype=np.uint64)
    prob = density_of_chunk(np.arange(words))
    for j in range(n):
        hit = rng.random(words) < prob
        vals = (np.uint64(1) << rng.integers(0, 64, size=words).astype(np.uint64)) | \
               (rng.integers(0, 1 << 62, size=words).astype(np.uint64) * (rng.random(words) < 0.05).astype(np.uint64))
        out[j, :words] = np.where(hit, vals, np.uint64(0))
    return out


@pytest.mark.parametrize("name,shots", [("steane", (1 << 27) + 64 * 2048 * 3 + 77), ("golay23", (1 << 25) + 12345)])
@pytest.mark.parametrize("profile", ["sparse", "dense_later", "dense"])
def test_decode_xz_host_compaction_equals_plain_copy(name, shots, profile):
    """qcss_decode_xz on host planes: the compacting path (host threads suppress zero words, k_zs_expand rebuilds the
    chunk in HBM; csrc/host_compact.h) against the plain chunked copy (option host_compact = 0): identical tallies
    for sparse planes, for planes whose later chunks are too dense to compact (those go over uncompacted), for dense
    planes (the probe declines), with a ragged last chunk and a padded stride; and against the oracle on the head."""
    code = CSSCode(*[np.array(h) for h in getattr(codes, name)()])
    ref = ocss.build_css(*[np.array(h) for h in getattr(codes, name)()])
    rng = np.random.default_rng(len(name) + shots % 97)
    words = (shots + 63) // 64
    stride = ((shots + 127) // 128) * 2 + 6
    if profile == "sparse":
        dens = lambda w: np.full(w.shape, 0.06)
    elif profile == "dense_later":
        dens = lambda w: np.where(w < words // 3, 0.03, 0.9)
    else:
        dens = lambda w: np.full(w.shape, 0.7)
    ex = _sparse_planes(rng, code.n, words, stride, dens)
    ez = _sparse_planes(rng, code.n, words, stride, dens)
    tail = shots % 64
    if tail:                                                  # bits past the last shot must not exist (plane contract)
        ex[:, words - 1] &= np.uint64((1 << tail) - 1)
        ez[:, words - 1] &= np.uint64((1 << tail) - 1)
    dev = code.device
    with _native.option("host_compact", 0):
        want = dev.decode_xz_planes(ex, ez, shots)
    got = dev.decode_xz_planes(ex, ez, shots)
    assert got == want and want["shots"] == shots
    with _native.option("host_threads", 5):                   # an odd team size: uneven task ranges
        assert dev.decode_xz_planes(ex, ez, shots) == want
    head = 64 * 1000
    bits = lambda p: np.unpackbits(np.ascontiguousarray(p[:, :1000]).view(np.uint8), axis=1, bitorder="little").T
    assert dev.decode_xz_planes(np.ascontiguousarray(ex[:, :1000]), np.ascontiguousarray(ez[:, :1000]), head) == \
        omc.tally_xz(ref, bits(ex), bits(ez))


@pytest.mark.parametrize("dtype,shots", [(np.uint8, 6_000_000 + 333), (np.int64, 900_000 + 77)])
def test_decode_xz_shots_host_compaction_equals_plain_copy(dtype, shots):
    """The reference's (shots, n) arrays through qcss_decode_xz_shots: whole 16384-shot groups go through the compacting
    pipeline, the ragged rest as plain copies; identical tallies with the option off, with an odd team, and the
    oracle's on the head."""
    code = CSSCode(*[np.array(h) for h in codes.steane()])
    ref = ocss.build_css(*[np.array(h) for h in codes.steane()])
    rng = np.random.default_rng(5)
    ex = (rng.random((shots, code.n)) < 2e-3).astype(dtype)
    ez = (rng.random((shots, code.n)) < 2e-3).astype(dtype)
    if dtype is np.int64:
        ex[::1000] *= 3                                      # only bit 0 counts (np.mod(x, 2))
    with _native.option("host_compact", 0):
        want = code.decode_xz(ex, ez)
        assert code.device.last_transfer()[1] == 0
    got = code.decode_xz(ex, ez)
    sent, team = code.device.last_transfer()
    assert got == want
    if team:                                                 # boxes with fewer than eight hardware threads keep the plain copies
        assert sent < ex.nbytes // 4
    with _native.option("host_threads", 5):
        assert code.decode_xz(ex, ez) == want and code.device.last_transfer()[1] == 5
    assert code.decode_xz(ex[:50_000], ez[:50_000]) == omc.tally_xz(ref, ex[:50_000] & 1, ez[:50_000] & 1)
