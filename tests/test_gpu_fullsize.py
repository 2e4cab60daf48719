"""GPU parity at BASELINE.json's FULL sizes, through size-independent properties (the oracle cannot run
1e10 shots): exact expected rates, additivity over shot ranges, GF(2) linearity of the syndrome map,
idempotence of the RREF and H.N^T = 0.  Device buffers only; comparisons with torch integer ops."""

import numpy as np
import pytest

from oracle import css as ocss, gf2 as ogf2, montecarlo as omc
from quantum_css_codes_b200 import CSSCode, SyndromeCode, codes, _native

pytestmark = pytest.mark.gpu

# exact per-shot probabilities at depolarising p = 1e-3 (exhaustive enumeration)
P_STEANE_SIDE, P_STEANE_ANY = 9.304338e-06, 1.6277421e-05


def test_c2_steane_1e10_shots_fused_sampler_rates_and_additivity():
    """Config 2: 1e10 shots at p = 1e-3.  Tallies must match the exact rates within 5 sigma, and be the
    sum of the tallies of any split of the shot range (counter-based Philox streams)."""
    import torch
    code = CSSCode(*[np.array(h) for h in codes.steane()])
    shots = 10_000_000_000
    whole = code.monte_carlo(1e-3, shots, seed=0x5EED)
    assert whole["shots"] == shots and whole["miss_x"] == 0 and whole["miss_z"] == 0
    for key, rate in (("fail_x", P_STEANE_SIDE), ("fail_z", P_STEANE_SIDE), ("fail_any", P_STEANE_ANY)):
        sigma = np.sqrt(rate * (1 - rate) / shots)
        assert abs(whole[key] / shots - rate) < 5 * sigma, (key, whole[key] / shots, rate)
    cuts = [0, 128 * 1_000_003, 128 * 40_000_000, shots]
    parts = [code.monte_carlo(1e-3, b - a, seed=0x5EED, first_shot=a) for a, b in zip(cuts[:-1], cuts[1:])]
    for key in whole:
        assert whole[key] == sum(p[key] for p in parts), key
    torch.cuda.empty_cache()


def test_c2_steane_1e10_shots_resident_decode_matches_fused_run():
    """The shared-input kernel on 1e10 resident shots (the bench workload) must tally exactly what the
    fused sampler + decode kernel tallies for the same Philox planes."""
    import torch
    code = CSSCode(*[np.array(h) for h in codes.steane()])
    dev = code.device
    shots = 10_000_000_000
    stride = ((shots + 127) // 128) * 2
    ex = torch.empty((7, stride), dtype=torch.int64, device="cuda")
    ez = torch.empty((7, stride), dtype=torch.int64, device="cuda")
    tally = torch.zeros(6, dtype=torch.int64, device="cuda")
    dev.mc_sample_dev(1e-3, shots, 77, 0, ex.data_ptr(), ez.data_ptr(), stride, 0)
    dev.decode_dev(shots, 0, ex=ex.data_ptr(), ez=ez.data_ptr(), e_stride=stride, tally=tally.data_ptr())
    torch.cuda.synchronize()
    fused = code.monte_carlo(1e-3, shots, seed=77)
    got = dict(zip(_native.TALLY_FIELDS, tally.cpu().tolist()))
    got["shots"] = shots                                   # the device form leaves the shot count to the caller
    assert got == fused
    del ex, ez
    torch.cuda.empty_cache()


@pytest.mark.parametrize("name,shots", [("qrm15", 1_000_000_000), ("golay23", 1_000_000_000)])
def test_c3_1e9_shots_lookup_decode_exact_rates(name, shots):
    """Config 3 at p = 1e-2 (p = 1e-3 expects < 4 events): rates vs the exact enumerators."""
    code = CSSCode(*[np.array(h) for h in getattr(codes, name)()])
    ref = ocss.build_css(*[np.array(h) for h in getattr(codes, name)()])
    p = 1e-2
    got = code.monte_carlo(p, shots, seed=3)
    q = 2 * p / 3
    for key, mkey, which in (("fail_x", "miss_x", 2), ("fail_z", "miss_z", 1)):
        h, table, lop = ocss.pauli_side(ref, which)
        flips, misses = omc.failure_enumerator(h, table, lop) if code.n <= 15 else (None, None)
        if flips is None:                                  # Golay: published enumerator
            flips = np.array([0, 0, 0, 0, 8855, 5313, 86779, 28589, 429088, 101200, 1005928, 171304, 1180774,
                              138138, 715990, 61226, 216568, 14168, 28336, 0, 1771, 253, 23, 1])
        rate = omc.exact_rate(flips, q)
        sigma = np.sqrt(rate * (1 - rate) / shots)
        assert abs(got[key] / shots - rate) < 5 * sigma, (name, key, got[key] / shots, rate)
        if misses is not None:
            mrate = omc.exact_rate(misses, q)
            msigma = max(np.sqrt(mrate * (1 - mrate) / shots), 1e-12)
            assert abs(got[mkey] / shots - mrate) < 5 * msigma + 1e-12, (name, mkey)


def test_c4_hgp_1e8_shots_syndrome_map_is_linear():
    """Config 4: 1e8 shots, n = 1600.  s(e1 ^ e2) == s(e1) ^ s(e2) over all 768 x 1e8 syndrome bits,
    for both Pauli types, and the zero batch maps to zero."""
    import torch
    hx, hz = codes.hgp1600()
    dev = SyndromeCode(hx, hz).device
    shots = 100_000_000
    stride = ((shots + 127) // 128) * 2
    gen = torch.Generator(device="cuda").manual_seed(1600)
    e1 = torch.randint(-2**62, 2**62, (1600, stride), dtype=torch.int64, device="cuda", generator=gen)
    e2 = torch.randint(-2**62, 2**62, (1600, stride), dtype=torch.int64, device="cuda", generator=gen)
    e2 &= torch.randint(-2**62, 2**62, (1600, stride), dtype=torch.int64, device="cuda", generator=gen)
    s1 = torch.empty((768, stride), dtype=torch.int64, device="cuda")
    s2 = torch.empty_like(s1)
    s3 = torch.empty_like(s1)
    for which in (1, 2):
        dev.syndrome_dev(which, e1.data_ptr(), stride, shots, s1.data_ptr(), stride, 0)
        dev.syndrome_dev(which, e2.data_ptr(), stride, shots, s2.data_ptr(), stride, 0)
        e1 ^= e2
        dev.syndrome_dev(which, e1.data_ptr(), stride, shots, s3.data_ptr(), stride, 0)
        e1 ^= e2
        torch.cuda.synchronize()
        assert bool(torch.equal(s3, s1 ^ s2)), which
        assert bool(s1.any())
    e1.zero_()
    dev.syndrome_dev(2, e1.data_ptr(), stride, shots, s1.data_ptr(), stride, 0)
    torch.cuda.synchronize()
    assert not bool(s1.any())
    del e1, e2, s1, s2, s3
    torch.cuda.empty_cache()


def test_c4_hgp_1e8_shots_tile_major_equals_plane_major_and_is_linear():
    """Config 4 in the tile-major layout: linear over all 768 x 1e8 syndrome bits, and bit-identical to the
    plane-major kernel on the same batch (transposed on the device)."""
    import torch
    hx, hz = codes.hgp1600()
    dev = SyndromeCode(hx, hz).device
    shots = 100_000_000
    tiles = (shots + 1023) // 1024
    gen = torch.Generator(device="cuda").manual_seed(1601)
    e1 = torch.randint(-2**62, 2**62, (tiles, 1600, 16), dtype=torch.int64, device="cuda", generator=gen)
    e2 = torch.randint(-2**62, 2**62, (tiles, 1600, 16), dtype=torch.int64, device="cuda", generator=gen)
    valid_last = shots - (tiles - 1) * 1024                      # padding bits of the last tile must be zero
    word, bit = divmod(valid_last, 64)
    for t in (e1, e2):
        t[-1, :, word + 1:] = 0
        t[-1, :, word] &= (1 << bit) - 1
    s1 = torch.empty((tiles, 768, 16), dtype=torch.int64, device="cuda")
    s2 = torch.empty_like(s1)
    s3 = torch.empty_like(s1)
    for which in (1, 2):
        dev.syndrome_tiles_dev(which, e1.data_ptr(), shots, s1.data_ptr(), 0)
        dev.syndrome_tiles_dev(which, e2.data_ptr(), shots, s2.data_ptr(), 0)
        e1 ^= e2
        dev.syndrome_tiles_dev(which, e1.data_ptr(), shots, s3.data_ptr(), 0)
        e1 ^= e2
        torch.cuda.synchronize()
        assert bool(torch.equal(s3, s1 ^ s2)), which
        assert bool(s1.any())
    del s2, s3, e2
    torch.cuda.empty_cache()
    # the same bits plane-major: (tiles, n, 16) -> (n, tiles * 16)
    stride = tiles * 16
    planes_e = e1.permute(1, 0, 2).contiguous().view(1600, stride)
    del e1
    planes_s = torch.zeros((768, stride), dtype=torch.int64, device="cuda")
    dev.syndrome_dev(2, planes_e.data_ptr(), stride, shots, planes_s.data_ptr(), stride, 0)
    torch.cuda.synchronize()
    del planes_e
    want = planes_s.view(768, tiles, 16).permute(1, 0, 2)
    assert bool(torch.equal(s1, want))                           # s1 holds which = 2 from the loop above
    del s1, planes_s
    torch.cuda.empty_cache()


def test_c4_hgp_1e8_shots_fused_sampler_equals_unfused():
    """Config 4 with the sampler fused in: 1e8 shots at p = 1e-3.  The syndromes written by the fused kernel equal
    the syndromes the tile-major kernel computes from the errors the same call wrote out (all 2 x 768 x 1e8 bits),
    and the sampled error rate is p within binomial noise."""
    import torch
    hx, hz = codes.hgp1600()
    dev = SyndromeCode(hx, hz).device
    shots, p = 100_000_000, 1e-3
    tiles = (shots + 1023) // 1024
    ex = torch.empty((tiles, 1600, 16), dtype=torch.int64, device="cuda")
    ez = torch.empty_like(ex)
    sx = torch.empty((tiles, 768, 16), dtype=torch.int64, device="cuda")
    sz = torch.empty_like(sx)
    dev.sample_syndrome_tiles_dev(p, shots, 0x5EED, 0, sx.data_ptr(), sz.data_ptr(), ex.data_ptr(), ez.data_ptr(), 0)
    want = torch.empty_like(sx)
    dev.syndrome_tiles_dev(2, ex.data_ptr(), shots, want.data_ptr(), 0)
    torch.cuda.synchronize()
    assert bool(torch.equal(sx, want))
    dev.syndrome_tiles_dev(1, ez.data_ptr(), shots, want.data_ptr(), 0)
    torch.cuda.synchronize()
    assert bool(torch.equal(sz, want))
    assert bool(sx.any()) and bool(sz.any())
    # a second call without error outputs gives the same syndromes
    dev.sample_syndrome_tiles_dev(p, shots, 0x5EED, 0, want.data_ptr(), 0, 0, 0, 0)
    torch.cuda.synchronize()
    assert bool(torch.equal(sx, want))
    del want, sx, sz
    any_err = ex | ez
    del ex, ez
    count = 0
    for chunk in torch.chunk(any_err.view(-1), 64):                   # popcount without a 20 GB temporary
        v = chunk.clone()
        v = (v & 0x5555555555555555) + ((v >> 1) & 0x5555555555555555)
        v = (v & 0x3333333333333333) + ((v >> 2) & 0x3333333333333333)
        v = (v + (v >> 4)) & 0x0F0F0F0F0F0F0F0F
        count += int(((v * 0x0101010101010101) >> 56 & 0xFF).sum().item())
    total = shots * 1600
    sigma = (total * p * (1 - p)) ** 0.5
    assert abs(count - total * p) < 6 * sigma, (count, total * p)
    del any_err
    torch.cuda.empty_cache()


def test_c5_4096_matrices_rref_idempotent_rank_and_nullspace():
    """Config 5: all 4096 random 1024 x 2048 matrices on the device.  RREF(RREF(A)) == RREF(A); rank ==
    number of non-zero rows; the pivot block of the RREF is the identity; A.N^T == 0 with N from
    qcss_gf2_nullspace_dev (checked through the RREF: R.N^T has the same null space); 16 matrices are
    compared bit for bit with the oracle."""
    import torch
    lib = _native.load()
    batch, m, n, words = 4096, 1024, 2048, 32
    host = codes.random_matrices_c5(16)
    gen = torch.Generator(device="cuda").manual_seed(5)
    mats = torch.randint(-2**62, 2**62, (batch, m, words), dtype=torch.int64, device="cuda", generator=gen)
    mats[:16] = torch.from_numpy(host.view(np.int64)).cuda()
    mats[100, 7] = mats[100, 3] ^ mats[100, 4]                        # one rank-deficient matrix
    out = torch.empty_like(mats)
    out2 = torch.empty_like(mats)
    rank = torch.zeros(batch, dtype=torch.int32, device="cuda")
    piv = torch.zeros((batch, m), dtype=torch.int32, device="cuda")
    _native.check(lib.qcss_gf2_rref_dev(mats.data_ptr(), batch, m, n, out.data_ptr(), rank.data_ptr(), piv.data_ptr(), 0))
    _native.check(lib.qcss_gf2_rref_dev(out.data_ptr(), batch, m, n, out2.data_ptr(), 0, 0, 0))
    torch.cuda.synchronize()
    assert bool(torch.equal(out, out2))
    nonzero_rows = (out != 0).any(dim=2).sum(dim=1).to(torch.int32)
    assert bool(torch.equal(nonzero_rows, rank))
    assert int(rank[100]) == 1023 and int((rank == 1024).sum()) == batch - 1
    got = out[:16].cpu().numpy().view(np.uint64)
    for b in range(16):
        want, pv = ogf2.rref_packed(host[b], n)
        assert np.array_equal(got[b], want), b
        assert np.array_equal(piv[b, : len(pv)].cpu().numpy(), pv)
    # null space of every matrix: basis rows n - rank, R . x = 0 for each basis vector x
    rows = n - m + 1
    basis = torch.empty((batch, rows, words), dtype=torch.int64, device="cuda")
    ovf = torch.zeros(1, dtype=torch.int32, device="cuda")
    _native.check(lib.qcss_gf2_nullspace_dev(mats.data_ptr(), batch, m, n, rows, basis.data_ptr(), 0, ovf.data_ptr(), 0))
    torch.cuda.synchronize()
    assert int(ovf.item()) == 0
    assert int((basis[:, : n - m] != 0).any(dim=2).sum()) == batch * (n - m)     # n - rank non-zero vectors
    assert bool((basis[100, n - m] != 0).any()) and not bool((basis[0, n - m] != 0).any())
    check = (0, 5, 100, 4095)                                          # parity of <row, x> for a sample of matrices
    for b in check:
        a = torch.from_numpy(_native.unpack_bits(mats[b].cpu().numpy().view(np.uint64), n).astype(np.float32)).cuda()
        x = torch.from_numpy(_native.unpack_bits(basis[b].cpu().numpy().view(np.uint64), n).astype(np.float32)).cuda()
        prod = (a @ x.T).to(torch.int64) & 1
        assert not bool(prod.any()), b
    del mats, out, out2, basis
    torch.cuda.empty_cache()
