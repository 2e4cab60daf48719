"""Pauli-frame Monte Carlo of repeated Steane error correction (the gadget of
CSSCode.error_correct, css_code.py:436-470, with the frame update of quil_classical_correct,
css_code.py:649-685).  The kernels are compared with oracle/ec_rounds.py -- an error-space numpy restatement -- on
identical Philox streams, bit-exact tallies; that oracle's round model is pinned against the circuit the unmodified
reference emits in tests/test_ec_gadget.py.

CPU: the kernels' per-thread code (csrc/ec_rounds.cuh, syndrome space) run on the host (tests/hostemu)
against the oracle; model sanity properties; world_size-2 gloo sharding.  GPU: the CUDA path through the
C ABI against the oracle, against qcss_mc_run, and sharding invariance at scale."""

import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import emu
from oracle import css as ocss, ec_rounds as oec, montecarlo as omc, philox as ophilox
from quantum_css_codes_b200 import codes, distributed as qdist, _native

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMED = {"steane": 0, "qrm15": 1, "golay23": 2}
CASES = [(0.05, 0.0, 1, 999), (0.03, 0.02, 3, 1000), (1e-3, 5e-3, 4, 2048), (0.2, 1e-3, 2, 333),
         (0.0, 0.1, 2, 640), (0.004, 0.3, 2, 500), (0.05, 0.05, 0, 256),
         (2e-3, 3e-3, 3, 5000), (0.0, 5e-3, 2, 3000), (5e-3, 0.0, 2, 1000), (7e-3, 7e-3, 6, 9000),
         (0.012, 0.015, 3, 4000)]
_cache = {}


def build(name):
    if name not in _cache:
        code = ocss.build_css(*[np.array(h) for h in getattr(codes, name)()])
        sx = emu.Side(code.parity_check_c2, code.lz[0], code.c2_syndromes)
        sz = emu.Side(code.parity_check_c1, code.lx[0], code.c1_syndromes)
        _cache[name] = (code, sx, sz)
    return _cache[name]


# ---- CPU: host emulation of the device code against the oracle ---------------------------------------------

@pytest.mark.parametrize("name", list(NAMED))
@pytest.mark.parametrize("static", [True, False])
def test_hostemu_matches_oracle(name, static):
    code, sx, sz = build(name)
    nid = NAMED[name] if static else -1
    for p, q, rounds, shots in CASES:
        got = emu.ec_run(sx, sz, p, q, rounds, shots, seed=0xABCDEF12345, first_shot=256, named_id=nid)
        want = oec.ec_rounds(code, p, q, rounds, shots, seed=0xABCDEF12345, first_shot=256)
        assert got == want, (p, q, rounds, shots)


@pytest.mark.parametrize("name", list(NAMED))
def test_hostemu_queue_form_matches_oracle(name):
    """The CTA-wide two-phase EC kernel replayed on the host over its own helpers (ec_fold_draw / ec_apply_round and
    the delta-row map): oracle tallies for every case with both rates below 1/64, ragged tails included."""
    code, sx, sz = build(name)
    for p, q, rounds, shots in [c for c in CASES if c[0] < 1 / 64 and c[1] < 1 / 64] + [(3e-3, 3e-3, 2, 128 * 9 + 5)]:
        got = emu.ec_run(sx, sz, p, q, rounds, shots, seed=0xABCDEF12345, first_shot=256, named_id=NAMED[name], queue_form=True)
        want = oec.ec_rounds(code, p, q, rounds, shots, seed=0xABCDEF12345, first_shot=256)
        assert got == want, (p, q, rounds, shots)


def test_hostemu_generic_shor9():
    """A code that only has the generic kernels (n = 9, m = 2 and 6)."""
    code = ocss.build_css(*[np.array(h) for h in codes.shor9()])
    sx = emu.Side(code.parity_check_c2, code.lz[0], code.c2_syndromes)
    sz = emu.Side(code.parity_check_c1, code.lx[0], code.c1_syndromes)
    for p, q, rounds, shots in CASES[:4]:
        assert emu.ec_run(sx, sz, p, q, rounds, shots, seed=9, first_shot=0) == \
            oec.ec_rounds(code, p, q, rounds, shots, seed=9, first_shot=0)


def test_one_round_clean_ancilla_is_the_single_shot_monte_carlo():
    code, sx, sz = build("steane")
    shots, seed, p = 4096, 77, 0.04
    ex, ez = ophilox.sample_bits(seed, 128, shots, code.n, p)
    want = omc.tally_xz(code, ex, ez)
    assert oec.ec_rounds(code, p, 0.0, 1, shots, seed, 128) == want
    assert emu.ec_run(sx, sz, p, 0.0, 1, shots, seed, 128, named_id=0) == want


def test_model_properties():
    """No noise, no failures; zero rounds, nothing to decode; with a clean ancilla every round corrects
    what the previous one left, so failures grow about linearly in the rounds; a noisy ancilla hurts."""
    code, _, _ = build("steane")
    zero = dict(shots=512, fail_x=0, fail_z=0, fail_any=0, miss_x=0, miss_z=0)
    assert oec.ec_rounds(code, 0.0, 0.0, 5, 512) == zero
    assert oec.ec_rounds(code, 0.3, 0.3, 0, 512) == zero
    shots = 20000
    one = oec.ec_rounds(code, 0.02, 0.0, 1, shots, seed=3)["fail_any"]
    four = oec.ec_rounds(code, 0.02, 0.0, 4, shots, seed=3)["fail_any"]
    noisy = oec.ec_rounds(code, 0.02, 0.02, 4, shots, seed=3)["fail_any"]
    assert 2.5 * one < four < 5.5 * one
    assert noisy > 1.5 * four


def test_shards_add_up_on_the_oracle():
    code, sx, sz = build("steane")
    args = (0.03, 0.01, 3)
    whole = oec.ec_rounds(code, *args, 128 * 9 + 17, seed=5)
    parts = [oec.ec_rounds(code, *args, shots, seed=5, first_shot=first)
             for first, shots in (qdist.shard_range(128 * 9 + 17, r, 3) for r in range(3))]
    for key in qdist.TALLY_FIELDS:
        assert sum(p[key] for p in parts) == whole[key]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_path):
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    from oracle import css as ocss_, ec_rounds as oec_
    from quantum_css_codes_b200 import codes as codes_, distributed as qd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref = ocss_.build_css(*[np.array(h) for h in codes_.steane()])
    got = qd.error_correct_sharded(None, 0.03, 0.01, 3, total, seed=11,
                                   local_run=lambda p, q, r, shots, seed, first: oec_.ec_rounds(ref, p, q, r, shots, seed, first))
    if rank == 0:
        np.save(out_path, np.array([got[k] for k in qd.TALLY_FIELDS], dtype=np.int64))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding(tmp_path):
    total = 128 * 21 + 7
    out = str(tmp_path / "ec.npy")
    mp.spawn(_worker, args=(2, _free_port(), total, out), nprocs=2, join=True)
    code, _, _ = build("steane")
    want = oec.ec_rounds(code, 0.03, 0.01, 3, total, seed=11)
    assert np.load(out).tolist() == [want[k] for k in qdist.TALLY_FIELDS]


# ---- GPU: the CUDA path through the C ABI -------------------------------------------------------------------

def device_code(name):
    import css_code
    key = "dev_" + name
    if key not in _cache:
        _cache[key] = css_code.CSSCode(*[np.array(h) for h in getattr(codes, name)()])
    return _cache[key]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["steane", "qrm15", "golay23", "shor9"])
def test_gpu_matches_oracle(name):
    dev = device_code(name)
    ref = ocss.build_css(*[np.array(h) for h in getattr(codes, name)()])
    for p, q, rounds, shots in CASES + [(0.01, 0.01, 5, 40000)]:
        got = dev.error_correct_monte_carlo(p, q, rounds, shots, seed=0xABCDEF12345, first_shot=256)
        assert got == oec.ec_rounds(ref, p, q, rounds, shots, seed=0xABCDEF12345, first_shot=256), (p, q, rounds, shots)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["steane", "qrm15", "golay23"])
@pytest.mark.parametrize("p", [1e-3, 0.05])
def test_gpu_one_round_equals_mc_run(name, p):
    dev = device_code(name)
    shots = 10**7 + 13
    assert dev.error_correct_monte_carlo(p, 0.0, 1, shots, seed=21, first_shot=1280) == \
        dev.monte_carlo(p, shots, seed=21, first_shot=1280)


@pytest.mark.gpu
def test_gpu_sharding_invariance_and_rates():
    """1e8 shots of 10 rounds at p = q = 1e-3 on Steane: shards add up to the whole run exactly, X and Z
    failure counts agree within binomial noise (the code and the model are X/Z symmetric up to the order of
    the two extractions), and the failure rate is below the unencoded 10 p (2.2e-3 against 1e-2)."""
    dev = device_code("steane")
    total, args = 10**8, (1e-3, 1e-3, 10)
    whole = dev.error_correct_monte_carlo(*args, total, seed=8)
    parts = [dev.error_correct_monte_carlo(*args, shots, seed=8, first_shot=first)
             for first, shots in (qdist.shard_range(total, r, 4) for r in range(4))]
    for key in qdist.TALLY_FIELDS:
        assert sum(p[key] for p in parts) == whole[key]
    fx, fz = whole["fail_x"], whole["fail_z"]
    assert fx > 1000 and abs(fx - fz) < 0.5 * max(fx, fz)
    assert whole["fail_any"] / total < 0.5 * 10 * 1e-3
    # the first 4e6 shots as computed in the build container by the host emulation of the same device code
    assert dev.error_correct_monte_carlo(*args, 4_000_000, seed=8) == dict(
        shots=4000000, fail_x=4766, fail_z=4797, fail_any=8756, miss_x=0, miss_z=0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["steane", "qrm15", "golay23"])
def test_gpu_queue_kernel_equals_in_place_kernel(name):
    """Both error rates below 1/64: the CTA-wide two-phase kernel (k_ec_named_q) against the in-place one
    (option "gapq" = 0) on the same Philox streams, many CTA iterations, ragged tail, offset shard."""
    dev = device_code(name)
    runs = [((1e-3, 1e-3, 10, 30_000_017), {}), ((5e-3, 2e-3, 3, 10_000_000), {}),
            ((1e-4, 7e-3, 7, 3_000_001), dict(first_shot=128 * 999))]
    with _native.option("gapq", 0):
        want = [dev.error_correct_monte_carlo(*args, seed=0xEC, **kw) for args, kw in runs]
    got = [dev.error_correct_monte_carlo(*args, seed=0xEC, **kw) for args, kw in runs]
    assert got == want
    assert want[0]["fail_any"] > 0


@pytest.mark.gpu
def test_gpu_argument_errors():
    dev = device_code("steane")
    with pytest.raises(ValueError, match="rounds"):
        dev.error_correct_monte_carlo(0.1, 0.1, -1, 100)
    with pytest.raises(ValueError, match="p must be in"):
        dev.error_correct_monte_carlo(1.5, 0.1, 1, 100)
    with pytest.raises(ValueError, match="first_shot"):
        dev.error_correct_monte_carlo(0.1, 0.1, 1, 100, first_shot=5)
    assert dev.error_correct_monte_carlo(0.1, 0.1, 3, 0)["fail_any"] == 0
