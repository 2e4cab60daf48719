"""N > 1 path on CPU: world_size-2 gloo process group, shards of one Monte-Carlo job whose local
work is done by the oracle (no GPU here); the reduced tallies must equal the single-process run."""

import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from quantum_css_codes_b200 import distributed as qdist


def test_shard_ranges_partition():
    for total in (0, 1, 127, 128, 129, 1000, 10**10, 12345678):
        for world in (1, 2, 3, 4, 8):
            spans = [qdist.shard_range(total, r, world) for r in range(world)]
            assert sum(s for _, s in spans) == total
            pos = 0
            for first, shots in spans:
                assert first % 128 == 0 or shots == 0
                if shots:
                    assert first == pos
                    pos += shots


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, seed, p, out_path):
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    from oracle import css as ocss, montecarlo as omc, philox as ophilox
    from quantum_css_codes_b200 import codes, distributed as qd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref = ocss.build_css(*[np.array(h) for h in codes.steane()])

    def local_run(p_, shots, seed_, first):
        ex, ez = ophilox.sample_bits(seed_, first, shots, ref.n, p_)
        return omc.tally_xz(ref, ex, ez)

    got = qd.monte_carlo_sharded(None, p, total, seed, local_run=local_run)
    # syndrome histogram of this rank's shard (oracle keys), summed over ranks
    first, shots = qd.shard_range(total, rank, world)
    ex, _ = ophilox.sample_bits(seed, first, shots, ref.n, p)
    h, _, _ = ocss.pauli_side(ref, 2)
    hist = np.bincount(omc.keys_batch(omc.syndromes_batch(h, ex)), minlength=1 << h.shape[0]).astype(np.uint64)
    hist = qd.allreduce_histogram(hist)
    if rank == 0:
        np.save(out_path, np.array([got[k] for k in qd.TALLY_FIELDS], dtype=np.int64))
        np.save(out_path + ".hist.npy", hist)
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(tmp_path):
    from oracle import css as ocss, montecarlo as omc, philox as ophilox
    from quantum_css_codes_b200 import codes
    total, seed, p = 128 * 37 + 5, 4242, 0.07
    out = str(tmp_path / "tally.npy")
    mp.spawn(_worker, args=(2, _free_port(), total, seed, p, out), nprocs=2, join=True)
    got = np.load(out).tolist()
    ref = ocss.build_css(*[np.array(h) for h in codes.steane()])
    ex, ez = ophilox.sample_bits(seed, 0, total, ref.n, p)
    want = omc.tally_xz(ref, ex, ez)
    assert got == [want[k] for k in qdist.TALLY_FIELDS]
    h, _, _ = ocss.pauli_side(ref, 2)
    want_hist = np.bincount(omc.keys_batch(omc.syndromes_batch(h, ex)), minlength=1 << h.shape[0])
    assert np.array_equal(np.load(out + ".hist.npy"), want_hist.astype(np.uint64))
