"""
Multi-GPU plumbing for the Monte-Carlo path: one process per GPU (torch.distributed), shots
sharded across ranks, ONE collective -- an all-reduce (sum) of the six tallies.

Shots are independent and the fused sampler's Philox streams are keyed by the *global* shot
index, so a sharded run returns exactly the tallies of a single-GPU run of the same seed
(tests/test_gpu_decode.py::test_monte_carlo_independent_of_sharding, tests/test_distributed.py).
"""

import numpy as np

ALIGN = 128          # the kernels work in units of 128 shots; shard boundaries stay on that grid

TALLY_FIELDS = ("shots", "fail_x", "fail_z", "fail_any", "miss_x", "miss_z")


def shard_range(total_shots, rank, world_size):
    """[first_shot, first_shot + shots) of `rank`: contiguous, 128-aligned starts, covering
    [0, total_shots) exactly once over all ranks."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    units = (int(total_shots) + ALIGN - 1) // ALIGN
    lo = (units * rank) // world_size * ALIGN
    hi = (units * (rank + 1)) // world_size * ALIGN
    hi = min(hi, int(total_shots))
    lo = min(lo, int(total_shots))
    return lo, hi - lo


def allreduce_tally(tally, device=None):
    """Sum a tally dict over all ranks of the default process group (NCCL on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(tally)
    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([int(tally[k]) for k in TALLY_FIELDS], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: int(v) for k, v in zip(TALLY_FIELDS, t.cpu().tolist())}


def allreduce_histogram(hist, device=None):
    """Sum a syndrome histogram (uint64[2^m] numpy array or int64 CUDA tensor) over all ranks."""
    import torch
    import torch.distributed as dist
    is_tensor = isinstance(hist, torch.Tensor)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return hist
    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = hist if is_tensor else torch.from_numpy(np.ascontiguousarray(hist).view(np.int64)).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t if is_tensor else t.cpu().numpy().view(np.uint64)


def monte_carlo_sharded(code, p, total_shots, seed=0, rank=None, world_size=None, local_run=None):
    """Run this rank's shard of a `total_shots` Monte-Carlo job and all-reduce the tallies.

    `local_run(p, shots, seed, first_shot) -> tally dict` defaults to ``code.monte_carlo`` (the
    CUDA kernels); tests inject an oracle-backed callable to exercise the sharding on CPU."""
    import torch.distributed as dist
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    first, shots = shard_range(total_shots, rank, world_size)
    run = local_run if local_run is not None else code.monte_carlo
    local = run(p, shots, seed, first) if shots > 0 else {k: 0 for k in TALLY_FIELDS}
    return allreduce_tally(local)


def error_correct_sharded(code, p_data, p_ancilla, rounds, total_shots, seed=0, rank=None, world_size=None,
                          local_run=None):
    """The repeated-error-correction Monte Carlo (``CSSCode.error_correct_monte_carlo``) sharded like
    ``monte_carlo_sharded``: Philox streams are keyed by the global shot word, so the all-reduced tally
    is the single-GPU tally of the same seed for any number of ranks."""
    import torch.distributed as dist
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    first, shots = shard_range(total_shots, rank, world_size)
    run = local_run if local_run is not None else code.error_correct_monte_carlo
    local = run(p_data, p_ancilla, rounds, shots, seed, first) if shots > 0 else {k: 0 for k in TALLY_FIELDS}
    return allreduce_tally(local)


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that pinned host buffers
    allocated afterwards are local to the GPU's PCIe root (host-to-device copies of the end-to-end
    path otherwise cross the socket interconnect for half of the ranks).  Returns the node or None."""
    import os
    try:
        bus_id = None
        try:
            import torch
            prop = torch.cuda.get_device_properties(device_index)
            bus_id = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        except Exception:
            import pynvml
            pynvml.nvmlInit()
            raw = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device_index)).busId
            raw = raw.decode() if isinstance(raw, bytes) else raw          # "00000000:1B:00.0"
            bus_id = raw.lower()[-12:]
        with open(f"/sys/bus/pci/devices/{bus_id}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None
