#!/usr/bin/env python
"""
bench.py -- decoded error shots/sec for the Monte-Carlo syndrome-extraction + lookup-decode path.

Workload (BASELINE.json configs[1]): Steane [[7,1,3]], depolarising p = 1e-3, 1e10 shots per GPU,
X and Z error planes bit-packed and RESIDENT in HBM (2 x 7 planes x 1.25 GB = 17.5 GB per GPU;
far larger than the 126 MB L2, so no flush is needed between steps).  One step = one pass of the
fused syndrome + decode + logical-check + tally kernel over all resident shots, then (N > 1) one
NCCL allreduce of the tallies.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--shots S] [--impl reference]

Prints ONE JSON line (rank 0).  `value` = whole-job shots/s from HBM-resident inputs; `e2e` = the
same metric through the host-buffer C-ABI call qcss_decode_xz (pinned host planes -> H2D -> kernel
-> D2H tallies inside the timed region); `roofline` = algorithmic bytes (2n/8 per shot) over the
kernel's event-timed duration against the measured HBM copy bandwidth; `cpu_baseline` = the numpy
oracle (port of the reference's arithmetic) timed on this host.
"""

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

P_ERR = 1e-3
SEED = 0x5EED
CPU_SAMPLE_SHOTS = 1 << 22        # per worker per step of the CPU baseline
# dram__bytes_read.sum + dram__bytes_write.sum of the decode kernel per shot, from the committed
# `ncu --set full` capture of this very command (profiles/r01_steane_bench_ncu_summary.txt:
# 17.500021 GB read + 3.6 MB written for 1e10 shots).  Algorithmic bytes are 1.75 B/shot.
NCU_TRAFFIC_BYTES_PER_SHOT = {"steane": 1.7503621}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--shots", type=float, default=1e10, help="shots per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--code", default="steane", choices=["steane", "qrm15", "golay23"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the informational C3/C4/C5 timings")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle's batched numpy restatement of the reference arithmetic
# ------------------------------------------------------------------------------------------------

def _cpu_worker(task):
    code_name, seed, shots = task
    from oracle import css as ocss, montecarlo as omc
    from quantum_css_codes_b200 import codes
    code = ocss.build_css(*[np.array(h) for h in getattr(codes, code_name)()])
    rng = np.random.default_rng(seed)
    ex, ez = omc.sample_depolarizing(rng, shots, code.n, P_ERR)
    t0 = time.perf_counter()
    tally = omc.tally_xz(code, ex, ez)
    return time.perf_counter() - t0, tally


def cpu_baseline(code_name, cores, steps=1, shots=CPU_SAMPLE_SHOTS):
    """shots/s of oracle.montecarlo.tally_xz (syndrome + key + table gather + logical check for
    both Pauli types) with `cores` worker processes.  Input generation is not timed: each worker
    times only its decode, and the parallel rate is total shots / (summed decode time / cores)."""
    tasks = [(code_name, 1000 + i, shots) for i in range(cores * steps)]
    if cores == 1:
        results = [_cpu_worker(t) for t in tasks]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(cores) as pool:
            results = pool.map(_cpu_worker, tasks, chunksize=1)
    busy = max(sum(r[0] for r in results) / cores, 1e-9)
    total = shots * len(tasks)
    return total / busy, total


def run_reference(args):
    """--impl reference: the reference's CPU arithmetic (oracle port, numpy) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    times = []
    shots_per_step = CPU_SAMPLE_SHOTS * cores
    cpu_baseline(args.code, cores, shots=1 << 16)                  # warm the workers / page cache
    for _ in range(args.steps):
        rate, total = cpu_baseline(args.code, cores)
        times.append(total / rate)
    ms = 1e3 * float(np.mean(times))
    value = shots_per_step / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "decoded error shots/sec", "value": value, "unit": "shots/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic",
        "config": {"workload": f"{args.code} depolarizing p=1e-3 syndrome+lookup decode+tally",
                   "shots_per_step": shots_per_step, "note": "bounded sample of the 1e10-shot workload"},
        "cpu_baseline": {"value": value, "unit": "shots/s", "cores": cores, "kind": "port",
                         "sample": f"{shots_per_step} shots/step, numpy batched oracle, {cores} processes"},
        "e2e": {"value": value, "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------

class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.handle, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------

def run_b200(args):
    import torch
    import torch.distributed as dist
    from quantum_css_codes_b200 import CSSCode, codes, _native

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    numa_node = None
    if world > 1:
        from quantum_css_codes_b200 import distributed as qdist
        numa_node = qdist.bind_to_gpu_numa_node(local_rank)      # pinned e2e buffers local to the GPU's root
    lib = _native.load()
    _native.check(lib.qcss_set_device(local_rank))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    code = CSSCode(*[np.array(h) for h in getattr(codes, args.code)()])
    dev = code.device
    n = code.n
    shots = int(args.shots)
    shots -= shots % 128
    stride = ((shots + 127) // 128) * 2                     # uint64 words per plane
    stream = torch.cuda.current_stream().cuda_stream

    ex = torch.empty((n, stride), dtype=torch.int64, device="cuda")
    ez = torch.empty((n, stride), dtype=torch.int64, device="cuda")
    tally = torch.zeros(6, dtype=torch.int64, device="cuda")
    # synthetic resident input: the library's own Philox depolarising sampler, distinct shots per rank
    dev.mc_sample_dev(P_ERR, shots, SEED, rank * shots, ex.data_ptr(), ez.data_ptr(), stride, stream)
    torch.cuda.synchronize()

    # One step = one pass of the fused kernel over the resident batch into that step's own six tallies.
    # The job's ONE collective (shots shard with no data-path exchange) is an NCCL allreduce of
    # all K steps' tallies at the end of the run, inside the timed region.  (Measured alternatives at
    # N = 8 / N = 2: an allreduce after every step re-synchronises the ranks every 2.4 ms, 6.7x of N = 1;
    # running it on a side stream under the next step's kernel takes SMs from the one-wave kernel, slower.)
    kstart = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    kstop = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]

    def run(steps, timed):
        tallies = torch.zeros((max(steps, 1), 6), dtype=torch.int64, device="cuda")
        for i in range(steps):
            if timed:
                kstart[i].record()
            dev.decode_dev(shots, stream, ex=ex.data_ptr(), ez=ez.data_ptr(), e_stride=stride,
                           tally=tallies[i].data_ptr())
            if timed:
                kstop[i].record()
        if world > 1:
            dist.all_reduce(tallies)
        return tallies

    run(args.warmup, False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        # align the ranks' streams on the device (host threads leave the barrier milliseconds apart): an
        # untimed one-word allreduce completes at the same moment everywhere, the start event follows it
        dist.all_reduce(torch.zeros(1, dtype=torch.int64, device="cuda"))
    start.record()
    tallies = run(args.steps, True)
    stop.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    sampler.join()
    tally = tallies[args.steps - 1]
    elapsed_ms = start.elapsed_time(stop)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(kstart, kstop)]))
    result = tally.cpu().numpy().astype(np.int64)
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = world * shots / (ms_per_step / 1e3)

    # ---- end-to-end through the host-buffer C ABI (pinned planes -> H2D -> kernel -> D2H) ----
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(torch, dist, world, dev, ex, ez, n, stride, shots, args, result)
    del ex, ez
    torch.cuda.empty_cache()

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        bytes_per_shot = 2 * n / 8.0
        achieved = bytes_per_shot * shots / (kernel_ms / 1e3) / 1e9
        others = None
        if world == 1 and not args.no_others:
            others = measure_other_configs(torch, peak)
        cpu = None
        if not args.no_cpu and world == 1:
            rate, total = cpu_baseline(args.code, 1, steps=16)      # ~11 s of single-core decode
            cpu = {"value": rate, "unit": "shots/s", "cores": 1, "kind": "port",
                   "sample": f"{total} shots (numpy batched oracle: syndrome+key+table gather+logical check, X and Z)"}
        line = {
            "metric": "decoded error shots/sec", "value": value, "unit": "shots/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": f"{args.code} [[{n},1]] depolarizing p={P_ERR} shared-input syndrome+lookup decode+tally",
                       "shots_per_gpu_per_step": shots, "layout": "bit-plane, 2 x %d planes resident in HBM" % n,
                       "resident_bytes_per_gpu": int(2 * n * stride * 8),
                       "l2_policy": "inputs (%.1f GB) larger than L2, no flush" % (2 * n * stride * 8 / 1e9),
                       "kernel": dev.kernel_name(), "parallelism": f"shots sharded over {world} GPU(s), weak",
                       "rank0_numa_node": numa_node},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": (NCU_TRAFFIC_BYTES_PER_SHOT[args.code] * shots
                                     if args.code in NCU_TRAFFIC_BYTES_PER_SHOT else None),
                         "traffic_source": "profiles/r01_steane_bench_ncu_summary.txt (ncu --set full, bytes/shot x shots)",
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                         "bytes_per_shot": bytes_per_shot, "kernel_ms": kernel_ms},
            "cpu_baseline": cpu,
            "clocks": sampler.summary(),
            "e2e": e2e,
            "gpu_launches": args.steps,
            "collective": (f"one nccl all_reduce of the {args.steps} x 6 int64 step tallies, inside the timed region"
                           if world > 1 else None),
            "tally": {k: int(v) for k, v in zip(_native.TALLY_FIELDS[1:], result[1:])},
            "other_configs": others,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_other_configs(torch, peak_gbs):
    """Informational: the other BASELINE configs (C2 fused, C3, C4, C5) timed once each with CUDA events on
    resident synthetic inputs.  Not the bench metric; failures are reported, never raised."""
    from quantum_css_codes_b200 import CSSCode, SyndromeCode, codes, _native
    lib = _native.load()
    stream = torch.cuda.current_stream().cuda_stream
    out = {}

    def timed(fn, iters=3):
        fn()
        torch.cuda.synchronize()
        best = None
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            best = ms if best is None else min(best, ms)
        return best

    def guarded(name, fn):
        try:
            out[name] = fn()
        except Exception as exc:                                   # informational section only
            out[name] = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()

    def c2_fused():
        code = CSSCode(*[np.array(h) for h in codes.steane()])
        tally = torch.zeros(6, dtype=torch.int64, device="cuda")
        shots = 10_000_000_000
        ms = timed(lambda: code.device.mc_run_dev(P_ERR, shots, SEED, 0, tally.data_ptr(), stream))
        return {"workload": "steane 1e10 shots, fused Philox sampler + decode + tally (no HBM input)", "ms": ms,
                "shots_per_s": shots / ms * 1e3, "bound": "int"}

    def c3(name):
        def run():
            code = CSSCode(*[np.array(h) for h in getattr(codes, name)()])
            dev, n, shots = code.device, code.n, 1_000_000_000
            stride = ((shots + 127) // 128) * 2
            ex = torch.empty((n, stride), dtype=torch.int64, device="cuda")
            ez = torch.empty((n, stride), dtype=torch.int64, device="cuda")
            tally = torch.zeros(6, dtype=torch.int64, device="cuda")
            dev.mc_sample_dev(P_ERR, shots, SEED, 0, ex.data_ptr(), ez.data_ptr(), stride, stream)
            ms = timed(lambda: dev.decode_dev(shots, stream, ex=ex.data_ptr(), ez=ez.data_ptr(), e_stride=stride,
                                              tally=tally.data_ptr()))
            gbs = 2 * n / 8 * shots / ms / 1e6
            return {"workload": f"{name} 1e9 shots resident, syndrome + lookup decode + tally", "ms": ms,
                    "shots_per_s": shots / ms * 1e3, "bound": "hbm", "gbs": gbs, "frac": gbs / peak_gbs,
                    "kernel": dev.kernel_name()}
        return run

    def c4():
        hx, hz = codes.hgp1600()
        dev = SyndromeCode(hx, hz).device
        shots = 100_000_000
        stride = ((shots + 127) // 128) * 2
        e = torch.randint(-2**62, 2**62, (1600, stride), dtype=torch.int64, device="cuda")
        s = torch.empty((768, stride), dtype=torch.int64, device="cuda")
        ms = timed(lambda: dev.syndrome_dev(2, e.data_ptr(), stride, shots, s.data_ptr(), stride, stream))
        gbs = (1600 + 768) / 8 * shots / ms / 1e6
        out4 = {"workload": "hgp n=1600 m=768, 1e8 shots resident, syndromes of one Pauli type", "ms": ms,
                "shots_per_s": shots / ms * 1e3, "bound": "hbm", "gbs": gbs, "frac": gbs / peak_gbs,
                "kernel": dev.kernel_name(), "layout": "plane-major"}
        del e, s
        tiles = (shots + 1023) // 1024
        e = torch.randint(-2**62, 2**62, (tiles, 1600, 16), dtype=torch.int64, device="cuda")
        s = torch.empty((tiles, 768, 16), dtype=torch.int64, device="cuda")
        per_type = {}
        for which in (1, 2):
            ms_t = timed(lambda: dev.syndrome_tiles_dev(which, e.data_ptr(), shots, s.data_ptr(), stream))
            per_type[which] = ms_t
        ms_t = (per_type[1] + per_type[2]) / 2
        gbs_t = (1600 + 768) / 8 * shots / ms_t / 1e6
        out4["tile_major"] = {"workload": "same batch stored [tile of 1024 shots][plane][128 B] (qcss_syndrome_tiles_dev), "
                                          "mean of the two Pauli types",
                              "ms": ms_t, "ms_which1": per_type[1], "ms_which2": per_type[2],
                              "shots_per_s": shots / ms_t * 1e3, "gbs": gbs_t, "frac": gbs_t / peak_gbs,
                              "kernel": "tiled-ring(tile-major, cp.async.bulk)"}
        return out4

    def c5():
        batch, m, n = 4096, 1024, 2048
        mats = torch.randint(-2**62, 2**62, (batch, m, n // 64), dtype=torch.int64, device="cuda")
        outm = torch.empty_like(mats)
        rank = torch.zeros(batch, dtype=torch.int32, device="cuda")
        piv = torch.zeros((batch, m), dtype=torch.int32, device="cuda")
        ms = timed(lambda: _native.check(lib.qcss_gf2_rref_dev(mats.data_ptr(), batch, m, n, outm.data_ptr(),
                                                               rank.data_ptr(), piv.data_ptr(), stream)))
        ops = 2.52e7 * batch / ms * 1e3
        res = {"workload": "4096 x (1024 x 2048) GF(2) RREF + rank + pivots", "ms": ms,
               "matrices_per_s": batch / ms * 1e3, "bound": "int", "xor_word_ops_per_s": ops,
               "frac_of_lop3_peak_1.85e13": ops / 1.85e13, "full_rank": int((rank == m).sum().item())}
        del outm, piv
        rows = n - m + 8
        basis = torch.empty((batch, rows, n // 64), dtype=torch.int64, device="cuda")
        ovf = torch.zeros(1, dtype=torch.int32, device="cuda")
        ms_ns = timed(lambda: _native.check(lib.qcss_gf2_nullspace_dev(mats.data_ptr(), batch, m, n, rows, basis.data_ptr(),
                                                                       rank.data_ptr(), ovf.data_ptr(), stream)))
        res["with_null_space"] = {"workload": "RREF + rank + null-space basis of every matrix", "ms": ms_ns,
                                  "matrices_per_s": batch / ms_ns * 1e3, "overflow": int(ovf.item())}
        return res

    guarded("c2_fused_sampler", c2_fused)
    guarded("c3_qrm15", c3("qrm15"))
    guarded("c3_golay23", c3("golay23"))
    guarded("c4_hgp1600", c4)
    guarded("c5_gf2_rref", c5)
    return out


def measure_e2e(torch, dist, world, dev, ex, ez, n, stride, shots, args, resident_tally):
    """qcss_decode_xz on pinned host planes: every step copies 2*n planes host->device (chunked,
    overlapped with the kernels) and reads the six tallies back."""
    from quantum_css_codes_b200 import _native
    if world > 1:
        # bound the pinned host memory of an N-rank run (N x 17.5 GB otherwise): the e2e rate is measured
        # on the first 2^31 shots of each rank's resident batch
        shots = min(shots, 1 << 31)
        stride_e2e = ((shots + 127) // 128) * 2
        ex, ez = ex[:, :stride_e2e].contiguous(), ez[:, :stride_e2e].contiguous()
        stride = stride_e2e
        resident_tally = None
    nbytes = n * stride * 8
    try:
        hx, hx_ptr = _native.host_alloc(nbytes)
        hz, hz_ptr = _native.host_alloc(nbytes)
    except Exception as exc:                                     # not enough pinnable memory
        return {"value": None, "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "error": str(exc)}
    try:
        torch.cuda.synchronize()
        tx = torch.from_numpy(hx.view(np.int64)).view(n, stride)
        tz = torch.from_numpy(hz.view(np.int64)).view(n, stride)
        tx.copy_(ex)
        tz.copy_(ez)
        torch.cuda.synchronize()
        tally = dev.decode_xz_host_ptr(hx_ptr, hz_ptr, stride, shots)        # warm-up (allocates slots)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            tally = dev.decode_xz_host_ptr(hx_ptr, hz_ptr, stride, shots)
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        ok = True
        if world == 1:
            ok = [tally[k] for k in _native.TALLY_FIELDS[1:]] == [int(v) for v in resident_tally[1:]]
        return {"value": world * shots * args.e2e_steps / dt, "unit": "shots/s", "shots_per_gpu_per_step": shots,
                "h2d_bytes_per_step": int(2 * nbytes), "d2h_bytes_per_step": 48,
                "steps": args.e2e_steps, "ms_per_step": 1e3 * dt / args.e2e_steps,
                "api": "qcss_decode_xz (host planes, chunked H2D overlapped with kernels)",
                "matches_resident_tally": bool(ok)}
    finally:
        _native.host_free(hx_ptr)
        _native.host_free(hz_ptr)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
