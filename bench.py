#!/usr/bin/env python3
"""bench.py -- headline benchmark of the hot path (BASELINE.json): suffix-array construction
throughput in MB/s of text, next to the CPU baseline, with the roofline of the dominant kernel.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3]

A "step" = one construction of the suffix array of one synthetic text.  At N=1 the workload is
BASELINE.json configs[1]: 1 GiB DNA-like text (sigma=4) on one B200.  N>1 (launched by torchrun,
one process per GPU): see DESIGN.md "Multi-GPU".  Prints ONE JSON line on rank 0.

  value  = whole-job MB/s with the text already resident in HBM (sab200_saca_device)
  e2e    = the same metric through the reference-facing C ABI call with HOST buffers
           (sab200_saca: pinned host text -> device -> pinned host SA inside the timed region)
  roofline     = the onesweep radix pass: 2*(K+V)*m algorithmic bytes / CUDA-event time on the
                 library's stream, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline = the oracle port (single thread, like divsufsort) on a bounded prefix of the text
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C4_BYTES = int(3.9 * (1 << 30))  # 4 187 593 113: the u32 index limit configuration

WORKLOADS = {
    "c4": ("3.9 GiB mixed text (first half uniform bytes, second half English-like), BASELINE.json configs[3]",
           C4_BYTES, "mixed"),
    "c2": ("1 GiB DNA-like text (sigma=4, planted repeats), BASELINE.json configs[1]", 1 << 30, "dna_like"),
    "c1": ("64 MiB uniform-random bytes (sigma=256), BASELINE.json configs[0]", 64 << 20, "uniform_bytes"),
    "c3": ("256 MiB repetitive text (1 MiB block, 1e-4 mutations), BASELINE.json configs[2]", 256 << 20, "repetitive"),
}
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


def make_text(workload, n, seed_shift=0):
    from suffix_array_b200 import gen
    fn = getattr(gen, WORKLOADS[workload][2])
    seed = {"c1": gen.SEED_C1, "c2": gen.SEED_C2, "c3": gen.SEED_C3, "c4": gen.SEED_C4}[workload] + seed_shift
    return fn(n, seed)


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (the recipe's clocks line).

    ONE long-lived `nvidia-smi -lms` process, started before the warm-up steps: a fresh nvidia-smi per sample
    initialises NVML on every GPU of the box each time and was seen to stall short multi-GPU steps.  Only the
    rows between begin() and end() are reported."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.all_rows = []
        self.t_begin = self.t_end = None
        self.proc = None

    def begin(self):
        self.t_begin = time.perf_counter()

    def end(self):
        self.t_end = time.perf_counter()
        self.stop_flag.set()
        if self.proc is not None:
            self.proc.terminate()

    @property
    def rows(self):
        lo = self.t_begin if self.t_begin is not None else 0.0
        hi = self.t_end if self.t_end is not None else float("inf")
        inside = [r for t, r in self.all_rows if lo <= t <= hi]
        return inside if inside else [r for _, r in self.all_rows[-1:]]  # region shorter than one period

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                parts = [x.strip() for x in line.strip().split(",")]
                if len(parts) >= 7:
                    self.all_rows.append((time.perf_counter(), parts))
                if self.stop_flag.is_set():
                    break
        except Exception:
            pass

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def cpu_baseline_run(text, sample_bytes):
    """Times the oracle port (oracle/sa_oracle.c, single thread) on the first sample_bytes of text."""
    from oracle import oracle
    sample = np.ascontiguousarray(text[:sample_bytes])
    oracle.lib()
    t0 = time.perf_counter()
    sa = oracle.saca(sample)
    dt = time.perf_counter() - t0
    assert int(sa[0]) == sample.size
    return sample.size / 1e6 / dt, dt


def run_reference(args, rank, world):
    """`--impl reference`: the reference's CPU path.  The reference cannot be compiled here (Rust crate;
    its SACA is the un-vendored cdivsufsort 2.0 = libdivsufsort, single-threaded), so this times the
    oracle port on the host cores -- one thread, as the reference uses."""
    if rank != 0:
        return
    desc, n_full, _ = WORKLOADS[args.workload]
    sample_bytes = min(n_full, args.ref_sample_mib << 20)
    text = make_text(args.workload, sample_bytes)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_baseline_run(text, min(sample_bytes, 1 << 20))
    times = []
    for _ in range(args.steps):
        _, dt = cpu_baseline_run(text, sample_bytes)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    v = sample_bytes / 1e6 / (ms / 1e3)
    sample = "first %d MiB of the %s text per step" % (sample_bytes >> 20, args.workload)
    print(json.dumps({
        "impl": "reference", "metric": "sa_construction_throughput", "value": round(v, 3), "unit": "MB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8/u32", "data": "synthetic",
        "config": {"workload": desc, "sample": sample},
        "cpu_baseline": {"value": round(v, 3), "unit": "MB/s", "cores": 1, "kind": "port", "sample": sample,
                         "note": "oracle SA-IS port; reference libdivsufsort (single thread) is not buildable here"},
        "e2e": {"value": round(v, 3), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: c2 (BASELINE.json configs[1]) at every N; c4 = the 3.9 GiB text of configs[3]")
    ap.add_argument("--n-mib", type=int, default=0, help="override the text size (MiB); 0 = the named config")
    ap.add_argument("--ref-sample-mib", type=int, default=16)
    ap.add_argument("--cpu-sample-mib", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-search", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload is None:
        # the same text at every N, so that the driver's 1 -> 8 GPU ratio is a strong-scaling figure of one
        # workload; the 3.9 GiB text of configs[3] is `--workload c4` (profiles/r01_multi_gpu.md)
        args.workload = "c2"
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from suffix_array_b200 import _lib
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = _lib.require_gpu()
    if world > 1:
        run_distributed(args, L, _lib, torch, dist, rank, local_rank, world)
        dist.destroy_process_group()
        return
    desc, n, _ = WORKLOADS[args.workload]
    if args.n_mib:
        n = args.n_mib << 20
    if n > (3 << 30):
        raise SystemExit("the 3.9 GiB text leaves no room for the device-resident AND host-buffer legs on one GPU: "
                         "run tools/c4_single.py (host-buffer entry, verified by sab200_check)")
    # N > 1: independent replicas, one text per rank (different seed) -- see DESIGN.md "Multi-GPU"
    text = make_text(args.workload, n, seed_shift=rank)
    dev = torch.device("cuda", local_rank)
    h_text = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_text.numpy()[:] = text
    h_sa = torch.empty(n + 1, dtype=torch.int32, pin_memory=True)
    d_text = h_text.to(dev)
    d_sa = torch.empty(n + 1, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    def step_device():
        _lib.check(L.sab200_saca_device(d_text.data_ptr(), n, d_sa.data_ptr(), local_rank), "sab200_saca_device")

    def step_e2e():
        _lib.check(L.sab200_saca(h_text.data_ptr(), n, h_sa.data_ptr(), 1), "sab200_saca")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    # ---- device-resident timing (value) + roofline of the radix pass
    L.sab200_set_profiling(1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler.begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pass_ms = pass_bytes = launches = pass_launches = 0
    e0.record()
    for _ in range(args.steps):
        step_device()   # blocks until the library's stream has drained
        st = _lib.last_stats()
        pass_ms += st["radix_pass_ms"]
        pass_bytes += st["radix_pass_bytes"]
        pass_launches += st["radix_pass_launches"]
        launches += st["kernel_launches"]
    e1.record()
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    stats = _lib.last_stats()
    # ---- end-to-end through the C ABI with host buffers
    if rank == 0 or world > 1:
        for _ in range(min(args.warmup, 2)):
            step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    sampler.end()
    sampler.join(timeout=2)
    if rank == 0:
        assert int(h_sa[0]) == n
    total_mb = world * n / 1e6
    value = total_mb / (dev_ms / 1e3)
    e2e = total_mb / (e2e_ms / 1e3)
    peak, peak_src = hbm_peak()
    achieved = (pass_bytes / 1e9) / (pass_ms / 1e3) if pass_ms > 0 else 0.0

    search = None
    if not args.no_search and rank == 0:
        search = bench_search(L, _lib, torch, dev, text, h_sa, n)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        sb = min(n, args.cpu_sample_mib << 20)
        v, dt = cpu_baseline_run(text, sb)
        cpu = {"value": round(v, 3), "unit": "MB/s", "cores": 1, "kind": "port",
               "sample": "first %d MiB of the same text, oracle SA-IS port, 1 thread (%.1f s)" % (sb >> 20, dt)}
    if rank == 0:
        print(json.dumps({
            "metric": "sa_construction_throughput", "value": round(value, 2), "unit": "MB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dev_ms, 3), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8 text / u64 keys / u32 ranks", "data": "synthetic",
            "config": {"workload": desc if not args.n_mib else "%s at %d MiB" % (args.workload, args.n_mib),
                       "text_bytes": n, "per_gpu": "independent replica (one text per rank)" if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (no flush needed)" if n > (126 << 20) else "working set 36n bytes > L2",
                       "rounds": stats["rounds"], "radix_passes": stats["passes"], "active": stats["active"],
                       "sigma": stats["sigma"], "symbols_per_key": stats["symbols_per_key"]},
            "roofline": {"bound": "hbm", "kernel": "onesweep_kernel<u64,u32> (LSD radix pass)", "achieved": round(achieved, 1),
                         "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": (measured_traffic() or {}).get("traffic") if args.workload == "c2" and not args.n_mib else None,
                         "traffic_note": "DRAM read+write bytes of one full-size pass (m = 2^30 records, algorithmic "
                                         "25.77e9 B) from the committed ncu capture, profiles/r01_traffic.json",
                         "peak_source": peak_src, "bytes_per_record": 24, "launches": int(pass_launches),
                         "avg_launch_ms": round(pass_ms / max(1, pass_launches), 4),
                         "share_of_step": round(pass_ms / args.steps / dev_ms, 3)},
            "e2e": {"value": round(e2e, 2), "unit": "MB/s", "ms_per_step": round(e2e_ms, 3), "h2d_bytes_per_step": n,
                    "d2h_bytes_per_step": 4 * (n + 1), "api": "sab200_saca (host buffers, pinned)"},
            "cpu_baseline": cpu, "gpu_launches": int(launches), "search": search,
            "breakdown_ms": {k: round(stats[k], 3) for k in ("total_ms", "radix_pass_ms", "hist_ms", "pack_ms", "rank_ms", "gather_ms", "group_sort_ms")},
            "group_sort": {"records": stats["group_sort_records"], "in_large_groups": stats["group_big_records"]},
            "clocks": sampler.summary()}))
    if world > 1:
        dist.destroy_process_group()


def run_distributed(args, L, _lib, torch, dist, rank, local_rank, world):
    """N > 1: ONE text sharded over the N GPUs (strong scaling: the total work is fixed as N grows).
    Distributed sample sort + prefix doubling, suffix_array_b200/dist.py; collectives = NCCL all_to_all."""
    from suffix_array_b200 import dist as sdist, gen
    dev = torch.device("cuda", local_rank)
    desc, n, _ = WORKLOADS[args.workload]
    if args.n_mib:
        n = args.n_mib << 20
    B, lo, hi = sdist.shard_bounds(n, rank, world)
    end = min(n, hi + sdist.HALO)
    if args.workload == "c4":
        shard = gen.mixed_range(n, lo, end)      # every rank generates only its own shard
    else:
        shard = make_text(args.workload, n)[lo:end]
    h_shard = torch.empty(shard.size, dtype=torch.uint8, pin_memory=True)
    h_shard.numpy()[:] = shard
    d_shard = h_shard.to(dev)
    torch.cuda.synchronize()

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    L.sab200_set_profiling(1)
    st = {}
    sampler = ClockSampler(local_rank)
    if rank == 0:  # one nvidia-smi poller per box, started before the warm-up (see ClockSampler)
        sampler.start()
    for _ in range(args.warmup):
        sdist.dist_saca(d_shard, n, dev, stats=st)
    barrier()
    sampler.begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pass_ms = pass_bytes = launches = pass_launches = 0
    e0.record()
    for _ in range(args.steps):
        sa_local, sa_off = sdist.dist_saca(d_shard, n, dev, stats=st)
        ls = _lib.last_stats()
        pass_ms += ls["radix_pass_ms"]
        pass_bytes += ls["radix_pass_bytes"]
        pass_launches += ls["radix_pass_launches"]
        launches += ls["kernel_launches"]
    e1.record()
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    # end to end: pinned host shard in, pinned host slice of the suffix array out
    h_out = torch.empty(int(st["slice"] * 1.25) + 1024, dtype=torch.int32, pin_memory=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sa_local, sa_off = sdist.dist_saca(h_shard.to(dev, non_blocking=True), n, dev, stats=st)
        if sa_local.numel() > h_out.numel():
            h_out = torch.empty(sa_local.numel(), dtype=torch.int32, pin_memory=True)
        h_out[:sa_local.numel()].copy_(sa_local, non_blocking=True)
        torch.cuda.synchronize()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    sampler.end()
    if rank == 0:
        sampler.join(timeout=2)
    slices = sum_over_ranks(st["slice"])
    assert int(slices) == n, "slices do not cover the suffix array"
    a2a = sum_over_ranks(st["all_to_all_bytes"])
    tot_pass_ms = max_over_ranks(pass_ms)
    tot_pass_bytes = sum_over_ranks(pass_bytes)
    tot_launches = sum_over_ranks(launches)
    tot_pass_launches = sum_over_ranks(pass_launches)
    peak, peak_src = hbm_peak()
    # per-GPU achieved bandwidth of the radix passes on the slowest rank
    achieved = (pass_bytes / 1e9) / (pass_ms / 1e3) if pass_ms > 0 else 0.0
    achieved = -max_over_ranks(-achieved)
    if rank == 0:
        value = n / 1e6 / (dev_ms / 1e3)
        print(json.dumps({
            "metric": "sa_construction_throughput", "value": round(value, 2), "unit": "MB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dev_ms, 3), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8 text / u64 keys / u32 ranks", "data": "synthetic",
            "config": {"workload": desc if not args.n_mib else "%s at %d MiB" % (args.workload, args.n_mib), "text_bytes": n,
                       "per_gpu": "one text block-sharded over %d GPUs: sample sort of the packed keys, then prefix "
                                  "doubling; NCCL all_to_all for keys, rank requests/answers and rank updates" % world,
                       "l2": "inputs larger than L2 (no flush needed)", "rounds": st["rounds"], "active": st["active"],
                       "symbols_per_key": st["symbols_per_key"], "collectives_per_step": st["collectives"],
                       "phase_ms_rank0": st.get("phase_ms"), "wall_ms_rank0": st.get("wall_ms"),
                       "all_to_all_bytes_per_step": int(a2a)},
            "roofline": {"bound": "hbm", "kernel": "onesweep_kernel (LSD radix pass, slowest rank)", "achieved": round(achieved, 1),
                         "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": None,
                         "peak_source": peak_src, "launches": int(tot_pass_launches),
                         "share_of_step": round(tot_pass_ms / args.steps / dev_ms, 3)},
            "e2e": {"value": round(n / 1e6 / (e2e_ms / 1e3), 2), "unit": "MB/s", "ms_per_step": round(e2e_ms, 3),
                    "h2d_bytes_per_step": n + world * sdist.HALO, "d2h_bytes_per_step": 4 * n,
                    "api": "suffix_array_b200.dist.dist_saca (pinned host shard in, pinned host SA slice out)"},
            "cpu_baseline": None, "gpu_launches": int(tot_launches), "search": None,
            "clocks": sampler.summary()}))


def bench_search(L, _lib, torch, dev, text, h_sa, n, npat=2_000_000):
    """Secondary metric (BASELINE.json configs[4] shape, reduced count): batched search_all queries/s
    on the finished index with buckets; kernel-only (patterns resident) and through the host ABI."""
    from suffix_array_b200 import gen
    sa = h_sa.numpy().view(np.uint32)
    bkt = np.empty(_lib.BKT_LEN, dtype=np.uint32)
    _lib.check(L.sab200_enable_buckets(text.ctypes.data, n, bkt.ctypes.data), "sab200_enable_buckets")
    ix = L.sab200_index_create(text.ctypes.data, n, sa.ctypes.data, n + 1, bkt.ctypes.data, 1)
    if not ix:
        return {"error": L.sab200_last_error().decode()}
    pats, offs = gen.patterns(text, npat)
    lo = np.empty(npat, dtype=np.uint32)
    hi = np.empty(npat, dtype=np.uint32)
    host_s = 1e9
    for _ in range(3):  # the first call also allocates the device-side pattern / result buffers
        t0 = time.perf_counter()
        _lib.check(L.sab200_search_all_batch(ix, pats.ctypes.data, offs.ctypes.data, npat, lo.ctypes.data, hi.ctypes.data), "search")
        host_s = min(host_s, time.perf_counter() - t0)
    d_p = torch.zeros(pats.size + 64, dtype=torch.uint8, device=dev)
    d_p[:pats.size] = torch.from_numpy(pats).to(dev)
    d_o = torch.from_numpy(offs.view(np.int64)).to(dev)
    d_lo = torch.empty(npat, dtype=torch.int32, device=dev)
    d_hi = torch.empty(npat, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        _lib.check(L.sab200_search_all_batch_device(ix, d_p.data_ptr(), d_o.data_ptr(), npat, d_lo.data_ptr(), d_hi.data_ptr()), "search_dev")
        best = min(best, time.perf_counter() - t0)
    ok = bool(np.array_equal(d_lo.cpu().numpy().view(np.uint32), lo))
    L.sab200_index_destroy(ix)
    hits = int((hi > lo).sum())
    return {"patterns": npat, "len": "8..64", "hit_fraction": round(hits / npat, 3), "queries_per_s_kernel": round(npat / best, 1),
            "queries_per_s_host_abi": round(npat / host_s, 1), "device_equals_host": ok}


if __name__ == "__main__":
    main()
