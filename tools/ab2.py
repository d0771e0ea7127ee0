#!/usr/bin/env python
"""Developer tool: A/B of variant builds of libsab200 inside ONE process (device-resident construction).

    python tools/ab2.py --workloads c2:1024,c3:256 --variants "" _t512x9 _dir2 --steps 4

Every variant is suffix_array_b200/libsab200<suffix>.so (built with `make -C suffix_array_b200/csrc variant
NAME=... EXTRA=...`).  The text of a workload is generated once and stays on the device; each variant builds the
suffix array `warmup` + `steps` times (CUDA events around the steps), then twice more with per-launch profiling
for the breakdown, verifies the result with its own sab200_check and releases its arena (sab200_shutdown).
Not part of the product or of the test-suite; numbers it prints are A/B evidence, not bench values."""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="c2:1024")
    ap.add_argument("--variants", nargs="*", default=[""])
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    import torch
    from suffix_array_b200 import _lib, gen
    makers = {"c1": gen.uniform_bytes, "c2": gen.dna_like, "c3": gen.repetitive, "c4": gen.mixed}
    dev = torch.device("cuda", 0)
    for wl in args.workloads.split(","):
        name, mib = wl.split(":")
        n = int(mib) << 20
        t0 = time.time()
        text = makers[name](n)
        d_text = torch.from_numpy(text).to(dev)
        d_sa = torch.empty(n + 1, dtype=torch.int32, device=dev)
        print("# %s %s MiB generated in %.1f s" % (name, mib, time.time() - t0), flush=True)
        ref = None
        for v in args.variants:
            path = os.path.join(ROOT, "suffix_array_b200", "libsab200%s.so" % v)
            if not os.path.exists(path):
                print("variant %r: missing" % v)
                continue
            L = _lib._bind(ctypes.CDLL(path))
            _lib._lib = L  # last_stats() reads through the module's handle

            def step():
                _lib.check(L.sab200_saca_device(d_text.data_ptr(), n, d_sa.data_ptr(), 0), "sab200_saca_device")

            L.sab200_set_profiling(0)
            for _ in range(args.warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            L.sab200_set_profiling(1)
            step()
            step()
            st = _lib.last_stats()
            L.sab200_set_profiling(0)
            gbs = st["radix_pass_bytes"] / 1e9 / (st["radix_pass_ms"] / 1e3) if st["radix_pass_ms"] > 0 else 0.0
            sa = d_sa.cpu().numpy().view(np.uint32)
            ok = L.sab200_check(text.ctypes.data, n, sa.ctypes.data, n + 1) == 1
            if ref is None:
                ref = sa.copy()
            same = bool(np.array_equal(ref, sa))
            print(json.dumps({"workload": wl, "variant": v, "ms": round(ms, 3), "GB/s": round(n / 1e6 / ms, 2),
                              "radix_pass_GBps": round(gbs, 1), "check": ok, "equals_first_variant": same,
                              "rounds": st["rounds"], "passes": st["passes"][:st["rounds"] + 1],
                              "breakdown": {k: round(st[k], 2) for k in ("total_ms", "radix_pass_ms", "hist_ms", "pack_ms", "rank_ms",
                                                                          "gather_ms", "group_sort_ms")},
                              "group_sort": [st["group_sort_records"], st["group_big_records"]]}), flush=True)
            L.sab200_shutdown()
        del d_text, d_sa
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
