"""Single-GPU run of the 3.9 GiB mixed text (BASELINE configs[3]) through sab200_saca with host buffers:
the N=1 point of the strong-scaling line in profiles/r01_multi_gpu.md.  Verified by sab200_check
(the GPU sufcheck of sab_search.cuh, itself parity-tested against the oracle)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from suffix_array_b200 import _lib, gen  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4187593113
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    L = _lib.require_gpu()
    t0 = time.time()
    text = gen.mixed_range(n, 0, n)
    gen_s = time.time() - t0
    sa = np.empty(n + 1, dtype=np.uint32)
    L.sab200_set_profiling(1)
    out = []
    for i in range(steps):
        t0 = time.time()
        _lib.check(L.sab200_saca(text.ctypes.data_as(C.c_void_p), n, sa.ctypes.data_as(C.c_void_p), 1), "sab200_saca")
        wall = time.time() - t0
        st = _lib.last_stats()
        out.append({"wall_s": round(wall, 3), **{k: round(st[k], 2) if isinstance(st[k], float) else st[k] for k in
                    ("total_ms", "radix_pass_ms", "hist_ms", "pack_ms", "rank_ms", "gather_ms", "rounds", "passes",
                     "sigma", "symbols_per_key", "active")}})
    t0 = time.time()
    ok = L.sab200_check(text.ctypes.data_as(C.c_void_p), n, sa.ctypes.data_as(C.c_void_p), n + 1)
    print(json.dumps({"n": n, "gen_s": round(gen_s, 1), "check": int(ok), "check_s": round(time.time() - t0, 1),
                      "MBps_device": round(n / 1e6 / (out[-1]["total_ms"] / 1e3), 1), "steps": out}))


if __name__ == "__main__":
    main()
