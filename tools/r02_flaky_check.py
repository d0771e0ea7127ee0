"""Developer tool (GPU): the three repetitive texts of tests/test_gpu_parity.py::test_filter_and_group_sort_rounds built
over and over in one process, every result checked by the oracle -- used once to confirm the fence in
split_filter_kernel after an intermittent failure of that test."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from suffix_array_b200 import SuffixArray, gen  # noqa: E402
from oracle import oracle  # noqa: E402

n = 12 << 20
texts = [gen.repetitive(n, block=b, mut_rate=m) for b, m in ((n * 10 // 44, 1e-2), (n // 3 + 17, 3e-2), (1 << 13, 1e-2))]
t0 = time.time()
bad = runs = 0
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
while time.time() - t0 < budget:
    for s in texts:
        sa = SuffixArray(s)
        runs += 1
        if not oracle.sufcheck(s, sa.sa):
            bad += 1
print("flaky check: %d constructions, %d wrong" % (runs, bad), flush=True)
