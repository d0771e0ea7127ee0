"""The N>1 construction path on CPU: world_size 2 and 3 over gloo, every kernel step running in the
SIMT-emulator build, the assembled suffix array compared bit-exactly with the oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,port,rebalance_min,layout", [(2, 29611, None, "block"), (3, 29612, None, "block"),
                                                             (3, 29613, "0", "block"), (2, 29614, "0", "block"),
                                                             (3, 29615, None, "cyclic"), (2, 29616, "0", "cyclic"),
                                                             (3, 29617, "0", "lazy"), (2, 29618, None, "lazy")])
def test_dist_construction_gloo(emu_lib, world, port, rebalance_min, layout):
    """layout: distribution of rank[] over the ranks (block / block-cyclic).
    rebalance_min="0": the active lists are evened out across the ranks whenever they are uneven, so newly
    unique suffixes are routed to the owners of their suffix-array slices (the large-text path)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py")]
    lazy = layout == "lazy"  # lazy inverse suffix array (block layout): only active ranks travel to their owners
    layout = "block" if lazy else layout
    env = dict(os.environ, OMP_NUM_THREADS="1", SAB_RANK_LAYOUT=layout, SAB_DIST_LAZY="1" if lazy else "0",
               SAB_DIST_LAZY_MAX_ACTIVE="1.0")  # the emulator build keeps many suffixes active: take the lazy path anyway
    if rebalance_min is not None:
        env["SAB_REBALANCE_MIN"] = rebalance_min
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("slices_ok=True") == 11, out.stdout
    assert out.stdout.count("layout=" + layout) == 11, out.stdout
    if lazy:
        assert out.stdout.count("lazy=True") == 10, out.stdout  # every text but the empty one
        resolved = [int(l.split("resolved=")[1].split()[0]) for l in out.stdout.splitlines() if "lazy=True" in l]
        assert sum(1 for r in resolved if r > 0) >= 1, out.stdout
    if rebalance_min == "0":
        assert out.stdout.count("rebalanced=True") >= 2, out.stdout


def test_shard_bounds():
    from suffix_array_b200.dist import shard_bounds
    for n in (0, 1, 5, 64, 1000, 1001):
        for P in (1, 2, 3, 8):
            cover = []
            for r in range(P):
                B, lo, hi = shard_bounds(n, r, P)
                assert 0 <= lo <= hi <= n and hi - lo <= B
                cover += list(range(lo, hi))
            assert cover == list(range(n))
