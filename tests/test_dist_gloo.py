"""The N>1 construction path on CPU: world_size 2 and 3 over gloo.  The distributed driver is the one inside
libsab200 (csrc/sab_dist.cuh) compiled against the SIMT emulator; its collectives are handed in as callbacks
(suffix_array_b200.dist.Comm("callbacks")) that run torch.distributed over gloo.  The assembled suffix array is
compared bit-exactly with the oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, port, env_extra, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1", **env_extra)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    return out.stdout


@pytest.mark.parametrize("world,port,lib", [(2, 29611, "libsab200_emu.so"), (2, 29613, "libsab200_emu_prod.so"),
                                            (3, 29614, "libsab200_emu_prod.so")])
def test_dist_construction_gloo(emu_lib, world, port, lib):
    """lib: the plain emulator build keeps most suffixes active after the initial sort (complete inverse suffix
    array, block-cyclic rank[] ownership); the build with the production cost-model constant leaves few
    (lazy inverse suffix array: EMPTY look-ups resolved through the keys)."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "suffix_array_b200", "csrc"), "emu-prod"], stdout=subprocess.DEVNULL)
    out = _run(world, port, {"SAB_EMU_LIB": lib, "SAB_REBALANCE_MIN": "0"})  # even out the active lists whenever they are uneven
    assert out.count("slices_ok=True") == 14, out
    if lib.endswith("prod.so"):
        assert out.count("lazy=True") >= 2, out
        resolved = [int(l.split("resolved=")[1].split()[0]) for l in out.splitlines() if "lazy=True" in l]
        assert sum(1 for r in resolved if r > 0) >= 1, out
        deep = [int(l.split("rounds=")[1].split()[0]) for l in out.splitlines() if "lazy=True" in l]
        assert max(deep) >= 4, out  # look-ups and memoised ranks over several rounds
        # evened-out lists (new SA entries routed to their slices), with and without the lazy inverse suffix array
        assert any("rebalanced=True" in l and "lazy=True" in l and "rounds=1 " not in l for l in out.splitlines()), out
        assert any("rebalanced=True" in l and "lazy=False" in l for l in out.splitlines()), out
    else:
        assert out.count("layout=cyclic") >= 5, out


@pytest.mark.parametrize("world,port,lib,layout", [(2, 29615, "libsab200_emu_prod.so", None), (3, 29616, "libsab200_emu.so", "block")])
def test_dist_randomized_gloo(emu_lib, world, port, lib, layout):
    """Fixed-seed randomized texts (tests/parity_cases.random_text) through the distributed driver."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "suffix_array_b200", "csrc"), "emu-prod"], stdout=subprocess.DEVNULL)
    env = {"SAB_EMU_LIB": lib, "SAB_DIST_FUZZ": "5"}
    if layout:
        env["SAB_RANK_LAYOUT"] = layout
    out = _run(world, port, env)
    assert out.count("slices_ok=True") == 5, out
    if layout:
        assert "layout=cyclic" not in out, out


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
