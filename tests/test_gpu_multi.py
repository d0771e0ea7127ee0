"""Multi-GPU construction on real GPUs, bit-exact against the oracle.  Skipped on boxes with a single GPU.

  * one process per GPU (torchrun, NCCL inside libsab200 through suffix_array_b200.dist.Comm): 2 and 4 ranks,
    default policy (lazy inverse suffix array when few suffixes stay active, block-cyclic rank[] otherwise) and
    the forced variants;
  * one process, several GPUs: sab200_saca(s, n, sa, ngpus) -- the reference seam (src/saca.rs:9-15) -- through
    ctypes, ngpus = 2 and all visible devices."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,port,env", [(2, 29621, {}), (2, 29623, {"SAB_DIST_LAZY": "0"}), (2, 29626, {"SAB_P2P_MAX_RECORDS": "0"}),
                                            (2, 29627, {"SAB_DIST_P2P": "0"}), (2, 29628, {"SAB_P2P_MAX_RECORDS": "300000", "SAB_REBALANCE_MIN": "1000"}),
                                            (2, 29624, {"SAB_DIST_LAZY": "0", "SAB_RANK_LAYOUT": "block"}),
                                            (4, 29622, {}), (4, 29625, {"SAB_DIST_FUZZ": "12"})])
def test_dist_construction_nccl(gpu_lib, world, port, env):
    if gpu_lib.sab200_device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, env=dict(os.environ, SAB_DIST_BACKEND="nccl", **env), cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    expect = int(env.get("SAB_DIST_FUZZ", 0)) or 19
    assert out.stdout.count("slices_ok=True") == expect, out.stdout
    if not env:
        assert "lazy=True" in out.stdout and "layout=cyclic" in out.stdout, out.stdout
        lines = out.stdout.splitlines()
        assert any("fused=1" in l for l in lines), out.stdout                      # key exchange fused into the partition
        assert any("lazy=True" in l and "p2p_rounds=0" not in l for l in lines), out.stdout   # peer-to-peer rounds, lazy look-ups
        assert any("lazy=False" in l and "p2p_rounds=0" not in l for l in lines), out.stdout  # ... and with the complete array
    if env.get("SAB_P2P_MAX_RECORDS") == "0":
        assert all("p2p_rounds=0" in l for l in out.stdout.splitlines() if "slices_ok" in l), out.stdout
    if env.get("SAB_DIST_P2P") == "0":
        assert all("fused=0" in l and "p2p_rounds=0" in l for l in out.stdout.splitlines() if "slices_ok" in l), out.stdout
    if env.get("SAB_DIST_LAZY") == "0":
        assert "lazy=True" not in out.stdout, out.stdout
    if env.get("SAB_RANK_LAYOUT") == "block":
        assert "layout=cyclic" not in out.stdout, out.stdout


@pytest.mark.parametrize("ngpus", [2, 0])
def test_saca_single_process_multi_gpu(gpu_lib, oracle, ngpus):
    """sab200_saca(s, n, sa, ngpus > 1): one host thread + stream per GPU, NCCL communicator inside the library."""
    from suffix_array_b200 import gen, _lib
    have = gpu_lib.sab200_device_count()
    if have < 2:
        pytest.skip("needs at least 2 GPUs")
    rng = np.random.default_rng(3)
    texts = [gen.dna_like(64 << 20), gen.mixed(24 << 20), gen.repetitive(6 << 20, block=1 << 13), np.full(300000, 7, dtype=np.uint8),
             np.frombuffer(b"", dtype=np.uint8), np.frombuffer(b"banana", dtype=np.uint8), rng.integers(0, 256, 100001, dtype=np.uint8)]
    for t in texts:
        n = int(t.size)
        sa = np.zeros(n + 1, dtype=np.uint32)
        t = np.ascontiguousarray(t)
        _lib.check(gpu_lib.sab200_saca(t.ctypes.data, n, sa.ctypes.data, ngpus), "sab200_saca(ngpus=%d)" % ngpus)
        assert np.array_equal(sa, oracle.saca(t)), (n, ngpus)
    st = _lib.DistStats()
    import ctypes
    assert gpu_lib.sab200_multi_stats(0, ctypes.byref(st)) == 0 and st.nranks == (ngpus or min(have, 16))
