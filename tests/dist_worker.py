"""Worker of tests/test_dist_gloo.py: one rank of the multi-GPU construction driver
(suffix_array_b200/dist.py) on CPU -- gloo collectives, the SIMT-emulator build as the "device" --
compared bit-exactly with the oracle.  Launched through torch.distributed.run."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    backend = os.environ.get("SAB_DIST_BACKEND", "gloo")
    from suffix_array_b200 import _lib, gen
    from suffix_array_b200 import dist as sdist
    from oracle import oracle
    rng = np.random.default_rng(1234)
    if backend == "nccl":  # real GPUs, real library (tests/test_gpu_multi.py)
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        device = "cuda:%d" % local
        _lib.require_gpu()
        big = [gen.dna_like(24 << 20), gen.mixed(16 << 20), gen.repetitive(8 << 20, block=1 << 14), gen.uniform_bytes(16 << 20),
               np.full(1 << 20, 65, dtype=np.uint8)]
    else:
        dist.init_process_group("gloo")
        device = "cpu"
        emu = _lib._bind(ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "libsab200_emu.so")))
        _lib._lib = emu  # test-only injection of the emulator build
        big = []
    rank, world = dist.get_rank(), dist.get_world_size()
    cases = big + [np.frombuffer(b"", dtype=np.uint8), np.frombuffer(b"a", dtype=np.uint8), np.frombuffer(b"banana", dtype=np.uint8),
             np.frombuffer(b"mississippi" * 7, dtype=np.uint8), np.full(3000, 97, dtype=np.uint8),
             np.frombuffer(b"\x00\xff" * 900 + b"\x00", dtype=np.uint8), gen.dna_like(20000), gen.uniform_bytes(9000),
             gen.repetitive(12000, block=257, mut_rate=2e-3), gen.mixed(7000),
             rng.integers(0, 3, 5001, dtype=np.uint8)]
    ok = True
    for t in cases:
        n = int(t.size)
        B, lo, hi = sdist.shard_bounds(n, rank, world)
        shard = t[lo:min(n, hi + sdist.HALO)]
        st = {}
        sa_local, sa_off = sdist.dist_saca(shard, n, device, stats=st, exchange=os.environ.get("SAB_DIST_EXCHANGE", "auto"))
        full = sdist.gather_sa(sa_local, n)
        if rank == 0:
            exp = oracle.saca(t)
            good = bool(np.array_equal(full, exp))
            print("n=%d P=%d rounds=%d slices_ok=%s a2a_bytes=%d exchange=%s rebalanced=%s layout=%s lazy=%s resolved=%d " % (
                n, world, st["rounds"], good, st["all_to_all_bytes"], st["exchange"], st.get("rebalanced"), st.get("rank_layout"),
                st.get("lazy_isa"), st.get("resolved_empty", 0)), flush=True)
            ok = ok and good
    flag = torch.tensor([1 if ok else 0], device=device)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
