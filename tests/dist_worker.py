"""Worker of tests/test_dist_gloo.py and tests/test_gpu_multi.py: one rank of the multi-GPU construction
(libsab200's distributed driver through suffix_array_b200.dist.Comm), assembled suffix array compared
bit-exactly with the oracle.  Launched through torch.distributed.run.

  SAB_DIST_BACKEND=gloo  (default) CPU: the SIMT-emulator build as the "device", the collectives handed to the
                         library as callbacks over gloo
  SAB_DIST_BACKEND=nccl  real GPUs, real library, NCCL inside the library"""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def planted(rng, n, sigma, seg, copies):
    """Random text with `copies` mutated copies of one segment: few suffixes stay active after the initial sort
    (lazy inverse suffix array), but they need several doubling rounds."""
    t = rng.integers(0, sigma, n, dtype=np.uint8)
    src = t[100:100 + seg].copy()
    for c in range(copies):
        at = int(rng.integers(2000, n - seg - 10))
        t[at:at + seg] = src
        t[at + int(rng.integers(0, seg))] ^= 1
    return t


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
        _lib.require_gpu()
        comm = sdist.Comm("nccl", device="cuda:%d" % local)
        big = [gen.dna_like(24 << 20), gen.mixed(16 << 20), gen.repetitive(8 << 20, block=1 << 14), gen.uniform_bytes(16 << 20),
               np.full(1 << 20, 65, dtype=np.uint8)]
    else:
        dist.init_process_group("gloo")
        emu = _lib._bind(ctypes.CDLL(os.path.join(ROOT, "tests", "emu", os.environ.get("SAB_EMU_LIB", "libsab200_emu.so"))))
        _lib._lib = emu  # test-only injection of the emulator build
        comm = sdist.Comm("callbacks")
        big = []
    rank, world = dist.get_rank(), dist.get_world_size()
    cases = big + [np.frombuffer(b"", dtype=np.uint8), np.frombuffer(b"a", dtype=np.uint8), np.frombuffer(b"banana", dtype=np.uint8),
             np.frombuffer(b"mississippi" * 7, dtype=np.uint8), np.full(3000, 97, dtype=np.uint8),
             np.frombuffer(b"\x00\xff" * 900 + b"\x00", dtype=np.uint8), gen.dna_like(20000), gen.uniform_bytes(9000),
             gen.repetitive(12000, block=257, mut_rate=2e-3), gen.mixed(7000),
             rng.integers(0, 3, 5001, dtype=np.uint8), planted(rng, 40000, 256, 1500, 4), planted(rng, 30000, 4, 900, 3),
             np.concatenate([gen.uniform_bytes(26000), gen.english_like(6000)])]
    fuzz = int(os.environ.get("SAB_DIST_FUZZ", "0"))
    if fuzz:  # fixed-seed randomized texts (runs, repeats with mutations, mixtures): same list on every rank
        from tests import parity_cases as pc
        frng = np.random.default_rng(int(os.environ.get("SAB_DIST_FUZZ_SEED", "7")))
        cases = [pc.random_text(frng) for _ in range(fuzz)]
    ok = True
    for t in cases:
        n = int(t.size)
        B, lo, hi = sdist.shard_bounds(n, rank, world)
        shard = np.ascontiguousarray(t[lo:min(n, hi + sdist.HALO)])
        sa_local, sa_off = comm.saca(shard, n)
        st = comm.stats()
        full = sdist.gather_sa(sa_local, n)
        if rank == 0:
            exp = oracle.saca(t)
            good = bool(np.array_equal(full, exp))
            print("n=%d P=%d rounds=%d slices_ok=%s a2a_bytes=%d collectives=%d rebalanced=%s layout=%s lazy=%s resolved=%d p2p_rounds=%d fused=%d " % (
                n, world, st["rounds"], good, st["all_to_all_bytes"], st["collectives"], bool(st["rebalanced"]),
                "cyclic" if st["rank_layout"] else "block", bool(st["lazy_isa"]), st["resolved_empty"], st["p2p_rounds"],
                st["fused_exchange"]), flush=True)
            ok = ok and good
    flag = torch.tensor([1 if ok else 0], device="cuda" if backend == "nccl" else "cpu")
    dist.broadcast(flag, 0)
    comm.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
