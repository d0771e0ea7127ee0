"""Multi-GPU suffix-array construction, one rank per GPU: a thin caller of the C ABI.

The whole distributed driver (sample sort of the packed keys, prefix doubling with NCCL all-to-all per
round) lives in libsab200.so (csrc/sab_dist.cuh, csrc/sab_comm.cuh).  The reference has nothing
distributed; this is the sharded form of saca() (/root/reference/src/saca.rs:9-15).  This module only

  * creates the communicator: `Comm("nccl")` broadcasts the 128-byte NCCL unique id through
    torch.distributed and calls sab200_comm_create_nccl (the collectives then run INSIDE the library, on
    its own stream); `Comm("callbacks")` hands the library three torch.distributed collectives as C
    callbacks -- that is how the CPU tests run the same driver over gloo with the SIMT-emulator build;
  * cuts / checks the shard (`shard_bounds`, HALO) and calls sab200_saca_sharded.

A single process that owns several GPUs does not need this module at all: sab200_saca(s, n, sa, ngpus).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

HALO = 64  # include/sab200.h SAB200_SHARD_HALO
PHASES = ("alphabet", "pack+splitters", "partition_keys", "exchange_keys", "local_sort", "init_ranks", "ranks_to_owners",
          "rebalance", "rounds/requests", "rounds/lazy_lookups", "rounds/sort", "rounds/rerank", "rounds/rank_updates",
          "rounds/route_sa", "h2d", "d2h")


def shard_bounds(n, rank, world):
    B = max(1, -(-n // world))
    lo = min(rank * B, n)
    hi = min(lo + B, n)
    return B, lo, hi


def _view(ptr, nbytes):
    if not nbytes:
        return torch.empty(0, dtype=torch.uint8)
    return torch.from_numpy(np.ctypeslib.as_array((C.c_uint8 * int(nbytes)).from_address(ptr)))


class Comm:
    """One rank's communicator (sab200_comm).  kind = "nccl" | "callbacks"."""

    def __init__(self, kind="nccl", device=None, group=None):
        self.L = _lib.lib()
        self.group = group
        self.rank = dist.get_rank(group)
        self.P = dist.get_world_size(group)
        self.kind = kind
        self.device = device
        self._keep = None
        if kind == "nccl":
            _lib.require_gpu()
            dev = torch.device(device if device is not None else "cuda")
            idx = dev.index if dev.index is not None else torch.cuda.current_device()
            uid = torch.zeros(128, dtype=torch.uint8)
            if self.rank == 0:
                buf = (C.c_uint8 * 128)()
                _lib.check(self.L.sab200_comm_unique_id(buf), "sab200_comm_unique_id")
                uid = torch.tensor(list(buf), dtype=torch.uint8)
            on_gpu = dist.get_backend(group) == "nccl"
            t = uid.to(dev) if on_gpu else uid
            dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            raw = bytes(t.cpu().tolist())
            self.h = self.L.sab200_comm_create_nccl(raw, self.rank, self.P, idx)
            self.dev_index = idx
        elif kind == "callbacks":
            AG = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64)
            AR = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_uint64)
            A2A = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)
            P = self.P

            def ag(_u, send, recv, nbytes):
                try:
                    outs = [_view(recv + r * nbytes, nbytes) for r in range(P)]
                    dist.all_gather(outs, _view(send, nbytes).clone(), group=group)
                    return 0
                except Exception as e:  # noqa: BLE001 -- reported through the return code
                    print("all_gather callback:", e, flush=True)
                    return -1

            def ar(_u, buf, count):
                try:
                    t = _view(buf, 8 * count).view(torch.int64)
                    dist.all_reduce(t, group=group)
                    return 0
                except Exception as e:  # noqa: BLE001
                    print("all_reduce callback:", e, flush=True)
                    return -1

            def a2a(_u, send, sbytes, soff, recv, rbytes, roff):
                try:
                    sb = list(np.ctypeslib.as_array((C.c_uint64 * P).from_address(sbytes)))
                    so = list(np.ctypeslib.as_array((C.c_uint64 * P).from_address(soff)))
                    rb = list(np.ctypeslib.as_array((C.c_uint64 * P).from_address(rbytes)))
                    ro = list(np.ctypeslib.as_array((C.c_uint64 * P).from_address(roff)))
                    ins = [_view(send + int(so[d]), int(sb[d])).clone() for d in range(P)]
                    out = torch.empty(int(sum(rb)), dtype=torch.uint8)
                    dist.all_to_all_single(out, torch.cat(ins), output_split_sizes=[int(x) for x in rb],
                                           input_split_sizes=[int(x) for x in sb], group=group)
                    at = 0
                    for s in range(P):
                        _view(recv + int(ro[s]), int(rb[s])).copy_(out[at:at + int(rb[s])])
                        at += int(rb[s])
                    return 0
                except Exception as e:  # noqa: BLE001
                    print("all_to_all_v callback:", e, flush=True)
                    return -1

            class CB(C.Structure):
                _fields_ = [("user", C.c_void_p), ("all_gather", AG), ("all_reduce_sum_u64", AR), ("all_to_all_v", A2A)]

            cb = CB(None, AG(ag), AR(ar), A2A(a2a))
            self._keep = cb  # the library keeps the function pointers: keep the thunks alive
            self.dev_index = 0
            self.h = self.L.sab200_comm_create_callbacks(C.byref(cb), self.rank, self.P, 0)
        else:
            raise ValueError(kind)
        if not self.h:
            raise _lib.SabError("sab200_comm_create_%s failed: %s" % (kind, self.L.sab200_last_error().decode("utf-8", "replace")))

    def close(self):
        if getattr(self, "h", None):
            self.L.sab200_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def saca(self, shard, n, out=None):
        """Collective.  shard: this rank's text positions [lo, hi) + up to HALO bytes of the next shard, as a CUDA
        uint8 tensor (device-resident), or a numpy array / CPU tensor (host buffer: the H2D copy is part of the
        call).  out: optional host/device buffer for the slice (same kind as the shard).
        Returns (slice, sa_off): `slice` = u32 suffix starts for SA positions [sa_off, sa_off + len); without
        `out` a device-resident call returns a torch view of the library's arena (valid until the next call)
        and a host call returns a fresh numpy array.  The sentinel sa[0] = n is implied."""
        L = self.L
        on_dev = isinstance(shard, torch.Tensor) and shard.is_cuda
        if on_dev:
            ptr, ln = shard.data_ptr(), shard.numel()
        else:
            shard = np.ascontiguousarray(shard.numpy() if isinstance(shard, torch.Tensor) else shard, dtype=np.uint8)
            ptr, ln = shard.ctypes.data, shard.size
        slen, off, dptr = C.c_uint64(), C.c_uint64(), C.c_void_p()
        if out is None and not on_dev:
            # host call without a buffer: first the construction, then the copy out of the arena
            rc = L.sab200_saca_sharded(self.h, ptr, ln, n, 0, None, 0, 0, C.byref(slen), C.byref(off), C.byref(dptr))
            _lib.check(rc, "sab200_saca_sharded")
            res = np.empty(slen.value, dtype=np.uint32)
            if slen.value:
                _lib.check(L.sab200_copy_from_device(res.ctypes.data, dptr, slen.value * 4, self.dev_index), "sab200_copy_from_device")
            return res, off.value
        optr = ocap = 0
        if out is not None:
            optr = out.data_ptr() if isinstance(out, torch.Tensor) else out.ctypes.data
            ocap = out.numel() if isinstance(out, torch.Tensor) else out.size
        rc = L.sab200_saca_sharded(self.h, ptr, ln, n, 1 if on_dev else 0, optr, ocap, 1 if on_dev else 0, C.byref(slen),
                                   C.byref(off), C.byref(dptr))
        _lib.check(rc, "sab200_saca_sharded")
        if out is not None:
            return out[:slen.value], off.value
        return (dptr.value or 0, slen.value), off.value  # device-resident, no copy: (arena pointer, entries)

    def stats(self):
        s = _lib.DistStats()
        _lib.check(self.L.sab200_comm_stats(self.h, C.byref(s)), "sab200_comm_stats")
        return s.as_dict()


def gather_sa(sa_local, n, group=None):
    """Assembles the full suffix array (n+1 entries incl. the sentinel) on every rank (tests / verification).
    sa_local: numpy uint32 array or tensor with this rank's slice."""
    P = dist.get_world_size(group)
    loc = torch.as_tensor(np.ascontiguousarray(sa_local).view(np.int32) if isinstance(sa_local, np.ndarray) else sa_local)
    dev = loc.device if loc.is_cuda else torch.device("cuda") if dist.get_backend(group) == "nccl" else torch.device("cpu")
    loc = loc.to(dev).view(torch.int32)
    size = torch.tensor([loc.numel()], dtype=torch.int64, device=dev)
    sizes = [torch.empty_like(size) for _ in range(P)]
    dist.all_gather(sizes, size, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes + [1])
    pad = torch.zeros(mx, dtype=torch.int32, device=dev)
    pad[:loc.numel()] = loc
    parts = [torch.empty_like(pad) for _ in range(P)]
    dist.all_gather(parts, pad, group=group)
    out = np.empty(n + 1, dtype=np.uint32)
    out[0] = n
    pos = 1
    for p, s in zip(parts, sizes):
        out[pos:pos + s] = p[:s].cpu().numpy().view(np.uint32)
        pos += s
    assert pos == n + 1, "slices do not cover the suffix array"
    return out
