"""Multi-GPU suffix-array construction: one process per GPU, torch.distributed for the plumbing.

The reference has nothing distributed; this is the sharded form of saca()
(/root/reference/src/saca.rs:9-15) for texts spread over the GPUs of one box:
distributed sample sort on the packed keys, then prefix doubling with two exchange steps per round.

    rank g owns text positions [g*B, (g+1)*B) and their rank[] entries            (B = ceil(n/P))
    1  byte histogram            all_reduce             -> common code table, key shape
    2  pack keys of own positions; P-1 splitters from an all_gather'ed key sample
    3  partition by destination (onesweep kernel, digit = #splitters <= key)       all_to_all (key, i)
    4  local radix sort -> this rank's contiguous slice of the suffix array; equal keys always land
       on one GPU, so groups never straddle GPUs and all re-ranking is local
    5  ranks to the owners of i                                                     all_to_all (i, rank)
    6  rounds h = k, 2k, ...:  requests i+h to their owners, answers back           2 x all_to_all
       local sort by (r1, r2), local re-rank, changed ranks to their owners         all_to_all

Every compute step is a CUDA kernel of libsab200 reached through include/sab200_dist.h; the
collectives are torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests, where the
"device" is the SIMT-emulator build).  Tensors are used as untyped device buffers.
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

HALO = 64
# Above this many records per rank an exchange goes through partition + all_to_all (bulk NVLink
# transfers, local random access); below it the kernels load / store the owners' blocks directly
# (no collective, but 4-byte remote accesses).  Measured cross-over on 8 B200: a few 10^7 records.
P2P_MAX_RECORDS = int(os.environ.get("SAB_P2P_MAX_RECORDS", 16 << 20))
# Active lists are evened out across the ranks (see _rebalance) when they average at least this many records
# per rank and the longest exceeds the mean by 10 %: below that a round is launch-bound anyway.
REBALANCE_MIN_RECORDS = int(os.environ.get("SAB_REBALANCE_MIN", 1 << 20))
# "block": GPU g owns the ranks of its own text shard; "cyclic": blocks of up to 1 Mi positions dealt
# round-robin (RankLayout).  Cyclic balances the owner-side work of the rounds on texts whose regions differ
# (profiles/r01_multi_gpu.md); it is exercised by the gloo tests and becomes the default once its
# peer-to-peer kernels have been validated on GPUs.
RANK_LAYOUT = os.environ.get("SAB_RANK_LAYOUT", "block")
# "1": lazy inverse suffix array (block layout): only active ranks travel to their owners; a request that
# finds EMPTY is resolved through the key of the suffix (see include/sab200_dist.h).  Pays when few suffixes
# stay active after the initial sort (1 GiB DNA-like text: 5 %; all ranks to owners is a third of the 2-GPU
# step).  Exercised by the gloo tests; not measured on GPUs yet, so off by default.
LAZY_ISA = os.environ.get("SAB_DIST_LAZY", "0") == "1"
LAZY_MAX_ACTIVE = float(os.environ.get("SAB_DIST_LAZY_MAX_ACTIVE", 0.25))  # lazy while at most this share of the suffixes is active


class RankLayout:
    """Distribution of rank[] over the GPUs (see include/sab200_dist.h): block (GPU g owns the ranks of its
    own text shard) or block-cyclic (blocks of 2^shift positions dealt round-robin).  `width`, `shift` are
    the (B, cyc_shift) arguments of the C steps; `local_len` = entries of every GPU's local array."""

    def __init__(self, n, P, kind):
        B = max(1, -(-n // P))
        if kind == "cyclic":
            # about 8 blocks per GPU at least, at most 1 Mi positions per block
            self.shift = max(0, min(20, (max(1, n // (8 * P))).bit_length() - 1))
            self.width = 1 << self.shift
            blocks = -(-(n + 1) // self.width)          # positions 0 .. n
            self.local_len = (-(-blocks // P)) << self.shift
        else:
            self.shift = -1
            self.width = B
            self.local_len = B + 1                      # the last GPU also owns position n
        self.kind = kind


def shard_bounds(n, rank, world):
    B = max(1, -(-n // world))
    lo = min(rank * B, n)
    hi = min(lo + B, n)
    return B, lo, hi


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() else C.c_void_p(0)


def _bind(L):
    if getattr(L, "_sab_dist_bound", False):
        return L
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32
    sig = {
        "sab200_dist_hist": [vp, u64, vp, i32],
        "sab200_dist_plan": [vp, u64, vp, C.POINTER(i32), C.POINTER(i32)],
        "sab200_dist_pack": [vp, u64, u64, u64, vp, i32, i32, vp, vp, i32],
        "sab200_dist_partition_keys": [vp, vp, u64, vp, i32, vp, vp, vp, i32],
        "sab200_dist_sort_pairs": [vp, vp, vp, vp, u64, i32, i32],
        "sab200_dist_init_ranks": [vp, vp, u64, u32, vp, vp, vp, vp, C.POINTER(u64), i32],
        "sab200_dist_partition_owner": [vp, vp, u64, u32, u32, i32, i32, vp, vp, vp, i32],
        "sab200_dist_scatter": [vp, vp, u64, u32, u32, i32, i32, vp, i32],
        "sab200_dist_gather": [vp, u64, u32, u32, u32, i32, i32, vp, vp, i32],
        "sab200_dist_make_keys": [vp, vp, u64, vp, i32],
        "sab200_dist_rerank": [vp, vp, u64, u32, vp, vp, vp, vp, vp, vp, C.POINTER(u64), i32],
        "sab200_dist_partition_slices": [vp, vp, u64, vp, i32, vp, vp, vp, i32],
        "sab200_dist_lazy_collect": [vp, vp, u64, u32, u64, vp, u64, vp, i32, i32, vp, vp, C.POINTER(u64), i32],
        "sab200_dist_lower_bound": [vp, u64, vp, u64, u32, vp, i32],
        "sab200_dist_lazy_fill": [vp, vp, u64, vp, u32, u32, vp, vp, i32],
        "sab200_dist_begin": [i32],
        "sab200_dist_end": [i32],
        "sab200_dist_gather_p2p": [vp, vp, u64, u32, u32, i32, i32, vp, vp, i32],
        "sab200_dist_count_keys": [vp, u64, vp, i32, vp, i32],
        "sab200_dist_partition_keys_p2p": [vp, vp, u64, vp, i32, vp, vp, vp, i32],
        "sab200_dist_scatter_p2p": [vp, vp, u64, u32, i32, i32, vp, i32],
    }
    for name, args in sig.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = i32
    L._sab_dist_bound = True
    return L


class _Ctx:
    def __init__(self, device, group):
        self.L = _bind(_lib.lib())
        self.device = torch.device(device)
        self.dev = self.device.index if self.device.type == "cuda" else 0
        self.group = group
        self.rank = dist.get_rank(group)
        self.P = dist.get_world_size(group)
        self.a2a_bytes = 0
        self.collectives = 0
        self.phase_ms = {}
        self.tracing = os.environ.get("SAB_DIST_TRACE", "0") == "1"
        self.resolved_empty = 0
        self._t = None

    def mark(self, name):
        """Phase timer (host clock around synchronised steps): time since the previous mark goes to `name`."""
        import time
        self.sync()
        now = time.perf_counter()
        if self._t is not None:
            self.phase_ms[name] = self.phase_ms.get(name, 0.0) + (now - self._t) * 1e3
        self._t = now

    def trace(self, name):
        """Sub-phase timer of the doubling rounds; only with SAB_DIST_TRACE=1 (it adds a device sync per mark,
        so traced runs are for attribution, not for the headline number)."""
        if self.tracing:
            self.mark(name)

    def check(self, rc, what):
        if rc < 0:
            _lib.check(rc, what)
        return rc

    def call(self, name, *args):
        """One library step.  The library runs on its own stream and synchronises it before returning;
        torch work queued on the current stream (fills, copies, collectives) must be complete before the
        step touches those buffers, hence the synchronise here."""
        self.sync()
        return self.check(getattr(self.L, name)(*args), name)

    def sync(self):
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()

    def empty(self, n, dtype):
        return torch.empty(max(int(n), 1), dtype=dtype, device=self.device)[:int(n)]

    def exchange_counts(self, send_counts):
        sc = torch.tensor([int(x) for x in send_counts], dtype=torch.int64, device=self.device)
        rc = torch.empty(self.P, dtype=torch.int64, device=self.device)
        dist.all_to_all_single(rc, sc, group=self.group)
        self.collectives += 1
        return [int(x) for x in rc.tolist()]

    def all_to_all(self, buf, send_counts, recv_counts):
        """buf: 1-D tensor laid out rank-major with send_counts elements per destination."""
        out = self.empty(sum(recv_counts), buf.dtype)
        src = buf[:sum(send_counts)]
        dist.all_to_all_single(out, src.contiguous(), output_split_sizes=list(recv_counts),
                               input_split_sizes=list(send_counts), group=self.group)
        self.sync()
        self.a2a_bytes += src.numel() * src.element_size()
        self.collectives += 1
        return out


_PEER_CACHE = {}


def _peer_recv(cx, cap):
    """Symmetric receive buffers of the fused key exchange: `cap` (key u64, index u32) records per GPU.
    Returns (keys tensor, idx tensor, peer key addresses, peer idx addresses)."""
    key = ("recv", str(cx.device), cx.P, cap, id(cx.group))
    if key not in _PEER_CACHE:
        for old in [k_ for k_ in _PEER_CACHE if k_[0] == "recv"]:  # one live set of receive buffers
            del _PEER_CACHE[old]
        import torch.distributed._symmetric_memory as symm
        grp = cx.group if cx.group is not None else dist.group.WORLD
        tk = symm.empty(cap, dtype=torch.int64, device=cx.device)
        hk = symm.rendezvous(tk, grp)
        ti = symm.empty(cap, dtype=torch.int32, device=cx.device)
        hi = symm.rendezvous(ti, grp)
        _PEER_CACHE[key] = (tk, ti, np.array([int(p) for p in hk.buffer_ptrs], dtype=np.uint64),
                            np.array([int(p) for p in hi.buffer_ptrs], dtype=np.uint64), hk, hi)
    return _PEER_CACHE[key][:4]


def _peer_ranks(cx, length):
    """The rank[] array of every GPU (`length` u32 each), mapped into this process through torch's
    symmetric memory: returns (local array tensor, uint64 array of the P peer addresses).
    Cached per (device, P, length): the rendezvous is a collective and not cheap."""
    key = ("rank", str(cx.device), cx.P, length, id(cx.group))
    if key not in _PEER_CACHE:
        for old in [k_ for k_ in _PEER_CACHE if k_[0] == "rank"]:
            del _PEER_CACHE[old]
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(length, dtype=torch.int32, device=cx.device)
        hdl = symm.rendezvous(t, cx.group if cx.group is not None else dist.group.WORLD)
        ptrs = np.array([int(p) for p in hdl.buffer_ptrs], dtype=np.uint64)
        _PEER_CACHE[key] = (t, ptrs, hdl)
    t, ptrs, _ = _PEER_CACHE[key]
    return t, ptrs


def _barrier(cx):
    dist.barrier(group=cx.group)
    cx.sync()
    cx.collectives += 1


def _to_owner(cx, keys, vals, count, add, lay):
    """Stable partition of (keys, vals) by the owner of position keys+add; returns the partitioned
    buffers and the per-destination counts (records whose key is 0xFFFFFFFF are dropped)."""
    kp = cx.empty(count, torch.int32)
    vp = cx.empty(count, torch.int32)
    cnt = np.zeros(cx.P, dtype=np.uint64)
    cx.call("sab200_dist_partition_owner", _p(keys), _p(vals), count, add, lay.width, cx.P, lay.shift, _p(kp), _p(vp),
                                              cnt.ctypes.data_as(C.c_void_p), cx.dev)
    return kp, vp, [int(x) for x in cnt]


def _send_ranks(cx, idx, ranks, count, lay, lo, rank_local):
    """rank[idx[t]] = ranks[t] on the GPU that owns text position idx[t]."""
    kp, vp, send = _to_owner(cx, idx, ranks, count, 0, lay)
    recv = cx.exchange_counts(send)
    ri = cx.all_to_all(kp, send, recv)
    rr = cx.all_to_all(vp, send, recv)
    cx.call("sab200_dist_scatter", _p(ri), _p(rr), ri.numel(), lo, lay.width, cx.P, lay.shift, _p(rank_local), cx.dev)


def _next_group_boundary(r1, c, m):
    """Smallest p >= c that starts a group of the rank-sorted list r1[:m] (p = m if none)."""
    if c <= 0:
        return 0
    if c >= m:
        return m
    v = r1[c - 1]
    p, w = c, 4096
    while p < m:
        seg = r1[p:min(m, p + w)]
        ne = (seg != v).nonzero()
        if ne.numel():
            return p + int(ne[0])
        p += seg.numel()
        w *= 4
    return m


def _rebalance(cx, act_r1, act_idx, m):
    """Evens out the active lists.  The suffix-array slices hold equal numbers of SUFFIXES, not of active
    ones (on the mixed text the English-like key ranges hold nearly all of them), and every round costs
    what its longest list costs.  The lists are globally sorted by rank and groups never straddle ranks,
    so moving cut points to group boundaries and shipping contiguous chunks keeps both properties.
    Returns (r1, idx, m, moved)."""
    P, rank = cx.P, cx.rank
    mine = torch.tensor([m], dtype=torch.int64, device=cx.device)
    allm = [torch.empty_like(mine) for _ in range(P)]
    dist.all_gather(allm, mine, group=cx.group)
    cx.collectives += 1
    ms = [int(x.item()) for x in allm]
    M = sum(ms)
    if M < REBALANCE_MIN_RECORDS * P or max(ms) * P <= 1.1 * M:
        return act_r1[:m], act_idx[:m], m, False
    Q = -(-M // P)
    off = sum(ms[:rank])
    bounds = [0]
    for j in range(1, P):
        c = min(max(j * Q - off, 0), m)
        bounds.append(max(bounds[-1], _next_group_boundary(act_r1, c, m)))
    bounds.append(m)
    send = [bounds[j + 1] - bounds[j] for j in range(P)]
    recv = cx.exchange_counts(send)
    r1 = cx.all_to_all(act_r1[:m], send, recv)
    idx = cx.all_to_all(act_idx[:m], send, recv)
    return r1, idx, r1.numel(), True


def _resolve_empty(cx, q, ans, h, lo, d_text, n, lut, b, k, splitters, sorted_keys, R, sa_off, rank_local):
    """Lazy inverse suffix array: the requests of this round that found EMPTY at this owner are resolved
    through their keys (owner -> GPU holding the key -> owner) and memoised.  Collective: every rank calls it."""
    cnt_q = q.numel()
    keys_u = cx.empty(cnt_q, torch.int64)
    slot_u = cx.empty(cnt_q, torch.int32)
    nu = C.c_uint64()
    cx.call("sab200_dist_lazy_collect", _p(q), _p(ans), cnt_q, h, lo, _p(d_text), n, lut.ctypes.data_as(C.c_void_p), b, k,
            _p(keys_u), _p(slot_u), C.byref(nu), cx.dev)
    nu = nu.value
    kp = cx.empty(nu, torch.int64)
    sp = cx.empty(nu, torch.int32)
    cnt = np.zeros(cx.P, dtype=np.uint64)
    cx.call("sab200_dist_partition_keys", _p(keys_u), _p(slot_u), nu, splitters.ctypes.data_as(C.c_void_p), cx.P - 1, _p(kp), _p(sp),
            cnt.ctypes.data_as(C.c_void_p), cx.dev)
    send = [int(x) for x in cnt]
    recv = cx.exchange_counts(send)
    kq = cx.all_to_all(kp, send, recv)
    rk = cx.empty(kq.numel(), torch.int32)
    cx.call("sab200_dist_lower_bound", _p(sorted_keys), R, _p(kq), kq.numel(), sa_off, _p(rk), cx.dev)
    back = cx.all_to_all(rk, recv, send)
    cx.call("sab200_dist_lazy_fill", _p(sp), _p(back), nu, _p(q), h, lo, _p(ans), _p(rank_local), cx.dev)
    cx.resolved_empty += nu


def _send_sa(cx, pos, idx, count, starts, sa_off, sa_local):
    """sa[pos[t]] = idx[t] on the GPU whose slice holds SA position pos[t] (entries with pos = 0xFFFFFFFF are dropped)."""
    kp = cx.empty(count, torch.int32)
    vp = cx.empty(count, torch.int32)
    cnt = np.zeros(cx.P, dtype=np.uint64)
    cx.call("sab200_dist_partition_slices", _p(pos), _p(idx), count, starts.ctypes.data_as(C.c_void_p), cx.P, _p(kp), _p(vp),
            cnt.ctypes.data_as(C.c_void_p), cx.dev)
    send = [int(x) for x in cnt]
    recv = cx.exchange_counts(send)
    rp = cx.all_to_all(kp, send, recv)
    ri = cx.all_to_all(vp, send, recv)
    cx.call("sab200_dist_scatter", _p(rp), _p(ri), rp.numel(), sa_off, 1, cx.P, -1, _p(sa_local), cx.dev)  # slices are contiguous


def dist_saca(shard, n, device, group=None, stats=None, exchange="auto"):
    """Builds the suffix array of a text of n bytes spread over the ranks of `group`.

    exchange: how rank[] crosses GPUs in the doubling rounds.  "p2p" = the gather / update kernels load
    and store the owners' blocks directly over NVLink (symmetric memory; no collective in the data path
    of a round); "nccl" = partition by owner + all_to_all (also the gloo path of the CPU tests);
    "auto" = p2p on CUDA devices when symmetric memory can be set up, else nccl.

    shard: uint8 numpy array or tensor with this rank's text positions [lo, hi) followed by up to HALO
    bytes of the next shard (text[lo : min(n, hi + HALO)], see shard_bounds).
    Returns (sa_local, sa_off): this rank's slice of the suffix array -- int32 tensor holding u32
    suffix indices for SA positions [sa_off, sa_off + len) -- the sentinel entry sa[0] = n is implied
    (rank 0's slice starts at position 1)."""
    import time
    t_enter = time.perf_counter()
    cx = _Ctx(device, group)
    L, P, rank = cx.L, cx.P, cx.rank
    if n > _lib.MAX_LENGTH:
        raise ValueError("text longer than MAX_LENGTH")
    B, lo, hi = shard_bounds(n, rank, P)
    count = hi - lo
    lay = RankLayout(n, P, RANK_LAYOUT)
    use_p2p = False
    if exchange in ("auto", "p2p") and cx.device.type == "cuda" and n > 0:
        try:
            rank_local, peer_ptrs = _peer_ranks(cx, lay.local_len)
            use_p2p = True
        except Exception as e:  # noqa: BLE001 -- any failure to map peers falls back to the collective path
            if exchange == "p2p":
                raise
            use_p2p = False
    flags = torch.tensor([1 if use_p2p else 0], dtype=torch.int64, device=cx.device)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN, group=cx.group)
    use_p2p = bool(int(flags.item()))
    peer_arg = peer_ptrs.ctypes.data_as(C.c_void_p) if use_p2p else None
    cx.call("sab200_dist_begin", cx.dev)
    cx.mark("start")
    d_text = torch.as_tensor(shard, dtype=torch.uint8).to(cx.device)
    need = min(count + HALO, n - lo)
    if d_text.numel() < need:
        raise ValueError("shard too short: %d bytes, need %d (own positions + halo)" % (d_text.numel(), need))
    cx.mark("upload")
    # 1. common alphabet / key shape
    d_hist = torch.zeros(256, dtype=torch.int64, device=cx.device)
    cx.call("sab200_dist_hist", _p(d_text), count, _p(d_hist), cx.dev)
    dist.all_reduce(d_hist, group=cx.group)
    hist = d_hist.cpu().numpy().astype(np.uint64)
    lut = np.zeros(256, dtype=np.uint16)
    b, k = C.c_int32(), C.c_int32()
    cx.call("sab200_dist_plan", hist.ctypes.data_as(C.c_void_p), n, lut.ctypes.data_as(C.c_void_p), C.byref(b), C.byref(k))
    b, k = b.value, k.value
    key_bits = max(1, (b ** k - 1).bit_length())  # keys are mixed-radix numbers in base b = sigma + 1
    cx.mark("alphabet")
    # 2. keys of own positions, splitters from a sample
    keys = cx.empty(count, torch.int64)
    idx = cx.empty(count, torch.int32)
    cx.call("sab200_dist_pack", _p(d_text), lo, count, n, lut.ctypes.data_as(C.c_void_p), b, k, _p(keys), _p(idx), cx.dev)
    S = 2048
    sample = torch.zeros(S + 1, dtype=torch.int64, device=cx.device)
    if count:
        step = max(1, count // S)
        s = keys[::step][:S]
        sample[:s.numel()] = s
        sample[S] = s.numel()
    gathered = [torch.empty_like(sample) for _ in range(P)]
    dist.all_gather(gathered, sample, group=cx.group)
    pool = np.concatenate([g.cpu().numpy()[:int(g[S])] for g in gathered]).view(np.uint64)
    pool.sort()
    splitters = np.zeros(max(P - 1, 1), dtype=np.uint64)
    for i in range(P - 1):
        splitters[i] = pool[min(pool.size - 1, (i + 1) * pool.size // P)] if pool.size else 0
    cx.mark("pack+splitters")
    # 3. partition by destination, exchange
    k0 = v0 = None
    if use_p2p:
        # fused: the partition kernel stores each record straight into its destination GPU
        cnt = np.zeros(P, dtype=np.uint64)
        cx.call("sab200_dist_count_keys", _p(keys), count, splitters.ctypes.data_as(C.c_void_p), P - 1,
                cnt.ctypes.data_as(C.c_void_p), cx.dev)
        mine = torch.tensor([int(x) for x in cnt], dtype=torch.int64, device=cx.device)
        allc = [torch.empty_like(mine) for _ in range(P)]
        dist.all_gather(allc, mine, group=cx.group)
        cx.collectives += 1
        mat = np.stack([a.cpu().numpy() for a in allc])          # mat[src][dst]
        recv_tot = mat.sum(axis=0)
        cap = int(1.25 * B) + 4096
        # the symmetric receive buffers stay allocated between calls: only use them when they are a small
        # part of the device memory (3.9 GiB on 2 GPUs needs every byte for the rounds)
        roomy = cap * 12 <= 0.12 * torch.cuda.get_device_properties(cx.device).total_memory
        if roomy and int(recv_tot.max()) <= cap:
            rk, ri, pk, pi = _peer_recv(cx, cap)
            offs = np.ascontiguousarray(mat[:rank].sum(axis=0) if rank else np.zeros(P, dtype=np.int64)).astype(np.uint64)
            _barrier(cx)  # every rank is done with the previous contents of its receive buffers
            cx.mark("partition_keys")
            cx.call("sab200_dist_partition_keys_p2p", _p(keys), _p(idx), count, splitters.ctypes.data_as(C.c_void_p), P - 1,
                    offs.ctypes.data_as(C.c_void_p), pk.ctypes.data_as(C.c_void_p), pi.ctypes.data_as(C.c_void_p), cx.dev)
            _barrier(cx)  # all peers have finished storing into this rank's buffers
            Rn = int(recv_tot[rank])
            k0, v0 = rk[:Rn], ri[:Rn]
            cx.a2a_bytes += count * 12
    if k0 is None:
        kp = cx.empty(count, torch.int64)
        ip = cx.empty(count, torch.int32)
        cnt = np.zeros(P, dtype=np.uint64)
        cx.call("sab200_dist_partition_keys", _p(keys), _p(idx), count, splitters.ctypes.data_as(C.c_void_p), P - 1, _p(kp), _p(ip),
                cnt.ctypes.data_as(C.c_void_p), cx.dev)
        cx.mark("partition_keys")
        send = [int(x) for x in cnt]
        recv = cx.exchange_counts(send)
        k0 = cx.all_to_all(kp, send, recv)
        v0 = cx.all_to_all(ip, send, recv)
        del kp, ip
    del keys, idx
    R = k0.numel()
    cx.mark("exchange_keys")
    # 4. local sort: this rank's slice of the suffix array
    k1 = cx.empty(R, torch.int64)
    v1 = cx.empty(R, torch.int32)
    which = cx.call("sab200_dist_sort_pairs", _p(k0), _p(k1), _p(v0), _p(v1), R, key_bits, cx.dev)
    ks, vs = (k0, v0) if which == 0 else (k1, v1)
    cx.mark("local_sort")
    sizes = torch.zeros(P, dtype=torch.int64, device=cx.device)
    sizes[rank] = R
    dist.all_reduce(sizes, group=cx.group)
    sizes = [int(x) for x in sizes.tolist()]
    sa_off = 1 + sum(sizes[:rank])
    sa_local = cx.empty(R, torch.int32)
    rank_seq = cx.empty(R, torch.int32)
    act_r1 = cx.empty(R, torch.int32)
    act_idx = cx.empty(R, torch.int32)
    m = C.c_uint64()
    cx.call("sab200_dist_init_ranks", _p(ks), _p(vs), R, sa_off, _p(sa_local), _p(rank_seq), _p(act_r1), _p(act_idx),
                                      C.byref(m), cx.dev)
    m = m.value
    cx.mark("init_ranks")
    # 5. ranks travel to the owners of their text positions: all of them, or (lazy) only the active ones
    lazy = False
    if LAZY_ISA and lay.kind == "block" and n > 0:
        tot0 = torch.tensor([m], dtype=torch.int64, device=cx.device)
        dist.all_reduce(tot0, group=cx.group)
        cx.collectives += 1
        lazy = int(tot0.item()) <= LAZY_MAX_ACTIVE * n
    src_idx, src_rank, src_cnt = (act_idx, act_r1, m) if lazy else (vs, rank_seq, R)
    if use_p2p:
        if lazy:
            rank_local.fill_(-1)  # EMPTY
        else:
            rank_local.zero_()
        _barrier(cx)        # nobody stores into a block that is still being cleared
        if B <= P2P_MAX_RECORDS:
            cx.call("sab200_dist_scatter_p2p", _p(src_idx), _p(src_rank), src_cnt, lay.width, P, lay.shift, peer_arg, cx.dev)
        else:
            _send_ranks(cx, src_idx, src_rank, src_cnt, lay, lo, rank_local)
    else:
        rank_local = torch.full((lay.local_len,), -1 if lazy else 0, dtype=torch.int32, device=cx.device)
        _send_ranks(cx, src_idx, src_rank, src_cnt, lay, lo, rank_local)
    if lazy:
        # the empty suffix (position n) has rank 0; every other EMPTY slot is resolved on demand
        o_n = min(n // lay.width, P - 1)
        if rank == o_n:
            rank_local[n - o_n * lay.width] = 0
        if use_p2p:
            _barrier(cx)
        sorted_keys = ks  # this rank's slice of the sorted keys answers the key look-ups of the rounds
    del src_idx, src_rank, ks, vs, k0, k1, v0, v1, rank_seq
    cx.mark("ranks_to_owners")
    # 6. doubling rounds
    rank_bits = max(1, int(n + 1).bit_length())
    cur_r1, cur_idx, m, rebalanced = _rebalance(cx, act_r1, act_idx, m)
    del act_r1, act_idx  # views (not moved) keep the storage alive; moved lists replace it
    starts = np.array([1 + sum(sizes[:g]) for g in range(P)], dtype=np.uint32)  # first SA position of every slice
    cx.mark("rebalance")
    h = k
    rounds = 0
    active = []
    while True:
        tot = torch.tensor([m], dtype=torch.int64, device=cx.device)
        dist.all_reduce(tot, group=cx.group)
        tot = int(tot.item())
        active.append(tot)
        if tot == 0:
            break
        rounds += 1
        if h > n or rounds > 64:
            raise RuntimeError("prefix doubling did not converge")
        cx.trace("rounds/count")
        key64 = cx.empty(m, torch.int64)
        p2p_round = use_p2p and tot <= P2P_MAX_RECORDS * P and not lazy  # same decision on every rank
        if p2p_round:
            # the all_reduce above ordered every rank's previous stores before these loads
            cx.call("sab200_dist_gather_p2p", _p(cur_r1), _p(cur_idx), m, h, lay.width, P, lay.shift, peer_arg, _p(key64), cx.dev)
            ipart = cur_idx
        else:
            # requests i+h to the owners, answers back in the same order
            ipart, rpart, send = _to_owner(cx, cur_idx, cur_r1, m, h, lay)
            cx.trace("rounds/partition_requests")
            recv = cx.exchange_counts(send)
            q = cx.all_to_all(ipart, send, recv)
            cx.trace("rounds/send_requests")
            ans = cx.empty(q.numel(), torch.int32)
            cx.call("sab200_dist_gather", _p(q), q.numel(), h, lo, lay.width, P, lay.shift, _p(rank_local), _p(ans), cx.dev)
            cx.trace("rounds/gather")
            if lazy:
                _resolve_empty(cx, q, ans, h, lo, d_text, n, lut, b, k, splitters, sorted_keys, R, sa_off, rank_local)
                cx.trace("rounds/resolve_empty")
            r2 = cx.all_to_all(ans, recv, send)
            cx.trace("rounds/send_answers")
            cx.call("sab200_dist_make_keys", _p(rpart), _p(r2), m, _p(key64), cx.dev)
        key_tmp = cx.empty(m, torch.int64)
        idx_tmp = cx.empty(m, torch.int32)
        which = cx.call("sab200_dist_sort_pairs", _p(key64), _p(key_tmp), _p(ipart), _p(idx_tmp), m, 32 + rank_bits, cx.dev)
        sk, si = (key64, ipart) if which == 0 else (key_tmp, idx_tmp)
        cx.trace("rounds/sort")
        out_r1 = cx.empty(m, torch.int32)
        out_idx = cx.empty(m, torch.int32)
        upd_idx = cx.empty(m, torch.int32)
        upd_r = cx.empty(m, torch.int32)
        kept = C.c_uint64()
        set_pos = cx.empty(m, torch.int32) if rebalanced else None
        cx.call("sab200_dist_rerank", _p(sk), _p(si), m, sa_off, _p(sa_local), _p(out_r1), _p(out_idx), _p(upd_idx), _p(upd_r),
                _p(set_pos) if rebalanced else None, C.byref(kept), cx.dev)
        cx.trace("rounds/rerank")
        # free what the round no longer needs before the exchanges allocate their buffers (the 3.9 GiB text on
        # 2 GPUs runs within a few GB of the device memory)
        del sk, key64, key_tmp
        if not p2p_round:
            del q, ans, r2, rpart
        if rebalanced:
            _send_sa(cx, set_pos, si, m, starts, sa_off, sa_local)
            cx.trace("rounds/route_sa")
        del set_pos, si, ipart, idx_tmp, cur_r1, cur_idx
        if p2p_round:
            _barrier(cx)  # every rank has finished loading ranks of this round
            cx.call("sab200_dist_scatter_p2p", _p(upd_idx), _p(upd_r), m, lay.width, P, lay.shift, peer_arg, cx.dev)
        else:
            _send_ranks(cx, upd_idx, upd_r, m, lay, lo, rank_local)
        del upd_idx, upd_r
        cx.trace("rounds/update_ranks")
        m = kept.value
        cur_r1, cur_idx = out_r1[:m], out_idx[:m]
        h *= 2
    cx.mark("rounds/count" if cx.tracing else "rounds")
    cx.call("sab200_dist_end", cx.dev)
    if stats is not None:
        stats.update({"rounds": rounds, "active": active, "slice": R, "sa_off": sa_off, "symbols_per_key": k,
                      "bits_per_symbol": b, "all_to_all_bytes": cx.a2a_bytes, "collectives": cx.collectives,
                      "exchange": "p2p" if use_p2p else "collective", "rebalanced": rebalanced, "rank_layout": lay.kind,
                      "lazy_isa": lazy, "resolved_empty": cx.resolved_empty,
                      "phase_ms": {k_: round(v_, 2) for k_, v_ in cx.phase_ms.items()},
                      "wall_ms": round((time.perf_counter() - t_enter) * 1e3, 2)})
    return sa_local, sa_off


def gather_sa(sa_local, n, group=None):
    """Assembles the full suffix array (n+1 entries incl. the sentinel) on every rank (tests / small n)."""
    P = dist.get_world_size(group)
    size = torch.tensor([sa_local.numel()], dtype=torch.int64, device=sa_local.device)
    sizes = [torch.empty_like(size) for _ in range(P)]
    dist.all_gather(sizes, size, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes + [1])
    pad = torch.zeros(mx, dtype=torch.int32, device=sa_local.device)
    pad[:sa_local.numel()] = sa_local
    parts = [torch.empty_like(pad) for _ in range(P)]
    dist.all_gather(parts, pad, group=group)
    out = np.empty(n + 1, dtype=np.uint32)
    out[0] = n
    pos = 1
    for p, s in zip(parts, sizes):
        out[pos:pos + s] = p[:s].cpu().numpy().view(np.uint32)
        pos += s
    assert pos == n + 1
    return out
