// sab_comm.cuh -- the communicator of the multi-GPU construction (include/sab200.h, "multi-GPU").
//
// One SabComm per rank (= per GPU).  Two transports behind the same four operations:
//   * NCCL (the product path): libnccl.so.2 is opened at run time (dlopen, so a single-GPU user needs no
//     NCCL installed, and a process that already carries torch's NCCL shares it).  Communicators come from
//     ncclCommInitRank (one process per GPU; the host language distributes the 128-byte unique id) or from
//     ncclCommInitAll (one process driving all GPUs: sab200_saca(..., ngpus > 1)).  Every collective is
//     enqueued on the library stream of the rank's device: nothing here synchronises the host.
//     all_to_all_v = grouped ncclSend / ncclRecv, the NCCL idiom for a variable-sized all-to-all.
//   * callbacks: the host language provides the collectives (tests run the whole distributed driver over
//     gloo with the SIMT-emulator build as the "device"; the buffers are then host memory).
#pragma once
#include "../../include/sab200.h"
#include "sab_context.cuh"
#include "sab_radix.cuh"

#ifndef SAB_EMU
#include <dlfcn.h>
#include <nccl.h>
#endif

struct sab200_comm {
    int rank = 0, P = 1, device = 0;
    int kind = 0;  // 0 = NCCL, 1 = callbacks
    void* nccl = nullptr;
    bool owns_nccl = true;
    sab200_comm_callbacks cb;
    // scratch for the small exchanges (counts, samples): device + pinned host, SAB_COMM_SCRATCH bytes each
    u64* d_small = nullptr;
    u64* h_small = nullptr;
    // accounting of the last construction
    u64 bytes_sent = 0;
    u32 collectives = 0;
    sab200_dist_stats last;
    // peers' device arenas mapped into this rank (fused key exchange): same process -> the pointer itself with peer
    // access enabled, another process -> cudaIpcOpenMemHandle.  Re-validated on every call (see sab_peers_update).
    struct Peer {
        u64 pid = 0, dev = 0, ptr = 0, bytes = 0;
        void* base = nullptr;
        bool ipc = false;
    } peers[SAB_MAX_RANKS];
    bool peers_ok = false;
    size_t last_want = 0;  // arena request of the previous construction on this communicator
};
#define SAB_COMM_SCRATCH ((size_t)1 << 20)

#ifndef SAB_EMU
// ---------------------------------------------------------------- NCCL through dlopen
struct SabNcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};
static SabNcclApi g_nccl;
static std::mutex g_nccl_mu;

static int sab_nccl_load() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.handle) return SAB_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        sab_set_error("multi-GPU construction needs NCCL: dlopen(libnccl.so.2) failed: %s", dlerror());
        return SAB_ERR_NCCL;
    }
#define SAB_NCCL_SYM(field, sym)                                                 \
    *(void**)(&g_nccl.field) = dlsym(h, sym);                                    \
    if (!g_nccl.field) {                                                         \
        sab_set_error("libnccl.so.2 lacks %s", sym);                             \
        return SAB_ERR_NCCL;                                                     \
    }
    SAB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    SAB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    SAB_NCCL_SYM(CommInitAll, "ncclCommInitAll")
    SAB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    SAB_NCCL_SYM(AllGather, "ncclAllGather")
    SAB_NCCL_SYM(AllReduce, "ncclAllReduce")
    SAB_NCCL_SYM(Send, "ncclSend")
    SAB_NCCL_SYM(Recv, "ncclRecv")
    SAB_NCCL_SYM(GroupStart, "ncclGroupStart")
    SAB_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    SAB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
    SAB_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef SAB_NCCL_SYM
    g_nccl.handle = h;
    return SAB_OK;
}
#define SAB_NCCL_TRY(expr)                                                                               \
    do {                                                                                                 \
        ncclResult_t _r = (expr);                                                                        \
        if (_r != ncclSuccess) {                                                                         \
            sab_set_error("%s:%d: %s -> NCCL: %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); \
            return SAB_ERR_NCCL;                                                                         \
        }                                                                                                \
    } while (0)
#endif  // !SAB_EMU

static int sab_comm_alloc_scratch(sab200_comm* cm) {
    SAB_CUDA_TRY(cudaSetDevice(cm->device));
    SAB_CUDA_TRY(cudaMalloc(&cm->d_small, SAB_COMM_SCRATCH));
    SAB_CUDA_TRY(cudaMallocHost(&cm->h_small, SAB_COMM_SCRATCH));
    return SAB_OK;
}

// Unmaps the peers' arenas (before any rank re-allocates its own: memory exported over CUDA IPC is only returned
// to the device once every importer has closed its mapping).
static void sab_comm_close_peers(sab200_comm* cm) {
#ifndef SAB_EMU
    for (int i = 0; i < SAB_MAX_RANKS; ++i) {
        if (cm->peers[i].ipc && cm->peers[i].base) cudaIpcCloseMemHandle(cm->peers[i].base);
        cm->peers[i] = sab200_comm::Peer();
    }
#endif
    cm->peers_ok = false;
}

static void sab_comm_free(sab200_comm* cm) {
    if (!cm) return;
    cudaSetDevice(cm->device);
#ifndef SAB_EMU
    if (cm->kind == 0 && cm->nccl && cm->owns_nccl && g_nccl.handle) g_nccl.CommDestroy((ncclComm_t)cm->nccl);
#endif
#ifndef SAB_EMU
    for (int i = 0; i < SAB_MAX_RANKS; ++i)
        if (cm->peers[i].ipc && cm->peers[i].base) cudaIpcCloseMemHandle(cm->peers[i].base);
#endif
    if (cm->d_small) cudaFree(cm->d_small);
    if (cm->h_small) cudaFreeHost(cm->h_small);
    delete cm;
}

// ---------------------------------------------------------------- the four operations (device buffers, stream-ordered)
// recv[r*bytes .. (r+1)*bytes) = rank r's send[0..bytes)
static int sab_comm_all_gather(sab200_comm* cm, cudaStream_t st, const void* send, void* recv, size_t bytes) {
    cm->collectives++;
    if (cm->P == 1) {
        if (send != recv) SAB_CUDA_TRY(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, st));
        return SAB_OK;
    }
    if (cm->kind == 1) {
        SAB_CUDA_TRY(cudaStreamSynchronize(st));
        if (cm->cb.all_gather(cm->cb.user, send, recv, bytes) != 0) {
            sab_set_error("all_gather callback failed");
            return SAB_ERR_NCCL;
        }
        return SAB_OK;
    }
#ifndef SAB_EMU
    SAB_NCCL_TRY(g_nccl.AllGather(send, recv, bytes, ncclUint8, (ncclComm_t)cm->nccl, st));
    return SAB_OK;
#else
    return SAB_ERR_NCCL;
#endif
}

static int sab_comm_all_reduce_u64(sab200_comm* cm, cudaStream_t st, u64* buf, size_t count) {
    cm->collectives++;
    if (cm->P == 1) return SAB_OK;
    if (cm->kind == 1) {
        SAB_CUDA_TRY(cudaStreamSynchronize(st));
        if (cm->cb.all_reduce_sum_u64(cm->cb.user, buf, count) != 0) {
            sab_set_error("all_reduce callback failed");
            return SAB_ERR_NCCL;
        }
        return SAB_OK;
    }
#ifndef SAB_EMU
    SAB_NCCL_TRY(g_nccl.AllReduce(buf, buf, count, ncclUint64, ncclSum, (ncclComm_t)cm->nccl, st));
    return SAB_OK;
#else
    return SAB_ERR_NCCL;
#endif
}

// Variable all-to-all in bytes: this rank sends send[soff[d] .. soff[d]+sbytes[d]) to rank d and receives
// rbytes[s] bytes from rank s at recv[roff[s]].  The part a rank sends to itself is a device copy.
static int sab_comm_all_to_all_v(sab200_comm* cm, cudaStream_t st, const void* send, const u64* sbytes, const u64* soff,
                                 void* recv, const u64* rbytes, const u64* roff) {
    cm->collectives++;
    const int P = cm->P, me = cm->rank;
    for (int d = 0; d < P; ++d)
        if (d != me) cm->bytes_sent += sbytes[d];
    if (cm->kind == 1 && P > 1) {
        SAB_CUDA_TRY(cudaStreamSynchronize(st));
        if (cm->cb.all_to_all_v(cm->cb.user, send, sbytes, soff, recv, rbytes, roff) != 0) {
            sab_set_error("all_to_all_v callback failed");
            return SAB_ERR_NCCL;
        }
        return SAB_OK;
    }
    if (sbytes[me])
        SAB_CUDA_TRY(cudaMemcpyAsync((char*)recv + roff[me], (const char*)send + soff[me], sbytes[me], cudaMemcpyDeviceToDevice, st));
    if (P == 1) return SAB_OK;
#ifndef SAB_EMU
    SAB_NCCL_TRY(g_nccl.GroupStart());
    for (int d = 0; d < P; ++d) {
        if (d == me) continue;
        if (sbytes[d]) SAB_NCCL_TRY(g_nccl.Send((const char*)send + soff[d], sbytes[d], ncclUint8, d, (ncclComm_t)cm->nccl, st));
        if (rbytes[d]) SAB_NCCL_TRY(g_nccl.Recv((char*)recv + roff[d], rbytes[d], ncclUint8, d, (ncclComm_t)cm->nccl, st));
    }
    SAB_NCCL_TRY(g_nccl.GroupEnd());
    return SAB_OK;
#else
    return SAB_ERR_NCCL;
#endif
}

// ---------------------------------------------------------------- host-visible small exchanges (one stream sync each)
// mat[s*P + d] = what rank s sends to rank d, on every rank
static int sab_comm_count_matrix(sab200_comm* cm, cudaStream_t st, const u64* mine, u64* mat) {
    const int P = cm->P;
    if (P == 1) {
        mat[0] = mine[0];
        return SAB_OK;
    }
    u64* h = cm->h_small;
    for (int d = 0; d < P; ++d) h[d] = mine[d];
    SAB_CUDA_TRY(cudaMemcpyAsync(cm->d_small, h, P * sizeof(u64), cudaMemcpyHostToDevice, st));
    SAB_TRY(sab_comm_all_gather(cm, st, cm->d_small, cm->d_small + P, P * sizeof(u64)));
    SAB_CUDA_TRY(cudaMemcpyAsync(h + P, cm->d_small + P, (size_t)P * P * sizeof(u64), cudaMemcpyDeviceToHost, st));
    SAB_CUDA_TRY(cudaStreamSynchronize(st));
    memcpy(mat, h + P, (size_t)P * P * sizeof(u64));
    return SAB_OK;
}

// Every rank contributes a row of W u64 values already in cm->d_small[0..W) (device); rows[s*W + i] = value i of
// rank s, on every rank's host.  One stream synchronisation.
static int sab_comm_gather_rows(sab200_comm* cm, cudaStream_t st, int W, u64* rows) {
    const int P = cm->P;
    u64* d_all = cm->d_small + 64;
    SAB_TRY(sab_comm_all_gather(cm, st, cm->d_small, d_all, (size_t)W * sizeof(u64)));
    SAB_CUDA_TRY(cudaMemcpyAsync(cm->h_small + 64, d_all, (size_t)P * W * sizeof(u64), cudaMemcpyDeviceToHost, st));
    SAB_CUDA_TRY(cudaStreamSynchronize(st));
    memcpy(rows, cm->h_small + 64, (size_t)P * W * sizeof(u64));
    return SAB_OK;
}

// vals[0..cnt) summed over the ranks, result on every rank's host
static int sab_comm_sum_u64(sab200_comm* cm, cudaStream_t st, u64* vals, int cnt) {
    if (cm->P == 1) return SAB_OK;
    u64* h = cm->h_small;
    for (int i = 0; i < cnt; ++i) h[i] = vals[i];
    SAB_CUDA_TRY(cudaMemcpyAsync(cm->d_small, h, cnt * sizeof(u64), cudaMemcpyHostToDevice, st));
    SAB_TRY(sab_comm_all_reduce_u64(cm, st, cm->d_small, (size_t)cnt));
    SAB_CUDA_TRY(cudaMemcpyAsync(h, cm->d_small, cnt * sizeof(u64), cudaMemcpyDeviceToHost, st));
    SAB_CUDA_TRY(cudaStreamSynchronize(st));
    for (int i = 0; i < cnt; ++i) vals[i] = h[i];
    return SAB_OK;
}

// Variable all-to-all of `elem`-byte records laid out destination-major in `send` (counts[d] records for rank d).
// Fills rcounts[] (records from every source) and returns the total received; `recv` must hold rtotal_cap records.
struct A2APlan {
    u64 scount[SAB_MAX_RANKS], rcount[SAB_MAX_RANKS];
    u64 stotal, rtotal;
};
static int sab_comm_plan(sab200_comm* cm, cudaStream_t st, const u64* send_counts, A2APlan* pl, u64* mat_out) {
    const int P = cm->P;
    u64 mat_[SAB_MAX_RANKS * SAB_MAX_RANKS];
    u64* mat = mat_out ? mat_out : mat_;
    SAB_TRY(sab_comm_count_matrix(cm, st, send_counts, mat));
    pl->stotal = pl->rtotal = 0;
    for (int d = 0; d < P; ++d) {
        pl->scount[d] = send_counts[d];
        pl->rcount[d] = mat[(size_t)d * P + cm->rank];
        pl->stotal += pl->scount[d];
        pl->rtotal += pl->rcount[d];
    }
    return SAB_OK;
}
// forward: records grouped by destination go out, records grouped by source come in
static int sab_comm_exchange(sab200_comm* cm, cudaStream_t st, const A2APlan& pl, const void* send, void* recv, size_t elem) {
    u64 sb[SAB_MAX_RANKS], so[SAB_MAX_RANKS], rb[SAB_MAX_RANKS], ro[SAB_MAX_RANKS];
    u64 a = 0, b = 0;
    for (int d = 0; d < cm->P; ++d) {
        sb[d] = pl.scount[d] * elem;
        so[d] = a;
        a += sb[d];
        rb[d] = pl.rcount[d] * elem;
        ro[d] = b;
        b += rb[d];
    }
    return sab_comm_all_to_all_v(cm, st, send, sb, so, recv, rb, ro);
}
// backward: the answers to what was received travel back along the same plan (same order)
static int sab_comm_exchange_back(sab200_comm* cm, cudaStream_t st, const A2APlan& pl, const void* send, void* recv, size_t elem) {
    u64 sb[SAB_MAX_RANKS], so[SAB_MAX_RANKS], rb[SAB_MAX_RANKS], ro[SAB_MAX_RANKS];
    u64 a = 0, b = 0;
    for (int d = 0; d < cm->P; ++d) {
        sb[d] = pl.rcount[d] * elem;
        so[d] = a;
        a += sb[d];
        rb[d] = pl.scount[d] * elem;
        ro[d] = b;
        b += rb[d];
    }
    return sab_comm_all_to_all_v(cm, st, send, sb, so, recv, rb, ro);
}
