// NCCL plumbing for the slab-sharded cycle (one process per GPU, NVLink 5 / NVSwitch).
//
// libnccl is resolved at run time with dlopen("libnccl.so.2"): inside a PyTorch process this
// binds to the copy torch already loaded (one NCCL per process), a plain C host gets the
// system library.  Replaces the MPI communicator handed to pmesh/PFFT (field.py:45-47) and the
// Alltoall(v)s inside Layout.exchange / PFFT's global transposes.
#include <dlfcn.h>
#include <nccl.h>

#include "ctx.cuh"

namespace hymd {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

struct Comm {
    ncclComm_t comm = nullptr;
    int P = 1, rank = 0;
};

static NcclApi g_nccl;

static int load_nccl() {
    if (g_nccl.handle) return HYMD_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        set_error("libnccl.so.2 could not be loaded: %s", dlerror());
        return HYMD_ERR_NCCL;
    }
#define HYMD_SYM(field, sym)                                             \
    *(void**)(&g_nccl.field) = dlsym(h, sym);                            \
    if (!g_nccl.field) {                                                 \
        set_error("libnccl is missing symbol %s", sym);                  \
        return HYMD_ERR_NCCL;                                            \
    }
    HYMD_SYM(GetUniqueId, "ncclGetUniqueId")
    HYMD_SYM(CommInitRank, "ncclCommInitRank")
    HYMD_SYM(CommDestroy, "ncclCommDestroy")
    HYMD_SYM(Send, "ncclSend")
    HYMD_SYM(Recv, "ncclRecv")
    HYMD_SYM(GroupStart, "ncclGroupStart")
    HYMD_SYM(GroupEnd, "ncclGroupEnd")
    HYMD_SYM(AllGather, "ncclAllGather")
    HYMD_SYM(GetErrorString, "ncclGetErrorString")
#undef HYMD_SYM
    g_nccl.handle = h;
    return HYMD_OK;
}

#define HYMD_NCCL(call)                                                                  \
    do {                                                                                 \
        ncclResult_t r_ = (call);                                                        \
        if (r_ != ncclSuccess) {                                                         \
            set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                      \
                      g_nccl.GetErrorString(r_));                                        \
            return HYMD_ERR_NCCL;                                                        \
        }                                                                                \
    } while (0)

int comm_unique_id(uint8_t* id) {
    HYMD_CHECK(load_nccl());
    static_assert(sizeof(ncclUniqueId) == HYMD_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    HYMD_NCCL(g_nccl.GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return HYMD_OK;
}

int comm_create(hymd_ctx* c, const uint8_t* id) {
    HYMD_CHECK(load_nccl());
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    Comm* cm = new Comm();
    cm->P = c->g.P;
    cm->rank = c->g.rank;
    ncclResult_t r = g_nccl.CommInitRank(&cm->comm, cm->P, u, cm->rank);
    if (r != ncclSuccess) {
        set_error("ncclCommInitRank(%d of %d) -> %s", cm->rank, cm->P, g_nccl.GetErrorString(r));
        delete cm;
        return HYMD_ERR_NCCL;
    }
    c->comm = cm;
    return HYMD_OK;
}

void comm_destroy(hymd_ctx* c) {
    if (!c->comm) return;
    if (c->comm->comm) g_nccl.CommDestroy(c->comm->comm);
    delete c->comm;
    c->comm = nullptr;
}

// Equal-size all-to-all: block q of `send` goes to rank q, block p of `recv` comes from rank p.
// The self block is NOT copied (callers write it in place).
int comm_alltoall(hymd_ctx* c, const void* send, void* recv, size_t bytes, cudaStream_t s) {
    Comm* cm = c->comm;
    if (!cm) { set_error("no communicator"); return HYMD_ERR_NCCL; }
    PhaseScope ps(c, HYMD_PHASE_ALLTOALL, s);
    HYMD_NCCL(g_nccl.GroupStart());
    for (int i = 1; i < cm->P; ++i) {
        const int to = (cm->rank + i) % cm->P, from = (cm->rank - i + cm->P) % cm->P;
        HYMD_NCCL(g_nccl.Send((const char*)send + (size_t)to * bytes, bytes, ncclInt8, to, cm->comm, s));
        HYMD_NCCL(g_nccl.Recv((char*)recv + (size_t)from * bytes, bytes, ncclInt8, from, cm->comm, s));
    }
    HYMD_NCCL(g_nccl.GroupEnd());
    c->launches += 1;
    return HYMD_OK;
}

int comm_alltoallv(hymd_ctx* c, const void* send, const size_t* send_off, const size_t* send_bytes,
                   void* recv, const size_t* recv_off, const size_t* recv_bytes, cudaStream_t s) {
    Comm* cm = c->comm;
    if (!cm) { set_error("no communicator"); return HYMD_ERR_NCCL; }
    HYMD_NCCL(g_nccl.GroupStart());
    for (int i = 1; i < cm->P; ++i) {
        const int to = (cm->rank + i) % cm->P, from = (cm->rank - i + cm->P) % cm->P;
        if (send_bytes[to])
            HYMD_NCCL(g_nccl.Send((const char*)send + send_off[to], send_bytes[to], ncclInt8, to, cm->comm, s));
        if (recv_bytes[from])
            HYMD_NCCL(g_nccl.Recv((char*)recv + recv_off[from], recv_bytes[from], ncclInt8, from, cm->comm, s));
    }
    HYMD_NCCL(g_nccl.GroupEnd());
    c->launches += 1;
    return HYMD_OK;
}

int comm_ring(hymd_ctx* c, int dir, void* const* sendp, void* const* recvp, int n, size_t bytes,
              cudaStream_t s) {
    Comm* cm = c->comm;
    if (!cm) { set_error("no communicator"); return HYMD_ERR_NCCL; }
    const int to = (cm->rank + dir + cm->P) % cm->P, from = (cm->rank - dir + cm->P) % cm->P;
    HYMD_NCCL(g_nccl.GroupStart());
    for (int i = 0; i < n; ++i) {
        HYMD_NCCL(g_nccl.Send(sendp[i], bytes, ncclInt8, to, cm->comm, s));
        HYMD_NCCL(g_nccl.Recv(recvp[i], bytes, ncclInt8, from, cm->comm, s));
    }
    HYMD_NCCL(g_nccl.GroupEnd());
    c->launches += 1;
    return HYMD_OK;
}

// Gathers `bytes` from every rank (device buffers) -- used for the migration counts.
int comm_allgather_host(hymd_ctx* c, const void* mine, void* all, size_t bytes, cudaStream_t s) {
    Comm* cm = c->comm;
    if (!cm) { set_error("no communicator"); return HYMD_ERR_NCCL; }
    HYMD_NCCL(g_nccl.AllGather(mine, all, bytes, ncclInt8, cm->comm, s));
    c->launches += 1;
    return HYMD_OK;
}

}  // namespace hymd
