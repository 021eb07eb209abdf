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
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

struct PeerMap {
    void* local;
    void* peer[HYMD_MAX_PEERS];
};

struct OpenedBlock {
    int rank;
    cudaIpcMemHandle_t handle;
    void* base;
};

constexpr size_t COMM_SCRATCH_BYTES = 512 + 128 * (HYMD_MAX_PEERS + 1);

struct Comm {
    ncclComm_t comm = nullptr;
    int P = 1, rank = 0;
    std::vector<PeerMap> maps;         // buffers of this rank and their addresses in every peer
    std::vector<OpenedBlock> opened;   // peer allocation blocks mapped through CUDA IPC
    void* d_scratch = nullptr;         // barrier word + handle exchange staging
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
    HYMD_SYM(AllReduce, "ncclAllReduce")
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
    for (auto& o : c->comm->opened) cudaIpcCloseMemHandle(o.base);
    if (c->comm->d_scratch) cudaFree(c->comm->d_scratch);
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

// ---- NVLink peer memory ------------------------------------------------------------------------
// Every rank exports `local` (a whole cudaMalloc allocation) through CUDA IPC and maps the
// corresponding buffer of every other rank, so kernels can store straight into a neighbour's HBM
// over NVLink.  Collective; cached per buffer.
struct IpcExport {
    cudaIpcMemHandle_t handle;       // of the allocation block that holds the buffer
    unsigned long long offset;       // of the buffer inside that block (small cudaMalloc
};                                   // allocations share one block)

int comm_peer_ptrs(hymd_ctx* c, void* local, void** peers, cudaStream_t s) {
    Comm* cm = c->comm;
    if (!cm) { set_error("no communicator"); return HYMD_ERR_NCCL; }
    if (cm->P > HYMD_MAX_PEERS) { set_error("more than %d slabs", HYMD_MAX_PEERS); return HYMD_ERR_INVALID; }
    for (auto& m : cm->maps)
        if (m.local == local) { memcpy(peers, m.peer, sizeof(void*) * cm->P); return HYMD_OK; }
    const size_t hb = sizeof(IpcExport);
    if (!cm->d_scratch) HYMD_CUDA(cudaMalloc(&cm->d_scratch, COMM_SCRATCH_BYTES));
    char* d_mine = (char*)cm->d_scratch + 256;
    char* d_all = d_mine + 256;
    // base of the allocation block (cuMemGetAddressRange through the runtime's driver entry point)
    typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
    static RangeFn range = nullptr;
    if (!range) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        HYMD_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !fn) { set_error("cuMemGetAddressRange unavailable"); return HYMD_ERR_CUDA; }
        range = (RangeFn)fn;
    }
    CUdeviceptr base = 0;
    size_t span = 0;
    if (range(&base, &span, (CUdeviceptr)local) != CUDA_SUCCESS) { set_error("cuMemGetAddressRange failed"); return HYMD_ERR_CUDA; }
    IpcExport mine;
    memset(&mine, 0, sizeof(mine));
    HYMD_CUDA(cudaIpcGetMemHandle(&mine.handle, (void*)base));
    mine.offset = (unsigned long long)((CUdeviceptr)local - base);
    HYMD_CUDA(cudaMemcpyAsync(d_mine, &mine, hb, cudaMemcpyHostToDevice, s));
    HYMD_NCCL(g_nccl.AllGather(d_mine, d_all, hb, ncclInt8, cm->comm, s));
    std::vector<IpcExport> all(cm->P);
    HYMD_CUDA(cudaMemcpyAsync(all.data(), d_all, hb * cm->P, cudaMemcpyDeviceToHost, s));
    HYMD_CUDA(cudaStreamSynchronize(s));
    PeerMap m;
    memset(&m, 0, sizeof(m));
    m.local = local;
    for (int q = 0; q < cm->P; ++q) {
        if (q == cm->rank) { m.peer[q] = local; continue; }
        void* opened = nullptr;
        for (auto& o : cm->opened)                      // a block can be opened only once per process
            if (o.rank == q && memcmp(&o.handle, &all[q].handle, sizeof(cudaIpcMemHandle_t)) == 0) opened = o.base;
        if (!opened) {
            cudaError_t e = cudaIpcOpenMemHandle(&opened, all[q].handle, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                set_error("cudaIpcOpenMemHandle(rank %d) -> %s (peer access between the GPUs of this "
                          "node is required)", q, cudaGetErrorString(e));
                return HYMD_ERR_CUDA;
            }
            cm->opened.push_back({q, all[q].handle, opened});
        }
        m.peer[q] = (char*)opened + all[q].offset;
    }
    cm->maps.push_back(m);
    memcpy(peers, m.peer, sizeof(void*) * cm->P);
    return HYMD_OK;
}

// All ranks have reached this point of the stream and their earlier kernels (including the stores
// they made into peer memory) are complete.
int comm_barrier(hymd_ctx* c, cudaStream_t s) {
    Comm* cm = c->comm;
    if (!cm) { set_error("no communicator"); return HYMD_ERR_NCCL; }
    if (!cm->d_scratch) HYMD_CUDA(cudaMalloc(&cm->d_scratch, COMM_SCRATCH_BYTES));
    HYMD_NCCL(g_nccl.AllReduce(cm->d_scratch, cm->d_scratch, 1, ncclInt32, ncclSum, cm->comm, s));
    c->launches += 1;
    c->peer_busy = 0;   // every consumer enqueued before this point has finished on every rank
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
