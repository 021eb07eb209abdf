// Communication layer of the slab-sharded cycle (one rank per GPU, NVLink 5 / NVSwitch).
//
// Replaces the MPI communicator handed to pmesh/PFFT (field.py:45-47) and the Alltoall(v)s inside
// Layout.exchange / PFFT's global transposes.  Two transports behind one
// interface:
//
//   NCCL  one process per GPU (torchrun).  libnccl is resolved at run time with
//         dlopen("libnccl.so.2"): inside a PyTorch process this binds to the copy torch already
//         loaded.  NCCL carries only the rendezvous (exchange of CUDA IPC handles, migration counts
//         and the all-to-all-v of domain_decomposition); the per-cycle data plane is stores into peer
//         HBM (CUDA IPC mappings) and the per-cycle synchronisation is a FLAG BARRIER in peer memory:
//         one 32-thread kernel per barrier, thread q release-stores the epoch into rank q's flag
//         word and acquire-spins on the word rank q writes here -- no NCCL launch, no host
//         involvement (HYMD_B200_NCCL_BARRIER=1 selects the round-1 4-byte all-reduce instead).
//   LOCAL "virtual slabs": the P ranks are P host threads of ONE process (hymd_local_group_id), on
//         one GPU or several.  Peer pointers are plain device pointers, barriers are stream
//         synchronise + a host barrier.  The whole sharded pipeline -- transposes, halos, per-step
//         routing -- then runs and is tested on a single-GPU box.
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <condition_variable>
#include <map>
#include <mutex>

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

// ---- LOCAL transport: ranks = threads of this process ----------------------------------------------
constexpr size_t LOCAL_SLOT_BYTES = 512;
static const char LOCAL_MAGIC[8] = {'H', 'Y', 'M', 'D', 'L', 'O', 'C', 'L'};

struct LocalGroup {
    int P = 0, refs = 0;
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0;
    unsigned long long generation = 0;
    bool broken = false;
    unsigned char slot[HYMD_MAX_PEERS][LOCAL_SLOT_BYTES];

    // false on timeout (a rank died or raised): the group is then broken for everybody
    bool barrier() {
        static const int timeout_s = [] {
            const char* e = getenv("HYMD_B200_LOCAL_TIMEOUT_S");
            const int v = e ? atoi(e) : 0;
            return v > 0 ? v : 120;
        }();
        std::unique_lock<std::mutex> lk(m);
        if (broken) return false;
        const unsigned long long gen = generation;
        if (++arrived == P) {
            arrived = 0;
            ++generation;
            cv.notify_all();
            return true;
        }
        const bool ok = cv.wait_for(lk, std::chrono::seconds(timeout_s), [&] { return generation != gen || broken; });
        if (!ok || broken) { broken = true; cv.notify_all(); return false; }
        return true;
    }
    // every rank contributes `bytes` (<= LOCAL_SLOT_BYTES) and receives everybody's
    bool allgather(int rank, const void* mine, void* all, size_t bytes) {
        memcpy(slot[rank], mine, bytes);
        if (!barrier()) return false;
        for (int q = 0; q < P; ++q) memcpy((char*)all + q * bytes, slot[q], bytes);
        return barrier();
    }
};

static std::mutex g_groups_mutex;
static std::map<unsigned long long, LocalGroup*> g_groups;
static unsigned long long g_next_group = 1;

constexpr size_t CTRL_BYTES = 4096;
constexpr int CTRL_FLAGS = 0;          // uint32[HYMD_MAX_PEERS]: barrier epochs, word q written by rank q
constexpr int CTRL_PAYLOAD = 64;       // uint32[HYMD_MAX_PEERS]: word q = payload published by rank q

struct Comm {
    int P = 1, rank = 0;
    bool local = false;
    ncclComm_t comm = nullptr;
    LocalGroup* group = nullptr;
    unsigned long long group_key = 0;
    std::vector<PeerMap> maps;         // buffers of this rank and their addresses in every peer
    std::vector<OpenedBlock> opened;   // peer allocation blocks mapped through CUDA IPC
    void* d_scratch = nullptr;         // NCCL barrier word + handle exchange staging
    // flag barrier
    bool flags = false;
    unsigned char* ctrl = nullptr;     // CTRL_BYTES, peer-mapped
    unsigned char* peer_ctrl[HYMD_MAX_PEERS] = {};
    uint32_t epoch = 0;
    unsigned int* h_status = nullptr;  // mapped pinned host words the kernels raise (barrier timeout)
    unsigned int* d_status = nullptr;  // device alias
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

#define HYMD_LOCAL(call)                                                                 \
    do {                                                                                 \
        if (!(call)) {                                                                   \
            set_error("%s:%d: a rank of the in-process group did not arrive (it failed "  \
                      "or left the call sequence)", __FILE__, __LINE__);                 \
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

int comm_local_group_id(int world_size, uint8_t* id) {
    if (world_size < 1 || world_size > HYMD_MAX_PEERS) {
        set_error("in-process group of %d ranks (1..%d supported)", world_size, HYMD_MAX_PEERS);
        return HYMD_ERR_INVALID;
    }
    LocalGroup* g = new LocalGroup();
    g->P = world_size;
    std::lock_guard<std::mutex> lk(g_groups_mutex);
    const unsigned long long key = g_next_group++;
    g_groups[key] = g;
    memset(id, 0, HYMD_NCCL_UNIQUE_ID_BYTES);
    memcpy(id, LOCAL_MAGIC, 8);
    memcpy(id + 8, &key, sizeof(key));
    return HYMD_OK;
}

static int setup_control(hymd_ctx* c);

int comm_create(hymd_ctx* c, const uint8_t* id) {
    Comm* cm = new Comm();
    cm->P = c->g.P;
    cm->rank = c->g.rank;
    if (memcmp(id, LOCAL_MAGIC, 8) == 0) {
        unsigned long long key = 0;
        memcpy(&key, id + 8, sizeof(key));
        std::lock_guard<std::mutex> lk(g_groups_mutex);
        auto it = g_groups.find(key);
        if (it == g_groups.end() || it->second->P != cm->P) {
            set_error("in-process group %llu does not exist or has another size", key);
            delete cm;
            return HYMD_ERR_INVALID;
        }
        cm->local = true;
        cm->group = it->second;
        cm->group_key = key;
        cm->group->refs++;
    } else {
        int st = load_nccl();
        if (st != HYMD_OK) { delete cm; return st; }
        ncclUniqueId u;
        memcpy(&u, id, sizeof(u));
        ncclResult_t r = g_nccl.CommInitRank(&cm->comm, cm->P, u, cm->rank);
        if (r != ncclSuccess) {
            set_error("ncclCommInitRank(%d of %d) -> %s", cm->rank, cm->P, g_nccl.GetErrorString(r));
            delete cm;
            return HYMD_ERR_NCCL;
        }
    }
    c->comm = cm;
    return setup_control(c);
}

void comm_destroy(hymd_ctx* c) {
    Comm* cm = c->comm;
    if (!cm) return;
    for (auto& o : cm->opened) cudaIpcCloseMemHandle(o.base);
    if (cm->d_scratch) cudaFree(cm->d_scratch);
    if (cm->ctrl) cudaFree(cm->ctrl);
    if (cm->h_status) cudaFreeHost(cm->h_status);
    if (cm->comm) g_nccl.CommDestroy(cm->comm);
    if (cm->group) {
        std::lock_guard<std::mutex> lk(g_groups_mutex);
        if (--cm->group->refs == 0) {
            g_groups.erase(cm->group_key);
            delete cm->group;
        }
    }
    delete cm;
    c->comm = nullptr;
}

bool comm_is_local(const hymd_ctx* c) { return c->comm && c->comm->local; }

// Sticky device-raised conditions (mapped pinned memory, no copies, no events): bit 0 = a flag barrier
// timed out, bit 1 = guest capacity exceeded (route.cu).  Read by every entry point that could return
// results computed after the condition.
unsigned int* comm_status_device(hymd_ctx* c) { return c->comm ? c->comm->d_status : nullptr; }

int comm_check_status(hymd_ctx* c) {
    Comm* cm = c->comm;
    if (!cm || !cm->h_status) return HYMD_OK;
    const unsigned int st = *(volatile unsigned int*)cm->h_status;
    if (st & 1u) {
        set_error("rank %d: a peer-memory barrier timed out (another rank failed or left the call sequence)",
                  cm->rank);
        return HYMD_ERR_NCCL;
    }
    if (st & 2u) {
        set_error("rank %d: more particles outside their home slab than the guest buffers hold "
                  "(%u dropped in one step); call domain_decomposition more often or raise "
                  "HYMD_B200_GUEST_CAPACITY", cm->rank, ((volatile unsigned int*)cm->h_status)[1]);
        return HYMD_ERR_CAPACITY;
    }
    return HYMD_OK;
}

// Equal-size all-to-all: block q of `send` goes to rank q, block p of `recv` comes from rank p.
// The self block is NOT copied (callers write it in place).
int comm_alltoall(hymd_ctx* c, const void* send, void* recv, size_t bytes, cudaStream_t s) {
    Comm* cm = c->comm;
    if (!cm) { set_error("no communicator"); return HYMD_ERR_NCCL; }
    PhaseScope ps(c, HYMD_PHASE_ALLTOALL, s);
    size_t off[HYMD_MAX_PEERS], len[HYMD_MAX_PEERS];
    for (int q = 0; q < cm->P; ++q) { off[q] = (size_t)q * bytes; len[q] = q == cm->rank ? 0 : bytes; }
    return comm_alltoallv(c, send, off, len, recv, off, len, s);
}

struct LocalV {
    const void* base;
    size_t off[HYMD_MAX_PEERS], bytes[HYMD_MAX_PEERS];
};

int comm_alltoallv(hymd_ctx* c, const void* send, const size_t* send_off, const size_t* send_bytes,
                   void* recv, const size_t* recv_off, const size_t* recv_bytes, cudaStream_t s) {
    Comm* cm = c->comm;
    if (!cm) { set_error("no communicator"); return HYMD_ERR_NCCL; }
    if (cm->local) {
        static_assert(sizeof(LocalV) <= LOCAL_SLOT_BYTES, "slot size");
        LocalV mine, all[HYMD_MAX_PEERS];
        mine.base = send;
        for (int q = 0; q < cm->P; ++q) { mine.off[q] = send_off[q]; mine.bytes[q] = send_bytes[q]; }
        HYMD_CUDA(cudaStreamSynchronize(s));                    // my send buffer is complete
        HYMD_LOCAL(cm->group->allgather(cm->rank, &mine, all, sizeof(LocalV)));
        for (int q = 0; q < cm->P; ++q) {
            if (q == cm->rank || recv_bytes[q] == 0) continue;
            if (all[q].bytes[cm->rank] != recv_bytes[q]) { set_error("all-to-all-v size mismatch"); return HYMD_ERR_NCCL; }
            HYMD_CUDA(cudaMemcpyAsync((char*)recv + recv_off[q], (const char*)all[q].base + all[q].off[cm->rank],
                                      recv_bytes[q], cudaMemcpyDefault, s));
        }
        HYMD_CUDA(cudaStreamSynchronize(s));
        HYMD_LOCAL(cm->group->barrier());                        // senders may reuse their buffers
        c->launches += 1;
        return HYMD_OK;
    }
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
    if (cm->local) { set_error("the NCCL exchange path needs one process per GPU"); return HYMD_ERR_INVALID; }
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
// over NVLink.  Collective; cached per buffer.  LOCAL transport: the ranks share one address space.
struct IpcExport {
    cudaIpcMemHandle_t handle;       // of the allocation block that holds the buffer
    unsigned long long offset;       // of the buffer inside that block (small cudaMalloc
};                                   // allocations share one block)

struct LocalExport {
    void* ptr;
    int dev;
};

int comm_peer_ptrs(hymd_ctx* c, void* local, void** peers, cudaStream_t s) {
    Comm* cm = c->comm;
    if (!cm) { set_error("no communicator"); return HYMD_ERR_NCCL; }
    if (cm->P > HYMD_MAX_PEERS) { set_error("more than %d slabs", HYMD_MAX_PEERS); return HYMD_ERR_INVALID; }
    for (auto& m : cm->maps)
        if (m.local == local) { memcpy(peers, m.peer, sizeof(void*) * cm->P); return HYMD_OK; }
    PeerMap m;
    memset(&m, 0, sizeof(m));
    m.local = local;
    if (cm->local) {
        LocalExport mine = {local, c->dev}, all[HYMD_MAX_PEERS];
        HYMD_LOCAL(cm->group->allgather(cm->rank, &mine, all, sizeof(mine)));
        for (int q = 0; q < cm->P; ++q) {
            if (all[q].dev != c->dev) {
                cudaError_t e = cudaDeviceEnablePeerAccess(all[q].dev, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    set_error("cudaDeviceEnablePeerAccess(%d -> %d) -> %s", c->dev, all[q].dev, cudaGetErrorString(e));
                    return HYMD_ERR_CUDA;
                }
                cudaGetLastError();
            }
            m.peer[q] = all[q].ptr;
        }
        cm->maps.push_back(m);
        memcpy(peers, m.peer, sizeof(void*) * cm->P);
        return HYMD_OK;
    }
    const size_t hb = sizeof(IpcExport);
    if (!cm->d_scratch) HYMD_CUDA(cudaMalloc(&cm->d_scratch, 512 + 128 * (HYMD_MAX_PEERS + 1)));
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

// ---- barriers ------------------------------------------------------------------------------------
struct CtrlPtrs {
    unsigned char* p[HYMD_MAX_PEERS];
};

// Thread q of the single warp: (optionally) publishes payload[q] into rank q's payload word for this
// rank, release-stores the epoch into rank q's flag word for this rank, and acquire-spins until rank q
// has done the same here.  Everything this rank's earlier kernels stored into peer memory is ordered
// before the flag by the system-scope fence + release; everything the peers stored before their flag
// is visible to the kernels launched after this one.  A 20 s watchdog turns a lost rank into an error
// (status bit 0) instead of a hung GPU.
__global__ void __launch_bounds__(32) flag_barrier_kernel(CtrlPtrs peers, unsigned char* mine, int P, int rank,
                                                          uint32_t epoch, const uint32_t* __restrict__ payload,
                                                          unsigned int* status) {
    const int q = threadIdx.x;
    if (q >= P || q == rank) return;
    if (payload) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(peers.p[q] + CTRL_PAYLOAD) + rank;
        *dst = payload[q];
    }
    __threadfence_system();
    uint32_t* flag = reinterpret_cast<uint32_t*>(peers.p[q] + CTRL_FLAGS) + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
    const uint32_t* wait = reinterpret_cast<const uint32_t*>(mine + CTRL_FLAGS) + q;
    unsigned long long t0 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (unsigned long long it = 0;; ++it) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(wait) : "memory");
        if ((int32_t)(v - epoch) >= 0) break;
        if ((it & 1023) == 1023) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 20000000000ULL) {
                if (status) { atomicOr(status, 1u); __threadfence_system(); }
                break;
            }
        }
    }
}

// payload without the flag barrier (NCCL-barrier and LOCAL transports)
__global__ void __launch_bounds__(32) publish_kernel(CtrlPtrs peers, int P, int rank,
                                                     const uint32_t* __restrict__ payload) {
    const int q = threadIdx.x;
    if (q >= P || q == rank) return;
    uint32_t* dst = reinterpret_cast<uint32_t*>(peers.p[q] + CTRL_PAYLOAD) + rank;
    *dst = payload[q];
    __threadfence_system();
}

static int setup_control(hymd_ctx* c) {
    Comm* cm = c->comm;
    HYMD_CUDA(cudaMalloc((void**)&cm->ctrl, CTRL_BYTES));
    HYMD_CUDA(cudaMemset(cm->ctrl, 0, CTRL_BYTES));
    HYMD_CUDA(cudaHostAlloc((void**)&cm->h_status, 64, cudaHostAllocMapped));
    memset(cm->h_status, 0, 64);
    HYMD_CUDA(cudaHostGetDevicePointer((void**)&cm->d_status, cm->h_status, 0));
    HYMD_CUDA(cudaDeviceSynchronize());
    void* peers[HYMD_MAX_PEERS] = {};
    HYMD_CHECK(comm_peer_ptrs(c, cm->ctrl, peers, 0));
    for (int q = 0; q < cm->P; ++q) cm->peer_ctrl[q] = (unsigned char*)peers[q];
    // (in-process ranks sharing one device must not spin on each other: a device-wide synchronisation
    // on one rank's host thread would wait for the other rank's spinning kernel -- they use the host barrier)
    const char* nb = getenv("HYMD_B200_NCCL_BARRIER");
    cm->flags = !cm->local && !(nb && nb[0] == '1');
    return HYMD_OK;
}

// All ranks have reached this point of the stream and their earlier kernels (including the stores
// they made into peer memory) are complete.  payload (device, P words) != NULL: word q is delivered
// to rank q (readable there as comm_payload()[sender]) together with the barrier.
int comm_barrier_payload(hymd_ctx* c, const uint32_t* d_payload, cudaStream_t s) {
    Comm* cm = c->comm;
    if (!cm) { set_error("no communicator"); return HYMD_ERR_NCCL; }
    CtrlPtrs peers;
    for (int q = 0; q < HYMD_MAX_PEERS; ++q) peers.p[q] = cm->peer_ctrl[q];
    if (cm->flags) {
        ++cm->epoch;
        flag_barrier_kernel<<<1, 32, 0, s>>>(peers, cm->ctrl, cm->P, cm->rank, cm->epoch, d_payload, cm->d_status);
        HYMD_LAUNCH_CHECK(c);
        c->peer_busy = 0;
        return HYMD_OK;
    }
    if (d_payload) {
        publish_kernel<<<1, 32, 0, s>>>(peers, cm->P, cm->rank, d_payload);
        HYMD_LAUNCH_CHECK(c);
    }
    if (cm->local) {
        HYMD_CUDA(cudaStreamSynchronize(s));
        HYMD_LOCAL(cm->group->barrier());
    } else {
        if (!cm->d_scratch) HYMD_CUDA(cudaMalloc(&cm->d_scratch, 512 + 128 * (HYMD_MAX_PEERS + 1)));
        HYMD_NCCL(g_nccl.AllReduce(cm->d_scratch, cm->d_scratch, 1, ncclInt32, ncclSum, cm->comm, s));
    }
    c->launches += 1;
    c->peer_busy = 0;   // every consumer enqueued before this point has finished on every rank
    return HYMD_OK;
}

int comm_barrier(hymd_ctx* c, cudaStream_t s) { return comm_barrier_payload(c, nullptr, s); }

// this rank's payload words (device pointer): word q = what rank q published for this rank
const uint32_t* comm_payload(hymd_ctx* c) {
    return reinterpret_cast<const uint32_t*>(c->comm->ctrl + CTRL_PAYLOAD);
}

// Gathers `bytes` from every rank (device buffers) -- used for the migration counts.
int comm_allgather_host(hymd_ctx* c, const void* mine, void* all, size_t bytes, cudaStream_t s) {
    Comm* cm = c->comm;
    if (!cm) { set_error("no communicator"); return HYMD_ERR_NCCL; }
    if (cm->local) {
        if (bytes > LOCAL_SLOT_BYTES) { set_error("local all-gather of %zu bytes", bytes); return HYMD_ERR_INVALID; }
        unsigned char h_mine[LOCAL_SLOT_BYTES], h_all[HYMD_MAX_PEERS * LOCAL_SLOT_BYTES];
        HYMD_CUDA(cudaMemcpyAsync(h_mine, mine, bytes, cudaMemcpyDeviceToHost, s));
        HYMD_CUDA(cudaStreamSynchronize(s));
        HYMD_LOCAL(cm->group->allgather(cm->rank, h_mine, h_all, bytes));
        HYMD_CUDA(cudaMemcpyAsync(all, h_all, bytes * cm->P, cudaMemcpyHostToDevice, s));
        HYMD_CUDA(cudaStreamSynchronize(s));
        c->launches += 1;
        return HYMD_OK;
    }
    HYMD_NCCL(g_nccl.AllGather(mine, all, bytes, ncclInt8, cm->comm, s));
    c->launches += 1;
    return HYMD_OK;
}

}  // namespace hymd
