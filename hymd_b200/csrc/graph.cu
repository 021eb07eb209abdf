// CUDA-graph replay of the per-step field update for launch-bound systems.
//
// What update_field (hymd/field.py:428-616) costs on a small system is not bandwidth but launches: C1
// (10 k particles, 24^3) and C2 (100 k, 64^3) run ~20 kernels / memsets of 2-15 us each per step, issued one
// by one from the host.  hymd_update_cycle() is hymd_sort_particles_ex + hymd_paint + hymd_field_cycle behind
// one entry point; when replay is enabled it records that sequence once per "shape" of the call into a CUDA
// graph (stream capture of the very same host code, so there is no second code path to keep in step) and
// afterwards submits the whole step with one cudaGraphLaunch.
//
// What makes a recorded step replayable:
//   * everything the step computes on the host and hands to kernels by value is a function of the call
//     arguments and of context state that the key covers: the particle count, the reuse flag, which half of
//     the record double buffer is current (it alternates every sort), and a digest of the configuration
//     (hymd_config incl. box and interaction matrix, geometry, the context's device pointers and buffer
//     sizes).  hymd_ctx_set_box / hymd_ctx_set_interaction / any re-allocation change the digest, so a
//     stale graph can never be hit;
//   * the positions pointer is NOT part of the key: the recorded kernels read a context-owned staging array,
//     which an ordinary device-to-device copy on the caller's stream fills right before the graph is
//     submitted.  MD loops that alternate between position buffers, and numpy callers whose upload lands at
//     a different address every step, replay the same two graphs (patching the source address of a copy
//     node inside the graph instead was tried first: cudaGraphExecMemcpyNodeSetParams1D rejects the node
//     of a captured copy -- gpurun_out/r4a);
//   * the host-side state the three entry points advance (record buffers swapped, "sorted", "have forces",
//     ... ) is re-applied from a snapshot taken when the graph was recorded.
// A key is run eagerly the first time it is seen (lazy allocations, cuFFT plans and kernel attributes are
// created then) and recorded the second time.  Several slabs, phase timing and compute_potential always run
// eagerly.
#include "ctx.cuh"

#include <stdlib.h>

namespace hymd {

// host state that hymd_sort_particles_ex / hymd_paint / hymd_field_cycle advance
struct StepState {
    int64_t np, order_n;
    bool sorted, has_charges;
    void* rec;
    void* rec_alt;
    bool have_lap, phi_is_filtered, have_phi_hat, have_phif, have_forces;
};

static StepState read_state(const hymd_ctx* c) {
    StepState s;
    s.np = c->np; s.order_n = c->order_n; s.sorted = c->sorted; s.has_charges = c->has_charges;
    s.rec = c->rec; s.rec_alt = c->rec_alt;
    s.have_lap = c->have_lap; s.phi_is_filtered = c->phi_is_filtered; s.have_phi_hat = c->have_phi_hat;
    s.have_phif = c->have_phif; s.have_forces = c->have_forces;
    return s;
}

static void write_state(hymd_ctx* c, const StepState& s) {
    c->np = s.np; c->order_n = s.order_n; c->sorted = s.sorted; c->has_charges = s.has_charges;
    c->rec = s.rec; c->rec_alt = s.rec_alt;
    c->have_lap = s.have_lap; c->phi_is_filtered = s.phi_is_filtered; c->have_phi_hat = s.have_phi_hat;
    c->have_phif = s.have_phif; c->have_forces = s.have_forces;
}

struct StepKey {
    const void* types;
    const void* charges;
    const void* rec;       // current half of the record double buffer before the call
    int64_t n;
    int reuse;
    uint64_t digest;
    bool operator==(const StepKey& o) const {
        return types == o.types && charges == o.charges && rec == o.rec && n == o.n && reuse == o.reuse &&
               digest == o.digest;
    }
};

struct StepGraph {
    StepKey key;
    int seen;                  // eager runs so far; < 0: recording failed once, never try again
    cudaGraphExec_t exec;      // nullptr until recorded
    StepState post;            // host state after the step
    int64_t launches;          // kernels in the graph (hymd_launch_count bookkeeping)
    uint64_t used;             // tick of the last use (eviction)
};

constexpr int GRAPH_SLOTS = 16;
constexpr int GRAPH_DEFAULT_MODE = -1;     // automatic; HYMD_B200_GRAPH = 0 | 1 | auto overrides

struct GraphCache {
    int mode;                  // 0 off, 1 on, -1 automatic (on for systems that are launch-bound)
    std::vector<StepGraph> slots;
    cudaStream_t cap;          // capture stream (the caller's may be the legacy stream, which cannot capture)
    void* stage;               // positions staging array
    size_t stage_bytes;
    uint64_t tick;
    int64_t replays, recorded, eager;
};

static GraphCache* cache(hymd_ctx* c) {
    if (!c->graphs) {
        GraphCache* g = new GraphCache();
        g->mode = GRAPH_DEFAULT_MODE;
        if (const char* e = getenv("HYMD_B200_GRAPH")) g->mode = strcmp(e, "auto") == 0 ? -1 : (atoi(e) != 0 ? 1 : 0);
        g->cap = nullptr; g->stage = nullptr; g->stage_bytes = 0;
        g->tick = 0; g->replays = g->recorded = g->eager = 0;
        c->graphs = g;
    }
    return c->graphs;
}

static void drop(StepGraph& e) {
    if (e.exec) cudaGraphExecDestroy(e.exec);
    e.exec = nullptr;
}

void graph_destroy(hymd_ctx* c) {
    GraphCache* g = c->graphs;
    if (!g) return;
    for (auto& e : g->slots) drop(e);
    if (g->cap) cudaStreamDestroy(g->cap);
    if (g->stage) cudaFree(g->stage);
    delete g;
    c->graphs = nullptr;
}

// FNV-1a over 8-byte words (the configuration block is 8.6 KB and is hashed on every call: bytewise that is
// ~10 us of host time, a fifth of a C1 cycle), trailing bytes one by one
static inline uint64_t fnv(uint64_t h, const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t w;
        memcpy(&w, b + i, 8);
        h = (h ^ w) * 1099511628211ull;
        h ^= h >> 29;
    }
    for (; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

// Everything besides the call arguments that the recorded launches depend on.
static uint64_t digest(const hymd_ctx* c, const GraphCache* g) {
    uint64_t h = 1469598103934665603ull;
    h = fnv(h, &c->cfg, sizeof(c->cfg));
    h = fnv(h, &c->g, sizeof(c->g));
    const int ints[] = {c->T, c->U, (int)c->f64, (int)c->fused, (int)c->plane, (int)c->grad2, (int)c->slab,
                        c->rtx, c->rty, c->rtz, c->rbz, c->rstages};
    h = fnv(h, ints, sizeof(ints));
    h = fnv(h, c->urow, sizeof(c->urow));
    // (rec / rec_alt enter as an unordered pair: which one is current is a key field of its own)
    const void* lo = c->rec < c->rec_alt ? c->rec : c->rec_alt;
    const void* hi = c->rec < c->rec_alt ? c->rec_alt : c->rec;
    const void* ptrs[] = {lo, hi, c->cell_start, c->q_sorted, c->scalars, c->scan_tmp, c->tab, c->xtw,
                          c->ytw, c->ztw, c->plane_scratch, c->Au, c->cu, c->d_urow, c->outscale, c->phi, c->phi_hat,
                          c->f_hat, c->gmesh, c->fft_work, c->wA, c->wS, g->stage};
    h = fnv(h, ptrs, sizeof(ptrs));
    const size_t sizes[] = {(size_t)c->cap, c->scan_tmp_bytes, c->plane_scratch_bytes, c->fft_work_bytes, c->wA_bytes,
                            c->wS_bytes, g->stage_bytes, c->plans ? c->plans->size() : 0};
    h = fnv(h, sizes, sizeof(sizes));
    return h;
}

static int eager_step(hymd_ctx* c, const void* d_pos, const int32_t* d_types, const void* d_charges, int64_t n,
                      int flags, int compute_potential, void* stream) {
    HYMD_CHECK(hymd_sort_particles_ex(c, d_pos, d_types, d_charges, n, flags, stream));
    HYMD_CHECK(hymd_paint(c, stream));
    return hymd_field_cycle(c, compute_potential, stream);
}

// Launch-bound systems: a step that moves less than ~100 MB is over in less time than its launches take to issue.
static bool small_system(const hymd_ctx* c, int64_t n) {
    return n <= (1 << 21) && c->g.real_elems <= (1ll << 21);
}

// Records one step into e (the host code runs for real, the device work does not).  On failure the host
// state is put back and the key is never recorded again.
static int record_step(hymd_ctx* c, GraphCache* g, StepGraph& e, const int32_t* d_types, const void* d_charges,
                       int64_t n, int flags) {
    const StepState pre = read_state(c);
    const int64_t launches0 = c->launches;
    cudaGraph_t graph = nullptr;
    bool ok = cudaStreamBeginCapture(g->cap, cudaStreamCaptureModeRelaxed) == cudaSuccess;
    if (ok) {
        const int st = eager_step(c, g->stage, d_types, d_charges, n, flags, 0, (void*)g->cap);
        const cudaError_t end = cudaStreamEndCapture(g->cap, &graph);
        ok = st == HYMD_OK && end == cudaSuccess && graph != nullptr;
    }
    if (ok) ok = cudaGraphInstantiate(&e.exec, graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
        cudaGetLastError();                 // a failed capture leaves a sticky-looking (non-fatal) error behind
        write_state(c, pre);
        c->launches = launches0;
        e.exec = nullptr;
        e.seen = -1;
        return HYMD_ERR_STATE;
    }
    e.post = read_state(c);
    e.launches = c->launches - launches0;
    write_state(c, pre);                    // nothing has run yet: the replay below advances the state
    c->launches = launches0;
    g->recorded++;
    return HYMD_OK;
}

static int replay_step(hymd_ctx* c, GraphCache* g, StepGraph& e, const void* d_pos, int64_t n, cudaStream_t s) {
    HYMD_CUDA(cudaMemcpyAsync(g->stage, d_pos, (size_t)n * 3 * c->rsz, cudaMemcpyDeviceToDevice, s));
    HYMD_CUDA(cudaGraphLaunch(e.exec, s));
    write_state(c, e.post);
    c->launches += e.launches;
    g->replays++;
    return HYMD_OK;
}

}  // namespace hymd

using namespace hymd;

extern "C" {

int hymd_ctx_set_graph(hymd_ctx* c, int mode) {
    if (!c) { set_error("null argument"); return HYMD_ERR_INVALID; }
    GraphCache* g = cache(c);
    g->mode = mode < 0 ? -1 : (mode ? 1 : 0);
    if (g->mode == 0) {
        cudaDeviceSynchronize();
        for (auto& e : g->slots) drop(e);
        g->slots.clear();
    }
    return HYMD_OK;
}

int hymd_ctx_graph_stats(hymd_ctx* c, int64_t out[4]) {
    if (!c || !out) { set_error("null argument"); return HYMD_ERR_INVALID; }
    GraphCache* g = c->graphs;
    out[0] = g ? g->replays : 0;
    out[1] = g ? g->recorded : 0;
    out[2] = g ? g->eager : 0;
    int64_t live = 0;
    if (g) for (auto& e : g->slots) live += e.exec != nullptr;
    out[3] = live;
    return HYMD_OK;
}

int hymd_update_cycle(hymd_ctx* c, const void* d_pos, const int32_t* d_types, const void* d_charges, int64_t n,
                      int flags, int compute_potential, void* stream) {
    if (!c) { set_error("null argument"); return HYMD_ERR_INVALID; }
    GraphCache* g = cache(c);
    const bool want = g->mode == 1 || (g->mode < 0 && small_system(c, n));
    if (!want || c->g.P != 1 || c->timing || compute_potential || n <= 0 || !d_pos) {
        g->eager++;
        return eager_step(c, d_pos, d_types, d_charges, n, flags, compute_potential, stream);
    }
    // the staging array follows the record capacity (its address is part of the digest)
    const size_t need = (size_t)(c->cap > n ? c->cap : n) * 3 * c->rsz;
    if (c->rec && g->stage_bytes < need) {
        if (g->stage) { HYMD_CUDA(cudaDeviceSynchronize()); cudaFree(g->stage); g->stage = nullptr; g->stage_bytes = 0; }
        HYMD_CUDA(cudaMalloc(&g->stage, need));
        g->stage_bytes = need;
    }
    if (!g->cap) HYMD_CUDA(cudaStreamCreateWithFlags(&g->cap, cudaStreamNonBlocking));
    const bool reuse = (flags & HYMD_SORT_REUSE_ORDER) && c->order_n == n;
    StepKey key = {reuse ? nullptr : (const void*)d_types, d_charges, c->rec, n, reuse ? 1 : 0, digest(c, g)};
    if (!reuse && !d_types) { set_error("null argument"); return HYMD_ERR_INVALID; }
    StepGraph* e = nullptr;
    for (auto& slot : g->slots)
        if (slot.key == key) { e = &slot; break; }
    if (!e) {
        if ((int)g->slots.size() >= GRAPH_SLOTS) {        // evict the least recently used slot
            size_t victim = 0;
            for (size_t i = 1; i < g->slots.size(); ++i)
                if (g->slots[i].used < g->slots[victim].used) victim = i;
            if (g->slots[victim].exec) HYMD_CUDA(cudaDeviceSynchronize());      // it may still be executing
            drop(g->slots[victim]);
            g->slots.erase(g->slots.begin() + victim);
        }
        StepGraph fresh;
        memset(&fresh, 0, sizeof(fresh));
        fresh.key = key;
        g->slots.push_back(fresh);
        e = &g->slots.back();
    }
    e->used = ++g->tick;
    // first sight of a key, a context that has not allocated its particle buffers yet, or a key whose
    // recording failed: the ordinary launches
    if (e->seen < 1 || !c->rec || !g->stage) {
        if (e->seen >= 0) e->seen++;
        g->eager++;
        return eager_step(c, d_pos, d_types, d_charges, n, flags, 0, stream);
    }
    if (!e->exec) {
        if (record_step(c, g, *e, reuse ? nullptr : d_types, d_charges, n, flags) != HYMD_OK) {
            g->eager++;
            return eager_step(c, d_pos, d_types, d_charges, n, flags, 0, stream);
        }
    }
    return replay_step(c, g, *e, d_pos, n, (cudaStream_t)stream);
}

}  // extern "C"
