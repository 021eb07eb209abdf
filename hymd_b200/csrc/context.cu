// C ABI of libhymd_b200.so: context life cycle, cuFFT plans, and the per-step entry points
// (declared in include/hymd_b200.h, which cites the reference call sites each one replaces).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include <vector>

#include "ctx.cuh"

namespace hymd {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static int ceil_log2(long long v) {
    int b = 0;
    while ((1LL << b) < v) ++b;
    return b;
}

static int dev_alloc(void** p, size_t bytes) {
    if (*p != nullptr) return HYMD_OK;
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        return HYMD_ERR_NOMEM;
    }
    return HYMD_OK;
}

template <typename real>
static int upload(void* dst, const std::vector<double>& v) {
    std::vector<real> h(v.size());
    for (size_t i = 0; i < v.size(); ++i) h[i] = (real)v[i];
    HYMD_CUDA(cudaMemcpy(dst, h.data(), h.size() * sizeof(real), cudaMemcpyHostToDevice));
    HYMD_CUDA(cudaDeviceSynchronize());   // callers launch on non-blocking streams: the DMA must have landed
    return HYMD_OK;
}

static int upload_real(hymd_ctx* c, void* dst, const std::vector<double>& v) {
    return c->f64 ? upload<double>(dst, v) : upload<float>(dst, v);
}

// Gaussian factors and wave numbers (fftfreq convention), separable per axis.
static int build_tables(hymd_ctx* c) {
    const Geometry& g = c->g;
    const double sigma = c->cfg.sigma;
    std::vector<double> tab;
    const int n[3] = {g.Nx, g.Ny, g.Nz};
    const int len[3] = {g.Nx, g.Ny, g.Nzc};
    std::vector<double> k[3];
    for (int a = 0; a < 3; ++a) {
        k[a].resize(len[a]);
        for (int i = 0; i < len[a]; ++i) {
            const int ni = (i < (n[a] + 1) / 2) ? i : i - n[a];   // numpy.fft.fftfreq ordering
            k[a][i] = 2.0 * M_PI * ni / g.box[a];
        }
    }
    for (int a = 0; a < 3; ++a)
        for (int i = 0; i < len[a]; ++i) tab.push_back(exp(-0.5 * sigma * sigma * k[a][i] * k[a][i]));
    for (int a = 0; a < 3; ++a)
        for (int i = 0; i < len[a]; ++i) tab.push_back(k[a][i]);
    HYMD_CHECK(dev_alloc(&c->tab, tab.size() * c->rsz));
    HYMD_CHECK(upload_real(c, c->tab, tab));
    std::vector<double> tw(2 * (size_t)g.Nx);
    for (int j = 0; j < g.Nx; ++j) {
        tw[2 * j] = cos(2.0 * M_PI * j / g.Nx);
        tw[2 * j + 1] = -sin(2.0 * M_PI * j / g.Nx);
    }
    HYMD_CHECK(dev_alloc(&c->xtw, tw.size() * c->rsz));
    return upload_real(c, c->xtw, tw);
}

static int build_interaction(hymd_ctx* c) {
    const int T = c->T;
    // group types with identical rows of A (they share one potential / force mesh triple)
    c->U = 0;
    for (int t = 0; t < T; ++t) {
        int found = -1;
        for (int u = 0; u < c->U && found < 0; ++u) {
            bool same = c->cfg.c[t] == c->cfg.c[c->rowrep[u]];
            for (int j = 0; j < T && same; ++j)
                same = c->cfg.A[t * T + j] == c->cfg.A[c->rowrep[u] * T + j];
            if (same) found = u;
        }
        if (found < 0) { found = c->U; c->rowrep[c->U++] = t; }
        c->urow[t] = found;
    }
    const Geometry& g = c->g;
    const double m = (double)g.Nx * g.Ny * g.Nz;
    std::vector<double> Au((size_t)c->U * T + 1), cu(c->U);
    for (int u = 0; u < c->U; ++u) {
        for (int j = 0; j < T; ++j) Au[u * T + j] = c->cfg.A[c->rowrep[u] * T + j] / m;
        cu[u] = c->cfg.c[c->rowrep[u]];
    }
    Au[(size_t)c->U * T] = 1.0 / m;
    HYMD_CHECK(dev_alloc(&c->Au, (size_t)(HYMD_MAX_TYPES * HYMD_MAX_TYPES + 1) * c->rsz));
    HYMD_CHECK(dev_alloc(&c->cu, (size_t)HYMD_MAX_TYPES * c->rsz));
    HYMD_CHECK(dev_alloc((void**)&c->d_urow, HYMD_MAX_TYPES * sizeof(int)));
    HYMD_CHECK(dev_alloc(&c->outscale, (size_t)(HYMD_MAX_TYPES + 1) * c->rsz));
    HYMD_CHECK(upload_real(c, c->Au, Au));
    HYMD_CHECK(upload_real(c, c->cu, cu));
    HYMD_CUDA(cudaMemcpy(c->d_urow, c->urow, T * sizeof(int), cudaMemcpyHostToDevice));
    HYMD_CUDA(cudaDeviceSynchronize());
    const double dv = g.box[0] * g.box[1] * g.box[2] / m;
    std::vector<double> os(T + 1);
    for (int t = 0; t < T; ++t) os[t] = c->cfg.m[t] / dv;
    os[T] = 1.0 / dv;
    return upload_real(c, c->outscale, os);
}

static int ensure_particle_capacity(hymd_ctx* c, int64_t n) {
    if (n <= c->cap && c->rec) return HYMD_OK;      // (a rank without particles still bins its guests)
    int64_t cap = n + n / 8 + 1024;
    const int64_t guests = route_guest_rows(c);      // several slabs: room for the guests of every peer
    void* bufs[] = {c->rec, c->rec_alt, c->q_sorted};
    for (void* b : bufs)
        if (b) cudaFree(b);
    c->rec = c->rec_alt = c->q_sorted = nullptr;
    c->order_n = -1;
    HYMD_CHECK(dev_alloc(&c->rec, (size_t)(cap + guests) * (c->f64 ? sizeof(Rec64) : sizeof(Rec32))));
    HYMD_CHECK(dev_alloc(&c->rec_alt, (size_t)(cap + guests) * (c->f64 ? sizeof(Rec64) : sizeof(Rec32))));
    HYMD_CHECK(dev_alloc(&c->q_sorted, (size_t)(cap + guests) * c->rsz));
    c->cap = cap;
    return HYMD_OK;
}

}  // namespace hymd

using namespace hymd;

extern "C" {

const char* hymd_last_error(void) { return g_err; }
int hymd_abi_version(void) { return HYMD_B200_ABI_VERSION; }

int hymd_ctx_create(const hymd_config* cfg, const uint8_t* nccl_id, hymd_ctx** out) {
    if (!cfg || !out) { set_error("null argument"); return HYMD_ERR_INVALID; }
    if (cfg->struct_size != (int32_t)sizeof(hymd_config)) {
        set_error("hymd_config size mismatch: caller %d, library %zu", cfg->struct_size,
                  sizeof(hymd_config));
        return HYMD_ERR_INVALID;
    }
    if (cfg->n_types < 1 || cfg->n_types > HYMD_MAX_TYPES) {
        set_error("n_types = %d outside [1, %d]", cfg->n_types, HYMD_MAX_TYPES);
        return HYMD_ERR_INVALID;
    }
    for (int a = 0; a < 3; ++a)
        if (cfg->mesh[a] < 2 || !(cfg->box[a] > 0.0)) {
            set_error("invalid mesh/box on axis %d: %d, %g", a, cfg->mesh[a], cfg->box[a]);
            return HYMD_ERR_INVALID;
        }
    const int P = cfg->world_size;
    if (P < 1 || cfg->rank < 0 || cfg->rank >= P || cfg->mesh[0] % P || (P > 1 && cfg->mesh[1] % P)) {
        set_error("mesh (%d,%d) not divisible into %d slabs (rank %d)", cfg->mesh[0], cfg->mesh[1],
                  P, cfg->rank);
        return HYMD_ERR_INVALID;
    }
    if (P > 1 && !nccl_id) { set_error("nccl_id required when world_size > 1"); return HYMD_ERR_INVALID; }

    hymd_ctx* c = new hymd_ctx();
    memset(c, 0, sizeof(*c));
    c->cfg = *cfg;
    c->plans = new std::vector<PlanEntry>();
    c->ev_pool = new std::vector<cudaEvent_t>();
    c->ev_open = new std::vector<PhaseInterval>();
    c->f64 = cfg->dtype == HYMD_F64;
    c->order_n = -1;
    c->rsz = c->f64 ? 8 : 4;
    c->T = cfg->n_types;
    cudaGetDevice(&c->dev);
    Geometry& g = c->g;
    g.Nx = cfg->mesh[0]; g.Ny = cfg->mesh[1]; g.Nz = cfg->mesh[2];
    g.P = P; g.rank = cfg->rank;
    g.nxl = g.Nx / P; g.x0 = g.rank * g.nxl;
    g.nyl = g.Ny / P; g.y0 = g.rank * g.nyl;
    g.Nzc = g.Nz / 2 + 1;
    g.Nzcp = g.Nzc + (g.Nzc & 1);
    g.Nzp = (g.Nz + 1 + 3) / 4 * 4;
    const int bits = c->f64 ? 64 : 32;
    g.fbx = bits - ceil_log2(g.nxl > 2 ? g.nxl : 2);
    g.fby = bits - ceil_log2(g.Ny);
    g.fbz = bits - ceil_log2(g.Nz);
    if (c->f64) {
        if (g.fbx > 52) g.fbx = 52;
        if (g.fby > 52) g.fby = 52;
        if (g.fbz > 52) g.fbz = 52;
    }
    for (int a = 0; a < 3; ++a) g.box[a] = cfg->box[a];
    g.nbz = zbins_per_row(g.Nz);
    g.ncell = (long long)g.nxl * g.Ny * g.nbz;
    g.vx = P == 1 ? g.nxl : g.nxl + 1;
    g.real_elems = (long long)g.vx * g.Ny * g.Nz;
    g.ghost_elems = (long long)(g.nxl + 1) * (g.Ny + 1) * g.Nzp;
    g.k_elems = (long long)g.Nx * g.nyl * g.Nzcp;
    if (g.ncell + 1 >= (1LL << 31) || g.ghost_elems >= (1LL << 31) || g.k_elems >= (1LL << 31)) {
        set_error("local mesh too large for 32-bit cell keys / cuFFT int strides");
        hymd_ctx_destroy(c);
        return HYMD_ERR_INVALID;
    }

    int st = HYMD_OK;
    auto fail = [&](int code) { hymd_ctx_destroy(c); return code; };
    if ((st = dev_alloc((void**)&c->cell_start, (size_t)(g.ncell + 3) * 4))) return fail(st);
    if ((st = dev_alloc((void**)&c->scalars, sizeof(DeviceScalars)))) return fail(st);
    c->scan_tmp_bytes = scan_temp_bytes(g.ncell + 2);
    if ((st = dev_alloc(&c->scan_tmp, c->scan_tmp_bytes))) return fail(st);
    if ((st = build_tables(c))) return fail(st);
    if ((st = build_interaction(c))) return fail(st);
    const size_t rb = (size_t)g.real_elems * c->rsz, kb = (size_t)g.k_elems * 2 * c->rsz,
                 gb = (size_t)g.ghost_elems * c->rsz;
    if ((st = dev_alloc(&c->phi, c->T * rb))) return fail(st);
    if ((st = dev_alloc(&c->phi_hat, c->T * kb))) return fail(st);
    if ((st = dev_alloc(&c->f_hat, 3 * (size_t)c->T * kb))) return fail(st);   // sized for U == T
    if ((st = dev_alloc(&c->gmesh, 3 * (size_t)c->T * gb))) return fail(st);
    if (cudaMemset(c->gmesh, 0, 3 * (size_t)c->T * gb) != cudaSuccess) return fail(HYMD_ERR_CUDA);
    if (cfg->pme) {
        if ((st = dev_alloc(&c->phi_q, rb))) return fail(st);
        if ((st = dev_alloc(&c->phiq_hat, kb))) return fail(st);
        if ((st = dev_alloc(&c->e_hat, 3 * kb))) return fail(st);
        if ((st = dev_alloc(&c->emesh, 3 * gb))) return fail(st);
        if (cudaMemset(c->emesh, 0, 3 * gb) != cudaSuccess) return fail(HYMD_ERR_CUDA);
    }
    // slab pipeline: always with several GPUs; on one GPU only when asked for (testing)
    const char* force_slab = getenv("HYMD_B200_FORCE_SLAB");
    const char* no_fused = getenv("HYMD_B200_NO_FUSED");
    c->fused = xline_supported(c) && !(no_fused && no_fused[0] == '1');
    c->slab = P > 1 || c->fused || (force_slab && force_slab[0] == '1');
    c->plane = c->slab && plane_supported(c);
    const char* no_grad2 = getenv("HYMD_B200_GRAD2");
    c->grad2 = c->fused && c->plane && !(no_grad2 && no_grad2[0] == '0');
    if (P > 1 && (st = comm_create(c, nccl_id))) return fail(st);
    const char* nccl_x = getenv("HYMD_B200_NCCL_EXCHANGE");
    c->p2p = P > 1 && (!(nccl_x && nccl_x[0] == '1') || comm_is_local(c));
    // Transposes of the slab FFT over NVLink: "blocked" (default with the plane kernels:
    // the kernels read / write per-destination blocks, each block crosses as ONE contiguous peer copy),
    // "fused" (remote stores issued by the plane r2c / x-line kernels), "kernels" (pack / unpack push kernels)
    const char* xm = getenv("HYMD_B200_EXCHANGE");
    const bool pow2 = (g.nxl & (g.nxl - 1)) == 0 && (g.nyl & (g.nyl - 1)) == 0;
    c->xmode = 0;
    if (c->p2p && c->plane && pow2) {
        c->xmode = 2;
        if (xm && xm[0] == 'f' && c->fused) c->xmode = 1;
        if (xm && xm[0] == 'k') c->xmode = 0;
    }
    c->fused_push = c->xmode == 1;
    c->xcopy_kernel = !(xm && strcmp(xm, "blockedm") == 0);     // "blockedm": cudaMemcpyAsync (copy engines)
    // pipelined blocked exchange (slabfft.cu): HYMD_B200_XPIPE = 0 off, 1 when the pieces are large (default), 2 always
    c->xpipe = 0; c->xstream = nullptr; c->plane_sm_reserve = 0; c->sm_count = 0;
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->dev);
    if (c->xmode == 2 && c->xcopy_kernel) {
        const char* xp = getenv("HYMD_B200_XPIPE");
        c->xpipe = xp ? atoi(xp) : 1;
        if (c->xpipe) {
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);       // lo = the numerically greatest = lowest priority
            if (cudaStreamCreateWithPriority(&c->xstream, cudaStreamNonBlocking, lo) != cudaSuccess) return fail(HYMD_ERR_CUDA);
            for (auto& e : c->xev)
                if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return fail(HYMD_ERR_CUDA);
            const char* xc = getenv("HYMD_B200_XPIPE_COPY");
            c->xpipe_ce = !(xc && xc[0] == 'k');
            for (int q = 0; q < P; ++q) {
                if (cudaStreamCreateWithFlags(&c->xpeer[q], cudaStreamNonBlocking) != cudaSuccess) return fail(HYMD_ERR_CUDA);
                for (auto& e : c->xdone[q])
                    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return fail(HYMD_ERR_CUDA);
            }
        }
    }
    if ((st = readout_setup(c))) return fail(st);
    if (cudaDeviceSynchronize() != cudaSuccess) return fail(HYMD_ERR_CUDA);
    *out = c;
    return HYMD_OK;
}

int hymd_ctx_destroy(hymd_ctx* c) {
    if (!c) return HYMD_OK;
    cudaDeviceSynchronize();
    if (c->plans) { destroy_plans(c); delete c->plans; c->plans = nullptr; }
    migrate_destroy(c);
    route_destroy(c);
    comm_destroy(c);
    gpe_destroy(c);
    graph_destroy(c);
    void* bufs[] = {c->rec, c->rec_alt, c->cell_start, c->q_sorted,
                    c->scalars, c->scan_tmp, c->tab, c->xtw, c->Au, c->cu, c->d_urow, c->outscale, c->phi,
                    c->phi_hat, c->f_hat, c->gmesh, c->v_hat, c->phif_hat, c->tmp_hat, c->v_ext,
                    c->phi_q, c->phiq_hat, c->phiqf_hat, c->e_hat, c->psi_hat, c->emesh, c->psi, c->fft_work, c->lap_hat, c->lap,
                    c->wA, c->wS, c->halo, c->ytw, c->ztw, c->plane_scratch};
    for (void* b : bufs)
        if (b) cudaFree(b);
    if (c->xstream) {
        for (auto& e : c->xev)
            if (e) cudaEventDestroy(e);
        cudaStreamDestroy(c->xstream);
        for (int q = 0; q < HYMD_MAX_PEERS; ++q) {
            for (auto& e : c->xdone[q])
                if (e) cudaEventDestroy(e);
            if (c->xpeer[q]) cudaStreamDestroy(c->xpeer[q]);
        }
    }
    if (c->ev_open) {
        for (auto& iv : *c->ev_open) { cudaEventDestroy(iv.a); cudaEventDestroy(iv.b); }
        delete c->ev_open;
    }
    if (c->ev_pool) {
        for (cudaEvent_t e : *c->ev_pool) cudaEventDestroy(e);
        delete c->ev_pool;
    }
    delete c;
    return HYMD_OK;
}

int hymd_ctx_set_timing(hymd_ctx* c, int enable) {
    if (!c) { set_error("null argument"); return HYMD_ERR_INVALID; }
    c->timing = enable != 0;
    return HYMD_OK;
}

int hymd_ctx_get_timings(hymd_ctx* c, double* ms, int64_t* calls) {
    if (!c || !ms || !calls) { set_error("null argument"); return HYMD_ERR_INVALID; }
    HYMD_CUDA(cudaDeviceSynchronize());
    for (auto& iv : *c->ev_open) {
        float t = 0.f;
        HYMD_CUDA(cudaEventElapsedTime(&t, iv.a, iv.b));
        if (iv.phase >= 0 && iv.phase < HYMD_PHASE_COUNT) { ms[iv.phase] += t; calls[iv.phase]++; }
        c->ev_pool->push_back(iv.a);
        c->ev_pool->push_back(iv.b);
    }
    c->ev_open->clear();
    return HYMD_OK;
}

int hymd_ctx_set_box(hymd_ctx* c, const double box[3]) {
    if (!c || !box) { set_error("null argument"); return HYMD_ERR_INVALID; }
    for (int a = 0; a < 3; ++a) { c->cfg.box[a] = box[a]; c->g.box[a] = box[a]; }
    c->sorted = false;
    HYMD_CHECK(build_tables(c));
    return build_interaction(c);
}

int hymd_ctx_set_interaction(hymd_ctx* c, const double* A, const double* cc, const double* m,
                             double sigma, double elec_conversion) {
    if (!c) { set_error("null argument"); return HYMD_ERR_INVALID; }
    const int T = c->T;
    if (A) memcpy(c->cfg.A, A, sizeof(double) * T * T);
    if (cc) memcpy(c->cfg.c, cc, sizeof(double) * T);
    if (m) memcpy(c->cfg.m, m, sizeof(double) * T);
    c->cfg.sigma = sigma;
    c->cfg.elec_conversion = elec_conversion;
    const int oldU = c->U;
    HYMD_CHECK(build_tables(c));
    HYMD_CHECK(build_interaction(c));
    if (c->U != oldU) {
        // the number of distinct potential rows changed: the TMA box set changes (cuFFT plans
        // are cached per batch size)
        HYMD_CHECK(readout_setup(c));
    }
    return HYMD_OK;
}

int hymd_sort_particles(hymd_ctx* c, const void* d_pos, const int32_t* d_types,
                        const void* d_charges, int64_t n, void* stream) {
    return hymd_sort_particles_ex(c, d_pos, d_types, d_charges, n, 0, stream);
}

int hymd_ctx_reset_order(hymd_ctx* c) {
    if (!c) { set_error("null argument"); return HYMD_ERR_INVALID; }
    c->order_n = -1;
    return HYMD_OK;
}

int hymd_sort_particles_ex(hymd_ctx* c, const void* d_pos, const int32_t* d_types,
                           const void* d_charges, int64_t n, int flags, void* stream) {
    const bool reuse = c && (flags & HYMD_SORT_REUSE_ORDER) && c->order_n == n && (n > 0 || c->g.P > 1);
    if (!c || (n > 0 && (!d_pos || (!d_types && !reuse)))) { set_error("null argument"); return HYMD_ERR_INVALID; }
    const int64_t lim = c->f64 ? (1LL << 31) : (1LL << REC32_IDX_BITS);
    if (n < 0 || n >= lim) {
        set_error("n = %lld particles exceeds the per-GPU limit %lld", (long long)n, (long long)lim);
        return HYMD_ERR_CAPACITY;
    }
    cudaStream_t s = (cudaStream_t)stream;
    HYMD_CHECK(comm_check_status(c));
    HYMD_CHECK(route_prepare(c, n, s));      // several slabs, first call: guest buffers (collective)
    HYMD_CHECK(ensure_particle_capacity(c, n));
    c->np = n;
    c->has_charges = d_charges != nullptr;
    {
        PhaseScope ps(c, HYMD_PHASE_SORT, s);
        HYMD_CHECK(sort_particles(c, d_pos, d_types, d_charges, n, reuse, s));
    }
    c->sorted = true;
    c->order_n = n;
    return HYMD_OK;
}

int hymd_exchange_cost(hymd_ctx* c, int64_t* sent_to, void* stream) {
    if (!c || !sent_to) { set_error("null argument"); return HYMD_ERR_INVALID; }
    for (int q = 0; q < c->g.P; ++q) sent_to[q] = 0;
    const uint32_t* d = route_send_counts(c);
    if (c->g.P == 1 || d == nullptr || !c->sorted) return HYMD_OK;
    uint32_t h[HYMD_MAX_PEERS] = {};
    HYMD_CUDA(cudaMemcpyAsync(h, d, sizeof(uint32_t) * c->g.P, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    HYMD_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    for (int q = 0; q < c->g.P; ++q) sent_to[q] = q == c->g.rank ? 0 : (int64_t)h[q];
    return HYMD_OK;
}

int hymd_set_charges(hymd_ctx* c, const void* d_charges, void* stream) {
    if (!c || (!d_charges && c->np > 0)) { set_error("null argument"); return HYMD_ERR_INVALID; }
    if (!c->sorted) { set_error("hymd_set_charges before hymd_sort_particles"); return HYMD_ERR_STATE; }
    HYMD_CHECK(gather_charges(c, d_charges, (cudaStream_t)stream));
    c->has_charges = true;
    return HYMD_OK;
}

int hymd_paint(hymd_ctx* c, void* stream) {
    if (!c) { set_error("null argument"); return HYMD_ERR_INVALID; }
    if (!c->sorted) { set_error("hymd_paint before hymd_sort_particles"); return HYMD_ERR_STATE; }
    {
        PhaseScope ps(c, HYMD_PHASE_PAINT, (cudaStream_t)stream);
        HYMD_CHECK(paint_types(c, (cudaStream_t)stream));
        HYMD_CHECK(halo_reduce(c, c->phi, c->T, (cudaStream_t)stream));
    }
    c->phi_is_filtered = false;
    c->have_phi_hat = false;
    c->have_lap = false;
    return HYMD_OK;
}

// Force spectra buffer the k-space stage writes: with the fused kernel on one GPU that is the
// work layout of the batched 2-D c2r, otherwise f_hat in the k layout.
static void* force_spectra(hymd_ctx* c) { return (c->fused && c->g.P == 1) ? c->wA : c->f_hat; }

// for_forces: the force spectra are consumed by the inverse (y,z) transforms right after (field / PME
// cycle); with several slabs the x-line kernel then stores them straight into the peers' work buffers.
// By-product calls (hymd_materialize) only want the k-space outputs and keep everything local.
static int run_kspace(hymd_ctx* c, bool want_v, bool want_phif, cudaStream_t s, bool for_forces = false) {
    if (!c->fused) return kspace_forces(c, want_v, want_phif, s);
    HYMD_CHECK(ensure_work(c, 3 * c->U > c->T ? 3 * c->U : c->T));
    if (for_forces && c->g.P > 1 && c->fused_push && c->plane) {
        void* peers[HYMD_MAX_PEERS];
        HYMD_CHECK(push_work_begin(c, (c->grad2 ? 2 : 3) * c->U, peers, s));
        HYMD_CHECK(xline_forces(c, c->phi_hat, nullptr, want_v ? c->v_hat : nullptr,
                                want_phif ? c->phif_hat : nullptr, s, peers));
        return push_work_end(c, s);
    }
    return xline_forces(c, c->phi_hat, force_spectra(c), want_v ? c->v_hat : nullptr,
                        want_phif ? c->phif_hat : nullptr, s);
}

static int run_kspace_pme(hymd_ctx* c, bool want_psi, cudaStream_t s, bool for_forces = false) {
    if (!c->fused) return kspace_pme(c, want_psi, s);
    HYMD_CHECK(ensure_work(c, 3 * c->U > c->T ? 3 * c->U : c->T));
    if (for_forces && c->g.P > 1 && c->fused_push && c->plane) {
        void* peers[HYMD_MAX_PEERS];
        HYMD_CHECK(push_work_begin(c, c->grad2 ? 2 : 3, peers, s));
        HYMD_CHECK(xline_pme(c, c->phiq_hat, nullptr, want_psi ? c->psi_hat : nullptr,
                             want_psi ? c->phiqf_hat : nullptr, s, peers));
        return push_work_end(c, s);
    }
    return xline_pme(c, c->phiq_hat, (c->g.P == 1) ? c->wA : c->e_hat,
                     want_psi ? c->psi_hat : nullptr, want_psi ? c->phiqf_hat : nullptr, s);
}

static int materialize_impl(hymd_ctx* c, bool want_phi, bool want_v, cudaStream_t s) {
    const Geometry& g = c->g;
    const size_t kb = (size_t)g.k_elems * 2 * c->rsz, rb = (size_t)g.real_elems * c->rsz;
    if (want_v) {
        HYMD_CHECK(dev_alloc(&c->v_ext, c->T * rb));
        HYMD_CHECK(fft_inverse(c, c->v_hat, c->U, c->v_ext, false, s));   // consumes v_hat
    }
    if (want_phi) {
        HYMD_CHECK(dev_alloc(&c->tmp_hat, c->T * kb));
        HYMD_CUDA(cudaMemcpyAsync(c->tmp_hat, c->phif_hat, c->T * kb, cudaMemcpyDeviceToDevice, s));
        HYMD_CHECK(fft_inverse(c, c->tmp_hat, c->T, c->phi, false, s));
        c->phi_is_filtered = true;
    }
    return HYMD_OK;
}

int hymd_field_cycle(hymd_ctx* c, int compute_potential, void* stream) {
    if (!c) { set_error("null argument"); return HYMD_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    const Geometry& g = c->g;
    const size_t kb = (size_t)g.k_elems * 2 * c->rsz;
    if (c->phi_is_filtered) { set_error("hymd_field_cycle needs a fresh hymd_paint"); return HYMD_ERR_STATE; }
    if (c->fused) HYMD_CHECK(ensure_work(c, 3 * c->U > c->T ? 3 * c->U : c->T));
    {
        PhaseScope ps(c, HYMD_PHASE_FFT_FWD, s);
        if (c->fused) HYMD_CHECK(fft_forward_yz(c, c->phi, c->T, c->phi_hat, s));
        else HYMD_CHECK(fft_forward(c, c->phi, c->T, c->phi_hat, s));
    }
    c->have_phi_hat = true;
    const bool cp = compute_potential != 0;
    if (cp) {
        HYMD_CHECK(dev_alloc(&c->v_hat, c->T * kb));
        HYMD_CHECK(dev_alloc(&c->phif_hat, c->T * kb));
    }
    {
        PhaseScope ps(c, HYMD_PHASE_KSPACE, s);
        HYMD_CHECK(run_kspace(c, cp, cp, s, true));
    }
    {
        PhaseScope ps(c, HYMD_PHASE_FFT_INV, s);
        if (c->fused) HYMD_CHECK(fft_inverse_xdone(c, force_spectra(c), 3 * c->U, c->gmesh, true, s, c->grad2));
        else HYMD_CHECK(fft_inverse(c, c->f_hat, 3 * c->U, c->gmesh, true, s));
    }
    {
        PhaseScope ps(c, HYMD_PHASE_GHOST, s);
        if (!c->plane) HYMD_CHECK(fill_ghosts(c, c->gmesh, 3 * c->U, s));
        HYMD_CHECK(halo_fetch(c, c->gmesh, 3 * c->U, s));
    }
    c->have_forces = true;
    c->have_phif = cp;
    if (cp) {
        PhaseScope ps(c, HYMD_PHASE_BYPRODUCTS, s);
        HYMD_CHECK(materialize_impl(c, true, true, s));
    }
    return HYMD_OK;
}

int hymd_materialize(hymd_ctx* c, int want_phi, int want_phi_fourier, int want_v_ext,
                     int want_psi, void* stream) {
    if (!c) { set_error("null argument"); return HYMD_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t kb = (size_t)c->g.k_elems * 2 * c->rsz;
    if (want_psi && !c->have_psi) {
        if (!c->cfg.pme || !c->have_phiq_hat) {
            set_error("hymd_materialize(psi) before hymd_pme_cycle");
            return HYMD_ERR_STATE;
        }
        HYMD_CHECK(dev_alloc(&c->phiqf_hat, kb));
        HYMD_CHECK(dev_alloc(&c->psi_hat, kb));
        HYMD_CHECK(dev_alloc(&c->psi, (size_t)c->g.real_elems * c->rsz));
        HYMD_CHECK(run_kspace_pme(c, true, s));   // e_hat / the work area are scratch between cycles
        HYMD_CHECK(fft_inverse(c, c->psi_hat, 1, c->psi, false, s));
        c->have_psi = true;
    }
    if (!(want_phi || want_phi_fourier || want_v_ext)) return HYMD_OK;
    if (!c->have_phi_hat) { set_error("hymd_materialize before hymd_field_cycle"); return HYMD_ERR_STATE; }
    const bool need_phi = want_phi && !c->phi_is_filtered;
    const bool need_v = want_v_ext != 0;
    const bool need_pf = (want_phi_fourier || need_phi) && !c->have_phif;
    if (need_pf || need_v) {
        HYMD_CHECK(dev_alloc(&c->v_hat, c->T * kb));
        HYMD_CHECK(dev_alloc(&c->phif_hat, c->T * kb));
        HYMD_CHECK(run_kspace(c, true, true, s));   // f_hat / the work area are scratch between cycles
        c->have_phif = true;
    }
    return materialize_impl(c, need_phi, need_v, s);
}

int hymd_laplacian(hymd_ctx* c, void* stream) {
    if (!c) { set_error("null argument"); return HYMD_ERR_INVALID; }
    if (c->have_lap) return HYMD_OK;
    cudaStream_t s = (cudaStream_t)stream;
    // phi_fourier of the current spectra (field.py:577), then -k_d^2 and 3T inverse transforms
    HYMD_CHECK(hymd_materialize(c, 0, 1, 0, 0, stream));
    const size_t kb = (size_t)c->g.k_elems * 2 * c->rsz, rb = (size_t)c->g.real_elems * c->rsz;
    HYMD_CHECK(dev_alloc(&c->lap_hat, 3 * (size_t)c->T * kb));
    HYMD_CHECK(dev_alloc(&c->lap, 3 * (size_t)c->T * rb));
    PhaseScope ps(c, HYMD_PHASE_BYPRODUCTS, s);
    HYMD_CHECK(kspace_laplacian(c, s));
    HYMD_CHECK(fft_inverse(c, c->lap_hat, 3 * c->T, c->lap, false, s));   // consumes lap_hat
    c->have_lap = true;
    return HYMD_OK;
}

int hymd_field_pressure(hymd_ctx* c, const double* A, const double* cc, const double* type_charges,
                        double out[4], void* stream) {
    if (!c || !A || !cc || !out) { set_error("null argument"); return HYMD_ERR_INVALID; }
    if (!c->phi_is_filtered || !c->have_lap) {
        set_error("hymd_field_pressure needs the filtered densities (hymd_materialize) and hymd_laplacian");
        return HYMD_ERR_STATE;
    }
    if (type_charges && !(c->cfg.pme && c->have_psi)) {
        set_error("hymd_field_pressure with type charges needs psi (hymd_pme_cycle with want_psi)");
        return HYMD_ERR_STATE;
    }
    return field_pressure(c, A, cc, type_charges, out, (cudaStream_t)stream);
}

int hymd_readout(hymd_ctx* c, void* d_force, void* stream) {
    if (!c || (c->np > 0 && !d_force)) { set_error("null argument"); return HYMD_ERR_INVALID; }
    if (!c->sorted || !c->have_forces) {
        set_error("hymd_readout needs hymd_sort_particles and hymd_field_cycle first");
        return HYMD_ERR_STATE;
    }
    HYMD_CHECK(comm_check_status(c));
    if (c->np == 0 && c->g.P == 1) return HYMD_OK;   // several slabs: guests may still be read out here
    PhaseScope ps(c, HYMD_PHASE_READOUT, (cudaStream_t)stream);
    return readout_forces(c, d_force, (cudaStream_t)stream);
}

int hymd_pme_cycle(hymd_ctx* c, void* d_elec_force, int want_psi, void* stream) {
    if (!c) { set_error("null argument"); return HYMD_ERR_INVALID; }
    if (!c->cfg.pme) { set_error("context created without PME buffers"); return HYMD_ERR_STATE; }
    if (!c->sorted || !c->has_charges) {
        set_error("hymd_pme_cycle needs hymd_sort_particles with charges");
        return HYMD_ERR_STATE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const Geometry& g = c->g;
    const size_t kb = (size_t)g.k_elems * 2 * c->rsz, rb = (size_t)g.real_elems * c->rsz;
    {
        PhaseScope ps(c, HYMD_PHASE_PME_PAINT, s);
        HYMD_CHECK(paint_charges(c, s));
        HYMD_CHECK(halo_reduce(c, c->phi_q, 1, s));
    }
    {
        PhaseScope ps(c, HYMD_PHASE_PME_FFT, s);
        if (c->fused) HYMD_CHECK(fft_forward_yz(c, c->phi_q, 1, c->phiq_hat, s));
        else HYMD_CHECK(fft_forward(c, c->phi_q, 1, c->phiq_hat, s));
    }
    c->have_phiq_hat = true;
    if (want_psi) {
        HYMD_CHECK(dev_alloc(&c->phiqf_hat, kb));
        HYMD_CHECK(dev_alloc(&c->psi_hat, kb));
        HYMD_CHECK(dev_alloc(&c->psi, rb));
    }
    {
        PhaseScope ps(c, HYMD_PHASE_PME_KSPACE, s);
        HYMD_CHECK(run_kspace_pme(c, want_psi != 0, s, true));
    }
    {
        PhaseScope ps(c, HYMD_PHASE_PME_FFT, s);
        if (c->fused)
            HYMD_CHECK(fft_inverse_xdone(c, (c->g.P == 1) ? c->wA : c->e_hat, 3, c->emesh, true, s, c->grad2));
        else HYMD_CHECK(fft_inverse(c, c->e_hat, 3, c->emesh, true, s));
        if (!c->plane) HYMD_CHECK(fill_ghosts(c, c->emesh, 3, s));
        HYMD_CHECK(halo_fetch(c, c->emesh, 3, s));
        if (want_psi) HYMD_CHECK(fft_inverse(c, c->psi_hat, 1, c->psi, false, s));
    }
    c->have_psi = want_psi != 0;
    HYMD_CHECK(comm_check_status(c));
    // (several slabs: a rank without particles of its own still reads out its guests -- collective)
    if (d_elec_force || (c->g.P > 1 && c->np == 0)) {
        PhaseScope ps(c, HYMD_PHASE_PME_READOUT, s);
        HYMD_CHECK(readout_pme(c, d_elec_force, s));
    }
    return HYMD_OK;
}

int hymd_field_energy(hymd_ctx* c, const double* chi, double kappa, double rho0, double a,
                      double out[2], void* stream) {
    if (!c || !chi || !out) { set_error("null argument"); return HYMD_ERR_INVALID; }
    if (!c->phi_is_filtered) {
        set_error("hymd_field_energy needs the filtered densities (hymd_materialize / compute_potential)");
        return HYMD_ERR_STATE;
    }
    if (c->cfg.pme && c->has_charges && !c->have_psi) {
        set_error("hymd_field_energy needs psi (hymd_pme_cycle with want_psi)");
        return HYMD_ERR_STATE;
    }
    return field_energy(c, chi, kappa, rho0, a, out, (cudaStream_t)stream);
}

int hymd_get_field(hymd_ctx* c, int field_id, int t, int d, void** d_ptr, int64_t dims[3],
                   int64_t pitch[3]) {
    if (!c || !d_ptr || !dims || !pitch) { set_error("null argument"); return HYMD_ERR_INVALID; }
    const Geometry& g = c->g;
    const size_t rb = (size_t)g.real_elems * c->rsz, gb = (size_t)g.ghost_elems * c->rsz;
    auto real_geom = [&]() {
        dims[0] = g.nxl; dims[1] = g.Ny; dims[2] = g.Nz;
        pitch[0] = (int64_t)g.Ny * g.Nz; pitch[1] = g.Nz; pitch[2] = 1;
    };
    auto ghost_geom = [&]() {
        dims[0] = g.nxl; dims[1] = g.Ny; dims[2] = g.Nz;
        pitch[0] = (int64_t)(g.Ny + 1) * g.Nzp; pitch[1] = g.Nzp; pitch[2] = 1;
    };
    auto k_geom = [&](int F) {
        dims[0] = g.Nx; dims[1] = g.nyl; dims[2] = g.Nzc;
        pitch[0] = klayout(c, F).xs; pitch[1] = g.Nzcp; pitch[2] = 1;
    };
    const size_t csz = 2 * c->rsz;
    const bool t_ok = t >= 0 && t < c->T, d_ok = d >= 0 && d < 3;
    char* p = nullptr;
    switch (field_id) {
        case HYMD_FIELD_PHI: if (!t_ok) break; p = (char*)c->phi + t * rb; real_geom(); break;
        case HYMD_FIELD_PHI_FOURIER:
            if (!t_ok || !c->phif_hat) break;
            p = (char*)c->phif_hat + (size_t)t * klayout(c, c->T).fs * csz; k_geom(c->T); break;
        case HYMD_FIELD_FORCE_MESH:
            if (!t_ok || !d_ok) break; p = (char*)c->gmesh + (3 * c->urow[t] + d) * gb; ghost_geom(); break;
        case HYMD_FIELD_V_EXT:
            if (!t_ok || !c->v_ext) break; p = (char*)c->v_ext + c->urow[t] * rb; real_geom(); break;
        case HYMD_FIELD_PHI_Q: if (!c->phi_q) break; p = (char*)c->phi_q; real_geom(); break;
        case HYMD_FIELD_PHI_Q_FOURIER: if (!c->phiqf_hat) break; p = (char*)c->phiqf_hat; k_geom(1); break;
        case HYMD_FIELD_PSI: if (!c->psi) break; p = (char*)c->psi; real_geom(); break;
        case HYMD_FIELD_ELEC_FIELD:
            if (!d_ok || !c->emesh) break; p = (char*)c->emesh + d * gb; ghost_geom(); break;
        case HYMD_FIELD_PHI_LAPLACIAN:
            if (!t_ok || !d_ok || !c->lap || !c->have_lap) break;
            p = (char*)c->lap + (size_t)(3 * t + d) * rb; real_geom(); break;
        case HYMD_FIELD_GPE_EPS: p = (char*)gpe_field(c, 0, 0); if (p) real_geom(); break;
        case HYMD_FIELD_GPE_ELEC_DOT: p = (char*)gpe_field(c, 1, 0); if (p) real_geom(); break;
        case HYMD_FIELD_GPE_VBAR: p = (char*)gpe_field(c, 2, t); if (p) real_geom(); break;
        default: break;
    }
    if (!p) {
        set_error("field %d [t=%d, d=%d] is not available (not allocated / not materialized)",
                  field_id, t, d);
        return HYMD_ERR_STATE;
    }
    *d_ptr = p;
    return HYMD_OK;
}

int64_t hymd_launch_count(hymd_ctx* c) { return c ? c->launches : 0; }

int hymd_ctx_status(hymd_ctx* c, int64_t out[4]) {
    if (!c || !out) { set_error("null argument"); return HYMD_ERR_INVALID; }
    DeviceScalars h;
    HYMD_CUDA(cudaDeviceSynchronize());
    HYMD_CUDA(cudaMemcpy(&h, c->scalars, sizeof(h), cudaMemcpyDeviceToHost));
    out[0] = h.max_cell_count; out[1] = h.out_of_slab; out[2] = c->np; out[3] = c->U;
    return HYMD_OK;
}

int hymd_ctx_paths(hymd_ctx* c, int32_t out[4]) {
    if (!c || !out) { set_error("null argument"); return HYMD_ERR_INVALID; }
    out[0] = c->fused; out[1] = c->plane; out[2] = c->slab; out[3] = c->p2p;
    return HYMD_OK;
}

int hymd_nccl_unique_id(uint8_t id[HYMD_NCCL_UNIQUE_ID_BYTES]) {
    if (!id) { set_error("null argument"); return HYMD_ERR_INVALID; }
    return comm_unique_id(id);
}

int hymd_local_group_id(int world_size, uint8_t id[HYMD_NCCL_UNIQUE_ID_BYTES]) {
    if (!id) { set_error("null argument"); return HYMD_ERR_INVALID; }
    return comm_local_group_id(world_size, id);
}

int hymd_ctx_check(hymd_ctx* c) {
    if (!c) { set_error("null argument"); return HYMD_ERR_INVALID; }
    HYMD_CUDA(cudaDeviceSynchronize());
    return comm_check_status(c);
}

int hymd_migrate_plan(hymd_ctx* c, const void* d_pos, int64_t n, int64_t* n_new, void* stream) {
    if (!c || !n_new || (n > 0 && !d_pos)) { set_error("null argument"); return HYMD_ERR_INVALID; }
    if (n < 0 || n >= (1LL << 31)) { set_error("n = %lld out of range", (long long)n); return HYMD_ERR_CAPACITY; }
    if (c->g.P == 1) { *n_new = n; return HYMD_OK; }   // one slab: every particle is home
    return migrate_plan(c, d_pos, n, n_new, (cudaStream_t)stream);
}

int hymd_migrate_apply(hymd_ctx* c, const void* d_in, void* d_out, int32_t row_bytes, void* stream) {
    if (!c || !d_in || !d_out) { set_error("null argument"); return HYMD_ERR_INVALID; }
    if (c->g.P == 1) { set_error("hymd_migrate_apply: nothing to move with one slab"); return HYMD_ERR_STATE; }
    return migrate_apply(c, d_in, d_out, row_bytes, (cudaStream_t)stream);
}

}  // extern "C"
