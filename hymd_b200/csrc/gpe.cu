// General-Poisson-equation electrostatics (coulombtype = "PIC_Spectral_GPE"), SURVEY.md section 8 row f3.
//
// Replaces update_field_force_q_GPE (hymd/field.py:964-1112) and compute_field_energy_q_GPE
// (field.py:706-760).  Per call: 1 charge paint, (6 + T + n_iter) forward and (11 + 3T + 3 n_iter)
// inverse transforms through the same plane / x-line / cuFFT pipeline as the force path (fft_forward /
// fft_inverse of slabfft.cu), one k-space kernel with four modes, a handful of pointwise kernels and
// a per-type readout of 3T electrostatic force meshes.  Several slabs: the same transforms (slabfft.cu), pointwise
// kernels on the owned planes, the convergence measure combined over the ranks on the device, guests read out and
// returned like in the force path.
//
// Reference semantics kept on purpose:
//  * the masked divisions `np.divide(x, y, where=y > 1e-6, out=o)` leave `o` untouched where the mask
//    is false (field.py:1019-1021, 1033-1035, 1085-1090), so phi_eps and elec_field_contrib are
//    persistent buffers of the context;
//  * the polarisation iteration starts from zero at every call: the reference rebinds its local
//    `phi_pol_prev` and never writes the caller's mesh (field.py:1046-1063);
//  * r2c carries 1/M, c2r none (every k-space mode applies `coef` = 1/M once per forward transform);
//  * Nyquist rule of kspace.cu for every i k_d product.
//
// Checked against the oracle (oracle/gpe_oracle.py, pinned on golden vectors of the reference's own function) in
// tests/test_zzgpu_gpe.py, on one GPU and on virtual slabs.
#include <stdlib.h>

#include "ctx.cuh"
#include "gpe.cuh"

namespace hymd {

int readout_custom(hymd_ctx* c, const void* mesh, const int* d_urow, void* d_force, cudaStream_t s);

struct GpeState {
    void *eps, *den, *eta, *pol, *tmp, *E, *dot, *contrib, *vbar;   // real meshes (eta, E: 3; vbar: T)
    void *kA, *kB, *kS;                                             // spectra: T, 3T, 1
    void* mesh;                                                     // 3T ghost-padded force meshes
    int* urow_id;
    double* d_par;                                                  // [T eps_t][T q_t]
    double* red;                                                    // block partials + result
    double* h_delta;                                                // pinned
    bool have;                                                      // eps / dot / psi valid
};

constexpr int GPE_BLOCKS = 148 * 4;

// The per-pair arithmetic lives in gpe.cuh (shared with the CPU check of tests/native/).
template <typename real>
__global__ void __launch_bounds__(256) gpe_kspace_kernel(const real* __restrict__ in, real* __restrict__ out_s,
                                                         real* __restrict__ out_v, const real* __restrict__ tab,
                                                         real coef, int use_h, int div_k2, real sign, GKParams p) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.npairs; i += stride)
        gpe_kspace_pair<real>(i, in, out_s, out_v, tab, coef, use_h, div_k2, sign, p);
}

// ---- pointwise kernels (n = cells of the local mesh) -------------------------------------------------
// phi_eps = sum_t eps_t phi_t / sum_t phi_t where the denominator exceeds 1e-6 (field.py:1012-1019)
template <typename real>
__global__ void __launch_bounds__(256) gpe_eps_kernel(const real* __restrict__ phi, long long fs, int T,
                                                      const double* __restrict__ eps_t, real* __restrict__ den,
                                                      real* __restrict__ eps, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        real num = 0, d = 0;
        for (int t = 0; t < T; ++t) {
            const real p = phi[t * fs + i];
            num = num + (real)eps_t[t] * p;
            d = d + p;
        }
        den[i] = d;
        if (d > (real)1e-6) eps[i] = num / d;
    }
}

// phi_q /= eps and eta_d /= eps where eps > 1e-6 (field.py:1021, 1033-1035)
template <typename real>
__global__ void __launch_bounds__(256) gpe_divide_kernel(real* __restrict__ phi_q, real* __restrict__ eta,
                                                         long long fs, const real* __restrict__ eps, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const real e = eps[i];
        if (e > (real)1e-6) {
            phi_q[i] = phi_q[i] / e;
            eta[i] = eta[i] / e; eta[fs + i] = eta[fs + i] / e; eta[2 * fs + i] = eta[2 * fs + i] / e;
        }
    }
}

// tmp = scale * (phi_q + pol) (field.py:1047, 1070)
template <typename real>
__global__ void __launch_bounds__(256) gpe_sum_kernel(const real* __restrict__ a, const real* __restrict__ b,
                                                      real scale, real* __restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = scale * (a[i] + b[i]);
}

// pol <- w * (-(eta . E)) + (1 - w) * pol and the convergence measure of |pol_new - pol_old|
// (field.py:1053-1061; mode 0 max, 1 sum, 2 sum of squares)
template <typename real>
__global__ void __launch_bounds__(256) gpe_pol_kernel(const real* __restrict__ eta, const real* __restrict__ E,
                                                      long long fs, real w, real* __restrict__ pol, int mode,
                                                      double* __restrict__ partial, long long n,
                                                      const double* __restrict__ state) {
    // state = {delta of the last counted iteration, iterations counted, conv_crit}: once delta <= conv_crit the
    // fixed point has stopped (field.py:1062-1064) and further launches of the un-synchronised batch change nothing
    if (state[0] <= state[2]) return;
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const real prev = pol[i];
        real p = -(eta[i] * E[i] + eta[fs + i] * E[fs + i] + eta[2 * fs + i] * E[2 * fs + i]);
        p = w * p + ((real)1.0 - w) * prev;
        pol[i] = p;
        const double d = fabs((double)(p - prev));
        if (mode == 0) acc = d > acc ? d : acc;
        else if (mode == 1) acc += d;
        else acc += d * d;
    }
    __shared__ double sh[256];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) {
            const double o = sh[threadIdx.x + k];
            sh[threadIdx.x] = mode == 0 ? (o > sh[threadIdx.x] ? o : sh[threadIdx.x]) : sh[threadIdx.x] + o;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// out = {delta, iterations, conv_crit} when `gate` (the polarisation loop): a converged loop is left alone
__global__ void __launch_bounds__(256) gpe_reduce_kernel(const double* __restrict__ partial, int nblocks, int mode,
                                                         double scale, double* __restrict__ out, int gate) {
    if (gate && out[0] <= out[2]) return;
    __shared__ double sh[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256) acc = mode == 0 ? (partial[i] > acc ? partial[i] : acc) : acc + partial[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) {
            const double o = sh[threadIdx.x + k];
            sh[threadIdx.x] = mode == 0 ? (o > sh[threadIdx.x] ? o : sh[threadIdx.x]) : sh[threadIdx.x] + o;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = sh[0] * scale;
        if (gate) out[1] += 1.0;
    }
}

// The convergence measure of all ranks (csum / cnorm are global sums, max_diff an MPI.MAX: main.py:141-163) into the
// loop state {delta, iterations, conv_crit}; a converged loop is left alone (every rank sees the same state).
__global__ void gpe_combine_kernel(const double* __restrict__ all, int P, int mode, double* __restrict__ state) {
    if (state[0] <= state[2]) return;
    double v = all[0];
    for (int q = 1; q < P; ++q) v = mode == 0 ? (all[q] > v ? all[q] : v) : v + all[q];
    state[0] = v;
    state[1] += 1.0;
}

// elec_dot = |E|^2, elec_field_contrib = elec_dot / den where den > 1e-6, and
// Vbar_t = q_t psi - (0.5 / eps0_inv) (eps_t - phi_eps) elec_field_contrib   (field.py:1079-1097)
template <typename real>
__global__ void __launch_bounds__(256) gpe_vbar_kernel(const real* __restrict__ E, long long fs,
                                                       const real* __restrict__ den, const real* __restrict__ eps,
                                                       const real* __restrict__ psi, int T,
                                                       const double* __restrict__ par, real half_eps0,
                                                       real* __restrict__ dot, real* __restrict__ contrib,
                                                       real* __restrict__ vbar, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const real d = E[i] * E[i] + E[fs + i] * E[fs + i] + E[2 * fs + i] * E[2 * fs + i];
        dot[i] = d;
        if (den[i] > (real)1e-6) contrib[i] = d / den[i];
        const real cb = contrib[i], ps = psi[i], e = eps[i];
        for (int t = 0; t < T; ++t)
            vbar[t * fs + i] = (real)par[T + t] * ps - half_eps0 * ((real)par[t] - e) * cb;
    }
}

// sum_cells eps * dot (field.py:757-759), fixed order
template <typename real>
__global__ void __launch_bounds__(256) gpe_energy_kernel(const real* __restrict__ eps, const real* __restrict__ dot,
                                                         double* __restrict__ partial, long long n) {
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride)
        acc += (double)eps[i] * (double)dot[i];
    __shared__ double sh[256];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

static int galloc(void** p, size_t bytes, bool zero) {
    if (*p) return HYMD_OK;
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        return HYMD_ERR_NOMEM;
    }
    if (zero) { cudaMemset(*p, 0, bytes); cudaDeviceSynchronize(); }
    return HYMD_OK;
}

static int gpe_state(hymd_ctx* c) {
    if (!c->gpe) { c->gpe = new GpeState(); memset(c->gpe, 0, sizeof(GpeState)); }
    GpeState* st = c->gpe;
    const Geometry& g = c->g;
    const size_t rb = (size_t)g.real_elems * c->rsz, kb = (size_t)g.k_elems * 2 * c->rsz,
                 gb = (size_t)g.ghost_elems * c->rsz;
    const int T = c->T;
    HYMD_CHECK(galloc(&st->eps, rb, true));
    HYMD_CHECK(galloc(&st->den, rb, true));
    HYMD_CHECK(galloc(&st->eta, 3 * rb, true));
    HYMD_CHECK(galloc(&st->pol, rb, true));
    HYMD_CHECK(galloc(&st->tmp, rb, true));
    HYMD_CHECK(galloc(&st->E, 3 * rb, true));
    HYMD_CHECK(galloc(&st->dot, rb, true));
    HYMD_CHECK(galloc(&st->contrib, rb, true));
    HYMD_CHECK(galloc(&st->vbar, (size_t)T * rb, true));
    HYMD_CHECK(galloc(&st->kA, (size_t)T * kb, true));
    HYMD_CHECK(galloc(&st->kB, (size_t)3 * T * kb, true));
    HYMD_CHECK(galloc(&st->kS, kb, true));
    HYMD_CHECK(galloc(&st->mesh, (size_t)3 * T * gb, true));
    HYMD_CHECK(galloc((void**)&st->d_par, sizeof(double) * 2 * HYMD_MAX_TYPES, true));
    HYMD_CHECK(galloc((void**)&st->red, sizeof(double) * (GPE_BLOCKS + 4 + HYMD_MAX_PEERS), true));
    HYMD_CHECK(galloc(&c->psi, rb, true));
    if (!st->urow_id) {
        HYMD_CHECK(galloc((void**)&st->urow_id, sizeof(int) * HYMD_MAX_TYPES, true));
        int id[HYMD_MAX_TYPES];
        for (int t = 0; t < HYMD_MAX_TYPES; ++t) id[t] = t;
        HYMD_CUDA(cudaMemcpy(st->urow_id, id, sizeof(id), cudaMemcpyHostToDevice));
        HYMD_CUDA(cudaDeviceSynchronize());
    }
    if (!st->h_delta) HYMD_CUDA(cudaMallocHost((void**)&st->h_delta, 4 * sizeof(double)));
    return HYMD_OK;
}

void gpe_destroy(hymd_ctx* c) {
    if (!c->gpe) return;
    GpeState* st = c->gpe;
    void* bufs[] = {st->eps, st->den, st->eta, st->pol, st->tmp, st->E, st->dot, st->contrib, st->vbar, st->kA,
                    st->kB, st->kS, st->mesh, st->urow_id, st->d_par, st->red};
    for (void* b : bufs)
        if (b) cudaFree(b);
    if (st->h_delta) cudaFreeHost(st->h_delta);
    delete st;
    c->gpe = nullptr;
}

template <typename real>
static int kspace(hymd_ctx* c, const void* in, int F, void* out_s, void* out_v, double coef, bool use_h,
                  bool div_k2, double sign, cudaStream_t s) {
    const Geometry& g = c->g;
    GKParams p;
    p.Nx = g.Nx; p.Ny = g.Ny; p.Nz = g.Nz; p.nyl = g.nyl; p.y0 = g.y0; p.Nzc = g.Nzc; p.Nzcp = g.Nzcp; p.F = F;
    p.npairs = g.k_elems / 2;
    const KLayout li = klayout(c, F), lv = klayout(c, 3 * F);
    p.xs_in = 2 * li.xs; p.fs_in = 2 * li.fs; p.xs_s = p.xs_in; p.fs_s = p.fs_in;
    p.xs_v = 2 * lv.xs; p.fs_v = 2 * lv.fs;
    long long want = (p.npairs + 255) / 256;
    const unsigned int grid = (unsigned int)(want < 148 * 16 ? (want > 0 ? want : 1) : 148 * 16);
    gpe_kspace_kernel<real><<<grid, 256, 0, s>>>((const real*)in, (real*)out_s, (real*)out_v, (const real*)c->tab,
                                                 (real)coef, use_h ? 1 : 0, div_k2 ? 1 : 0, (real)sign, p);
    HYMD_LAUNCH_CHECK(c);
    return HYMD_OK;
}

template <typename real>
static int gpe_cycle_t(hymd_ctx* c, const hymd_gpe_params* prm, void* d_force, int32_t* iterations, cudaStream_t s) {
    GpeState* st = c->gpe;
    const Geometry& g = c->g;
    const int T = c->T;
    // pointwise kernels and reductions run over the owned planes (the first nxl of a field's nxl + 1 planes on
    // several slabs: the ghost plane belongs to the next slab's sums)
    const long long n = (long long)g.nxl * g.Ny * g.Nz, fs = g.real_elems;
    const double M = (double)g.Nx * g.Ny * g.Nz;
    const double eps0_inv = prm->coulomb_constant * 4.0 * 3.14159265358979323846;
    double par[2 * HYMD_MAX_TYPES];
    for (int t = 0; t < T; ++t) { par[t] = prm->dielectric_type[t]; par[T + t] = prm->type_charges[t]; }
    HYMD_CUDA(cudaMemcpyAsync(st->d_par, par, sizeof(double) * 2 * T, cudaMemcpyHostToDevice, s));
    real* phi_q = (real*)c->phi_q;
    PhaseScope ps(c, HYMD_PHASE_PME_KSPACE, s);
    // smeared charge density (field.py:1006-1010)
    HYMD_CHECK(paint_charges(c, s));
    HYMD_CHECK(halo_reduce(c, c->phi_q, 1, s));
    HYMD_CHECK(fft_forward(c, c->phi_q, 1, st->kA, s));
    HYMD_CHECK(kspace<real>(c, st->kA, 1, st->kS, nullptr, 1.0 / M, true, false, 1.0, s));
    HYMD_CHECK(fft_inverse(c, st->kS, 1, c->phi_q, false, s));
    // dielectric field and its scaled gradient (field.py:1012-1035)
    gpe_eps_kernel<real><<<GPE_BLOCKS, 256, 0, s>>>((const real*)c->phi, fs, T, st->d_par, (real*)st->den,
                                                    (real*)st->eps, n);
    HYMD_LAUNCH_CHECK(c);
    HYMD_CHECK(fft_forward(c, st->eps, 1, st->kA, s));
    HYMD_CHECK(kspace<real>(c, st->kA, 1, nullptr, st->kB, 1.0 / M, false, false, 1.0, s));
    HYMD_CHECK(fft_inverse(c, st->kB, 3, st->eta, false, s));
    gpe_divide_kernel<real><<<GPE_BLOCKS, 256, 0, s>>>(phi_q, (real*)st->eta, fs, (const real*)st->eps, n);
    HYMD_LAUNCH_CHECK(c);
    // polarisation-charge fixed point (field.py:1037-1064), from zero at every call
    HYMD_CUDA(cudaMemsetAsync(st->pol, 0, (size_t)n * sizeof(real), s));
    // The convergence test runs on the device: {delta, iterations, conv_crit} live in st->red + GPE_BLOCKS, the
    // update and the reduction of an iteration are skipped once delta <= conv_crit, and the host looks at the state
    // only once per batch of GPE_BATCH iterations (it was one stream synchronisation per iteration).  The iterations
    // of a batch that come after convergence only recompute scratch (tmp, E), which the next stage overwrites.
    constexpr int GPE_BATCH = 4;
    const int max_iter = prm->max_iter > 0 ? prm->max_iter : 100;
    double* state = st->red + GPE_BLOCKS;
    st->h_delta[0] = 1.0; st->h_delta[1] = 0.0; st->h_delta[2] = prm->conv_crit;     // the reference starts from delta = 1
    HYMD_CUDA(cudaMemcpyAsync(state, st->h_delta, 3 * sizeof(double), cudaMemcpyHostToDevice, s));
    int it = 0, launched = 0;
    double delta = 1.0;
    while (launched < max_iter && delta > prm->conv_crit) {
        const int batch = max_iter - launched < GPE_BATCH ? max_iter - launched : GPE_BATCH;
        for (int b = 0; b < batch; ++b) {
            gpe_sum_kernel<real><<<GPE_BLOCKS, 256, 0, s>>>(phi_q, (const real*)st->pol, (real)1, (real*)st->tmp, n);
            HYMD_LAUNCH_CHECK(c);
            HYMD_CHECK(fft_forward(c, st->tmp, 1, st->kA, s));
            HYMD_CHECK(kspace<real>(c, st->kA, 1, nullptr, st->kB, 1.0 / M, false, true, -1.0, s));
            HYMD_CHECK(fft_inverse(c, st->kB, 3, st->E, false, s));
            gpe_pol_kernel<real><<<GPE_BLOCKS, 256, 0, s>>>((const real*)st->eta, (const real*)st->E, fs,
                                                            (real)prm->pol_mixing, (real*)st->pol,
                                                            prm->convergence_type, st->red, n, state);
            HYMD_LAUNCH_CHECK(c);
            double* xloc = st->red + GPE_BLOCKS + 3;
            double* xall = xloc + 1;
            gpe_reduce_kernel<<<1, 256, 0, s>>>(st->red, GPE_BLOCKS, prm->convergence_type, 1.0, xloc, 0);
            HYMD_LAUNCH_CHECK(c);
            if (g.P > 1) HYMD_CHECK(comm_allgather_host(c, xloc, xall, sizeof(double), s));
            gpe_combine_kernel<<<1, 1, 0, s>>>(g.P > 1 ? xall : xloc, g.P, prm->convergence_type, state);
            HYMD_LAUNCH_CHECK(c);
        }
        launched += batch;
        HYMD_CUDA(cudaMemcpyAsync(st->h_delta, state, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
        HYMD_CUDA(cudaStreamSynchronize(s));
        delta = st->h_delta[0];
        it = (int)st->h_delta[1];
    }
    if (iterations) *iterations = it;
    // potential and field (field.py:1066-1084)
    gpe_sum_kernel<real><<<GPE_BLOCKS, 256, 0, s>>>(phi_q, (const real*)st->pol, (real)eps0_inv, (real*)st->tmp, n);
    HYMD_LAUNCH_CHECK(c);
    HYMD_CHECK(fft_forward(c, st->tmp, 1, st->kA, s));
    HYMD_CHECK(kspace<real>(c, st->kA, 1, st->kS, st->kB, 1.0 / M, false, true, -1.0, s));
    HYMD_CHECK(fft_inverse(c, st->kS, 1, c->psi, false, s));
    HYMD_CHECK(fft_inverse(c, st->kB, 3, st->E, false, s));
    // |E|^2, its density-weighted form and the per-type electrostatic potential (field.py:1079-1097)
    gpe_vbar_kernel<real><<<GPE_BLOCKS, 256, 0, s>>>((const real*)st->E, fs, (const real*)st->den, (const real*)st->eps,
                                                     (const real*)c->psi, T, st->d_par, (real)(0.5 / eps0_inv),
                                                     (real*)st->dot, (real*)st->contrib, (real*)st->vbar, n);
    HYMD_LAUNCH_CHECK(c);
    // filtered -grad Vbar_t on the mesh and at the particles (field.py:1099-1111)
    HYMD_CHECK(fft_forward(c, st->vbar, T, st->kA, s));
    HYMD_CHECK(kspace<real>(c, st->kA, T, nullptr, st->kB, 1.0 / M, true, false, -1.0, s));
    HYMD_CHECK(fft_inverse(c, st->kB, 3 * T, st->mesh, true, s));
    if (!c->plane) HYMD_CHECK(fill_ghosts(c, st->mesh, 3 * T, s));
    HYMD_CHECK(halo_fetch(c, st->mesh, 3 * T, s));
    c->have_psi = true;
    c->have_phiq_hat = false;      // phi_q now holds the filtered, eps-scaled density (reference semantics)
    st->have = true;
    // several slabs: collective (a rank without particles still reads out its guests)
    if (g.P > 1 || (c->np > 0 && d_force)) HYMD_CHECK(readout_custom(c, st->mesh, st->urow_id, d_force, s));
    return HYMD_OK;
}

int gpe_cycle(hymd_ctx* c, const hymd_gpe_params* prm, void* d_force, int32_t* iterations, cudaStream_t s) {
    HYMD_CHECK(gpe_state(c));
    return c->f64 ? gpe_cycle_t<double>(c, prm, d_force, iterations, s)
                  : gpe_cycle_t<float>(c, prm, d_force, iterations, s);
}

int gpe_energy(hymd_ctx* c, double coulomb_constant, double* out, cudaStream_t s) {
    GpeState* st = c->gpe;
    const Geometry& g = c->g;
    const long long n = (long long)g.nxl * g.Ny * g.Nz;       // owned planes: the Python layer sums over the ranks
    if (c->f64) gpe_energy_kernel<double><<<GPE_BLOCKS, 256, 0, s>>>((const double*)st->eps, (const double*)st->dot, st->red, n);
    else gpe_energy_kernel<float><<<GPE_BLOCKS, 256, 0, s>>>((const float*)st->eps, (const float*)st->dot, st->red, n);
    HYMD_LAUNCH_CHECK(c);
    const double dv = g.box[0] * g.box[1] * g.box[2] / ((double)g.Nx * g.Ny * g.Nz);
    const double eps_0 = 1.0 / (coulomb_constant * 4.0 * 3.14159265358979323846);
    gpe_reduce_kernel<<<1, 256, 0, s>>>(st->red, GPE_BLOCKS, 1, dv * 0.5 * eps_0, st->red + GPE_BLOCKS, 0);
    HYMD_LAUNCH_CHECK(c);
    HYMD_CUDA(cudaMemcpyAsync(st->h_delta, st->red + GPE_BLOCKS, sizeof(double), cudaMemcpyDeviceToHost, s));
    HYMD_CUDA(cudaStreamSynchronize(s));
    *out = st->h_delta[0];
    return HYMD_OK;
}

void* gpe_field(hymd_ctx* c, int which, int t) {
    GpeState* st = c->gpe;
    if (!st || !st->have) return nullptr;
    const size_t rb = (size_t)c->g.real_elems * c->rsz;
    switch (which) {
        case 0: return st->eps;
        case 1: return st->dot;
        case 2: return (t >= 0 && t < c->T) ? (char*)st->vbar + (size_t)t * rb : nullptr;
        default: return nullptr;
    }
}

}  // namespace hymd

using namespace hymd;

extern "C" {

int hymd_gpe_cycle(hymd_ctx* c, const hymd_gpe_params* prm, void* d_elec_force, int32_t* iterations, void* stream) {
    if (!c || !prm) { set_error("null argument"); return HYMD_ERR_INVALID; }
    if (prm->struct_size != (int32_t)sizeof(hymd_gpe_params)) {
        set_error("hymd_gpe_params size mismatch: caller %d, library %zu", prm->struct_size, sizeof(hymd_gpe_params));
        return HYMD_ERR_INVALID;
    }
    if (!c->cfg.pme) { set_error("context created without the charge-density buffers (pme = 0)"); return HYMD_ERR_STATE; }
    if (!c->sorted || !c->has_charges) { set_error("hymd_gpe_cycle needs hymd_sort_particles with charges"); return HYMD_ERR_STATE; }
    if (prm->convergence_type < 0 || prm->convergence_type > 2 || !(prm->conv_crit > 0) ||
        !(prm->pol_mixing > 0) || !(prm->coulomb_constant > 0)) {
        set_error("hymd_gpe_cycle: bad parameters (convergence_type 0..2, conv_crit, pol_mixing, coulomb_constant > 0)");
        return HYMD_ERR_INVALID;
    }
    // the dielectric field is built from the FILTERED type densities of the last update_field
    HYMD_CHECK(hymd_materialize(c, 1, 0, 0, 0, stream));
    return gpe_cycle(c, prm, d_elec_force, iterations, (cudaStream_t)stream);
}

int hymd_gpe_energy(hymd_ctx* c, double coulomb_constant, double* out, void* stream) {
    if (!c || !out) { set_error("null argument"); return HYMD_ERR_INVALID; }
    if (!c->gpe || !c->gpe->have) { set_error("hymd_gpe_energy before hymd_gpe_cycle"); return HYMD_ERR_STATE; }
    return gpe_energy(c, coulomb_constant, out, (cudaStream_t)stream);
}

}  // extern "C"
