// Fused k-space kernels (one pass over the spectra, 128-bit accesses).
//
// Density path, replaces per type (field.py:577, 582-612):
//   phi_fourier[t].apply(H); v = v_ext[t](phi~); v.r2c; apply(H); 3 x copy; 3 x apply(-i k_d)
// Because every shipped Hamiltonian has an affine v_ext (hamiltonian.py:188-191, 303-306,
// 470-473), V^_u(k) = H(k)^2 * sum_j A[u][j] phi^_j(k) (+ c_u at k = 0), evaluated directly on
// the raw density spectra; F^_{u,d} = -i k_d V^_u.  Types whose rows of A coincide share one
// potential ("unique row" u).
//
// PME path (field.py:369-396): psi^ = 4 pi c_e H rho^ / k^2 (k = 0 divisor replaced by 1),
// E^_d = -i k_d psi^.
//
// Nyquist rule (tests/test_oracle_analytic.py): the reference multiplies the
// stored half spectrum by -i k_d with the fftfreq sign convention and hands it to FFTW's c2r.
// The Hermitian-consistent equivalent is: the z wave number is 0 on the k_z = N_z/2 plane and
// the x (y) wave number is 0 on the index N_x/2 (N_y/2) line inside the self-conjugate planes
// k_z in {0, N_z/2}; everywhere else index N/2 carries -pi N/L.
#include "ctx.cuh"

namespace hymd {

struct KParams {
    int Nx, Ny, Nz, nyl, y0, Nzc, Nzcp;
    int T, U;
    long long npairs;    // Nx*nyl*Nzcp/2
    // strides in REALS (2 x complex elements) of each buffer: x stride, field stride
    long long xs_in, fs_in, xs_f, fs_f, xs_v, fs_v, xs_pf, fs_pf;
};

__device__ __forceinline__ void load4(const float* p, float v[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const double* p, double v[4]) {
    double2 a = *reinterpret_cast<const double2*>(p);
    double2 b = *reinterpret_cast<const double2*>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void store4(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void store4(double* p, double a, double b, double c, double d) {
    *reinterpret_cast<double2*>(p) = make_double2(a, b);
    *reinterpret_cast<double2*>(p + 2) = make_double2(c, d);
}

template <typename real>
struct KTables {
    const real *hx, *hy, *hz, *kx, *ky, *kz;
};

template <typename real>
__device__ __forceinline__ KTables<real> make_tables(const real* tab, const KParams& p) {
    KTables<real> t;
    t.hx = tab; t.hy = t.hx + p.Nx; t.hz = t.hy + p.Ny;
    t.kx = t.hz + p.Nzc; t.ky = t.kx + p.Nx; t.kz = t.ky + p.Ny;
    return t;
}

// Effective wave numbers of the two z entries handled by a thread, with the Nyquist rule.
template <typename real>
__device__ __forceinline__ void wave_numbers(const KTables<real>& tb, const KParams& p, int ix,
                                             int iy, int iz, real kxe[2], real kye[2],
                                             real kze[2], real h[2], bool valid[2]) {
    const real hxy = tb.hx[ix] * tb.hy[iy];
    const real kx = tb.kx[ix], ky = tb.ky[iy];
    const bool x_nyq = (p.Nx % 2 == 0) && ix == p.Nx / 2;
    const bool y_nyq = (p.Ny % 2 == 0) && iy == p.Ny / 2;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int z = iz + j;
        valid[j] = z < p.Nzc;
        const int zc = valid[j] ? z : 0;
        const bool z_nyq = (p.Nz % 2 == 0) && zc == p.Nz / 2;
        const bool self_conj = zc == 0 || z_nyq;
        h[j] = hxy * tb.hz[zc];
        kxe[j] = (x_nyq && self_conj) ? (real)0 : kx;
        kye[j] = (y_nyq && self_conj) ? (real)0 : ky;
        kze[j] = z_nyq ? (real)0 : tb.kz[zc];
    }
}

// TT > 0: compile-time number of types (inputs held in registers); TT == 0: runtime T.
template <typename real, int TT>
__global__ void __launch_bounds__(256) kspace_force_kernel(
    const real* __restrict__ phi_hat, real* __restrict__ f_hat, real* __restrict__ v_hat,
    real* __restrict__ phif_hat, const real* __restrict__ tab, const real* __restrict__ Au,
    const real* __restrict__ cu, KParams p) {
    const KTables<real> tb = make_tables(tab, p);
    const int hz2 = p.Nzcp / 2;
    const int T = TT > 0 ? TT : p.T;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.npairs; i += stride) {
        const int iz = (int)(i % hz2) * 2;
        const long long r = i / hz2;
        const int iyl = (int)(r % p.nyl);
        const int ix = (int)(r / p.nyl);
        const int iy = iyl + p.y0;
        real kxe[2], kye[2], kze[2], h[2];
        bool valid[2];
        wave_numbers(tb, p, ix, iy, iz, kxe, kye, kze, h, valid);
        const long long col = 2 * ((long long)iyl * p.Nzcp + iz);   // in reals
        const long long off = ix * p.xs_in + col, fs = p.fs_in;
        real in[TT > 0 ? TT : 1][4];
        if (TT > 0) {
#pragma unroll
            for (int t = 0; t < TT; ++t) load4(phi_hat + t * fs + off, in[t]);
        }
        if (phif_hat != nullptr) {
            const real s = Au[p.U * p.T];   // 1/M
            for (int t = 0; t < T; ++t) {
                real v[4];
                if (TT > 0) { v[0] = in[t][0]; v[1] = in[t][1]; v[2] = in[t][2]; v[3] = in[t][3]; }
                else load4(phi_hat + t * fs + off, v);
                const real s0 = valid[0] ? h[0] * s : (real)0, s1 = valid[1] ? h[1] * s : (real)0;
                store4(phif_hat + t * p.fs_pf + ix * p.xs_pf + col, v[0] * s0, v[1] * s0, v[2] * s1,
                       v[3] * s1);
            }
        }
        const real g0 = valid[0] ? h[0] * h[0] : (real)0, g1 = valid[1] ? h[1] * h[1] : (real)0;
        const bool origin = (ix == 0 && iy == 0 && iz == 0);
        for (int u = 0; u < p.U; ++u) {
            real a0 = 0, b0 = 0, a1 = 0, b1 = 0;
            if (TT > 0) {
#pragma unroll
                for (int t = 0; t < TT; ++t) {
                    const real a = Au[u * TT + t];
                    a0 += a * in[t][0]; b0 += a * in[t][1]; a1 += a * in[t][2]; b1 += a * in[t][3];
                }
            } else {
                for (int t = 0; t < T; ++t) {
                    real v[4];
                    load4(phi_hat + t * fs + off, v);
                    const real a = Au[u * T + t];
                    a0 += a * v[0]; b0 += a * v[1]; a1 += a * v[2]; b1 += a * v[3];
                }
            }
            a0 *= g0; b0 *= g0; a1 *= g1; b1 *= g1;
            // F_d = -i k_d (a + i b) = k_d b - i k_d a
            real* f = f_hat + (long long)(3 * u) * p.fs_f + ix * p.xs_f + col;
            store4(f, kxe[0] * b0, -kxe[0] * a0, kxe[1] * b1, -kxe[1] * a1);
            store4(f + p.fs_f, kye[0] * b0, -kye[0] * a0, kye[1] * b1, -kye[1] * a1);
            store4(f + 2 * p.fs_f, kze[0] * b0, -kze[0] * a0, kze[1] * b1, -kze[1] * a1);
            if (v_hat != nullptr) {
                if (origin) a0 += cu[u];
                store4(v_hat + (long long)u * p.fs_v + ix * p.xs_v + col, a0, b0, a1, b1);
            }
        }
    }
}

template <typename real>
__global__ void __launch_bounds__(256) kspace_pme_kernel(
    const real* __restrict__ rho_hat, real* __restrict__ e_hat, real* __restrict__ psi_hat,
    real* __restrict__ rhof_hat, const real* __restrict__ tab, real coef /* 4 pi c_e / M */,
    real inv_m, KParams p) {
    const KTables<real> tb = make_tables(tab, p);
    const int hz2 = p.Nzcp / 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.npairs; i += stride) {
        const int iz = (int)(i % hz2) * 2;
        const long long r = i / hz2;
        const int iyl = (int)(r % p.nyl);
        const int ix = (int)(r / p.nyl);
        const int iy = iyl + p.y0;
        real kxe[2], kye[2], kze[2], h[2];
        bool valid[2];
        wave_numbers(tb, p, ix, iy, iz, kxe, kye, kze, h, valid);
        const long long col = 2 * ((long long)iyl * p.Nzcp + iz);
        const long long off = ix * p.xs_in + col;       // single-field buffers (rho, psi, rho_f)
        const long long offe = ix * p.xs_f + col, fs = p.fs_f;
        real v[4];
        load4(rho_hat + off, v);
        const real kx = tb.kx[ix], ky = tb.ky[iy];
        real g[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int zc = valid[j] ? iz + j : 0;
            const real kz = tb.kz[zc];
            real k2 = kx * kx + ky * ky + kz * kz;
            if (ix == 0 && iy == 0 && zc == 0) k2 = (real)1;   // normp(p=2, zeromode=1)
            g[j] = valid[j] ? coef * h[j] / k2 : (real)0;
        }
        if (rhof_hat != nullptr) {
            const real s0 = valid[0] ? h[0] * inv_m : (real)0, s1 = valid[1] ? h[1] * inv_m : (real)0;
            store4(rhof_hat + off, v[0] * s0, v[1] * s0, v[2] * s1, v[3] * s1);
        }
        const real a0 = v[0] * g[0], b0 = v[1] * g[0], a1 = v[2] * g[1], b1 = v[3] * g[1];
        store4(e_hat + offe, kxe[0] * b0, -kxe[0] * a0, kxe[1] * b1, -kxe[1] * a1);
        store4(e_hat + fs + offe, kye[0] * b0, -kye[0] * a0, kye[1] * b1, -kye[1] * a1);
        store4(e_hat + 2 * fs + offe, kze[0] * b0, -kze[0] * a0, kze[1] * b1, -kze[1] * a1);
        if (psi_hat != nullptr) store4(psi_hat + off, a0, b0, a1, b1);
    }
}

// comp_laplacian (field.py:406-425): out[3t+d] = -k_d^2 * phi_fourier[t] (phi_fourier = the
// filtered, 1/M-normalised density spectra), one pass for all types and directions.  -k^2 is
// even, so the products stay Hermitian and the Nyquist entries need no special rule.
template <typename real>
__global__ void __launch_bounds__(256) kspace_laplacian_kernel(
    const real* __restrict__ phif_hat, real* __restrict__ lap_hat, const real* __restrict__ tab, KParams p) {
    const KTables<real> tb = make_tables(tab, p);
    const int hz2 = p.Nzcp / 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.npairs; i += stride) {
        const int iz = (int)(i % hz2) * 2;
        const long long r = i / hz2;
        const int iyl = (int)(r % p.nyl);
        const int ix = (int)(r / p.nyl);
        const int iy = iyl + p.y0;
        const real kx2 = -tb.kx[ix] * tb.kx[ix], ky2 = -tb.ky[iy] * tb.ky[iy];
        real kz2[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const bool valid = iz + j < p.Nzc;
            const real kz = tb.kz[valid ? iz + j : 0];
            kz2[j] = valid ? -kz * kz : (real)0;
        }
        const bool v1 = iz + 1 < p.Nzc;
        const long long col = 2 * ((long long)iyl * p.Nzcp + iz);
        for (int t = 0; t < p.T; ++t) {
            real v[4];
            load4(phif_hat + t * p.fs_in + ix * p.xs_in + col, v);
            if (!v1) { v[2] = 0; v[3] = 0; }
            real* o = lap_hat + (long long)(3 * t) * p.fs_f + ix * p.xs_f + col;
            store4(o, kx2 * v[0], kx2 * v[1], kx2 * v[2], kx2 * v[3]);
            store4(o + p.fs_f, ky2 * v[0], ky2 * v[1], ky2 * v[2], ky2 * v[3]);
            store4(o + 2 * p.fs_f, kz2[0] * v[0], kz2[0] * v[1], kz2[1] * v[2], kz2[1] * v[3]);
        }
    }
}

static KParams make_kparams(const hymd_ctx* c) {
    const Geometry& g = c->g;
    KParams p;
    p.Nx = g.Nx; p.Ny = g.Ny; p.Nz = g.Nz; p.nyl = g.nyl; p.y0 = g.y0;
    p.Nzc = g.Nzc; p.Nzcp = g.Nzcp; p.T = c->T; p.U = c->U;
    p.npairs = g.k_elems / 2;
    const KLayout lin = klayout(c, c->T), lf = klayout(c, 3 * c->U), lv = klayout(c, c->U);
    p.xs_in = 2 * lin.xs; p.fs_in = 2 * lin.fs;
    p.xs_f = 2 * lf.xs; p.fs_f = 2 * lf.fs;
    p.xs_v = 2 * lv.xs; p.fs_v = 2 * lv.fs;
    p.xs_pf = p.xs_in; p.fs_pf = p.fs_in;
    return p;
}

static unsigned int kgrid(long long npairs) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long want = (npairs + 255) / 256;
    long long cap = (long long)sms * 16;   // grid-stride: a multiple of the SM count
    return (unsigned int)(want < cap ? (want > 0 ? want : 1) : cap);
}

template <typename real>
static int launch_force(hymd_ctx* c, bool want_v, bool want_phif, cudaStream_t s) {
    KParams p = make_kparams(c);
    const unsigned int grid = kgrid(p.npairs);
    const real* in = (const real*)c->phi_hat;
    real* f = (real*)c->f_hat;
    real* v = want_v ? (real*)c->v_hat : nullptr;
    real* pf = want_phif ? (real*)c->phif_hat : nullptr;
    const real* tab = (const real*)c->tab;
    const real* Au = (const real*)c->Au;
    const real* cu = (const real*)c->cu;
#define HYMD_KCASE(TT)                                                                         \
    case TT:                                                                                   \
        kspace_force_kernel<real, TT><<<grid, 256, 0, s>>>(in, f, v, pf, tab, Au, cu, p);      \
        break;
    switch (c->T) {
        HYMD_KCASE(1) HYMD_KCASE(2) HYMD_KCASE(3) HYMD_KCASE(4) HYMD_KCASE(5) HYMD_KCASE(6)
        HYMD_KCASE(7) HYMD_KCASE(8)
        default:
            kspace_force_kernel<real, 0><<<grid, 256, 0, s>>>(in, f, v, pf, tab, Au, cu, p);
    }
#undef HYMD_KCASE
    HYMD_LAUNCH_CHECK(c);
    return HYMD_OK;
}

int kspace_forces(hymd_ctx* c, bool want_v, bool want_phif, cudaStream_t s) {
    return c->f64 ? launch_force<double>(c, want_v, want_phif, s)
                  : launch_force<float>(c, want_v, want_phif, s);
}

// phif_hat (T fields) -> lap_hat (3T fields, k layout of 3T fields)
int kspace_laplacian(hymd_ctx* c, cudaStream_t s) {
    KParams p = make_kparams(c);
    const KLayout lin = klayout(c, c->T), lo = klayout(c, 3 * c->T);
    p.xs_in = 2 * lin.xs; p.fs_in = 2 * lin.fs;
    p.xs_f = 2 * lo.xs; p.fs_f = 2 * lo.fs;
    const unsigned int grid = kgrid(p.npairs);
    if (c->f64)
        kspace_laplacian_kernel<double><<<grid, 256, 0, s>>>((const double*)c->phif_hat, (double*)c->lap_hat,
                                                            (const double*)c->tab, p);
    else
        kspace_laplacian_kernel<float><<<grid, 256, 0, s>>>((const float*)c->phif_hat, (float*)c->lap_hat,
                                                           (const float*)c->tab, p);
    HYMD_LAUNCH_CHECK(c);
    return HYMD_OK;
}

int kspace_pme(hymd_ctx* c, bool want_psi, cudaStream_t s) {
    KParams p = make_kparams(c);
    const KLayout l1 = klayout(c, 1), l3 = klayout(c, 3);
    p.xs_in = 2 * l1.xs; p.fs_in = 2 * l1.fs;
    p.xs_f = 2 * l3.xs; p.fs_f = 2 * l3.fs;
    const unsigned int grid = kgrid(p.npairs);
    const double m = (double)c->g.Nx * c->g.Ny * c->g.Nz;
    const double coef = 4.0 * 3.14159265358979323846 * c->cfg.elec_conversion / m;
    if (c->f64) {
        kspace_pme_kernel<double><<<grid, 256, 0, s>>>(
            (const double*)c->phiq_hat, (double*)c->e_hat, want_psi ? (double*)c->psi_hat : nullptr,
            want_psi ? (double*)c->phiqf_hat : nullptr, (const double*)c->tab, coef, 1.0 / m, p);
    } else {
        kspace_pme_kernel<float><<<grid, 256, 0, s>>>(
            (const float*)c->phiq_hat, (float*)c->e_hat, want_psi ? (float*)c->psi_hat : nullptr,
            want_psi ? (float*)c->phiqf_hat : nullptr, (const float*)c->tab, (float)coef,
            (float)(1.0 / m), p);
    }
    HYMD_LAUNCH_CHECK(c);
    return HYMD_OK;
}

}  // namespace hymd
