// k-space arithmetic of the GPE electrostatics (gpe.cu), __host__ __device__ so that
// tests/native/host_check.cpp can run exactly this source on the CPU against numpy.
#pragma once
#include <stdint.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#endif

namespace hymd {

struct GKParams {
    int Nx, Ny, Nz, nyl, y0, Nzc, Nzcp, F;
    long long npairs, xs_in, fs_in, xs_s, fs_s, xs_v, fs_v;   // strides in reals
};

// One pair of adjacent k_z entries (index i of Nx * nyl * Nzcp/2):
//   out_s[f] = in[f] * g,  out_v[3f+d] = sign * i k_d * in[f] * g,
//   g = coef * (use_h ? H : 1) / (div_k2 ? k^2 : 1)     (k^2 with the origin replaced by 1: normp(p=2, zeromode=1))
// tab = [hx[Nx] hy[Ny] hz[Nzc] kx[Nx] ky[Ny] kz[Nzc]] (context.cu build_tables); Nyquist rule of kspace.cu.
template <typename real>
__host__ __device__ inline void gpe_kspace_pair(long long i, const real* __restrict__ in, real* __restrict__ out_s,
                                                real* __restrict__ out_v, const real* __restrict__ tab, real coef,
                                                int use_h, int div_k2, real sign, const GKParams& p) {
    const real* hx = tab; const real* hy = hx + p.Nx; const real* hz = hy + p.Ny;
    const real* kxt = hz + p.Nzc; const real* kyt = kxt + p.Nx; const real* kzt = kyt + p.Ny;
    const int hz2 = p.Nzcp / 2;
    const int iz = (int)(i % hz2) * 2;
    const long long r = i / hz2;
    const int iyl = (int)(r % p.nyl);
    const int ix = (int)(r / p.nyl);
    const int iy = iyl + p.y0;
    const real kx = kxt[ix], ky = kyt[iy];
    const bool x_nyq = (p.Nx % 2 == 0) && ix == p.Nx / 2;
    const bool y_nyq = (p.Ny % 2 == 0) && iy == p.Ny / 2;
    real g[2], kxe[2], kye[2], kze[2];
    for (int j = 0; j < 2; ++j) {
        const int z = iz + j;
        const bool valid = z < p.Nzc;
        const int zc = valid ? z : 0;
        const bool z_nyq = (p.Nz % 2 == 0) && zc == p.Nz / 2;
        const bool self_conj = zc == 0 || z_nyq;
        const real kz = kzt[zc];
        real k2 = kx * kx + ky * ky + kz * kz;
        if (ix == 0 && iy == 0 && zc == 0) k2 = (real)1;
        real gg = coef;
        if (use_h) gg *= hx[ix] * hy[iy] * hz[zc];
        if (div_k2) gg /= k2;
        g[j] = valid ? gg : (real)0;
        kxe[j] = (x_nyq && self_conj) ? (real)0 : kx;
        kye[j] = (y_nyq && self_conj) ? (real)0 : ky;
        kze[j] = z_nyq ? (real)0 : kz;
    }
    const long long col = 2 * ((long long)iyl * p.Nzcp + iz);
    for (int f = 0; f < p.F; ++f) {
        const real* src = in + f * p.fs_in + ix * p.xs_in + col;
        const real a0 = src[0] * g[0], b0 = src[1] * g[0], a1 = src[2] * g[1], b1 = src[3] * g[1];
        if (out_s != nullptr) {
            real* o = out_s + f * p.fs_s + ix * p.xs_s + col;
            o[0] = a0; o[1] = b0; o[2] = a1; o[3] = b1;
        }
        if (out_v != nullptr) {
            // sign * i k (a + i b) = sign * (-k b + i k a)
            real* o = out_v + (long long)(3 * f) * p.fs_v + ix * p.xs_v + col;
            o[0] = -sign * kxe[0] * b0; o[1] = sign * kxe[0] * a0; o[2] = -sign * kxe[1] * b1; o[3] = sign * kxe[1] * a1;
            o += p.fs_v;
            o[0] = -sign * kye[0] * b0; o[1] = sign * kye[0] * a0; o[2] = -sign * kye[1] * b1; o[3] = sign * kye[1] * a1;
            o += p.fs_v;
            o[0] = -sign * kze[0] * b0; o[1] = sign * kze[0] * a0; o[2] = -sign * kze[1] * b1; o[3] = sign * kze[1] * a1;
        }
    }
}

}  // namespace hymd
