// Per-particle evaluation of HyMD's intramolecular forces (bonds, angles, dihedrals).
//
// Reference kernels (Fortran, sequential over terms, read-modify-write of f per term):
//   hymd/compute_bond_forces.f90:1-61        cbf
//   hymd/compute_angle_forces.f90:1-93       caf
//   hymd/compute_dihedral_forces.f90:1-137   cdf   (dtype 0: cosine series, 1: combined bending-torsion, 2: improper)
//   hymd/dipole_reconstruction.f90:50-221    reconstruct (bending term, backbone dipoles, transfer matrices)
//   hymd/dipole_reconstruction.f90:37-48     cosine_series
//
// B200 design: no scatter, no atomics.  A host-built CSR lists, for every particle, the terms it
// takes part in (term index and slot a/b/c/d, ascending term order = the Fortran accumulation
// order); one thread owns one particle, re-evaluates each of its terms from the (L1/L2-resident,
// molecule-contiguous) neighbour positions and writes its force exactly once.  Re-evaluating a term
// 2-4 times costs flops the kernel has to spare; it saves the term-force round trip through HBM and
// makes the result bitwise reproducible.  Energy and the pressure by-products are counted by the
// thread that holds slot 0 of a term and reduced in a fixed order.
//
// Arithmetic follows the Fortran: position differences in the position type (real(4) for the fp32
// build), everything after that in double (the Fortran locals are real(8) in both builds).
//
// The per-particle functions are __host__ __device__ so that tests/native/bonded_host_check.cu can
// run exactly this source on the CPU against the oracle when no GPU is present.
#pragma once
#include <math.h>
#include <stdint.h>

#include <vector>

#include "md.cuh"

#ifndef __CUDACC__
#define __host__
#define __device__
#endif

namespace hymd {

constexpr int DIH_ROWS = 6, DIH_COLS = 5;   // prepare_bonds: bonds_4_coeff (D,6,5), force.py:678-690

struct Vec3d {
    double x, y, z;
};
__host__ __device__ inline Vec3d operator+(Vec3d a, Vec3d b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__host__ __device__ inline Vec3d operator-(Vec3d a, Vec3d b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__host__ __device__ inline Vec3d operator*(Vec3d a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__host__ __device__ inline Vec3d mul(Vec3d a, Vec3d b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
__host__ __device__ inline double dot(Vec3d a, Vec3d b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ inline Vec3d cross(Vec3d a, Vec3d b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// r(i,:) - r(j,:) in the position type, then minimum image `d - box * nint(d / box)` in double
// (compute_bond_forces.f90:47-48; nint rounds half away from zero like round()).
template <typename real>
__host__ __device__ inline Vec3d mic_diff(const real* __restrict__ pos, long long i, long long j, Vec3d box) {
    Vec3d d = {(double)(pos[3 * i + 0] - pos[3 * j + 0]), (double)(pos[3 * i + 1] - pos[3 * j + 1]),
               (double)(pos[3 * i + 2] - pos[3 * j + 2])};
    d.x -= box.x * round(d.x / box.x);
    d.y -= box.y * round(d.y / box.y);
    d.z -= box.z * round(d.z / box.z);
    return d;
}

// Positions of a CTA's own particles [p0, p1) staged in shared memory, everything else from global
// memory (atoms of molecules that straddle the CTA boundary).
template <typename real>
struct PosTile {
    const real* g;
    const real* s;
    long long p0, p1;
    __host__ __device__ inline real get(long long i, int d) const {
        return (i >= p0 && i < p1) ? s[3 * (i - p0) + d] : g[3 * i + d];
    }
};
template <typename real>
__host__ __device__ inline Vec3d mic_diff(const PosTile<real>& pos, long long i, long long j, Vec3d box) {
    Vec3d d = {(double)(pos.get(i, 0) - pos.get(j, 0)), (double)(pos.get(i, 1) - pos.get(j, 1)),
               (double)(pos.get(i, 2) - pos.get(j, 2))};
    d.x -= box.x * round(d.x / box.x);
    d.y -= box.y * round(d.y / box.y);
    d.z -= box.z * round(d.z / box.z);
    return d;
}

struct BondAcc {      // what one particle accumulates
    Vec3d f;          // force on the particle
    double e;         // energy of the terms it holds slot 0 of
    Vec3d pr;         // pressure by-product of those terms
};

// ---- two-particle bonds ------------------------------------------------------------------------
// Every *_eval function evaluates one term completely (all slot forces, energy, pressure by-product);
// the per-particle path keeps the share of its slot, the CTA-cooperative path stores all of them.
template <typename P>
__host__ __device__ inline void bond_eval(const P& pos, Vec3d box, int ia, int ib, double r0,
                                          double k, Vec3d& fa, double& e, Vec3d& pr) {
    const Vec3d rab = mic_diff(pos, (long long)ib, (long long)ia, box);
    const double n = sqrt(dot(rab, rab));
    const double df = k * (n - r0);
    fa = rab * (-df / n);
    e = 0.5 * k * (n - r0) * (n - r0);
    pr = mul(fa, rab);
}

__host__ __device__ inline void bond_apply(int slot, Vec3d fa, Vec3d& f) {
    f = slot == 0 ? f - fa : f + fa;       // f(aa) -= fa, f(bb) += fa
}

template <typename real>
__host__ __device__ inline void bond_term(const real* __restrict__ pos, Vec3d box, int ia, int ib, double r0,
                                          double k, int slot, BondAcc& acc) {
    Vec3d fa, pr;
    double e;
    bond_eval(pos, box, ia, ib, r0, k, fa, e, pr);
    bond_apply(slot, fa, acc.f);
    if (slot == 0) {
        acc.e += e;
        acc.pr = acc.pr + pr;
    }
}

// ---- three-particle angles ---------------------------------------------------------------------
// Returns false (nothing to add) when cos^2 >= 1, like the Fortran's `if (cosphi2 < 1.0)`.
template <typename P>
__host__ __device__ inline bool angle_eval(const P& pos, Vec3d box, int ia, int ib, int ic,
                                           double t0, double k, Vec3d& fa, Vec3d& fc, double& e, Vec3d& pr) {
    const Vec3d ra = mic_diff(pos, (long long)ia, (long long)ib, box);
    const Vec3d rc = mic_diff(pos, (long long)ic, (long long)ib, box);
    const double na = sqrt(dot(ra, ra)), nc = sqrt(dot(rc, rc));
    const Vec3d ea = ra * (1.0 / na), ec = rc * (1.0 / nc);
    const double cosphi = dot(ea, ec);
    if (!(cosphi * cosphi < 1.0)) return false;
    const double theta = acos(cosphi);
    const double sinphi = sin(theta);
    const double d = theta - t0;
    const double ff = k * d;
    const double xra = -ff / (na * sinphi), xrc = -ff / (nc * sinphi);
    fa = (ec - ea * cosphi) * xra;
    fc = (ea - ec * cosphi) * xrc;
    e = 0.5 * ff * d;
    const Vec3d zero = {0.0, 0.0, 0.0};
    pr = zero - mul(fa, ra) - mul(fc, rc);
    return true;
}

__host__ __device__ inline void angle_apply(int slot, Vec3d fa, Vec3d fc, Vec3d& f) {
    if (slot == 0) f = f - fa;             // f(aa) -= fa
    else if (slot == 2) f = f - fc;        // f(cc) -= fc
    else f = f + fa + fc;                  // f(bb) += fa + fc
}

template <typename real>
__host__ __device__ inline void angle_term(const real* __restrict__ pos, Vec3d box, int ia, int ib, int ic,
                                           double t0, double k, int slot, BondAcc& acc) {
    Vec3d fa, fc, pr;
    double e;
    if (angle_eval(pos, box, ia, ib, ic, t0, k, fa, fc, e, pr)) {
        angle_apply(slot, fa, fc, acc.f);
        if (slot == 0) {
            acc.e += e;
            acc.pr = acc.pr + pr;
        }
    }
}

// ---- four-particle dihedrals -------------------------------------------------------------------
__host__ __device__ inline void cosine_series(const double* __restrict__ c_n, const double* __restrict__ d_n,
                                              double phi, double& energy, double& de) {
    for (int i = 0; i < DIH_COLS; ++i) {
        energy += c_n[i] * (1.0 + cos(i * phi - d_n[i]));
        de -= i * c_n[i] * sin(i * phi - d_n[i]);
    }
}

// out[slot] = what is ADDED to the force of the particle in that slot (compute_dihedral_forces.f90:121-134)
template <typename P>
__host__ __device__ inline void dihedral_eval(const P& pos, Vec3d box, int ia, int ib, int ic,
                                              int id, const double* __restrict__ coeff, int dtype, Vec3d* out,
                                              double& e) {
    const Vec3d f = mic_diff(pos, (long long)ia, (long long)ib, box);
    const Vec3d g = mic_diff(pos, (long long)ib, (long long)ic, box);
    const Vec3d h = mic_diff(pos, (long long)id, (long long)ic, box);
    const Vec3d v = cross(f, g), w = cross(h, g);
    const double v_sq = dot(v, v), w_sq = dot(w, w);
    const double g_norm = sqrt(dot(g, g));
    const double cos_phi = dot(v, w);
    const double sin_phi = dot(w, f) * g_norm;
    const double phi = atan2(sin_phi, cos_phi);
    const double f_dot_g = dot(f, g), h_dot_g = dot(h, g);
    double df = 0.0;
    e = 0.0;
    if (dtype == 0 || dtype == 1) {    // dtype 1: the propensity series here, the bending term in cbt_eval below
        cosine_series(coeff, coeff + DIH_COLS, phi, e, df);
        const double* c_coil = coeff + 2 * DIH_COLS;
        const double* d_coil = coeff + 3 * DIH_COLS;
        bool c_any = false, d_any = false;
        for (int i = 0; i < DIH_COLS; ++i) {
            c_any |= (c_coil[i] != 0.0);
            d_any |= (d_coil[i] != 0.0);
        }
        if (c_any && d_any) cosine_series(c_coil, d_coil, phi, e, df);
    } else {   // dtype 2 (improper): coeff(1,1) = equilibrium, coeff(1,2) = force constant
        const double eq = coeff[0], fc = coeff[1];
        df = fc * (phi - eq);
        e = 0.5 * fc * (phi - eq) * (phi - eq);
    }
    const Vec3d sc = v * (f_dot_g / (v_sq * g_norm)) - w * (h_dot_g / (w_sq * g_norm));
    const Vec3d fa = v * (-df * g_norm / v_sq);
    const Vec3d fd = w * (df * g_norm / w_sq);
    out[0] = fa;
    out[1] = sc * df - fa;
    out[2] = sc * (-df) - fd;
    out[3] = fd;
}

// ---- combined bending-torsion dihedrals (dtype 1) and the backbone dipoles -------------------------
// compute_dihedral_forces.f90:77-112 + dipole_reconstruction.f90:50-221 (reconstruct): besides the propensity
// series, a dihedral a-b-c-d of dtype 1 carries V = 1/2 k(phi) (gamma - gamma_0(phi))^2 on the angle gamma = a-b-c
// (and, for the last dihedral of a backbone, on b-c-d as well), with k(phi) a cosine series (coefficient rows 4, 5)
// and gamma_0(phi) = 1.85 - 0.227 cos(phi - 0.785).  The bending term acts on the three beads of its angle AND adds
// dV/dphi to the dihedral force.  The reconstructed dipole (two charges 0.3 nm apart, centred on the b-c bond) and
// its transfer matrices follow the Fortran literally, including its sign convention for the gamma-derivative terms
// (oracle/bonded_oracle.py::reconstruct) and its single-precision constants cos(1.392947), sin(1.392947), 0.1.
struct CbtGeom {
    Vec3d w, v, dga, dgb, dgc;      // unit vectors of b->a and b->c; d gamma / d r_a, r_b, r_c
    double norm_a, norm_c, cos_gamma, sin_gamma, gamm;
    double energy, df_cbt, df_ang;
};

// false for collinear bonds (cos^2 gamma >= 1): the Fortran then leaves every output untouched
__host__ __device__ inline bool cbt_angle(Vec3d rab, Vec3d rcb, const double* __restrict__ c_k,
                                          const double* __restrict__ d_k, double phi, CbtGeom& q) {
    double k = 0.0, dk = 0.0;
    cosine_series(c_k, d_k, phi, k, dk);
    const double gamma_0 = 1.85 - 0.227 * cos(phi - 0.785);
    const double dg = 0.227 * sin(phi - 0.785);
    q.norm_a = sqrt(dot(rab, rab));
    q.norm_c = sqrt(dot(rcb, rcb));
    q.w = {rab.x / q.norm_a, rab.y / q.norm_a, rab.z / q.norm_a};
    q.v = {rcb.x / q.norm_c, rcb.y / q.norm_c, rcb.z / q.norm_c};
    q.cos_gamma = dot(q.w, q.v);
    const double cos2 = q.cos_gamma * q.cos_gamma;
    if (!(cos2 < 1.0)) return false;
    q.gamm = acos(q.cos_gamma);
    q.sin_gamma = sqrt(1.0 - cos2);
    if (q.sin_gamma < 0.1) q.sin_gamma = 0.10000000149011612;      // "sin_gamma = 0.1", a default-real literal
    const Vec3d fa = (q.v - q.w * q.cos_gamma), fc = (q.w - q.v * q.cos_gamma);
    q.dga = {-(fa.x / q.norm_a) / q.sin_gamma, -(fa.y / q.norm_a) / q.sin_gamma, -(fa.z / q.norm_a) / q.sin_gamma};
    q.dgc = {-(fc.x / q.norm_c) / q.sin_gamma, -(fc.y / q.norm_c) / q.sin_gamma, -(fc.z / q.norm_c) / q.sin_gamma};
    q.dgb = {-(q.dga.x + q.dgc.x), -(q.dga.y + q.dgc.y), -(q.dga.z + q.dgc.z)};
    q.df_ang = k * (q.gamm - gamma_0);
    const double var_sq = (q.gamm - gamma_0) * (q.gamm - gamma_0);
    q.energy = 0.5 * k * var_sq;
    q.df_cbt = 0.5 * dk * var_sq - q.df_ang * dg;
    return true;
}

// what the bending term ADDS to the force of the particle in each slot (on top of dihedral_eval's propensity part),
// and its energy; last != 0: the angle b-c-d is treated as well
template <typename P>
__host__ __device__ inline void cbt_eval(const P& pos, Vec3d box, int ia, int ib, int ic, int id,
                                         const double* __restrict__ coeff, int last, Vec3d* out, double& e) {
    const Vec3d f = mic_diff(pos, (long long)ia, (long long)ib, box);
    const Vec3d g = mic_diff(pos, (long long)ib, (long long)ic, box);
    const Vec3d h = mic_diff(pos, (long long)id, (long long)ic, box);
    const Vec3d v = cross(f, g), w = cross(h, g);
    const double v_sq = dot(v, v), w_sq = dot(w, w);
    const double g_norm = sqrt(dot(g, g));
    const double phi = atan2(dot(w, f) * g_norm, dot(v, w));
    const double f_dot_g = dot(f, g), h_dot_g = dot(h, g);
    const double* c_k = coeff + 4 * DIH_COLS;
    const double* d_k = coeff + 5 * DIH_COLS;
    const Vec3d zero = {0.0, 0.0, 0.0};
    out[0] = out[1] = out[2] = out[3] = zero;
    double df = 0.0;
    e = 0.0;
    CbtGeom q;
    if (cbt_angle(f, zero - g, c_k, d_k, phi, q)) {
        e += q.energy;
        df += q.df_cbt;
        out[0] = out[0] - q.dga * q.df_ang;
        out[1] = out[1] - q.dgb * q.df_ang;
        out[2] = out[2] - q.dgc * q.df_ang;
    }
    if (last && cbt_angle(g, h, c_k, d_k, phi, q)) {
        e += q.energy;
        df += q.df_cbt;
        out[1] = out[1] - q.dga * q.df_ang;
        out[2] = out[2] - q.dgb * q.df_ang;
        out[3] = out[3] - q.dgc * q.df_ang;
    }
    const Vec3d sc = v * (f_dot_g / (v_sq * g_norm)) - w * (h_dot_g / (w_sq * g_norm));
    const Vec3d fa = v * (-df * g_norm / v_sq);
    const Vec3d fd = w * (df * g_norm / w_sq);
    out[0] = out[0] + fa;
    out[1] = out[1] + (sc * df - fa);
    out[2] = out[2] + (sc * (-df) - fd);
    out[3] = out[3] + fd;
}

struct Mat3d {
    double m[3][3];
};
__host__ __device__ inline Vec3d row(const Mat3d& a, int i) { return {a.m[i][0], a.m[i][1], a.m[i][2]}; }
__host__ __device__ inline void set_row(Mat3d& a, int i, Vec3d r) { a.m[i][0] = r.x; a.m[i][1] = r.y; a.m[i][2] = r.z; }
__host__ __device__ inline double comp(Vec3d a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
// row i of the result = row i of a x vec (dipole_reconstruction.f90:14-24)
__host__ __device__ inline Mat3d cross_matrix(const Mat3d& a, Vec3d vec) {
    Mat3d o;
    for (int i = 0; i < 3; ++i) set_row(o, i, cross(row(a, i), vec));
    return o;
}
// row i = a_i * b (dipole_reconstruction.f90:26-35)
__host__ __device__ inline Mat3d outer(Vec3d a, Vec3d b) {
    Mat3d o;
    for (int i = 0; i < 3; ++i) set_row(o, i, b * comp(a, i));
    return o;
}
__host__ __device__ inline Mat3d axpby(double x, const Mat3d& a, double y, const Mat3d& b) {
    Mat3d o;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) o.m[i][j] = x * a.m[i][j] + y * b.m[i][j];
    return o;
}

// The dipole of the angle a-b-c (rab = r_a - r_b, rcb = r_c - r_b, rb = r_b): the two charge positions (wrapped,
// rounded to the position type like the Fortran's real(4) array) and the three transfer matrices D_a, D_b, D_c.
// Returns false (outputs untouched) for collinear bonds.
template <typename real>
__host__ __device__ inline bool cbt_dipole(Vec3d rab, Vec3d rb, Vec3d rcb, Vec3d box, const double* __restrict__ c_k,
                                           const double* __restrict__ d_k, double phi, real* __restrict__ dipole,
                                           real* __restrict__ transfer) {
    CbtGeom q;
    if (!cbt_angle(rab, rcb, c_k, d_k, phi, q)) return false;
    const double delta = 0.3, cos_psi = 0.1769132763147354, sin_psi = 0.9842264652252197;   // cos / sin(1.392947) in real(4)
    const double fac = exp((q.gamm - 1.73) / 0.025);
    const double theta = -1.607 * q.gamm + 0.094 + 1.883 / (1.0 + fac);
    const double d_theta = -1.607 - 1.883 / 0.025 * fac / ((1.0 + fac) * (1.0 + fac));
    const double cos_theta = cos(theta), sin_theta = sin(theta);
    const Vec3d wxv = cross(q.w, q.v);
    const Vec3d n = {wxv.x / q.sin_gamma, wxv.y / q.sin_gamma, wxv.z / q.sin_gamma};
    const Vec3d m = cross(n, q.v);
    const Vec3d r0 = rb + rcb * 0.5;
    const Vec3d d = (q.v * cos_psi + (n * cos_theta + m * sin_theta) * sin_psi) * (0.5 * delta);
    const double bx[3] = {box.x, box.y, box.z};
    for (int s = 0; s < 2; ++s) {
        const Vec3d p = s == 0 ? r0 + d : r0 - d;
        for (int k = 0; k < 3; ++k) {
            const double x = (double)(real)comp(p, k);                  // dipole(s,:) is real(4)
            const double qn = x / bx[k];
            const double nint = qn >= 0.0 ? floor(qn + 0.5) : -floor(-qn + 0.5);
            dipole[3 * s + k] = (real)(x - bx[k] * nint);
        }
    }
    Mat3d V_b, W_b;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            V_b.m[i][j] = (comp(q.v, i) * comp(q.v, j) - (i == j ? 1.0 : 0.0)) / q.norm_c;
            W_b.m[i][j] = (comp(q.w, i) * comp(q.w, j) - (i == j ? 1.0 : 0.0)) / q.norm_a;
        }
    const Mat3d Z = {{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}};
    const Mat3d V_c = axpby(-1.0, V_b, 0.0, Z), W_a = axpby(-1.0, W_b, 0.0, Z);
    // N_i, M_i as written in the Fortran (lines 186-192)
    Mat3d N_a = axpby(q.cos_gamma, outer(q.dga, n), 1.0, cross_matrix(W_a, q.v));
    Mat3d N_b = axpby(1.0, axpby(q.cos_gamma, outer(q.dgb, n), 1.0, cross_matrix(W_b, q.v)), -1.0, cross_matrix(V_b, q.w));
    Mat3d N_c = axpby(q.cos_gamma, outer(q.dgc, n), -1.0, cross_matrix(V_c, q.w));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            N_a.m[i][j] = N_a.m[i][j] / q.sin_gamma; N_b.m[i][j] = N_b.m[i][j] / q.sin_gamma; N_c.m[i][j] = N_c.m[i][j] / q.sin_gamma;
        }
    const Mat3d M_a = cross_matrix(N_a, q.v);
    const Mat3d M_b = axpby(1.0, cross_matrix(N_b, q.v), -1.0, cross_matrix(V_b, n));
    const Mat3d M_c = axpby(1.0, cross_matrix(N_c, q.v), -1.0, cross_matrix(V_c, n));
    const Vec3d dg[3] = {q.dga, q.dgb, q.dgc};
    const Mat3d* Nm[3] = {&N_a, &N_b, &N_c};
    const Mat3d* Mm[3] = {&M_a, &M_b, &M_c};
    const Mat3d* Vm[3] = {&Z, &V_b, &V_c};
    for (int t = 0; t < 3; ++t) {
        const Mat3d FN = outer(dg[t], n), FM = outer(dg[t], m);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                const double inner = cos_theta * Nm[t]->m[i][j] + sin_theta * Mm[t]->m[i][j] +
                                     sin_theta * d_theta * FN.m[i][j] - cos_theta * d_theta * FM.m[i][j];
                transfer[9 * t + 3 * i + j] = (real)(0.5 * delta * (cos_psi * Vm[t]->m[i][j] + sin_psi * inner));
            }
    }
    return true;
}

// dipoles (4,3) and transfer matrices (6,3,3) of one dihedral (zero unless dtype 1; rows 2-3 / matrices 3-5 only for
// the last dihedral of a backbone): compute_dihedral_forces.f90:27-28, 84-111 with dipole_flag = 1
template <typename real>
__host__ __device__ inline void dipole_term(const real* __restrict__ pos, Vec3d box, int ia, int ib, int ic, int id,
                                            const double* __restrict__ coeff, int dtype, int last,
                                            real* __restrict__ dipoles, real* __restrict__ transfer) {
    for (int k = 0; k < 12; ++k) dipoles[k] = (real)0;
    for (int k = 0; k < 54; ++k) transfer[k] = (real)0;
    if (dtype != 1) return;
    const Vec3d f = mic_diff(pos, (long long)ia, (long long)ib, box);
    const Vec3d g = mic_diff(pos, (long long)ib, (long long)ic, box);
    const Vec3d h = mic_diff(pos, (long long)id, (long long)ic, box);
    const Vec3d v = cross(f, g), w = cross(h, g);
    const double g_norm = sqrt(dot(g, g));
    const double phi = atan2(dot(w, f) * g_norm, dot(v, w));
    const double* c_k = coeff + 4 * DIH_COLS;
    const double* d_k = coeff + 5 * DIH_COLS;
    const Vec3d zero = {0.0, 0.0, 0.0};
    const Vec3d rb = {(double)pos[3 * ib + 0], (double)pos[3 * ib + 1], (double)pos[3 * ib + 2]};
    cbt_dipole<real>(f, rb, zero - g, box, c_k, d_k, phi, dipoles, transfer);
    if (last) {
        const Vec3d rc = {(double)pos[3 * ic + 0], (double)pos[3 * ic + 1], (double)pos[3 * ic + 2]};
        cbt_dipole<real>(g, rc, h, box, c_k, d_k, phi, dipoles + 6, transfer + 27);
    }
}

// dipole_forces_redistribution (hymd/force.py:855-880): what the forces fd (4,3) on the dipole charges of ONE
// dihedral add to the bead in `slot` of that dihedral (D = its six transfer matrices; f += D_i @ (f+ - f-), the
// two beads of the bond carrying the dipole also take half of f+ + f-)
template <typename real>
__host__ __device__ inline Vec3d redistribute_term(int slot, int last, const real* __restrict__ fd,
                                                   const real* __restrict__ D) {
    Vec3d out = {0.0, 0.0, 0.0};
    auto mv = [&](int mat, Vec3d x) {
        const real* M = D + 9 * mat;
        return Vec3d{(double)M[0] * x.x + (double)M[1] * x.y + (double)M[2] * x.z,
                     (double)M[3] * x.x + (double)M[4] * x.y + (double)M[5] * x.z,
                     (double)M[6] * x.x + (double)M[7] * x.y + (double)M[8] * x.z};
    };
    const Vec3d f0 = {(double)fd[0], (double)fd[1], (double)fd[2]}, f1 = {(double)fd[3], (double)fd[4], (double)fd[5]};
    const Vec3d s01 = f0 + f1, d01 = f0 - f1;
    if (slot == 0) out = mv(0, d01);
    else if (slot == 1) out = mv(1, d01) + s01 * 0.5;
    else if (slot == 2) out = mv(2, d01) + s01 * 0.5;
    if (last) {
        const Vec3d f2 = {(double)fd[6], (double)fd[7], (double)fd[8]}, f3 = {(double)fd[9], (double)fd[10], (double)fd[11]};
        const Vec3d s23 = f2 + f3, d23 = f2 - f3;
        if (slot == 1) out = out + mv(3, d23);
        else if (slot == 2) out = out + mv(4, d23) + s23 * 0.5;
        else if (slot == 3) out = mv(5, d23) + s23 * 0.5;
    }
    return out;
}

template <typename real>
__host__ __device__ inline void dihedral_term(const real* __restrict__ pos, Vec3d box, int ia, int ib, int ic,
                                              int id, const double* __restrict__ coeff, int dtype, int slot,
                                              BondAcc& acc) {
    Vec3d out[4];
    double e;
    dihedral_eval(pos, box, ia, ib, ic, id, coeff, dtype, out, e);
    acc.f = acc.f + out[slot];
    if (slot == 0) acc.e += e;
}

// ---- per-particle term lists -------------------------------------------------------------------
// refs[start[p] .. start[p+1]) = term * 4 + slot for every (term, slot) with index[slot][term] == p,
// ascending in term.  Returns false if an index is out of range or a particle occurs twice in a term.
inline bool build_particle_csr(long long n_particles, long long n_terms, int n_slots,
                               const int32_t* const* index, std::vector<uint32_t>& start,
                               std::vector<uint32_t>& refs) {
    start.assign((size_t)n_particles + 1, 0u);
    if (n_terms >= (1LL << 30)) return false;
    for (long long t = 0; t < n_terms; ++t)
        for (int s = 0; s < n_slots; ++s) {
            const long long p = index[s][t];
            if (p < 0 || p >= n_particles) return false;
            for (int s2 = 0; s2 < s; ++s2)
                if (index[s2][t] == p) return false;
            start[(size_t)p + 1]++;
        }
    for (long long p = 0; p < n_particles; ++p) start[(size_t)p + 1] += start[(size_t)p];
    refs.assign((size_t)n_terms * n_slots, 0u);
    std::vector<uint32_t> cur(start.begin(), start.end() - 1);
    for (long long t = 0; t < n_terms; ++t)
        for (int s = 0; s < n_slots; ++s) refs[cur[(size_t)index[s][t]]++] = (uint32_t)(t * 4 + s);
    return true;
}

// One particle's bonded forces of one kind (KIND = 2, 3, 4 particles per term).
template <typename real, int KIND>
__host__ __device__ inline BondAcc particle_terms(long long p, const real* __restrict__ pos, Vec3d box,
                                                  const uint32_t* __restrict__ start,
                                                  const uint32_t* __restrict__ refs,
                                                  const int32_t* __restrict__ idx,     // [term][4]
                                                  const double* __restrict__ par,      // [term][2] or [term][30]
                                                  const int32_t* __restrict__ dtype) {
    BondAcc acc = {{0.0, 0.0, 0.0}, 0.0, {0.0, 0.0, 0.0}};
    for (uint32_t r = start[p]; r < start[p + 1]; ++r) {
        const uint32_t ref = refs[r];
        const long long t = ref >> 2;
        const int slot = (int)(ref & 3u);
        const int32_t* ix = idx + 4 * t;
        if (KIND == 2)
            bond_term(pos, box, ix[0], ix[1], par[2 * t], par[2 * t + 1], slot, acc);
        else if (KIND == 3)
            angle_term(pos, box, ix[0], ix[1], ix[2], par[2 * t], par[2 * t + 1], slot, acc);
        else
            dihedral_term(pos, box, ix[0], ix[1], ix[2], ix[3], par + (long long)DIH_ROWS * DIH_COLS * t,
                          dtype[t], slot, acc);
    }
    return acc;
}

// All three kinds of one rank's molecules (device or host pointers).
struct TermLists {
    const uint32_t* start[3];
    const uint32_t* refs[3];
    const int32_t* idx[3];
    const double* par[3];
    const int32_t* dih_type;
    const int32_t* dih_last;     // bonds_4_last (only read for dih_type 1)
    long long n_terms[3];
};

// Tail of a fused step for one particle: round each kind's force to the array type (the Fortran's f
// arrays are real(4) in the default build), optional per-kind output, kick(s), drift + wrap.
template <typename real>
__host__ __device__ inline void finish_particle(long long p, const real* __restrict__ x_in,
                                                real* __restrict__ x_out, real* __restrict__ vel, Vec3d box,
                                                real mass, real half_dt, int n_kicks, real dt,
                                                real* const* f_out, const BondAcc* acc) {
    const real L[3] = {(real)box.x, (real)box.y, (real)box.z};
    for (int d = 0; d < 3; ++d) {
        real ft[3];
        for (int k = 0; k < 3; ++k) {
            const double fk = d == 0 ? acc[k].f.x : (d == 1 ? acc[k].f.y : acc[k].f.z);
            ft[k] = (real)fk;
            if (f_out != nullptr && f_out[k] != nullptr) f_out[k][3 * p + d] = ft[k];
        }
        if (vel == nullptr) continue;
        real v = vel[3 * p + d];
        for (int r = 0; r < n_kicks; ++r) v = kick(v, ft, 3, mass, half_dt);
        if (n_kicks > 0) vel[3 * p + d] = v;
        if (x_out != nullptr) x_out[3 * p + d] = drift_wrap(x_in[3 * p + d], v, dt, L[d]);
    }
}

// One particle's share of a fused inner rRESPA step (main.py:829-893):
//   F = bonded forces at x_in (each kind rounded to the array type like the Fortran's f arrays),
//   n_kicks x  v += half_dt * (f_bond + f_angle + f_dihedral) / mass   (closing kick of the previous
//              inner step and opening kick of the next one: same forces, two roundings like the
//              reference's two integrate_velocity calls),
//   x_out = mod(x_in + dt * v, box)   if x_out != nullptr (double-buffered: other threads still read x_in).
// acc[k] returns the energy / pressure by-products of the terms this particle owns.
template <typename real>
__host__ __device__ inline void inner_step_particle(long long p, const real* __restrict__ x_in,
                                                    real* __restrict__ x_out, real* __restrict__ vel,
                                                    Vec3d box, const TermLists& t, real mass, real half_dt,
                                                    int n_kicks, real dt, real* const* f_out, BondAcc* acc) {
    const BondAcc zero = {{0.0, 0.0, 0.0}, 0.0, {0.0, 0.0, 0.0}};
    acc[0] = t.n_terms[0] ? particle_terms<real, 2>(p, x_in, box, t.start[0], t.refs[0], t.idx[0], t.par[0], nullptr) : zero;
    acc[1] = t.n_terms[1] ? particle_terms<real, 3>(p, x_in, box, t.start[1], t.refs[1], t.idx[1], t.par[1], nullptr) : zero;
    acc[2] = t.n_terms[2] ? particle_terms<real, 4>(p, x_in, box, t.start[2], t.refs[2], t.idx[2], t.par[2], t.dih_type) : zero;
    finish_particle<real>(p, x_in, x_out, vel, box, mass, half_dt, n_kicks, dt, f_out, acc);
}

// The bending terms of one particle's dtype-1 dihedrals (what cbt_kernel / the CBT variant of the fused step add)
template <typename real>
__host__ __device__ inline void particle_cbt(long long p, const real* __restrict__ pos, Vec3d box, const TermLists& t,
                                             BondAcc& acc) {
    for (uint32_t r = t.start[2][p]; r < t.start[2][p + 1]; ++r) {
        const uint32_t ref = t.refs[2][r];
        const long long term = ref >> 2;
        if (t.dih_type[term] != 1) continue;
        const int32_t* ix = t.idx[2] + 4 * term;
        Vec3d out[4];
        double e;
        cbt_eval(pos, box, ix[0], ix[1], ix[2], ix[3], t.par[2] + (long long)DIH_ROWS * DIH_COLS * term, t.dih_last[term],
                 out, e);
        acc.f = acc.f + out[ref & 3u];
        if ((ref & 3u) == 0) acc.e += e;
    }
}

// inner_step_particle for topologies with dtype-1 dihedrals: the bending terms join the dihedral kind before the
// rounding to the array type (a separate instantiation: the common kernels do not carry this code)
template <typename real>
__host__ __device__ inline void inner_step_particle_cbt(long long p, const real* __restrict__ x_in,
                                                        real* __restrict__ x_out, real* __restrict__ vel,
                                                        Vec3d box, const TermLists& t, real mass, real half_dt,
                                                        int n_kicks, real dt, real* const* f_out, BondAcc* acc) {
    const BondAcc zero = {{0.0, 0.0, 0.0}, 0.0, {0.0, 0.0, 0.0}};
    acc[0] = t.n_terms[0] ? particle_terms<real, 2>(p, x_in, box, t.start[0], t.refs[0], t.idx[0], t.par[0], nullptr) : zero;
    acc[1] = t.n_terms[1] ? particle_terms<real, 3>(p, x_in, box, t.start[1], t.refs[1], t.idx[1], t.par[1], nullptr) : zero;
    acc[2] = zero;
    if (t.n_terms[2]) {
        acc[2] = particle_terms<real, 4>(p, x_in, box, t.start[2], t.refs[2], t.idx[2], t.par[2], t.dih_type);
        particle_cbt<real>(p, x_in, box, t, acc[2]);
    }
    finish_particle<real>(p, x_in, x_out, vel, box, mass, half_dt, n_kicks, dt, f_out, acc);
}

// ---- CTA-cooperative evaluation ----------------------------------------------------------------
// The per-particle path evaluates every term once per participant (2x, 3x, 4x).  Here a CTA of
// `cta_size` consecutive particles first evaluates every term that touches it ONCE (its threads
// stride over the CTA's term list) into shared memory -- 3 / 6 / 12 doubles per bond / angle /
// dihedral -- and every particle then sums its own (term, slot) references in the same ascending
// term order with the same additions as the per-particle path, so the forces are bitwise identical.
// Terms of molecules that straddle a CTA boundary are evaluated by each CTA they touch.
struct CtaLists {
    const uint32_t* cta_start[3];   // [n_cta + 1]
    const uint32_t* cta_terms[3];   // term ids per CTA, ascending
    const uint32_t* lrefs[3];       // parallel to TermLists::refs: (position in the CTA list << 2) | slot
    int max_terms[3];               // largest CTA list per kind (sizes the shared memory)
};
constexpr int CTA_DOUBLES[3] = {3, 6, 12};

inline void build_cta_lists(long long n_particles, long long n_terms, int n_slots, const int32_t* const* index,
                            int cta_size, const std::vector<uint32_t>& start, std::vector<uint32_t>& cta_start,
                            std::vector<uint32_t>& cta_terms, std::vector<uint32_t>& lrefs, int& max_terms) {
    const long long n_cta = (n_particles + cta_size - 1) / cta_size;
    cta_start.assign((size_t)n_cta + 1, 0u);
    auto distinct = [&](long long t, long long* c) {
        int m = 0;
        for (int s = 0; s < n_slots; ++s) {
            const long long cs = index[s][t] / cta_size;
            bool seen = false;
            for (int j = 0; j < m; ++j) seen |= (c[j] == cs);
            if (!seen) c[m++] = cs;
        }
        return m;
    };
    long long c[4];
    for (long long t = 0; t < n_terms; ++t) {
        const int m = distinct(t, c);
        for (int j = 0; j < m; ++j) cta_start[(size_t)c[j] + 1]++;
    }
    max_terms = 0;
    for (long long i = 0; i < n_cta; ++i) {
        if ((int)cta_start[(size_t)i + 1] > max_terms) max_terms = (int)cta_start[(size_t)i + 1];
        cta_start[(size_t)i + 1] += cta_start[(size_t)i];
    }
    cta_terms.assign(cta_start[(size_t)n_cta], 0u);
    lrefs.assign((size_t)n_terms * n_slots, 0u);
    std::vector<uint32_t> cur(cta_start.begin(), cta_start.end() - 1);
    std::vector<uint32_t> pcur(start.begin(), start.end() - 1);
    for (long long t = 0; t < n_terms; ++t) {
        const int m = distinct(t, c);
        uint32_t lpos[4];
        for (int j = 0; j < m; ++j) {
            const uint32_t pos = cur[(size_t)c[j]]++;
            cta_terms[pos] = (uint32_t)t;
            lpos[j] = pos - cta_start[(size_t)c[j]];
        }
        // refs of a particle are filled in (term, slot) ascending order: the next free one is (t, s)
        for (int s = 0; s < n_slots; ++s) {
            const long long p = index[s][t];
            int j = 0;
            while (c[j] != p / cta_size) ++j;
            lrefs[pcur[(size_t)p]++] = (lpos[j] << 2) | (uint32_t)s;
        }
    }
}

// Phase 1: thread `tid` of `nthreads` evaluates its share of CTA `cta`'s terms of all kinds into `sm`
// (layout: [bonds: max_terms[0]*3][angles: max_terms[1]*6][dihedrals: max_terms[2]*12] doubles) and
// accumulates energy / pressure of the terms whose slot-0 particle lies in [p0, p1) into own[12].
template <typename real>
__host__ __device__ inline void cta_eval_terms(int tid, int nthreads, long long cta, long long p0, long long p1,
                                               const real* __restrict__ x, Vec3d box, const TermLists& t,
                                               const CtaLists& c, double* __restrict__ sm, double* own) {
    double* sm2 = sm;
    double* sm3 = sm2 + (long long)c.max_terms[0] * CTA_DOUBLES[0];
    double* sm4 = sm3 + (long long)c.max_terms[1] * CTA_DOUBLES[1];
    // one flat work list over the three kinds, so that a CTA with few terms of each kind (chains and
    // solvent interleave in domain_decomposition order) still keeps all of its threads busy in one round
    const uint32_t b2 = t.n_terms[0] ? c.cta_start[0][cta] : 0u, n2 = t.n_terms[0] ? c.cta_start[0][cta + 1] - b2 : 0u;
    const uint32_t b3 = t.n_terms[1] ? c.cta_start[1][cta] : 0u, n3 = t.n_terms[1] ? c.cta_start[1][cta + 1] - b3 : 0u;
    const uint32_t b4 = t.n_terms[2] ? c.cta_start[2][cta] : 0u, n4 = t.n_terms[2] ? c.cta_start[2][cta + 1] - b4 : 0u;
    for (uint32_t j = (uint32_t)tid; j < n2 + n3 + n4; j += (uint32_t)nthreads) {
        if (j < n2) {
            const uint32_t i = j;
            const long long term = c.cta_terms[0][b2 + i];
            const int32_t* ix = t.idx[0] + 4 * term;
            Vec3d fa, pr;
            double en;
            bond_eval(x, box, ix[0], ix[1], t.par[0][2 * term], t.par[0][2 * term + 1], fa, en, pr);
            double* o = sm2 + (long long)i * 3;
            o[0] = fa.x; o[1] = fa.y; o[2] = fa.z;
            if (ix[0] >= p0 && ix[0] < p1) { own[0] += en; own[1] += pr.x; own[2] += pr.y; own[3] += pr.z; }
        } else if (j < n2 + n3) {
            const uint32_t i = j - n2;
            const long long term = c.cta_terms[1][b3 + i];
            const int32_t* ix = t.idx[1] + 4 * term;
            Vec3d fa = {0.0, 0.0, 0.0}, fc = {0.0, 0.0, 0.0}, pr = {0.0, 0.0, 0.0};
            double en = 0.0;
            const bool ok = angle_eval(x, box, ix[0], ix[1], ix[2], t.par[1][2 * term], t.par[1][2 * term + 1],
                                       fa, fc, en, pr);
            double* o = sm3 + (long long)i * 6;
            // an invalid angle is marked with a NaN in the first slot: the particle phase skips it,
            // exactly like the per-particle path skips the additions
            o[0] = ok ? fa.x : nan(""); o[1] = fa.y; o[2] = fa.z; o[3] = fc.x; o[4] = fc.y; o[5] = fc.z;
            if (ok && ix[0] >= p0 && ix[0] < p1) { own[4] += en; own[5] += pr.x; own[6] += pr.y; own[7] += pr.z; }
        } else {
            const uint32_t i = j - n2 - n3;
            const long long term = c.cta_terms[2][b4 + i];
            const int32_t* ix = t.idx[2] + 4 * term;
            Vec3d out[4];
            double en;
            dihedral_eval(x, box, ix[0], ix[1], ix[2], ix[3], t.par[2] + (long long)DIH_ROWS * DIH_COLS * term,
                          t.dih_type[term], out, en);
            double* o = sm4 + (long long)i * 12;
            for (int s = 0; s < 4; ++s) { o[3 * s] = out[s].x; o[3 * s + 1] = out[s].y; o[3 * s + 2] = out[s].z; }
            if (ix[0] >= p0 && ix[0] < p1) own[8] += en;
        }
    }
}

// ---- CTA-cooperative evaluation, second layout ("mode 2") ----------------------------------------
// Cuts the dependent global loads of phase 1 from four levels (list bounds -> term id -> indices and
// parameters -> positions) to two: the bond / angle records {indices, parameters} are stored INLINE in
// the CTA's list (32 B, one coalesced load per thread) and the CTA's own positions are staged in
// shared memory first, so a term only goes back to global memory for atoms outside the CTA.
struct TermRec {
    int32_t i[4];
    double p[2];
};
struct CtaRecs {
    const TermRec* rec[2];      // bonds, angles: parallel to CtaLists::cta_terms[0], [1]
};

inline void build_cta_records(const std::vector<uint32_t>& cta_terms, const int32_t* idx4, const double* par2,
                              std::vector<TermRec>& rec) {
    rec.resize(cta_terms.size());
    for (size_t i = 0; i < cta_terms.size(); ++i) {
        const size_t t = cta_terms[i];
        for (int s = 0; s < 4; ++s) rec[i].i[s] = idx4[4 * t + s];
        rec[i].p[0] = par2[2 * t];
        rec[i].p[1] = par2[2 * t + 1];
    }
}

// P = PosTile<real> (mode 2: own positions staged in shared memory) or const real* (mode 3: inline
// records only, positions straight from global memory / L2, no staging barrier).
template <typename real, typename P>
__host__ __device__ inline void cta2_eval_terms(int tid, int nthreads, long long cta, long long p0, long long p1,
                                                const P& x, Vec3d box, const TermLists& t,
                                                const CtaLists& c, const CtaRecs& rc, double* __restrict__ sm,
                                                double* own) {
    double* sm2 = sm;
    double* sm3 = sm2 + (long long)c.max_terms[0] * CTA_DOUBLES[0];
    double* sm4 = sm3 + (long long)c.max_terms[1] * CTA_DOUBLES[1];
    const uint32_t b2 = t.n_terms[0] ? c.cta_start[0][cta] : 0u, n2 = t.n_terms[0] ? c.cta_start[0][cta + 1] - b2 : 0u;
    const uint32_t b3 = t.n_terms[1] ? c.cta_start[1][cta] : 0u, n3 = t.n_terms[1] ? c.cta_start[1][cta + 1] - b3 : 0u;
    const uint32_t b4 = t.n_terms[2] ? c.cta_start[2][cta] : 0u, n4 = t.n_terms[2] ? c.cta_start[2][cta + 1] - b4 : 0u;
    for (uint32_t j = (uint32_t)tid; j < n2 + n3 + n4; j += (uint32_t)nthreads) {
        if (j < n2) {
            const uint32_t i = j;
            const TermRec r = rc.rec[0][b2 + i];
            Vec3d fa, pr;
            double en;
            bond_eval(x, box, r.i[0], r.i[1], r.p[0], r.p[1], fa, en, pr);
            double* o = sm2 + (long long)i * 3;
            o[0] = fa.x; o[1] = fa.y; o[2] = fa.z;
            if (r.i[0] >= p0 && r.i[0] < p1) { own[0] += en; own[1] += pr.x; own[2] += pr.y; own[3] += pr.z; }
        } else if (j < n2 + n3) {
            const uint32_t i = j - n2;
            const TermRec r = rc.rec[1][b3 + i];
            Vec3d fa = {0.0, 0.0, 0.0}, fc = {0.0, 0.0, 0.0}, pr = {0.0, 0.0, 0.0};
            double en = 0.0;
            const bool ok = angle_eval(x, box, r.i[0], r.i[1], r.i[2], r.p[0], r.p[1], fa, fc, en, pr);
            double* o = sm3 + (long long)i * 6;
            o[0] = ok ? fa.x : nan(""); o[1] = fa.y; o[2] = fa.z; o[3] = fc.x; o[4] = fc.y; o[5] = fc.z;
            if (ok && r.i[0] >= p0 && r.i[0] < p1) { own[4] += en; own[5] += pr.x; own[6] += pr.y; own[7] += pr.z; }
        } else {
            const uint32_t i = j - n2 - n3;
            const long long term = c.cta_terms[2][b4 + i];
            const int32_t* ix = t.idx[2] + 4 * term;
            Vec3d out[4];
            double en;
            dihedral_eval(x, box, ix[0], ix[1], ix[2], ix[3], t.par[2] + (long long)DIH_ROWS * DIH_COLS * term,
                          t.dih_type[term], out, en);
            double* o = sm4 + (long long)i * 12;
            for (int s = 0; s < 4; ++s) { o[3 * s] = out[s].x; o[3 * s + 1] = out[s].y; o[3 * s + 2] = out[s].z; }
            if (ix[0] >= p0 && ix[0] < p1) own[8] += en;
        }
    }
}

// Phase 2: particle p sums its references out of shared memory (same order and additions as
// particle_terms); acc[k].e / .pr stay zero (phase 1 accounts for them).  The list bounds can be
// fetched ahead of the barrier (RefBounds).
struct RefBounds {
    uint32_t b[3], e[3];
};
__host__ __device__ inline RefBounds ref_bounds(long long p, const TermLists& t) {
    RefBounds r;
    for (int k = 0; k < 3; ++k) {
        r.b[k] = t.n_terms[k] ? t.start[k][p] : 0u;
        r.e[k] = t.n_terms[k] ? t.start[k][p + 1] : 0u;
    }
    return r;
}

__host__ __device__ inline void cta_gather_bounds(const RefBounds& rb, const CtaLists& c,
                                                  const double* __restrict__ sm, BondAcc* acc) {
    const BondAcc zero = {{0.0, 0.0, 0.0}, 0.0, {0.0, 0.0, 0.0}};
    const double* sm2 = sm;
    const double* sm3 = sm2 + (long long)c.max_terms[0] * CTA_DOUBLES[0];
    const double* sm4 = sm3 + (long long)c.max_terms[1] * CTA_DOUBLES[1];
    acc[0] = acc[1] = acc[2] = zero;
    for (uint32_t r = rb.b[0]; r < rb.e[0]; ++r) {
        const uint32_t lr = c.lrefs[0][r];
        const double* o = sm2 + (long long)(lr >> 2) * 3;
        const Vec3d fa = {o[0], o[1], o[2]};
        bond_apply((int)(lr & 3u), fa, acc[0].f);
    }
    for (uint32_t r = rb.b[1]; r < rb.e[1]; ++r) {
        const uint32_t lr = c.lrefs[1][r];
        const double* o = sm3 + (long long)(lr >> 2) * 6;
        if (o[0] != o[0]) continue;
        const Vec3d fa = {o[0], o[1], o[2]}, fc = {o[3], o[4], o[5]};
        angle_apply((int)(lr & 3u), fa, fc, acc[1].f);
    }
    for (uint32_t r = rb.b[2]; r < rb.e[2]; ++r) {
        const uint32_t lr = c.lrefs[2][r];
        const double* o = sm4 + (long long)(lr >> 2) * 12 + 3 * (lr & 3u);
        const Vec3d add = {o[0], o[1], o[2]};
        acc[2].f = acc[2].f + add;
    }
}

__host__ __device__ inline void cta_gather_particle(long long p, const TermLists& t, const CtaLists& c,
                                                    const double* __restrict__ sm, BondAcc* acc) {
    cta_gather_bounds(ref_bounds(p, t), c, sm, acc);
}

}  // namespace hymd
