// Per-particle evaluation of HyMD's intramolecular forces (bonds, angles, dihedrals).
//
// Reference kernels (Fortran, sequential over terms, read-modify-write of f per term):
//   hymd/compute_bond_forces.f90:1-61        cbf
//   hymd/compute_angle_forces.f90:1-93       caf
//   hymd/compute_dihedral_forces.f90:1-137   cdf   (dtype 0: cosine series, 2: improper)
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
    if (dtype == 0) {
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
