// Single-precision evaluation of bonds and angles for the fp32 build (opt-in: hymd_bonded_set_math(b, 1)
// or HYMD_B200_BONDED_F32MATH=1; per-particle fused inner step only).
//
// The default path (bonded.cuh) follows the Fortran: float positions, double arithmetic.  It is bound by
// the fp64 instruction rate (angles: ~250 fp64 instructions per evaluation).  This
// variant does the arithmetic in float with formulas that do not lose accuracy where the textbook ones
// do in single precision:
//   * theta = atan2(|ea x ec|, ea . ec) instead of acos(ea . ec)      (acos is ill-conditioned at 0 / pi,
//     where the 180-degree equilibrium angles of HyMD's lipid models live),
//   * sin(theta) = |ea x ec| directly,
//   * ec - (ea.ec) ea = (ea x ec) x ea  and  ea - (ea.ec) ec = ec x (ea x ec): no cancellation.
// Accuracy against the float64 oracle on float32 positions: <= 1e-5 of the largest force (the north_star
// tolerance of the fp32 build), tests/test_native_host_check.py::test_f32_math_accuracy.  Dihedrals keep
// the double evaluator.  __host__ __device__ like bonded.cuh.
#pragma once
#include "bonded.cuh"

namespace hymd {

struct Vec3f {
    float x, y, z;
};
__host__ __device__ inline Vec3f operator+(Vec3f a, Vec3f b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__host__ __device__ inline Vec3f operator-(Vec3f a, Vec3f b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__host__ __device__ inline Vec3f operator*(Vec3f a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__host__ __device__ inline float dotf(Vec3f a, Vec3f b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ inline Vec3f crossf(Vec3f a, Vec3f b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__host__ __device__ inline Vec3f mic_diff_f(const float* __restrict__ pos, long long i, long long j, Vec3f box) {
    Vec3f d = {pos[3 * i + 0] - pos[3 * j + 0], pos[3 * i + 1] - pos[3 * j + 1], pos[3 * i + 2] - pos[3 * j + 2]};
    d.x -= box.x * roundf(d.x / box.x);
    d.y -= box.y * roundf(d.y / box.y);
    d.z -= box.z * roundf(d.z / box.z);
    return d;
}

struct BondAccF {
    Vec3f f;
    double e;        // energies / pressure terms are summed in double (few per particle)
    Vec3d pr;
};

__host__ __device__ inline void bond_term_f(const float* __restrict__ pos, Vec3f box, int ia, int ib, float r0,
                                            float k, int slot, BondAccF& acc) {
    const Vec3f rab = mic_diff_f(pos, ib, ia, box);
    const float n = sqrtf(dotf(rab, rab));
    const float dn = n - r0;
    const Vec3f fa = rab * (-(k * dn) / n);
    if (slot == 0) {
        acc.f = acc.f - fa;
        acc.e += 0.5 * (double)k * (double)dn * (double)dn;
        acc.pr = acc.pr + Vec3d{(double)fa.x * rab.x, (double)fa.y * rab.y, (double)fa.z * rab.z};
    } else {
        acc.f = acc.f + fa;
    }
}

__host__ __device__ inline void angle_term_f(const float* __restrict__ pos, Vec3f box, int ia, int ib, int ic,
                                             float t0, float k, int slot, BondAccF& acc) {
    const Vec3f ra = mic_diff_f(pos, ia, ib, box);
    const Vec3f rc = mic_diff_f(pos, ic, ib, box);
    const float na = sqrtf(dotf(ra, ra)), nc = sqrtf(dotf(rc, rc));
    const Vec3f ea = ra * (1.0f / na), ec = rc * (1.0f / nc);
    const Vec3f cr = crossf(ea, ec);
    const float s = sqrtf(dotf(cr, cr));          // sin(theta) >= 0
    if (!(s > 0.0f)) return;                      // cos^2 == 1: the Fortran skips the term
    const float c = dotf(ea, ec);
    const float theta = atan2f(s, c);
    const float d = theta - t0;
    const float ff = k * d;
    const float xra = -ff / (na * s), xrc = -ff / (nc * s);
    const Vec3f fa = crossf(cr, ea) * xra;        // (ec - c ea) * xra
    const Vec3f fc = crossf(ec, cr) * xrc;        // (ea - c ec) * xrc
    if (slot == 0) {
        acc.f = acc.f - fa;
        acc.e += 0.5 * (double)ff * (double)d;
        acc.pr = acc.pr - Vec3d{(double)fa.x * ra.x, (double)fa.y * ra.y, (double)fa.z * ra.z} -
                 Vec3d{(double)fc.x * rc.x, (double)fc.y * rc.y, (double)fc.z * rc.z};
    } else if (slot == 2) {
        acc.f = acc.f - fc;
    } else {
        acc.f = acc.f + fa + fc;
    }
}

// One particle's bonds (KIND 2) or angles (KIND 3) with the float evaluators; same result layout as
// particle_terms (bonded.cuh).
template <int KIND>
__host__ __device__ inline BondAcc particle_terms_f32(long long p, const float* __restrict__ pos, Vec3d box,
                                                      const uint32_t* __restrict__ start,
                                                      const uint32_t* __restrict__ refs,
                                                      const int32_t* __restrict__ idx,
                                                      const double* __restrict__ par) {
    const Vec3f bf = {(float)box.x, (float)box.y, (float)box.z};
    BondAccF a = {{0.0f, 0.0f, 0.0f}, 0.0, {0.0, 0.0, 0.0}};
    for (uint32_t r = start[p]; r < start[p + 1]; ++r) {
        const uint32_t ref = refs[r];
        const long long term = ref >> 2;
        const int slot = (int)(ref & 3u);
        const int32_t* ix = idx + 4 * term;
        const float p0 = (float)par[2 * term], p1 = (float)par[2 * term + 1];
        if (KIND == 2) bond_term_f(pos, bf, ix[0], ix[1], p0, p1, slot, a);
        else angle_term_f(pos, bf, ix[0], ix[1], ix[2], p0, p1, slot, a);
    }
    BondAcc out;
    out.f = {(double)a.f.x, (double)a.f.y, (double)a.f.z};
    out.e = a.e;
    out.pr = a.pr;
    return out;
}

// Same contract as inner_step_particle<float> (bonded.cuh) with the float evaluators for bonds and angles.
__host__ __device__ inline void inner_step_particle_f32(long long p, const float* __restrict__ x_in,
                                                        float* __restrict__ x_out, float* __restrict__ vel,
                                                        Vec3d box, const TermLists& t, float mass, float half_dt,
                                                        int n_kicks, float dt, float* const* f_out, BondAcc* acc) {
    const BondAcc zero = {{0.0, 0.0, 0.0}, 0.0, {0.0, 0.0, 0.0}};
    acc[0] = t.n_terms[0] ? particle_terms_f32<2>(p, x_in, box, t.start[0], t.refs[0], t.idx[0], t.par[0]) : zero;
    acc[1] = t.n_terms[1] ? particle_terms_f32<3>(p, x_in, box, t.start[1], t.refs[1], t.idx[1], t.par[1]) : zero;
    acc[2] = t.n_terms[2] ? particle_terms<float, 4>(p, x_in, box, t.start[2], t.refs[2], t.idx[2], t.par[2], t.dih_type)
                          : zero;
    finish_particle<float>(p, x_in, x_out, vel, box, mass, half_dt, n_kicks, dt, f_out, acc);
}

}  // namespace hymd
