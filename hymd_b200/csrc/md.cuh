// Per-particle arithmetic of the velocity-Verlet / CSVR caller side of the hot path (row f2).
// __host__ __device__ so that tests/native/ can run exactly this source on the CPU.
//   hymd/integrator.py:9-75          integrate_velocity / integrate_position
//   hymd/main.py:829-837             inner rRESPA step (kick, drift, np.mod wrap)
//   hymd/thermostat.py:12-15, 177-219  cancel_com_momentum / csvr_thermostat
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#endif

namespace hymd {

constexpr int MD_MAX_FORCES = 8;

// v + 0.5*dt * (sum_k f_k) / m, evaluated in the array type like numpy does for float32 arrays with
// Python-float scalars (main.py:803-827, 830-834: the forces are summed first, then divided by mass).
template <typename real>
__host__ __device__ inline real kick(real v, const real* f_terms, int nf, real mass, real half_dt) {
    real a = f_terms[0];
    for (int k = 1; k < nf; ++k) a = a + f_terms[k];
    return v + half_dt * (a / mass);
}

// np.mod(x + dt*v, L) (main.py:836-837); a result that rounds up to L maps to 0 so x stays in [0, L).
template <typename real>
__host__ __device__ inline real drift_wrap(real x, real v, real dt, real L) {
    real y = x + dt * v;
    y = y - floor(y / L) * L;
    if (y >= L || y < (real)0) y = (real)0;
    return y;
}

// Moments of a velocity set: {count, sum vx, sum vy, sum vz, sum |v|^2}
constexpr int MOM = 5;

// csvr_thermostat for one coupling group, given the (globally reduced) moments
//   mom[0..4]  = moments of the group, mom[5..9] = moments of all particles.
// Returns alpha and the centre-of-mass velocity to remove/re-add; *dK = thermostat work.
struct CsvrScale {
    double alpha, cx, cy, cz;
    int group_only;     // 1: rescale the group's particles about their c.o.m.; 0: rescale ALL velocities
};
__host__ __device__ inline CsvrScale csvr_scale(const double* mom, double mass, double kT15, double c,
                                                double R, double SNf, int remove_com, double* dK) {
    const double n_g = mom[0];
    CsvrScale s;
    double K;
    if (remove_com && n_g > 1.0) {          // thermostat.py:186-189
        s.cx = mom[1] / n_g; s.cy = mom[2] / n_g; s.cz = mom[3] / n_g;
        K = 0.5 * mass * (mom[4] - n_g * (s.cx * s.cx + s.cy * s.cy + s.cz * s.cz));
        s.group_only = 1;
    } else {                                 // thermostat.py:190-191: all local velocities
        s.cx = s.cy = s.cz = 0.0;
        K = 0.5 * mass * mom[MOM + 4];
        s.group_only = 0;
    }
    const double K_target = kT15 * n_g;      // 1.5 R T0 n_g
    const double N_f = 3.0 * n_g;
    const double alpha2 = c + (1.0 - c) * (SNf + R * R) * K_target / (N_f * K) +
                          2.0 * R * sqrt(c * (1.0 - c) * K_target / (N_f * K));
    *dK = K * (alpha2 - 1.0);
    s.alpha = sqrt(alpha2);
    return s;
}

template <typename real>
__host__ __device__ inline void csvr_apply_particle(real* v3, const CsvrScale& s, bool in_group) {
    if (s.group_only) {
        if (in_group) {          // thermostat.py:211-215
            v3[0] = (real)(((double)v3[0] - s.cx) * s.alpha + s.cx);
            v3[1] = (real)(((double)v3[1] - s.cy) * s.alpha + s.cy);
            v3[2] = (real)(((double)v3[2] - s.cz) * s.alpha + s.cz);
        }
    } else {                     // thermostat.py:216-217
        v3[0] = (real)((double)v3[0] * s.alpha);
        v3[1] = (real)((double)v3[1] * s.alpha);
        v3[2] = (real)((double)v3[2] * s.alpha);
    }
}

}  // namespace hymd
