// Intramolecular forces on the device: hymd_bonded_create / _forces.
// Replaces the f2py kernels cbf / caf / cdf (hymd/compute_bond_forces.f90, compute_angle_forces.f90,
// compute_dihedral_forces.f90) that main.py:841-887 calls respa_inner times per outer step.
// The arithmetic lives in bonded.cuh (shared with the CPU check of tests/native/).
#include <stdlib.h>

#include "bonded.cuh"
#include "bonded_f32.cuh"
#include "ctx.cuh"

struct hymd_bonded {
    long long n_particles;
    long long n_terms[3];        // bonds, angles, dihedrals
    uint32_t* start[3];          // [n_particles + 1]
    uint32_t* refs[3];           // [n_terms * slots]
    int32_t* idx[3];             // [n_terms][4]
    double* par[3];              // [n_terms][2] / [n_terms][30]
    int32_t* dih_type;           // [n4]
    int32_t* dih_last;           // [n4] bonds_4_last: 1 = last dihedral of a backbone (hymd_bonded_set_last)
    long long n_cbt;             // dihedrals of dih_type 1 (combined bending-torsion)
    uint32_t* cta_start[3];      // CTA-cooperative evaluation (bonded.cuh, CtaLists)
    uint32_t* cta_terms[3];
    uint32_t* lrefs[3];
    hymd::TermRec* rec[2];       // inline bond / angle records of the CTA lists (mode 2)
    int max_terms[3];
    size_t cta_smem;             // dynamic shared memory of inner_step_cta_kernel
    int use_cta;                 // HYMD_B200_BONDED_CTA / hymd_bonded_set_cta: 0, 1, 2 or 3
    int tile;                    // particles per CTA of the cooperative kernels (HYMD_B200_BONDED_TILE)
    int f32math;                 // hymd_bonded_set_math / HYMD_B200_BONDED_F32MATH: float arithmetic for bonds
                                 // and angles in the fp32 build's per-particle fused step (bonded_f32.cuh)
    int occ;                     // HYMD_B200_BONDED_OCC: 0 (default, no register limit), 6 or 8 resident CTAs per SM
    double* out12;               // scratch result of the fused kernels
    double* partial;             // [max_blocks][4] block partials of {energy, pr_x, pr_y, pr_z}
    int max_blocks;
    int64_t launches;
};

namespace hymd {

constexpr int BONDED_THREADS = 128;

template <typename real, int KIND>
__global__ void __launch_bounds__(BONDED_THREADS) bonded_kernel(
    const real* __restrict__ pos, long long n, Vec3d box, const uint32_t* __restrict__ start,
    const uint32_t* __restrict__ refs, const int32_t* __restrict__ idx, const double* __restrict__ par,
    const int32_t* __restrict__ dtype, real* __restrict__ force, double* __restrict__ partial) {
    const long long p = (long long)blockIdx.x * BONDED_THREADS + threadIdx.x;
    BondAcc acc = {{0.0, 0.0, 0.0}, 0.0, {0.0, 0.0, 0.0}};
    if (p < n) {
        acc = particle_terms<real, KIND>(p, pos, box, start, refs, idx, par, dtype);
        force[3 * p + 0] = (real)acc.f.x;
        force[3 * p + 1] = (real)acc.f.y;
        force[3 * p + 2] = (real)acc.f.z;
    }
    // fixed-order block reduction of {e, pr}: lanes by shuffle, warps through shared memory
    double v[4] = {acc.e, acc.pr.x, acc.pr.y, acc.pr.z};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
    __shared__ double sh[BONDED_THREADS / 32][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
        for (int k = 0; k < 4; ++k) sh[warp][k] = v[k];
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0.0;
        for (int w2 = 0; w2 < BONDED_THREADS / 32; ++w2) s += sh[w2][threadIdx.x];
        partial[4 * (long long)blockIdx.x + threadIdx.x] = s;
    }
}

// bonds / angles with the single-precision evaluators (bonded_f32.cuh; fp32 build, opt-in)
template <int KIND>
__global__ void __launch_bounds__(BONDED_THREADS) bonded_f32_kernel(
    const float* __restrict__ pos, long long n, Vec3d box, const uint32_t* __restrict__ start,
    const uint32_t* __restrict__ refs, const int32_t* __restrict__ idx, const double* __restrict__ par,
    float* __restrict__ force, double* __restrict__ partial) {
    const long long p = (long long)blockIdx.x * BONDED_THREADS + threadIdx.x;
    BondAcc acc = {{0.0, 0.0, 0.0}, 0.0, {0.0, 0.0, 0.0}};
    if (p < n) {
        acc = particle_terms_f32<KIND>(p, pos, box, start, refs, idx, par);
        force[3 * p + 0] = (float)acc.f.x;
        force[3 * p + 1] = (float)acc.f.y;
        force[3 * p + 2] = (float)acc.f.z;
    }
    double v[4] = {acc.e, acc.pr.x, acc.pr.y, acc.pr.z};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
    __shared__ double sh[BONDED_THREADS / 32][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
        for (int k = 0; k < 4; ++k) sh[warp][k] = v[k];
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0.0;
        for (int w2 = 0; w2 < BONDED_THREADS / 32; ++w2) s += sh[w2][threadIdx.x];
        partial[4 * (long long)blockIdx.x + threadIdx.x] = s;
    }
}

struct ForceOut {
    void* f[3];
};

// Fused inner rRESPA step: bonded forces of all kinds + velocity kick(s) + drift/wrap in one pass
// (48 B per particle of HBM traffic in fp32 instead of 176 B for the four separate launches; the
// per-kind force arrays are only written on request).
// CBT: topologies with dtype-1 dihedrals -- the bending terms (bonded.cuh: particle_cbt) join the dihedral kind; a
// separate instantiation, so that the common kernel keeps its registers.
template <typename real, bool CBT>
__global__ void __launch_bounds__(BONDED_THREADS) inner_step_kernel(
    const real* __restrict__ x_in, real* __restrict__ x_out, real* __restrict__ vel, long long n, Vec3d box,
    TermLists t, real mass, real half_dt, int n_kicks, real dt, ForceOut fo, double* __restrict__ partial) {
    const long long p = (long long)blockIdx.x * BONDED_THREADS + threadIdx.x;
    BondAcc acc[3];
    double v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = 0.0;
    if (p < n) {
        real* f_out[3] = {(real*)fo.f[0], (real*)fo.f[1], (real*)fo.f[2]};
        if (CBT) inner_step_particle_cbt<real>(p, x_in, x_out, vel, box, t, mass, half_dt, n_kicks, dt, f_out, acc);
        else inner_step_particle<real>(p, x_in, x_out, vel, box, t, mass, half_dt, n_kicks, dt, f_out, acc);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            v[4 * k] = acc[k].e; v[4 * k + 1] = acc[k].pr.x; v[4 * k + 2] = acc[k].pr.y; v[4 * k + 3] = acc[k].pr.z;
        }
    }
#pragma unroll
    for (int k = 0; k < 12; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
    __shared__ double sh[BONDED_THREADS / 32][12];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
        for (int k = 0; k < 12; ++k) sh[warp][k] = v[k];
    __syncthreads();
    if (threadIdx.x < 12) {
        double s = 0.0;
        for (int w2 = 0; w2 < BONDED_THREADS / 32; ++w2) s += sh[w2][threadIdx.x];
        partial[12 * (long long)blockIdx.x + threadIdx.x] = s;
    }
}

// Same step with single-precision bond / angle arithmetic (bonded_f32.cuh; fp32 build, opt-in).
__global__ void __launch_bounds__(BONDED_THREADS) inner_step_f32_kernel(
    const float* __restrict__ x_in, float* __restrict__ x_out, float* __restrict__ vel, long long n, Vec3d box,
    TermLists t, float mass, float half_dt, int n_kicks, float dt, ForceOut fo, double* __restrict__ partial) {
    const long long p = (long long)blockIdx.x * BONDED_THREADS + threadIdx.x;
    BondAcc acc[3];
    double v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = 0.0;
    if (p < n) {
        float* f_out[3] = {(float*)fo.f[0], (float*)fo.f[1], (float*)fo.f[2]};
        inner_step_particle_f32(p, x_in, x_out, vel, box, t, mass, half_dt, n_kicks, dt, f_out, acc);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            v[4 * k] = acc[k].e; v[4 * k + 1] = acc[k].pr.x; v[4 * k + 2] = acc[k].pr.y; v[4 * k + 3] = acc[k].pr.z;
        }
    }
#pragma unroll
    for (int k = 0; k < 12; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
    __shared__ double sh[BONDED_THREADS / 32][12];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
        for (int k = 0; k < 12; ++k) sh[warp][k] = v[k];
    __syncthreads();
    if (threadIdx.x < 12) {
        double s = 0.0;
        for (int w2 = 0; w2 < BONDED_THREADS / 32; ++w2) s += sh[w2][threadIdx.x];
        partial[12 * (long long)blockIdx.x + threadIdx.x] = s;
    }
}

// Same step, CTA-cooperative term evaluation (bonded.cuh): every term touching the CTA's 128
// particles is evaluated once into shared memory, then each particle gathers its slots.
// MINB = resident CTAs per SM the register allocation is limited for (0: no limit, 94 registers,
// 5 CTAs; 6: 80 registers; 8: 64 registers with spills) -- HYMD_B200_BONDED_OCC, to be measured.
template <typename real, int MINB>
__global__ void __launch_bounds__(BONDED_THREADS, MINB) inner_step_cta_kernel(
    const real* __restrict__ x_in, real* __restrict__ x_out, real* __restrict__ vel, long long n, Vec3d box,
    TermLists t, CtaLists c, int tile, real mass, real half_dt, int n_kicks, real dt, ForceOut fo,
    double* __restrict__ partial) {
    extern __shared__ double sm[];
    const long long cta = blockIdx.x;
    const long long p0 = cta * tile;                  // tile = 128 * (particles per thread)
    const long long p1 = p0 + tile < n ? p0 + tile : n;
    double v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = 0.0;
    cta_eval_terms<real>((int)threadIdx.x, BONDED_THREADS, cta, p0, p1, x_in, box, t, c, sm, v);
    __syncthreads();
    real* f_out[3] = {(real*)fo.f[0], (real*)fo.f[1], (real*)fo.f[2]};
    for (long long p = p0 + threadIdx.x; p < p1; p += BONDED_THREADS) {
        BondAcc acc[3];
        cta_gather_particle(p, t, c, sm, acc);
        finish_particle<real>(p, x_in, x_out, vel, box, mass, half_dt, n_kicks, dt, f_out, acc);
    }
#pragma unroll
    for (int k = 0; k < 12; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
    __shared__ double sh[BONDED_THREADS / 32][12];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
        for (int k = 0; k < 12; ++k) sh[warp][k] = v[k];
    __syncthreads();
    if (threadIdx.x < 12) {
        double s = 0.0;
        for (int w2 = 0; w2 < BONDED_THREADS / 32; ++w2) s += sh[w2][threadIdx.x];
        partial[12 * (long long)blockIdx.x + threadIdx.x] = s;
    }
}

// Mode 2: own positions staged in shared memory, inline term records (bonded.cuh).  The per-particle
// list bounds are fetched before the first barrier so their latency overlaps phase 1.
template <typename real>
__global__ void __launch_bounds__(BONDED_THREADS) inner_step_cta2_kernel(
    const real* __restrict__ x_in, real* __restrict__ x_out, real* __restrict__ vel, long long n, Vec3d box,
    TermLists t, CtaLists c, CtaRecs rc, int tile, real mass, real half_dt, int n_kicks, real dt, ForceOut fo,
    double* __restrict__ partial) {
    extern __shared__ double sm[];
    real* own_pos = (real*)sm;                                         // [tile * 3] reals
    double* vecs = sm + ((size_t)tile * 3 * sizeof(real) + 7) / 8;
    const long long cta = blockIdx.x;
    const long long p0 = cta * tile;
    const long long p1 = p0 + tile < n ? p0 + tile : n;
    const int n_own = (int)(p1 - p0);
    for (int i = threadIdx.x; i < 3 * n_own; i += BONDED_THREADS) own_pos[i] = x_in[3 * p0 + i];
    __syncthreads();
    const PosTile<real> x = {x_in, own_pos, p0, p1};
    const long long first = p0 + threadIdx.x;
    RefBounds rb0;
    if (first < p1) rb0 = ref_bounds(first, t);     // in flight during phase 1
    double v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = 0.0;
    cta2_eval_terms<real, PosTile<real>>((int)threadIdx.x, BONDED_THREADS, cta, p0, p1, x, box, t, c, rc, vecs, v);
    __syncthreads();
    real* f_out[3] = {(real*)fo.f[0], (real*)fo.f[1], (real*)fo.f[2]};
    for (long long p = first; p < p1; p += BONDED_THREADS) {
        BondAcc acc[3];
        cta_gather_bounds(p == first ? rb0 : ref_bounds(p, t), c, vecs, acc);
        finish_particle<real>(p, x_in, x_out, vel, box, mass, half_dt, n_kicks, dt, f_out, acc);
    }
#pragma unroll
    for (int k = 0; k < 12; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
    __shared__ double sh[BONDED_THREADS / 32][12];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
        for (int k = 0; k < 12; ++k) sh[warp][k] = v[k];
    __syncthreads();
    if (threadIdx.x < 12) {
        double s = 0.0;
        for (int w2 = 0; w2 < BONDED_THREADS / 32; ++w2) s += sh[w2][threadIdx.x];
        partial[12 * (long long)blockIdx.x + threadIdx.x] = s;
    }
}

// Mode 3: inline term records like mode 2, positions read from global memory like mode 1: one dependent
// load level less than mode 1 without mode 2's staging barrier.
template <typename real>
__global__ void __launch_bounds__(BONDED_THREADS) inner_step_cta3_kernel(
    const real* __restrict__ x_in, real* __restrict__ x_out, real* __restrict__ vel, long long n, Vec3d box,
    TermLists t, CtaLists c, CtaRecs rc, int tile, real mass, real half_dt, int n_kicks, real dt, ForceOut fo,
    double* __restrict__ partial) {
    extern __shared__ double sm[];
    const long long cta = blockIdx.x;
    const long long p0 = cta * tile;
    const long long p1 = p0 + tile < n ? p0 + tile : n;
    const long long first = p0 + threadIdx.x;
    RefBounds rb0;
    if (first < p1) rb0 = ref_bounds(first, t);
    double v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = 0.0;
    const real* xg = x_in;     // plain pointer: the accessor parameter is a const reference
    cta2_eval_terms<real, const real*>((int)threadIdx.x, BONDED_THREADS, cta, p0, p1, xg, box, t, c, rc, sm, v);
    __syncthreads();
    real* f_out[3] = {(real*)fo.f[0], (real*)fo.f[1], (real*)fo.f[2]};
    for (long long p = first; p < p1; p += BONDED_THREADS) {
        BondAcc acc[3];
        cta_gather_bounds(p == first ? rb0 : ref_bounds(p, t), c, sm, acc);
        finish_particle<real>(p, x_in, x_out, vel, box, mass, half_dt, n_kicks, dt, f_out, acc);
    }
#pragma unroll
    for (int k = 0; k < 12; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
    __shared__ double sh[BONDED_THREADS / 32][12];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
        for (int k = 0; k < 12; ++k) sh[warp][k] = v[k];
    __syncthreads();
    if (threadIdx.x < 12) {
        double s = 0.0;
        for (int w2 = 0; w2 < BONDED_THREADS / 32; ++w2) s += sh[w2][threadIdx.x];
        partial[12 * (long long)blockIdx.x + threadIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256) inner_final_kernel(const double* __restrict__ partial, int nblocks,
                                                          double* __restrict__ out) {
    __shared__ double sh[12][256];
    double acc[12];
    for (int k = 0; k < 12; ++k) acc[k] = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256)
        for (int k = 0; k < 12; ++k) acc[k] += partial[12 * (long long)i + k];
    for (int k = 0; k < 12; ++k) sh[k][threadIdx.x] = acc[k];
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w)
            for (int k = 0; k < 12; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x < 12) out[threadIdx.x] = sh[threadIdx.x][0];
}

__global__ void __launch_bounds__(256) bonded_final_kernel(const double* __restrict__ partial, int nblocks,
                                                           double* __restrict__ out) {
    __shared__ double sh[4][256];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < nblocks; i += 256)
        for (int k = 0; k < 4; ++k) acc[k] += partial[4 * (long long)i + k];
    for (int k = 0; k < 4; ++k) sh[k][threadIdx.x] = acc[k];
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w)
            for (int k = 0; k < 4; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x < 4) out[threadIdx.x] = sh[threadIdx.x][0];
}

// ---- dtype-1 dihedrals (bonded.cuh: cbt_eval, dipole_term, redistribute_term) ---------------------------
// The bending term of the combined bending-torsion dihedrals, gathered per particle like everything else here (no
// atomics, fixed order) and ADDED to the force array the dihedral kernel has just written; energy into `partial`.
// A separate pass: topologies without dtype 1 never launch it and the hot kernels keep their registers.
template <typename real>
__global__ void __launch_bounds__(BONDED_THREADS) cbt_kernel(
    const real* __restrict__ pos, long long n, Vec3d box, const uint32_t* __restrict__ start,
    const uint32_t* __restrict__ refs, const int32_t* __restrict__ idx, const double* __restrict__ par,
    const int32_t* __restrict__ dtype, const int32_t* __restrict__ last, real* __restrict__ force,
    double* __restrict__ partial) {
    const long long p = (long long)blockIdx.x * BONDED_THREADS + threadIdx.x;
    double e_acc = 0.0;
    if (p < n) {
        Vec3d acc = {0.0, 0.0, 0.0};
        bool any = false;
        for (uint32_t k = start[p]; k < start[p + 1]; ++k) {
            const long long t = refs[k] >> 2;
            const int slot = (int)(refs[k] & 3u);
            if (dtype[t] != 1) continue;
            const int32_t* ix = idx + 4 * t;
            Vec3d out[4];
            double e;
            cbt_eval(pos, box, ix[0], ix[1], ix[2], ix[3], par + (long long)DIH_ROWS * DIH_COLS * t, last[t], out, e);
            acc = acc + out[slot];
            if (slot == 0) e_acc += e;
            any = true;
        }
        if (any) {
            force[3 * p + 0] = (real)((double)force[3 * p + 0] + acc.x);
            force[3 * p + 1] = (real)((double)force[3 * p + 1] + acc.y);
            force[3 * p + 2] = (real)((double)force[3 * p + 2] + acc.z);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e_acc += __shfl_down_sync(0xffffffffu, e_acc, o);
    __shared__ double sh[BONDED_THREADS / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sh[warp] = e_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w2 = 0; w2 < BONDED_THREADS / 32; ++w2) s += sh[w2];
        partial[blockIdx.x] = s;
    }
}

// out[0] += sum of the block partials (fixed order)
__global__ void __launch_bounds__(256) cbt_final_kernel(const double* __restrict__ partial, int nblocks,
                                                        double* __restrict__ out) {
    __shared__ double sh[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256) acc += partial[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] += sh[0];
}

// one thread per dihedral: its (4,3) dipole positions and (6,3,3) transfer matrices (zeros unless dtype 1)
template <typename real>
__global__ void __launch_bounds__(BONDED_THREADS) dipole_kernel(
    const real* __restrict__ pos, long long n4, Vec3d box, const int32_t* __restrict__ idx,
    const double* __restrict__ par, const int32_t* __restrict__ dtype, const int32_t* __restrict__ last,
    real* __restrict__ dipoles, real* __restrict__ transfer) {
    const long long t = (long long)blockIdx.x * BONDED_THREADS + threadIdx.x;
    if (t >= n4) return;
    const int32_t* ix = idx + 4 * t;
    dipole_term<real>(pos, box, ix[0], ix[1], ix[2], ix[3], par + (long long)DIH_ROWS * DIH_COLS * t, dtype[t], last[t],
                      dipoles + 12 * t, transfer + 54 * t);
}

// dipole_forces_redistribution (force.py:855-880), gathered per bead in ascending dihedral order
template <typename real>
__global__ void __launch_bounds__(BONDED_THREADS) redistribute_kernel(
    long long n, const uint32_t* __restrict__ start, const uint32_t* __restrict__ refs,
    const int32_t* __restrict__ dtype, const int32_t* __restrict__ last, const real* __restrict__ f_dipoles,
    const real* __restrict__ transfer, real* __restrict__ f_beads) {
    const long long p = (long long)blockIdx.x * BONDED_THREADS + threadIdx.x;
    if (p >= n) return;
    Vec3d acc = {0.0, 0.0, 0.0};
    for (uint32_t k = start[p]; k < start[p + 1]; ++k) {
        const long long t = refs[k] >> 2;
        if (dtype[t] != 1) continue;
        acc = acc + redistribute_term<real>((int)(refs[k] & 3u), last[t], f_dipoles + 12 * t, transfer + 54 * t);
    }
    f_beads[3 * p + 0] = (real)acc.x;
    f_beads[3 * p + 1] = (real)acc.y;
    f_beads[3 * p + 2] = (real)acc.z;
}

template <typename T>
static int to_device(T** dst, const T* src, size_t count) {
    *dst = nullptr;
    HYMD_CUDA(cudaMalloc((void**)dst, (count ? count : 1) * sizeof(T)));
    if (count) HYMD_CUDA(cudaMemcpy(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice));
    return HYMD_OK;
}

static int upload_kind(hymd_bonded* b, int kind, long long n_terms, int slots,
                       const int32_t* const* index, const double* par, size_t par_per_term) {
    std::vector<uint32_t> start, refs;
    if (!build_particle_csr(b->n_particles, n_terms, slots, index, start, refs)) {
        set_error("bonded terms of %d particles: index outside [0, %lld), a particle twice in one term, "
                  "or more than 2^30 terms", slots, b->n_particles);
        return HYMD_ERR_INVALID;
    }
    std::vector<int32_t> idx((size_t)n_terms * 4, 0);
    for (long long t = 0; t < n_terms; ++t)
        for (int s = 0; s < slots; ++s) idx[(size_t)4 * t + s] = index[s][t];
    std::vector<uint32_t> cta_start, cta_terms, lrefs;
    build_cta_lists(b->n_particles, n_terms, slots, index, b->tile, start, cta_start, cta_terms, lrefs,
                    b->max_terms[kind]);
    HYMD_CHECK(to_device(&b->cta_start[kind], cta_start.data(), cta_start.size()));
    HYMD_CHECK(to_device(&b->cta_terms[kind], cta_terms.data(), cta_terms.size()));
    HYMD_CHECK(to_device(&b->lrefs[kind], lrefs.data(), lrefs.size()));
    if (kind < 2) {
        std::vector<TermRec> rec;
        build_cta_records(cta_terms, idx.data(), par, rec);
        HYMD_CHECK(to_device(&b->rec[kind], rec.data(), rec.size()));
    }
    b->n_terms[kind] = n_terms;
    HYMD_CHECK(to_device(&b->start[kind], start.data(), start.size()));
    HYMD_CHECK(to_device(&b->refs[kind], refs.data(), refs.size()));
    HYMD_CHECK(to_device(&b->idx[kind], idx.data(), idx.size()));
    HYMD_CHECK(to_device(&b->par[kind], par, (size_t)n_terms * par_per_term));
    return HYMD_OK;
}

template <typename real>
static int launch_cbt(hymd_bonded* b, const real* pos, Vec3d box, real* force, double* d_out, cudaStream_t s);

template <typename real>
static int launch_kind(hymd_bonded* b, int kind, const real* pos, Vec3d box, real* force, double* d_out,
                       cudaStream_t s) {
    const long long n = b->n_particles;
    const int blocks = (int)((n + BONDED_THREADS - 1) / BONDED_THREADS);
    if (blocks > 0 && b->f32math && sizeof(real) == 4 && kind < 2) {
        if (kind == 0)
            bonded_f32_kernel<2><<<blocks, BONDED_THREADS, 0, s>>>((const float*)pos, n, box, b->start[0], b->refs[0],
                                                                 b->idx[0], b->par[0], (float*)force, b->partial);
        else
            bonded_f32_kernel<3><<<blocks, BONDED_THREADS, 0, s>>>((const float*)pos, n, box, b->start[1], b->refs[1],
                                                                 b->idx[1], b->par[1], (float*)force, b->partial);
        HYMD_LAUNCH_CHECK(b);
    } else if (blocks > 0) {
        if (kind == 0)
            bonded_kernel<real, 2><<<blocks, BONDED_THREADS, 0, s>>>(
                pos, n, box, b->start[0], b->refs[0], b->idx[0], b->par[0], nullptr, force, b->partial);
        else if (kind == 1)
            bonded_kernel<real, 3><<<blocks, BONDED_THREADS, 0, s>>>(
                pos, n, box, b->start[1], b->refs[1], b->idx[1], b->par[1], nullptr, force, b->partial);
        else
            bonded_kernel<real, 4><<<blocks, BONDED_THREADS, 0, s>>>(
                pos, n, box, b->start[2], b->refs[2], b->idx[2], b->par[2], b->dih_type, force, b->partial);
        HYMD_LAUNCH_CHECK(b);
    }
    bonded_final_kernel<<<1, 256, 0, s>>>(b->partial, blocks, d_out);
    HYMD_LAUNCH_CHECK(b);
    if (kind == 2) return launch_cbt<real>(b, pos, box, force, d_out, s);
    return HYMD_OK;
}

// the bending term of the dtype-1 dihedrals on top of a dihedral force array / energy just computed on stream s
template <typename real>
static int launch_cbt(hymd_bonded* b, const real* pos, Vec3d box, real* force, double* d_out, cudaStream_t s) {
    if (b->n_cbt == 0 || b->n_particles == 0) return HYMD_OK;
    const long long n = b->n_particles;
    const int blocks = (int)((n + BONDED_THREADS - 1) / BONDED_THREADS);
    cbt_kernel<real><<<blocks, BONDED_THREADS, 0, s>>>(pos, n, box, b->start[2], b->refs[2], b->idx[2], b->par[2],
                                                       b->dih_type, b->dih_last, force, b->partial);
    HYMD_LAUNCH_CHECK(b);
    cbt_final_kernel<<<1, 256, 0, s>>>(b->partial, blocks, d_out);
    HYMD_LAUNCH_CHECK(b);
    return HYMD_OK;
}

constexpr size_t CTA_SMEM_LIMIT = 160 * 1024;

template <typename real>
static int launch_inner(hymd_bonded* b, int kind_mask, const real* x_in, real* x_out, real* vel, Vec3d box,
                        double mass, double kick_dt, int n_kicks, double drift_dt, void* const* d_force_out,
                        double* d_out, cudaStream_t s) {
    const long long n = b->n_particles;
    // dtype-1 dihedrals: the per-particle kernel that carries the bending term, whatever evaluation mode is set
    const bool cbt = b->n_cbt > 0 && ((kind_mask >> 2) & 1);
    const int per_cta = (b->use_cta && !cbt) ? b->tile : BONDED_THREADS;
    const int blocks = (int)((n + per_cta - 1) / per_cta);
    TermLists t;
    CtaLists c;
    for (int k = 0; k < 3; ++k) {
        t.start[k] = b->start[k]; t.refs[k] = b->refs[k]; t.idx[k] = b->idx[k]; t.par[k] = b->par[k];
        t.n_terms[k] = (kind_mask >> k) & 1 ? b->n_terms[k] : 0;
        c.cta_start[k] = b->cta_start[k]; c.cta_terms[k] = b->cta_terms[k]; c.lrefs[k] = b->lrefs[k];
        c.max_terms[k] = b->max_terms[k];
    }
    t.dih_type = b->dih_type;
    t.dih_last = b->dih_last;
    ForceOut fo;
    for (int k = 0; k < 3; ++k) fo.f[k] = d_force_out ? d_force_out[k] : nullptr;
    if (blocks > 0 && cbt) {
        inner_step_kernel<real, true><<<blocks, BONDED_THREADS, 0, s>>>(x_in, x_out, vel, n, box, t, (real)mass,
                                                                        (real)(0.5 * kick_dt), n_kicks,
                                                                        (real)drift_dt, fo, b->partial);
        HYMD_LAUNCH_CHECK(b);
    } else if (blocks > 0 && b->f32math && sizeof(real) == 4 && !b->use_cta) {
        inner_step_f32_kernel<<<blocks, BONDED_THREADS, 0, s>>>((const float*)x_in, (float*)x_out, (float*)vel, n, box,
                                                              t, (float)mass, (float)(0.5 * kick_dt), n_kicks,
                                                              (float)drift_dt, fo, b->partial);
        HYMD_LAUNCH_CHECK(b);
    } else if (blocks > 0) {
        if (b->use_cta == 3) {
            CtaRecs rc;
            rc.rec[0] = b->rec[0];
            rc.rec[1] = b->rec[1];
            inner_step_cta3_kernel<real><<<blocks, BONDED_THREADS, b->cta_smem, s>>>(
                x_in, x_out, vel, n, box, t, c, rc, b->tile, (real)mass, (real)(0.5 * kick_dt), n_kicks,
                (real)drift_dt, fo, b->partial);
        } else if (b->use_cta == 2) {
            CtaRecs rc;
            rc.rec[0] = b->rec[0];
            rc.rec[1] = b->rec[1];
            inner_step_cta2_kernel<real><<<blocks, BONDED_THREADS, b->cta_smem + (size_t)b->tile * 3 * sizeof(real) + 8, s>>>(
                x_in, x_out, vel, n, box, t, c, rc, b->tile, (real)mass, (real)(0.5 * kick_dt), n_kicks, (real)drift_dt,
                fo, b->partial);
        } else if (b->use_cta) {
#define HYMD_CTA_LAUNCH(MINB)                                                                              \
    inner_step_cta_kernel<real, MINB><<<blocks, BONDED_THREADS, b->cta_smem, s>>>(                         \
        x_in, x_out, vel, n, box, t, c, b->tile, (real)mass, (real)(0.5 * kick_dt), n_kicks, (real)drift_dt, fo, \
        b->partial)
            if (b->occ == 8) HYMD_CTA_LAUNCH(8);
            else if (b->occ == 6) HYMD_CTA_LAUNCH(6);
            else HYMD_CTA_LAUNCH(0);
#undef HYMD_CTA_LAUNCH
        }
        else
            inner_step_kernel<real, false><<<blocks, BONDED_THREADS, 0, s>>>(x_in, x_out, vel, n, box, t, (real)mass,
                                                                             (real)(0.5 * kick_dt), n_kicks,
                                                                             (real)drift_dt, fo, b->partial);
        HYMD_LAUNCH_CHECK(b);
    }
    if (d_out) {
        inner_final_kernel<<<1, 256, 0, s>>>(b->partial, blocks, d_out);
        HYMD_LAUNCH_CHECK(b);
    }
    return HYMD_OK;
}

static int set_cta(hymd_bonded* b, int enable) {
    if (!enable) { b->use_cta = 0; return HYMD_OK; }
    if (b->cta_smem > CTA_SMEM_LIMIT) {
        set_error("CTA-cooperative bonded evaluation needs %zu bytes of shared memory per CTA (limit %zu): "
                  "a block of %d consecutive particles takes part in too many terms", b->cta_smem,
                  CTA_SMEM_LIMIT, b->tile);
        return HYMD_ERR_CAPACITY;
    }
    if (b->cta_smem + 4096 > 48 * 1024) {
        const int bytes = (int)b->cta_smem + b->tile * 3 * (int)sizeof(double) + 8;
        HYMD_CUDA(cudaFuncSetAttribute(inner_step_cta_kernel<float, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        HYMD_CUDA(cudaFuncSetAttribute(inner_step_cta_kernel<double, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        HYMD_CUDA(cudaFuncSetAttribute(inner_step_cta_kernel<float, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        HYMD_CUDA(cudaFuncSetAttribute(inner_step_cta_kernel<double, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        HYMD_CUDA(cudaFuncSetAttribute(inner_step_cta_kernel<float, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        HYMD_CUDA(cudaFuncSetAttribute(inner_step_cta_kernel<double, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        HYMD_CUDA(cudaFuncSetAttribute(inner_step_cta2_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        HYMD_CUDA(cudaFuncSetAttribute(inner_step_cta2_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        HYMD_CUDA(cudaFuncSetAttribute(inner_step_cta3_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        HYMD_CUDA(cudaFuncSetAttribute(inner_step_cta3_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    }
    b->use_cta = (enable == 2 || enable == 3) ? enable : 1;
    return HYMD_OK;
}

}  // namespace hymd

using namespace hymd;

extern "C" {

int hymd_bonded_create(int64_t n_particles, int64_t n2, const int32_t* a2, const int32_t* b2,
                       const double* r0_2, const double* k_2, int64_t n3, const int32_t* a3,
                       const int32_t* b3, const int32_t* c3, const double* t0_3, const double* k_3,
                       int64_t n4, const int32_t* a4, const int32_t* b4, const int32_t* c4,
                       const int32_t* d4, const double* coeff4, const int32_t* type4, hymd_bonded** out) {
    if (!out || n_particles < 0 || n2 < 0 || n3 < 0 || n4 < 0 || (n2 && (!a2 || !b2 || !r0_2 || !k_2)) ||
        (n3 && (!a3 || !b3 || !c3 || !t0_3 || !k_3)) || (n4 && (!a4 || !b4 || !c4 || !d4 || !coeff4 || !type4))) {
        set_error("hymd_bonded_create: null or negative argument");
        return HYMD_ERR_INVALID;
    }
    if (n_particles >= (1LL << 31)) { set_error("more than 2^31 particles per GPU"); return HYMD_ERR_INVALID; }
    long long n_cbt = 0;
    for (int64_t t = 0; t < n4; ++t) {
        if (type4[t] < 0 || type4[t] > 2) {
            set_error("dihedral %lld has dih_type %d: expected 0 (cosine series), 1 (combined bending-torsion) or "
                      "2 (improper)", (long long)t, type4[t]);
            return HYMD_ERR_INVALID;
        }
        n_cbt += type4[t] == 1;
    }
    hymd_bonded* b = new hymd_bonded();
    memset(b, 0, sizeof(*b));
    b->n_particles = n_particles;
    b->n_cbt = n_cbt;
    b->tile = BONDED_THREADS;
    if (const char* env = getenv("HYMD_B200_BONDED_TILE")) {
        const int tile = atoi(env);
        if (tile < BONDED_THREADS || tile > 2048 || tile % BONDED_THREADS != 0) {
            set_error("HYMD_B200_BONDED_TILE = %s: expected a multiple of %d up to 2048", env, BONDED_THREADS);
            delete b;
            return HYMD_ERR_INVALID;
        }
        b->tile = tile;
    }
    if (const char* env = getenv("HYMD_B200_BONDED_F32MATH")) b->f32math = atoi(env) != 0;
    b->occ = 0;
    if (const char* env = getenv("HYMD_B200_BONDED_OCC")) {
        b->occ = atoi(env);
        if (b->occ != 0 && b->occ != 6 && b->occ != 8) {
            set_error("HYMD_B200_BONDED_OCC = %s: expected 0, 6 or 8", env);
            delete b;
            return HYMD_ERR_INVALID;
        }
    }
    int st = HYMD_OK;
    {
        std::vector<double> par((size_t)n2 * 2);
        for (int64_t t = 0; t < n2; ++t) { par[2 * t] = r0_2[t]; par[2 * t + 1] = k_2[t]; }
        const int32_t* index[2] = {a2, b2};
        st = upload_kind(b, 0, n2, 2, index, par.data(), 2);
    }
    if (st == HYMD_OK) {
        std::vector<double> par((size_t)n3 * 2);
        for (int64_t t = 0; t < n3; ++t) { par[2 * t] = t0_3[t]; par[2 * t + 1] = k_3[t]; }
        const int32_t* index[3] = {a3, b3, c3};
        st = upload_kind(b, 1, n3, 3, index, par.data(), 2);
    }
    if (st == HYMD_OK) {
        const int32_t* index[4] = {a4, b4, c4, d4};
        st = upload_kind(b, 2, n4, 4, index, coeff4, (size_t)DIH_ROWS * DIH_COLS);
    }
    if (st == HYMD_OK) st = to_device(&b->dih_type, type4, (size_t)n4);
    if (st == HYMD_OK) {
        std::vector<int32_t> zeros((size_t)n4, 0);
        st = to_device(&b->dih_last, zeros.data(), (size_t)n4);
    }
    if (st == HYMD_OK) {
        b->max_blocks = (int)((n_particles + BONDED_THREADS - 1) / BONDED_THREADS);
        cudaError_t e = cudaMalloc((void**)&b->partial, sizeof(double) * 12 * (size_t)(b->max_blocks + 1));
        if (e != cudaSuccess) { set_error("cudaMalloc failed: %s", cudaGetErrorString(e)); st = HYMD_ERR_NOMEM; }
    }
    if (st == HYMD_OK) {
        cudaError_t e = cudaMalloc((void**)&b->out12, sizeof(double) * 12);
        if (e != cudaSuccess) { set_error("cudaMalloc failed: %s", cudaGetErrorString(e)); st = HYMD_ERR_NOMEM; }
    }
    if (st == HYMD_OK) {
        b->cta_smem = 0;
        for (int k = 0; k < 3; ++k) b->cta_smem += sizeof(double) * CTA_DOUBLES[k] * (size_t)b->max_terms[k];
        const char* env = getenv("HYMD_B200_BONDED_CTA");
        if (env && atoi(env) != 0 && b->cta_smem <= CTA_SMEM_LIMIT) st = set_cta(b, atoi(env));
    }
    if (st != HYMD_OK) { hymd_bonded_destroy(b); return st; }
    *out = b;
    return HYMD_OK;
}

int hymd_bonded_destroy(hymd_bonded* b) {
    if (!b) return HYMD_OK;
    for (int k = 0; k < 3; ++k) {
        cudaFree(b->start[k]);
        cudaFree(b->refs[k]);
        cudaFree(b->idx[k]);
        cudaFree(b->par[k]);
        cudaFree(b->cta_start[k]);
        cudaFree(b->cta_terms[k]);
        cudaFree(b->lrefs[k]);
        if (k < 2) cudaFree(b->rec[k]);
    }
    cudaFree(b->out12);
    cudaFree(b->dih_type);
    cudaFree(b->dih_last);
    cudaFree(b->partial);
    delete b;
    return HYMD_OK;
}

int hymd_bonded_forces(hymd_bonded* b, int kind, int dtype, const void* d_pos, const double box[3],
                       void* d_force, double* d_out, void* stream) {
    if (!b || !box || !d_out || (b->n_particles > 0 && (!d_pos || !d_force))) {
        set_error("hymd_bonded_forces: null argument");
        return HYMD_ERR_INVALID;
    }
    if (kind < 2 || kind > 4) { set_error("kind = %d, expected 2, 3 or 4 particles per term", kind); return HYMD_ERR_INVALID; }
    if (dtype != HYMD_F32 && dtype != HYMD_F64) { set_error("bad dtype %d", dtype); return HYMD_ERR_INVALID; }
    const Vec3d bx = {box[0], box[1], box[2]};
    cudaStream_t s = (cudaStream_t)stream;
    if (b->use_cta) {   // one kind through the CTA-cooperative kernel: no velocities, force array on request
        void* fo[3] = {nullptr, nullptr, nullptr};
        fo[kind - 2] = d_force;
        const int st = dtype == HYMD_F64
            ? launch_inner<double>(b, 1 << (kind - 2), (const double*)d_pos, nullptr, nullptr, bx, 1.0, 0.0, 0,
                                   0.0, fo, b->out12, s)
            : launch_inner<float>(b, 1 << (kind - 2), (const float*)d_pos, nullptr, nullptr, bx, 1.0, 0.0, 0,
                                  0.0, fo, b->out12, s);
        if (st != HYMD_OK) return st;
        HYMD_CUDA(cudaMemcpyAsync(d_out, b->out12 + 4 * (kind - 2), 4 * sizeof(double),
                                  cudaMemcpyDeviceToDevice, s));
        if (kind == 4)
            return dtype == HYMD_F64 ? launch_cbt<double>(b, (const double*)d_pos, bx, (double*)d_force, d_out, s)
                                     : launch_cbt<float>(b, (const float*)d_pos, bx, (float*)d_force, d_out, s);
        return HYMD_OK;
    }
    if (dtype == HYMD_F64)
        return launch_kind<double>(b, kind - 2, (const double*)d_pos, bx, (double*)d_force, d_out, s);
    return launch_kind<float>(b, kind - 2, (const float*)d_pos, bx, (float*)d_force, d_out, s);
}

int hymd_bonded_inner_step(hymd_bonded* b, int dtype, const void* d_pos_in, void* d_pos_out, void* d_vel,
                           const double box[3], double mass, double kick_dt, int n_kicks, double drift_dt,
                           void* const* d_force_out, double* d_out, void* stream) {
    if (!b || !box || (b->n_particles > 0 && (!d_pos_in || !d_vel))) {
        set_error("hymd_bonded_inner_step: null argument");
        return HYMD_ERR_INVALID;
    }
    if (n_kicks < 0 || n_kicks > 2) { set_error("n_kicks = %d, expected 0, 1 or 2", n_kicks); return HYMD_ERR_INVALID; }
    if (d_pos_out == d_pos_in) { set_error("hymd_bonded_inner_step: d_pos_out must not alias d_pos_in"); return HYMD_ERR_INVALID; }
    if (dtype != HYMD_F32 && dtype != HYMD_F64) { set_error("bad dtype %d", dtype); return HYMD_ERR_INVALID; }
    const Vec3d bx = {box[0], box[1], box[2]};
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == HYMD_F64)
        return launch_inner<double>(b, 7, (const double*)d_pos_in, (double*)d_pos_out, (double*)d_vel, bx, mass,
                                    kick_dt, n_kicks, drift_dt, d_force_out, d_out, s);
    return launch_inner<float>(b, 7, (const float*)d_pos_in, (float*)d_pos_out, (float*)d_vel, bx, mass,
                               kick_dt, n_kicks, drift_dt, d_force_out, d_out, s);
}

int hymd_bonded_set_last(hymd_bonded* b, const int32_t* last4) {
    if (!b || (b->n_terms[2] > 0 && !last4)) { set_error("hymd_bonded_set_last: null argument"); return HYMD_ERR_INVALID; }
    if (b->n_terms[2] > 0)
        HYMD_CUDA(cudaMemcpy(b->dih_last, last4, sizeof(int32_t) * (size_t)b->n_terms[2], cudaMemcpyHostToDevice));
    return HYMD_OK;
}

int hymd_bonded_dipoles(hymd_bonded* b, int dtype, const void* d_pos, const double box[3], void* d_dipoles,
                        void* d_transfer, void* stream) {
    if (!b || !box || (b->n_terms[2] > 0 && (!d_pos || !d_dipoles || !d_transfer))) {
        set_error("hymd_bonded_dipoles: null argument");
        return HYMD_ERR_INVALID;
    }
    if (dtype != HYMD_F32 && dtype != HYMD_F64) { set_error("bad dtype %d", dtype); return HYMD_ERR_INVALID; }
    const long long n4 = b->n_terms[2];
    if (n4 == 0) return HYMD_OK;
    const Vec3d bx = {box[0], box[1], box[2]};
    cudaStream_t s = (cudaStream_t)stream;
    const int blocks = (int)((n4 + BONDED_THREADS - 1) / BONDED_THREADS);
    if (dtype == HYMD_F64)
        dipole_kernel<double><<<blocks, BONDED_THREADS, 0, s>>>((const double*)d_pos, n4, bx, b->idx[2], b->par[2], b->dih_type,
                                                                b->dih_last, (double*)d_dipoles, (double*)d_transfer);
    else
        dipole_kernel<float><<<blocks, BONDED_THREADS, 0, s>>>((const float*)d_pos, n4, bx, b->idx[2], b->par[2], b->dih_type,
                                                               b->dih_last, (float*)d_dipoles, (float*)d_transfer);
    HYMD_LAUNCH_CHECK(b);
    return HYMD_OK;
}

int hymd_dipole_redistribute(hymd_bonded* b, int dtype, const void* d_f_dipoles, const void* d_transfer,
                             void* d_f_beads, void* stream) {
    if (!b || (b->n_particles > 0 && !d_f_beads) || (b->n_terms[2] > 0 && (!d_f_dipoles || !d_transfer))) {
        set_error("hymd_dipole_redistribute: null argument");
        return HYMD_ERR_INVALID;
    }
    if (dtype != HYMD_F32 && dtype != HYMD_F64) { set_error("bad dtype %d", dtype); return HYMD_ERR_INVALID; }
    const long long n = b->n_particles;
    if (n == 0) return HYMD_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int blocks = (int)((n + BONDED_THREADS - 1) / BONDED_THREADS);
    if (dtype == HYMD_F64)
        redistribute_kernel<double><<<blocks, BONDED_THREADS, 0, s>>>(n, b->start[2], b->refs[2], b->dih_type, b->dih_last,
                                                                      (const double*)d_f_dipoles, (const double*)d_transfer,
                                                                      (double*)d_f_beads);
    else
        redistribute_kernel<float><<<blocks, BONDED_THREADS, 0, s>>>(n, b->start[2], b->refs[2], b->dih_type, b->dih_last,
                                                                     (const float*)d_f_dipoles, (const float*)d_transfer,
                                                                     (float*)d_f_beads);
    HYMD_LAUNCH_CHECK(b);
    return HYMD_OK;
}

int hymd_bonded_set_math(hymd_bonded* b, int f32math) {
    if (!b) { set_error("null argument"); return HYMD_ERR_INVALID; }
    b->f32math = f32math ? 1 : 0;
    return HYMD_OK;
}

int hymd_bonded_set_cta(hymd_bonded* b, int enable) {
    if (!b) { set_error("null argument"); return HYMD_ERR_INVALID; }
    return set_cta(b, enable);
}

int64_t hymd_bonded_launch_count(hymd_bonded* b) { return b ? b->launches : 0; }

}  // extern "C"
