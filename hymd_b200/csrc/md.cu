// Velocity-Verlet / rRESPA updates and the CSVR thermostat on the device:
// fused elementwise kernels that keep positions and velocities resident in HBM between field-force
// cycles.  Replaces the numpy expressions of hymd/main.py:803-837, 889-893, 1144-1169,
// hymd/integrator.py:9-75 and hymd/thermostat.py:12-15, 177-219.  Stateless: every entry point takes
// device pointers and a stream.
#include "ctx.cuh"
#include "md.cuh"

namespace hymd {

struct ForcePtrs {
    const void* f[MD_MAX_FORCES];
};

constexpr int MD_THREADS = 256;
constexpr int MD_MOM_BLOCKS = 148 * 4;

// v <- v + half_dt * (sum_k f_k) / mass              (sequential == 0; inner rRESPA kick, main.py:830-834)
// v <- ((v + half_dt*f_0/mass) + half_dt*f_1/mass) … (sequential == 1; outer kicks, main.py:803-827)
// and, when pos != nullptr, x <- mod(x + dt*v, L)    (main.py:836-837) in the same pass.
template <typename real>
__global__ void __launch_bounds__(MD_THREADS) kick_drift_kernel(
    real* __restrict__ vel, real* __restrict__ pos, ForcePtrs forces, int nf, int sequential, real mass,
    real half_dt, real dt, real Lx, real Ly, real Lz, long long n3) {
    const long long stride = (long long)gridDim.x * MD_THREADS;
    for (long long i = (long long)blockIdx.x * MD_THREADS + threadIdx.x; i < n3; i += stride) {
        real v = vel[i];
        if (nf > 0) {
            real ft[MD_MAX_FORCES];
#pragma unroll
            for (int k = 0; k < MD_MAX_FORCES; ++k)
                if (k < nf) ft[k] = ((const real*)forces.f[k])[i];
            if (sequential) {
#pragma unroll
                for (int k = 0; k < MD_MAX_FORCES; ++k)
                    if (k < nf) v = kick(v, &ft[k], 1, mass, half_dt);
            } else {
                v = kick(v, ft, nf, mass, half_dt);
            }
            vel[i] = v;
        }
        if (pos != nullptr) {
            const int d = (int)(i % 3);
            const real L = d == 0 ? Lx : (d == 1 ? Ly : Lz);
            pos[i] = drift_wrap(pos[i], v, dt, L);
        }
    }
}

// Moments {count, sum v, sum v^2} of the particles with group[i] == g (all particles when group ==
// nullptr or g < 0) -> out[0..4], and of ALL particles -> out[5..9].  Fixed-order two-stage double
// reduction (bitwise reproducible).
template <typename real>
__global__ void __launch_bounds__(MD_THREADS) moments_partial_kernel(
    const real* __restrict__ vel, const int32_t* __restrict__ group, int g, long long n,
    double* __restrict__ partial) {
    double acc[2 * MOM] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long stride = (long long)gridDim.x * MD_THREADS;
    for (long long i = (long long)blockIdx.x * MD_THREADS + threadIdx.x; i < n; i += stride) {
        const double vx = (double)vel[3 * i], vy = (double)vel[3 * i + 1], vz = (double)vel[3 * i + 2];
        const double v2 = vx * vx + vy * vy + vz * vz;
        const bool in_g = (group == nullptr || g < 0) ? true : (group[i] == g);
        if (in_g) { acc[0] += 1.0; acc[1] += vx; acc[2] += vy; acc[3] += vz; acc[4] += v2; }
        acc[5] += 1.0; acc[6] += vx; acc[7] += vy; acc[8] += vz; acc[9] += v2;
    }
    __shared__ double sh[2 * MOM][MD_THREADS];
    for (int k = 0; k < 2 * MOM; ++k) sh[k][threadIdx.x] = acc[k];
    __syncthreads();
    for (int w = MD_THREADS / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w)
            for (int k = 0; k < 2 * MOM; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x < 2 * MOM) partial[2 * MOM * blockIdx.x + threadIdx.x] = sh[threadIdx.x][0];
}

__global__ void __launch_bounds__(MD_THREADS) moments_final_kernel(const double* __restrict__ partial,
                                                                   int nblocks, double* __restrict__ out) {
    __shared__ double sh[2 * MOM][MD_THREADS];
    double acc[2 * MOM] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < nblocks; i += MD_THREADS)
        for (int k = 0; k < 2 * MOM; ++k) acc[k] += partial[2 * MOM * i + k];
    for (int k = 0; k < 2 * MOM; ++k) sh[k][threadIdx.x] = acc[k];
    __syncthreads();
    for (int w = MD_THREADS / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w)
            for (int k = 0; k < 2 * MOM; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x < 2 * MOM) out[threadIdx.x] = sh[threadIdx.x][0];
}

// csvr_thermostat for one coupling group (thermostat.py:184-219), alpha evaluated on the device from
// the reduced moments so the host never waits for the kinetic energy.
template <typename real>
__global__ void __launch_bounds__(MD_THREADS) csvr_apply_kernel(
    real* __restrict__ vel, const int32_t* __restrict__ group, int g, long long n,
    const double* __restrict__ mom, double mass, double kT15, double c, double R, double SNf,
    int remove_com, double* __restrict__ work) {
    double dK;
    const CsvrScale s = csvr_scale(mom, mass, kT15, c, R, SNf, remove_com, &dK);
    const long long stride = (long long)gridDim.x * MD_THREADS;
    for (long long i = (long long)blockIdx.x * MD_THREADS + threadIdx.x; i < n; i += stride) {
        const bool in_g = (group == nullptr || g < 0) ? true : (group[i] == g);
        csvr_apply_particle(vel + 3 * i, s, in_g);
    }
    if (work != nullptr && blockIdx.x == 0 && threadIdx.x == 0) work[0] += dK;
}

// cancel_com_momentum (thermostat.py:12-15): v <- v - (sum v) / n_particles, sums = mom[6..8].
template <typename real>
__global__ void __launch_bounds__(MD_THREADS) cancel_com_kernel(real* __restrict__ vel, long long n,
                                                                const double* __restrict__ mom,
                                                                double n_particles) {
    const double cx = mom[MOM + 1] / n_particles, cy = mom[MOM + 2] / n_particles,
                 cz = mom[MOM + 3] / n_particles;
    const long long stride = (long long)gridDim.x * MD_THREADS;
    for (long long i = (long long)blockIdx.x * MD_THREADS + threadIdx.x; i < n; i += stride) {
        vel[3 * i + 0] = (real)((double)vel[3 * i + 0] - cx);
        vel[3 * i + 1] = (real)((double)vel[3 * i + 1] - cy);
        vel[3 * i + 2] = (real)((double)vel[3 * i + 2] - cz);
    }
}

static int grid_for(long long work) {
    long long b = (work + MD_THREADS - 1) / MD_THREADS;
    const long long cap = 148LL * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

static int launch_ok(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: kernel launch -> %s", what, cudaGetErrorString(e));
        return HYMD_ERR_CUDA;
    }
    return HYMD_OK;
}

}  // namespace hymd

using namespace hymd;

extern "C" {

int hymd_md_kick_drift(int dtype, void* d_vel, void* d_pos, const void* const* d_forces, int n_forces,
                       int sequential, double mass, double kick_dt, double drift_dt, const double box[3],
                       int64_t n, void* stream) {
    if ((n > 0 && !d_vel) || n_forces < 0 || n_forces > MD_MAX_FORCES || (n_forces && !d_forces) ||
        (d_pos && !box) || n < 0) {
        set_error("hymd_md_kick_drift: bad argument (n_forces <= %d)", MD_MAX_FORCES);
        return HYMD_ERR_INVALID;
    }
    if (dtype != HYMD_F32 && dtype != HYMD_F64) { set_error("bad dtype %d", dtype); return HYMD_ERR_INVALID; }
    if (n == 0 || (n_forces == 0 && !d_pos)) return HYMD_OK;
    ForcePtrs fp;
    for (int k = 0; k < MD_MAX_FORCES; ++k) fp.f[k] = k < n_forces ? d_forces[k] : nullptr;
    for (int k = 0; k < n_forces; ++k)
        if (!fp.f[k]) { set_error("hymd_md_kick_drift: null force array %d", k); return HYMD_ERR_INVALID; }
    const long long n3 = 3 * (long long)n;
    cudaStream_t s = (cudaStream_t)stream;
    const double L[3] = {box ? box[0] : 1.0, box ? box[1] : 1.0, box ? box[2] : 1.0};
    if (dtype == HYMD_F64)
        kick_drift_kernel<double><<<grid_for(n3), MD_THREADS, 0, s>>>(
            (double*)d_vel, (double*)d_pos, fp, n_forces, sequential, mass, 0.5 * kick_dt, drift_dt, L[0],
            L[1], L[2], n3);
    else
        kick_drift_kernel<float><<<grid_for(n3), MD_THREADS, 0, s>>>(
            (float*)d_vel, (float*)d_pos, fp, n_forces, sequential, (float)mass, (float)(0.5 * kick_dt),
            (float)drift_dt, (float)L[0], (float)L[1], (float)L[2], n3);
    return launch_ok("hymd_md_kick_drift");
}

int hymd_velocity_moments(int dtype, const void* d_vel, const int32_t* d_group, int group, int64_t n,
                          double* d_scratch, double* d_out, void* stream) {
    if (!d_out || !d_scratch || (n > 0 && !d_vel) || n < 0) { set_error("hymd_velocity_moments: null argument"); return HYMD_ERR_INVALID; }
    if (dtype != HYMD_F32 && dtype != HYMD_F64) { set_error("bad dtype %d", dtype); return HYMD_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == HYMD_F64)
        moments_partial_kernel<double><<<MD_MOM_BLOCKS, MD_THREADS, 0, s>>>((const double*)d_vel, d_group, group, n, d_scratch);
    else
        moments_partial_kernel<float><<<MD_MOM_BLOCKS, MD_THREADS, 0, s>>>((const float*)d_vel, d_group, group, n, d_scratch);
    HYMD_CHECK(launch_ok("hymd_velocity_moments"));
    moments_final_kernel<<<1, MD_THREADS, 0, s>>>(d_scratch, MD_MOM_BLOCKS, d_out);
    return launch_ok("hymd_velocity_moments");
}

int64_t hymd_velocity_moments_scratch_doubles(void) { return (int64_t)MD_MOM_BLOCKS * 2 * MOM; }

int hymd_csvr_apply(int dtype, void* d_vel, const int32_t* d_group, int group, int64_t n,
                    const double* d_moments, double mass, double kT15, double c, double R, double SNf,
                    int remove_com, double* d_work, void* stream) {
    if (!d_moments || (n > 0 && !d_vel) || n < 0) { set_error("hymd_csvr_apply: null argument"); return HYMD_ERR_INVALID; }
    if (dtype != HYMD_F32 && dtype != HYMD_F64) { set_error("bad dtype %d", dtype); return HYMD_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == HYMD_F64)
        csvr_apply_kernel<double><<<grid_for(n), MD_THREADS, 0, s>>>((double*)d_vel, d_group, group, n, d_moments,
                                                                   mass, kT15, c, R, SNf, remove_com, d_work);
    else
        csvr_apply_kernel<float><<<grid_for(n), MD_THREADS, 0, s>>>((float*)d_vel, d_group, group, n, d_moments,
                                                                  mass, kT15, c, R, SNf, remove_com, d_work);
    return launch_ok("hymd_csvr_apply");
}

int hymd_cancel_com(int dtype, void* d_vel, int64_t n, const double* d_moments, double n_particles,
                    void* stream) {
    if (!d_moments || (n > 0 && !d_vel) || n < 0 || !(n_particles > 0)) { set_error("hymd_cancel_com: bad argument"); return HYMD_ERR_INVALID; }
    if (dtype != HYMD_F32 && dtype != HYMD_F64) { set_error("bad dtype %d", dtype); return HYMD_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == HYMD_F64)
        cancel_com_kernel<double><<<grid_for(n), MD_THREADS, 0, s>>>((double*)d_vel, n, d_moments, n_particles);
    else
        cancel_com_kernel<float><<<grid_for(n), MD_THREADS, 0, s>>>((float*)d_vel, n, d_moments, n_particles);
    return launch_ok("hymd_cancel_com");
}

}  // extern "C"
