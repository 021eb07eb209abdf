// Field energy reductions: compute_field_and_kinetic_energy (field.py:688-703).
//   E_field = dV * sum_cells [ sum_{i<j} chi_ij phi~_i phi~_j / rho0 + (sum_t phi~_t - a)^2 / (2 kappa rho0) ]
//   E_q     = dV * sum_cells 0.5 * phi_q * psi            (self energy subtracted by the caller)
// Two-stage reduction in double with a fixed combination order -> bitwise reproducible.
#include "ctx.cuh"

namespace hymd {

constexpr int EN_BLOCKS = 148 * 4;

template <typename real>
__global__ void __launch_bounds__(256) energy_partial_kernel(
    const real* __restrict__ phi, long long field_stride, long long n, int T,
    const double* __restrict__ chi, double inv_rho0, double half_inv_kr, double a,
    const real* __restrict__ phi_q, const real* __restrict__ psi, double* __restrict__ partial) {
    double e0 = 0.0, e1 = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        if (phi != nullptr) {
            double tot = 0.0, inter = 0.0;
            for (int t = 0; t < T; ++t) {
                const double pt = (double)phi[t * field_stride + i];
                tot += pt;
                for (int u = t + 1; u < T; ++u) {
                    const double ch = chi[t * T + u];
                    if (ch != 0.0) inter += ch * pt * (double)phi[u * field_stride + i];
                }
            }
            const double d = tot - a;
            e0 += inter * inv_rho0 + half_inv_kr * d * d;
        }
        if (phi_q != nullptr) e1 += 0.5 * (double)phi_q[i] * (double)psi[i];
    }
    __shared__ double s0[256], s1[256];
    s0[threadIdx.x] = e0; s1[threadIdx.x] = e1;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) { s0[threadIdx.x] += s0[threadIdx.x + w]; s1[threadIdx.x] += s1[threadIdx.x + w]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = s0[0]; partial[2 * blockIdx.x + 1] = s1[0]; }
}

__global__ void energy_final_kernel(const double* __restrict__ partial, int nblocks, double dv,
                                    double* __restrict__ out) {
    __shared__ double s0[256], s1[256];
    double e0 = 0.0, e1 = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256) { e0 += partial[2 * i]; e1 += partial[2 * i + 1]; }
    s0[threadIdx.x] = e0; s1[threadIdx.x] = e1;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) { s0[threadIdx.x] += s0[threadIdx.x + w]; s1[threadIdx.x] += s1[threadIdx.x + w]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = s0[0] * dv; out[1] = s1[0] * dv; }
}

// Field terms of comp_pressure (pressure.py:105-127): per cell, with V_t = c_t + sum_j A_tj phi~_j
// (+ q_t psi with electrostatics; V_bar_0 / V_bar of hamiltonian.py:157-186, 271-301, 423-475),
//   s0 = sum_t V_t phi~_t,   s_{1+d} = sum_t V_t lap[t][d];
// same fixed-order two-stage double reduction as the energies.
constexpr int PR_TERMS = 4;

template <typename real>
__global__ void __launch_bounds__(256) pressure_partial_kernel(
    const real* __restrict__ phi, const real* __restrict__ lap, long long field_stride, long long n,
    int T, const double* __restrict__ A, const double* __restrict__ cc, const double* __restrict__ qt,
    const real* __restrict__ psi, double* __restrict__ partial) {
    double acc[PR_TERMS] = {0.0, 0.0, 0.0, 0.0};
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const double ps = (qt != nullptr) ? (double)psi[i] : 0.0;
        for (int t = 0; t < T; ++t) {
            double v = cc[t];
            for (int j = 0; j < T; ++j) v += A[t * T + j] * (double)phi[j * field_stride + i];
            if (qt != nullptr) v += qt[t] * ps;
            acc[0] += v * (double)phi[t * field_stride + i];
#pragma unroll
            for (int d = 0; d < 3; ++d) acc[1 + d] += v * (double)lap[(3 * t + d) * field_stride + i];
        }
    }
    __shared__ double sh[PR_TERMS][256];
#pragma unroll
    for (int k = 0; k < PR_TERMS; ++k) sh[k][threadIdx.x] = acc[k];
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
#pragma unroll
            for (int k = 0; k < PR_TERMS; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x < PR_TERMS) partial[PR_TERMS * blockIdx.x + threadIdx.x] = sh[threadIdx.x][0];
}

__global__ void pressure_final_kernel(const double* __restrict__ partial, int nblocks, double* __restrict__ out) {
    __shared__ double sh[PR_TERMS][256];
    double acc[PR_TERMS] = {0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < nblocks; i += 256)
        for (int k = 0; k < PR_TERMS; ++k) acc[k] += partial[PR_TERMS * i + k];
    for (int k = 0; k < PR_TERMS; ++k) sh[k][threadIdx.x] = acc[k];
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w)
            for (int k = 0; k < PR_TERMS; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x < PR_TERMS) out[threadIdx.x] = sh[threadIdx.x][0];
}

int field_pressure(hymd_ctx* c, const double* A, const double* cc, const double* qt, double out[4],
                   cudaStream_t s) {
    const Geometry& g = c->g;
    const long long n = (long long)g.nxl * g.Ny * g.Nz;
    const int T = c->T;
    double* scratch = nullptr;   // [T*T A][T c][T q][4*EN_BLOCKS partial][4 out]
    const size_t nd = (size_t)T * T + 2 * T + PR_TERMS * EN_BLOCKS + PR_TERMS;
    HYMD_CUDA(cudaMallocAsync((void**)&scratch, sizeof(double) * nd, s));
    double* d_A = scratch;
    double* d_c = d_A + (size_t)T * T;
    double* d_q = d_c + T;
    double* d_part = d_q + T;
    double* d_out = d_part + PR_TERMS * EN_BLOCKS;
    HYMD_CUDA(cudaMemcpyAsync(d_A, A, sizeof(double) * T * T, cudaMemcpyHostToDevice, s));
    HYMD_CUDA(cudaMemcpyAsync(d_c, cc, sizeof(double) * T, cudaMemcpyHostToDevice, s));
    if (qt) HYMD_CUDA(cudaMemcpyAsync(d_q, qt, sizeof(double) * T, cudaMemcpyHostToDevice, s));
    if (c->f64)
        pressure_partial_kernel<double><<<EN_BLOCKS, 256, 0, s>>>(
            (const double*)c->phi, (const double*)c->lap, g.real_elems, n, T, d_A, d_c,
            qt ? d_q : nullptr, (const double*)c->psi, d_part);
    else
        pressure_partial_kernel<float><<<EN_BLOCKS, 256, 0, s>>>(
            (const float*)c->phi, (const float*)c->lap, g.real_elems, n, T, d_A, d_c,
            qt ? d_q : nullptr, (const float*)c->psi, d_part);
    HYMD_LAUNCH_CHECK(c);
    pressure_final_kernel<<<1, 256, 0, s>>>(d_part, EN_BLOCKS, d_out);
    HYMD_LAUNCH_CHECK(c);
    HYMD_CUDA(cudaMemcpyAsync(out, d_out, PR_TERMS * sizeof(double), cudaMemcpyDeviceToHost, s));
    HYMD_CUDA(cudaStreamSynchronize(s));
    HYMD_CUDA(cudaFreeAsync(scratch, s));
    return HYMD_OK;
}

int field_energy(hymd_ctx* c, const double* chi, double kappa, double rho0, double a,
                 double out[2], cudaStream_t s) {
    const Geometry& g = c->g;
    const long long n = (long long)g.nxl * g.Ny * g.Nz;   // owned planes only (no ghost plane)
    double* scratch = nullptr;   // [T*T chi][2*EN_BLOCKS partial][2 out]
    const size_t bytes = sizeof(double) * ((size_t)c->T * c->T + 2 * EN_BLOCKS + 2);
    HYMD_CUDA(cudaMallocAsync((void**)&scratch, bytes, s));
    double* d_chi = scratch;
    double* d_part = scratch + (size_t)c->T * c->T;
    double* d_out = d_part + 2 * EN_BLOCKS;
    HYMD_CUDA(cudaMemcpyAsync(d_chi, chi, sizeof(double) * c->T * c->T, cudaMemcpyHostToDevice, s));
    const bool have_phi = c->phi_is_filtered;
    const bool have_q = c->cfg.pme && c->psi != nullptr;
    const double dv = g.box[0] * g.box[1] * g.box[2] / ((double)g.Nx * g.Ny * g.Nz);
    if (c->f64)
        energy_partial_kernel<double><<<EN_BLOCKS, 256, 0, s>>>(
            have_phi ? (const double*)c->phi : nullptr, g.real_elems, n, c->T, d_chi, 1.0 / rho0,
            0.5 / (kappa * rho0), a, have_q ? (const double*)c->phi_q : nullptr,
            (const double*)c->psi, d_part);
    else
        energy_partial_kernel<float><<<EN_BLOCKS, 256, 0, s>>>(
            have_phi ? (const float*)c->phi : nullptr, g.real_elems, n, c->T, d_chi, 1.0 / rho0,
            0.5 / (kappa * rho0), a, have_q ? (const float*)c->phi_q : nullptr,
            (const float*)c->psi, d_part);
    HYMD_LAUNCH_CHECK(c);
    energy_final_kernel<<<1, 256, 0, s>>>(d_part, EN_BLOCKS, dv, d_out);
    HYMD_LAUNCH_CHECK(c);
    HYMD_CUDA(cudaMemcpyAsync(out, d_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    HYMD_CUDA(cudaStreamSynchronize(s));
    HYMD_CUDA(cudaFreeAsync(scratch, s));
    return HYMD_OK;
}

}  // namespace hymd
