// Cell binning of the local particles: replaces pm.decompose (main.py:977-980, 1007).
//
// A counting sort keyed by the (local-slab, row-major) z-bin of each particle's mesh cell (ctx.cuh,
// ZBIN: two bins per 32 cells of an (x,y) cell row), on ONE array a[0 .. nbins] (a[0] stays 0,
// cur = a + 1):
//   count   : atomicAdd(cur[key], 1) per particle                        -> cur[k] = count of cell k
//   scan    : exclusive prefix sum of cur[0 .. ncell) in place           -> cur[k] = start of cell k
//   scatter : slot = atomicAdd(cur[key], 1); record[slot] = fixed-point coordinates | index | type
// after which cur[k] = start[k] + count[k] = start[k+1], i.e. a[] IS the cell_start array that
// paint and readout index.  No per-particle key / rank arrays are written or re-read.
// The order of particles inside a cell depends on atomic arrival order; paint accumulates in
// integer fixed point (order independent) and readout is a pure gather, so results are
// bitwise reproducible anyway.
#include <cub/device/device_scan.cuh>

#include "ctx.cuh"

namespace hymd {

struct SortParams {
    int Nx, Ny, Nz, nxl, x0, nbz;
    int fbx, fby, fbz;
    double sx, sy, sz;  // N/L per axis
};

__device__ __forceinline__ void split_coord(double x, int n, int& cell, double& frac) {
    double f = floor(x);
    double d = x - f;
    long long c = (long long)f % n;
    if (c < 0) c += n;
    if (d >= 1.0) {  // x = -tiny rounds to d == 1
        d = 0.0;
        c = (c + 1 == n) ? 0 : c + 1;
    }
    cell = (int)c;
    frac = d;
}

template <typename UT>
__device__ __forceinline__ UT pack_coord(int cell, double frac, int fb) {
    UT f = (UT)(frac * (double)((UT)1 << fb));
    if (f >> fb) f = ((UT)1 << fb) - 1;  // frac*2^fb rounded up to 2^fb
    return ((UT)cell << fb) | f;
}

// Consecutive lanes with the same key form a run (in REUSE mode the lanes walk the previous bin
// order, so a warp usually holds two or three runs): one counter atomic per run instead of one per
// lane.  Returns this lane's run head, its rank inside the run and the run length.
__device__ __forceinline__ void warp_runs(uint32_t key, unsigned& head_lane, unsigned& rank, unsigned& count) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = lane == 0 || key != prev;
    const unsigned H = __ballot_sync(0xffffffffu, head);
    const unsigned upto = (2u << lane) - 1u;          // lanes 0 .. lane (lane 31: all)
    head_lane = 31u - (unsigned)__clz((int)(H & upto));
    const unsigned above = H & ~upto;
    const unsigned next = above ? (unsigned)__ffs((int)above) - 1u : 32u;
    rank = lane - head_lane;
    count = next - head_lane;
}

// Pass 1.  Thread j handles one particle: in REUSE mode the particle that sat at sorted slot j
// in the previous call (its index and type come from the previous record), otherwise particle j
// of the caller's arrays.  It converts the CURRENT position to the fixed-point record, leaves it
// in stage[j] (in place over the previous record) and counts its cell.
template <typename real, typename RecT, typename UT, int IDX_BITS, bool REUSE>
__global__ void __launch_bounds__(256) count_kernel(const real* __restrict__ pos,
                                                    const int32_t* __restrict__ types, long long n,
                                                    SortParams p, RecT* __restrict__ stage,
                                                    uint32_t* __restrict__ cnt,
                                                    DeviceScalars* __restrict__ sc) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    unsigned int r1 = 0, bad = 0;
    uint32_t key = 0xffffffffu;                        // lanes past the end: a run of their own
    if (j < n) {
        UT idx, type;
        if (REUSE) {
            const UT meta = stage[j].meta;
            idx = meta & (((UT)1 << IDX_BITS) - 1);
            type = meta >> IDX_BITS;
        } else {
            idx = (UT)j;
            type = (UT)(uint32_t)types[j];
        }
        int cx, cy, cz;
        double dx, dy, dz;
        split_coord((double)pos[3 * idx + 0] * p.sx, p.Nx, cx, dx);
        split_coord((double)pos[3 * idx + 1] * p.sy, p.Ny, cy, dy);
        split_coord((double)pos[3 * idx + 2] * p.sz, p.Nz, cz, dz);
        int lx = cx - p.x0;
        if (lx < 0 || lx >= p.nxl) {
            bad = 1;
            lx = lx < 0 ? 0 : p.nxl - 1;
        }
        RecT r;
        r.ux = pack_coord<UT>(lx, dx, p.fbx);
        r.uy = pack_coord<UT>(cy, dy, p.fby);
        r.uz = pack_coord<UT>(cz, dz, p.fbz);
        r.meta = idx | (type << IDX_BITS);
        stage[j] = r;
        key = (uint32_t)(((long long)lx * p.Ny + cy) * p.nbz + zbin_of(cz, p.Nz));
    }
    unsigned head_lane, rank, count;
    warp_runs(key, head_lane, rank, count);
    if (rank == 0 && j < n) r1 = atomicAdd(&cnt[key], count) + count;
    unsigned int m = __reduce_max_sync(0xffffffffu, r1);
    unsigned int b = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
        if (m > sc->max_cell_count) atomicMax(&sc->max_cell_count, m);
        if (b) atomicAdd(&sc->out_of_slab, b);
    }
}

// Pass 2 (after the scan): staged record j goes to the next free slot of its cell.
template <typename real, typename RecT, typename UT, int IDX_BITS>
__global__ void __launch_bounds__(256) scatter_kernel(
    const RecT* __restrict__ stage, const real* __restrict__ q, long long n, SortParams p,
    uint32_t* __restrict__ cur, RecT* __restrict__ rec, real* __restrict__ q_sorted,
    DeviceScalars* __restrict__ sc) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    float aq = 0.f;
    RecT r;
    uint32_t key = 0xffffffffu;
    if (j < n) {
        r = stage[j];
        const long long lx = (long long)(r.ux >> p.fbx), cy = (long long)(r.uy >> p.fby),
                        cz = (long long)(r.uz >> p.fbz);
        key = (uint32_t)((lx * p.Ny + cy) * p.nbz + zbin_of((int)cz, p.Nz));
    }
    unsigned head_lane, rank, count;
    warp_runs(key, head_lane, rank, count);
    uint32_t base = 0;
    if (rank == 0 && j < n) base = atomicAdd(&cur[key], count);
    base = __shfl_sync(0xffffffffu, base, (int)head_lane);
    if (j < n) {
        const size_t slot = (size_t)base + rank;
        rec[slot] = r;
        if (q != nullptr) {
            const real qi = q[r.meta & (((UT)1 << IDX_BITS) - 1)];
            q_sorted[slot] = qi;
            aq = fabsf((float)qi);
        }
    }
    if (q != nullptr) {
        unsigned int m = __reduce_max_sync(0xffffffffu, __float_as_uint(aq));
        if ((threadIdx.x & 31) == 0 && m > sc->qmax_bits) atomicMax(&sc->qmax_bits, m);
    }
}

// Charges into sorted order for a sort that was made without them (update_field_force_q is
// called after update_field on the same positions: main.py:1006-1058).
template <typename real, typename RecT, typename UT, int IDX_BITS>
__global__ void __launch_bounds__(256) gather_charges_kernel(const RecT* __restrict__ rec,
                                                             const real* __restrict__ q, long long n,
                                                             real* __restrict__ q_sorted,
                                                             DeviceScalars* __restrict__ sc) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    float aq = 0.f;
    if (i < n) {
        const UT idx = rec[i].meta & (((UT)1 << IDX_BITS) - 1);
        const real qi = q[idx];
        q_sorted[i] = qi;
        aq = fabsf((float)qi);
    }
    unsigned int m = __reduce_max_sync(0xffffffffu, __float_as_uint(aq));
    if ((threadIdx.x & 31) == 0 && m > sc->qmax_bits) atomicMax(&sc->qmax_bits, m);
}

int gather_charges(hymd_ctx* c, const void* d_q, cudaStream_t s) {
    const long long n = c->np;
    if (n == 0) return HYMD_OK;
    const unsigned int blocks = (unsigned int)((n + 255) / 256);
    if (c->f64)
        gather_charges_kernel<double, Rec64, unsigned long long, REC64_IDX_BITS>
            <<<blocks, 256, 0, s>>>((const Rec64*)c->rec, (const double*)d_q, n,
                                    (double*)c->q_sorted, c->scalars);
    else
        gather_charges_kernel<float, Rec32, uint32_t, REC32_IDX_BITS>
            <<<blocks, 256, 0, s>>>((const Rec32*)c->rec, (const float*)d_q, n,
                                    (float*)c->q_sorted, c->scalars);
    HYMD_LAUNCH_CHECK(c);
    return HYMD_OK;
}

size_t scan_temp_bytes(long long n) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
    return bytes;
}

template <typename real, typename RecT, typename UT, int IDX_BITS>
static int sort_impl(hymd_ctx* c, const void* d_pos, const int32_t* d_types, const void* d_q,
                     int64_t n, bool reuse, cudaStream_t s) {
    const Geometry& g = c->g;
    SortParams p;
    p.Nx = g.Nx; p.Ny = g.Ny; p.Nz = g.Nz; p.nxl = g.nxl; p.x0 = g.x0; p.nbz = g.nbz;
    p.fbx = g.fbx; p.fby = g.fby; p.fbz = g.fbz;
    p.sx = g.Nx / g.box[0]; p.sy = g.Ny / g.box[1]; p.sz = g.Nz / g.box[2];
    const long long ncell = g.ncell;
    uint32_t* cur = c->cell_start + 1;
    HYMD_CUDA(cudaMemsetAsync(c->cell_start, 0, (size_t)(ncell + 2) * sizeof(uint32_t), s));
    HYMD_CUDA(cudaMemsetAsync(c->scalars, 0, sizeof(DeviceScalars), s));
    // stage = the buffer holding the previous sorted records (overwritten in place), out = the other
    RecT* stage = (RecT*)c->rec;
    RecT* out = (RecT*)c->rec_alt;
    const unsigned int blocks = (unsigned int)((n + 255) / 256);
    if (n > 0) {
        if (reuse)
            count_kernel<real, RecT, UT, IDX_BITS, true><<<blocks, 256, 0, s>>>(
                (const real*)d_pos, d_types, n, p, stage, cur, c->scalars);
        else
            count_kernel<real, RecT, UT, IDX_BITS, false><<<blocks, 256, 0, s>>>(
                (const real*)d_pos, d_types, n, p, stage, cur, c->scalars);
        HYMD_LAUNCH_CHECK(c);
    }
    size_t tmp = c->scan_tmp_bytes;
    HYMD_CUDA(cub::DeviceScan::ExclusiveSum(c->scan_tmp, tmp, cur, cur, (int)ncell, s));
    c->launches += 2;  // cub scan: init + scan kernels
    if (n > 0) {
        scatter_kernel<real, RecT, UT, IDX_BITS><<<blocks, 256, 0, s>>>(
            stage, (const real*)d_q, n, p, cur, out, (real*)c->q_sorted, c->scalars);
        HYMD_LAUNCH_CHECK(c);
    }
    c->rec = out;
    c->rec_alt = stage;
    return HYMD_OK;
}

// reuse: start from the order of the previous call (same n, same per-index types): consecutive
// MD steps move particles by a fraction of a cell, so the staged records are almost sorted and
// the counter atomics and record writes of both passes hit neighbouring addresses.
int sort_particles(hymd_ctx* c, const void* d_pos, const int32_t* d_types, const void* d_q,
                   int64_t n, bool reuse, cudaStream_t s) {
    return c->f64 ? sort_impl<double, Rec64, unsigned long long, REC64_IDX_BITS>(c, d_pos, d_types, d_q, n, reuse, s)
                  : sort_impl<float, Rec32, uint32_t, REC32_IDX_BITS>(c, d_pos, d_types, d_q, n, reuse, s);
}

}  // namespace hymd
