// Force readout: compute_field_force (field.py:197-200) and the readout half of
// update_field_force_q (field.py:400-403).
//
// The force meshes live in a ghost-padded layout (nxl+1, Ny+1, Nzp): the planes x = nxl, y = Ny
// and z = Nz hold the periodic images (or, with several GPUs, the first plane of the next slab),
// so the (tile + 1)^3 neighbourhood a tile of cells needs is ONE dense box.  A CTA takes a tile
// of cells, pulls that box for all three force components (and every distinct potential row)
// into shared memory with TMA (cp.async.bulk.tensor, completion on an mbarrier), then walks the
// tile's cell-sorted particles -- one warp per (x,y) row, whose particles are contiguous -- and
// interpolates from shared memory.  Results are written to the caller's (n,3) array at the
// particle's original index, so the caller's order is preserved.
#include <stdlib.h>

#include "ctx.cuh"

namespace hymd {

struct ReadoutParams {
    int Nx, Ny, Nz, nxl;
    int fbx, fby, fbz;
    int tx, ty, tz, bz;      // tile of cells and box z extent (tz + 1 rounded up for TMA)
    int ntx, nty, ntz;
    int U, T;
    unsigned int box_bytes;  // bytes of one TMA box (3 components)
    unsigned int box_stride; // box_bytes rounded up to 128 B (TMA destination alignment)
    int debug_seq;           // TIMING EXPERIMENT ONLY: write forces in sorted order
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

template <typename real> struct RTraits;
template <> struct RTraits<float> {
    using Rec = Rec32; using UT = uint32_t;
    static constexpr int IDX_BITS = REC32_IDX_BITS;
};
template <> struct RTraits<double> {
    using Rec = Rec64; using UT = unsigned long long;
    static constexpr int IDX_BITS = REC64_IDX_BITS;
};

constexpr int READOUT_MAX_ROWS = 64;   // tx * ty of the largest tile

template <typename real, bool CHARGE>
__global__ void __launch_bounds__(256) readout_kernel(
    const __grid_constant__ CUtensorMap tmap, const typename RTraits<real>::Rec* __restrict__ rec,
    const real* __restrict__ q_sorted, const uint32_t* __restrict__ start,
    const int* __restrict__ urow, real* __restrict__ force, ReadoutParams p) {
    using Tr = RTraits<real>;
    using UT = typename Tr::UT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    real* S = reinterpret_cast<real*>(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)p.U * p.box_stride);
    int* s_urow = reinterpret_cast<int*>(bar + 1);
    __shared__ uint32_t s_begin[READOUT_MAX_ROWS];
    __shared__ uint32_t s_off[READOUT_MAX_ROWS];
    __shared__ uint32_t s_wsum[2];

    int b = blockIdx.x;
    const int tz_i = b % p.ntz; b /= p.ntz;
    const int ty_i = b % p.nty; b /= p.nty;
    const int tx_i = b;
    const int x0 = tx_i * p.tx, y0 = ty_i * p.ty, z0 = tz_i * p.tz;
    const int rows = p.tx * p.ty;
    const int zb = min(z0 + p.tz, p.Nz);

    // the TMA boxes of all potential rows fly while the particle runs of the tile are located
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar, (uint32_t)p.U * p.box_bytes);
        for (int u = 0; u < p.U; ++u)
            tma_load_4d(smem_raw + (size_t)u * p.box_stride, &tmap, bar, z0, y0, x0, 3 * u);
    }
    if (!CHARGE)
        for (int i = threadIdx.x; i < p.T; i += blockDim.x) s_urow[i] = urow[i];
    // one (x,y) row of the tile = one contiguous run of the cell-sorted records
    if (threadIdx.x < 64) {
        const int r = threadIdx.x;
        uint32_t pa = 0, len = 0;
        if (r < rows) {
            const int gx = x0 + r / p.ty, gy = y0 + r % p.ty;
            if (gx < p.nxl && gy < p.Ny) {
                const long long rowbase = ((long long)gx * p.Ny + gy) * p.Nz;
                pa = start[rowbase + z0];
                len = start[rowbase + zb] - pa;
            }
        }
        // exclusive prefix sum over the 64 rows (two warps, combined through shared memory)
        uint32_t inc = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc, d);
            if ((r & 31) >= d) inc += v;
        }
        if ((r & 31) == 31) s_wsum[r >> 5] = inc;
        s_begin[r] = pa;
        s_off[r] = inc - len;                             // exclusive within its warp
    }
    __syncthreads();
    if (threadIdx.x >= 32 && threadIdx.x < 64) s_off[threadIdx.x] += s_wsum[0];
    __syncthreads();
    const uint32_t total = s_wsum[0] + s_wsum[1];

    const UT mx = ((UT)1 << p.fbx) - 1, my = ((UT)1 << p.fby) - 1, mz = ((UT)1 << p.fbz) - 1;
    const real ifx = (real)1 / (real)((UT)1 << p.fbx), ify = (real)1 / (real)((UT)1 << p.fby),
               ifz = (real)1 / (real)((UT)1 << p.fbz);
    const UT idx_mask = ((UT)1 << Tr::IDX_BITS) - 1;
    const int py = p.ty + 1, px = p.tx + 1;
    const int comp_stride = px * py * p.bz;                 // elements between components
    const int box_elems = (int)(p.box_stride / sizeof(real));

    bool waited = false;
    for (uint32_t j = threadIdx.x; j < total; j += blockDim.x) {
        // row of particle j: largest r with s_off[r] <= j  (64 entries: 6 halvings)
        int r = 0;
#pragma unroll
        for (int step = 32; step > 0; step >>= 1)
            if (s_off[r + step] <= j && r + step < READOUT_MAX_ROWS) r += step;
        const uint32_t i = s_begin[r] + (j - s_off[r]);
        const typename Tr::Rec rc = rec[i];
        real q = (real)1;
        if (CHARGE) q = q_sorted[i];
        const int lx = r / p.ty, ly = r % p.ty;
        const real dx = (real)(rc.ux & mx) * ifx, dy = (real)(rc.uy & my) * ify,
                   dz = (real)(rc.uz & mz) * ifz;
        const int lz = (int)(rc.uz >> p.fbz) - z0;
        int u = 0;
        if (!CHARGE) u = s_urow[(int)(rc.meta >> Tr::IDX_BITS)];
        if (!waited) { mbar_wait(bar, 0); waited = true; }
        const real* B = S + (size_t)u * box_elems + (lx * py + ly) * p.bz + lz;
        real f0 = 0, f1 = 0, f2 = 0;
#pragma unroll
        for (int ax = 0; ax < 2; ++ax) {
            const real wx = ax ? dx : (real)1 - dx;
#pragma unroll
            for (int ay = 0; ay < 2; ++ay) {
                const real wxy = wx * (ay ? dy : (real)1 - dy);
                const real* c = B + (ax * py + ay) * p.bz;
                const real w0 = wxy * ((real)1 - dz), w1 = wxy * dz;
                f0 += w0 * c[0] + w1 * c[1];
                f1 += w0 * c[comp_stride] + w1 * c[comp_stride + 1];
                f2 += w0 * c[2 * comp_stride] + w1 * c[2 * comp_stride + 1];
            }
        }
        if (CHARGE) { f0 *= q; f1 *= q; f2 *= q; }
        const size_t o = p.debug_seq ? (size_t)i * 3 : (size_t)(rc.meta & idx_mask) * 3;
        force[o] = f0; force[o + 1] = f1; force[o + 2] = f2;
    }
    // the CTA must not retire with the bulk copies still in flight
    if (!waited) mbar_wait(bar, 0);
}

// Periodic images into the ghost planes of nfields ghost-padded meshes (single-GPU x; y and z
// always).  dest (x,y,z) with x == nxl or y == Ny or z == Nz  <-  src (x%nxl, y%Ny, z%Nz).
template <typename real>
__global__ void __launch_bounds__(256) fill_ghost_kernel(real* __restrict__ mesh, int nfields,
                                                         int nxl, int Ny, int Nz, int Nzp,
                                                         int fill_x, long long ghost_elems) {
    const long long nA = fill_x ? (long long)(Ny + 1) * (Nz + 1) : 0;   // x = nxl plane
    const long long nB = (long long)nxl * (Nz + 1);                     // y = Ny rows, x < nxl
    const long long nC = (long long)nxl * Ny;                           // z = Nz, x < nxl, y < Ny
    const long long per = nA + nB + nC;
    const long long total = per * nfields;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
        const int f = (int)(i / per);
        long long j = i % per;
        int x, y, z;
        if (j < nA) { x = nxl; y = (int)(j / (Nz + 1)); z = (int)(j % (Nz + 1)); }
        else if ((j -= nA) < nB) { x = (int)(j / (Nz + 1)); y = Ny; z = (int)(j % (Nz + 1)); }
        else { j -= nB; x = (int)(j / Ny); y = (int)(j % Ny); z = Nz; }
        real* m = mesh + (long long)f * ghost_elems;
        const int sx = x == nxl ? 0 : x, sy = y == Ny ? 0 : y, sz = z == Nz ? 0 : z;
        m[((long long)x * (Ny + 1) + y) * Nzp + z] = m[((long long)sx * (Ny + 1) + sy) * Nzp + sz];
    }
}

int fill_ghosts(hymd_ctx* c, void* mesh, int nfields, cudaStream_t s) {
    const Geometry& g = c->g;
    const int fill_x = g.P == 1 ? 1 : 0;
    const long long per = (long long)(g.Ny + 1) * (g.Nz + 1) + (long long)g.nxl * (g.Nz + 1) +
                          (long long)g.nxl * g.Ny;
    long long blocks = (per * nfields + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (c->f64)
        fill_ghost_kernel<double><<<(unsigned)blocks, 256, 0, s>>>(
            (double*)mesh, nfields, g.nxl, g.Ny, g.Nz, g.Nzp, fill_x, g.ghost_elems);
    else
        fill_ghost_kernel<float><<<(unsigned)blocks, 256, 0, s>>>(
            (float*)mesh, nfields, g.nxl, g.Ny, g.Nz, g.Nzp, fill_x, g.ghost_elems);
    HYMD_LAUNCH_CHECK(c);
    return HYMD_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_map(hymd_ctx* c, CUtensorMap* map, void* base, int nfields) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        HYMD_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !fn) {
            set_error("cuTensorMapEncodeTiled not available from the driver");
            return HYMD_ERR_CUDA;
        }
        encode = (EncodeTiledFn)fn;
    }
    const Geometry& g = c->g;
    const cuuint64_t sz = c->rsz;
    cuuint64_t dims[4] = {(cuuint64_t)g.Nzp, (cuuint64_t)(g.Ny + 1), (cuuint64_t)(g.nxl + 1),
                          (cuuint64_t)nfields};
    cuuint64_t strides[3] = {g.Nzp * sz, (cuuint64_t)(g.Ny + 1) * g.Nzp * sz,
                             (cuuint64_t)g.ghost_elems * sz};
    cuuint32_t box[4] = {(cuuint32_t)c->rbz, (cuuint32_t)(c->rty + 1), (cuuint32_t)(c->rtx + 1), 3};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(map, c->f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                        4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return HYMD_ERR_CUDA;
    }
    return HYMD_OK;
}

int readout_setup(hymd_ctx* c) {
    // largest tile whose boxes for all U potential rows fit ~100 KB (two CTAs per SM)
    static const int cand[][2] = {{8, 8}, {4, 8}, {4, 4}, {2, 4}, {2, 2}, {1, 2}, {1, 1}};
    const int tz = 32;
    const int bz = ((tz + 1 + 3) / 4) * 4;
    size_t budget = 100 * 1024;
    if (const char* e = getenv("HYMD_B200_READOUT_KB")) budget = (size_t)atoi(e) * 1024;   // tuning
    int pick = 6;
    for (int i = 0; i < 7; ++i) {
        size_t bytes = (size_t)c->U * 3 * (cand[i][0] + 1) * (cand[i][1] + 1) * bz * c->rsz;
        if (bytes <= budget) { pick = i; break; }
    }
    c->rtx = cand[pick][0]; c->rty = cand[pick][1]; c->rtz = tz; c->rbz = bz;
    const size_t box_bytes = (size_t)3 * (c->rtx + 1) * (c->rty + 1) * bz * c->rsz;
    const size_t box_stride = (box_bytes + 127) / 128 * 128;
    c->readout_smem = (size_t)c->U * box_stride + 16 + HYMD_MAX_TYPES * sizeof(int);
    c->readout_smem_pme = box_stride + 16 + HYMD_MAX_TYPES * sizeof(int);
    if (c->readout_smem > 227 * 1024) {
        set_error("readout: %d distinct potential rows need %zu B of shared memory", c->U,
                  c->readout_smem);
        return HYMD_ERR_INVALID;
    }
    HYMD_CHECK(encode_map(c, &c->tmap_gmesh, c->gmesh, 3 * c->U));
    if (c->cfg.pme) HYMD_CHECK(encode_map(c, &c->tmap_emesh, c->emesh, 3));
    return HYMD_OK;
}

template <typename real, bool CHARGE>
static int launch_readout(hymd_ctx* c, void* d_force, cudaStream_t s) {
    using Tr = RTraits<real>;
    const Geometry& g = c->g;
    ReadoutParams p;
    p.Nx = g.Nx; p.Ny = g.Ny; p.Nz = g.Nz; p.nxl = g.nxl;
    p.fbx = g.fbx; p.fby = g.fby; p.fbz = g.fbz;
    p.tx = c->rtx; p.ty = c->rty; p.tz = c->rtz; p.bz = c->rbz;
    p.ntx = (g.nxl + p.tx - 1) / p.tx;
    p.nty = (g.Ny + p.ty - 1) / p.ty;
    p.ntz = (g.Nz + p.tz - 1) / p.tz;
    p.U = CHARGE ? 1 : c->U;
    p.T = c->T;
    p.box_bytes = (unsigned int)((size_t)3 * (p.tx + 1) * (p.ty + 1) * p.bz * sizeof(real));
    p.box_stride = (p.box_bytes + 127u) / 128u * 128u;
    p.debug_seq = getenv("HYMD_B200_DEBUG_SEQ") != nullptr;
    const size_t smem = CHARGE ? c->readout_smem_pme : c->readout_smem;
    auto kern = readout_kernel<real, CHARGE>;
    HYMD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long blocks = (long long)p.ntx * p.nty * p.ntz;
    kern<<<(unsigned int)blocks, 256, smem, s>>>(CHARGE ? c->tmap_emesh : c->tmap_gmesh,
                                                 (const typename Tr::Rec*)c->rec,
                                                 (const real*)c->q_sorted, c->cell_start, c->d_urow,
                                                 (real*)d_force, p);
    HYMD_LAUNCH_CHECK(c);
    return HYMD_OK;
}

int readout_forces(hymd_ctx* c, void* d_force, cudaStream_t s) {
    return c->f64 ? launch_readout<double, false>(c, d_force, s)
                  : launch_readout<float, false>(c, d_force, s);
}

int readout_pme(hymd_ctx* c, void* d_force, cudaStream_t s) {
    return c->f64 ? launch_readout<double, true>(c, d_force, s)
                  : launch_readout<float, true>(c, d_force, s);
}

}  // namespace hymd
