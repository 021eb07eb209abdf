// Deterministic cloud-in-cell density painting: pm.paint(...) / volume_per_cell for all types
// (field.py:574-575) and for the charge density (field.py:363-364).
//
// One CTA owns a tile of TX x TY x TZ mesh vertices (for a chunk of particle types) held in
// shared memory as integer fixed-point accumulators.  It gathers the particles of every cell
// that can touch the tile -- the tile's own cells plus a one-cell skirt on the low side of each
// axis -- from the cell-sorted record array (one contiguous z-run per (x,y) row, one warp per
// row), and adds each particle's weights with shared-memory integer atomics.  Integer addition
// is associative, so the result does not depend on the order in which particles arrive:
// the field is bitwise reproducible run to run, needs no global atomics, and every vertex is
// written exactly once (zeros included), so there is no separate clear pass.
//
// The fixed-point scale is 2^e with e chosen on the device from the maximum cell occupancy
// found by the binning pass, so that no accumulator can overflow:
//   |sum| <= 8 * max_cell_count * wmax * 2^e < 2^(bits-1).
#include <cub/block/block_scan.cuh>

#include "ctx.cuh"

namespace hymd {

struct PaintParams {
    int Nx, Ny, Nz, nxl, P;
    int fbx, fby, fbz;
    int vx;            // number of vertex planes painted: nxl (periodic) or nxl+1 (ghost plane)
    int T;             // number of fields
    int tchunk;        // types per CTA
    int ntx, nty, ntz; // tiles per axis
    long long out_field_stride;   // elements between fields in the output
    long long out_plane_stride;   // elements between x planes
};

template <typename real> struct PaintTraits;
template <> struct PaintTraits<float> {
    using Rec = Rec32; using UT = uint32_t; using Acc = int;
    static constexpr int IDX_BITS = REC32_IDX_BITS;
    static constexpr int ACC_BITS = 30;
    __device__ static __forceinline__ Acc to_fixed(float w, float scale) { return __float2int_rn(w * scale); }
    __device__ static __forceinline__ void add(Acc* p, Acc v) { atomicAdd(p, v); }
    __device__ static __forceinline__ float to_real(Acc v) { return __int2float_rn(v); }
};
template <> struct PaintTraits<double> {
    using Rec = Rec64; using UT = unsigned long long; using Acc = long long;
    static constexpr int IDX_BITS = REC64_IDX_BITS;
    static constexpr int ACC_BITS = 61;
    __device__ static __forceinline__ Acc to_fixed(double w, double scale) { return __double2ll_rn(w * scale); }
    __device__ static __forceinline__ void add(Acc* p, Acc v) {
        atomicAdd((unsigned long long*)p, (unsigned long long)v);
    }
    __device__ static __forceinline__ double to_real(Acc v) { return __ll2double_rn(v); }
};

__device__ __forceinline__ void store_vec4(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store_vec4(double* p, const double (&v)[4]) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
}

__device__ __forceinline__ int ceil_log2_u32(unsigned int v) {
    return v <= 1 ? 0 : 32 - __clz(v - 1);
}

constexpr int PAINT_ROWS = (PAINT_TX + 1) * (PAINT_TY + 1);   // (x,y) rows of cells feeding a tile
constexpr int PAINT_RUNS = 256;                                 // 2 z-segments per row, padded

// CHARGE = false: field index = particle type, unit weight (per-type mass applied on output).
// CHARGE = true : single field, weight = sorted charge.
template <typename real, bool CHARGE>
__global__ void __launch_bounds__(256) paint_kernel(
    const typename PaintTraits<real>::Rec* __restrict__ rec, const real* __restrict__ q_sorted,
    const uint32_t* __restrict__ start, const DeviceScalars* __restrict__ sc,
    const real* __restrict__ outscale, real* __restrict__ out, PaintParams p) {
    using Tr = PaintTraits<real>;
    using Acc = typename Tr::Acc;
    using UT = typename Tr::UT;
    using Scan = cub::BlockScan<uint32_t, 256>;
    constexpr int TX = PAINT_TX, TY = PAINT_TY, TZ = PAINT_TZ;
    constexpr int TILE = TX * TY * TZ;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Acc* V = reinterpret_cast<Acc*>(smem_raw);
    __shared__ typename Scan::TempStorage scan_tmp;
    __shared__ uint32_t s_begin[PAINT_RUNS], s_off[PAINT_RUNS];
    __shared__ uint32_t s_total;

    // tile / type-chunk of this CTA (z tiles fastest so that neighbouring CTAs share rows in L2)
    int b = blockIdx.x;
    const int tz_i = b % p.ntz; b /= p.ntz;
    const int ty_i = b % p.nty; b /= p.nty;
    const int tx_i = b % p.ntx; b /= p.ntx;
    const int t0 = b * p.tchunk;
    const int nt = min(p.tchunk, p.T - t0);
    const int x0 = tx_i * TX, y0 = ty_i * TY, z0 = tz_i * TZ;
    const bool periodic_x = (p.P == 1);
    const int zb = min(z0 + TZ, p.Nz);

    // ---- locate the particle runs feeding this tile: run s = 2*row + segment ----------------
    // row r = (rx+1)*(TY+1) + (ry+1) with rx, ry in -1..T-1 (the -1 skirt holds the cells whose
    // upper vertices fall into the tile); segment 0 = cells z0-1 .. zb-1, segment 1 = the
    // periodic wrap cell Nz-1 seen as lz = -1 (only for z0 == 0).
    {
        const int sidx = threadIdx.x;
        uint32_t pa = 0, len = 0;
        const int r = sidx >> 1, seg = sidx & 1;
        if (r < PAINT_ROWS && (seg == 0 || z0 == 0)) {
            const int rx = r / (TY + 1) - 1, ry = r % (TY + 1) - 1;
            int gx = x0 + rx, gy = y0 + ry;
            bool ok = true;
            if (gx < 0) { if (periodic_x) gx += p.nxl; else ok = false; }
            if (gx >= p.nxl) ok = false;
            if (gy < 0) gy += p.Ny;
            if (gy >= p.Ny) ok = false;
            const bool x_lo = rx >= 0 && (x0 + rx) < p.vx, x_hi = rx + 1 < TX && (x0 + rx + 1) < p.vx;
            const bool y_lo = ry >= 0 && (y0 + ry) < p.Ny, y_hi = ry + 1 < TY && (y0 + ry + 1) < p.Ny;
            if (!(x_lo || x_hi) || !(y_lo || y_hi)) ok = false;
            if (ok) {
                const long long rowbase = ((long long)gx * p.Ny + gy) * p.Nz;
                int ca, cb;
                if (seg == 0) { ca = z0 == 0 ? 0 : z0 - 1; cb = zb; }
                else { ca = p.Nz - 1; cb = p.Nz; }
                pa = start[rowbase + ca];
                len = start[rowbase + cb] - pa;
            }
        }
        uint32_t off, total;
        Scan(scan_tmp).ExclusiveSum(len, off, total);
        s_begin[sidx] = pa;
        s_off[sidx] = off;
        if (sidx == 0) s_total = total;
    }
    // ---- clear the accumulators (16-byte stores) ---------------------------------------------
    {
        int4* V4 = reinterpret_cast<int4*>(V);
        const int n4 = nt * TILE * (int)sizeof(Acc) / 16;
        for (int i = threadIdx.x; i < n4; i += blockDim.x) V4[i] = make_int4(0, 0, 0, 0);
    }

    // fixed-point scale from the occupancy bound (identical in every CTA)
    float wmax = 1.0f;
    if (CHARGE) wmax = fmaxf(__uint_as_float(sc->qmax_bits), 1e-30f);
    int wexp;
    frexpf(wmax, &wexp);                       // wmax <= 2^wexp
    const int e = Tr::ACC_BITS - ceil_log2_u32(8u * max(sc->max_cell_count, 1u)) - wexp;
    const real scale = (real)exp2((double)e);
    __syncthreads();

    const UT mx = ((UT)1 << p.fbx) - 1, my = ((UT)1 << p.fby) - 1, mz = ((UT)1 << p.fbz) - 1;
    const real ifx = (real)1 / (real)((UT)1 << p.fbx), ify = (real)1 / (real)((UT)1 << p.fby),
               ifz = (real)1 / (real)((UT)1 << p.fbz);
    const uint32_t total = s_total;

    // ---- deposit: every thread takes particles from the flattened list ----------------------
    for (uint32_t j = threadIdx.x; j < total; j += blockDim.x) {
        int sidx = 0;      // largest run with s_off[run] <= j (empty runs share offsets: the last wins)
#pragma unroll
        for (int step = PAINT_RUNS / 2; step > 0; step >>= 1)
            if (s_off[sidx + step] <= j) sidx += step;
        const uint32_t i = s_begin[sidx] + (j - s_off[sidx]);
        const typename Tr::Rec rc = rec[i];
        const int r = sidx >> 1, seg = sidx & 1;
        const int rx = r / (TY + 1) - 1, ry = r % (TY + 1) - 1;
        int tl = 0;
        real w = (real)1;
        if (CHARGE) {
            w = q_sorted[i];
        } else {
            tl = (int)(rc.meta >> Tr::IDX_BITS) - t0;
            if (tl < 0 || tl >= nt) continue;
        }
        const bool x_lo = rx >= 0 && (x0 + rx) < p.vx, x_hi = rx + 1 < TX && (x0 + rx + 1) < p.vx;
        const bool y_lo = ry >= 0 && (y0 + ry) < p.Ny, y_hi = ry + 1 < TY && (y0 + ry + 1) < p.Ny;
        const real dx = (real)(rc.ux & mx) * ifx, dy = (real)(rc.uy & my) * ify,
                   dz = (real)(rc.uz & mz) * ifz;
        int lz = (int)(rc.uz >> p.fbz) - z0;   // -1 .. TZ-1
        if (seg == 1) lz = -1;                 // the wrap cell Nz-1
        const real wz0 = (real)1 - dz, wz1 = dz;
        const bool z_lo = lz >= 0, z_hi = (lz + 1 < TZ) && (z0 + lz + 1 < p.Nz);
        Acc* Vt = V + tl * TILE;
#pragma unroll
        for (int ax = 0; ax < 2; ++ax) {
            if (!(ax ? x_hi : x_lo)) continue;
            const real wx = w * (ax ? dx : (real)1 - dx);
#pragma unroll
            for (int ay = 0; ay < 2; ++ay) {
                if (!(ay ? y_hi : y_lo)) continue;
                const real wxy = wx * (ay ? dy : (real)1 - dy);
                Acc* row = Vt + ((rx + ax) * TY + (ry + ay)) * TZ;
                if (z_lo) Tr::add(row + lz, Tr::to_fixed(wxy * wz0, scale));
                if (z_hi) Tr::add(row + lz + 1, Tr::to_fixed(wxy * wz1, scale));
            }
        }
    }
    __syncthreads();

    // ---- flush: every owned vertex is written once; per-field scale = m_t / (dV * 2^e) -------
    const real inv_scale = (real)exp2((double)-e);
    constexpr int VEC = 4;
    const bool vec_ok = (p.Nz % VEC) == 0;     // rows start 16/32-byte aligned and end on a group
    for (int i4 = threadIdx.x; i4 < nt * TILE / VEC; i4 += blockDim.x) {
        const int i = i4 * VEC;
        const int tl = i / TILE, v = i % TILE;
        const int lz = v % TZ, ly = (v / TZ) % TY, lx = v / (TZ * TY);
        const int gx = x0 + lx, gy = y0 + ly, gz = z0 + lz;
        if (gx >= p.vx || gy >= p.Ny || gz >= p.Nz) continue;
        const real sc_t = (CHARGE ? outscale[0] : outscale[t0 + tl]) * inv_scale;
        real* o = out + (long long)(t0 + tl) * p.out_field_stride + (long long)gx * p.out_plane_stride +
                  (long long)gy * p.Nz + gz;
        real r4[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) r4[k] = Tr::to_real(V[i + k]) * sc_t;
        if (vec_ok) {
            store_vec4(o, r4);
        } else {
#pragma unroll
            for (int k = 0; k < VEC; ++k)
                if (gz + k < p.Nz) o[k] = r4[k];
        }
    }
}

template <typename real, bool CHARGE>
static int launch_paint(hymd_ctx* c, int nfields, void* out, const void* outscale, cudaStream_t s) {
    using Tr = PaintTraits<real>;
    const Geometry& g = c->g;
    PaintParams p;
    p.Nx = g.Nx; p.Ny = g.Ny; p.Nz = g.Nz; p.nxl = g.nxl; p.P = g.P;
    p.fbx = g.fbx; p.fby = g.fby; p.fbz = g.fbz;
    p.vx = g.P == 1 ? g.nxl : g.nxl + 1;
    p.T = nfields;
    const size_t tile_bytes = (size_t)PAINT_TX * PAINT_TY * PAINT_TZ * sizeof(typename Tr::Acc);
    int tchunk = (int)((64 * 1024) / tile_bytes);      // <= 64 KB per CTA: 3 CTAs per SM
    if (tchunk < 1) tchunk = 1;
    if (tchunk > nfields) tchunk = nfields;
    p.tchunk = tchunk;
    p.ntx = (p.vx + PAINT_TX - 1) / PAINT_TX;
    p.nty = (g.Ny + PAINT_TY - 1) / PAINT_TY;
    p.ntz = (g.Nz + PAINT_TZ - 1) / PAINT_TZ;
    // multi-GPU: the ghost plane lives in a separate buffer right after the nxl owned planes of
    // each field, so the per-field stride is (nxl+1) planes there; single GPU: nxl planes.
    p.out_plane_stride = (long long)g.Ny * g.Nz;
    p.out_field_stride = (long long)p.vx * p.out_plane_stride;
    const int nchunks = (nfields + tchunk - 1) / tchunk;
    const long long blocks = (long long)p.ntx * p.nty * p.ntz * nchunks;
    const size_t smem = tile_bytes * tchunk;
    auto kern = paint_kernel<real, CHARGE>;
    HYMD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned int)blocks, 256, smem, s>>>(
        (const typename Tr::Rec*)c->rec, (const real*)c->q_sorted, c->cell_start, c->scalars,
        (const real*)outscale, (real*)out, p);
    HYMD_LAUNCH_CHECK(c);
    return HYMD_OK;
}

int paint_types(hymd_ctx* c, cudaStream_t s) {
    return c->f64 ? launch_paint<double, false>(c, c->T, c->phi, c->outscale, s)
                  : launch_paint<float, false>(c, c->T, c->phi, c->outscale, s);
}

int paint_charges(hymd_ctx* c, cudaStream_t s) {
    // outscale[T] holds 1/dV for the charge density
    const void* os = (const char*)c->outscale + (size_t)c->T * c->rsz;
    return c->f64 ? launch_paint<double, true>(c, 1, c->phi_q, os, s)
                  : launch_paint<float, true>(c, 1, c->phi_q, os, s);
}

}  // namespace hymd
