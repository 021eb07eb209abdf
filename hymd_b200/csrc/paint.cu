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
//   |sum| <= 8 * max_cell_count * wmax * 2^e < 2^bits   (bits: 32 for the unsigned density sums,
//   30 for signed charge sums, 61 in the fp64 build; the occupancy bound is the largest sort bin).
#include <cub/block/block_scan.cuh>

#include "ctx.cuh"

namespace hymd {

struct PaintParams {
    int Nx, Ny, Nz, nxl, P, nbz;
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
    // usable magnitude bits: signed sums (charges) keep one bit of head room below the sign, sums of
    // non-negative weights (densities) use the full unsigned word
    static constexpr int ACC_BITS = 30, ACC_BITS_UNSIGNED = 32;
    // round-to-nearest of a * b (|a b| < 2^22) without a conversion instruction: the integer is
    // the mantissa of a b + 1.5 * 2^23
    static constexpr int MAX_EXP = 22;
    __device__ static __forceinline__ Acc to_fixed(float a, float b) {
        return __float_as_int(fmaf(a, b, 12582912.0f)) - 0x4B400000;
    }
    __device__ static __forceinline__ void add(Acc* p, Acc v) { atomicAdd(p, v); }
    __device__ static __forceinline__ float to_real(Acc v) { return __int2float_rn(v); }
    __device__ static __forceinline__ float to_real_unsigned(Acc v) { return __uint2float_rn((unsigned int)v); }
};
template <> struct PaintTraits<double> {
    using Rec = Rec64; using UT = unsigned long long; using Acc = long long;
    static constexpr int IDX_BITS = REC64_IDX_BITS;
    static constexpr int ACC_BITS = 61, ACC_BITS_UNSIGNED = 61;
    static constexpr int MAX_EXP = 1000;
    __device__ static __forceinline__ Acc to_fixed(double a, double b) { return __double2ll_rn(a * b); }
    __device__ static __forceinline__ void add(Acc* p, Acc v) {
        atomicAdd((unsigned long long*)p, (unsigned long long)v);
    }
    __device__ static __forceinline__ double to_real(Acc v) { return __ll2double_rn(v); }
    __device__ static __forceinline__ double to_real_unsigned(Acc v) { return __ll2double_rn(v); }
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
constexpr int PAINT_JMAP = 4096;   // positions of the flattened particle list with a direct run lookup

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
    __shared__ uint32_t s_begin[PAINT_RUNS], s_off[PAINT_RUNS], s_info[PAINT_RUNS];
    __shared__ uint8_t s_run[PAINT_JMAP];
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
    // s_info[run] = everything the deposit needs to know about the run's (x,y) cell column:
    // bits 0..11 accumulator offset of vertex (rx, ry, 0) + 512, bits 12..15 which of the four
    // (x,y) vertices lie inside the tile, bit 16 wrap segment.
    {
        const int sidx = threadIdx.x;
        uint32_t pa = 0, len = 0, info = 0;
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
                // bins of this row (ctx.cuh, ZBIN): the skirt cell z0-1 is bin 2 tz - 1, cells z0 .. z0+30
                // bin 2 tz, cell z0+31 (or Nz-1) bin 2 tz + 1; the wrap cell Nz-1 is the row's last bin
                const long long rowbase = ((long long)gx * p.Ny + gy) * p.nbz;
                int ba, bb;
                if (seg == 0) { ba = tz_i == 0 ? 0 : 2 * tz_i - 1; bb = min(2 * tz_i + 2, p.nbz); }
                else { ba = p.nbz - 1; bb = p.nbz; }
                pa = start[rowbase + ba];
                len = start[rowbase + bb] - pa;
                info = (uint32_t)((rx * TY + ry) * TZ + 512) | (x_lo ? 1u << 12 : 0u) | (x_hi ? 1u << 13 : 0u) |
                       (y_lo ? 1u << 14 : 0u) | (y_hi ? 1u << 15 : 0u) | (seg ? 1u << 16 : 0u);
            }
        }
        uint32_t off, total;
        Scan(scan_tmp).ExclusiveSum(len, off, total);
        s_begin[sidx] = pa;
        s_off[sidx] = off;
        s_info[sidx] = info;
        if (sidx == 0) s_total = total;
        // direct lookup "position in the flattened list -> run" for the first PAINT_JMAP positions
        const uint32_t hi = min(off + len, (uint32_t)PAINT_JMAP);
        for (uint32_t k = off; k < hi; ++k) s_run[k] = (uint8_t)sidx;
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
    int wexp = 0;                              // wmax <= 2^wexp (unit weights: exactly 2^0)
    if (CHARGE) frexpf(wmax, &wexp);
    // sc->max_cell_count is the largest sort bin (>= the largest cell, ctx.cuh ZBIN)
    int e = (CHARGE ? Tr::ACC_BITS : Tr::ACC_BITS_UNSIGNED) - ceil_log2_u32(8u * max(sc->max_cell_count, 1u)) - wexp;
    e = min(e, Tr::MAX_EXP - wexp);            // a single contribution stays inside to_fixed's range
    const real scale = (real)exp2((double)e);
    __syncthreads();

    const UT mx = ((UT)1 << p.fbx) - 1, my = ((UT)1 << p.fby) - 1, mz = ((UT)1 << p.fbz) - 1;
    const real ifx = (real)1 / (real)((UT)1 << p.fbx), ify = (real)1 / (real)((UT)1 << p.fby),
               ifz = scale / (real)((UT)1 << p.fbz);          // the z weights carry the fixed-point scale
    const uint32_t total = s_total;

    // ---- deposit: every thread takes particles from the flattened list ----------------------
    // straight-line code: 8 predicated shared-memory integer atomics per particle
    for (uint32_t j = threadIdx.x; j < total; j += blockDim.x) {
        int sidx;
        if (j < (uint32_t)PAINT_JMAP) {
            sidx = s_run[j];
        } else {       // largest run with s_off[run] <= j (empty runs share offsets: the last wins)
            sidx = 0;
#pragma unroll
            for (int step = PAINT_RUNS / 2; step > 0; step >>= 1)
                if (s_off[sidx + step] <= j) sidx += step;
        }
        const uint32_t i = s_begin[sidx] + (j - s_off[sidx]);
        const uint32_t info = s_info[sidx];
        const typename Tr::Rec rc = rec[i];
        int tl = 0;
        real w = (real)1;
        if (CHARGE) {
            w = q_sorted[i];
        } else {
            tl = (int)(rc.meta >> Tr::IDX_BITS) - t0;
            if ((unsigned)tl >= (unsigned)nt) continue;
        }
        const real dx = (real)(rc.ux & mx) * ifx, dy = (real)(rc.uy & my) * ify;
        const real dzs = (real)(rc.uz & mz) * ifz;             // dz * scale
        int lz = (int)(rc.uz >> p.fbz) - z0;                   // -1 .. TZ-1
        if (info & (1u << 16)) lz = -1;                        // the wrap cell Nz-1
        const bool z_lo = lz >= 0, z_hi = (lz + 1 < TZ) && (z0 + lz + 1 < p.Nz);
        const bool x_lo = info & (1u << 12), x_hi = info & (1u << 13);
        const bool y_lo = info & (1u << 14), y_hi = info & (1u << 15);
        const real wx1 = w * dx, wx0 = w - wx1;
        const real w00 = wx0 - wx0 * dy, w01 = wx0 * dy, w10 = wx1 - wx1 * dy, w11 = wx1 * dy;
        const real wz1 = dzs, wz0 = scale - dzs;
        Acc* v = V + tl * TILE + ((int)(info & 0xfffu) - 512) + lz;       // vertex (rx, ry, lz)
        if (x_lo && y_lo && z_lo) Tr::add(v, Tr::to_fixed(w00, wz0));
        if (x_lo && y_lo && z_hi) Tr::add(v + 1, Tr::to_fixed(w00, wz1));
        if (x_lo && y_hi && z_lo) Tr::add(v + TZ, Tr::to_fixed(w01, wz0));
        if (x_lo && y_hi && z_hi) Tr::add(v + TZ + 1, Tr::to_fixed(w01, wz1));
        if (x_hi && y_lo && z_lo) Tr::add(v + TY * TZ, Tr::to_fixed(w10, wz0));
        if (x_hi && y_lo && z_hi) Tr::add(v + TY * TZ + 1, Tr::to_fixed(w10, wz1));
        if (x_hi && y_hi && z_lo) Tr::add(v + TY * TZ + TZ, Tr::to_fixed(w11, wz0));
        if (x_hi && y_hi && z_hi) Tr::add(v + TY * TZ + TZ + 1, Tr::to_fixed(w11, wz1));
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
        for (int k = 0; k < VEC; ++k) r4[k] = (CHARGE ? Tr::to_real(V[i + k]) : Tr::to_real_unsigned(V[i + k])) * sc_t;
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
    p.Nx = g.Nx; p.Ny = g.Ny; p.Nz = g.Nz; p.nxl = g.nxl; p.P = g.P; p.nbz = g.nbz;
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
