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

// 2^e assembled from the exponent bits (|e| stays far inside the normal range; exp2() in double was ~150
// instructions per thread, a tenth of what a CTA executes)
template <typename real> __device__ __forceinline__ real pow2_real(int e);
template <> __device__ __forceinline__ float pow2_real<float>(int e) { return __int_as_float((e + 127) << 23); }
template <> __device__ __forceinline__ double pow2_real<double>(int e) { return __longlong_as_double((long long)(e + 1023) << 52); }

__device__ __forceinline__ int ceil_log2_u32(unsigned int v) {
    return v <= 1 ? 0 : 32 - __clz(v - 1);
}

// Accumulators of one field: 64 vertex columns (lx, ly) + one TRASH column, each holding the z vertices -1 .. 32 of
// the tile at words 0 .. 33 (the two ends are a halo that is never flushed).  A particle of a feeding cell adds all
// eight contributions unconditionally: the (x,y) vertices that fall outside the tile are redirected to the trash
// column by the per-run column table, the z vertices outside it land in the halo.  No predicates, no branches in
// the deposit (it was 8 branch / reconvergence pairs per particle: ptxas does not if-convert shared atomics).
constexpr int PAINT_ZS = PAINT_TZ + 2;                          // words per column
constexpr int PAINT_COLS = PAINT_TX * PAINT_TY + 1;             // + trash
constexpr int PAINT_TILEP = PAINT_COLS * PAINT_ZS;              // accumulators per field (2210)
constexpr int PAINT_ROWS = (PAINT_TX + 1) * (PAINT_TY + 1);   // (x,y) rows of cells feeding a tile
constexpr int PAINT_RUNS = 256;                                 // 2 z-segments per row, padded
constexpr int PAINT_NRUN = (2 * PAINT_ROWS + 15) / 16 * 16;     // runs that exist
constexpr int PAINT_JMAP = 3584;   // positions of the flattened particle list with a direct run lookup

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
    constexpr int TX = PAINT_TX, TY = PAINT_TY, TZ = PAINT_TZ, ZS = PAINT_ZS;
    constexpr int TILE = PAINT_TILEP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Acc* V = reinterpret_cast<Acc*>(smem_raw);
    __shared__ typename Scan::TempStorage scan_tmp;
    // per run: {first record - position of the run in the flattened list, wrap flag}: record index = j + delta;
    // accumulator offsets (bytes) of the run's four (x,y) vertices (00, 01, 10, 11; outside the tile: trash column)
    __shared__ uint2 s_rinfo[PAINT_NRUN];
    __shared__ uint4 s_cols[PAINT_NRUN];
    __shared__ uint32_t s_off[PAINT_RUNS];
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

    // ---- locate the particle runs feeding this tile: run s = 2*row + segment ----------------
    // row r = (rx+1)*(TY+1) + (ry+1) with rx, ry in -1..T-1 (the -1 skirt holds the cells whose
    // upper vertices fall into the tile); segment 0 = cells z0-1 .. zb-1, segment 1 = the
    // periodic wrap cell Nz-1 seen as lz = -1 (only for z0 == 0).
    {
        const int sidx = threadIdx.x;
        uint32_t pa = 0, len = 0, wrap = 0;
        uint4 cols = make_uint4(0, 0, 0, 0);
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
                wrap = (uint32_t)seg;
                constexpr int AB = (int)sizeof(Acc);
                const int trash = TX * TY * ZS * AB;
                const int c00 = (rx * TY + ry) * ZS * AB;      // vertex column (rx, ry); only used where valid
                cols.x = (uint32_t)((x_lo && y_lo) ? c00 : trash);
                cols.y = (uint32_t)((x_lo && y_hi) ? c00 + ZS * AB : trash);
                cols.z = (uint32_t)((x_hi && y_lo) ? c00 + TY * ZS * AB : trash);
                cols.w = (uint32_t)((x_hi && y_hi) ? c00 + (TY * ZS + ZS) * AB : trash);
            }
        }
        uint32_t off, total;
        Scan(scan_tmp).ExclusiveSum(len, off, total);
        if (sidx < PAINT_NRUN) {
            s_rinfo[sidx] = make_uint2(pa - off, wrap);
            s_cols[sidx] = cols;
        }
        s_off[sidx] = off;
        if (sidx == 0) s_total = total;
        // direct lookup "position in the flattened list -> run" for the first PAINT_JMAP positions
        const uint32_t hi = min(off + len, (uint32_t)PAINT_JMAP);
        for (uint32_t k = off; k < hi; ++k) s_run[k] = (uint8_t)sidx;
    }
    // ---- clear the accumulators (16-byte stores) ---------------------------------------------
    {
        int4* V4 = reinterpret_cast<int4*>(V);
        const int n4 = (nt * TILE * (int)sizeof(Acc) + 15) / 16;
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
    const real scale = pow2_real<real>(e);
    __syncthreads();

    const UT mx = ((UT)1 << p.fbx) - 1, my = ((UT)1 << p.fby) - 1, mz = ((UT)1 << p.fbz) - 1;
    const real ifx = pow2_real<real>(-p.fbx), ify = pow2_real<real>(-p.fby),
               ifz = pow2_real<real>(e - p.fbz);              // the z weights carry the fixed-point scale
    const uint32_t total = s_total;

    // ---- deposit: every thread takes particles from the flattened list ----------------------
    // straight-line code: 8 shared-memory integer reductions per particle
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
        const uint2 ri = s_rinfo[sidx];
        const uint4 co = s_cols[sidx];
        const uint32_t i = j + ri.x;
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
        int lz1 = (int)(rc.uz >> p.fbz) - z0 + 1;              // cell z0-1 .. z0+31 -> word 0 .. 32 of its column
        if (ri.y) lz1 = 0;                                     // the wrap cell Nz-1
        const real wx1 = w * dx, wx0 = w - wx1;
        const real w00 = wx0 - wx0 * dy, w01 = wx0 * dy, w10 = wx1 - wx1 * dy, w11 = wx1 * dy;
        const real wz1 = dzs, wz0 = scale - dzs;
        unsigned char* v = reinterpret_cast<unsigned char*>(V + tl * TILE + lz1);
        Tr::add(reinterpret_cast<Acc*>(v + co.x), Tr::to_fixed(w00, wz0));
        Tr::add(reinterpret_cast<Acc*>(v + co.x) + 1, Tr::to_fixed(w00, wz1));
        Tr::add(reinterpret_cast<Acc*>(v + co.y), Tr::to_fixed(w01, wz0));
        Tr::add(reinterpret_cast<Acc*>(v + co.y) + 1, Tr::to_fixed(w01, wz1));
        Tr::add(reinterpret_cast<Acc*>(v + co.z), Tr::to_fixed(w10, wz0));
        Tr::add(reinterpret_cast<Acc*>(v + co.z) + 1, Tr::to_fixed(w10, wz1));
        Tr::add(reinterpret_cast<Acc*>(v + co.w), Tr::to_fixed(w11, wz0));
        Tr::add(reinterpret_cast<Acc*>(v + co.w) + 1, Tr::to_fixed(w11, wz1));
    }
    __syncthreads();

    // ---- flush: every owned vertex is written once; per-field scale = m_t / (dV * 2^e) -------
    const real inv_scale = pow2_real<real>(-e);
    // thread t owns the four consecutive z vertices 4 (t % 8) of row (lx, ly) = (t / 64 + 4 h, (t / 8) % 8), h = 0, 1,
    // of every field: one pointer per thread, constant strides between its stores
    static_assert(TX == 8 && TY == 8 && TZ == 32, "flush: thread <-> vertex map");
    constexpr int VEC = 4;
    const bool vec_ok = (p.Nz % VEC) == 0;     // rows start 16/32-byte aligned and end on a group
    {
        const int t = threadIdx.x;
        const int lz = (t & 7) * VEC, ly = (t >> 3) & 7, lx0 = t >> 6;
        const int gy = y0 + ly, gz = z0 + lz;
        if (gy < p.Ny && gz < p.Nz) {
            real* o0 = out + (long long)t0 * p.out_field_stride + (long long)(x0 + lx0) * p.out_plane_stride +
                       (long long)gy * p.Nz + gz;
            const Acc* v0 = V + (lx0 * TY + ly) * ZS + lz + 1;
            for (int tl = 0; tl < nt; ++tl) {
                const real sc_t = (CHARGE ? outscale[0] : outscale[t0 + tl]) * inv_scale;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (x0 + lx0 + 4 * h >= p.vx) continue;
                    const Acc* v = v0 + tl * TILE + h * (4 * TY * ZS);
                    real* o = o0 + (long long)tl * p.out_field_stride + (long long)(4 * h) * p.out_plane_stride;
                    real r4[VEC];
#pragma unroll
                    for (int k = 0; k < VEC; ++k) r4[k] = (CHARGE ? Tr::to_real(v[k]) : Tr::to_real_unsigned(v[k])) * sc_t;
                    if (vec_ok) {
                        store_vec4(o, r4);
                    } else {
#pragma unroll
                        for (int k = 0; k < VEC; ++k)
                            if (gz + k < p.Nz) o[k] = r4[k];
                    }
                }
            }
        }
    }
}


// =====================================================================================================
// Row-walker paint (HYMD_B200_PAINT=rows; not the default).  What bounds the flattened kernel above is the
// shared-memory atomic pipe:
// 32 consecutive particles of the flattened list sit in the same (x,y) cell row at random heights, so the 32
// lanes of a deposit hit random banks (ncu: 4 extra wavefronts per ATOMS, l1tex 80 % busy).  Here a LANE owns
// a cell ROW and walks its z-run, and the accumulators are stored z-major, V[z][column], so the bank of a
// deposit is the lane's column -- distinct for the 32 rows a warp owns, whatever the heights:
//   warps 0-2: rows rx 0..3, ry 0..7        warps 3-5: rows rx 4..7, ry 0..7        warps 6-7: the 17 skirt rows
// (the warps of a group interleave over the run positions k = sub, sub + nsub, ...).  The tile's feed -- one
// contiguous run of 16-byte records per row -- is brought into shared memory by bulk asynchronous copies (one
// cp.async.bulk per run, completion counted on an mbarrier) while the accumulators are cleared; runs beyond the
// staging capacity are read from global memory.  Same fixed-point arithmetic and the same sums as the kernel
// above: the two are bitwise identical.
// Measured at C4 on a B200 (profiles/r3b_paint_rows.md): the atomic wavefronts per ATOMS drop from 5.1 to 1.4 and
// the shared-memory pipe from 80 % to 32 % busy -- but the kernel takes 0.223 ms against 0.154 ms: 70 KB of shared
// memory per CTA leave 3 CTAs (24 warps) per SM, a lane's run is a serial chain (wait + short-scoreboard stalls),
// only 68 % of the lanes of a deposit are active (run lengths are Poisson) and the warps of a CTA wait for the
// longest run at the barrier.  Kept as a tested variant and as the record of that experiment.
// =====================================================================================================
constexpr int PAINT_NC = PAINT_TX * PAINT_TY;                   // vertex columns of a tile
constexpr int PAINT_TILEZ = PAINT_ZS * PAINT_NC;                // accumulators per field, z-major (2176)
constexpr int PAINT_STAGE_BYTES = 36 * 1024;                    // staged records: 3 CTAs per SM with 4 fp32 fields

__device__ __forceinline__ uint32_t paint_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void paint_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(paint_smem_addr(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void paint_mbar_arrive_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(paint_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void paint_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PAINT_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra PAINT_WAIT_DONE;\n"
        "bra PAINT_WAIT_LOOP;\n"
        "PAINT_WAIT_DONE:\n"
        "}\n" ::"r"(paint_smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void paint_bulk_load(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(paint_smem_addr(dst_smem)), "l"(src), "r"(bytes), "r"(paint_smem_addr(bar)) : "memory");
}

__device__ __forceinline__ void store_vec8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store_vec8(double* p, const double (&v)[8]) {
#pragma unroll
    for (int k = 0; k < 8; k += 2) *reinterpret_cast<double2*>(p + k) = make_double2(v[k], v[k + 1]);
}

template <typename real, bool CHARGE>
__global__ void __launch_bounds__(256) paint_rows_kernel(
    const typename PaintTraits<real>::Rec* __restrict__ rec, const real* __restrict__ q_sorted,
    const uint32_t* __restrict__ start, const DeviceScalars* __restrict__ sc,
    const real* __restrict__ outscale, real* __restrict__ out, PaintParams p) {
    using Tr = PaintTraits<real>;
    using Acc = typename Tr::Acc;
    using UT = typename Tr::UT;
    using Rec = typename Tr::Rec;
    using Scan = cub::BlockScan<uint32_t, 256>;
    constexpr int TX = PAINT_TX, TY = PAINT_TY, TZ = PAINT_TZ, NC = PAINT_NC;
    constexpr int TILE = PAINT_TILEZ;
    constexpr uint32_t CAP = PAINT_STAGE_BYTES / sizeof(Rec);    // records that fit the staging buffer
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Rec* stage = reinterpret_cast<Rec*>(smem_raw);                // first: 16-byte aligned for the bulk copies
    Acc* V = reinterpret_cast<Acc*>(smem_raw + PAINT_STAGE_BYTES);
    __shared__ typename Scan::TempStorage scan_tmp;
    __shared__ uint32_t s_start[PAINT_NRUN], s_len[PAINT_NRUN], s_soff[PAINT_NRUN];
    __shared__ __align__(8) uint64_t s_bar;

    int b = blockIdx.x;
    const int tz_i = b % p.ntz; b /= p.ntz;
    const int ty_i = b % p.nty; b /= p.nty;
    const int tx_i = b % p.ntx; b /= p.ntx;
    const int t0 = b * p.tchunk;
    const int nt = min(p.tchunk, p.T - t0);
    const int x0 = tx_i * TX, y0 = ty_i * TY, z0 = tz_i * TZ;
    const bool periodic_x = (p.P == 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // which of the four (x,y) vertices of cell row (rx, ry) lie inside the tile (and inside the mesh)
    auto row_flags = [&](int rx, int ry, bool& x_lo, bool& x_hi, bool& y_lo, bool& y_hi) {
        x_lo = rx >= 0 && (x0 + rx) < p.vx; x_hi = rx + 1 < TX && (x0 + rx + 1) < p.vx;
        y_lo = ry >= 0 && (y0 + ry) < p.Ny; y_hi = ry + 1 < TY && (y0 + ry + 1) < p.Ny;
    };

    if (tid == 0) paint_mbar_init(&s_bar, 1);
    // ---- the particle runs feeding this tile: run 2 r + seg of row r = (rx+1) (TY+1) + (ry+1), as above ----
    {
        uint32_t pa = 0, len = 0;
        const int r = tid >> 1, seg = tid & 1;
        if (r < PAINT_ROWS && (seg == 0 || z0 == 0)) {
            const int rx = r / (TY + 1) - 1, ry = r % (TY + 1) - 1;
            int gx = x0 + rx, gy = y0 + ry;
            bool ok = true;
            if (gx < 0) { if (periodic_x) gx += p.nxl; else ok = false; }
            if (gx >= p.nxl) ok = false;
            if (gy < 0) gy += p.Ny;
            if (gy >= p.Ny) ok = false;
            bool x_lo, x_hi, y_lo, y_hi;
            row_flags(rx, ry, x_lo, x_hi, y_lo, y_hi);
            if (!(x_lo || x_hi) || !(y_lo || y_hi)) ok = false;
            if (ok) {
                const long long rowbase = ((long long)gx * p.Ny + gy) * p.nbz;
                int ba, bb;
                if (seg == 0) { ba = tz_i == 0 ? 0 : 2 * tz_i - 1; bb = min(2 * tz_i + 2, p.nbz); }
                else { ba = p.nbz - 1; bb = p.nbz; }
                pa = start[rowbase + ba];
                len = start[rowbase + bb] - pa;
            }
        }
        uint32_t off, total;
        Scan(scan_tmp).ExclusiveSum(len, off, total);
        if (tid < PAINT_NRUN) { s_start[tid] = pa; s_len[tid] = len; s_soff[tid] = off; }
        __syncthreads();                                       // the mbarrier is initialised, the table is complete
        if (tid == 0) paint_mbar_arrive_expect(&s_bar, (uint32_t)(min(total, CAP) * sizeof(Rec)));
        if (len > 0 && off < CAP) {                            // the part of the run that fits the staging buffer
            const uint32_t n_st = min(len, CAP - off);
            paint_bulk_load(stage + off, rec + pa, (uint32_t)(n_st * sizeof(Rec)), &s_bar);
        }
    }
    // ---- clear the accumulators while the copies fly ------------------------------------------
    {
        int4* V4 = reinterpret_cast<int4*>(V);
        const int n4 = (nt * TILE * (int)sizeof(Acc) + 15) / 16;
        for (int i = tid; i < n4; i += blockDim.x) V4[i] = make_int4(0, 0, 0, 0);
    }
    float wmax = 1.0f;
    if (CHARGE) wmax = fmaxf(__uint_as_float(sc->qmax_bits), 1e-30f);
    int wexp = 0;
    if (CHARGE) frexpf(wmax, &wexp);
    int e = (CHARGE ? Tr::ACC_BITS : Tr::ACC_BITS_UNSIGNED) - ceil_log2_u32(8u * max(sc->max_cell_count, 1u)) - wexp;
    e = min(e, Tr::MAX_EXP - wexp);
    const real scale = pow2_real<real>(e);
    const UT mx = ((UT)1 << p.fbx) - 1, my = ((UT)1 << p.fby) - 1, mz = ((UT)1 << p.fbz) - 1;
    const real ifx = pow2_real<real>(-p.fbx), ify = pow2_real<real>(-p.fby), ifz = pow2_real<real>(e - p.fbz);

    // ---- this lane's cell row ---------------------------------------------------------------------
    int rx = 0, ry = 0, sub, nsub;
    bool have = true;
    if (warp < 6) { rx = (warp < 3 ? 0 : 4) + (lane >> 3); ry = lane & 7; sub = warp % 3; nsub = 3; }
    else {
        sub = warp - 6; nsub = 2;
        if (lane <= TY) { rx = -1; ry = lane - 1; }
        else if (lane <= TY + TX) { rx = lane - TY - 1; ry = -1; }
        else have = false;
    }
    bool x_lo, x_hi, y_lo, y_hi;
    row_flags(rx, ry, x_lo, x_hi, y_lo, y_hi);
    const int r = (rx + 1) * (TY + 1) + (ry + 1);
    uint32_t pa0 = 0, len0 = 0, so0 = 0, pa1 = 0, len1 = 0, so1 = 0;
    __syncthreads();                                           // accumulators cleared (and the table long since visible)
    if (have) {
        pa0 = s_start[2 * r]; len0 = s_len[2 * r]; so0 = s_soff[2 * r];
        pa1 = s_start[2 * r + 1]; len1 = s_len[2 * r + 1]; so1 = s_soff[2 * r + 1];
    }
    const uint32_t lent = len0 + len1;
    const bool v00 = x_lo && y_lo, v01 = x_lo && y_hi, v10 = x_hi && y_lo, v11 = x_hi && y_hi;
    Acc* vcol = V + (rx * TY + ry);                            // column of vertex (rx, ry): dereferenced only where valid
    paint_mbar_wait(&s_bar, 0);                                // the staged records have landed

    // ---- deposit: the warp walks its 32 runs in lock step -----------------------------------------
    for (uint32_t k = sub;; k += nsub) {
        const bool act = k < lent;
        if (!__any_sync(0xffffffffu, act)) break;
        if (act) {
            const bool wrapseg = k >= len0;
            const uint32_t kk = wrapseg ? k - len0 : k;
            const uint32_t spos = (wrapseg ? so1 : so0) + kk, gidx = (wrapseg ? pa1 : pa0) + kk;
            const Rec rc = spos < CAP ? stage[spos] : rec[gidx];
            int tl = 0;
            real w = (real)1;
            bool mine = true;
            if (CHARGE) w = q_sorted[gidx];
            else { tl = (int)(rc.meta >> Tr::IDX_BITS) - t0; mine = (unsigned)tl < (unsigned)nt; }
            if (mine) {
                const real dx = (real)(rc.ux & mx) * ifx, dy = (real)(rc.uy & my) * ify;
                const real dzs = (real)(rc.uz & mz) * ifz;     // dz * scale
                const int lz1 = wrapseg ? 0 : (int)(rc.uz >> p.fbz) - z0 + 1;   // cell z0-1 .. z0+31 -> plane 0 .. 32
                const real wx1 = w * dx, wx0 = w - wx1;
                const real w00 = wx0 - wx0 * dy, w01 = wx0 * dy, w10 = wx1 - wx1 * dy, w11 = wx1 * dy;
                const real wz1 = dzs, wz0 = scale - dzs;
                Acc* v = vcol + tl * TILE + lz1 * NC;
                if (v00) { Tr::add(v, Tr::to_fixed(w00, wz0)); Tr::add(v + NC, Tr::to_fixed(w00, wz1)); }
                if (v01) { Tr::add(v + 1, Tr::to_fixed(w01, wz0)); Tr::add(v + 1 + NC, Tr::to_fixed(w01, wz1)); }
                if (v10) { Tr::add(v + TY, Tr::to_fixed(w10, wz0)); Tr::add(v + TY + NC, Tr::to_fixed(w10, wz1)); }
                if (v11) { Tr::add(v + TY + 1, Tr::to_fixed(w11, wz0)); Tr::add(v + TY + 1 + NC, Tr::to_fixed(w11, wz1)); }
            }
        }
    }
    __syncthreads();

    // ---- flush: lane = column (conflict-free reads of the z-major planes), 8 consecutive z per lane = one full
    // 32-byte sector per row; warp w takes column group w & 1 and z octet w >> 1 of every field ----
    const real inv_scale = pow2_real<real>(-e);
    {
        const int col = 32 * (warp & 1) + lane, oct = warp >> 1;
        const int lx = col / TY, ly = col % TY;
        const int gx = x0 + lx, gy = y0 + ly, gz = z0 + 8 * oct;
        if (gx < p.vx && gy < p.Ny && gz < p.Nz) {
            const bool vec_ok = (p.Nz % 8) == 0;
            real* o0 = out + (long long)t0 * p.out_field_stride + (long long)gx * p.out_plane_stride + (long long)gy * p.Nz + gz;
            const Acc* v0 = V + (8 * oct + 1) * NC + col;
            for (int tl = 0; tl < nt; ++tl) {
                const real sc_t = (CHARGE ? outscale[0] : outscale[t0 + tl]) * inv_scale;
                real r8[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const Acc a = v0[tl * TILE + k * NC];
                    r8[k] = (CHARGE ? Tr::to_real(a) : Tr::to_real_unsigned(a)) * sc_t;
                }
                real* o = o0 + (long long)tl * p.out_field_stride;
                if (vec_ok) {
                    store_vec8(o, r8);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (gz + k < p.Nz) o[k] = r8[k];
                }
            }
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
    const char* pv = getenv("HYMD_B200_PAINT");             // "rows": the row-walker kernel (measured slower, see above)
    const bool flat = !(pv && strcmp(pv, "rows") == 0);
    const size_t tile_bytes = (size_t)(flat ? PAINT_TILEP : PAINT_TILEZ) * sizeof(typename Tr::Acc);
    int tchunk = (int)((64 * 1024) / tile_bytes);      // <= 64 KB of accumulators per CTA
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
    const size_t smem = (tile_bytes * tchunk + 15) / 16 * 16 + (flat ? 0 : PAINT_STAGE_BYTES);
    auto kern = flat ? paint_kernel<real, CHARGE> : paint_rows_kernel<real, CHARGE>;
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
