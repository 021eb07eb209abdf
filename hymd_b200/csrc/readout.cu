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

#include <type_traits>

#include "ctx.cuh"

namespace hymd {

struct ReadoutParams {
    int Nx, Ny, Nz, nxl, nbz;
    int fbx, fby, fbz;
    int tx, ty, tz, bz;      // tile of cells and box z extent (tz + 1 rounded up for TMA)
    int ntx, nty, ntz;
    int U, T;
    unsigned int box_bytes;  // bytes of one TMA box (3 components)
    unsigned int box_stride; // box_bytes rounded up to 128 B (TMA destination alignment)
    int nstage;              // box buffers in the TMA ring (tiles in flight + the one in use)
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

constexpr int READOUT_NT = 512, READOUT_NW = READOUT_NT / 32;
constexpr int READOUT_TZ = 32, READOUT_BZ = 36;   // tile z extent and box z extent (tz + 1 -> x4)
static_assert(READOUT_TZ == ZBIN, "readout tiles and sort bins share their z extent");
constexpr int READOUT_MAX_STAGES = 8;

// Persistent CTAs (one per SM) walk the tiles of TX x TY x 32 cells in a software pipeline:
//   iteration i :  publish the particle runs of tile i+3, load those of tile i+4   (threads < rows)
//                  issue the TMA boxes of tile i+NS-1        (one thread; ring of NS box buffers)
//                  prefetch the records of tile i+2          (registers, three rotating sets)
//                  wait for the boxes of tile i, interpolate the records prefetched in i-2
// so the HBM latency of all three input streams is hidden behind the previous tiles' work.
// One (x,y) row of a tile is one contiguous run of the cell-sorted records; warp w owns the rows
// w, w+16, ... and lane l the l-th particle of the run (longer runs: an on-demand tail loop), so
// no search is needed to map particles to rows.
template <typename real, bool CHARGE, int TX, int TY>
__global__ void __launch_bounds__(READOUT_NT, 1) readout_kernel(
    const __grid_constant__ CUtensorMap tmap, const typename RTraits<real>::Rec* __restrict__ rec,
    const real* __restrict__ q_sorted, const uint32_t* __restrict__ start,
    const int* __restrict__ urow, real* __restrict__ force, ReadoutParams p) {
    using Tr = RTraits<real>;
    using UT = typename Tr::UT;
    using Rec = typename Tr::Rec;
    constexpr int NT = READOUT_NT, NW = READOUT_NW, TZ = READOUT_TZ, BZ = READOUT_BZ;
    constexpr int ROWS = TX * TY, RPW = (ROWS + NW - 1) / NW;
    constexpr int PY = TY + 1, PX = TX + 1;
    constexpr int COMP = PX * PY * BZ;                      // elements between components
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const size_t buf_bytes = (size_t)p.U * p.box_stride;
    const int NS = p.nstage;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + NS * buf_bytes);
    int* s_urow = reinterpret_cast<int*>(bar + READOUT_MAX_STAGES);
    __shared__ uint32_t s_begin[4][ROWS], s_len[4][ROWS];

    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const int ntiles = p.ntx * p.nty * p.ntz;
    const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    auto tile_origin = [&](int it, int& x0, int& y0, int& z0) {
        int b = (int)blockIdx.x + it * (int)gridDim.x;
        z0 = (b % p.ntz) * TZ; b /= p.ntz;
        y0 = (b % p.nty) * TY; b /= p.nty;
        x0 = b * TX;
    };
    // Threads 0..ROWS-1 load the run bounds of a tile into registers one iteration before they
    // publish them, so the start[] latency never sits on the critical path.
    uint32_t run_pa = 0, run_len = 0;
    auto load_runs = [&](int it) {
        int x0, y0, z0;
        tile_origin(it, x0, y0, z0);
        run_pa = 0; run_len = 0;
        const int gx = x0 + tid / TY, gy = y0 + tid % TY;
        if (gx < p.nxl && gy < p.Ny) {
            // cells z0 .. z0+31 of the row = sort bins 2 tz and 2 tz + 1 (ctx.cuh, ZBIN)
            const long long rowbase = ((long long)gx * p.Ny + gy) * p.nbz;
            const int tzi = z0 / TZ;
            run_pa = start[rowbase + 2 * tzi];
            run_len = start[rowbase + min(2 * tzi + 2, p.nbz)] - run_pa;
        }
    };
    auto publish_runs = [&](int it) {
        s_begin[it & 3][tid] = run_pa;
        s_len[it & 3][tid] = run_len;
    };
    auto issue_boxes = [&](int it) {
        int x0, y0, z0;
        tile_origin(it, x0, y0, z0);
        uint64_t* b = bar + (it % NS);
        mbar_expect_tx(b, (uint32_t)p.U * p.box_bytes);
        for (int u = 0; u < p.U; ++u)
            tma_load_4d(smem_raw + (it % NS) * buf_bytes + (size_t)u * p.box_stride, &tmap, b, z0, y0, x0, 3 * u);
    };

    if (tid == 0) {
        for (int i = 0; i < NS; ++i) mbar_init(bar + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (!CHARGE)
        for (int i = tid; i < p.T; i += NT) s_urow[i] = urow[i];
    if (my_tiles == 0) return;
    if (tid < ROWS) {
        load_runs(0); publish_runs(0);
        if (my_tiles > 1) { load_runs(1); publish_runs(1); }
        if (my_tiles > 2) { load_runs(2); publish_runs(2); }
        if (my_tiles > 3) load_runs(3);
    }
    __syncthreads();
    if (tid == 0)
        for (int i = 0; i < NS - 1 && i < my_tiles; ++i) issue_boxes(i);

    const UT mx = ((UT)1 << p.fbx) - 1, my = ((UT)1 << p.fby) - 1, mz = ((UT)1 << p.fbz) - 1;
    const real ifx = (real)1 / (real)((UT)1 << p.fbx), ify = (real)1 / (real)((UT)1 << p.fby),
               ifz = (real)1 / (real)((UT)1 << p.fbz);
    const UT idx_mask = ((UT)1 << Tr::IDX_BITS) - 1;
    const int box_elems = (int)(p.box_stride / sizeof(real));

    // Three register sets of prefetched records rotate through the loop (unrolled by three, so the
    // rotation is a renaming and no loaded register is touched before its tile is processed).
    Rec recs[3][RPW];
    real recq[3][RPW];
    bool have[3][RPW];
    auto prefetch = [&](auto setc, int it) {
        constexpr int SET = decltype(setc)::value;
#pragma unroll
        for (int k = 0; k < RPW; ++k) {
            const int row = warp + k * NW;
            have[SET][k] = false;
            recq[SET][k] = (real)1;
            if (row < ROWS && it < my_tiles && (uint32_t)lane < s_len[it & 3][row]) {
                const uint32_t i = s_begin[it & 3][row] + lane;
                have[SET][k] = true;
                recs[SET][k] = rec[i];
                if (CHARGE) recq[SET][k] = q_sorted[i];
            }
        }
    };
    prefetch(std::integral_constant<int, 0>{}, 0);
    prefetch(std::integral_constant<int, 1>{}, 1);

    auto body = [&](auto setc, int it) {
        constexpr int SET = decltype(setc)::value;
        if (it >= my_tiles) return;
        if (it > 0) __syncthreads();     // everyone is done with tile it-1: its buffers may be reused
        if (tid < ROWS) {
            if (it + 3 < my_tiles) publish_runs(it + 3);     // loaded during the previous iteration
            if (it + 4 < my_tiles) load_runs(it + 4);
        }
        if (tid == 64 && it + NS - 1 < my_tiles) issue_boxes(it + NS - 1);   // into the buffer tile it-1 used
        prefetch(std::integral_constant<int, (SET + 2) % 3>{}, it + 2);
        const int z0 = (((int)blockIdx.x + it * (int)gridDim.x) % p.ntz) * TZ;
        const real* S = reinterpret_cast<const real*>(smem_raw + (it % NS) * buf_bytes);
        mbar_wait(bar + (it % NS), (uint32_t)((it / NS) & 1));

        auto interpolate = [&](const Rec& rc, int rowoff, real q) {
            const real dx = (real)(rc.ux & mx) * ifx, dy = (real)(rc.uy & my) * ify,
                       dz = (real)(rc.uz & mz) * ifz;
            const int lz = (int)(rc.uz >> p.fbz) - z0;
            int u = 0;
            if (!CHARGE) u = s_urow[(int)(rc.meta >> Tr::IDX_BITS)];
            const real* B = S + u * box_elems + rowoff + lz;
            real f0 = 0, f1 = 0, f2 = 0;
#pragma unroll
            for (int ax = 0; ax < 2; ++ax) {
                const real wx = ax ? dx : (real)1 - dx;
#pragma unroll
                for (int ay = 0; ay < 2; ++ay) {
                    const real wxy = wx * (ay ? dy : (real)1 - dy);
                    const real* c = B + (ax * PY + ay) * BZ;
                    const real w0 = wxy * ((real)1 - dz), w1 = wxy * dz;
                    f0 += w0 * c[0] + w1 * c[1];
                    f1 += w0 * c[COMP] + w1 * c[COMP + 1];
                    f2 += w0 * c[2 * COMP] + w1 * c[2 * COMP + 1];
                }
            }
            if (CHARGE) { f0 *= q; f1 *= q; f2 *= q; }
            real* o = force + (size_t)(rc.meta & idx_mask) * 3;
            o[0] = f0; o[1] = f1; o[2] = f2;
        };
#pragma unroll
        for (int k = 0; k < RPW; ++k) {
            const int row = warp + k * NW;
            if (row >= ROWS) continue;
            const int rowoff = ((row / TY) * PY + row % TY) * BZ;
            if (have[SET][k]) interpolate(recs[SET][k], rowoff, recq[SET][k]);
            // crowded rows: the records past the first 32 are fetched on demand
            const uint32_t len = s_len[it & 3][row];
            if (len > 32) {
                const uint32_t b0 = s_begin[it & 3][row];
                for (uint32_t j = lane + 32; j < len; j += 32)
                    interpolate(rec[b0 + j], rowoff, CHARGE ? q_sorted[b0 + j] : (real)1);
            }
        }
    };
    for (int it = 0; it < my_tiles; it += 3) {
        body(std::integral_constant<int, 0>{}, it);
        body(std::integral_constant<int, 1>{}, it + 1);
        body(std::integral_constant<int, 2>{}, it + 2);
    }
}

// Alternative without staging: one thread per cell-sorted particle gathers its 8 x 3 mesh values
// straight from the ghost-padded meshes through L1 (neighbouring particles share cache lines,
// the ghost planes remove all wrap logic).  Full occupancy hides the latency.
struct GatherParams {
    long long n, ghost_elems;
    int Ny1, Nzp;            // Ny + 1, padded z pitch
    int fbx, fby, fbz;
    // several slabs (sort.cu, per-step routing): n is a capacity bound, the live count is rt->n_work;
    // records with idx >= n_home are guests: guest e = idx - n_home came from rank e / G as its row
    // e % G, and its force goes into that rank's return section for this rank
    const RouteTotals* rt;
    long long n_home, G;
    int rank;
    void* ret[HYMD_MAX_PEERS];
};

// 2^e from the exponent bits (the CIC fraction is fraction_bits * 2^-fb exactly, as the division was)
template <typename real> __device__ __forceinline__ real gather_pow2(int e);
template <> __device__ __forceinline__ float gather_pow2<float>(int e) { return __int_as_float((e + 127) << 23); }
template <> __device__ __forceinline__ double gather_pow2<double>(int e) { return __longlong_as_double((long long)(e + 1023) << 52); }

template <typename real, bool CHARGE>
__global__ void __launch_bounds__(256) readout_gather_kernel(
    const real* __restrict__ mesh, const typename RTraits<real>::Rec* __restrict__ rec,
    const real* __restrict__ q_sorted, const int* __restrict__ urow, real* __restrict__ force,
    GatherParams p) {
    using Tr = RTraits<real>;
    using UT = typename Tr::UT;
    // Grid-stride walk over the sorted records; a thread loads the record of its NEXT particle before it gathers for
    // the current one (only matters with the persistent grid, launch_gather: HYMD_B200_READOUT_PERSIST=1).  Several
    // slabs: the live count is rt->n_work (home particles + guests), known only on the device.
    const long long limit = p.rt ? (long long)p.rt->n_work : p.n;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= limit) return;
    const UT mx = ((UT)1 << p.fbx) - 1, my = ((UT)1 << p.fby) - 1, mz = ((UT)1 << p.fbz) - 1;
    const real ifx = gather_pow2<real>(-p.fbx), ify = gather_pow2<real>(-p.fby), ifz = gather_pow2<real>(-p.fbz);
    typename Tr::Rec rc = rec[i];
    for (;;) {
    const long long inext = i + stride;
    const bool more = inext < limit;
    typename Tr::Rec rn = rc;
    if (more) rn = rec[inext];
    const real dx = (real)(rc.ux & mx) * ifx, dy = (real)(rc.uy & my) * ify, dz = (real)(rc.uz & mz) * ifz;
    const long long cx = (long long)(rc.ux >> p.fbx), cy = (long long)(rc.uy >> p.fby), cz = (long long)(rc.uz >> p.fbz);
    int u = 0;
    if (!CHARGE) u = urow[(int)(rc.meta >> Tr::IDX_BITS)];
    const real* B = mesh + (long long)(3 * u) * p.ghost_elems + (cx * p.Ny1 + cy) * p.Nzp + cz;
    real f0 = 0, f1 = 0, f2 = 0;
#pragma unroll
    for (int ax = 0; ax < 2; ++ax) {
        const real wx = ax ? dx : (real)1 - dx;
#pragma unroll
        for (int ay = 0; ay < 2; ++ay) {
            const real wxy = wx * (ay ? dy : (real)1 - dy);
            const real* c = B + ((long long)ax * p.Ny1 + ay) * p.Nzp;
            const real w0 = wxy * ((real)1 - dz), w1 = wxy * dz;
            f0 += w0 * __ldg(c) + w1 * __ldg(c + 1);
            f1 += w0 * __ldg(c + p.ghost_elems) + w1 * __ldg(c + p.ghost_elems + 1);
            f2 += w0 * __ldg(c + 2 * p.ghost_elems) + w1 * __ldg(c + 2 * p.ghost_elems + 1);
        }
    }
    if (CHARGE) { const real q = q_sorted[i]; f0 *= q; f1 *= q; f2 *= q; }
    const long long idx = (long long)(rc.meta & (((UT)1 << Tr::IDX_BITS) - 1));
    real* o;
    if (p.rt != nullptr && idx >= p.n_home) {
        const long long e = idx - p.n_home;
        o = reinterpret_cast<real*>(p.ret[(int)(e / p.G)]) + 3 * ((long long)p.rank * p.G + e % p.G);
    } else {
        o = force + (size_t)idx * 3;
    }
    o[0] = f0; o[1] = f1; o[2] = f2;
    if (!more) break;
    rc = rn; i = inext;
    }
}

// HYMD_B200_READOUT_PERSIST=1: persistent grid of 8 CTAs of 256 threads per SM, every thread prefetching its next
// record.  Measured at C4 on one B200 (profiles/r3f_variants.jsonl): 0.261 ms against 0.250 ms with one thread per
// particle (the default) -- the hardware's CTA turnover already keeps enough record loads in flight.
static unsigned int gather_grid(hymd_ctx* c, unsigned int blocks) {
    const char* e = getenv("HYMD_B200_READOUT_PERSIST");
    if (!(e && e[0] == '1')) return blocks;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->dev) != cudaSuccess || sms <= 0) return blocks;
    const unsigned int cap = (unsigned int)sms * 8u;
    return blocks > cap ? cap : blocks;
}

static void gather_params(hymd_ctx* c, GatherParams& p) {
    const Geometry& g = c->g;
    p.n = c->np; p.ghost_elems = g.ghost_elems; p.Ny1 = g.Ny + 1; p.Nzp = g.Nzp;
    p.fbx = g.fbx; p.fby = g.fby; p.fbz = g.fbz;
    p.rt = route_totals(c);
    p.n_home = c->np; p.rank = g.rank;
    route_peer_ret(c, p.ret, &p.G);
    if (p.rt) p.n = c->np > 0 ? c->np : 1;     // grid size only: the live count is rt->n_work
}

template <typename real, bool CHARGE>
static int launch_gather(hymd_ctx* c, void* d_force, cudaStream_t s) {
    using Tr = RTraits<real>;
    GatherParams p;
    gather_params(c, p);
    HYMD_CHECK(route_acquire_return(c, s));
    unsigned int blocks = (unsigned int)((p.n + 255) / 256);
    blocks = gather_grid(c, blocks);
    if (blocks > 0) {
        readout_gather_kernel<real, CHARGE><<<blocks, 256, 0, s>>>(
            (const real*)(CHARGE ? c->emesh : c->gmesh), (const typename Tr::Rec*)c->rec,
            (const real*)c->q_sorted, c->d_urow, (real*)d_force, p);
        HYMD_LAUNCH_CHECK(c);
    }
    return route_return(c, d_force, s);
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
    // A ring of NS box buffers per CTA: NS-1 tiles' boxes are in flight while one is in use.
    // Prefer four stages of <= 50 KB (enough bytes in flight to cover the HBM latency); fall back
    // to two larger stages when that would shrink the tile below 4 x 4 cells.
    static const int cand[][2] = {{4, 8}, {4, 4}, {2, 4}, {2, 2}, {1, 2}, {1, 1}};
    const int tz = READOUT_TZ, bz = READOUT_BZ;
    auto pick_tile = [&](size_t budget) {
        for (int i = 0; i < 6; ++i)
            if ((size_t)c->U * 3 * (cand[i][0] + 1) * (cand[i][1] + 1) * bz * c->rsz <= budget) return i;
        return 5;
    };
    int ns = 2;
    int pick = pick_tile(100 * 1024);
    if (const char* e = getenv("HYMD_B200_READOUT_TILE")) {   // tuning: "<candidate index>,<stages>"
        int a = 0, g = 0;
        if (sscanf(e, "%d,%d", &a, &g) == 2 && a >= 0 && a < 6 && g >= 2 && g <= READOUT_MAX_STAGES) {
            pick = a; ns = g;
        }
    }
    c->rtx = cand[pick][0]; c->rty = cand[pick][1]; c->rtz = tz; c->rbz = bz; c->rstages = ns;
    const size_t box_bytes = (size_t)3 * (c->rtx + 1) * (c->rty + 1) * bz * c->rsz;
    const size_t box_stride = (box_bytes + 127) / 128 * 128;
    const size_t tail = READOUT_MAX_STAGES * sizeof(uint64_t) + HYMD_MAX_TYPES * sizeof(int);
    c->readout_smem = ns * (size_t)c->U * box_stride + tail;
    c->readout_smem_pme = ns * box_stride + tail;
    if (c->readout_smem > 225 * 1024) {
        set_error("readout: %d distinct potential rows need %zu B of shared memory", c->U,
                  c->readout_smem);
        return HYMD_ERR_INVALID;
    }
    HYMD_CHECK(encode_map(c, &c->tmap_gmesh, c->gmesh, 3 * c->U));
    if (c->cfg.pme) HYMD_CHECK(encode_map(c, &c->tmap_emesh, c->emesh, 3));
    return HYMD_OK;
}

template <typename real, bool CHARGE, int TX, int TY>
static int launch_readout_tile(hymd_ctx* c, void* d_force, cudaStream_t s) {
    using Tr = RTraits<real>;
    const Geometry& g = c->g;
    ReadoutParams p;
    p.Nx = g.Nx; p.Ny = g.Ny; p.Nz = g.Nz; p.nxl = g.nxl; p.nbz = g.nbz;
    p.fbx = g.fbx; p.fby = g.fby; p.fbz = g.fbz;
    p.tx = TX; p.ty = TY; p.tz = c->rtz; p.bz = c->rbz;
    p.ntx = (g.nxl + p.tx - 1) / p.tx;
    p.nty = (g.Ny + p.ty - 1) / p.ty;
    p.ntz = (g.Nz + p.tz - 1) / p.tz;
    p.U = CHARGE ? 1 : c->U;
    p.T = c->T;
    p.box_bytes = (unsigned int)((size_t)3 * (p.tx + 1) * (p.ty + 1) * p.bz * sizeof(real));
    p.box_stride = (p.box_bytes + 127u) / 128u * 128u;
    p.nstage = c->rstages;
    const size_t smem = CHARGE ? c->readout_smem_pme : c->readout_smem;
    auto kern = readout_kernel<real, CHARGE, TX, TY>;
    HYMD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long blocks = (long long)p.ntx * p.nty * p.ntz;
    int sms = 0;
    HYMD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->dev));
    if (blocks > sms) blocks = sms;           // persistent: one CTA per SM walks the tiles
    kern<<<(unsigned int)blocks, READOUT_NT, smem, s>>>(CHARGE ? c->tmap_emesh : c->tmap_gmesh,
                                                 (const typename Tr::Rec*)c->rec,
                                                 (const real*)c->q_sorted, c->cell_start, c->d_urow,
                                                 (real*)d_force, p);
    HYMD_LAUNCH_CHECK(c);
    return HYMD_OK;
}

// Default: the direct gather (measured faster at C4: 0.25 ms vs 0.30 ms for the TMA-staged tiles
// with cell-ordered callers); HYMD_B200_READOUT=tma selects the TMA-staged kernel.
static bool use_gather(const hymd_ctx* c) {
    if (c->g.P > 1) return true;      // guests' forces go to their owners (per-step routing): gather kernel only
    const char* e = getenv("HYMD_B200_READOUT");
    return !(e && e[0] == 't');
}

template <typename real, bool CHARGE>
static int launch_readout(hymd_ctx* c, void* d_force, cudaStream_t s) {
    if (use_gather(c)) return launch_gather<real, CHARGE>(c, d_force, s);
    switch (c->rtx * 16 + c->rty) {
        case 4 * 16 + 8: return launch_readout_tile<real, CHARGE, 4, 8>(c, d_force, s);
        case 4 * 16 + 4: return launch_readout_tile<real, CHARGE, 4, 4>(c, d_force, s);
        case 2 * 16 + 4: return launch_readout_tile<real, CHARGE, 2, 4>(c, d_force, s);
        case 2 * 16 + 2: return launch_readout_tile<real, CHARGE, 2, 2>(c, d_force, s);
        case 1 * 16 + 2: return launch_readout_tile<real, CHARGE, 1, 2>(c, d_force, s);
        case 1 * 16 + 1: return launch_readout_tile<real, CHARGE, 1, 1>(c, d_force, s);
        default: break;
    }
    set_error("readout: no kernel for tile %d x %d", c->rtx, c->rty);
    return HYMD_ERR_INVALID;
}

// Per-type gather from a caller-chosen set of ghost-padded force meshes (3 per row of d_urow): the
// GPE electrostatic forces (gpe.cu) read 3T meshes with the identity type -> row map.
template <typename real>
static int launch_gather_custom(hymd_ctx* c, const void* mesh, const int* d_urow, void* d_force, cudaStream_t s) {
    using Tr = RTraits<real>;
    GatherParams p;
    gather_params(c, p);
    HYMD_CHECK(route_acquire_return(c, s));
    const unsigned int blocks = (unsigned int)((p.n + 255) / 256);
    if (blocks > 0) {
        readout_gather_kernel<real, false><<<blocks, 256, 0, s>>>(
            (const real*)mesh, (const typename Tr::Rec*)c->rec, (const real*)c->q_sorted, d_urow, (real*)d_force, p);
        HYMD_LAUNCH_CHECK(c);
    }
    return route_return(c, d_force, s);       // several slabs: guests' rows go back to their owners
}

int readout_custom(hymd_ctx* c, const void* mesh, const int* d_urow, void* d_force, cudaStream_t s) {
    return c->f64 ? launch_gather_custom<double>(c, mesh, d_urow, d_force, s)
                  : launch_gather_custom<float>(c, mesh, d_urow, d_force, s);
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
