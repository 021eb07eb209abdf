// (y,z) plane transforms of the slab pipeline: the batched 2-D r2c / c2r of RealField.r2c /
// ComplexField.c2r (field.py:576-578, 584, 613, 616, 366, 377, 397) done in ONE pass over HBM
// per direction instead of the two passes (y c2c, z r2c/c2r) a library 2-D plan makes.
//
// One persistent CTA per SM takes a whole (field, x-plane) unit at a time:
//
//   inverse (c2r)   phase 1: chunks of 8 kz-columns of the spectrum, inverse FFT along y (four
//                            groups of 128 threads, each on its own chunk), stored to a per-CTA
//                            scratch plane that is reused for every unit and therefore lives in L2;
//                   phase 2: row pairs (y, y+1) of the scratch plane -> ONE complex inverse FFT
//                            along z of A + iB (Hermitian extension built on the fly; the imaginary
//                            parts of k_z = 0 and k_z = Nz/2 are dropped = c2r semantics), real part
//                            -> row y, imaginary part -> row y+1, every warp on its own row pairs,
//                            written straight into the ghost-padded force-mesh layout INCLUDING
//                            the periodic images (z = Nz element, y = Ny row, x = nxl plane), so
//                            no separate ghost-fill pass is needed.
//   forward (r2c)   the mirror image: row pairs of the real plane -> one complex FFT along z,
//                   split into the two half spectra, scratch plane, then column chunks along y.
//
// HBM traffic per unit = read the plane once + write it once; the y<->z exchange goes through
// the L2-resident scratch (148 CTAs x Ny x (Nz/2+1) complex = 39 MB at 256^2 fp32).
// Power-of-two Ny, Nz only; other mesh sizes keep the cuFFT plans of slabfft.cu.
#include <stdlib.h>

#include "ctx.cuh"
#include "fft.cuh"

namespace hymd {

// ---- shared-memory layouts ---------------------------------------------------------------------
// LayA: FFT along the strided axis, CH columns contiguous (column chunks).
template <int N, int CH>
struct LayA {
    static constexpr int R2 = Radix<N>::R2;
    static constexpr int ELEMS = (N + N / R2) * CH;
    __device__ static __forceinline__ int at(int pos, int c) { return (pos + pos / R2) * CH + c; }
    // pos = hi * R2 + lo with lo < R2 (no division; with a compile-time hi or lo the compiler folds
    // the product into the address offset)
    __device__ static __forceinline__ int at2(int hi, int lo, int c) { return (hi * (R2 + 1) + lo) * CH + c; }
};
// LayB: FFT along the contiguous axis, one padded row per transform (row pairs).
template <int N, int CP>
struct LayB {
    static constexpr int R2 = Radix<N>::R2;
    static constexpr int PITCH = N + N / R2;
    static constexpr int ELEMS = PITCH * CP;
    __device__ static __forceinline__ int at(int pos, int c) { return c * PITCH + pos + pos / R2; }
    __device__ static __forceinline__ int at2(int hi, int lo, int c) { return c * PITCH + hi * (R2 + 1) + lo; }
};

// ---- memory access helpers ---------------------------------------------------------------------
__device__ __forceinline__ Cx<float> ld_stream(const Cx<float>* p) {      // read once from HBM
    const float2 t = __ldcs(reinterpret_cast<const float2*>(p)); return {t.x, t.y};
}
__device__ __forceinline__ Cx<double> ld_stream(const Cx<double>* p) {
    const double2 t = __ldcs(reinterpret_cast<const double2*>(p)); return {t.x, t.y};
}
__device__ __forceinline__ Cx<float> ld_l2(const Cx<float>* p) {          // scratch plane: L2 only
    const float2 t = __ldcg(reinterpret_cast<const float2*>(p)); return {t.x, t.y};
}
__device__ __forceinline__ Cx<double> ld_l2(const Cx<double>* p) {
    const double2 t = __ldcg(reinterpret_cast<const double2*>(p)); return {t.x, t.y};
}
__device__ __forceinline__ void st_stream(Cx<float>* p, Cx<float> v) {
    __stcs(reinterpret_cast<float2*>(p), make_float2(v.x, v.y));
}
__device__ __forceinline__ void st_stream(Cx<double>* p, Cx<double> v) {
    __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y));
}
__device__ __forceinline__ void st_l2(Cx<float>* p, Cx<float> v) {
    __stcg(reinterpret_cast<float2*>(p), make_float2(v.x, v.y));
}
__device__ __forceinline__ void st_l2(Cx<double>* p, Cx<double> v) {
    __stcg(reinterpret_cast<double2*>(p), make_double2(v.x, v.y));
}
// Scratch plane with explicit L2 eviction priorities (HYMD_B200_SCR_HINT, PlaneParams::scr_hint):
// written "evict last" (it must survive in L2 until the other phase reads it, next to the
// streaming traffic of 148 CTAs), read "evict first" (dead after the read).
__device__ __forceinline__ uint64_t l2_policy_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ Cx<float> ld_hint(const Cx<float>* p, uint64_t pol) {
    Cx<float> v;
    asm volatile("ld.global.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ Cx<double> ld_hint(const Cx<double>* p, uint64_t pol) {
    Cx<double> v;
    asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_hint(Cx<float>* p, Cx<float> v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint(Cx<double>* p, Cx<double> v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void group_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// ---- bulk async copies (TMA, non-tensor form) + mbarrier -------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PLANE_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra PLANE_WAIT_DONE;\n"
        "bra PLANE_WAIT_LOOP;\n"
        "PLANE_WAIT_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(parity)
        : "memory");
}
// global -> shared, completion counted in bytes on an mbarrier (16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst_smem)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_store(void* dst, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(smem_addr(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy writes (st.shared / st.global) made visible to the async proxy (the copy engine)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__device__ __forceinline__ float shfl(float v, int lane) { return __shfl_sync(0xffffffffu, v, lane); }
__device__ __forceinline__ double shfl(double v, int lane) { return __shfl_sync(0xffffffffu, v, lane); }

struct PlaneParams {
    int nunits, nplanes;            // units = fields x planes
    int f0;                         // first field (potential row in derive mode) of this launch: the slab pipeline
                                    // transforms field by field while the previous field crosses NVLink
    long long k_fs, k_xs;           // complex strides (field, plane) of the spectra, row pitch Nzcp
    long long r_fs, r_xs;           // real strides (field, plane)
    int r_ys;                       // real row pitch
    int ghost;                      // inverse: also write the z = Nz element and the y = Ny row
    int xdup_plane;                 // inverse: plane 0 is written to this plane as well (-1: no)
    // inverse, "derive" mode: a unit is (potential row u, plane) and produces the three force
    // meshes 3u+d from TWO input spectra: slot 2u = -i k_x V (x-inverted), slot 2u+1 = -i V;
    // k_y / k_z (with the Nyquist rule of SURVEY.md section 7) are applied while loading.
    int derive;
    double dky, dkz;                // 2 pi / L_y, 2 pi / L_z
    int scr_hint;                   // scratch plane accesses carry L2 eviction priorities
    int scr_alt;                    // experiment: alternate between two scratch planes per CTA (L2 footprint x2)
    int row_tma;                    // staged row phase (Cfg::TMA_ROWS): bit 0 = inputs, bit 1 = outputs via bulk copies
    // forward, several slabs: the last butterfly stage stores every spectrum row straight into the k
    // buffer of the rank that owns its k_y range (the forward transpose of the slab FFT, fused):
    // row ky of local plane x, field f  ->  peer[ky >> nyl_shift] + (x0 + x) pk_xs + f pk_fs + (ky & (nyl-1)) Nzcp
    int push, nyl_shift, x0;
    long long pk_xs, pk_fs;
    void* peer[HYMD_MAX_PEERS];
    // inverse, several slabs, "blocked" exchange: the input is the receive buffer of the inverse transpose as the
    // peers' contiguous copies left it, W[q][x][f][kyl][kz] (block q = the k_y range rank q transformed along x):
    // row ky of plane x, spectrum f  ->  in + (ky >> nyl_shift) blk_q + x blk_x + f blk_f + (ky & (nyl-1)) Nzcp
    int vec_out;                    // tensor-memory inverse: finished rows staged in shared memory, 16-byte stores
    int blk_in;
    long long blk_q, blk_x, blk_f;
    long long blk_d;                // blk_q - nyl Nzcp: what crossing into the next block adds to a row offset
};

// first element of spectrum row ky (plane x, spectrum f) of the inverse kernels' input
template <typename real>
__device__ __forceinline__ const Cx<real>* in_row(const Cx<real>* __restrict__ in, const PlaneParams& p, int f, int x,
                                                   int ky, int nzcp) {
    if (p.blk_in)
        return in + (long long)(ky >> p.nyl_shift) * p.blk_q + (long long)x * p.blk_x + (long long)f * p.blk_f +
               (long long)(ky & ((1 << p.nyl_shift) - 1)) * nzcp;
    return in + f * p.k_fs + x * p.k_xs + (long long)ky * nzcp;
}

// The column phase of the inverse kernels reads, per thread, rows k1 + R k2 (k2 = 0 .. R2-1) of one plane.  Their
// addresses are one base pointer (row k1: in_row, once per transform) plus row_step: a compile-time multiple of the
// row pitch -- an immediate of the load instruction -- and, for the blocked layout only, the block term.  (It was
// ~17 integer instructions and a branch per 8-byte load, as many as the butterflies themselves: profiles/r2_sass_tmem.md.)
template <bool BLK, int R, int NZCP_>
__device__ __forceinline__ long long row_step(int k1, int k2, int q0, int blk_shift, long long blk_d) {
    long long o = (long long)k2 * (R * NZCP_);
    if (BLK) {
        int q = ((k1 + R * k2) >> blk_shift) - q0;
        asm volatile("" : "+r"(q));        // evaluated at the load: sixteen hoisted 64-bit offsets would spill
        o += (long long)q * blk_d;
    }
    return o;
}

// Work decomposition inside the 512-thread CTA.
//   column phase: NG groups of GT threads, each on its own chunk of CG kz-columns (64 bytes per
//                 spectrum row), synchronised with a named barrier per group;
//   row phase   : every warp on its own CW row pairs, synchronised with __syncwarp only.
// Butterfly inputs are loaded from global memory straight into registers and the last butterfly
// stage stores straight to global memory, so each FFT makes exactly one trip through shared memory.
template <typename real, int NY, int NZ, int NT_, int TILES> struct PlaneCfg {
    static constexpr int NT = NT_, NW = NT / 32, CTAS = 512 / NT;            // CTAs per SM
    static constexpr int R1y = Radix<NY>::R1, R2y = Radix<NY>::R2, R1z = Radix<NZ>::R1, R2z = Radix<NZ>::R2;
    static constexpr int NZC = NZ / 2 + 1, NZCP = NZC + (NZC & 1);
    static constexpr int CG = (TILES == 3 ? 128 : 64) / (int)sizeof(Cx<real>);   // columns per chunk: 64- or 128-byte rows
    static constexpr int GT = CG * 16, NG = NT / GT;
    using LA = LayA<NY, CG>;
    static constexpr int NTILE = (TILES == 2 && 2 * NG * LA::ELEMS * (int)sizeof(Cx<real>) <= 144 * 1024 / CTAS) ? 2 : 1;
    static constexpr int CW = 32 / R1z;                                    // row pairs per warp pass
    using LB = LayB<NZ, CW>;
    static constexpr int COL_ELEMS = NG * NTILE * LA::ELEMS, ROW_ELEMS = NW * LB::ELEMS;
    // Row phase of the inverse, fp32: every warp stages its input rows (CW row pairs of the scratch
    // plane = ONE contiguous block) and its output rows in shared memory and moves them with bulk
    // async copies (TMA) instead of 8-byte loads / 4-byte stores: the load/store unit was the
    // limiter (ncu: lg_throttle + long_scoreboard = 45 % of the stall samples).
    static constexpr int IN_ELEMS = CW * 2 * NZCP;                    // complex, per warp
    static constexpr bool HAS_IN = TILES == 2;                        // single-tile build: no staged inputs (L1 matters more)
    static constexpr int IN_ALLOC = HAS_IN ? IN_ELEMS : 0;
    static constexpr int OUT_REALS = CW * 2 * (NZ + 4) + 32 * CW;     // reals, per warp (ghost pitch + bank gaps)
    static constexpr size_t ROW_BYTES = sizeof(Cx<real>) * (size_t)ROW_ELEMS;
    static constexpr size_t ROW_IN_BYTES = ROW_BYTES + (size_t)NW * IN_ALLOC * sizeof(Cx<real>);
    static constexpr size_t ROW_TMA_BYTES = ROW_IN_BYTES + (size_t)NW * OUT_REALS * sizeof(real);
    static constexpr bool TMA_ROWS = TILES == 2 && sizeof(real) == 4 && (NZCP % 2 == 0) && (NZ % 4 == 0) &&
                                     ROW_TMA_BYTES + sizeof(Cx<real>) * (NY + NZ) <= (size_t)216 * 1024 / CTAS;
    static constexpr size_t COL_BYTES = sizeof(Cx<real>) * (size_t)COL_ELEMS;
    static constexpr int TILE = COL_ELEMS > ROW_ELEMS ? COL_ELEMS : ROW_ELEMS;
    static constexpr size_t SMEM = sizeof(Cx<real>) * (size_t)(NY + NZ + TILE);
    // inverse kernel: staged inputs only / staged inputs + outputs
    static constexpr size_t smem_inv(bool staged_out) {
        size_t row = TMA_ROWS ? (staged_out ? ROW_TMA_BYTES : ROW_IN_BYTES) : ROW_BYTES;
        return sizeof(Cx<real>) * (size_t)(NY + NZ) + (COL_BYTES > row ? COL_BYTES : row);
    }
};

// ---- inverse: spectra [ky][kz] -> real plane ------------------------------------------------------
template <typename real, int NY, int NZ, int NTH, int TILES, bool BLK>
__global__ void __launch_bounds__(NTH, 512 / NTH) plane_c2r_kernel(
    const Cx<real>* __restrict__ in, Cx<real>* __restrict__ scratch, real* __restrict__ out,
    const Cx<real>* __restrict__ twy_g, const Cx<real>* __restrict__ twz_g, PlaneParams p) {
    using Cfg = PlaneCfg<real, NY, NZ, NTH, TILES>;
    using LA = typename Cfg::LA;
    using LB = typename Cfg::LB;
    constexpr int NT = Cfg::NT, NW = Cfg::NW, CG = Cfg::CG, GT = Cfg::GT, NG = Cfg::NG, NTILE = Cfg::NTILE,
                  CW = Cfg::CW, NZC = Cfg::NZC, NZCP = Cfg::NZCP;
    constexpr int R1y = Cfg::R1y, R2y = Cfg::R2y, R1z = Cfg::R1z, R2z = Cfg::R2z;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cx<real>* twy = reinterpret_cast<Cx<real>*>(smem_raw);
    Cx<real>* twz = twy + NY;
    Cx<real>* tile = twz + NZ;
    const int tid = threadIdx.x;
    const uint64_t pol_last = l2_policy_last(), pol_first = l2_policy_first();
    for (int i = tid; i < NY; i += NT) twy[i] = twy_g[i];
    for (int i = tid; i < NZ; i += NT) twz[i] = twz_g[i];
    __syncthreads();
    Cx<real>* scr0 = scratch + (size_t)blockIdx.x * NY * NZCP;
    Cx<real>* scr = scr0;
    constexpr int NCHG = (NZC + CG - 1) / CG;             // the pad column is never read
    const int g = tid / GT, gt = tid % GT;
    const int warp = tid / 32, lane = tid % 32;
    Cx<real>* gtile = tile + g * NTILE * LA::ELEMS;
    Cx<real>* wtile = tile + warp * LB::ELEMS;
    // one step-A task per thread (k1 fixed for the whole kernel): prefetchable, twiddles in registers
    constexpr bool ONE_TASK = (CG * R1y <= GT) && sizeof(real) == 4 && R2y <= 16;
    constexpr bool REGC = ONE_TASK && R2y >= 4, REGR = sizeof(real) == 4 && R2z <= 16 && R2z >= 4;
    TwiddleRegs<real, R2y> twc;
    TwiddleRegs<real, R2z> twr;
    if constexpr (REGC) twc.init(twy, (gt / CG) % R1y, NY);
    if constexpr (REGR) twr.init(twz, lane % R1z, NZ);
    // staged row phase (Cfg::TMA_ROWS): per-warp input rows, output rows and one mbarrier
    __shared__ __align__(8) uint64_t row_bar[NW];
    unsigned char* row_stage = reinterpret_cast<unsigned char*>(tile) + Cfg::ROW_BYTES;
    Cx<real>* inb = reinterpret_cast<Cx<real>*>(row_stage) + warp * Cfg::IN_ALLOC;
    real* outb = reinterpret_cast<real*>(row_stage + (size_t)NW * Cfg::IN_ALLOC * sizeof(Cx<real>)) +
                 warp * Cfg::OUT_REALS;
    // output row rr of this warp starts at rr * r_ys + (rr / 2) * ogap: row pairs stay contiguous
    // (one bulk store each) and the two pairs of a pass fall into different banks
    const int ogap = (48 - (2 * p.r_ys) % 32) % 32;
    uint32_t row_parity = 0;
    if constexpr (Cfg::TMA_ROWS) {
        if (lane == 0) mbar_init(&row_bar[warp], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }

    const int ND = p.derive ? 3 : 1;
    const real dky = (real)p.dky, dkz = (real)p.dkz;
    const int blk_shift = BLK ? p.nyl_shift : 0;
    const long long blk_d = BLK ? p.blk_d : 0;
    for (int unit = blockIdx.x; unit < p.nunits; unit += gridDim.x) {
      const int fu = unit / p.nplanes + p.f0, x = unit % p.nplanes;
      for (int d = 0; d < ND; ++d) {
        // derive: d = 1, 2 read the same spectrum plane (the second time from L2)
        const int fin = p.derive ? 2 * fu + (d > 0 ? 1 : 0) : fu;
        const int f = p.derive ? 3 * fu + d : fu;
        const int mode = p.derive ? d : 0;
        if (p.scr_alt) scr = scr0 + (size_t)((unit / gridDim.x * 3 + d) & 1) * gridDim.x * NY * NZCP;
        const Cx<real>* src = in + fin * p.k_fs + x * p.k_xs;
        // v[k2] = spectrum element (k_y = k1 + R1y k2, column col): times k_y (mode 1) or k_z (mode 2)
        auto scale_chunk = [&](Cx<real> (&v)[R2y], int k1, int col) {
            if (mode == 1) {
                const bool selfc = col == 0 || 2 * col == NZ;
                const real fk1 = (real)k1;
#pragma unroll
                for (int k2 = 0; k2 < R2y; ++k2) {
                    // signed frequency of row k1 + R1y k2 (rows >= NY/2 <=> k2 >= R2y/2)
                    const real nn = fk1 + (real)(R1y * k2 - (2 * k2 >= R2y ? NY : 0));
                    real sk = dky * nn;
                    if (2 * k2 == R2y && k1 == 0 && selfc) sk = 0;      // y-Nyquist on a self-conjugate column
                    v[k2].x *= sk; v[k2].y *= sk;
                }
            } else if (mode == 2) {
                const real sk = (2 * col == NZ) ? (real)0 : dkz * (real)col;
#pragma unroll
                for (int k2 = 0; k2 < R2y; ++k2) { v[k2].x *= sk; v[k2].y *= sk; }
            }
        };
        // ---------------- column phase: inverse FFT along y ----------------
        // (the spectrum comes from HBM: the next chunk's butterfly inputs are loaded into a second
        // register set while the current chunk is transformed)
        {
            Cx<real> nx[R2y];
            auto load_chunk = [&](int ch, int task, Cx<real> (&dst)[R2y]) {
                const int c = task % CG, k1 = task / CG;
                const bool valid = ch < NCHG && task < CG * R1y && ch * CG + c < NZC;
                const Cx<real>* e0 = in_row<real>(in, p, fin, x, k1, NZCP) + ch * CG + c;
                const int q0 = k1 >> blk_shift;
#pragma unroll
                for (int k2 = 0; k2 < R2y; ++k2)
                {
                    const Cx<real>* e = e0 + row_step<BLK, R1y, NZCP>(k1, k2, q0, blk_shift, blk_d);
                    // derive: the plane read for the k_y component is read again for k_z: keep it in L2
                    dst[k2] = valid ? (mode == 1 ? ld_l2(e) : ld_stream(e)) : Cx<real>{0, 0};
                }
            };
            int it = 0;
            const int ch0 = (g + unit + d) % NG;
            if (ONE_TASK) load_chunk(ch0, gt, nx);
            for (int ch = ch0; ch < NCHG; ch += NG, ++it) {
                Cx<real>* cur = gtile + (it & (NTILE - 1)) * LA::ELEMS;
                const int c0 = ch * CG;
                for (int task = gt; task < CG * R1y; task += GT) {
                    const int c = task % CG, k1 = task / CG;
                    Cx<real> v[R2y];
                    if (ONE_TASK) {
#pragma unroll
                        for (int k2 = 0; k2 < R2y; ++k2) v[k2] = nx[k2];
                        load_chunk(ch + NG, gt, nx);
                    } else {
                        load_chunk(ch, task, v);
                    }
                    scale_chunk(v, k1, c0 + c);
                    dft_reg<real, R2y, +1>(v);
                    if constexpr (REGC) {
                        twc.template apply<true>(v);
#pragma unroll
                        for (int n2 = 0; n2 < R2y; ++n2) cur[LA::at2(k1, n2, c)] = v[n2];
                    } else {
#pragma unroll
                        for (int n2 = 0; n2 < R2y; ++n2) {
                            Cx<real> w = twy[(n2 * k1) & (NY - 1)];
                            w.y = -w.y;
                            cur[LA::at2(k1, n2, c)] = (k1 == 0) ? v[n2] : cmul(w, v[n2]);
                        }
                    }
                }
                group_sync(g + 1, GT);
                for (int task = gt; task < CG * R2y; task += GT) {
                    const int c = task % CG, n2 = task / CG;
                    Cx<real> v[R1y];
#pragma unroll
                    for (int k1 = 0; k1 < R1y; ++k1) v[k1] = cur[LA::at2(k1, n2, c)];
                    dft_reg<real, R1y, +1>(v);
                    if (c0 + c < NZC) {
#pragma unroll
                        for (int n1 = 0; n1 < R1y; ++n1)
                        {
                            Cx<real>* e = scr + (long long)(n1 * R2y + n2) * NZCP + c0 + c;
                            if (p.scr_hint) st_hint(e, v[n1], pol_last); else st_l2(e, v[n1]);
                        }
                    }
                }
                if (NTILE == 1) group_sync(g + 1, GT);
            }
        }
        if constexpr (Cfg::TMA_ROWS) fence_async_global();   // scratch stores -> visible to the copy engine
        __syncthreads();
        // ---------------- row phase: c2r along z on row pairs (A + iB) ----------------
        real* obase = out + f * p.r_fs + x * p.r_xs;
        real* obase2 = (p.xdup_plane >= 0 && x == 0) ? out + f * p.r_fs + p.xdup_plane * p.r_xs : nullptr;
        if constexpr (Cfg::TMA_ROWS) {
            constexpr int NPG = (NY / 2) / CW;
            constexpr uint32_t IN_BYTES = Cfg::IN_ELEMS * sizeof(Cx<real>);
            // the CW row pairs of a pass are 2 CW consecutive rows of the scratch plane: one copy
            const bool tin = Cfg::HAS_IN && (p.row_tma & 1), tout = p.row_tma & 2;
            if (tin && lane == 0 && warp < NPG) {
                mbar_expect_tx(&row_bar[warp], IN_BYTES);
                bulk_load(inb, scr + (long long)(2 * warp * CW) * NZCP, IN_BYTES, &row_bar[warp]);
            }
            for (int pg = warp; pg < NPG; pg += NW) {
                const int c = lane / R1z, k1 = lane % R1z;
                Cx<real> v[R2z];
                if (tin) {
                    mbar_wait(&row_bar[warp], row_parity);
                    row_parity ^= 1;
                }
                {
                    const Cx<real>* rowA = tin ? inb + (2 * c) * NZCP : scr + (long long)(2 * (pg * CW + c)) * NZCP;
                    const Cx<real>* rowB = rowA + NZCP;
#pragma unroll
                    for (int k2 = 0; k2 < R2z; ++k2) {
                        const int k = k1 + R1z * k2;
                        const int kk = (2 * k <= NZ) ? k : NZ - k;
                        const Cx<real> A = tin ? rowA[kk] : ld_l2(rowA + kk), B = tin ? rowB[kk] : ld_l2(rowB + kk);
                        real sg = (2 * k < NZ) ? (real)1 : (real)-1;         // mirrored half: conjugates
                        if (k == 0 || 2 * k == NZ) sg = 0;                   // c2r drops these imaginary parts
                        v[k2] = {A.x - sg * B.y, sg * A.y + B.x};
                    }
                }
                __syncwarp();                                               // staged rows consumed
                if (tin && lane == 0 && pg + NW < NPG) {                    // next pass flies during this one
                    mbar_expect_tx(&row_bar[warp], IN_BYTES);
                    bulk_load(inb, scr + (long long)(2 * (pg + NW) * CW) * NZCP, IN_BYTES, &row_bar[warp]);
                }
                dft_reg<real, R2z, +1>(v);
                if constexpr (REGR) {
                    twr.template apply<true>(v);
#pragma unroll
                    for (int n2 = 0; n2 < R2z; ++n2) wtile[LB::at2(k1, n2, c)] = v[n2];
                } else {
#pragma unroll
                    for (int n2 = 0; n2 < R2z; ++n2) {
                        Cx<real> w = twz[(n2 * k1) & (NZ - 1)];
                        w.y = -w.y;
                        wtile[LB::at2(k1, n2, c)] = (k1 == 0) ? v[n2] : cmul(w, v[n2]);
                    }
                }
                __syncwarp();
                if (tout && lane < CW) bulk_wait_read();                    // previous pass's stores have left outb
                __syncwarp();
                for (int task = lane; task < CW * R2z; task += 32) {
                    const int c2 = task / R2z, n2 = task % R2z;
                    Cx<real> u[R1z];
#pragma unroll
                    for (int k1b = 0; k1b < R1z; ++k1b) u[k1b] = wtile[LB::at2(k1b, n2, c2)];
                    dft_reg<real, R1z, +1>(u);
                    if (!tout) {                                            // straight from the registers
                        const int y0 = 2 * (pg * CW + c2);
#pragma unroll
                        for (int dup = 0; dup < 2; ++dup) {
                            real* ob = dup ? obase2 : obase;
                            if (ob == nullptr) continue;
                            real* og = ob + (long long)y0 * p.r_ys + n2;
#pragma unroll
                            for (int n1 = 0; n1 < R1z; ++n1) {
                                __stcs(og + n1 * R2z, u[n1].x);
                                __stcs(og + p.r_ys + n1 * R2z, u[n1].y);
                            }
                            if (p.ghost) {
                                if (n2 == 0) { og[NZ] = u[0].x; og[p.r_ys + NZ] = u[0].y; }
                                if (y0 == 0) {
                                    real* oy = ob + (long long)NY * p.r_ys + n2;
#pragma unroll
                                    for (int n1 = 0; n1 < R1z; ++n1) __stcs(oy + n1 * R2z, u[n1].x);
                                    if (n2 == 0) oy[NZ] = u[0].x;
                                }
                            }
                        }
                        continue;
                    }
                    real* oa = outb + (2 * c2) * p.r_ys + c2 * ogap + n2;
#pragma unroll
                    for (int n1 = 0; n1 < R1z; ++n1) {
                        oa[n1 * R2z] = u[n1].x;
                        oa[p.r_ys + n1 * R2z] = u[n1].y;
                    }
                    if (p.ghost && n2 < p.r_ys - NZ) {      // periodic image z = Nz, zeros in the row padding
                        oa[NZ] = n2 == 0 ? u[0].x : (real)0;
                        oa[p.r_ys + NZ] = n2 == 0 ? u[0].y : (real)0;
                    }
                }
                if (tout) fence_async_smem();
                __syncwarp();
                if (tout && lane < CW) {                                    // lane = row pair: 2 rows, contiguous
                    const int y0 = 2 * (pg * CW + lane);
                    const real* sb = outb + (2 * lane) * p.r_ys + lane * ogap;
                    const uint32_t pair_bytes = 2u * (uint32_t)p.r_ys * (uint32_t)sizeof(real);
#pragma unroll
                    for (int dup = 0; dup < 2; ++dup) {
                        real* ob = dup ? obase2 : obase;
                        if (ob == nullptr) continue;
                        bulk_store(ob + (long long)y0 * p.r_ys, sb, pair_bytes);
                        if (p.ghost && y0 == 0) bulk_store(ob + (long long)NY * p.r_ys, sb, pair_bytes / 2);
                    }
                    bulk_commit();
                }
            }
            if (tout && lane < CW) bulk_wait_read();                        // before the tiles are reused
        } else
        for (int pg = warp; pg < (NY / 2) / CW; pg += NW) {
            {   // exactly 32 tasks: (row pair c, k1)
                const int c = lane / R1z, k1 = lane % R1z;
                const Cx<real>* rowA = scr + (long long)(2 * (pg * CW + c)) * NZCP;
                const Cx<real>* rowB = rowA + NZCP;
                Cx<real> v[R2z];
#pragma unroll
                for (int k2 = 0; k2 < R2z; ++k2) {
                    const int k = k1 + R1z * k2;
                    const int kk = (2 * k <= NZ) ? k : NZ - k;
                    const Cx<real> A = p.scr_hint ? ld_hint(rowA + kk, pol_first) : ld_l2(rowA + kk),
                                   B = p.scr_hint ? ld_hint(rowB + kk, pol_first) : ld_l2(rowB + kk);
                    real s = (2 * k < NZ) ? (real)1 : (real)-1;          // mirrored half: conjugates
                    if (k == 0 || 2 * k == NZ) s = 0;                    // c2r drops these imaginary parts
                    v[k2] = {A.x - s * B.y, s * A.y + B.x};
                }
                dft_reg<real, R2z, +1>(v);
                if constexpr (REGR) {
                    twr.template apply<true>(v);
#pragma unroll
                    for (int n2 = 0; n2 < R2z; ++n2) wtile[LB::at2(k1, n2, c)] = v[n2];
                } else {
#pragma unroll
                    for (int n2 = 0; n2 < R2z; ++n2) {
                        Cx<real> w = twz[(n2 * k1) & (NZ - 1)];
                        w.y = -w.y;
                        wtile[LB::at2(k1, n2, c)] = (k1 == 0) ? v[n2] : cmul(w, v[n2]);
                    }
                }
            }
            __syncwarp();
            for (int task = lane; task < CW * R2z; task += 32) {
                const int c = task / R2z, n2 = task % R2z;
                Cx<real> v[R1z];
#pragma unroll
                for (int k1 = 0; k1 < R1z; ++k1) v[k1] = wtile[LB::at2(k1, n2, c)];
                dft_reg<real, R1z, +1>(v);
                const int y0 = 2 * (pg * CW + c);
#pragma unroll
                for (int dup = 0; dup < 2; ++dup) {
                    real* ob = dup ? obase2 : obase;
                    if (ob == nullptr) continue;
                    real* oa = ob + (long long)y0 * p.r_ys + n2;
#pragma unroll
                    for (int n1 = 0; n1 < R1z; ++n1) {
                        __stcs(oa + n1 * R2z, v[n1].x);
                        __stcs(oa + p.r_ys + n1 * R2z, v[n1].y);
                    }
                    if (p.ghost) {
                        if (n2 == 0) { oa[NZ] = v[0].x; oa[p.r_ys + NZ] = v[0].y; }
                        if (y0 == 0) {
                            real* og = ob + (long long)NY * p.r_ys + n2;
#pragma unroll
                            for (int n1 = 0; n1 < R1z; ++n1) __stcs(og + n1 * R2z, v[n1].x);
                            if (n2 == 0) og[NZ] = v[0].x;
                        }
                    }
                }
            }
            __syncwarp();
        }
        __syncthreads();
      }
    }
}

// ---- forward: real plane -> spectra [ky][kz] -----------------------------------------------------
template <typename real, int NY, int NZ, int NTH, int TILES>
__global__ void __launch_bounds__(NTH, 512 / NTH) plane_r2c_kernel(
    const real* __restrict__ in, Cx<real>* __restrict__ scratch, Cx<real>* __restrict__ out,
    const Cx<real>* __restrict__ twy_g, const Cx<real>* __restrict__ twz_g, PlaneParams p) {
    using Cfg = PlaneCfg<real, NY, NZ, NTH, TILES>;
    using LA = typename Cfg::LA;
    using LB = typename Cfg::LB;
    constexpr int NT = Cfg::NT, NW = Cfg::NW, CG = Cfg::CG, GT = Cfg::GT, NG = Cfg::NG, NTILE = Cfg::NTILE,
                  CW = Cfg::CW, NZC = Cfg::NZC, NZCP = Cfg::NZCP;
    constexpr int R1y = Cfg::R1y, R2y = Cfg::R2y, R1z = Cfg::R1z, R2z = Cfg::R2z;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cx<real>* twy = reinterpret_cast<Cx<real>*>(smem_raw);
    Cx<real>* twz = twy + NY;
    Cx<real>* tile = twz + NZ;
    const int tid = threadIdx.x;
    for (int i = tid; i < NY; i += NT) twy[i] = twy_g[i];
    for (int i = tid; i < NZ; i += NT) twz[i] = twz_g[i];
    __syncthreads();
    Cx<real>* scr = scratch + (size_t)blockIdx.x * NY * NZCP;
    constexpr int NCHG = (NZCP + CG - 1) / CG;            // the pad column is written (zeros)
    const int g = tid / GT, gt = tid % GT;
    const int warp = tid / 32, lane = tid % 32;
    Cx<real>* gtile = tile + g * NTILE * LA::ELEMS;
    Cx<real>* wtile = tile + warp * LB::ELEMS;
    // one task per thread in the twiddled stages (n2 fixed for the whole kernel): twiddles in registers
    constexpr bool ONE_ROW_TASK = (CW * R2z == 32) && sizeof(real) == 4;
    constexpr bool REGR = ONE_ROW_TASK && R1z <= 16 && R1z >= 4;
    constexpr bool REGC = sizeof(real) == 4 && (CG * R2y <= GT) && R1y <= 16 && R1y >= 4;
    TwiddleRegs<real, R1z> twr;
    TwiddleRegs<real, R1y> twc;
    if constexpr (REGR) twr.init(twz, lane % R2z, NZ);
    if constexpr (REGC) twc.init(twy, (gt / CG) % R2y, NY);

    for (int unit = blockIdx.x; unit < p.nunits; unit += gridDim.x) {
        const int f = unit / p.nplanes + p.f0, x = unit % p.nplanes;
        const real* src = in + f * p.r_fs + x * p.r_xs;
        // ---------------- row phase: one complex FFT along z per row pair (a + ib) ----------------
        // (the plane comes from HBM: the next pass's inputs are loaded into a second register set
        // while the current pass is transformed)
        constexpr int NPG = (NY / 2) / CW;
        Cx<real> nx[R1z];
        auto load_pair = [&](int pg, int task, Cx<real> (&dst)[R1z]) {
            const int c = task / R2z, n2 = task % R2z;
            const bool valid = pg < NPG;
            const real* ra = src + (long long)(2 * (pg * CW + c)) * p.r_ys + n2;
#pragma unroll
            for (int n1 = 0; n1 < R1z; ++n1)
                dst[n1] = valid ? Cx<real>{__ldcs(ra + n1 * R2z), __ldcs(ra + p.r_ys + n1 * R2z)} : Cx<real>{0, 0};
        };
        if (ONE_ROW_TASK) load_pair(warp, lane, nx);
        for (int pg = warp; pg < NPG; pg += NW) {
            for (int task = lane; task < CW * R2z; task += 32) {
                const int c = task / R2z, n2 = task % R2z;
                Cx<real> v[R1z];
                if (ONE_ROW_TASK) {
#pragma unroll
                    for (int n1 = 0; n1 < R1z; ++n1) v[n1] = nx[n1];
                    load_pair(pg + NW, lane, nx);
                } else {
                    load_pair(pg, task, v);
                }
                dft_reg<real, R1z, -1>(v);
                if constexpr (REGR) {
                    twr.template apply<false>(v);
#pragma unroll
                    for (int k1 = 0; k1 < R1z; ++k1) wtile[LB::at2(k1, n2, c)] = v[k1];
                } else {
#pragma unroll
                    for (int k1 = 0; k1 < R1z; ++k1) {
                        const Cx<real> w = twz[(n2 * k1) & (NZ - 1)];
                        wtile[LB::at2(k1, n2, c)] = (k1 == 0) ? v[k1] : cmul(w, v[k1]);
                    }
                }
            }
            __syncwarp();
            {   // exactly 32 tasks: (row pair c, k1); v[k2] = Z[k1 + R1z k2]
                const int c = lane / R1z, k1 = lane % R1z;
                Cx<real> v[R2z];
#pragma unroll
                for (int n2 = 0; n2 < R2z; ++n2) v[n2] = wtile[LB::at2(k1, n2, c)];
                dft_reg<real, R2z, -1>(v);
                // Z[NZ - k] lives in lane (R1z - k1) % R1z of the same row, register R2z-1-k2
                // (k1 == 0: own register (R2z - k2) % R2z)
                const int partner = lane - k1 + ((R1z - k1) % R1z);
                Cx<real>* rowA = scr + (long long)(2 * (pg * CW + c)) * NZCP;
                Cx<real>* rowB = rowA + NZCP;
                static_for<0, R2z / 2 + 1>([&](auto kc) {
                    constexpr int k2 = decltype(kc)::value;
                    const Cx<real> give = (k1 == 0) ? v[(R2z - k2) % R2z] : v[R2z - 1 - k2];
                    const Cx<real> zn = {shfl(give.x, partner), shfl(give.y, partner)};
                    const Cx<real> zk = v[k2];
                    const int k = k1 + R1z * k2;
                    if (2 * k <= NZ) {
                        st_l2(rowA + k, Cx<real>{(real)0.5 * (zk.x + zn.x), (real)0.5 * (zk.y - zn.y)});
                        st_l2(rowB + k, Cx<real>{(real)0.5 * (zk.y + zn.y), (real)0.5 * (zn.x - zk.x)});
                    }
                });
                if (NZCP > NZC && k1 == 0) {
                    st_l2(rowA + NZC, Cx<real>{0, 0});
                    st_l2(rowB + NZC, Cx<real>{0, 0});
                }
            }
            __syncwarp();
        }
        __syncthreads();
        // ---------------- column phase: FFT along y from the scratch plane ----------------
        Cx<real>* dstp = out + f * p.k_fs + x * p.k_xs;
        int it = 0;
        for (int ch = (g + unit) % NG; ch < NCHG; ch += NG, ++it) {
            Cx<real>* cur = gtile + (it & (NTILE - 1)) * LA::ELEMS;
            const int c0 = ch * CG;
            for (int task = gt; task < CG * R2y; task += GT) {
                const int c = task % CG, n2 = task / CG;
                const bool valid = c0 + c < NZCP;
                Cx<real> v[R1y];
#pragma unroll
                for (int n1 = 0; n1 < R1y; ++n1)
                    v[n1] = valid ? ld_l2(scr + (long long)(n1 * R2y + n2) * NZCP + c0 + c) : Cx<real>{0, 0};
                dft_reg<real, R1y, -1>(v);
                if constexpr (REGC) {
                    twc.template apply<false>(v);
#pragma unroll
                    for (int k1 = 0; k1 < R1y; ++k1) cur[LA::at2(k1, n2, c)] = v[k1];
                } else {
#pragma unroll
                    for (int k1 = 0; k1 < R1y; ++k1) {
                        const Cx<real> w = twy[(n2 * k1) & (NY - 1)];
                        cur[LA::at2(k1, n2, c)] = (k1 == 0) ? v[k1] : cmul(w, v[k1]);
                    }
                }
            }
            group_sync(g + 1, GT);
            for (int task = gt; task < CG * R1y; task += GT) {
                const int c = task % CG, k1 = task / CG;
                Cx<real> v[R2y];
#pragma unroll
                for (int n2 = 0; n2 < R2y; ++n2) v[n2] = cur[LA::at2(k1, n2, c)];
                dft_reg<real, R2y, -1>(v);
                if (c0 + c < NZCP) {
                    if (p.push) {
                        const long long off = (long long)(p.x0 + x) * p.pk_xs + f * p.pk_fs + c0 + c;
                        const int nyl_mask = (1 << p.nyl_shift) - 1;
#pragma unroll
                        for (int k2 = 0; k2 < R2y; ++k2) {
                            const int ky = k1 + R1y * k2;
                            Cx<real>* dq = reinterpret_cast<Cx<real>*>(p.peer[ky >> p.nyl_shift]);
                            st_stream(dq + off + (long long)(ky & nyl_mask) * NZCP, v[k2]);
                        }
                    } else {
#pragma unroll
                        for (int k2 = 0; k2 < R2y; ++k2)
                            st_stream(dstp + (long long)(k1 + R1y * k2) * NZCP + c0 + c, v[k2]);
                    }
                }
            }
            if (NTILE == 1) group_sync(g + 1, GT);
        }
        __syncthreads();
    }
}

// =====================================================================================================
// Tensor-memory variant (fp32, 256 x 256 planes: the benchmark mesh): the y <-> z exchange between the two
// phases goes through TMEM instead of an L2-resident scratch plane.
//
// A plane of half spectra is 256 rows x 129 columns of complex fp32.  Columns kz = 0 .. 127 are
// 256 * 128 * 8 B = 256 KB -- exactly the tensor memory of one SM (128 lanes x 512 columns x 32 bit), which
// nothing else in this library uses; the Nyquist column kz = 128 (2 KB) lives in shared memory.  tcgen05.st /
// tcgen05.ld (shape 32x32b) let thread t of warp w touch TMEM lane 32 (w % 4) + t only, so the two phases
// are laid out such that the thread that WRITES element (y, kz) in one phase and the thread that READS it in
// the other sit in the same lane of warps with the same w % 4:
//
//      y = y7 .. y0,  kz = kz6 .. kz0            lane   = (y1, kz3 kz2 kz1 kz0)        (thread lane)
//                                                quarter = (y3, y2)                     (warp % 4)
//                                                column = 2 * (16 * (y7 y6 y5 y4) + (kz6 kz5 kz4 y0)) + {re, im}
//
//   column phase (FFT along y, 32 columns at a time over the whole CTA): the thread of the last butterfly
//       stage owns (column c = kz4..kz0, n2 = y3..y0) and all sixteen n1 = y7..y4: sixteen 2-column
//       accesses;
//   row phase (FFT along z on row pairs): the thread owns k1 = kz3..kz0 and two rows y0 = 0, 1 of one row
//       pair, for k2 = kz6 kz5 kz4 = 0..7: ONE 32-column access; the mirrored half k > 128 comes from the
//       partner lane 16 - k1 of the same warp by shuffles.
//
// No scratch plane, no L2 round trip (it was two thirds of the kernel's global load/store instructions
// and half of its L2 traffic; ncu: lg_throttle was the top stall), and the global accesses of the column
// phase become 256-byte rows.  Everything else -- butterflies, register twiddles, the derive mode, the
// ghost-padded output, the fused forward transpose -- is the code of the kernels above.
// =====================================================================================================
constexpr int TM_N = 256, TM_CS = 32;                      // plane edge, columns per super-chunk
using TmLA = LayA<TM_N, TM_CS>;
using TmLB = LayB<TM_N, 2>;
constexpr int TM_TILE = TmLA::ELEMS > 16 * TmLB::ELEMS ? TmLA::ELEMS : 16 * TmLB::ELEMS;
constexpr size_t TM_SMEM = sizeof(Cx<float>) * (size_t)(4 * TM_N + TM_TILE);
constexpr size_t TM_SMEM_DBUF = sizeof(Cx<float>) * (size_t)(4 * TM_N + 2 * TmLA::ELEMS);

__device__ __forceinline__ void tmem_st2(uint32_t taddr, float a, float b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(__float_as_uint(a)),
                 "r"(__float_as_uint(b)) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float& a, float& b) {
    uint32_t x, y;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(taddr) : "memory");
    a = __uint_as_float(x); b = __uint_as_float(y);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]),
        "f"(r[8]), "f"(r[9]), "f"(r[10]), "f"(r[11]), "f"(r[12]), "f"(r[13]), "f"(r[14]), "f"(r[15]), "f"(r[16]),
        "f"(r[17]), "f"(r[18]), "f"(r[19]), "f"(r[20]), "f"(r[21]), "f"(r[22]), "f"(r[23]), "f"(r[24]), "f"(r[25]),
        "f"(r[26]), "f"(r[27]), "f"(r[28]), "f"(r[29]), "f"(r[30]), "f"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
          "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]),
          "=f"(r[16]), "=f"(r[17]), "=f"(r[18]), "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]),
          "=f"(r[24]), "=f"(r[25]), "=f"(r[26]), "=f"(r[27]), "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// CTA barrier that orders tensor-memory accesses before / after it
__device__ __forceinline__ void tmem_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ uint32_t tmem_alloc_all(uint32_t* slot, int warp) {
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tmem_sync();
    return *reinterpret_cast<volatile uint32_t*>(slot);
}
__device__ __forceinline__ void tmem_free_all(uint32_t base, int warp) {
    tmem_sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

// thread <-> task maps shared by both kernels (512 threads, warp w = 0..15, lane l)
struct TmMap {
    int c, n2, combo_lo;        // column-phase TMEM side: column within the super-chunk, y % 16, (kz4, y0)
    int k1, y1, q, n1lo;        // row-phase: kz % 16, y1, (y3 y2), (y5 y4)
    uint32_t lane_base;         // TMEM lane field of this warp
    __device__ __forceinline__ TmMap(int w, int l) {
        c = ((w >> 2) & 1) * 16 + (l & 15);
        n2 = ((w & 3) << 2) | ((l >> 4) << 1) | (w >> 3);
        combo_lo = ((w >> 2) & 1) * 2 + (w >> 3);
        k1 = l & 15; y1 = l >> 4; q = w & 3; n1lo = w >> 2;
        lane_base = (uint32_t)(32 * (w & 3)) << 16;
    }
};

// BSH: rows k1 + 16 k2 (k1 < 16) of the input lie in block k2 >> BSH of the blocked layout (Ny / P = 16 << BSH rows
// per block); BSH = 4: the plain layout (one block).
template <int BSH>
__global__ void __launch_bounds__(512, 1) plane_c2r_tmem_kernel(
    const Cx<float>* __restrict__ in, float* __restrict__ out, const Cx<float>* __restrict__ twy_g,
    const Cx<float>* __restrict__ twz_g, PlaneParams p) {
    using real = float;
    constexpr int NY = TM_N, NZ = TM_N, NZC = NZ / 2 + 1, NZCP = NZC + 1, R = 16, CS = TM_CS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cx<real>* twy = reinterpret_cast<Cx<real>*>(smem_raw);
    Cx<real>* twz = twy + NY;
    Cx<real>* nyq = twz + NZ;                               // column kz = 128 after the y transform (x2: derive)
    Cx<real>* tile = nyq + 2 * NY;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const bool dbuf = p.scr_alt != 0;                       // two column tiles: one CTA barrier per round
    for (int i = tid; i < NY; i += 512) { twy[i] = twy_g[i]; twz[i] = twz_g[i]; }
    const uint32_t tbase = tmem_alloc_all(&tmem_slot, warp);
    const TmMap m(warp, lane);
    TwiddleRegs<real, R> twc, twr;
    const int a_c = lane, a_k1 = warp;                      // first column stage: (column, k1)
    twc.init(twy, a_k1, NY);
    twr.init(twz, m.k1, NZ);
    Cx<real>* wtile = tile + warp * TmLB::ELEMS;
    const int ND = p.derive ? 3 : 1;
    const real dky = (real)p.dky, dkz = (real)p.dkz;

    for (int unit = blockIdx.x; unit < p.nunits; unit += gridDim.x) {
      const int fu = unit / p.nplanes + p.f0, x = unit % p.nplanes;
      for (int d = 0; d < ND; ++d) {
        const int fin = p.derive ? 2 * fu + (d > 0 ? 1 : 0) : fu;
        const int f = p.derive ? 3 * fu + d : fu;
        const int mode = p.derive ? d : 0;
        const Cx<real>* src = in + fin * p.k_fs + x * p.k_xs;
        auto scale_chunk = [&](Cx<real> (&v)[R], int k1, int col) {
            if (mode == 1) {
                const bool selfc = col == 0 || 2 * col == NZ;
                const real fk1 = (real)k1;
#pragma unroll
                for (int k2 = 0; k2 < R; ++k2) {
                    const real nn = fk1 + (real)(R * k2 - (2 * k2 >= R ? NY : 0));
                    real sk = dky * nn;
                    if (2 * k2 == R && k1 == 0 && selfc) sk = 0;
                    v[k2].x *= sk; v[k2].y *= sk;
                }
            } else if (mode == 2) {
                const real sk = (2 * col == NZ) ? (real)0 : dkz * (real)col;
#pragma unroll
                for (int k2 = 0; k2 < R; ++k2) { v[k2].x *= sk; v[k2].y *= sk; }
            }
        };
        // ---------------- column phase: inverse FFT along y, super-chunks 0..3 = columns 0..127, 4 = Nyquist ----
        Cx<real> nx[R];
        // row a_k1, column a_c of the input plane; rows a_k1 + 16 k2 and the column blocks follow by immediates
        const Cx<real>* cbase = in_row<real>(in, p, fin, x, a_k1, NZCP) + a_c;
        auto load_chunk = [&](int sc, Cx<real> (&dst)[R]) {
            const bool valid = sc < 4 || (sc == 4 && a_c == 0);
#pragma unroll
            for (int k2 = 0; k2 < R; ++k2) {
                const Cx<real>* e = cbase + sc * CS + (k2 * (R * NZCP) + (long long)(k2 >> BSH) * p.blk_d);
                dst[k2] = valid ? (mode == 1 ? ld_l2(e) : ld_stream(e)) : Cx<real>{0, 0};
            }
        };
        // Nyquist column (round 4).  derive: mode 2 multiplies it by k_z = 0 (all zeros: no round), and the
        // columns of modes 0 and 1 are transformed together in the round of d = 0 (lanes c = 0 and c = 1).
        const int nround = (p.derive && d > 0) ? 4 : 5;
        const int nyq_cols = p.derive ? 2 : 1;
        auto load_nyq = [&](Cx<real> (&dst)[R]) {
            const int fn = p.derive ? 2 * fu + (a_c < nyq_cols ? a_c : 0) : fin;
            const Cx<real>* nbase = in_row<real>(in, p, fn, x, a_k1, NZCP) + NZ / 2;
#pragma unroll
            for (int k2 = 0; k2 < R; ++k2)
                dst[k2] = a_c < nyq_cols ? ld_l2(nbase + (k2 * (R * NZCP) + (long long)(k2 >> BSH) * p.blk_d)) : Cx<real>{0, 0};
        };
        load_chunk(0, nx);
        // rounds 0..3 (compile-time round index: the prefetch target, the TMEM columns and the tile buffer fold)
        static_for<0, 4>([&](auto scc) {
            constexpr int sc = decltype(scc)::value;
            Cx<real>* cur = tile + (dbuf ? (sc & 1) * TmLA::ELEMS : 0);
            Cx<real> v[R];
#pragma unroll
            for (int k2 = 0; k2 < R; ++k2) v[k2] = nx[k2];
            if constexpr (sc < 3) load_chunk(sc + 1, nx);
            else if (nround == 5) load_nyq(nx);
            scale_chunk(v, a_k1, sc * CS + a_c);
            dft_reg<real, R, +1>(v);
            twc.template apply<true>(v);
#pragma unroll
            for (int n2 = 0; n2 < R; ++n2) cur[TmLA::at2(a_k1, n2, a_c)] = v[n2];
            __syncthreads();
            Cx<real> u[R];
#pragma unroll
            for (int k1 = 0; k1 < R; ++k1) u[k1] = cur[TmLA::at2(k1, m.n2, m.c)];
            dft_reg<real, R, +1>(u);
            const uint32_t t0 = tbase + m.lane_base + 2u * (uint32_t)(sc * 4 + m.combo_lo);
#pragma unroll
            for (int n1 = 0; n1 < R; ++n1) tmem_st2(t0 + 32u * n1, u[n1].x, u[n1].y);
            if (!dbuf) __syncthreads();
        });
        if (nround == 5) {                                  // Nyquist column(s): 16 (32) threads per stage
            Cx<real>* cur = tile;
            if (a_c < nyq_cols) {
                Cx<real> v[R];
#pragma unroll
                for (int k2 = 0; k2 < R; ++k2) v[k2] = nx[k2];
                if (p.derive && a_c == 1) {                 // mode 1 on the Nyquist column: k_y, self-conjugate rule
                    const real fk1 = (real)a_k1;
#pragma unroll
                    for (int k2 = 0; k2 < R; ++k2) {
                        real sk = dky * (fk1 + (real)(R * k2 - (2 * k2 >= R ? NY : 0)));
                        if (2 * k2 == R && a_k1 == 0) sk = 0;
                        v[k2].x *= sk; v[k2].y *= sk;
                    }
                }
                dft_reg<real, R, +1>(v);
                twc.template apply<true>(v);
#pragma unroll
                for (int n2 = 0; n2 < R; ++n2) cur[TmLA::at2(a_k1, n2, a_c)] = v[n2];
            }
            __syncthreads();
            if (m.c < nyq_cols) {
                Cx<real> u[R];
#pragma unroll
                for (int k1 = 0; k1 < R; ++k1) u[k1] = cur[TmLA::at2(k1, m.n2, m.c)];
                dft_reg<real, R, +1>(u);
                Cx<real>* nq = nyq + m.c * NY;
#pragma unroll
                for (int n1 = 0; n1 < R; ++n1) nq[n1 * R + m.n2] = u[n1];
            }
        }
        tmem_wait_st();
        tmem_sync();
        // ---------------- row phase: c2r along z on row pairs (A + iB) ----------------
        real* obase = out + f * p.r_fs + x * p.r_xs;
        real* obase2 = (p.xdup_plane >= 0 && x == 0) ? out + f * p.r_fs + p.xdup_plane * p.r_xs : nullptr;
        const int partner = (lane & 16) | ((16 - m.k1) & 15);
        for (int pass = 0; pass < 4; ++pass) {
            const int n1 = pass * 4 + m.n1lo;
            const int ybase = n1 * 16 + m.q * 4;            // rows ybase + 2 y1 (+ 1) belong to this lane
            float r[32];
            tmem_ld32(tbase + m.lane_base + 32u * (uint32_t)n1, r);
            tmem_wait_ld();
            Cx<real> v[R];
            // k = k1 + 16 k2 < 128: A = (r[4k2], r[4k2+1]), B = (r[4k2+2], r[4k2+3])
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) {
                const real s = (k2 == 0 && m.k1 == 0) ? (real)0 : (real)1;
                v[k2] = {r[4 * k2] - s * r[4 * k2 + 3], s * r[4 * k2 + 1] + r[4 * k2 + 2]};
            }
            // mirrored half k = k1 + 16 (8 + j): conj(A) + i conj(B) of element 256 - k, held by the partner lane
            // (register 7 - j; lane k1 = 0 mirrors onto itself, register 8 - j, and j = 0 is the Nyquist column)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int own = (j == 0) ? 0 : 8 - j;       // k1 == 0 source register (j = 0 unused)
                const real gx = (m.k1 == 0) ? r[4 * own] + r[4 * own + 3] : r[4 * (7 - j)] + r[4 * (7 - j) + 3];
                const real gy = (m.k1 == 0) ? r[4 * own + 2] - r[4 * own + 1] : r[4 * (7 - j) + 2] - r[4 * (7 - j) + 1];
                v[8 + j] = {shfl(gx, partner), shfl(gy, partner)};
            }
            if (m.k1 == 0) {                                // k = 128: imaginary parts dropped (c2r semantics)
                const int y = ybase + 2 * m.y1;
                if (p.derive && d == 2) v[8] = {0, 0};
                else { const Cx<real>* nq = nyq + ((p.derive && d == 1) ? NY : 0); v[8] = {nq[y].x, nq[y + 1].x}; }
            }
            dft_reg<real, R, +1>(v);
            twr.template apply<true>(v);
#pragma unroll
            for (int n2 = 0; n2 < R; ++n2) wtile[TmLB::at2(m.k1, n2, m.y1)] = v[n2];
            __syncwarp();
            {
                const int c2 = lane / R, n2 = lane % R;
                Cx<real> u[R];
#pragma unroll
                for (int k1b = 0; k1b < R; ++k1b) u[k1b] = wtile[TmLB::at2(k1b, n2, c2)];
                dft_reg<real, R, +1>(u);
                const int y0 = ybase + 2 * c2;
                if (p.vec_out) {
                    // The four finished rows ybase .. ybase + 3 of this warp go through the (now free) warp tile and
                    // leave as 16-byte stores, whole 512-byte runs per instruction, instead of 32 4-byte stores per
                    // lane: the load/store queue was the kernel's top stall (ncu: lg_throttle).  Row pitch 264 words:
                    // the two row pairs of a pass fall into different banks.
                    constexpr int OP = NZ + 8;
                    __syncwarp();                            // every lane has read its butterfly inputs
                    float* st = reinterpret_cast<float*>(wtile);
#pragma unroll
                    for (int n1b = 0; n1b < R; ++n1b) {
                        st[(2 * c2) * OP + n1b * R + n2] = u[n1b].x;
                        st[(2 * c2 + 1) * OP + n1b * R + n2] = u[n1b].y;
                    }
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float4 a = *reinterpret_cast<const float4*>(st + r * OP + 4 * lane);
                        const float4 b = *reinterpret_cast<const float4*>(st + r * OP + NZ / 2 + 4 * lane);
#pragma unroll
                        for (int dup = 0; dup < 2; ++dup) {
                            real* ob = dup ? obase2 : obase;
                            if (ob == nullptr) continue;
#pragma unroll
                            for (int gh = 0; gh < 2; ++gh) {     // gh = 1: row 0 again as the periodic image y = NY
                                if (gh && !(p.ghost && ybase + r == 0)) continue;
                                real* row = ob + (long long)(gh ? NY : ybase + r) * p.r_ys;
                                __stcs(reinterpret_cast<float4*>(row + 4 * lane), a);
                                __stcs(reinterpret_cast<float4*>(row + NZ / 2 + 4 * lane), b);
                                if (p.ghost && lane == 0) row[NZ] = a.x;
                            }
                        }
                    }
                } else
#pragma unroll
                for (int dup = 0; dup < 2; ++dup) {
                    real* ob = dup ? obase2 : obase;
                    if (ob == nullptr) continue;
                    real* oa = ob + (long long)y0 * p.r_ys + n2;
#pragma unroll
                    for (int n1b = 0; n1b < R; ++n1b) {
                        __stcs(oa + n1b * R, u[n1b].x);
                        __stcs(oa + p.r_ys + n1b * R, u[n1b].y);
                    }
                    if (p.ghost) {
                        if (n2 == 0) { oa[NZ] = u[0].x; oa[p.r_ys + NZ] = u[0].y; }
                        if (y0 == 0) {
                            real* og = ob + (long long)NY * p.r_ys + n2;
#pragma unroll
                            for (int n1b = 0; n1b < R; ++n1b) __stcs(og + n1b * R, u[n1b].x);
                            if (n2 == 0) og[NZ] = u[0].x;
                        }
                    }
                }
            }
            __syncwarp();
        }
        tmem_sync();                                        // TMEM and the tile are free for the next transform
      }
    }
    tmem_free_all(tbase, warp);
}

__global__ void __launch_bounds__(512, 1) plane_r2c_tmem_kernel(
    const float* __restrict__ in, Cx<float>* __restrict__ out, const Cx<float>* __restrict__ twy_g,
    const Cx<float>* __restrict__ twz_g, PlaneParams p) {
    using real = float;
    constexpr int NY = TM_N, NZ = TM_N, NZC = NZ / 2 + 1, NZCP = NZC + 1, R = 16, CS = TM_CS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cx<real>* twy = reinterpret_cast<Cx<real>*>(smem_raw);
    Cx<real>* twz = twy + NY;
    Cx<real>* nyq = twz + NZ;
    Cx<real>* tile = nyq + 2 * NY;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const bool dbuf = p.scr_alt != 0;
    for (int i = tid; i < NY; i += 512) { twy[i] = twy_g[i]; twz[i] = twz_g[i]; }
    const uint32_t tbase = tmem_alloc_all(&tmem_slot, warp);
    const TmMap m(warp, lane);
    TwiddleRegs<real, R> twc, twr;
    twr.init(twz, lane % R, NZ);            // row phase, first stage: n2 = lane % 16
    twc.init(twy, m.n2, NY);                // column phase, first stage: n2 = y % 16
    Cx<real>* wtile = tile + warp * TmLB::ELEMS;
    const int b_c = lane, b_k1 = warp;      // last column stage: (column, k1)

    for (int unit = blockIdx.x; unit < p.nunits; unit += gridDim.x) {
        const int f = unit / p.nplanes + p.f0, x = unit % p.nplanes;
        const real* src = in + f * p.r_fs + x * p.r_xs;
        // ---------------- row phase: one complex FFT along z per row pair (a + ib) ----------------
        Cx<real> nx[R];
        auto load_pair = [&](int pass, Cx<real> (&dst)[R]) {
            const int c = lane / R, n2 = lane % R;
            const int y = (pass * 4 + m.n1lo) * 16 + m.q * 4 + 2 * c;
            const real* ra = src + (long long)y * p.r_ys + n2;
#pragma unroll
            for (int n1 = 0; n1 < R; ++n1)
                dst[n1] = pass < 4 ? Cx<real>{__ldcs(ra + n1 * R), __ldcs(ra + p.r_ys + n1 * R)} : Cx<real>{0, 0};
        };
        load_pair(0, nx);
        const int partner = (lane & 16) | ((16 - m.k1) & 15);
        for (int pass = 0; pass < 4; ++pass) {
            const int n1r = pass * 4 + m.n1lo;
            {
                const int c = lane / R, n2 = lane % R;
                Cx<real> v[R];
#pragma unroll
                for (int n1 = 0; n1 < R; ++n1) v[n1] = nx[n1];
                load_pair(pass + 1, nx);
                dft_reg<real, R, -1>(v);
                twr.template apply<false>(v);
#pragma unroll
                for (int k1 = 0; k1 < R; ++k1) wtile[TmLB::at2(k1, n2, c)] = v[k1];
            }
            __syncwarp();
            {   // (row pair y1, k1): v[k2] = Z[k1 + 16 k2]
                Cx<real> v[R];
#pragma unroll
                for (int n2 = 0; n2 < R; ++n2) v[n2] = wtile[TmLB::at2(m.k1, n2, m.y1)];
                dft_reg<real, R, -1>(v);
                float r[32];
                Cx<real> nA = {0, 0}, nB = {0, 0};
                static_for<0, 9>([&](auto kc) {
                    constexpr int k2 = decltype(kc)::value;
                    // Z[NZ - k] lives in lane (16 - k1) % 16 of the same row pair, register 15 - k2
                    // (k1 == 0: own register (16 - k2) % 16)
                    const Cx<real> give = (m.k1 == 0) ? v[(R - k2) % R] : v[R - 1 - k2];
                    const Cx<real> zn = {shfl(give.x, partner), shfl(give.y, partner)};
                    const Cx<real> zk = v[k2];
                    const Cx<real> A = {(real)0.5 * (zk.x + zn.x), (real)0.5 * (zk.y - zn.y)};
                    const Cx<real> B = {(real)0.5 * (zk.y + zn.y), (real)0.5 * (zn.x - zk.x)};
                    if constexpr (k2 < 8) {
                        r[4 * k2] = A.x; r[4 * k2 + 1] = A.y; r[4 * k2 + 2] = B.x; r[4 * k2 + 3] = B.y;
                    } else {
                        nA = A; nB = B;                     // k = 128 for k1 == 0
                    }
                });
                tmem_st32(tbase + m.lane_base + 32u * (uint32_t)n1r, r);
                if (m.k1 == 0) {
                    const int y = n1r * 16 + m.q * 4 + 2 * m.y1;
                    nyq[y] = nA; nyq[y + 1] = nB;
                }
            }
            __syncwarp();
        }
        tmem_wait_st();
        tmem_sync();
        // ---------------- column phase: FFT along y, super-chunks 0..3 = columns 0..127, 4 = Nyquist + pad ----
        Cx<real>* dstp = out + f * p.k_fs + x * p.k_xs;
        for (int sc = 0; sc < 5; ++sc) {
            Cx<real>* cur = tile + (dbuf ? (sc & 1) * TmLA::ELEMS : 0);
            if (sc < 4 || m.c == 0) {
                Cx<real> v[R];
                if (sc < 4) {
                    const uint32_t t0 = tbase + m.lane_base + 2u * (uint32_t)(sc * 4 + m.combo_lo);
#pragma unroll
                    for (int n1 = 0; n1 < R; ++n1) tmem_ld2(t0 + 32u * n1, v[n1].x, v[n1].y);
                    tmem_wait_ld();
                } else {
#pragma unroll
                    for (int n1 = 0; n1 < R; ++n1) v[n1] = nyq[n1 * R + m.n2];
                }
                dft_reg<real, R, -1>(v);
                twc.template apply<false>(v);
#pragma unroll
                for (int k1 = 0; k1 < R; ++k1) cur[TmLA::at2(k1, m.n2, m.c)] = v[k1];
            }
            __syncthreads();
            const int col = sc * CS + b_c;
            if (sc < 4 || b_c < 2) {                        // column 128 and the zero pad column 129
                Cx<real> v[R];
#pragma unroll
                for (int n2 = 0; n2 < R; ++n2) v[n2] = (sc < 4 || b_c == 0) ? cur[TmLA::at2(b_k1, n2, b_c)] : Cx<real>{0, 0};
                if (sc < 4 || b_c == 0) dft_reg<real, R, -1>(v);
                if (p.push) {
                    const long long off = (long long)(p.x0 + x) * p.pk_xs + f * p.pk_fs + col;
                    const int nyl_mask = (1 << p.nyl_shift) - 1;
#pragma unroll
                    for (int k2 = 0; k2 < R; ++k2) {
                        const int ky = b_k1 + R * k2;
                        Cx<real>* dq = reinterpret_cast<Cx<real>*>(p.peer[ky >> p.nyl_shift]);
                        st_stream(dq + off + (long long)(ky & nyl_mask) * NZCP, v[k2]);
                    }
                } else {
#pragma unroll
                    for (int k2 = 0; k2 < R; ++k2) st_stream(dstp + (long long)(b_k1 + R * k2) * NZCP + col, v[k2]);
                }
            }
            if (!dbuf) __syncthreads();
        }
        tmem_sync();
    }
    tmem_free_all(tbase, warp);
}

static bool plane_tmem_enabled() {
    const char* e = getenv("HYMD_B200_PLANE_TMEM");
    return !(e && e[0] == '0');
}

static int launch_plane_tmem(hymd_ctx* c, bool inverse, const void* in, void* out, const PlaneParams& p_in, cudaStream_t s) {
    int sms = 0;
    HYMD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->dev));
    int grid = sms;                     // one CTA per SM: it owns the SM's whole tensor memory
    if (c->plane_sm_reserve > 0 && grid > sms - c->plane_sm_reserve) grid = sms - c->plane_sm_reserve;
    if (grid > p_in.nunits) grid = p_in.nunits;
    if (grid < 1) return HYMD_OK;
    PlaneParams p = p_in;
    // double-buffered column tiles (one CTA barrier per round instead of two): measured -2 % on the forward
    // kernel, nothing on the inverse (profiles/r2h_*), so only the forward kernel pays the extra 70 KB
    p.scr_alt = inverse ? 0 : 1;
    if (const char* e = getenv("HYMD_B200_TMEM_DBUF")) p.scr_alt = atoi(e) != 0;
    p.vec_out = 0;                      // measured slower (2 slabs, C4: 0.489 vs 0.475 ms per cycle in the inverse transform)
    if (const char* e = getenv("HYMD_B200_C2R_VEC")) p.vec_out = atoi(e) != 0;
    if (p.r_ys % 4 != 0) p.vec_out = 0;
    const size_t smem = p.scr_alt ? TM_SMEM_DBUF : TM_SMEM;
    if (inverse) {
        const int bsh = p.blk_in ? p.nyl_shift - 4 : 4;
        if (bsh < 0 || bsh > 4) { set_error("tensor-memory c2r: blocked input needs 16 <= Ny / P"); return HYMD_ERR_INVALID; }
        auto kern = bsh == 0 ? plane_c2r_tmem_kernel<0> : bsh == 1 ? plane_c2r_tmem_kernel<1> : bsh == 2 ? plane_c2r_tmem_kernel<2>
                  : bsh == 3 ? plane_c2r_tmem_kernel<3> : plane_c2r_tmem_kernel<4>;
        HYMD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 512, smem, s>>>((const Cx<float>*)in, (float*)out, (const Cx<float>*)c->ytw,
                                     (const Cx<float>*)c->ztw, p);
    } else {
        HYMD_CUDA(cudaFuncSetAttribute(plane_r2c_tmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        plane_r2c_tmem_kernel<<<grid, 512, smem, s>>>((const float*)in, (Cx<float>*)out, (const Cx<float>*)c->ytw,
                                                         (const Cx<float>*)c->ztw, p);
    }
    HYMD_LAUNCH_CHECK(c);
    return HYMD_OK;
}

// ---- host side ------------------------------------------------------------------------------------
static bool pow2_in(int n, int lo, int hi) { return n >= lo && n <= hi && (n & (n - 1)) == 0; }

bool plane_supported(const hymd_ctx* c) {
    const Geometry& g = c->g;
    if (const char* e = getenv("HYMD_B200_NO_PLANE")) if (e[0] == '1') return false;
    if (g.Ny != g.Nz) return false;                       // square planes are instantiated
    return pow2_in(g.Ny, 16, c->f64 ? 256 : 512);
}

static int plane_tables(hymd_ctx* c) {
    if (c->ytw) return HYMD_OK;
    const Geometry& g = c->g;
    for (int a = 0; a < 2; ++a) {
        const int n = a ? g.Nz : g.Ny;
        std::vector<double> tw(2 * (size_t)n);
        for (int j = 0; j < n; ++j) {
            tw[2 * j] = cos(2.0 * M_PI * j / n);
            tw[2 * j + 1] = -sin(2.0 * M_PI * j / n);
        }
        void* d = nullptr;
        HYMD_CUDA(cudaMalloc(&d, tw.size() * c->rsz));
        if (c->f64) {
            HYMD_CUDA(cudaMemcpy(d, tw.data(), tw.size() * 8, cudaMemcpyHostToDevice));
        } else {
            std::vector<float> h(tw.begin(), tw.end());
            HYMD_CUDA(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
        }
        (a ? c->ztw : c->ytw) = d;
    }
    HYMD_CUDA(cudaDeviceSynchronize());   // uploads ran on the legacy stream, the kernels use the caller's
    return HYMD_OK;
}

// Build variants of the plane kernels (HYMD_B200_PLANE_TILES):
//   1 (default): one column tile per group (two group barriers per chunk), row phase straight
//                from / to global memory; 74 KB of shared memory at 256^2 fp32;
//   2          : double-buffered column tiles (one barrier per chunk) and, in the inverse, row
//                inputs (and with HYMD_B200_ROW_TMA=3 outputs) staged through bulk async copies;
//                143 - 208 KB of shared memory.
// Measured at C4 (12 inverse transforms): 0.623 ms (1) vs 0.656 (2, staged inputs) vs 0.68 - 0.82
// (2, staged outputs).  These kernels speed up with every KB of L1 left to the load/store unit
// (padding the allocation of the same kernel by 64 KB: 0.66 -> 0.85 ms), which outweighs what the
// copy engine saves in load/store instructions.  Two 256-thread CTAs per SM instead of one of 512
// were slower as well (0.84 vs 0.69 ms: the second scratch plane per SM) and are gone.
static int plane_tiles() {
    if (const char* e = getenv("HYMD_B200_PLANE_TILES")) { const int v = atoi(e); return v >= 1 && v <= 3 ? v : 1; }
    return 1;
}

template <typename real, int N, int NTH, int TILES, bool INVERSE>
static int launch_plane(hymd_ctx* c, const void* in, void* out, const PlaneParams& p, cudaStream_t s) {
    using Cfg = PlaneCfg<real, N, N, NTH, TILES>;
    HYMD_CHECK(plane_tables(c));
    int sms = 0;
    HYMD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->dev));
    int grid = sms * Cfg::CTAS;
    if (c->plane_sm_reserve > 0 && sms > c->plane_sm_reserve) grid = (sms - c->plane_sm_reserve) * Cfg::CTAS;
    if (const char* e = getenv("HYMD_B200_PLANE_GRID")) grid = atoi(e) > 0 ? atoi(e) : grid;   // tuning
    if (grid > p.nunits) grid = p.nunits;
    if (grid < 1) return HYMD_OK;
    const size_t need = (size_t)sms * 2 * N * Cfg::NZCP * sizeof(Cx<real>);
    if (c->plane_scratch_bytes < need) {
        if (c->plane_scratch) { HYMD_CUDA(cudaDeviceSynchronize()); cudaFree(c->plane_scratch); c->plane_scratch = nullptr; }
        HYMD_CUDA(cudaMalloc(&c->plane_scratch, need));
        c->plane_scratch_bytes = need;
    }
    if ((size_t)grid * N * Cfg::NZCP * sizeof(Cx<real>) > c->plane_scratch_bytes) grid = 2 * sms;
    if (INVERSE) {
        auto kern = p.blk_in ? plane_c2r_kernel<real, N, N, NTH, TILES, true> : plane_c2r_kernel<real, N, N, NTH, TILES, false>;
        size_t smem = Cfg::smem_inv((p.row_tma & 2) != 0);
        if (const char* e = getenv("HYMD_B200_PLANE_SMEM_PAD")) smem += (size_t)atoi(e) * 1024;   // L1 carve-out experiment
        HYMD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, Cfg::NT, smem, s>>>((const Cx<real>*)in, (Cx<real>*)c->plane_scratch, (real*)out,
                                             (const Cx<real>*)c->ytw, (const Cx<real>*)c->ztw, p);
    } else {
        auto kern = plane_r2c_kernel<real, N, N, NTH, TILES>;
        size_t smem = Cfg::SMEM;
        if (const char* e = getenv("HYMD_B200_PLANE_SMEM_PAD")) smem += (size_t)atoi(e) * 1024;
        HYMD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, Cfg::NT, smem, s>>>((const real*)in, (Cx<real>*)c->plane_scratch, (Cx<real>*)out,
                                             (const Cx<real>*)c->ytw, (const Cx<real>*)c->ztw, p);
    }
    HYMD_LAUNCH_CHECK(c);
    return HYMD_OK;
}

// 64^2 / 128^2 fp32 planes: two 256-thread CTAs per SM instead of one of 512 when the launch has work for four waves
// of SMs (>= 4 x SMs (field, plane) units).  Measured at C3 (1 M particles, 128^3, T = 6; profiles/r3p_*): forward
// 0.052 -> 0.046 ms, inverse 0.195 -> 0.164 ms per cycle; the PME transforms of the same system (128 / 384 units) are
// slower that way (0.052 -> 0.060 ms) and keep the 512-thread CTA.  HYMD_B200_PLANE_THREADS = 256 / 512 forces either.
static bool plane_half_ctas(const hymd_ctx* c, int nunits) {
    if (const char* e = getenv("HYMD_B200_PLANE_THREADS")) return atoi(e) == 256;
    const int sms = c->sm_count > 0 ? c->sm_count : 148;
    return nunits >= 4 * sms;
}

template <typename real, int N, bool INVERSE>
static int launch_plane_nt(hymd_ctx* c, const void* in, void* out, const PlaneParams& p, cudaStream_t s) {
    if constexpr (sizeof(real) == 4 && (N == 64 || N == 128)) {
        if (plane_tiles() == 1 && plane_half_ctas(c, p.nunits)) return launch_plane<real, N, 256, 1, INVERSE>(c, in, out, p, s);
    }
    if constexpr (sizeof(real) == 4 && N == 512) {      // experiment (C5): only when forced
        const char* e = getenv("HYMD_B200_PLANE_THREADS");
        if (plane_tiles() == 1 && e && atoi(e) == 256) return launch_plane<real, N, 256, 1, INVERSE>(c, in, out, p, s);
    }
    switch (plane_tiles()) {
        case 2: return launch_plane<real, N, 512, 2, INVERSE>(c, in, out, p, s);
        case 3: return launch_plane<real, N, 512, 3, INVERSE>(c, in, out, p, s);
        default: return launch_plane<real, N, 512, 1, INVERSE>(c, in, out, p, s);
    }
}

template <typename real, bool INVERSE>
static int dispatch_plane(hymd_ctx* c, const void* in, void* out, const PlaneParams& p, cudaStream_t s) {
    if (sizeof(real) == 4 && c->g.Ny == TM_N && c->g.Nz == TM_N && plane_tmem_enabled()) {
        HYMD_CHECK(plane_tables(c));
        return launch_plane_tmem(c, INVERSE, in, out, p, s);
    }
    switch (c->g.Ny) {
        case 16: return launch_plane_nt<real, 16, INVERSE>(c, in, out, p, s);
        case 32: return launch_plane_nt<real, 32, INVERSE>(c, in, out, p, s);
        case 64: return launch_plane_nt<real, 64, INVERSE>(c, in, out, p, s);
        case 128: return launch_plane_nt<real, 128, INVERSE>(c, in, out, p, s);
        case 256: return launch_plane_nt<real, 256, INVERSE>(c, in, out, p, s);
        default: break;
    }
    if (sizeof(real) == 4 && c->g.Ny == 512) return launch_plane_nt<float, 512, INVERSE>(c, in, out, p, s);
    set_error("plane transform: unsupported plane %d x %d", c->g.Ny, c->g.Nz);
    return HYMD_ERR_INVALID;
}

// real [f][plane][Ny][Nz] (strides r_*) -> spectra [f][plane][Ny][Nzcp] (strides k_*)
int plane_forward(hymd_ctx* c, const void* real_in, long long r_fs, int F, int nplanes, void* k_out,
                  long long k_fs, cudaStream_t s, void* const* push_peers, int f0, int nf) {
    const Geometry& g = c->g;
    PlaneParams p;
    memset(&p, 0, sizeof(p));
    if (push_peers) {      // k_out is ignored: rows go to the peers' k buffers (k layout of F fields)
        int sh = 0;
        while ((1 << sh) < g.nyl) ++sh;
        if ((1 << sh) != g.nyl) { set_error("fused forward transpose needs a power-of-two Ny / P"); return HYMD_ERR_INVALID; }
        const KLayout l = klayout(c, F);
        p.push = 1; p.nyl_shift = sh; p.x0 = g.x0; p.pk_xs = l.xs; p.pk_fs = l.fs;
        for (int q = 0; q < g.P; ++q) p.peer[q] = push_peers[q];
    }
    if (nf < 0) nf = F - f0;          // fields f0 .. f0 + nf - 1 of the F-field layouts
    p.nunits = nf * nplanes; p.nplanes = nplanes; p.f0 = f0;
    p.k_fs = k_fs; p.k_xs = (long long)g.Ny * g.Nzcp;
    p.r_fs = r_fs; p.r_xs = (long long)g.Ny * g.Nz; p.r_ys = g.Nz;
    p.ghost = 0; p.xdup_plane = -1; p.derive = 0; p.dky = p.dkz = 0; p.row_tma = 0; p.scr_alt = 0;
    p.scr_hint = 1;
    if (const char* e = getenv("HYMD_B200_SCR_HINT")) p.scr_hint = atoi(e) != 0;
    return c->f64 ? dispatch_plane<double, false>(c, real_in, k_out, p, s)
                  : dispatch_plane<float, false>(c, real_in, k_out, p, s);
}

// spectra [f][plane][Ny][Nzcp] -> F real planes; ghost: the ghost-padded force-mesh layout with the
// periodic y/z images (and plane 0 duplicated into plane nxl on a single GPU).
// derive: k_in holds 2F/3 spectra (per potential row: -i k_x V and -i V, x-inverted) and the
// kernel forms the k_y / k_z components itself (PlaneParams::derive).
int plane_inverse(hymd_ctx* c, const void* k_in, long long k_fs, int F, int nplanes, void* real_out,
                  bool ghost, bool derive, cudaStream_t s, bool blocked, int f0, int nf) {
    const Geometry& g = c->g;
    PlaneParams p;
    memset(&p, 0, sizeof(p));
    if (blocked) {         // k_in = W[q][x][f][kyl][kz] with Fin spectra (see PlaneParams::blk_in)
        const int Fin = derive ? F / 3 * 2 : F;
        int sh = 0;
        while ((1 << sh) < g.nyl) ++sh;
        if ((1 << sh) != g.nyl) { set_error("blocked inverse transpose needs a power-of-two Ny / P"); return HYMD_ERR_INVALID; }
        p.blk_in = 1; p.nyl_shift = sh;
        p.blk_f = (long long)g.nyl * g.Nzcp; p.blk_x = Fin * p.blk_f; p.blk_q = g.nxl * p.blk_x;
        p.blk_d = p.blk_q - (long long)g.nyl * g.Nzcp;
    }
    if (derive && F % 3 != 0) { set_error("plane_inverse: derive needs 3 outputs per row"); return HYMD_ERR_INVALID; }
    if (nf < 0) nf = (derive ? F / 3 : F) - f0;      // fields (derive: potential rows) f0 .. f0 + nf - 1
    p.nunits = nf * nplanes; p.nplanes = nplanes; p.f0 = f0;
    p.derive = derive ? 1 : 0;
    p.dky = 2.0 * M_PI / g.box[1]; p.dkz = 2.0 * M_PI / g.box[2];
    p.row_tma = 1;
    if (const char* e = getenv("HYMD_B200_ROW_TMA")) p.row_tma = atoi(e) & 3;
    p.scr_alt = 0;
    if (const char* e = getenv("HYMD_B200_SCR_ALT")) p.scr_alt = atoi(e) != 0;
    p.scr_hint = 1;
    if (const char* e = getenv("HYMD_B200_SCR_HINT")) p.scr_hint = atoi(e) != 0;
    p.k_fs = k_fs; p.k_xs = (long long)g.Ny * g.Nzcp;
    if (ghost) {
        p.r_ys = g.Nzp; p.r_xs = (long long)(g.Ny + 1) * g.Nzp; p.r_fs = g.ghost_elems;
        p.ghost = 1; p.xdup_plane = g.P == 1 ? g.nxl : -1;
    } else {
        p.r_ys = g.Nz; p.r_xs = (long long)g.Ny * g.Nz; p.r_fs = g.real_elems;
        p.ghost = 0; p.xdup_plane = -1;
    }
    return c->f64 ? dispatch_plane<double, true>(c, k_in, real_out, p, s)
                  : dispatch_plane<float, true>(c, k_in, real_out, p, s);
}

}  // namespace hymd
