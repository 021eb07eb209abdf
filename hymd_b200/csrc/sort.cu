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
    int c;
    if (f >= 0.0 && f < (double)n) {       // the usual case: callers wrap positions into [0, L) (main.py:187, 837)
        c = (int)f;
    } else {                               // images further out: 64-bit modulo (some 60 instructions), rarely taken
        long long cc = (long long)f % n;
        if (cc < 0) cc += n;
        c = (int)cc;
    }
    if (d >= 1.0) {  // x = -tiny rounds to d == 1
        d = 0.0;
        c = (c + 1 == n) ? 0 : c + 1;
    }
    cell = c;
    frac = d;
}

template <typename UT>
__device__ __forceinline__ UT pack_coord(int cell, double frac, int fb) {
    UT f = (UT)(frac * (double)((UT)1 << fb));
    if (f >> fb) f = ((UT)1 << fb) - 1;  // frac*2^fb rounded up to 2^fb
    return ((UT)cell << fb) | f;
}

// Peer addresses of the guest buffers (per-step routing, below) and what the binning pass needs to send a
// home particle that is away to the rank owning its cell.
struct RoutePeers {
    void* pos[HYMD_MAX_PEERS];
    void* type[HYMD_MAX_PEERS];
    void* q[HYMD_MAX_PEERS];
    void* ret[HYMD_MAX_PEERS];
};
struct RouteOut {
    RoutePeers peers;
    uint32_t* send_count;      // [P] guests sent to each rank this step
    int32_t* sent_idx;         // [P*G] home index of guest k sent to rank d at d*G + k
    unsigned int* status;      // mapped host words: bit 1 = capacity exceeded
    const void* q;             // caller-order charges or NULL
    long long G;
    int rank;
};

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
//
// ROUTED (several slabs, route.cu): the staged array of the previous call also holds that call's guests
// (records with idx >= n, dropped here: this step's guests are binned by guest_count_kernel) and, in an
// extra bin `ncell` at its end, the home particles that were away; REUSE then walks all
// rt->n_total records.  A home particle outside this rank's slab goes to the away bin -- it is painted
// and read out on the rank that owns its cell.  keys[j] keeps the bin for the scatter pass.
template <typename real, typename RecT, typename UT, int IDX_BITS, bool REUSE, bool ROUTED, int PER>
__global__ void __launch_bounds__(256, ROUTED ? 5 : 8) count_kernel(const real* __restrict__ pos,
                                                    const int32_t* __restrict__ types, long long n,
                                                    SortParams p, RecT* __restrict__ stage,
                                                    uint32_t* __restrict__ cnt,
                                                    DeviceScalars* __restrict__ sc,
                                                    const RouteTotals* __restrict__ rt,
                                                    uint32_t* __restrict__ keys, uint32_t away_bin, RouteOut ro) {
    const long long limit = (ROUTED && REUSE) ? (long long)rt->n_total : n;
    // A block takes PER x 256 consecutive records per trip (thread t: records t, t + 256, ...), with the loads of
    // all of them issued before the first is processed (the pass is a chain of dependent memory operations).
    // ROUTED: the grid covers the n home particles; the (few) blocks whose range is followed by staged guest
    // records of the previous step take a second trip (block-uniform loop: the warp shuffles stay whole)
    const long long trip = (long long)gridDim.x * (256 * PER);
    for (long long base = blockIdx.x * (long long)(256 * PER);; base += trip) {
    UT idx[PER], type[PER];
    bool live[PER], inr[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const long long j = base + 256 * u + threadIdx.x;
        inr[u] = j < limit;
        live[u] = inr[u];
        idx[u] = 0; type[u] = 0;
        if (live[u]) {
            if (REUSE) {
                const UT meta = stage[j].meta;
                idx[u] = meta & (((UT)1 << IDX_BITS) - 1);
                type[u] = meta >> IDX_BITS;
                if (ROUTED && (long long)idx[u] >= n) live[u] = false;      // a guest of the previous step
            } else {
                idx[u] = (UT)j;
                type[u] = (UT)(uint32_t)types[j];
            }
        }
    }
    real x[PER][3];
#pragma unroll
    for (int u = 0; u < PER; ++u)
#pragma unroll
        for (int a = 0; a < 3; ++a) x[u][a] = live[u] ? pos[3 * idx[u] + a] : (real)0;
    unsigned int r1 = 0, badsum = 0;
#pragma unroll
    for (int u = 0; u < PER; ++u) {
    const long long j = base + 256 * u + threadIdx.x;
    unsigned int bad = 0;
    uint32_t key = 0xffffffffu;                        // lanes past the end: a run of their own
    if (ROUTED && inr[u] && !live[u]) keys[j] = 0xffffffffu;
    if (live[u]) {
        int cx, cy, cz;
        double dx, dy, dz;
        split_coord((double)x[u][0] * p.sx, p.Nx, cx, dx);
        split_coord((double)x[u][1] * p.sy, p.Ny, cy, dy);
        split_coord((double)x[u][2] * p.sz, p.Nz, cz, dz);
        int lx = cx - p.x0;
        if (lx < 0 || lx >= p.nxl) {
            bad = 1;
            lx = lx < 0 ? 0 : p.nxl - 1;
        }
        RecT r;
        r.ux = pack_coord<UT>(lx, dx, p.fbx);
        r.uy = pack_coord<UT>(cy, dy, p.fby);
        r.uz = pack_coord<UT>(cz, dz, p.fbz);
        r.meta = idx[u] | (type[u] << IDX_BITS);
        stage[j] = r;
        key = (uint32_t)(((long long)lx * p.Ny + cy) * p.nbz + zbin_of(cz, p.Nz));
        if (ROUTED) {
            if (bad) {
                // away: goes into rank d's inbox section for this rank (stores into d's HBM over NVLink);
                // here it only occupies the extra bin so that the next REUSE pass finds it again
                key = away_bin;
                const int d = cx / p.nxl;
                const uint32_t k = atomicAdd(&ro.send_count[d], 1u);
                if (k >= (uint32_t)ro.G) {         // never silently: raised to the host through mapped memory
                    if (ro.status) { atomicOr(ro.status, 2u); atomicAdd(ro.status + 1, 1u); }
                } else {
                    const long long row = (long long)ro.rank * ro.G + k;
                    real* dp = reinterpret_cast<real*>(ro.peers.pos[d]) + 3 * row;
                    dp[0] = x[u][0]; dp[1] = x[u][1]; dp[2] = x[u][2];
                    reinterpret_cast<int32_t*>(ro.peers.type[d])[row] = (int32_t)type[u];
                    if (ro.q != nullptr)
                        reinterpret_cast<real*>(ro.peers.q[d])[row] = reinterpret_cast<const real*>(ro.q)[idx[u]];
                    ro.sent_idx[(long long)d * ro.G + k] = (int32_t)idx[u];
                }
            }
            keys[j] = key;
        }
    }
    unsigned head_lane, rank, count;
    warp_runs(key, head_lane, rank, count);
    unsigned int r1u = 0;
    if (rank == 0 && live[u]) r1u = atomicAdd(&cnt[key], count) + count;
    if (ROUTED && key == away_bin) r1u = 0;            // the away bin does not bound the paint scale
    r1 = max(r1, r1u);
    badsum += bad;
    }
    unsigned int m = __reduce_max_sync(0xffffffffu, r1);
    unsigned int b = __reduce_add_sync(0xffffffffu, badsum);
    if ((threadIdx.x & 31) == 0) {
        if (m > sc->max_cell_count) atomicMax(&sc->max_cell_count, m);
        if (b) atomicAdd(&sc->out_of_slab, b);
    }
    if (!ROUTED || base + trip >= limit) break;
    }
}

// Pass 1, one slab: the same as count_kernel<..., false>, two particles per thread (j and j + 256 of a 512-particle
// block).  The pass is a chain of dependent memory operations per particle (previous record -> position -> staged
// record, counter): with both chains of a thread in flight the latency is paid once per pair.
template <typename real, typename RecT, typename UT, int IDX_BITS, bool REUSE>
__global__ void __launch_bounds__(256, 6) count2_kernel(const real* __restrict__ pos, const int32_t* __restrict__ types,
                                                     long long n, SortParams p, RecT* __restrict__ stage,
                                                     uint32_t* __restrict__ cnt, DeviceScalars* __restrict__ sc) {
    const long long j0 = blockIdx.x * 512LL + threadIdx.x;
    UT idx[2], type[2];
    bool live[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const long long j = j0 + 256 * u;
        live[u] = j < n;
        idx[u] = 0; type[u] = 0;
        if (live[u]) {
            if (REUSE) {
                const UT meta = stage[j].meta;
                idx[u] = meta & (((UT)1 << IDX_BITS) - 1);
                type[u] = meta >> IDX_BITS;
            } else {
                idx[u] = (UT)j;
                type[u] = (UT)(uint32_t)types[j];
            }
        }
    }
    real x[2][3];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int a = 0; a < 3; ++a) x[u][a] = live[u] ? pos[3 * idx[u] + a] : (real)0;
    unsigned int r1 = 0, bad = 0;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        uint32_t key = 0xffffffffu;                    // lanes past the end: a run of their own
        if (live[u]) {
            int cx, cy, cz;
            double dx, dy, dz;
            split_coord((double)x[u][0] * p.sx, p.Nx, cx, dx);
            split_coord((double)x[u][1] * p.sy, p.Ny, cy, dy);
            split_coord((double)x[u][2] * p.sz, p.Nz, cz, dz);
            int lx = cx - p.x0;
            if (lx < 0 || lx >= p.nxl) {
                bad += 1;
                lx = lx < 0 ? 0 : p.nxl - 1;
            }
            RecT r;
            r.ux = pack_coord<UT>(lx, dx, p.fbx);
            r.uy = pack_coord<UT>(cy, dy, p.fby);
            r.uz = pack_coord<UT>(cz, dz, p.fbz);
            r.meta = idx[u] | (type[u] << IDX_BITS);
            stage[j0 + 256 * u] = r;
            key = (uint32_t)(((long long)lx * p.Ny + cy) * p.nbz + zbin_of(cz, p.Nz));
        }
        unsigned head_lane, rank, count;
        warp_runs(key, head_lane, rank, count);
        if (rank == 0 && live[u]) r1 = max(r1, atomicAdd(&cnt[key], count) + count);
    }
    unsigned int m = __reduce_max_sync(0xffffffffu, r1);
    unsigned int b = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
        if (m > sc->max_cell_count) atomicMax(&sc->max_cell_count, m);
        if (b) atomicAdd(&sc->out_of_slab, b);
    }
}

// Pass 2 (after the scan): staged record j goes to the next free slot of its cell.
template <typename real, typename RecT, typename UT, int IDX_BITS, bool ROUTED, int PER>
__global__ void __launch_bounds__(256) scatter_kernel(
    const RecT* __restrict__ stage, const real* __restrict__ q, long long n, SortParams p,
    uint32_t* __restrict__ cur, RecT* __restrict__ rec, real* __restrict__ q_sorted,
    DeviceScalars* __restrict__ sc, const RouteTotals* __restrict__ rt, const uint32_t* __restrict__ keys,
    int reuse) {
    const long long limit = (ROUTED && reuse) ? (long long)rt->n_total : n;
    const long long trip = (long long)gridDim.x * (256 * PER);
    for (long long base = blockIdx.x * (long long)(256 * PER);; base += trip) {
    float aq = 0.f;
    RecT r[PER];
    uint32_t key[PER];
    bool live[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const long long j = base + 256 * u + threadIdx.x;
        key[u] = 0xffffffffu;
        live[u] = j < limit;
        if (ROUTED) {
            if (live[u]) {
                key[u] = keys[j];
                live[u] = key[u] != 0xffffffffu;
                if (live[u]) r[u] = stage[j];
            }
        } else if (live[u]) {
            r[u] = stage[j];
        }
    }
    uint32_t bs[PER];
    unsigned head_lane[PER], rank[PER], count[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        if (!ROUTED && live[u]) {
            const long long lx = (long long)(r[u].ux >> p.fbx), cy = (long long)(r[u].uy >> p.fby),
                            cz = (long long)(r[u].uz >> p.fbz);
            key[u] = (uint32_t)((lx * p.Ny + cy) * p.nbz + zbin_of((int)cz, p.Nz));
        }
        warp_runs(key[u], head_lane[u], rank[u], count[u]);
        bs[u] = 0;
        if (rank[u] == 0 && live[u]) bs[u] = atomicAdd(&cur[key[u]], count[u]);
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        bs[u] = __shfl_sync(0xffffffffu, bs[u], (int)head_lane[u]);
        if (live[u]) {
            const size_t slot = (size_t)bs[u] + rank[u];
            rec[slot] = r[u];
            if (q != nullptr) {
                const real qi = q[r[u].meta & (((UT)1 << IDX_BITS) - 1)];
                q_sorted[slot] = qi;
                aq = fmaxf(aq, fabsf((float)qi));
            }
        }
    }
    if (q != nullptr) {
        unsigned int m = __reduce_max_sync(0xffffffffu, __float_as_uint(aq));
        if ((threadIdx.x & 31) == 0 && m > sc->qmax_bits) atomicMax(&sc->qmax_bits, m);
    }
    if (!ROUTED || base + trip >= limit) break;
    }
}

// Pass 2, one slab: the same as scatter_kernel<..., false>, two staged records per thread (j and j + 256 of a
// 512-record block) with both loads, then both cursor atomics, in flight at once: the pass is a chain of three
// dependent memory operations per record (staged record -> cursor -> slot) and was latency-bound at one record per
// thread (ncu: long_scoreboard 36 of 44 stall cycles per issue).
template <typename real, typename RecT, typename UT, int IDX_BITS>
__global__ void __launch_bounds__(256) scatter2_kernel(
    const RecT* __restrict__ stage, const real* __restrict__ q, long long n, SortParams p,
    uint32_t* __restrict__ cur, RecT* __restrict__ rec, real* __restrict__ q_sorted, DeviceScalars* __restrict__ sc) {
    const long long j0 = blockIdx.x * 512LL + threadIdx.x;
    RecT r[2];
    uint32_t key[2];
    bool live[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const long long j = j0 + 256 * u;
        live[u] = j < n;
        key[u] = 0xffffffffu;
        if (live[u]) r[u] = stage[j];
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
        if (live[u]) {
            const long long lx = (long long)(r[u].ux >> p.fbx), cy = (long long)(r[u].uy >> p.fby),
                            cz = (long long)(r[u].uz >> p.fbz);
            key[u] = (uint32_t)((lx * p.Ny + cy) * p.nbz + zbin_of((int)cz, p.Nz));
        }
    unsigned head_lane[2], rank[2], count[2];
    uint32_t base[2] = {0, 0};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        warp_runs(key[u], head_lane[u], rank[u], count[u]);
        if (rank[u] == 0 && live[u]) base[u] = atomicAdd(&cur[key[u]], count[u]);
    }
    float aq = 0.f;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        base[u] = __shfl_sync(0xffffffffu, base[u], (int)head_lane[u]);
        if (live[u]) {
            const size_t slot = (size_t)base[u] + rank[u];
            rec[slot] = r[u];
            if (q != nullptr) {
                const real qi = q[r[u].meta & (((UT)1 << IDX_BITS) - 1)];
                q_sorted[slot] = qi;
                aq = fmaxf(aq, fabsf((float)qi));
            }
        }
    }
    if (q != nullptr) {
        unsigned int m = __reduce_max_sync(0xffffffffu, __float_as_uint(aq));
        if ((threadIdx.x & 31) == 0 && m > sc->qmax_bits) atomicMax(&sc->qmax_bits, m);
    }
}

// Charges into sorted order for a sort that was made without them (update_field_force_q is
// called after update_field on the same positions: main.py:1006-1058).  Several slabs: n = capacity
// bound, the live count is rt->n_total, guests read the charges their owners sent (gq).
template <typename real, typename RecT, typename UT, int IDX_BITS>
__global__ void __launch_bounds__(256) gather_charges_kernel(const RecT* __restrict__ rec,
                                                             const real* __restrict__ q, long long n,
                                                             real* __restrict__ q_sorted,
                                                             DeviceScalars* __restrict__ sc,
                                                             const RouteTotals* __restrict__ rt,
                                                             const real* __restrict__ gq, long long n_home) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    float aq = 0.f;
    if (rt != nullptr) n = (long long)rt->n_total;
    if (i < n) {
        const long long idx = (long long)(rec[i].meta & (((UT)1 << IDX_BITS) - 1));
        const real qi = (rt != nullptr && idx >= n_home) ? gq[idx - n_home] : q[idx];
        q_sorted[i] = qi;
        aq = fabsf((float)qi);
    }
    unsigned int m = __reduce_max_sync(0xffffffffu, __float_as_uint(aq));
    if ((threadIdx.x & 31) == 0 && m > sc->qmax_bits) atomicMax(&sc->qmax_bits, m);
}

// HYMD_B200_SCATTER=1: one staged record per thread (the first version of the pass; A/B)
static bool scatter_two() {
    const char* e = getenv("HYMD_B200_SCATTER");
    return !(e && e[0] == '1');
}

size_t scan_temp_bytes(long long n) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
    return bytes;
}

// ---- per-step routing between slabs (several GPUs) ----------------------------------------------------
// pmesh routes, on every paint / readout, a copy of each particle to the rank that owns its mesh cell and
// brings the read-out value back (pm.decompose + Layout.exchange / Layout.gather: main.py:977-980,
// field.py:574, 200).  The caller's particles therefore need not lie in their rank's slab: molecules live
// on the rank of their first bead (field.py:1156-1163) and every particle drifts between two
// domain_decomposition calls.  Here, per step and entirely on the device:
//   route_scan   every home particle whose x cell belongs to another slab q is appended (atomic slot) to
//                rank q's INBOX section for this rank -- position, type, (charge) stored straight into
//                q's HBM over NVLink -- and remembered in sent_idx;
//   barrier      carries the per-destination counts (comm_barrier_payload);
//   binning      home particles that are present + the guests of all inbox sections share one counting
//                sort; home particles that are away sit in an extra bin nobody paints;
//   readout      a guest's force is stored into its owner's RETURN section; after a barrier the owner
//                scatters the returned rows to force[sent_idx].
// No host synchronisation, no NCCL call; capacities are fixed (G guests per rank pair), an overflow
// raises status bit 1 (comm_check_status) instead of dropping particles silently.
struct RouteState {
    long long G = 0;                 // guest rows per (source, destination) pair
    void* in_pos = nullptr;          // [P*G][3] real   inbox: section r written by rank r
    int32_t* in_type = nullptr;      // [P*G]
    void* in_q = nullptr;            // [P*G] real
    void* ret = nullptr;             // [P*G][3] real   section q = forces of the guests sent to rank q
    RoutePeers peers;
    uint32_t* send_count = nullptr;  // [HYMD_MAX_PEERS]
    int32_t* sent_idx = nullptr;     // [P*G]: home index of guest k sent to rank q at q*G + k
    void* gstage = nullptr;          // [P*G] staged guest records
    uint32_t* gkeys = nullptr;       // [P*G]
    uint32_t* keys = nullptr;        // [cap + P*G] bins of the staged home records
    long long keys_cap = 0;
    RouteTotals* totals = nullptr;   // {n_work, n_total} of the last sort
    bool sent_charges = false;
};

// Flat index over the guests actually present (count[r] rows of every section r != rank, each capped at G)
// -> row e = r * G + k of the [P][G] guest arrays; false past the end.
__device__ __forceinline__ bool guest_row(long long flat, const uint32_t* __restrict__ count, int P, int rank,
                                          long long G, long long& e) {
    for (int r = 0; r < P; ++r) {
        if (r == rank) continue;
        const long long c = (long long)min(count[r], (uint32_t)G);
        if (flat < c) { e = (long long)r * G + flat; return true; }
        flat -= c;
    }
    return false;
}
constexpr int GUEST_BLOCKS = 148 * 2;      // grid of the guest passes: their work is the guest count, not P * G

// charges of the guests already sent (hymd_set_charges after a sort without charges)
template <typename real>
__global__ void __launch_bounds__(256) route_charges_kernel(const real* __restrict__ q, int P, int rank, long long G,
                                                            RoutePeers peers, const uint32_t* __restrict__ send_count,
                                                            const int32_t* __restrict__ sent_idx) {
    for (long long flat = blockIdx.x * (long long)blockDim.x + threadIdx.x;; flat += (long long)gridDim.x * blockDim.x) {
        long long e;
        if (!guest_row(flat, send_count, P, rank, G, e)) return;
        reinterpret_cast<real*>(peers.q[(int)(e / G)])[(long long)rank * G + e % G] = q[sent_idx[e]];
    }
}

template <typename real, typename RecT, typename UT, int IDX_BITS>
__global__ void __launch_bounds__(256) guest_count_kernel(
    const real* __restrict__ gpos, const int32_t* __restrict__ gtype, const uint32_t* __restrict__ recv_count,
    int P, int rank, long long G, long long n_home, SortParams p, RecT* __restrict__ gstage,
    uint32_t* __restrict__ gkeys, uint32_t* __restrict__ cnt, DeviceScalars* __restrict__ sc, unsigned int* status) {
    for (long long flat = blockIdx.x * (long long)blockDim.x + threadIdx.x;; flat += (long long)gridDim.x * blockDim.x) {
    long long e;
    if (!guest_row(flat, recv_count, P, rank, G, e)) return;
    int cx, cy, cz;
    double dx, dy, dz;
    split_coord((double)gpos[3 * e + 0] * p.sx, p.Nx, cx, dx);
    split_coord((double)gpos[3 * e + 1] * p.sy, p.Ny, cy, dy);
    split_coord((double)gpos[3 * e + 2] * p.sz, p.Nz, cz, dz);
    int lx = cx - p.x0;
    if (lx < 0 || lx >= p.nxl) {               // cannot happen: sender and receiver use the same rule
        if (status) atomicOr(status, 4u);
        gkeys[e] = 0xffffffffu;
        continue;
    }
    RecT rc;
    rc.ux = pack_coord<UT>(lx, dx, p.fbx);
    rc.uy = pack_coord<UT>(cy, dy, p.fby);
    rc.uz = pack_coord<UT>(cz, dz, p.fbz);
    rc.meta = (UT)(n_home + e) | ((UT)(uint32_t)gtype[e] << IDX_BITS);
    gstage[e] = rc;
    const uint32_t key = (uint32_t)(((long long)lx * p.Ny + cy) * p.nbz + zbin_of(cz, p.Nz));
    gkeys[e] = key;
    const uint32_t c1 = atomicAdd(&cnt[key], 1u) + 1u;
    if (c1 > sc->max_cell_count) atomicMax(&sc->max_cell_count, c1);
    }
}

template <typename real, typename RecT>
__global__ void __launch_bounds__(256) guest_scatter_kernel(
    const RecT* __restrict__ gstage, const uint32_t* __restrict__ gkeys, const real* __restrict__ gq,
    const uint32_t* __restrict__ recv_count, int P, int rank, long long G, uint32_t* __restrict__ cur,
    RecT* __restrict__ rec, real* __restrict__ q_sorted, DeviceScalars* __restrict__ sc, int with_q) {
    for (long long flat = blockIdx.x * (long long)blockDim.x + threadIdx.x;; flat += (long long)gridDim.x * blockDim.x) {
        long long e;
        if (!guest_row(flat, recv_count, P, rank, G, e)) return;
        const uint32_t key = gkeys[e];
        if (key == 0xffffffffu) continue;
        const uint32_t slot = atomicAdd(&cur[key], 1u);
        rec[slot] = gstage[e];
        if (with_q) {
            const real qi = gq[e];
            q_sorted[slot] = qi;
            atomicMax(&sc->qmax_bits, __float_as_uint(fabsf((float)qi)));
        }
    }
}

// after the scatter passes: cell_start[k] = start of bin k, bin ncell = the away bin
__global__ void route_totals_kernel(const uint32_t* __restrict__ cell_start, long long ncell,
                                    RouteTotals* __restrict__ rt) {
    rt->n_work = cell_start[ncell];
    rt->n_total = cell_start[ncell + 1];
}

// forces of the guests this rank sent away, back at their home index
template <typename real>
__global__ void __launch_bounds__(256) route_return_kernel(const real* __restrict__ ret, int P, int rank, long long G,
                                                           const uint32_t* __restrict__ send_count,
                                                           const int32_t* __restrict__ sent_idx,
                                                           real* __restrict__ force) {
    for (long long flat = blockIdx.x * (long long)blockDim.x + threadIdx.x;; flat += (long long)gridDim.x * blockDim.x) {
        long long e;
        if (!guest_row(flat, send_count, P, rank, G, e)) return;
        real* o = force + 3 * (long long)sent_idx[e];
        o[0] = ret[3 * e]; o[1] = ret[3 * e + 1]; o[2] = ret[3 * e + 2];
    }
}

static int route_alloc(void** p, size_t bytes) {
    if (cudaMalloc(p, bytes ? bytes : 16) != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) failed in the routing buffers", bytes);
        return HYMD_ERR_NOMEM;
    }
    cudaMemset(*p, 0, bytes ? bytes : 16);
    return HYMD_OK;
}

void route_destroy(hymd_ctx* c) {
    RouteState* r = c->route;
    if (!r) return;
    void* bufs[] = {r->in_pos, r->in_type, r->in_q, r->ret, r->send_count, r->sent_idx,
                    r->gstage, r->gkeys, r->keys, r->totals};
    for (void* b : bufs)
        if (b) cudaFree(b);
    delete r;
    c->route = nullptr;
}

long long route_guest_rows(const hymd_ctx* c) { return c->route ? c->route->G * c->g.P : 0; }
const RouteTotals* route_totals(const hymd_ctx* c) { return c->route ? c->route->totals : nullptr; }
const uint32_t* route_send_counts(const hymd_ctx* c) { return c->route ? c->route->send_count : nullptr; }

// Collective (first sort of a context with several slabs): sizes the guest buffers from this rank's
// particle count -- G = max(16384, n / 4) rows per rank pair, or HYMD_B200_GUEST_CAPACITY -- and
// exchanges their peer addresses.  The buffers are never re-allocated (peers hold their addresses).
int route_prepare(hymd_ctx* c, int64_t n, cudaStream_t s) {
    if (c->route || c->g.P == 1) return HYMD_OK;
    RouteState* r = new RouteState();
    c->route = r;
    const int P = c->g.P;
    long long G = n / 4 > 16384 ? n / 4 : 16384;
    if (const char* e = getenv("HYMD_B200_GUEST_CAPACITY")) {
        const long long v = atoll(e);
        if (v > 0) G = v;
    }
    // all ranks must agree on G (it is part of the peer addressing): take the maximum
    {
        uint32_t* d_tmp = nullptr;
        HYMD_CHECK(route_alloc((void**)&d_tmp, 2 * HYMD_MAX_PEERS * sizeof(unsigned long long)));
        unsigned long long mine = (unsigned long long)G, all[HYMD_MAX_PEERS];
        HYMD_CUDA(cudaMemcpyAsync(d_tmp, &mine, sizeof(mine), cudaMemcpyHostToDevice, s));
        HYMD_CHECK(comm_allgather_host(c, d_tmp, (unsigned long long*)d_tmp + 1, sizeof(mine), s));
        HYMD_CUDA(cudaMemcpyAsync(all, (unsigned long long*)d_tmp + 1, sizeof(mine) * P, cudaMemcpyDeviceToHost, s));
        HYMD_CUDA(cudaStreamSynchronize(s));
        for (int q = 0; q < P; ++q) if ((long long)all[q] > G) G = (long long)all[q];
        cudaFree(d_tmp);
    }
    r->G = G;
    const size_t rows = (size_t)P * G;
    HYMD_CHECK(route_alloc(&r->in_pos, rows * 3 * c->rsz));
    HYMD_CHECK(route_alloc((void**)&r->in_type, rows * 4));
    HYMD_CHECK(route_alloc(&r->in_q, rows * c->rsz));
    HYMD_CHECK(route_alloc(&r->ret, rows * 3 * c->rsz));
    HYMD_CHECK(route_alloc((void**)&r->send_count, HYMD_MAX_PEERS * 4));
    HYMD_CHECK(route_alloc((void**)&r->sent_idx, rows * 4));
    HYMD_CHECK(route_alloc(&r->gstage, rows * (c->f64 ? sizeof(Rec64) : sizeof(Rec32))));
    HYMD_CHECK(route_alloc((void**)&r->gkeys, rows * 4));
    HYMD_CHECK(route_alloc((void**)&r->totals, sizeof(RouteTotals)));
    HYMD_CUDA(cudaDeviceSynchronize());
    HYMD_CHECK(comm_peer_ptrs(c, r->in_pos, r->peers.pos, s));
    HYMD_CHECK(comm_peer_ptrs(c, r->in_type, r->peers.type, s));
    HYMD_CHECK(comm_peer_ptrs(c, r->in_q, r->peers.q, s));
    HYMD_CHECK(comm_peer_ptrs(c, r->ret, r->peers.ret, s));
    return HYMD_OK;
}

template <typename real>
static int route_gather_charges(hymd_ctx* c, const void* d_q, cudaStream_t s) {
    // guests were sent without charges: send them now (same slots), then gather into sorted order
    RouteState* r = c->route;
    const Geometry& g = c->g;
    if (!r->sent_charges) {
        if (c->peer_busy & PEER_INBOX) HYMD_CHECK(comm_barrier(c, s));
        const long long rows = (long long)g.P * r->G;
        route_charges_kernel<real><<<GUEST_BLOCKS, 256, 0, s>>>(
            (const real*)d_q, g.P, g.rank, r->G, r->peers, r->send_count, r->sent_idx);
        HYMD_LAUNCH_CHECK(c);
        HYMD_CHECK(comm_barrier(c, s));
        c->peer_busy |= PEER_INBOX;
        r->sent_charges = true;
    }
    return HYMD_OK;
}

int gather_charges(hymd_ctx* c, const void* d_q, cudaStream_t s) {
    const long long n = c->np;
    RouteState* r = c->route;
    if (r) HYMD_CHECK(c->f64 ? route_gather_charges<double>(c, d_q, s) : route_gather_charges<float>(c, d_q, s));
    const long long bound = r ? n + (long long)c->g.P * r->G : n;
    if (bound == 0) return HYMD_OK;
    const unsigned int blocks = (unsigned int)((bound + 255) / 256);
    if (c->f64)
        gather_charges_kernel<double, Rec64, unsigned long long, REC64_IDX_BITS>
            <<<blocks, 256, 0, s>>>((const Rec64*)c->rec, (const double*)d_q, n, (double*)c->q_sorted, c->scalars,
                                    r ? r->totals : nullptr, r ? (const double*)r->in_q : nullptr, n);
    else
        gather_charges_kernel<float, Rec32, uint32_t, REC32_IDX_BITS>
            <<<blocks, 256, 0, s>>>((const Rec32*)c->rec, (const float*)d_q, n, (float*)c->q_sorted, c->scalars,
                                    r ? r->totals : nullptr, r ? (const float*)r->in_q : nullptr, n);
    HYMD_LAUNCH_CHECK(c);
    return HYMD_OK;
}

// readout side: forces of the guests that were sent away, back into the caller's array
int route_return(hymd_ctx* c, void* d_force, cudaStream_t s) {
    RouteState* r = c->route;
    if (!r) return HYMD_OK;
    const Geometry& g = c->g;
    HYMD_CHECK(comm_barrier(c, s));                   // every rank's readout has stored its guests' rows
    const unsigned blocks = GUEST_BLOCKS;
    if (c->f64) route_return_kernel<double><<<blocks, 256, 0, s>>>((const double*)r->ret, g.P, g.rank, r->G,
                                                                   r->send_count, r->sent_idx, (double*)d_force);
    else route_return_kernel<float><<<blocks, 256, 0, s>>>((const float*)r->ret, g.P, g.rank, r->G,
                                                           r->send_count, r->sent_idx, (float*)d_force);
    HYMD_LAUNCH_CHECK(c);
    c->peer_busy |= PEER_RET;
    return HYMD_OK;
}

int route_acquire_return(hymd_ctx* c, cudaStream_t s) {
    // before a readout stores guests' rows into the owners' return sections: the owners must have consumed
    // the previous contents (a barrier has passed since their route_return)
    if (c->route && (c->peer_busy & PEER_RET)) return comm_barrier(c, s);
    return HYMD_OK;
}

void route_peer_ret(const hymd_ctx* c, void** out, long long* G) {
    for (int q = 0; q < HYMD_MAX_PEERS; ++q) out[q] = c->route ? c->route->peers.ret[q] : nullptr;
    *G = c->route ? c->route->G : 0;
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
    const bool routed = g.P > 1;
    RouteState* r = nullptr;
    long long rows = 0;
    if (routed) {
        HYMD_CHECK(route_prepare(c, n, s));
        r = c->route;
        rows = (long long)g.P * r->G;
        if (n + rows >= ((long long)1 << IDX_BITS)) {
            set_error("n + guest rows = %lld exceeds the record index range", n + rows);
            return HYMD_ERR_CAPACITY;
        }
        if (r->keys_cap < c->cap + rows) {
            if (r->keys) { cudaStreamSynchronize(s); cudaFree(r->keys); r->keys = nullptr; }
            r->keys_cap = c->cap + rows;
            HYMD_CHECK(route_alloc((void**)&r->keys, (size_t)r->keys_cap * 4));
        }
        // guests go out from inside the binning pass (count_kernel): the owners' inboxes must be free
        if (c->peer_busy & PEER_INBOX) HYMD_CHECK(comm_barrier(c, s));
        HYMD_CUDA(cudaMemsetAsync(r->send_count, 0, HYMD_MAX_PEERS * 4, s));
        r->sent_charges = d_q != nullptr;
    }
    RouteOut ro;
    memset(&ro, 0, sizeof(ro));
    if (routed) {
        ro.peers = r->peers; ro.send_count = r->send_count; ro.sent_idx = r->sent_idx;
        ro.status = comm_status_device(c); ro.q = d_q; ro.G = r->G; ro.rank = g.rank;
    }
    uint32_t* cur = c->cell_start + 1;
    HYMD_CUDA(cudaMemsetAsync(c->cell_start, 0, (size_t)(ncell + 3) * sizeof(uint32_t), s));
    HYMD_CUDA(cudaMemsetAsync(c->scalars, 0, sizeof(DeviceScalars), s));
    // stage = the buffer holding the previous sorted records (overwritten in place), out = the other
    RecT* stage = (RecT*)c->rec;
    RecT* out = (RecT*)c->rec_alt;
    // routed: the grid covers the home particles, the staged guests behind them are reached by a second
    // (block-uniform) trip of the first blocks; at least one block so that a rank without particles of
    // its own still retires the guests it staged last step
    const long long span = routed ? (n > 0 ? n : 1) : n;
    const unsigned int blocks = (unsigned int)((span + 255) / 256);
    const unsigned int blocks2 = (unsigned int)((span + 511) / 512);      // two records per thread
    if (span > 0) {
        if (routed) {
            if (reuse)
                count_kernel<real, RecT, UT, IDX_BITS, true, true, 2><<<blocks2, 256, 0, s>>>(
                    (const real*)d_pos, d_types, n, p, stage, cur, c->scalars, r->totals, r->keys, (uint32_t)ncell, ro);
            else
                count_kernel<real, RecT, UT, IDX_BITS, false, true, 2><<<blocks2, 256, 0, s>>>(
                    (const real*)d_pos, d_types, n, p, stage, cur, c->scalars, r->totals, r->keys, (uint32_t)ncell, ro);
        } else if (scatter_two()) {
            if (reuse)
                count2_kernel<real, RecT, UT, IDX_BITS, true><<<blocks2, 256, 0, s>>>(
                    (const real*)d_pos, d_types, n, p, stage, cur, c->scalars);
            else
                count2_kernel<real, RecT, UT, IDX_BITS, false><<<blocks2, 256, 0, s>>>(
                    (const real*)d_pos, d_types, n, p, stage, cur, c->scalars);
        } else if (reuse) {
            count_kernel<real, RecT, UT, IDX_BITS, true, false, 1><<<blocks, 256, 0, s>>>(
                (const real*)d_pos, d_types, n, p, stage, cur, c->scalars, nullptr, nullptr, 0u, ro);
        } else {
            count_kernel<real, RecT, UT, IDX_BITS, false, false, 1><<<blocks, 256, 0, s>>>(
                (const real*)d_pos, d_types, n, p, stage, cur, c->scalars, nullptr, nullptr, 0u, ro);
        }
        HYMD_LAUNCH_CHECK(c);
    }
    const unsigned gblocks = GUEST_BLOCKS;
    if (routed) {
        // the per-destination counts ride on the barrier; after it every inbox section is complete
        HYMD_CHECK(comm_barrier_payload(c, r->send_count, s));
        c->peer_busy |= PEER_INBOX;
        guest_count_kernel<real, RecT, UT, IDX_BITS><<<gblocks, 256, 0, s>>>(
            (const real*)r->in_pos, r->in_type, comm_payload(c), g.P, g.rank, r->G, n, p, (RecT*)r->gstage, r->gkeys,
            cur, c->scalars, comm_status_device(c));
        HYMD_LAUNCH_CHECK(c);
    }
    size_t tmp = c->scan_tmp_bytes;
    HYMD_CUDA(cub::DeviceScan::ExclusiveSum(c->scan_tmp, tmp, cur, cur, (int)(routed ? ncell + 1 : ncell), s));
    c->launches += 2;  // cub scan: init + scan kernels
    if (span > 0) {
        if (routed)
            scatter_kernel<real, RecT, UT, IDX_BITS, true, 2><<<blocks2, 256, 0, s>>>(
                stage, (const real*)d_q, n, p, cur, out, (real*)c->q_sorted, c->scalars, r->totals, r->keys, reuse ? 1 : 0);
        else if (scatter_two())
            scatter2_kernel<real, RecT, UT, IDX_BITS><<<(unsigned int)((span + 511) / 512), 256, 0, s>>>(
                stage, (const real*)d_q, n, p, cur, out, (real*)c->q_sorted, c->scalars);
        else
            scatter_kernel<real, RecT, UT, IDX_BITS, false, 1><<<blocks, 256, 0, s>>>(
                stage, (const real*)d_q, n, p, cur, out, (real*)c->q_sorted, c->scalars, nullptr, nullptr, 0);
        HYMD_LAUNCH_CHECK(c);
    }
    if (routed) {
        guest_scatter_kernel<real, RecT><<<gblocks, 256, 0, s>>>(
            (const RecT*)r->gstage, r->gkeys, (const real*)r->in_q, comm_payload(c), g.P, g.rank, r->G, cur, out,
            (real*)c->q_sorted, c->scalars, d_q != nullptr ? 1 : 0);
        HYMD_LAUNCH_CHECK(c);
        route_totals_kernel<<<1, 1, 0, s>>>(c->cell_start, ncell, r->totals);
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
