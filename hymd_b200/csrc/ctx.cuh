// Internal context shared by the translation units of libhymd_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cufft.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/hymd_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libhymd_b200 is written for sm_100a (B200) only"
#endif

namespace hymd {

void set_error(const char* fmt, ...);

#define HYMD_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            hymd::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                 \
                            cudaGetErrorString(e_));                                      \
            return HYMD_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

#define HYMD_CUFFT(call)                                                                  \
    do {                                                                                  \
        cufftResult r_ = (call);                                                          \
        if (r_ != CUFFT_SUCCESS) {                                                        \
            hymd::set_error("%s:%d: %s -> cufft error %d", __FILE__, __LINE__, #call,     \
                            (int)r_);                                                     \
            return HYMD_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

#define HYMD_CHECK(call)                                                                  \
    do {                                                                                  \
        int s_ = (call);                                                                  \
        if (s_ != HYMD_OK) return s_;                                                     \
    } while (0)

#define HYMD_LAUNCH_CHECK(ctx)                                                            \
    do {                                                                                  \
        (ctx)->launches++;                                                                \
        cudaError_t e_ = cudaGetLastError();                                              \
        if (e_ != cudaSuccess) {                                                          \
            hymd::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__,             \
                            cudaGetErrorString(e_));                                      \
            return HYMD_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

// Tile of mesh cells a paint CTA owns (vertices) / a readout CTA stages through TMA.
constexpr int PAINT_TX = 8, PAINT_TY = 8, PAINT_TZ = 32;

// Binning granularity of the counting sort (sort.cu).  Paint and readout never need the particles
// of a single cell, only those of the z-range a 32-vertex tile draws from -- cells z0-1 .. z0+31
// for the paint, z0 .. z0+31 for the tiled readout -- so every (x,y) cell row is cut into bins
//     [0..30] [31] [32..62] [63] ...          (the cell Nz-1 is always a bin of its own)
// i.e. 2 bins per tile: bin 2t = cells 32t .. 32t+30, bin 2t+1 = the boundary cell shared by the
// tiles t and t+1 (and by the periodic wrap).  1/16 of the counters of a per-cell sort: the
// counter array (4 MB instead of 67 MB at 256^3) stays in L2 and the scan is negligible.
constexpr int ZBIN = 32;
static_assert(PAINT_TZ == ZBIN, "paint tiles and sort bins share their z extent");
__host__ __device__ inline int zbins_per_row(int Nz) { return 2 * ((Nz + ZBIN - 1) / ZBIN); }
__host__ __device__ inline int zbin_of(int cz, int Nz) {
    return 2 * (cz / ZBIN) + ((cz % ZBIN == ZBIN - 1 || cz == Nz - 1) ? 1 : 0);
}

// Sorted particle record.  Coordinates are unsigned fixed point in local-slab grid units:
// integer part = mesh cell, fraction = CIC offset d.  meta = original index | type << idx_bits.
struct __align__(16) Rec32 {
    uint32_t ux, uy, uz, meta;
};
struct __align__(16) Rec64 {
    unsigned long long ux, uy, uz, meta;
};
enum PeerBuffer : unsigned { PEER_WORK = 1u, PEER_HALO = 2u, PEER_K = 4u, PEER_MESH = 8u, PEER_INBOX = 16u, PEER_RET = 32u };
constexpr int HYMD_MAX_PEERS = 8;    // slabs (GPUs of one NVLink domain)
constexpr int REC32_IDX_BITS = 27;  // <= 134M particles per GPU, <= 32 types
constexpr int REC64_IDX_BITS = 40;

struct DeviceScalars {
    unsigned int max_cell_count;   // max particles in one sort bin (>= any cell: bound for the paint fixed-point scale)
    unsigned int qmax_bits;        // max |charge| as float bits
    unsigned int out_of_slab;      // particles outside the local slab (multi-GPU)
    unsigned int pad;
};

struct Geometry {
    int Nx, Ny, Nz;        // global mesh
    int nxl, x0;           // local slab [x0, x0+nxl)
    int P, rank;
    int nyl, y0;           // k-space (transposed) local y range
    int Nzc, Nzcp;         // complex z extent and padded pitch (even)
    int Nzp;               // padded real z pitch of ghost meshes
    int fbx, fby, fbz;     // fraction bits of the fixed-point coordinates
    double box[3];
    long long ncell;       // sort bins: nxl*Ny*zbins_per_row(Nz)
    int nbz;               // bins per (x,y) cell row
    long long real_elems;  // nxl*Ny*Nz
    long long ghost_elems; // (nxl+1)*(Ny+1)*Nzp
    long long k_elems;     // Nx*nyl*Nzcp  (complex elements per spectrum)
    int vx;                // x planes stored per real field: nxl (P == 1) or nxl+1 (ghost plane)
};

// Strides (in complex elements) of a k-space buffer holding F spectra.  Element (f, ix, iyl, iz)
// lives at f*fs + ix*xs + iyl*Nzcp + iz.
//   single GPU : [f][x][ky][kz]   fs = Nx*Ny*Nzcp, xs = Ny*Nzcp      (cuFFT batch order)
//   P slabs    : [x][f][kyl][kz]  fs = nyl*Nzcp,   xs = F*nyl*Nzcp   (all-to-all receive order)
struct KLayout {
    long long fs, xs;
};

struct PlanEntry {
    int kind, batch;
    cufftHandle h;
};
struct Comm;
struct MigrateState;
struct RouteState;
// device-side particle counts of the last sort with several slabs (sort.cu, per-step routing)
struct RouteTotals {
    unsigned int n_work;    // records painted / read out here: home particles present + guests
    unsigned int n_total;   // n_work + the home particles that are away (kept in an extra bin)
};
struct GpeState;
struct GraphCache;

struct PhaseInterval {
    int phase;
    cudaEvent_t a, b;
};

}  // namespace hymd

struct hymd_ctx {
    hymd_config cfg;
    hymd::Geometry g;
    int dev;
    bool f64;
    int T, U;
    int urow[HYMD_MAX_TYPES];     // type -> unique potential row
    int rowrep[HYMD_MAX_TYPES];   // unique row -> representative type
    size_t rsz;                   // sizeof(real)

    // particles
    int64_t np, cap;
    bool sorted, has_charges;
    void* rec;              // cell-sorted records of the last sort
    void* rec_alt;          // the other half of the double buffer (staging area of the next sort)
    int64_t order_n;        // particle count the order in `rec` is valid for (-1: none)
    uint32_t* cell_start;   // ncell + 2: [0] = 0, then the per-bin cursors (see sort.cu)
    void* q_sorted;
    hymd::DeviceScalars* scalars;
    void* scan_tmp;
    size_t scan_tmp_bytes;

    // k-space tables (real): hx,hy,hz Gaussian factors, kx,ky,kz wave numbers; Au (U x T) / M
    void* tab;            // packed: hx[Nx] hy[Ny] hz[Nzc] kx[Nx] ky[Ny] kz[Nzc]
    void* xtw;            // Nx complex twiddles exp(-2 pi i j / Nx) for the fused x-line kernel
    bool fused;           // fused x-line kernel in use (power-of-two Nx)
    bool plane;           // one-pass (y,z) plane transforms in use (planefft.cu)
    void* ytw;            // Ny / Nz complex twiddles of the plane transforms
    void* ztw;
    void* plane_scratch;  // per-CTA scratch planes of the plane transforms (L2 resident)
    size_t plane_scratch_bytes;
    void* Au;             // U*T reals, already divided by M
    void* cu;             // U offsets (added at k = 0 for v_ext)
    int* d_urow;          // T ints
    void* outscale;       // T reals: m_t / dV (paint output scale, without the fixed-point factor)

    // fields
    void* phi;            // T x real_elems
    void* phi_hat;        // T x k_elems complex (raw spectra of phi, unnormalised)
    void* f_hat;          // 3U x k_elems complex
    void* gmesh;          // 3U x ghost_elems real
    void* v_hat;          // U x k_elems (lazy)
    void* phif_hat;       // T x k_elems (lazy, filtered & normalised = reference phi_fourier)
    void* tmp_hat;        // max(T,U) x k_elems (lazy scratch for c2r of by-products)
    void* v_ext;          // T x real_elems (lazy)
    void* lap_hat;        // 3T x k_elems (lazy scratch): -k_d^2 phi_fourier
    void* lap;            // 3T x real_elems (lazy): phi_laplacian[t][d], field.py:406-425
    bool have_lap;          // lap valid for the current spectra
    bool phi_is_filtered;   // phi holds the filtered densities (reference semantics after update_field)
    bool have_phi_hat;      // raw density spectra of the last paint are valid
    bool have_phif;         // phif_hat valid for the current spectra
    bool have_forces;       // gmesh valid
    bool have_psi;
    bool have_phiq_hat;
    // PME
    void* phi_q;          // real_elems
    void* phiq_hat;       // k_elems raw
    void* phiqf_hat;      // k_elems filtered/normalised (reference phi_q_fourier)
    void* e_hat;          // 3 x k_elems: E_x,E_y,E_z
    void* psi_hat;        // k_elems (lazy)
    void* emesh;          // 3 x ghost_elems
    void* psi;            // real_elems

    // cuFFT: plans are created on first use and cached by (kind, batch); they share one work
    // area (everything runs on one stream)
    std::vector<hymd::PlanEntry>* plans;
    void* fft_work;
    size_t fft_work_bytes;
    bool grad2;             // fused x-line + plane kernels: 2 spectra per potential row cross HBM (and the
                            // inverse transpose) instead of 3; the plane c2r applies k_y / k_z itself
    bool slab;              // slab pipeline (2-D cuFFT per plane + x transform) instead of 3-D plans
    void* wA;               // slab work: F x (nxl+1) x Ny x Nzcp complex (2-D transform side)
    void* wS;               // slab work: all-to-all staging, F x nxl x Ny x Nzcp complex
    size_t wA_bytes, wS_bytes;
    void* halo;             // ghost-plane exchange staging
    size_t halo_bytes;
    hymd::Comm* comm;       // NCCL communicator (world_size > 1)
    bool p2p;               // exchanges store into peer memory over NVLink (CUDA IPC) instead of NCCL send/recv
    int xmode;              // how the FFT transposes cross NVLink (slabfft.cu): 0 pack / unpack push kernels,
                            // 1 stores issued by the plane r2c / x-line kernels, 2 blocked layouts + contiguous peer copies
    bool fused_push;        // xmode == 1
    bool xcopy_kernel;      // xmode == 2: the blocks are moved by an SM copy kernel instead of the copy engines
    // xmode == 2, pipelined exchange (slabfft.cu): the copies of field f run on a second, low-priority stream while
    // the plane kernel transforms field f + 1 (forward) / row u - 1 (inverse)
    int xpipe;              // 0 off, 1 on when a field's block is large enough, 2 always (tests)
    int sm_count;           // multiprocessors of the device
    int plane_sm_reserve;   // > 0 while the pipeline runs: the persistent plane kernels leave this many SMs to the copies
    cudaStream_t xstream;                               // SM copy kernel variant (HYMD_B200_XPIPE_COPY=kernel)
    cudaEvent_t xev[2 * HYMD_MAX_TYPES + 2];             // "piece ready" on the main stream / copied (kernel variant)
    cudaStream_t xpeer[hymd::HYMD_MAX_PEERS];                  // copy-engine variant (default): one stream per destination
    cudaEvent_t xdone[hymd::HYMD_MAX_PEERS][HYMD_MAX_TYPES + 1];
    bool xpipe_ce;
    bool xpushed;           // the x-line kernel has already stored its output into the peers' work buffers
    unsigned peer_busy;     // PEER_* buffers whose local consumers were enqueued after the last barrier:
                            // a peer may not overwrite them before another barrier (same call sequence
                            // on every rank, so the flags agree)
    hymd::MigrateState* mig;
    hymd::RouteState* route;   // per-step routing of particles outside their home slab (several slabs)
    hymd::GpeState* gpe;    // general-Poisson-equation electrostatics (gpe.cu), allocated on first use
    hymd::GraphCache* graphs;  // CUDA-graph replay of the per-step field update (graph.cu), allocated on first use

    // readout TMA
    CUtensorMap tmap_gmesh, tmap_emesh;
    int rtx, rty, rtz, rbz;       // readout tile and box z extent
    int rstages;                  // box buffers in the readout TMA ring
    size_t readout_smem, readout_smem_pme;

    int64_t launches;

    // phase timing (hymd_ctx_set_timing)
    bool timing;
    std::vector<cudaEvent_t>* ev_pool;      // free events
    std::vector<hymd::PhaseInterval>* ev_open;   // recorded, not yet read
};

namespace hymd {
// Brackets a phase with CUDA events on stream s while timing is enabled (destructor closes it,
// so early returns through HYMD_CHECK are covered).
struct PhaseScope {
    hymd_ctx* c;
    cudaStream_t s;
    cudaEvent_t a, b;
    int phase;
    bool on;
    PhaseScope(hymd_ctx* c_, int phase_, cudaStream_t s_) : c(c_), s(s_), phase(phase_), on(c_->timing) {
        if (!on) return;
        a = take(); b = take();
        cudaEventRecord(a, s);
    }
    ~PhaseScope() {
        if (!on) return;
        cudaEventRecord(b, s);
        c->ev_open->push_back({phase, a, b});
    }
    cudaEvent_t take() {
        cudaEvent_t e;
        if (!c->ev_pool->empty()) { e = c->ev_pool->back(); c->ev_pool->pop_back(); }
        else cudaEventCreate(&e);
        return e;
    }
};

// sort.cu
int sort_particles(hymd_ctx* c, const void* d_pos, const int32_t* d_types, const void* d_q,
                   int64_t n, bool reuse, cudaStream_t s);
size_t scan_temp_bytes(long long n);
int gather_charges(hymd_ctx* c, const void* d_q, cudaStream_t s);
// per-step routing (sort.cu)
int route_prepare(hymd_ctx* c, int64_t n, cudaStream_t s);
void route_destroy(hymd_ctx* c);
long long route_guest_rows(const hymd_ctx* c);
const RouteTotals* route_totals(const hymd_ctx* c);
const uint32_t* route_send_counts(const hymd_ctx* c);   // [HYMD_MAX_PEERS] guests sent to each slab by the last sort (device)
int route_return(hymd_ctx* c, void* d_force, cudaStream_t s);
int route_acquire_return(hymd_ctx* c, cudaStream_t s);
void route_peer_ret(const hymd_ctx* c, void** out, long long* G);
// paint.cu
int paint_types(hymd_ctx* c, cudaStream_t s);
int paint_charges(hymd_ctx* c, cudaStream_t s);
// kspace.cu
int kspace_forces(hymd_ctx* c, bool want_v, bool want_phif, cudaStream_t s);
int kspace_pme(hymd_ctx* c, bool want_psi, cudaStream_t s);
int kspace_laplacian(hymd_ctx* c, cudaStream_t s);
// slabfft.cu: 3-D transforms of F fields between the real layout [f][vx][Ny][Nz] (or the
// ghost-padded force-mesh layout) and the k layout klayout(c, F)
KLayout klayout(const hymd_ctx* c, int F);
int fft_forward(hymd_ctx* c, void* real_in, int F, void* k_out, cudaStream_t s);
int fft_forward_yz(hymd_ctx* c, void* real_in, int F, void* k_out, cudaStream_t s);
int fft_inverse(hymd_ctx* c, void* k_in, int F, void* real_out, bool ghost, cudaStream_t s);
int fft_inverse_xdone(hymd_ctx* c, void* k_in, int F, void* real_out, bool ghost, cudaStream_t s,
                      bool derive = false);
int ensure_work(hymd_ctx* c, int F);
void destroy_plans(hymd_ctx* c);
int halo_reduce(hymd_ctx* c, void* fields, int F, cudaStream_t s);
int halo_fetch(hymd_ctx* c, void* ghost_meshes, int F, cudaStream_t s);
// planefft.cu
bool plane_supported(const hymd_ctx* c);
// f0, nf: only fields (derive: potential rows) f0 .. f0 + nf - 1 of the F-field layouts (nf < 0: all from f0)
int plane_forward(hymd_ctx* c, const void* real_in, long long r_fs, int F, int nplanes, void* k_out,
                  long long k_fs, cudaStream_t s, void* const* push_peers = nullptr, int f0 = 0, int nf = -1);
int plane_inverse(hymd_ctx* c, const void* k_in, long long k_fs, int F, int nplanes, void* real_out,
                  bool ghost, bool derive, cudaStream_t s, bool blocked = false, int f0 = 0, int nf = -1);
// xline.cu
bool xline_supported(const hymd_ctx* c);
int xline_forces(hymd_ctx* c, const void* in, void* fout, void* vout, void* pfout, cudaStream_t s,
                 void* const* push_peers = nullptr);
int xline_pme(hymd_ctx* c, const void* in, void* fout, void* psi_out, void* rhof_out, cudaStream_t s,
              void* const* push_peers = nullptr);
// slabfft.cu: peer addresses of the work buffer + acquire, for the fused inverse transpose; and its closing barrier
int push_work_begin(hymd_ctx* c, int F, void** peers, cudaStream_t s);
int push_work_end(hymd_ctx* c, cudaStream_t s);
// migrate.cu
int migrate_plan(hymd_ctx* c, const void* d_pos, int64_t n, int64_t* n_new, cudaStream_t s);
int migrate_apply(hymd_ctx* c, const void* d_in, void* d_out, int row_bytes, cudaStream_t s);
void migrate_destroy(hymd_ctx* c);
// comm.cu
int comm_unique_id(uint8_t* id);
int comm_local_group_id(int world_size, uint8_t* id);
bool comm_is_local(const hymd_ctx* c);
unsigned int* comm_status_device(hymd_ctx* c);
int comm_check_status(hymd_ctx* c);
int comm_barrier_payload(hymd_ctx* c, const uint32_t* d_payload, cudaStream_t s);
const uint32_t* comm_payload(hymd_ctx* c);
int comm_create(hymd_ctx* c, const uint8_t* id);
void comm_destroy(hymd_ctx* c);
int comm_alltoall(hymd_ctx* c, const void* send, void* recv, size_t bytes_per_peer, cudaStream_t s);
// ring exchange: send n blocks (sendp[i], bytes each) to rank+dir, receive n blocks from rank-dir
int comm_ring(hymd_ctx* c, int dir, void* const* sendp, void* const* recvp, int n, size_t bytes,
              cudaStream_t s);
int comm_alltoallv(hymd_ctx* c, const void* send, const size_t* send_off, const size_t* send_bytes,
                   void* recv, const size_t* recv_off, const size_t* recv_bytes, cudaStream_t s);
int comm_peer_ptrs(hymd_ctx* c, void* local, void** peers, cudaStream_t s);
int comm_barrier(hymd_ctx* c, cudaStream_t s);
int comm_allgather_host(hymd_ctx* c, const void* mine, void* all, size_t bytes, cudaStream_t s);
// readout.cu
int readout_setup(hymd_ctx* c);
int readout_forces(hymd_ctx* c, void* d_force, cudaStream_t s);
int readout_pme(hymd_ctx* c, void* d_force, cudaStream_t s);
int fill_ghosts(hymd_ctx* c, void* mesh, int nfields, cudaStream_t s);
// graph.cu
void graph_destroy(hymd_ctx* c);
// gpe.cu
void gpe_destroy(hymd_ctx* c);
void* gpe_field(hymd_ctx* c, int which, int t);
// energy.cu
int field_energy(hymd_ctx* c, const double* chi, double kappa, double rho0, double a,
                 double out[2], cudaStream_t s);
int field_pressure(hymd_ctx* c, const double* A, const double* cc, const double* qt, double out[4],
                   cudaStream_t s);
}  // namespace hymd
