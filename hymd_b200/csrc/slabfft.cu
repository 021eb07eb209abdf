// 3-D real transforms of the field-force cycle (RealField.r2c / ComplexField.c2r at
// field.py:576-578, 584, 613, 616, 366, 377, 397), slab-decomposed so that one code path serves
// 1..8 GPUs:
//
//   forward :  batched 2-D r2c over the (y,z) planes of the local x-slab            (cuFFT)
//              pack [f][xl][ky][kz] -> [peer][xl][f][kyl][kz] + all-to-all          (ours + NCCL)
//              1-D transform along x on the k-slab [x][f][kyl][kz]                  (cuFFT or the
//              fused x-line kernel of xline.cu, which also does the k-space math)
//   inverse :  the same steps backwards; the final 2-D c2r writes straight into the ghost-padded
//              force-mesh layout the readout kernel stages through TMA.
//
// The receive order of the all-to-all IS the k layout, so the forward transpose needs one pack
// and no unpack, the inverse one unpack and no pack.  With one GPU and no fused kernel for the
// mesh size, plain 3-D cuFFT plans are used instead.
#include "ctx.cuh"

namespace hymd {

enum PlanKind {
    PK_3D_R2C = 0, PK_3D_C2R, PK_3D_C2R_GHOST,
    PK_2D_R2C,          // real [F*vx planes]            -> [F*vx][Ny][Nzcp]
    PK_2D_C2R,          // [F*vx][Ny][Nzcp]              -> real [F*vx planes]
    PK_2D_C2R_GHOST,    // [F*(nxl+1)][Ny][Nzcp]         -> ghost [F*(nxl+1)][Ny+1][Nzp]
    PK_1D_FWD_INV       // c2c along x, stride xs, batch = columns (direction given at exec)
};

KLayout klayout(const hymd_ctx* c, int F) {
    const Geometry& g = c->g;
    KLayout l;
    if (g.P == 1) { l.xs = (long long)g.Ny * g.Nzcp; l.fs = (long long)g.Nx * l.xs; }
    else { l.fs = (long long)g.nyl * g.Nzcp; l.xs = (long long)F * l.fs; }
    return l;
}

void destroy_plans(hymd_ctx* c) {
    if (!c->plans) return;
    for (auto& p : *c->plans) cufftDestroy(p.h);
    c->plans->clear();
}

static int grow(void** p, size_t* have, size_t want) {
    if (*have >= want) return HYMD_OK;
    if (*p) { cudaDeviceSynchronize(); cudaFree(*p); *p = nullptr; *have = 0; }
    cudaError_t e = cudaMalloc(p, want);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
        return HYMD_ERR_NOMEM;
    }
    cudaMemset(*p, 0, want);
    cudaDeviceSynchronize();     // the memset runs on the legacy stream, the users on non-blocking ones
    *have = want;
    return HYMD_OK;
}

int ensure_work(hymd_ctx* c, int F) {
    const Geometry& g = c->g;
    const size_t csz = 2 * c->rsz;
    // peer-mapped buffers must never be re-allocated: size them for the largest batch (3T fields)
    if (c->p2p && F < 3 * c->T) F = 3 * c->T;
    HYMD_CHECK(grow(&c->wA, &c->wA_bytes, (size_t)F * (g.nxl + 1) * g.Ny * g.Nzcp * csz));
    if (g.P > 1 && (!c->p2p || c->xmode == 2))
        HYMD_CHECK(grow(&c->wS, &c->wS_bytes, (size_t)F * g.nxl * g.Ny * g.Nzcp * csz));
    return HYMD_OK;
}

static int plan_get(hymd_ctx* c, int kind, int F, cufftHandle* out) {
    for (auto& p : *c->plans)
        if (p.kind == kind && p.batch == F) { *out = p.h; return HYMD_OK; }
    const Geometry& g = c->g;
    cufftHandle h;
    HYMD_CUFFT(cufftCreate(&h));
    HYMD_CUFFT(cufftSetAutoAllocation(h, 0));
    size_t ws = 0;
    const cufftType r2c = c->f64 ? CUFFT_D2Z : CUFFT_R2C, c2r = c->f64 ? CUFFT_Z2D : CUFFT_C2R,
                    c2c = c->f64 ? CUFFT_Z2Z : CUFFT_C2C;
    long long n3[3] = {g.Nx, g.Ny, g.Nz}, n2[2] = {g.Ny, g.Nz}, n1[1] = {g.Nx};
    long long real3[3] = {g.Nx, g.Ny, g.Nz}, ghost3[3] = {g.Nx + 1, g.Ny + 1, g.Nzp},
              k3[3] = {g.Nx, g.Ny, g.Nzcp};
    long long real2[2] = {g.Ny, g.Nz}, ghost2[2] = {g.Ny + 1, g.Nzp}, k2[2] = {g.Ny, g.Nzcp};
    const long long plane_r = (long long)g.Ny * g.Nz, plane_k = (long long)g.Ny * g.Nzcp,
                    plane_g = (long long)(g.Ny + 1) * g.Nzp;
    switch (kind) {
        case PK_3D_R2C:
            HYMD_CUFFT(cufftMakePlanMany64(h, 3, n3, real3, 1, g.real_elems, k3, 1, g.k_elems, r2c, F, &ws));
            break;
        case PK_3D_C2R:
            HYMD_CUFFT(cufftMakePlanMany64(h, 3, n3, k3, 1, g.k_elems, real3, 1, g.real_elems, c2r, F, &ws));
            break;
        case PK_3D_C2R_GHOST:
            HYMD_CUFFT(cufftMakePlanMany64(h, 3, n3, k3, 1, g.k_elems, ghost3, 1, g.ghost_elems, c2r, F, &ws));
            break;
        case PK_2D_R2C:
            HYMD_CUFFT(cufftMakePlanMany64(h, 2, n2, real2, 1, plane_r, k2, 1, plane_k, r2c,
                                           (long long)F * g.vx, &ws));
            break;
        case PK_2D_C2R:
            HYMD_CUFFT(cufftMakePlanMany64(h, 2, n2, k2, 1, plane_k, real2, 1, plane_r, c2r,
                                           (long long)F * g.vx, &ws));
            break;
        case PK_2D_C2R_GHOST:
            HYMD_CUFFT(cufftMakePlanMany64(h, 2, n2, k2, 1, plane_k, ghost2, 1, plane_g, c2r,
                                           (long long)F * (g.nxl + 1), &ws));
            break;
        case PK_1D_FWD_INV: {
            // P == 1: one field per call (F ignored, batch = Ny*Nzcp columns); P > 1: all F fields
            const KLayout l = klayout(c, F);
            const long long batch = g.P == 1 ? l.xs : l.xs;
            long long embed[1] = {g.Nx};
            HYMD_CUFFT(cufftMakePlanMany64(h, 1, n1, embed, l.xs, 1, embed, l.xs, 1, c2c, batch, &ws));
            break;
        }
        default:
            set_error("unknown plan kind %d", kind);
            return HYMD_ERR_INVALID;
    }
    c->plans->push_back({kind, F, h});
    if (ws > c->fft_work_bytes) {
        HYMD_CHECK(grow(&c->fft_work, &c->fft_work_bytes, ws));
        for (auto& p : *c->plans) HYMD_CUFFT(cufftSetWorkArea(p.h, c->fft_work));
    } else {
        HYMD_CUFFT(cufftSetWorkArea(h, c->fft_work));
    }
    *out = h;
    return HYMD_OK;
}

static int exec_r2c(hymd_ctx* c, cufftHandle h, void* in, void* out, cudaStream_t s) {
    HYMD_CUFFT(cufftSetStream(h, s));
    if (c->f64) HYMD_CUFFT(cufftExecD2Z(h, (cufftDoubleReal*)in, (cufftDoubleComplex*)out));
    else HYMD_CUFFT(cufftExecR2C(h, (cufftReal*)in, (cufftComplex*)out));
    c->launches += 2;   // cuFFT launches >= 2 kernels per multi-dimensional transform (lower bound)
    return HYMD_OK;
}

static int exec_c2r(hymd_ctx* c, cufftHandle h, void* in, void* out, cudaStream_t s) {
    HYMD_CUFFT(cufftSetStream(h, s));
    if (c->f64) HYMD_CUFFT(cufftExecZ2D(h, (cufftDoubleComplex*)in, (cufftDoubleReal*)out));
    else HYMD_CUFFT(cufftExecC2R(h, (cufftComplex*)in, (cufftReal*)out));
    c->launches += 2;
    return HYMD_OK;
}

static int exec_c2c(hymd_ctx* c, cufftHandle h, void* data, int dir, cudaStream_t s) {
    HYMD_CUFFT(cufftSetStream(h, s));
    if (c->f64) HYMD_CUFFT(cufftExecZ2Z(h, (cufftDoubleComplex*)data, (cufftDoubleComplex*)data, dir));
    else HYMD_CUFFT(cufftExecC2C(h, (cufftComplex*)data, (cufftComplex*)data, dir));
    c->launches += 1;
    return HYMD_OK;
}

// ---- pack / unpack around the all-to-all ------------------------------------------------------
struct PackParams {
    int P, rank, nxl, nyl, Ny, Nzcp, F, x0;
    long long total;     // F*nxl*Ny*Nzcp
    long long xs, fs;    // k layout strides
};

// forward: A[f][xl'][ky][kz] (nxl+1 planes per field) -> S[q][xl][f][kyl][kz]; the block for this
// rank goes straight to its place in the k buffer.
template <typename cx>
__global__ void __launch_bounds__(256) pack_kernel(const cx* __restrict__ A, cx* __restrict__ S,
                                                   cx* __restrict__ K, PackParams p) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < p.total; o += stride) {
        long long r = o;
        const int kz = (int)(r % p.Nzcp); r /= p.Nzcp;
        const int kyl = (int)(r % p.nyl); r /= p.nyl;
        const int f = (int)(r % p.F); r /= p.F;
        const int xl = (int)(r % p.nxl);
        const int q = (int)(r / p.nxl);
        const cx v = A[(((long long)f * (p.nxl + 1) + xl) * p.Ny + (q * p.nyl + kyl)) * p.Nzcp + kz];
        if (q == p.rank) K[(long long)(p.x0 + xl) * p.xs + f * p.fs + (long long)kyl * p.Nzcp + kz] = v;
        else S[o] = v;
    }
}

// inverse: S[q][xl][f][kyl][kz] (block q received from rank q; own block read from the k buffer)
// -> A[f][xl'][ky][kz]
template <typename cx>
__global__ void __launch_bounds__(256) unpack_kernel(const cx* __restrict__ S, const cx* __restrict__ K,
                                                     cx* __restrict__ A, PackParams p) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < p.total; o += stride) {
        long long r = o;
        const int kz = (int)(r % p.Nzcp); r /= p.Nzcp;
        const int kyl = (int)(r % p.nyl); r /= p.nyl;
        const int f = (int)(r % p.F); r /= p.F;
        const int xl = (int)(r % p.nxl);
        const int q = (int)(r / p.nxl);
        const cx v = q == p.rank
                         ? K[(long long)(p.x0 + xl) * p.xs + f * p.fs + (long long)kyl * p.Nzcp + kz]
                         : S[o];
        A[(((long long)f * (p.nxl + 1) + xl) * p.Ny + (q * p.nyl + kyl)) * p.Nzcp + kz] = v;
    }
}

static PackParams make_pack(const hymd_ctx* c, int F) {
    const Geometry& g = c->g;
    PackParams p;
    p.P = g.P; p.rank = g.rank; p.nxl = g.nxl; p.nyl = g.nyl; p.Ny = g.Ny; p.Nzcp = g.Nzcp;
    p.F = F; p.x0 = g.x0;
    p.total = (long long)F * g.nxl * g.Ny * g.Nzcp;
    const KLayout l = klayout(c, F);
    p.xs = l.xs; p.fs = l.fs;
    return p;
}

static unsigned int stream_grid(long long n) {
    long long b = (n + 255) / 256;
    const long long cap = 148LL * 16;
    return (unsigned int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---- exchange over NVLink peer memory ------------------------------------------------------------
// The pack / unpack kernels store straight into the destination rank's buffer (mapped through
// CUDA IPC), so the transpose is ONE pass: read local HBM, write remote HBM over NVLink, no
// staging buffer, no separate unpack, every SM drives the links.  A barrier (comm_barrier) after
// the kernel makes the data visible to the consumers.  16-byte accesses (Nzcp is even).
struct PeerPtrs {
    void* p[HYMD_MAX_PEERS];
};

// forward: A[f][xl'][ky][kz] (nxl+1 planes per field)  ->  rank q = ky / nyl:
//          K_q[(x0 + xl)][f][kyl][kz]   (the k layout of the receiver)
template <typename vec, int CPV>
__global__ void __launch_bounds__(256) pack_push_kernel(const vec* __restrict__ A, PeerPtrs K, PackParams p) {
    const int rowv = p.Nzcp / CPV;                       // 16-byte vectors per spectrum row
    const long long total = p.total / CPV, stride = (long long)gridDim.x * blockDim.x;
    for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total; o += stride) {
        long long r = o;
        const int kv = (int)(r % rowv); r /= rowv;
        const int ky = (int)(r % p.Ny); r /= p.Ny;
        const int xl = (int)(r % p.nxl);
        const int f = (int)(r / p.nxl);
        const int q = ky / p.nyl, kyl = ky % p.nyl;
        const vec v = A[(((long long)f * (p.nxl + 1) + xl) * p.Ny + ky) * rowv + kv];
        vec* dst = reinterpret_cast<vec*>(K.p[q]);
        dst[((long long)(p.x0 + xl) * p.xs + f * p.fs + (long long)kyl * p.Nzcp) / CPV + kv] = v;
    }
}

// inverse: local K[x][f][kyl][kz]  ->  rank q = x / nxl:  A_q[f][xl][ky = rank*nyl + kyl][kz]
template <typename vec, int CPV>
__global__ void __launch_bounds__(256) unpack_push_kernel(const vec* __restrict__ Kl, PeerPtrs A, PackParams p) {
    const int rowv = p.Nzcp / CPV;
    const long long total = p.total / CPV, stride = (long long)gridDim.x * blockDim.x;
    for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total; o += stride) {
        long long r = o;
        const int kv = (int)(r % rowv); r /= rowv;
        const int kyl = (int)(r % p.nyl); r /= p.nyl;
        const int f = (int)(r % p.F);
        const int x = (int)(r / p.F);
        const int q = x / p.nxl, xl = x % p.nxl;
        const vec v = Kl[((long long)x * p.xs + f * p.fs + (long long)kyl * p.Nzcp) / CPV + kv];
        vec* dst = reinterpret_cast<vec*>(A.p[q]);
        dst[(((long long)f * (p.nxl + 1) + xl) * p.Ny + (p.rank * p.nyl + kyl)) * rowv + kv] = v;
    }
}

// Blocked exchange with SM copies (HYMD_B200_EXCHANGE=blockedk): block q of `src` (contiguous, `block16` 16-byte
// words) goes to dst.p[q] + dst_off16; every SM streams to every peer at once, 16 bytes per lane, four loads in
// flight per thread (the pattern that reaches 700+ GB/s per direction in tools/microbench/p2p.cu).
__global__ void __launch_bounds__(256) block_copy_kernel(const uint4* __restrict__ src, PeerPtrs dst, long long block16,
                                                         long long dst_off16, int P, int rank, int include_self) {
    // blockIdx.y = which peer (all peers are written at the same time, each by gridDim.x CTAs)
    const int q = (rank + (int)blockIdx.y + (include_self ? 0 : 1)) % P;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const uint4* s = src + (long long)q * block16;
    uint4* d = reinterpret_cast<uint4*>(dst.p[q]) + dst_off16;
    long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (; j + 3 * stride < block16; j += 4 * stride) {
        const uint4 a = __ldcs(s + j), b = __ldcs(s + j + stride), c2 = __ldcs(s + j + 2 * stride), e = __ldcs(s + j + 3 * stride);
        d[j] = a; d[j + stride] = b; d[j + 2 * stride] = c2; d[j + 3 * stride] = e;
    }
    for (; j < block16; j += stride) d[j] = __ldcs(s + j);
}

// grid of the block copy: ~148 x 8 CTAs in total, split over the peers; CTAs per peer bounded by the block size
static dim3 block_copy_grid(long long block16, int npeers) {
    long long per = (148LL * 8 + npeers - 1) / npeers;
    const long long need = (block16 + 4 * 256 - 1) / (4 * 256);
    if (per > need) per = need;
    if (per < 1) per = 1;
    return dim3((unsigned)per, (unsigned)npeers, 1);
}

// Pipelined exchange: the part of every per-destination block that belongs to ONE field (forward) or one potential
// row (inverse) -- inside block q that is `nseg` segments (one per local x plane) of seg16 16-byte words, stride16
// apart, starting off16 into the block.  Same roles of src / dst / include_self as block_copy_kernel.  A small fixed
// grid (16 SMs' worth): it runs beside the plane kernel that transforms the next field.
__global__ void __launch_bounds__(256) block_copy2d_kernel(const uint4* __restrict__ src, PeerPtrs dst, long long block16,
                                                           long long dst_off16, long long off16, unsigned int seg16,
                                                           long long stride16, unsigned int nseg, int P, int rank,
                                                           int include_self) {
    const int q = (rank + (int)blockIdx.y + (include_self ? 0 : 1)) % P;
    const uint4* s = src + (long long)q * block16 + off16;
    uint4* d = reinterpret_cast<uint4*>(dst.p[q]) + dst_off16 + off16;
    const unsigned int total = seg16 * nseg, stride = gridDim.x * blockDim.x;
    auto at = [&](unsigned int w) { return (long long)(w / seg16) * stride16 + (w % seg16); };
    unsigned int w = blockIdx.x * blockDim.x + threadIdx.x;
    for (; w + 3ull * stride < total; w += 4 * stride) {
        const long long o0 = at(w), o1 = at(w + stride), o2 = at(w + 2 * stride), o3 = at(w + 3 * stride);
        const uint4 a = __ldcs(s + o0), b = __ldcs(s + o1), c2 = __ldcs(s + o2), e = __ldcs(s + o3);
        d[o0] = a; d[o1] = b; d[o2] = c2; d[o3] = e;
    }
    for (; w < total; w += stride) { const long long o = at(w); d[o] = __ldcs(s + o); }
}

// Pipeline plan: how many fields (potential rows) make one piece; 0 = no pipeline (one piece).  A piece is one launch
// of the persistent plane kernel, so one field must bring about as many (field, plane) units as there are SMs:
//   2 ranks, C4 (128 planes per field on 148 SMs): 1.364 -> 1.264 ms per cycle with per-field pieces (profiles/r3e_*)
//   4 ranks, C4 ( 64 planes): two fields per piece 0.907 ms against 0.899 ms without the pipeline (profiles/r3k_*)
//   8 ranks, C5 ( 64 planes): per-field pieces 5.09 ms against 3.59 ms without (profiles/r3i_*)
// so the pipeline is used when a slab has >= 0.75 x SMs planes (and a piece is >= 4 MB per destination: below that the
// exchange is latency, not bandwidth).  HYMD_B200_XPIPE=2 forces it, HYMD_B200_XPIPE_GROUP sets the piece (tests).
static int pipe_group(const hymd_ctx* c, int nfields, long long field_bytes) {
    if (c->xpipe == 0 || nfields < 2 || nfields > HYMD_MAX_TYPES || c->xstream == nullptr || field_bytes % 16 != 0) return 0;
    int group = 1;
    if (c->xpipe != 2) {
        const int sms = c->sm_count > 0 ? c->sm_count : 148;
        if (4 * c->g.nxl < 3 * sms || field_bytes < (4LL << 20)) return 0;
    }
    if (const char* e = getenv("HYMD_B200_XPIPE_GROUP")) group = atoi(e) > 0 ? atoi(e) : group;     // tests
    return group < nfields ? group : 0;
}

// while alive, the persistent plane kernels keep 16 SMs free for the copy kernel of the second stream
struct SmReserve {
    hymd_ctx* c;
    explicit SmReserve(hymd_ctx* ctx) : c(ctx) { c->plane_sm_reserve = 16; }
    ~SmReserve() { c->plane_sm_reserve = 0; }
};

static int launch_copy2d(hymd_ctx* c, const void* src, const PeerPtrs& dst, long long block16, long long dst_off16,
                         long long off16, long long seg16, long long stride16, int nseg, int npeers, int include_self,
                         cudaStream_t s) {
    if (seg16 * nseg >= (1LL << 32)) { set_error("pipelined exchange: piece too large"); return HYMD_ERR_INVALID; }
    long long per = 128 / npeers;                       // ~128 CTAs in total
    const long long need = (seg16 * nseg + 4 * 256 - 1) / (4 * 256);
    if (per > need) per = need;
    if (per < 1) per = 1;
    block_copy2d_kernel<<<dim3((unsigned)per, (unsigned)npeers, 1), 256, 0, s>>>(
        (const uint4*)src, dst, block16, dst_off16, off16, (unsigned int)seg16, stride16, (unsigned int)nseg,
        c->g.P, c->g.rank, include_self);
    HYMD_LAUNCH_CHECK(c);
    return HYMD_OK;
}

static int peer_table(hymd_ctx* c, void* local, PeerPtrs* t, cudaStream_t s) {
    memset(t, 0, sizeof(*t));
    return comm_peer_ptrs(c, local, t->p, s);
}

// Before storing into a peer's buffer: its previous contents must have been consumed there.
// Normally another barrier of the cycle has passed since (peer_busy cleared); back-to-back
// transforms (by-product materialisation) need an extra one.
static int peer_acquire(hymd_ctx* c, unsigned which, cudaStream_t s) {
    if (c->peer_busy & which) return comm_barrier(c, s);
    return HYMD_OK;
}

static int transpose_forward(hymd_ctx* c, int F, void* k_out, cudaStream_t s) {
    const PackParams p = make_pack(c, F);
    if (c->p2p) {
        PhaseScope ps(c, HYMD_PHASE_ALLTOALL, s);
        PeerPtrs K;
        HYMD_CHECK(peer_table(c, k_out, &K, s));
        HYMD_CHECK(peer_acquire(c, PEER_K, s));
        const unsigned int grid = stream_grid(p.total / 2);
        if (c->f64) pack_push_kernel<double2, 1><<<grid, 256, 0, s>>>((const double2*)c->wA, K, p);
        else pack_push_kernel<float4, 2><<<grid, 256, 0, s>>>((const float4*)c->wA, K, p);
        HYMD_LAUNCH_CHECK(c);
        HYMD_CHECK(comm_barrier(c, s));
        c->peer_busy |= PEER_K;
        return HYMD_OK;
    }
    if (c->f64) pack_kernel<double2><<<stream_grid(p.total), 256, 0, s>>>(
        (const double2*)c->wA, (double2*)c->wS, (double2*)k_out, p);
    else pack_kernel<float2><<<stream_grid(p.total), 256, 0, s>>>(
        (const float2*)c->wA, (float2*)c->wS, (float2*)k_out, p);
    HYMD_LAUNCH_CHECK(c);
    const size_t block = (size_t)p.total / p.P * 2 * c->rsz;
    // block q of wS -> rank q; block p of the k buffer (x in slab p) <- rank p
    return comm_alltoall(c, c->wS, k_out, block, s);
}

static int transpose_inverse(hymd_ctx* c, int F, void* k_in, cudaStream_t s) {
    const PackParams p = make_pack(c, F);
    if (c->p2p) {
        PhaseScope ps(c, HYMD_PHASE_ALLTOALL, s);
        PeerPtrs A;
        HYMD_CHECK(peer_table(c, c->wA, &A, s));
        HYMD_CHECK(peer_acquire(c, PEER_WORK, s));
        const unsigned int grid = stream_grid(p.total / 2);
        if (c->f64) unpack_push_kernel<double2, 1><<<grid, 256, 0, s>>>((const double2*)k_in, A, p);
        else unpack_push_kernel<float4, 2><<<grid, 256, 0, s>>>((const float4*)k_in, A, p);
        HYMD_LAUNCH_CHECK(c);
        HYMD_CHECK(comm_barrier(c, s));
        c->peer_busy |= PEER_WORK;
        return HYMD_OK;
    }
    const size_t block = (size_t)p.total / p.P * 2 * c->rsz;
    HYMD_CHECK(comm_alltoall(c, k_in, c->wS, block, s));
    if (c->f64) unpack_kernel<double2><<<stream_grid(p.total), 256, 0, s>>>(
        (const double2*)c->wS, (const double2*)k_in, (double2*)c->wA, p);
    else unpack_kernel<float2><<<stream_grid(p.total), 256, 0, s>>>(
        (const float2*)c->wS, (const float2*)k_in, (float2*)c->wA, p);
    HYMD_LAUNCH_CHECK(c);
    return HYMD_OK;
}

// Fused inverse transpose (xline.cu stores into the peers' work buffers): peer addresses + acquire before
// the launch, barrier after it.
int push_work_begin(hymd_ctx* c, int F, void** peers, cudaStream_t s) {
    HYMD_CHECK(ensure_work(c, F));
    PeerPtrs A;
    HYMD_CHECK(peer_table(c, c->wA, &A, s));
    HYMD_CHECK(peer_acquire(c, PEER_WORK, s));
    for (int q = 0; q < HYMD_MAX_PEERS; ++q) peers[q] = A.p[q];
    return HYMD_OK;
}

int push_work_end(hymd_ctx* c, cudaStream_t s) {
    PhaseScope ps(c, HYMD_PHASE_ALLTOALL, s);
    HYMD_CHECK(comm_barrier(c, s));
    c->peer_busy |= PEER_WORK;
    c->xpushed = true;
    return HYMD_OK;
}

// 2-D (y,z) transforms of the local planes: the one-pass plane kernels (planefft.cu) when the
// plane size is supported, batched cuFFT 2-D plans otherwise.
static int yz_forward(hymd_ctx* c, void* real_in, int F, void* k_out, long long k_fs, cudaStream_t s) {
    const Geometry& g = c->g;
    if (c->plane) return plane_forward(c, real_in, g.real_elems, F, g.nxl, k_out, k_fs, s);
    cufftHandle h;
    HYMD_CHECK(plan_get(c, PK_2D_R2C, F, &h));
    return exec_r2c(c, h, real_in, k_out, s);
}

static int yz_inverse(hymd_ctx* c, void* k_in, long long k_fs, int F, void* real_out, bool ghost,
                      cudaStream_t s, bool derive = false) {
    const Geometry& g = c->g;
    if (c->plane) return plane_inverse(c, k_in, k_fs, F, g.nxl, real_out, ghost, derive, s);
    if (derive) { set_error("derived force components need the plane kernels"); return HYMD_ERR_STATE; }
    cufftHandle h;
    HYMD_CHECK(plan_get(c, ghost ? PK_2D_C2R_GHOST : PK_2D_C2R, F, &h));
    return exec_c2r(c, h, k_in, real_out, s);
}

// ---- public transforms ------------------------------------------------------------------------
int fft_forward(hymd_ctx* c, void* real_in, int F, void* k_out, cudaStream_t s) {
    const Geometry& g = c->g;
    cufftHandle h;
    if (!c->slab) {
        HYMD_CHECK(plan_get(c, PK_3D_R2C, F, &h));
        return exec_r2c(c, h, real_in, k_out, s);
    }
    const size_t csz = 2 * c->rsz;
    const KLayout l = klayout(c, F);
    HYMD_CHECK(fft_forward_yz(c, real_in, F, k_out, s));
    if (g.P == 1) {
        HYMD_CHECK(plan_get(c, PK_1D_FWD_INV, 1, &h));
        for (int f = 0; f < F; ++f)
            HYMD_CHECK(exec_c2c(c, h, (char*)k_out + (size_t)f * l.fs * csz, CUFFT_FORWARD, s));
        return HYMD_OK;
    }
    HYMD_CHECK(plan_get(c, PK_1D_FWD_INV, F, &h));
    return exec_c2c(c, h, k_out, CUFFT_FORWARD, s);
}

// Forward transform over y and z only (the fused x-line kernel does the rest): k_out receives
// the k layout with x still in real space.
int fft_forward_yz(hymd_ctx* c, void* real_in, int F, void* k_out, cudaStream_t s) {
    const Geometry& g = c->g;
    if (g.P == 1) return yz_forward(c, real_in, F, k_out, klayout(c, F).fs, s);
    HYMD_CHECK(ensure_work(c, F));
    if (c->xmode == 2) {
        // blocked exchange: the plane kernel stores every spectrum row into the staging block of the rank that
        // owns its k_y range (own rows straight into the local k buffer); a block is exactly the contiguous
        // range [x0, x0 + nxl) of the receiver's k buffer, so it crosses NVLink as ONE contiguous copy
        const size_t csz = 2 * c->rsz;
        const KLayout l = klayout(c, F);
        const long long block = (long long)g.nxl * l.xs;               // elements per destination
        PeerPtrs K, T;
        HYMD_CHECK(peer_table(c, k_out, &K, s));
        HYMD_CHECK(peer_acquire(c, PEER_K, s));
        for (int q = 0; q < g.P; ++q)
            T.p[q] = q == g.rank ? k_out : (char*)c->wS + ((long long)q * block - (long long)g.x0 * l.xs) * (long long)csz;
        const int grp = (l.fs * csz) % 16 == 0 ? pipe_group(c, F, (long long)g.nxl * l.fs * (long long)csz) : 0;
        if (grp > 0) {
            // fields f .. f + cnt - 1 cross NVLink while the plane kernel transforms the next group
            int piece = 0;
            if (c->xpipe_ce) {
                // copy engines, one stream per destination: the SMs stay with the plane kernel (an SM copy kernel
                // squeezed onto the 16-20 SMs the plane kernel leaves free reached a third of the NVLink rate)
                for (int f = 0; f < F; f += grp, ++piece) {
                    const int cnt = F - f < grp ? F - f : grp;
                    HYMD_CHECK(plane_forward(c, real_in, g.real_elems, F, g.nxl, nullptr, 0, s, T.p, f, cnt));
                    HYMD_CUDA(cudaEventRecord(c->xev[piece], s));
                    for (int i = 1; i < g.P; ++i) {
                        const int q = (g.rank + i) % g.P;
                        HYMD_CUDA(cudaStreamWaitEvent(c->xpeer[q], c->xev[piece], 0));
                        HYMD_CUDA(cudaMemcpy2DAsync((char*)K.p[q] + ((size_t)g.x0 * l.xs + (size_t)f * l.fs) * csz, (size_t)l.xs * csz,
                                                    (char*)c->wS + ((size_t)q * block + (size_t)f * l.fs) * csz, (size_t)l.xs * csz,
                                                    (size_t)cnt * l.fs * csz, (size_t)g.nxl, cudaMemcpyDeviceToDevice, c->xpeer[q]));
                    }
                    c->launches += g.P - 1;
                }
                PhaseScope ps(c, HYMD_PHASE_ALLTOALL, s);
                for (int i = 1; i < g.P; ++i) {
                    const int q = (g.rank + i) % g.P;
                    HYMD_CUDA(cudaEventRecord(c->xdone[q][0], c->xpeer[q]));
                    HYMD_CUDA(cudaStreamWaitEvent(s, c->xdone[q][0], 0));
                }
                HYMD_CHECK(comm_barrier(c, s));
                c->peer_busy |= PEER_K;
                return HYMD_OK;
            }
            SmReserve reserve(c);
            for (int f = 0; f < F; f += grp, ++piece) {
                const int cnt = F - f < grp ? F - f : grp;
                HYMD_CHECK(plane_forward(c, real_in, g.real_elems, F, g.nxl, nullptr, 0, s, T.p, f, cnt));
                HYMD_CUDA(cudaEventRecord(c->xev[piece], s));
                HYMD_CUDA(cudaStreamWaitEvent(c->xstream, c->xev[piece], 0));
                HYMD_CHECK(launch_copy2d(c, c->wS, K, (long long)(block * csz / 16), (long long)((size_t)g.x0 * l.xs * csz / 16),
                                         (long long)((size_t)f * l.fs * csz / 16), (long long)((size_t)cnt * l.fs * csz / 16),
                                         (long long)(l.xs * csz / 16), g.nxl, g.P - 1, 0, c->xstream));
            }
            PhaseScope ps(c, HYMD_PHASE_ALLTOALL, s);
            HYMD_CUDA(cudaEventRecord(c->xev[piece], c->xstream));
            HYMD_CUDA(cudaStreamWaitEvent(s, c->xev[piece], 0));
            HYMD_CHECK(comm_barrier(c, s));
            c->peer_busy |= PEER_K;
            return HYMD_OK;
        }
        HYMD_CHECK(plane_forward(c, real_in, g.real_elems, F, g.nxl, nullptr, 0, s, T.p));
        PhaseScope ps(c, HYMD_PHASE_ALLTOALL, s);
        if (c->xcopy_kernel && (block * csz) % 16 == 0) {
            block_copy_kernel<<<block_copy_grid(block * csz / 16, g.P - 1), 256, 0, s>>>(
                (const uint4*)c->wS, K, (long long)(block * csz / 16), (long long)((size_t)g.x0 * l.xs * csz / 16),
                g.P, g.rank, 0);
            HYMD_LAUNCH_CHECK(c);
        } else {
            for (int i = 1; i < g.P; ++i) {
                const int q = (g.rank + i) % g.P;
                HYMD_CUDA(cudaMemcpyAsync((char*)K.p[q] + (size_t)g.x0 * l.xs * csz, (char*)c->wS + (size_t)q * block * csz,
                                          (size_t)block * csz, cudaMemcpyDeviceToDevice, s));
            }
            c->launches += g.P - 1;
        }
        HYMD_CHECK(comm_barrier(c, s));
        c->peer_busy |= PEER_K;
        return HYMD_OK;
    }
    if (c->fused_push && c->plane) {
        // the forward transpose is the plane kernel's own epilogue: every spectrum row is stored into the
        // k buffer of the rank owning its k_y range while the next rows are still being transformed
        PeerPtrs K;
        HYMD_CHECK(peer_table(c, k_out, &K, s));
        HYMD_CHECK(peer_acquire(c, PEER_K, s));
        HYMD_CHECK(plane_forward(c, real_in, g.real_elems, F, g.nxl, nullptr, 0, s, K.p));
        PhaseScope ps(c, HYMD_PHASE_ALLTOALL, s);
        HYMD_CHECK(comm_barrier(c, s));
        c->peer_busy |= PEER_K;
        return HYMD_OK;
    }
    HYMD_CHECK(yz_forward(c, real_in, F, c->wA, (long long)(g.nxl + 1) * g.Ny * g.Nzcp, s));
    return transpose_forward(c, F, k_out, s);
}

// Copies [f][Nx][Ny][Nzcp] into the [f][Nx+1][Ny][Nzcp] work layout (single GPU, ghost output).
static int copy_to_work(hymd_ctx* c, const void* k_in, int F, cudaStream_t s) {
    const Geometry& g = c->g;
    const size_t row = (size_t)g.Nx * g.Ny * g.Nzcp * 2 * c->rsz;
    const size_t pitch = (size_t)(g.Nx + 1) * g.Ny * g.Nzcp * 2 * c->rsz;
    HYMD_CUDA(cudaMemcpy2DAsync(c->wA, pitch, k_in, row, row, F, cudaMemcpyDeviceToDevice, s));
    return HYMD_OK;
}

int fft_inverse_xdone(hymd_ctx* c, void* k_in, int F, void* real_out, bool ghost, cudaStream_t s,
                      bool derive) {
    // derive: k_in holds 2F/3 spectra (xline.cu, XParams::two); only those cross the transpose
    // the x transform has been applied.  P > 1: k_in is in the k layout and goes through the
    // inverse transpose into the work layout.  P == 1: k_in is the work layout [f][Nx+1][Ny][Nzcp]
    // for ghost outputs and the k layout [f][Nx][Ny][Nzcp] otherwise.
    const Geometry& g = c->g;
    const long long plane = (long long)g.Ny * g.Nzcp;
    const int Fin = derive ? F / 3 * 2 : F;
    if (g.P > 1) {
        HYMD_CHECK(ensure_work(c, Fin));
        if (c->xmode == 2) {
            // blocked exchange: the x-range of rank q is one contiguous block of the k layout; it is copied as
            // it is into block `rank` of q's work buffer, and the plane kernel reads W[q][x][f][kyl][kz]
            const size_t csz = 2 * c->rsz;
            const KLayout lk = klayout(c, Fin);
            const long long block = (long long)g.nxl * lk.xs;
            PeerPtrs W;
            HYMD_CHECK(peer_table(c, c->wA, &W, s));
            const int U = derive ? F / 3 : 0;
            const int grp = (derive && (lk.fs * csz) % 16 == 0) ? pipe_group(c, U, 2LL * g.nxl * lk.fs * (long long)csz) : 0;
            if (grp > 0) {
                // the spectra of potential rows u .. u + cnt - 1 cross NVLink while the plane kernel turns the previous
                // group into its force meshes; one barrier per group tells every rank that the group has landed
                const int npiece = (U + grp - 1) / grp;
                if (c->xpipe_ce) {
                    {
                        PhaseScope ps(c, HYMD_PHASE_ALLTOALL, s);
                        HYMD_CHECK(peer_acquire(c, PEER_WORK, s));
                        HYMD_CUDA(cudaEventRecord(c->xev[0], s));               // the x-line kernel has written k_in
                        for (int i = 0; i < g.P; ++i) {
                            const int q = (g.rank + i) % g.P;
                            HYMD_CUDA(cudaStreamWaitEvent(c->xpeer[q], c->xev[0], 0));
                            for (int pc = 0; pc < npiece; ++pc) {
                                const int u = pc * grp, cnt = U - u < grp ? U - u : grp;
                                HYMD_CUDA(cudaMemcpy2DAsync((char*)W.p[q] + ((size_t)g.rank * block + (size_t)2 * u * lk.fs) * csz, (size_t)lk.xs * csz,
                                                            (char*)k_in + ((size_t)q * block + (size_t)2 * u * lk.fs) * csz, (size_t)lk.xs * csz,
                                                            (size_t)2 * cnt * lk.fs * csz, (size_t)g.nxl, cudaMemcpyDeviceToDevice, c->xpeer[q]));
                                HYMD_CUDA(cudaEventRecord(c->xdone[q][pc], c->xpeer[q]));
                            }
                        }
                        c->launches += g.P * npiece;
                    }
                    for (int pc = 0; pc < npiece; ++pc) {
                        const int u = pc * grp, cnt = U - u < grp ? U - u : grp;
                        {
                            PhaseScope ps(c, HYMD_PHASE_ALLTOALL, s);
                            for (int q = 0; q < g.P; ++q) HYMD_CUDA(cudaStreamWaitEvent(s, c->xdone[q][pc], 0));
                            HYMD_CHECK(comm_barrier(c, s));
                        }
                        HYMD_CHECK(plane_inverse(c, c->wA, 0, F, g.nxl, real_out, ghost, derive, s, true, u, cnt));
                    }
                    c->peer_busy |= PEER_WORK;
                    return HYMD_OK;
                }
                {
                    PhaseScope ps(c, HYMD_PHASE_ALLTOALL, s);
                    HYMD_CHECK(peer_acquire(c, PEER_WORK, s));
                    HYMD_CUDA(cudaEventRecord(c->xev[0], s));                   // the x-line kernel has written k_in
                    HYMD_CUDA(cudaStreamWaitEvent(c->xstream, c->xev[0], 0));
                    for (int pc = 0; pc < npiece; ++pc) {
                        const int u = pc * grp, cnt = U - u < grp ? U - u : grp;
                        HYMD_CHECK(launch_copy2d(c, k_in, W, (long long)(block * csz / 16),
                                                 (long long)((size_t)g.rank * block * csz / 16),
                                                 (long long)((size_t)2 * u * lk.fs * csz / 16), (long long)((size_t)2 * cnt * lk.fs * csz / 16),
                                                 (long long)(lk.xs * csz / 16), g.nxl, g.P, 1, c->xstream));
                        HYMD_CUDA(cudaEventRecord(c->xev[1 + pc], c->xstream));
                    }
                }
                SmReserve reserve(c);
                for (int pc = 0; pc < npiece; ++pc) {
                    const int u = pc * grp, cnt = U - u < grp ? U - u : grp;
                    {
                        PhaseScope ps(c, HYMD_PHASE_ALLTOALL, s);
                        HYMD_CUDA(cudaStreamWaitEvent(s, c->xev[1 + pc], 0));
                        HYMD_CHECK(comm_barrier(c, s));
                    }
                    HYMD_CHECK(plane_inverse(c, c->wA, 0, F, g.nxl, real_out, ghost, derive, s, true, u, cnt));
                }
                c->peer_busy |= PEER_WORK;
                return HYMD_OK;
            }
            PhaseScope ps(c, HYMD_PHASE_ALLTOALL, s);
            HYMD_CHECK(peer_acquire(c, PEER_WORK, s));
            if (c->xcopy_kernel && (block * csz) % 16 == 0) {
                block_copy_kernel<<<block_copy_grid(block * csz / 16, g.P), 256, 0, s>>>(
                    (const uint4*)k_in, W, (long long)(block * csz / 16), (long long)((size_t)g.rank * block * csz / 16),
                    g.P, g.rank, 1);
                HYMD_LAUNCH_CHECK(c);
            } else {
                for (int i = 0; i < g.P; ++i) {
                    const int q = (g.rank + i) % g.P;
                    HYMD_CUDA(cudaMemcpyAsync((char*)W.p[q] + (size_t)g.rank * block * csz, (char*)k_in + (size_t)q * block * csz,
                                              (size_t)block * csz, cudaMemcpyDeviceToDevice, s));
                }
                c->launches += g.P;
            }
            HYMD_CHECK(comm_barrier(c, s));
            c->peer_busy |= PEER_WORK;
            return plane_inverse(c, c->wA, 0, F, g.nxl, real_out, ghost, derive, s, true);
        }
        if (c->xpushed) c->xpushed = false;     // the x-line kernel stored into the peers' work buffers itself
        else HYMD_CHECK(transpose_inverse(c, Fin, k_in, s));
        return yz_inverse(c, c->wA, (g.nxl + 1) * plane, F, real_out, ghost, s, derive);
    }
    return yz_inverse(c, k_in, (ghost ? g.Nx + 1 : g.Nx) * plane, F, real_out, ghost, s, derive);
}

int fft_inverse(hymd_ctx* c, void* k_in, int F, void* real_out, bool ghost, cudaStream_t s) {
    const Geometry& g = c->g;
    cufftHandle h;
    if (!c->slab) {
        HYMD_CHECK(plan_get(c, ghost ? PK_3D_C2R_GHOST : PK_3D_C2R, F, &h));
        return exec_c2r(c, h, k_in, real_out, s);
    }
    const size_t csz = 2 * c->rsz;
    const KLayout l = klayout(c, F);
    if (g.P == 1) {
        HYMD_CHECK(plan_get(c, PK_1D_FWD_INV, 1, &h));
        for (int f = 0; f < F; ++f)
            HYMD_CHECK(exec_c2c(c, h, (char*)k_in + (size_t)f * l.fs * csz, CUFFT_INVERSE, s));
        if (!ghost) return fft_inverse_xdone(c, k_in, F, real_out, false, s);
        HYMD_CHECK(ensure_work(c, F));
        HYMD_CHECK(copy_to_work(c, k_in, F, s));
        return fft_inverse_xdone(c, c->wA, F, real_out, true, s);
    }
    HYMD_CHECK(plan_get(c, PK_1D_FWD_INV, F, &h));
    HYMD_CHECK(exec_c2c(c, h, k_in, CUFFT_INVERSE, s));
    return fft_inverse_xdone(c, k_in, F, real_out, ghost, s);
}

// ---- ghost-plane exchange between neighbouring slabs ------------------------------------------
template <typename real>
__global__ void __launch_bounds__(256) halo_add_kernel(real* __restrict__ fields,
                                                       const real* __restrict__ halo, int F,
                                                       long long plane, long long field_stride) {
    const long long total = plane * F, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long f = i / plane, j = i % plane;
        fields[f * field_stride + j] += halo[i];
    }
}

// paint: plane nxl of every field (the ghost plane) is added into plane 0 of the next slab
// (the Layout.exchange ghost copies of pm.paint, field.py:574).
int halo_reduce(hymd_ctx* c, void* fields, int F, cudaStream_t s) {
    const Geometry& g = c->g;
    if (g.P == 1) return HYMD_OK;
    PhaseScope ps(c, HYMD_PHASE_HALO, s);
    const long long plane = (long long)g.Ny * g.Nz;
    // sized once for T fields: the buffer is mapped into the neighbour
    HYMD_CHECK(grow(&c->halo, &c->halo_bytes, (size_t)(c->T > F ? c->T : F) * plane * c->rsz));
    if (c->p2p) {
        PeerPtrs H;
        HYMD_CHECK(peer_table(c, c->halo, &H, s));
        HYMD_CHECK(peer_acquire(c, PEER_HALO, s));
        const int to = (g.rank + 1) % g.P;
        // ghost plane nxl of every field -> the next slab's staging buffer, over NVLink
        HYMD_CUDA(cudaMemcpy2DAsync(H.p[to], (size_t)plane * c->rsz,
                                    (char*)fields + (size_t)g.nxl * plane * c->rsz,
                                    (size_t)g.real_elems * c->rsz, (size_t)plane * c->rsz, F,
                                    cudaMemcpyDeviceToDevice, s));
        HYMD_CHECK(comm_barrier(c, s));
        c->peer_busy |= PEER_HALO;
    } else {
        void* sp[HYMD_MAX_TYPES];
        void* rp[HYMD_MAX_TYPES];
        for (int f = 0; f < F; ++f) {
            sp[f] = (char*)fields + ((size_t)f * g.real_elems + (size_t)g.nxl * plane) * c->rsz;
            rp[f] = (char*)c->halo + (size_t)f * plane * c->rsz;
        }
        HYMD_CHECK(comm_ring(c, +1, sp, rp, F, (size_t)plane * c->rsz, s));
    }
    if (c->f64) halo_add_kernel<double><<<stream_grid(plane * F), 256, 0, s>>>(
        (double*)fields, (const double*)c->halo, F, plane, g.real_elems);
    else halo_add_kernel<float><<<stream_grid(plane * F), 256, 0, s>>>(
        (float*)fields, (const float*)c->halo, F, plane, g.real_elems);
    HYMD_LAUNCH_CHECK(c);
    return HYMD_OK;
}

// readout: plane nxl of every ghost-padded mesh <- plane 0 (with its y/z ghosts) of the next slab
int halo_fetch(hymd_ctx* c, void* meshes, int F, cudaStream_t s) {
    const Geometry& g = c->g;
    if (g.P == 1) return HYMD_OK;
    const size_t plane = (size_t)(g.Ny + 1) * g.Nzp * c->rsz;
    if (c->p2p) {
        // push: my plane 0 is the ghost plane of the previous slab
        PeerPtrs M;
        HYMD_CHECK(peer_table(c, meshes, &M, s));
        HYMD_CHECK(peer_acquire(c, PEER_MESH, s));
        // cuFFT path: the receiver's own batched 2-D c2r writes all nxl + 1 planes of its ghost meshes (the
        // last one from a stale work plane), so nobody may store into a neighbour's plane nxl before every
        // rank has finished its transform.  (The plane kernels write planes 0 .. nxl-1 only.)
        if (!c->plane) HYMD_CHECK(comm_barrier(c, s));
        const int to = (g.rank - 1 + g.P) % g.P;
        HYMD_CUDA(cudaMemcpy2DAsync((char*)M.p[to] + (size_t)g.nxl * plane, (size_t)g.ghost_elems * c->rsz,
                                    meshes, (size_t)g.ghost_elems * c->rsz, plane, F,
                                    cudaMemcpyDeviceToDevice, s));
        HYMD_CHECK(comm_barrier(c, s));
        c->peer_busy |= PEER_MESH;
        return HYMD_OK;
    }
    void* sp[3 * HYMD_MAX_TYPES];
    void* rp[3 * HYMD_MAX_TYPES];
    for (int f = 0; f < F; ++f) {
        sp[f] = (char*)meshes + (size_t)f * g.ghost_elems * c->rsz;
        rp[f] = (char*)sp[f] + (size_t)g.nxl * plane;
    }
    return comm_ring(c, -1, sp, rp, F, plane, s);
}

}  // namespace hymd
