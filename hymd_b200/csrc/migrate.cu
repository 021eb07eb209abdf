// GPU-side particle exchange between slabs: replaces pmesh's Layout.exchange as used by
// domain_decomposition (field.py:1165-1178, main.py:1171-1201).
//
//   plan  : destination slab of every local particle from its (wrapped) x cell, a stable
//           grouping of the particle indices by destination (one radix pass over log2 P bits),
//           and an all-gather of the P x P count matrix so every rank knows what it receives.
//   apply : for one per-particle array, gather the rows that stay (in their original relative
//           order) to the front of the output, the rows that leave into a send buffer grouped by
//           destination, all-to-all-v them over NCCL, arrivals appended in source-rank order.
//
// The new local order is a deterministic function of the inputs (stable grouping, fixed
// source-rank order), and every array passed to apply() is permuted identically.
#include <cub/device/device_radix_sort.cuh>

#include "ctx.cuh"

namespace hymd {

struct MigrateState {
    int64_t n = 0, n_new = 0, n_stay = 0, cap = 0;
    int32_t *dest = nullptr, *dest_sorted = nullptr, *idx = nullptr, *perm = nullptr;
    void* sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    unsigned int* d_counts = nullptr;       // P counters (this rank's sends)
    unsigned int* d_all = nullptr;          // P x P matrix, row r = sends of rank r
    unsigned int* h_all = nullptr;          // pinned copy
    void* sendbuf = nullptr;
    size_t sendbuf_bytes = 0;
    size_t send_rows_off[9], recv_rows_off[9];
    size_t send_rows[8], recv_rows[8];
    bool planned = false;
};

template <typename real>
__global__ void __launch_bounds__(256) dest_kernel(const real* __restrict__ pos, long long n,
                                                   double sx, int Nx, int nxl, int32_t* __restrict__ dest,
                                                   int32_t* __restrict__ idx,
                                                   unsigned int* __restrict__ counts) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = (double)pos[3 * i] * sx;
    double f = floor(x);
    long long c = (long long)f % Nx;
    if (c < 0) c += Nx;
    if (x - f >= 1.0) c = (c + 1 == Nx) ? 0 : c + 1;   // same rounding rule as the cell binning
    const int d = (int)(c / nxl);
    dest[i] = d;
    idx[i] = (int32_t)i;
    atomicAdd(&counts[d], 1u);
}

// out rows [0, n_stay) <- stayers, sendbuf rows <- leavers (perm is grouped by destination:
// [dest 0 | dest 1 | ... ]; the group of this rank starts at stay_off)
template <typename W>
__global__ void __launch_bounds__(256) gather_rows_kernel(const W* __restrict__ in, W* __restrict__ out,
                                                          W* __restrict__ send,
                                                          const int32_t* __restrict__ perm, long long n,
                                                          long long stay_off, long long n_stay, int words) {
    const long long total = n * words, stride = (long long)gridDim.x * blockDim.x;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += stride) {
        const long long j = e / words;
        const int w = (int)(e % words);
        const W v = in[(long long)perm[j] * words + w];
        if (j < stay_off) send[j * words + w] = v;
        else if (j < stay_off + n_stay) out[(j - stay_off) * words + w] = v;
        else send[(j - n_stay) * words + w] = v;
    }
}

static int mig_alloc(void** p, size_t bytes) {
    if (cudaMalloc(p, bytes ? bytes : 16) != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) failed in migrate", bytes);
        return HYMD_ERR_NOMEM;
    }
    return HYMD_OK;
}

void migrate_destroy(hymd_ctx* c) {
    MigrateState* m = c->mig;
    if (!m) return;
    void* bufs[] = {m->dest, m->dest_sorted, m->idx, m->perm, m->sort_tmp, m->d_counts, m->d_all, m->sendbuf};
    for (void* b : bufs)
        if (b) cudaFree(b);
    if (m->h_all) cudaFreeHost(m->h_all);
    delete m;
    c->mig = nullptr;
}

int migrate_plan(hymd_ctx* c, const void* d_pos, int64_t n, int64_t* n_new, cudaStream_t s) {
    const Geometry& g = c->g;
    const int P = g.P;
    if (!c->mig) c->mig = new MigrateState();
    MigrateState* m = c->mig;
    m->planned = false;
    if (n > m->cap) {
        void* bufs[] = {m->dest, m->dest_sorted, m->idx, m->perm, m->sort_tmp};
        for (void* b : bufs)
            if (b) cudaFree(b);
        m->cap = n + n / 8 + 1024;
        HYMD_CHECK(mig_alloc((void**)&m->dest, m->cap * 4));
        HYMD_CHECK(mig_alloc((void**)&m->dest_sorted, m->cap * 4));
        HYMD_CHECK(mig_alloc((void**)&m->idx, m->cap * 4));
        HYMD_CHECK(mig_alloc((void**)&m->perm, m->cap * 4));
        m->sort_tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, m->sort_tmp_bytes, m->dest, m->dest_sorted, m->idx,
                                        m->perm, (int)m->cap, 0, 4);
        HYMD_CHECK(mig_alloc(&m->sort_tmp, m->sort_tmp_bytes));
    }
    if (!m->d_counts) {
        HYMD_CHECK(mig_alloc((void**)&m->d_counts, 8 * sizeof(unsigned int)));
        HYMD_CHECK(mig_alloc((void**)&m->d_all, 64 * sizeof(unsigned int)));
        HYMD_CUDA(cudaMallocHost((void**)&m->h_all, 64 * sizeof(unsigned int)));
    }
    if (P > 8) { set_error("migrate supports up to 8 slabs"); return HYMD_ERR_INVALID; }
    HYMD_CUDA(cudaMemsetAsync(m->d_counts, 0, 8 * sizeof(unsigned int), s));
    if (n > 0) {
        const unsigned int blocks = (unsigned int)((n + 255) / 256);
        const double sx = g.Nx / g.box[0];
        if (c->f64) dest_kernel<double><<<blocks, 256, 0, s>>>((const double*)d_pos, n, sx, g.Nx, g.nxl,
                                                               m->dest, m->idx, m->d_counts);
        else dest_kernel<float><<<blocks, 256, 0, s>>>((const float*)d_pos, n, sx, g.Nx, g.nxl,
                                                       m->dest, m->idx, m->d_counts);
        HYMD_LAUNCH_CHECK(c);
        size_t tmp = m->sort_tmp_bytes;
        int bits = 1;
        while ((1 << bits) < P) ++bits;
        HYMD_CUDA(cub::DeviceRadixSort::SortPairs(m->sort_tmp, tmp, m->dest, m->dest_sorted, m->idx,
                                                  m->perm, (int)n, 0, bits, s));
        c->launches += 3;
    }
    HYMD_CHECK(comm_allgather_host(c, m->d_counts, m->d_all, 8 * sizeof(unsigned int), s));
    // d_all row r (8 entries) = sends of rank r
    HYMD_CUDA(cudaMemcpyAsync(m->h_all, m->d_all, 64 * sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    HYMD_CUDA(cudaStreamSynchronize(s));
    const int me = g.rank;
    m->n = n;
    m->n_stay = m->h_all[me * 8 + me];
    size_t so = 0, ro = 0;
    for (int r = 0; r < P; ++r) {
        m->send_rows[r] = r == me ? 0 : m->h_all[me * 8 + r];
        m->recv_rows[r] = r == me ? 0 : m->h_all[r * 8 + me];
        m->send_rows_off[r] = so; so += m->send_rows[r];
        m->recv_rows_off[r] = ro; ro += m->recv_rows[r];
    }
    m->n_new = m->n_stay + (int64_t)ro;
    m->planned = true;
    *n_new = m->n_new;
    return HYMD_OK;
}

int migrate_apply(hymd_ctx* c, const void* d_in, void* d_out, int row_bytes, cudaStream_t s) {
    MigrateState* m = c->mig;
    if (!m || !m->planned) { set_error("hymd_migrate_apply before hymd_migrate_plan"); return HYMD_ERR_STATE; }
    if (row_bytes <= 0) { set_error("row_bytes must be positive"); return HYMD_ERR_INVALID; }
    const int P = c->g.P, me = c->g.rank;
    const size_t n_send = (size_t)(m->n - m->n_stay);
    if (n_send * row_bytes > m->sendbuf_bytes) {
        if (m->sendbuf) { cudaStreamSynchronize(s); cudaFree(m->sendbuf); }
        m->sendbuf_bytes = n_send * row_bytes + (1 << 20);
        HYMD_CHECK(mig_alloc(&m->sendbuf, m->sendbuf_bytes));
    }
    // start of this rank's group inside perm
    long long stay_off = 0;
    for (int r = 0; r < me; ++r) stay_off += m->h_all[me * 8 + r];
    if (m->n > 0) {
        long long blocks = (m->n * (row_bytes / 4 + 1) + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        if (row_bytes % 8 == 0 && ((uintptr_t)d_in % 8 == 0) && ((uintptr_t)d_out % 8 == 0))
            gather_rows_kernel<unsigned long long><<<(unsigned)blocks, 256, 0, s>>>(
                (const unsigned long long*)d_in, (unsigned long long*)d_out,
                (unsigned long long*)m->sendbuf, m->perm, m->n, stay_off, m->n_stay, row_bytes / 8);
        else if (row_bytes % 4 == 0)
            gather_rows_kernel<uint32_t><<<(unsigned)blocks, 256, 0, s>>>(
                (const uint32_t*)d_in, (uint32_t*)d_out, (uint32_t*)m->sendbuf, m->perm, m->n,
                stay_off, m->n_stay, row_bytes / 4);
        else
            gather_rows_kernel<unsigned char><<<(unsigned)blocks, 256, 0, s>>>(
                (const unsigned char*)d_in, (unsigned char*)d_out, (unsigned char*)m->sendbuf,
                m->perm, m->n, stay_off, m->n_stay, row_bytes);
        HYMD_LAUNCH_CHECK(c);
    }
    size_t so[8], sb[8], ro[8], rb[8];
    for (int r = 0; r < P; ++r) {
        so[r] = m->send_rows_off[r] * row_bytes; sb[r] = m->send_rows[r] * row_bytes;
        ro[r] = ((size_t)m->n_stay + m->recv_rows_off[r]) * row_bytes; rb[r] = m->recv_rows[r] * row_bytes;
    }
    PhaseScope ps(c, HYMD_PHASE_MIGRATE, s);
    return comm_alltoallv(c, m->sendbuf, so, sb, d_out, ro, rb, s);
}

}  // namespace hymd
