// Fused x-line kernel: the third axis of the forward FFT, ALL the k-space arithmetic and the
// first axis of the inverse FFT in one pass over the spectra.
//
// After the batched 2-D (y,z) transforms every rank holds complete x-lines of all T density
// spectra for its (ky,kz) columns.  One CTA takes CH neighbouring columns, loads the T x Nx x CH
// tile into shared memory, runs the Nx-point FFTs there, forms for every distinct potential row
//     V^_u = H^2 sum_t A[u][t] phi^_t            (field.py:577-585, hamiltonian.py:402-412)
// applies -i k_d (field.py:607-612; Nyquist rule of SURVEY.md section 7), transforms back along x
// and writes the 3U force spectra, ready for the batched 2-D c2r.  This replaces three full
// passes over HBM (x-FFT, k-space kernel, x-IFFT: T*3 + 3U*3 spectrum transfers) by one
// (T reads + 3U writes), and because k_y, k_z are constant along an x-line only two inverse
// transforms per row are needed (V and k_x V) instead of three.
//
// FFT: Nx = R1*R2, four-step inside the CTA, radix-R butterflies in registers.  The forward
// transform leaves frequency k1 + R1*k2 at position k1*R2 + k2 (no reordering pass); the
// inverse consumes exactly that order and ends in natural x order.
//
// PME variant (field.py:369-396): T = U = 1, G = 4 pi c_e H / k^2 (k = 0 divisor -> 1).
#include "ctx.cuh"
#include "fft.cuh"

namespace hymd {

__device__ __forceinline__ void store_cx(Cx<float>* p, Cx<float> v) {
    __stcs(reinterpret_cast<float2*>(p), make_float2(v.x, v.y));
}
__device__ __forceinline__ void store_cx(Cx<double>* p, Cx<double> v) {
    __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y));
}

struct XParams {
    int Nx, Ny, Nz, nyl, y0, Nzc, Nzcp;
    int T, U;
    long long ncols;                 // nyl * Nzcp
    long long xs_in, fs_in;          // complex strides of the input spectra
    long long xs_f, fs_f;            // ... of the 3U (or 3) force outputs
    long long xs_v, fs_v, xs_pf, fs_pf;   // optional k-space outputs (not x-inverted)
    int pme;
    int two;                         // 2 outputs per row (-i k_x V, -i V): the plane c2r applies k_y, k_z
    int no_inplace;                  // HYMD_B200_XLINE_INPLACE=0: one combine pass per potential row (first version; A/B)
    // several slabs: the force spectra go straight into the work buffer of the rank owning plane x (the
    // inverse transpose of the slab FFT, fused): element (f, x, col) -> peer[x / nxl] + (f (nxl+1) + x % nxl)
    // plane + ky0 Nzcp + col, the layout the plane c2r reads
    int push, nxl_shift;
    long long push_plane, push_fs, push_y0;
    void* peer[HYMD_MAX_PEERS];
};

// shared-memory position of element (pos, c): rows of CH columns, one extra row of padding per
// R2 positions so that the contiguous-radix step spreads over all banks
template <int NX, int CH>
__device__ __forceinline__ int spos(int pos, int c) {
    return (pos + pos / Radix<NX>::R2) * CH + c;
}
// pos = hi * R2 + lo with lo < R2: no division
template <int NX, int CH>
__device__ __forceinline__ int spos2(int hi, int lo, int c) {
    return (hi * (Radix<NX>::R2 + 1) + lo) * CH + c;
}
template <int NX, int CH>
constexpr int field_elems() { return (NX + NX / Radix<NX>::R2) * CH; }

// forward: strided radix-R1 over n1 (+ twiddle), then contiguous radix-R2 over n2.
// REG: the twiddles w^(n2 k1) of this thread's (constant) n2 come from registers (TwiddleRegs)
template <typename real, int NX, int CH, bool REG>
__device__ __forceinline__ void fft_fwd_step1(Cx<real>* buf, const Cx<real>* __restrict__ tw,
                                              const TwiddleRegs<real, Radix<NX>::R1>& twr, int task) {
    constexpr int R1 = Radix<NX>::R1, R2 = Radix<NX>::R2;
    const int c = task % CH, n2 = task / CH;
    Cx<real> v[R1];
#pragma unroll
    for (int n1 = 0; n1 < R1; ++n1) v[n1] = buf[spos2<NX, CH>(n1, n2, c)];
    dft_reg<real, R1, -1>(v);
    if constexpr (REG) {
        twr.template apply<false>(v);
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) buf[spos2<NX, CH>(k1, n2, c)] = v[k1];
    } else {
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            const Cx<real> w = tw[(n2 * k1) & (NX - 1)];   // exp(-2 pi i n2 k1 / NX)
            buf[spos2<NX, CH>(k1, n2, c)] = (k1 == 0) ? v[k1] : cmul(w, v[k1]);
        }
    }
}
template <typename real, int NX, int CH>
__device__ __forceinline__ void fft_fwd_step2(Cx<real>* buf, int task) {
    constexpr int R2 = Radix<NX>::R2;
    const int c = task % CH, k1 = task / CH;
    Cx<real> v[R2];
#pragma unroll
    for (int n2 = 0; n2 < R2; ++n2) v[n2] = buf[spos2<NX, CH>(k1, n2, c)];
    dft_reg<real, R2, -1>(v);
#pragma unroll
    for (int k2 = 0; k2 < R2; ++k2) buf[spos2<NX, CH>(k1, k2, c)] = v[k2];
}
// inverse: contiguous radix-R2 over k2 (+ conjugate twiddle), then strided radix-R1 over k1.
template <typename real, int NX, int CH, bool REG>
__device__ __forceinline__ void fft_inv_stepA(Cx<real>* buf, const Cx<real>* __restrict__ tw,
                                              const TwiddleRegs<real, Radix<NX>::R2>& twr, int task) {
    constexpr int R2 = Radix<NX>::R2;
    const int c = task % CH, k1 = task / CH;
    Cx<real> v[R2];
#pragma unroll
    for (int k2 = 0; k2 < R2; ++k2) v[k2] = buf[spos2<NX, CH>(k1, k2, c)];
    dft_reg<real, R2, +1>(v);
    if constexpr (REG) {
        twr.template apply<true>(v);
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) buf[spos2<NX, CH>(k1, n2, c)] = v[n2];
    } else {
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) {
            Cx<real> w = tw[(n2 * k1) & (NX - 1)];
            w.y = -w.y;
            buf[spos2<NX, CH>(k1, n2, c)] = (k1 == 0) ? v[n2] : cmul(w, v[n2]);
        }
    }
}
template <typename real, int NX, int CH>
__device__ __forceinline__ void fft_inv_stepB(Cx<real>* buf, int task) {
    constexpr int R1 = Radix<NX>::R1, R2 = Radix<NX>::R2;
    const int c = task % CH, n2 = task / CH;
    Cx<real> v[R1];
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) v[k1] = buf[spos2<NX, CH>(k1, n2, c)];
    dft_reg<real, R1, +1>(v);
#pragma unroll
    for (int n1 = 0; n1 < R1; ++n1) buf[spos2<NX, CH>(n1, n2, c)] = v[n1];
}

template <typename real, int NX, int CH, int NTH>
__global__ void __launch_bounds__(NTH) xline_kernel(
    const Cx<real>* __restrict__ in, Cx<real>* __restrict__ fout, Cx<real>* __restrict__ vout,
    Cx<real>* __restrict__ pfout, const real* __restrict__ tab, const Cx<real>* __restrict__ twg,
    const real* __restrict__ Au, const real* __restrict__ cu, real coef, real inv_m, XParams p) {
    constexpr int R1 = Radix<NX>::R1, R2 = Radix<NX>::R2;
    constexpr int FE = field_elems<NX, CH>();
    constexpr int NT = NTH;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cx<real>* tw = reinterpret_cast<Cx<real>*>(smem_raw);      // NX twiddles
    Cx<real>* wV = tw + NX;                                     // FE: V^ (then its x-inverse)
    Cx<real>* wK = wV + FE;                                     // FE: k_x V^
    Cx<real>* data = wK + FE;                                   // T x FE input spectra
    __shared__ real s_hk[2][CH][4];   // per column: {h_y h_z, k_y eff, k_z eff, self-conjugate}, {k_y, k_z, origin}
    __shared__ real s_hx[NX], s_kx[NX];
    __shared__ real s_A[HYMD_MAX_TYPES * HYMD_MAX_TYPES];

    const real* hx = tab; const real* hy = hx + p.Nx; const real* hz = hy + p.Ny;
    const real* kxt = hz + p.Nzc; const real* kyt = kxt + p.Nx; const real* kzt = kyt + p.Ny;

    const int tid = threadIdx.x;
    const long long col0 = (long long)blockIdx.x * CH;

    for (int i = tid; i < NX; i += NT) {      // s_hx / s_kx are indexed by POSITION (frequency order of the in-place FFT)
        const int kx = i / R2 + R1 * (i % R2);
        tw[i] = twg[i]; s_hx[i] = hx[kx]; s_kx[i] = kxt[kx];
    }
    for (int i = tid; i < p.U * p.T; i += NT) s_A[i] = Au[i];
    // per-column constants
    if (tid < CH) {
        long long col = col0 + tid;
        if (col >= p.ncols) col = p.ncols - 1;
        const int iz = (int)(col % p.Nzcp), iyl = (int)(col / p.Nzcp);
        const int iy = iyl + p.y0;
        const bool zvalid = iz < p.Nzc;
        const int zc = zvalid ? iz : 0;
        const bool z_nyq = (p.Nz % 2 == 0) && zc == p.Nz / 2;
        const bool self_conj = zc == 0 || z_nyq;
        const bool y_nyq = (p.Ny % 2 == 0) && iy == p.Ny / 2;
        s_hk[0][tid][0] = zvalid ? hy[iy] * hz[zc] : (real)0;           // h_y h_z (0 on the pad column)
        s_hk[0][tid][1] = (y_nyq && self_conj) ? (real)0 : kyt[iy];     // effective k_y
        s_hk[0][tid][2] = z_nyq ? (real)0 : kzt[zc];                    // effective k_z
        s_hk[0][tid][3] = self_conj ? (real)1 : (real)0;
        s_hk[1][tid][0] = kyt[iy];                                      // raw k_y, k_z (PME k^2)
        s_hk[1][tid][1] = kzt[zc];
        s_hk[1][tid][2] = (zvalid && iy == 0 && zc == 0) ? (real)1 : (real)0;   // column with k = 0
        s_hk[1][tid][3] = 0;
    }

    // ---- load T x NX x CH (16-byte cp.async chunks) ----
    constexpr int VEC = 16 / (int)sizeof(Cx<real>);     // complex elements per 16 bytes
    const bool full = col0 + CH <= p.ncols;
    for (int e = tid; e < p.T * NX * (CH / VEC); e += NT) {
        const int c = (e % (CH / VEC)) * VEC, x = (e / (CH / VEC)) % NX, t = e / ((CH / VEC) * NX);
        Cx<real>* dst = data + t * FE + spos<NX, CH>(x, c);
        if (full || col0 + c < p.ncols) {
            const Cx<real>* src = in + t * p.fs_in + x * p.xs_in + col0 + c;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
        } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) dst[i] = {0, 0};
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // twiddles of this thread's constant sub-index (R*CH divides the CTA size, so every task a
    // thread gets in the strided task loops below has the same n2 / k1), loaded while the tile flies
    constexpr bool REG = sizeof(real) == 4 && R1 <= 16 && R2 <= 16 && R1 >= 4 && NT % (R1 * CH) == 0 &&
                         NT % (R2 * CH) == 0;
    TwiddleRegs<real, R1> twf;      // forward: n2 constant, k1 = 0 .. R1-1
    TwiddleRegs<real, R2> twi;      // inverse: k1 constant, n2 = 0 .. R2-1
    if constexpr (REG) {
        twf.init(twg, (tid % (R2 * CH)) / CH, NX);
        twi.init(twg, (tid % (R1 * CH)) / CH, NX);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // ---- forward FFT along x, all fields ----
    for (int task = tid; task < p.T * R2 * CH; task += NT)
        fft_fwd_step1<real, NX, CH, REG>(data + (task / (R2 * CH)) * FE, tw, twf, task % (R2 * CH));
    __syncthreads();
    for (int task = tid; task < p.T * R1 * CH; task += NT)
        fft_fwd_step2<real, NX, CH>(data + (task / (R1 * CH)) * FE, task % (R1 * CH));
    __syncthreads();

    // optional: filtered density spectra H phi^ / M in natural k order (phi_fourier, field.py:577)
    if (pfout != nullptr) {
        for (int e = tid; e < p.T * NX * CH; e += NT) {
            const int c = e % CH, pos = (e / CH) % NX, t = e / (CH * NX);
            const int kx = pos / R2 + R1 * (pos % R2);
            if (full || col0 + c < p.ncols) {
                const real s = s_hx[pos] * s_hk[0][c][0] * inv_m;
                const Cx<real> v = data[t * FE + spos<NX, CH>(pos, c)];
                pfout[t * p.fs_pf + kx * p.xs_pf + col0 + c] = {v.x * s, v.y * s};
            }
        }
    }

    // ---- U <= 4 potential rows (every shipped Hamiltonian): all rows in ONE pass over the T spectra, V^_u written in
    // place over spectrum u (U <= T), k_x V^_u formed by the inverse butterfly's K tasks while they load.  The
    // per-row pass below read the T spectra once per row and stored k_x V^ as a second spectrum: a third of the
    // kernel's shared-memory traffic (ncu: l1tex 70 % busy, the kernel's second limiter after instruction issue).
    constexpr int UF = 4;
    const bool inplace = p.U <= UF && p.U <= p.T && !p.no_inplace;
    if (inplace) {
        for (int e = tid; e < NX * (CH / VEC); e += NT) {
            const int c = (e % (CH / VEC)) * VEC, pos = e / (CH / VEC);
            const int kx = pos / R2 + R1 * (pos % R2);
            const int sp = spos<NX, CH>(pos, c);
            const real hxv = s_hx[pos], kxr = s_kx[pos];
            Cx<real> acc[UF][VEC];
#pragma unroll
            for (int u = 0; u < UF; ++u)
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[u][i] = {0, 0};
            for (int t = 0; t < p.T; ++t) {
                Cx<real> v[VEC];
#pragma unroll
                for (int i = 0; i < VEC; ++i) v[i] = data[t * FE + sp + i];
#pragma unroll
                for (int u = 0; u < UF; ++u) {
                    if (u < p.U) {
                        const real a = s_A[u * p.T + t];
#pragma unroll
                        for (int i = 0; i < VEC; ++i) { acc[u][i].x += a * v[i].x; acc[u][i].y += a * v[i].y; }
                    }
                }
            }
            real g[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const int ci = c + i;
                const real h = hxv * s_hk[0][ci][0];
                if (p.pme) {
                    const real kyr = s_hk[1][ci][0], kzr = s_hk[1][ci][1];
                    real k2 = kxr * kxr + kyr * kyr + kzr * kzr;
                    if (kx == 0 && s_hk[1][ci][2] != (real)0) k2 = (real)1;   // normp(p=2, zeromode=1)
                    g[i] = coef * h / k2;
                } else {
                    g[i] = h * h;
                }
            }
#pragma unroll
            for (int u = 0; u < UF; ++u) {
                if (u < p.U) {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const int ci = c + i;
                        const real ar = acc[u][i].x * g[i], ai = acc[u][i].y * g[i];
                        data[u * FE + sp + i] = {ar, ai};
                        if (vout != nullptr && (full || col0 + ci < p.ncols)) {
                            real vr = ar;
                            if (!p.pme && kx == 0 && s_hk[1][ci][2] != (real)0) vr += cu[u];
                            vout[u * p.fs_v + kx * p.xs_v + col0 + ci] = {vr, ai};
                        }
                    }
                }
            }
        }
        __syncthreads();
    }

    for (int u = 0; u < p.U; ++u) {
        if (inplace) {
            // ---- first inverse butterfly of V^_u (tasks < R1 CH) and of k_x V^_u (the others), both from spectrum u ----
            for (int task = tid; task < 2 * R1 * CH; task += NT) {
                const bool isK = task >= R1 * CH;
                const int rem = task % (R1 * CH), c = rem % CH, k1 = rem / CH;
                const Cx<real>* src = data + u * FE;
                Cx<real>* dst = isK ? wK : wV;
                Cx<real> v[R2];
#pragma unroll
                for (int k2 = 0; k2 < R2; ++k2) v[k2] = src[spos2<NX, CH>(k1, k2, c)];
                if (isK) {
                    const bool selfc = s_hk[0][c][3] != (real)0;
#pragma unroll
                    for (int k2 = 0; k2 < R2; ++k2) {
                        // position k1 R2 + k2 holds frequency k1 + R1 k2; x-Nyquist rule on self-conjugate columns
                        real kxe = s_kx[k1 * R2 + k2];
                        if ((NX % 2 == 0) && k1 + R1 * k2 == NX / 2 && selfc) kxe = 0;
                        v[k2].x *= kxe; v[k2].y *= kxe;
                    }
                }
                dft_reg<real, R2, +1>(v);
                if constexpr (REG) {
                    twi.template apply<true>(v);
#pragma unroll
                    for (int n2 = 0; n2 < R2; ++n2) dst[spos2<NX, CH>(k1, n2, c)] = v[n2];
                } else {
#pragma unroll
                    for (int n2 = 0; n2 < R2; ++n2) {
                        Cx<real> w = tw[(n2 * k1) & (NX - 1)];
                        w.y = -w.y;
                        dst[spos2<NX, CH>(k1, n2, c)] = (k1 == 0) ? v[n2] : cmul(w, v[n2]);
                    }
                }
            }
            __syncthreads();
        } else {
        // ---- potential and k_x * potential in frequency space ----
        for (int e = tid; e < NX * (CH / VEC); e += NT) {
            const int c = (e % (CH / VEC)) * VEC, pos = e / (CH / VEC);
            const int kx = pos / R2 + R1 * (pos % R2);     // frequency index held at this position
            const int sp = spos<NX, CH>(pos, c);
            const real hxv = s_hx[pos], kxr = s_kx[pos];
            const bool x_nyq = (NX % 2 == 0) && kx == NX / 2;
            Cx<real> acc[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] = {0, 0};
            for (int t = 0; t < p.T; ++t) {
                const real a = s_A[u * p.T + t];
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const Cx<real> v = data[t * FE + sp + i];
                    acc[i].x += a * v.x; acc[i].y += a * v.y;
                }
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const int ci = c + i;
                const real h = hxv * s_hk[0][ci][0];
                real g;
                if (p.pme) {
                    const real kyr = s_hk[1][ci][0], kzr = s_hk[1][ci][1];
                    real k2 = kxr * kxr + kyr * kyr + kzr * kzr;
                    if (kx == 0 && s_hk[1][ci][2] != (real)0) k2 = (real)1;   // normp(p=2, zeromode=1)
                    g = coef * h / k2;
                } else {
                    g = h * h;
                }
                const real ar = acc[i].x * g, ai = acc[i].y * g;
                const real kxe = (x_nyq && s_hk[0][ci][3] != (real)0) ? (real)0 : kxr;
                wV[sp + i] = {ar, ai};
                wK[sp + i] = {kxe * ar, kxe * ai};
                if (vout != nullptr && (full || col0 + ci < p.ncols)) {
                    real vr = ar;
                    if (!p.pme && kx == 0 && s_hk[1][ci][2] != (real)0) vr += cu[u];
                    vout[u * p.fs_v + kx * p.xs_v + col0 + ci] = {vr, ai};
                }
            }
        }
        __syncthreads();
        // ---- inverse FFT along x of V^ and k_x V^ (2 x R1 x CH tasks) ----
        for (int task = tid; task < 2 * R1 * CH; task += NT)
            fft_inv_stepA<real, NX, CH, REG>(task < R1 * CH ? wV : wK, tw, twi, task % (R1 * CH));
        __syncthreads();
        }
        // ---- last butterfly stage + F_d = -i k_d V:  -i (a + i b) = b - i a, stored straight from
        // the registers (V -> F_y, F_z; k_x V -> F_x); lanes = neighbouring columns: 64-byte runs ----
        Cx<real>* f0 = fout + (long long)((p.two ? 2 : 3) * u) * p.fs_f;
        for (int task = tid; task < 2 * R2 * CH; task += NT) {
            const bool isK = task >= R2 * CH;
            const int rem = task % (R2 * CH), c = rem % CH, n2 = rem / CH;
            const Cx<real>* buf = isK ? wK : wV;
            Cx<real> v[R1];
#pragma unroll
            for (int k1 = 0; k1 < R1; ++k1) v[k1] = buf[spos2<NX, CH>(k1, n2, c)];
            dft_reg<real, R1, +1>(v);
            if (!(full || col0 + c < p.ncols)) continue;
            const real ky = s_hk[0][c][1], kz = s_hk[0][c][2];
            Cx<real>* o = f0 + (long long)n2 * p.xs_f + col0 + c;
            if (p.push) {
                const int fo = (p.two ? 2 : 3) * u + (isK ? 0 : 1);
                const int nxl_mask = (1 << p.nxl_shift) - 1;
                const long long off = (long long)fo * p.push_fs + p.push_y0 + col0 + c;
#pragma unroll
                for (int n1 = 0; n1 < R1; ++n1) {
                    const int x = n1 * R2 + n2;
                    Cx<real>* dq = reinterpret_cast<Cx<real>*>(p.peer[x >> p.nxl_shift]) + off +
                                   (long long)(x & nxl_mask) * p.push_plane;
                    const Cx<real> w = {v[n1].y, -v[n1].x};
                    if (isK || p.two) {
                        store_cx(dq, w);
                    } else {
                        store_cx(dq, Cx<real>{ky * w.x, ky * w.y});
                        store_cx(dq + p.push_fs, Cx<real>{kz * w.x, kz * w.y});
                    }
                }
                continue;
            }
#pragma unroll
            for (int n1 = 0; n1 < R1; ++n1) {
                Cx<real>* ox = o + (long long)(n1 * R2) * p.xs_f;
                if (isK) {
                    store_cx(ox, Cx<real>{v[n1].y, -v[n1].x});
                } else if (p.two) {
                    store_cx(ox + p.fs_f, Cx<real>{v[n1].y, -v[n1].x});
                } else {
                    store_cx(ox + p.fs_f, Cx<real>{ky * v[n1].y, -ky * v[n1].x});
                    store_cx(ox + 2 * p.fs_f, Cx<real>{kz * v[n1].y, -kz * v[n1].x});
                }
            }
        }
        __syncthreads();
    }
}

// ---- host side --------------------------------------------------------------------------------
static int xline_ch(int n, bool f64) { return n <= 256 ? 8 : (n == 512 ? 4 : 2); }

static size_t xline_smem(int n, int T, bool f64) {
    int r2 = 4;
    switch (n) { case 32: case 64: r2 = 8; break; case 128: case 256: r2 = 16; break;
                 case 512: case 1024: r2 = 32; break; default: break; }
    const size_t fe = (size_t)(n + n / r2) * xline_ch(n, f64);
    return (f64 ? 16 : 8) * ((size_t)n + (size_t)(2 + T) * fe);
}

bool xline_supported(const hymd_ctx* c) {
    const int n = c->g.Nx;
    if (n < 16 || n > 1024 || (n & (n - 1))) return false;
    if (c->f64 && n > 256) return false;     // radix-32 butterflies in fp64 would spill
    return xline_smem(n, c->T, c->f64) <= 220 * 1024;
}

template <typename real, int NX, int CH, int NTH = 256>
static int launch_x(hymd_ctx* c, bool pme, const void* in, void* fout, void* vout, void* pfout,
                    int T, int U, cudaStream_t s, void* const* push_peers) {
    const Geometry& g = c->g;
    XParams p;
    memset(&p, 0, sizeof(p));
    p.Nx = g.Nx; p.Ny = g.Ny; p.Nz = g.Nz; p.nyl = g.nyl; p.y0 = g.y0; p.Nzc = g.Nzc; p.Nzcp = g.Nzcp;
    p.T = T; p.U = U; p.pme = pme ? 1 : 0; p.two = c->grad2 ? 1 : 0;
    if (const char* e = getenv("HYMD_B200_XLINE_INPLACE")) p.no_inplace = e[0] == '0';
    p.ncols = (long long)g.nyl * g.Nzcp;
    const KLayout lin = klayout(c, T), lv = klayout(c, U);
    p.xs_in = lin.xs; p.fs_in = lin.fs;
    p.xs_pf = lin.xs; p.fs_pf = lin.fs;
    p.xs_v = lv.xs; p.fs_v = lv.fs;
    if (g.P == 1) {   // straight into the work layout of the batched 2-D c2r: [f][Nx+1][Ny][Nzcp]
        p.xs_f = (long long)g.Ny * g.Nzcp;
        p.fs_f = (long long)(g.Nx + 1) * p.xs_f;
    } else {
        const KLayout lf = klayout(c, (c->grad2 ? 2 : 3) * U);
        p.xs_f = lf.xs; p.fs_f = lf.fs;
        if (push_peers) {
            int sh = 0;
            while ((1 << sh) < g.nxl) ++sh;
            if ((1 << sh) != g.nxl) { set_error("fused inverse transpose needs a power-of-two Nx / P"); return HYMD_ERR_INVALID; }
            p.push = 1; p.nxl_shift = sh;
            p.push_plane = (long long)g.Ny * g.Nzcp;
            p.push_fs = (long long)(g.nxl + 1) * p.push_plane;
            p.push_y0 = (long long)g.y0 * g.Nzcp;
            for (int q = 0; q < g.P; ++q) p.peer[q] = push_peers[q];
        }
    }
    const size_t smem = sizeof(Cx<real>) * ((size_t)NX + (size_t)(2 + T) * field_elems<NX, CH>());
    if (smem > 227 * 1024) {
        set_error("xline: %d types need %zu B of shared memory", T, smem);
        return HYMD_ERR_INVALID;
    }
    auto kern = xline_kernel<real, NX, CH, NTH>;
    HYMD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const double m = (double)g.Nx * g.Ny * g.Nz;
    const real coef = (real)(4.0 * 3.14159265358979323846 * c->cfg.elec_conversion / m);
    const unsigned int blocks = (unsigned int)((p.ncols + CH - 1) / CH);
    // PME: A = [1/M] lives right after the U x T matrix (build_interaction); its row "0" is used
    const real* Au = pme ? (const real*)c->Au + (size_t)c->U * c->T : (const real*)c->Au;
    kern<<<blocks, NTH, smem, s>>>((const Cx<real>*)in, (Cx<real>*)fout, (Cx<real>*)vout,
                                   (Cx<real>*)pfout, (const real*)c->tab,
                                   (const Cx<real>*)c->xtw, Au, (const real*)c->cu,
                                   pme ? (real)(coef * m) : (real)0, (real)(1.0 / m), p);
    HYMD_LAUNCH_CHECK(c);
    return HYMD_OK;
}

template <typename real>
static int dispatch_x(hymd_ctx* c, bool pme, const void* in, void* fout, void* vout, void* pfout,
                      int T, int U, cudaStream_t s, void* const* push_peers) {
    // columns per CTA: 8 complex = 64 B (fp32) / 128 B (fp64) contiguous per x row
    switch (c->g.Nx) {
        case 16: return launch_x<real, 16, 8>(c, pme, in, fout, vout, pfout, T, U, s, push_peers);
        case 32: return launch_x<real, 32, 8>(c, pme, in, fout, vout, pfout, T, U, s, push_peers);
        case 64: return launch_x<real, 64, 8>(c, pme, in, fout, vout, pfout, T, U, s, push_peers);
        case 128: return launch_x<real, 128, 8>(c, pme, in, fout, vout, pfout, T, U, s, push_peers);
        case 256: return launch_x<real, 256, 8>(c, pme, in, fout, vout, pfout, T, U, s, push_peers);
        default: break;
    }
    if (sizeof(real) == 4) {
        if (c->g.Nx == 512) return launch_x<float, 512, 4>(c, pme, in, fout, vout, pfout, T, U, s, push_peers);
        if (c->g.Nx == 1024) return launch_x<float, 1024, 2>(c, pme, in, fout, vout, pfout, T, U, s, push_peers);
    }
    set_error("xline: unsupported Nx = %d", c->g.Nx);
    return HYMD_ERR_INVALID;
}

// in: spectra after the 2-D (y,z) transforms, k layout of T fields.  fout: 3U x-inverted force
// spectra (P == 1: work layout of the ghost c2r; P > 1: k layout of 3U fields).
// push_peers != NULL (several slabs): fout is ignored, the force spectra are stored into the peers' work
// buffers (the fused inverse transpose; the caller brackets the launch with the acquire / barrier).
int xline_forces(hymd_ctx* c, const void* in, void* fout, void* vout, void* pfout, cudaStream_t s,
                 void* const* push_peers) {
    return c->f64 ? dispatch_x<double>(c, false, in, fout, vout, pfout, c->T, c->U, s, push_peers)
                  : dispatch_x<float>(c, false, in, fout, vout, pfout, c->T, c->U, s, push_peers);
}

int xline_pme(hymd_ctx* c, const void* in, void* fout, void* psi_out, void* rhof_out, cudaStream_t s,
              void* const* push_peers) {
    return c->f64 ? dispatch_x<double>(c, true, in, fout, psi_out, rhof_out, 1, 1, s, push_peers)
                  : dispatch_x<float>(c, true, in, fout, psi_out, rhof_out, 1, 1, s, push_peers);
}

}  // namespace hymd
