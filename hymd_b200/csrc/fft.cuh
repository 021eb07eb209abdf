// In-register / shared-memory FFT building blocks shared by the fused x-line kernel (xline.cu)
// and the (y,z) plane transforms (planefft.cu).  Power-of-two lengths N = R1*R2, four-step inside
// a CTA with radix-R butterflies held in registers.
#pragma once
#include <type_traits>

namespace hymd {

template <typename real> struct Cx { real x, y; };

template <typename real>
__device__ __forceinline__ Cx<real> cmul(Cx<real> a, Cx<real> b) {
    return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}

// conj(a) * b
template <typename real>
__device__ __forceinline__ Cx<real> cmulc(Cx<real> a, Cx<real> b) {
    return {a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x};
}

// cos(2 pi j / 32); j is a compile-time constant after unrolling, so the switch folds away
template <typename real>
__device__ __forceinline__ real cos32(int j) {
    j &= 31;
    if (j > 16) j = 32 - j;
    switch (j) {
        case 0: return (real)1.0;
        case 1: return (real)0.98078528040323044913;
        case 2: return (real)0.92387953251128675613;
        case 3: return (real)0.83146961230254523708;
        case 4: return (real)0.70710678118654752440;
        case 5: return (real)0.55557023301960222474;
        case 6: return (real)0.38268343236508977173;
        case 7: return (real)0.19509032201612826785;
        case 8: return (real)0.0;
        case 9: return (real)-0.19509032201612826785;
        case 10: return (real)-0.38268343236508977173;
        case 11: return (real)-0.55557023301960222474;
        case 12: return (real)-0.70710678118654752440;
        case 13: return (real)-0.83146961230254523708;
        case 14: return (real)-0.92387953251128675613;
        case 15: return (real)-0.98078528040323044913;
        default: return (real)-1.0;
    }
}
template <typename real>
__device__ __forceinline__ real sin32(int j) { return cos32<real>(j - 8); }

template <int R> __host__ __device__ constexpr int log2c() { return R <= 1 ? 0 : 1 + log2c<R / 2>(); }
template <int R> __host__ __device__ constexpr int bitrev(int i) {
    int r = 0;
    for (int b = 0; b < log2c<R>(); ++b) r |= ((i >> b) & 1) << (log2c<R>() - 1 - b);
    return r;
}

// compile-time unrolled helpers (every register-array index is a constant expression)
template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

// In-register radix-2 DIT DFT of size R (power of two <= 32), sign = -1 forward, +1 inverse.
template <typename real, int R, int SIGN>
__device__ __forceinline__ void dft_reg(Cx<real> (&v)[R]) {
    static_for<0, R>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr int r = bitrev<R>(i);
        if constexpr (r > i) { const Cx<real> tmp = v[i]; v[i] = v[r]; v[r] = tmp; }
    });
    static_for<1, log2c<R>() + 1>([&](auto sc) {
        constexpr int m = 1 << decltype(sc)::value;
        static_for<0, R / 2>([&](auto bc) {
            constexpr int bfly = decltype(bc)::value;          // butterfly number 0 .. R/2-1
            constexpr int k = (bfly / (m / 2)) * m, j = bfly % (m / 2);
            constexpr int tj = j * (32 / m);
            const Cx<real> a = v[k + j];
            Cx<real> b = v[k + j + m / 2];
            // w = exp(SIGN * 2 pi i tj / 32); quarter and eighth turns need no general multiply
            if constexpr (tj == 8) {
                b = {-(real)SIGN * b.y, (real)SIGN * b.x};
            } else if constexpr (tj == 4) {
                constexpr real h = (real)0.70710678118654752440;
                b = {h * (b.x - (real)SIGN * b.y), h * ((real)SIGN * b.x + b.y)};
            } else if constexpr (tj == 12) {
                constexpr real h = (real)0.70710678118654752440;
                b = {-h * (b.x + (real)SIGN * b.y), h * ((real)SIGN * b.x - b.y)};
            } else if constexpr (j != 0) {
                const Cx<real> w = {cos32<real>(tj), (real)SIGN * sin32<real>(tj)};
                b = cmul(w, b);
            }
            v[k + j] = {a.x + b.x, a.y + b.y};
            v[k + j + m / 2] = {a.x - b.x, a.y - b.y};
        });
    });
}

// Four-step twiddles w^(i j), i = 0 .. LEN-1, for a THREAD-CONSTANT j (w = exp(-2 pi i / N)), held in
// registers as two short tables: S[b-1] = w^(b j), b = 1..3, and T[a-1] = w^(4 a j), a = 1..LEN/4-1,
// so that w^(i j) = T[i >> 2] * S[i & 3].  Replaces LEN dependent shared-memory loads per
// butterfly group (a third of the shared-memory instructions of a four-step pass, and the ones
// whose latency the few resident warps cannot hide) by at most LEN/2 extra complex multiplies.
template <typename real, int LEN>
struct TwiddleRegs {
    static constexpr int NA = LEN / 4;
    Cx<real> S[3], T[NA > 1 ? NA - 1 : 1];
    __device__ __forceinline__ void init(const Cx<real>* __restrict__ tw, int j, int n) {
#pragma unroll
        for (int b = 1; b < 4; ++b) S[b - 1] = tw[(b * j) & (n - 1)];
#pragma unroll
        for (int a = 1; a < NA; ++a) T[a - 1] = tw[(4 * a * j) & (n - 1)];
    }
    // v[i] *= w^(i j)  (CONJ: the conjugate, inverse transforms)
    template <bool CONJ>
    __device__ __forceinline__ void apply(Cx<real> (&v)[LEN]) const {
        static_for<1, LEN>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            constexpr int b = i & 3, a = i >> 2;
            if constexpr (b != 0) v[i] = CONJ ? cmulc(S[b - 1], v[i]) : cmul(S[b - 1], v[i]);
            if constexpr (a != 0) v[i] = CONJ ? cmulc(T[a - 1], v[i]) : cmul(T[a - 1], v[i]);
        });
    }
};

__device__ __forceinline__ void store16(Cx<float>* p, const Cx<float> (&v)[2]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
}
__device__ __forceinline__ void store16(Cx<double>* p, const Cx<double> (&v)[1]) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0].x, v[0].y);
}

template <int NX> struct Radix;
template <> struct Radix<16> { static constexpr int R1 = 4, R2 = 4; };
template <> struct Radix<32> { static constexpr int R1 = 4, R2 = 8; };
template <> struct Radix<64> { static constexpr int R1 = 8, R2 = 8; };
template <> struct Radix<128> { static constexpr int R1 = 8, R2 = 16; };
template <> struct Radix<256> { static constexpr int R1 = 16, R2 = 16; };
template <> struct Radix<512> { static constexpr int R1 = 16, R2 = 32; };
template <> struct Radix<1024> { static constexpr int R1 = 32, R2 = 32; };

}  // namespace hymd
