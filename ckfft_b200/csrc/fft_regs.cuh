// fft_regs.cuh -- register-resident radix-2..32 butterflies for sm_100a.
//
// Replaces the reference's scalar / NEON butterfly code (src/ckfft/fft_default.cpp:12-266,
// src/ckfft/fft_neon.cpp:16-256, src/ckfft/math_util.h:17-80).  Nothing here is a port of
// that recursion: a radix-R transform (R <= 32) lives entirely in one thread's registers as
// a fully unrolled decimation-in-time network whose twiddles are compile-time immediates,
// written so that every general butterfly is 6 FFMA (a + w*b, then 2a - (a + w*b)).
#pragma once
#include <cuda_runtime.h>

namespace ckb {

typedef float2 cf;

template <int V> struct Int { static constexpr int value = V; };

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f)
{
    if constexpr (I < N) {
        f(Int<I>{});
        static_for<I + 1, N>(static_cast<F&&>(f));
    }
}

__host__ __device__ constexpr int ilog2(int x) { int l = 0; while ((1 << l) < x) ++l; return l; }

// reverse the low log2(R) bits of x
template <int R>
__host__ __device__ constexpr int bitrev(int x)
{
    int r = 0;
    for (int b = 1; b < R; b <<= 1) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

// cos(2*pi*k/32), k any integer, from the first octant + symmetry
__host__ __device__ constexpr float cos32(int k)
{
    constexpr float t[9] = {
        1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
        0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f,
        0.19509032201612826785f, 0.0f };
    k &= 31;
    if (k > 16) k = 32 - k;
    return k > 8 ? -t[16 - k] : t[k];
}
__host__ __device__ constexpr float sin32(int k) { return cos32(k - 8); }

// cos / sin(2*pi*k/64), k any integer (the real split / twist factors a thread needs are its own W_2M^j times these)
__host__ __device__ constexpr float cos64(int k)
{
    constexpr float t[17] = { 1.0f, 0.99518472667219692873f, 0.98078528040323043058f, 0.95694033573220882438f, 0.92387953251128673848f, 0.88192126434835504956f, 0.83146961230254523567f, 0.77301045336273699338f, 0.70710678118654757274f, 0.63439328416364548779f, 0.55557023301960228867f, 0.47139673682599780857f, 0.38268343236508983729f, 0.29028467725446233105f, 0.19509032201612833135f, 0.09801714032956077016f, 0.0f };
    k &= 63;
    if (k > 32) k = 64 - k;
    return k > 16 ? -t[32 - k] : t[k];
}
__host__ __device__ constexpr float sin64(int k) { return cos64(k - 16); }

// Packed single precision (sm_100: add / mul / fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2): one instruction works on both
// halves of an aligned register pair, i.e. on a whole complex value.  ptxas folds the half swap, a per-half sign and a
// scalar broadcast into operand modifiers (R.F32x2.LO_HI, .NP, R.F32), so a complex add is 1 instruction instead of 2, a
// complex multiply 2 instead of 4, a twiddled butterfly 4 instead of 6 (the 1024-point loop: 760 instead of 1095
// instructions).  Each half is rounded exactly as the scalar instruction would round it, and the GPU parity tests pass
// with it.  MEASURED ON B200 (-DCKB_PACKED_MATH=1 for the whole library): slower almost everywhere -- the aligned-pair
// operands cost registers, and the kernels at the 128-register cap spill (16384 points: 96 -> 432 bytes, 0.77 -> 0.49 of
// the copy peak; 8192: 0.87 -> 0.78; 1024: 0.91 -> 0.88), only C2R n = 1024 (0.85 -> 0.90) and C2R n = 64 gain.  It
// therefore stays OFF; a per-kernel switch for the few plans with register headroom is the follow-up.
#ifndef CKB_PACKED_MATH
#define CKB_PACKED_MATH 0
#endif

__device__ __forceinline__ unsigned long long pk2(cf a)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ cf upk2(unsigned long long a)
{
    cf r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(a));
    return r;
}
__device__ __forceinline__ cf add2(cf a, cf b) { unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b))); return upk2(d); }
__device__ __forceinline__ cf sub2(cf a, cf b) { unsigned long long d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b))); return upk2(d); }
__device__ __forceinline__ cf mul2(cf a, cf b) { unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b))); return upk2(d); }
__device__ __forceinline__ cf fma2(cf a, cf b, cf c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b)), "l"(pk2(c)));
    return upk2(d);
}
__device__ __forceinline__ cf swp(cf a) { return make_float2(a.y, a.x); }
__device__ __forceinline__ cf bc(float x) { return make_float2(x, x); }

// CKB_PACKED_CMUL: packed arithmetic in the complex multiplications only (twiddles, split / twist factors).  Their operands
// are register pairs that 64-bit loads already aligned, so this costs no extra registers, unlike the packed butterflies.
#ifndef CKB_PACKED_CMUL
#define CKB_PACKED_CMUL CKB_PACKED_MATH
#endif

__device__ __forceinline__ cf cmul(cf a, cf w)
{
#if CKB_PACKED_CMUL
    const cf u = mul2(bc(w.y), swp(a));                              // (w.y a.y, w.y a.x)
    return fma2(bc(w.x), a, make_float2(-u.x, u.y));                 // (w.x a.x - w.y a.y, w.x a.y + w.y a.x)
#else
    return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
#endif
}

// a * W_64^K (forward sign, exp(-2 pi i K/64)), K a compile-time constant: trivial factors cost nothing
template <int K>
__device__ __forceinline__ cf cmul_w64(cf a)
{
    constexpr int k = K & 63;
    if constexpr (k == 0) return a;
    else if constexpr (k == 16) return make_float2(a.y, -a.x);      // * -i
    else if constexpr (k == 32) return make_float2(-a.x, -a.y);
    else if constexpr (k == 48) return make_float2(-a.y, a.x);      // * +i
    else return cmul(a, make_float2(cos64(k), -sin64(k)));
}

// (a, b) <- (a + w b, a - w b) with w = exp(-+ 2 pi i K/32), K in [0, 16), sign by INV
template <int K, bool INV>
__device__ __forceinline__ void bfly(cf& a, cf& b)
{
    static_assert(K >= 0 && K < 16, "DIT twiddles live in the upper half plane");
#if CKB_PACKED_MATH
    if constexpr (K == 0) {
        const cf s = add2(a, b), d = sub2(a, b);
        a = s; b = d;
    } else if constexpr (K == 8) {
        // w b = (b.y, -b.x) forward, (-b.y, b.x) inverse
        const cf bs = swp(b);
        const cf s = fma2(bs, INV ? make_float2(-1.f, 1.f) : make_float2(1.f, -1.f), a);
        const cf d = fma2(bs, INV ? make_float2(1.f, -1.f) : make_float2(-1.f, 1.f), a);
        a = s; b = d;
    } else if constexpr (K == 4 || K == 12) {
        // w b = +-c (p, q) with (p, q) = b + (+-b.y, -+b.x)
        constexpr float c = 0.70710678118654752440f;
        constexpr bool plus_minus = (K == 4) != INV;                 // (p, q) = (b.x + b.y, b.y - b.x), else (b.x - b.y, b.y + b.x)
        const cf pq = fma2(swp(b), plus_minus ? make_float2(1.f, -1.f) : make_float2(-1.f, 1.f), b);
        constexpr float cs = K == 4 ? c : -c;
        const cf s = fma2(bc(cs), pq, a), d = fma2(bc(-cs), pq, a);
        a = s; b = d;
    } else {
        constexpr float wr = cos32(K);
        constexpr float wi = INV ? sin32(K) : -sin32(K);
        const cf bs = swp(b);
        const cf s = fma2(bc(wr), b, fma2(make_float2(-wi, wi), bs, a));      // a + w b
        const cf d = fma2(bc(-wr), b, fma2(make_float2(wi, -wi), bs, a));     // a - w b
        a = s; b = d;
    }
#else
    if constexpr (K == 0) {
        cf s = make_float2(a.x + b.x, a.y + b.y);
        cf d = make_float2(a.x - b.x, a.y - b.y);
        a = s; b = d;
    } else if constexpr (K == 8) {
        // w = -i (forward) / +i (inverse):  w b = (b.y, -b.x) / (-b.y, b.x)
        cf s, d;
        if constexpr (!INV) { s = make_float2(a.x + b.y, a.y - b.x); d = make_float2(a.x - b.y, a.y + b.x); }
        else                { s = make_float2(a.x - b.y, a.y + b.x); d = make_float2(a.x + b.y, a.y - b.x); }
        a = s; b = d;
    } else if constexpr (K == 4 || K == 12) {
        // w = c (+-1 -+ i) : two adds, then four FFMA with the immediate c
        constexpr float c = 0.70710678118654752440f;
        float p, q;   // w b = c * (p, q)
        if constexpr (K == 4) {
            if constexpr (!INV) { p = b.x + b.y; q = b.y - b.x; } else { p = b.x - b.y; q = b.x + b.y; }
        } else {
            if constexpr (!INV) { p = b.y - b.x; q = -(b.x + b.y); } else { p = -(b.x + b.y); q = b.x - b.y; }
        }
        cf s = make_float2(fmaf(c, p, a.x), fmaf(c, q, a.y));
        cf d = make_float2(fmaf(-c, p, a.x), fmaf(-c, q, a.y));
        a = s; b = d;
    } else {
        constexpr float wr = cos32(K);
        constexpr float wi = INV ? sin32(K) : -sin32(K);
        cf s = make_float2(fmaf(wr, b.x, fmaf(-wi, b.y, a.x)), fmaf(wr, b.y, fmaf(wi, b.x, a.y)));
        cf d = make_float2(fmaf(2.0f, a.x, -s.x), fmaf(2.0f, a.y, -s.y));
        a = s; b = d;
    }
#endif
}

// In-register DFT of R points held in v[OFF .. OFF+R).  Input sample t must sit in slot
// OFF + bitrev<R>(t); output bin u is left in slot OFF + u.
template <int R, int OFF, bool INV, int E>
__device__ __forceinline__ void fft_regs(cf (&v)[E])
{
    static_assert(R >= 1 && R <= 32 && (R & (R - 1)) == 0, "radix must be a power of two <= 32");
    static_for<1, ilog2(R) + 1>([&](auto s_) {
        constexpr int len = 1 << decltype(s_)::value;
        constexpr int half = len / 2;
        static_for<0, R / len>([&](auto b_) {
            constexpr int base = OFF + decltype(b_)::value * len;
            static_for<0, half>([&](auto k_) {
                constexpr int k = decltype(k_)::value;
                bfly<k * (32 / len), INV>(v[base + k], v[base + k + half]);
            });
        });
    });
}

}  // namespace ckb
