// fft_kernel.cuh -- single-pass batched FFT kernel family for sm_100a (M = 16 .. 16384 complex points).
//
// One launch does one HBM read and one HBM write per transform.  It replaces the reference's
// complex driver + recursion (src/ckfft/fft.cpp:13-46, src/ckfft/fft_default.cpp:12-266) and, in
// the real modes, the split / twist passes (src/ckfft/fft_real_default.cpp:13-114) which are fused
// into the same launch through shared memory.
//
// Structure (Stockham autosort, 2 or 3 radix stages, no bit-reversal pass):
//   * a "group" of T = M/E threads owns one transform; every thread keeps E complex values in
//     registers; G groups share a CTA;  CTAs are persistent and walk the batch grid-stride;
//   * stage s loads x[j + t*M/R], multiplies by the stage twiddle W_{Ns*R}^{t*(j mod Ns)}, runs an R-point
//     register DFT (fft_regs.cuh) and scatters to (j/Ns)*Ns*R + (j mod Ns) + u*Ns.  Twiddles come from a
//     shared-memory LUT laid out [t][j mod Ns] (immediate-offset, conflict-free LDS.64) or, where the plan
//     says so (TWR, power bases), from a few table values the thread keeps in registers for the whole kernel;
//   * stages exchange through a padded shared-memory buffer (index p -> p + p/R0: conflict-free for the
//     stride-R0 scatter and for the unit-stride gather); groups of <= 32 threads synchronise with __syncwarp
//     only, larger groups with a named barrier;
//   * global input: every row is one contiguous run of M*8 bytes.  The prefetching variants (Cfg::PF) have
//     one elected thread per group fetch the group's NEXT row with a bulk asynchronous copy (cp.async.bulk,
//     TMA's 1-D mode, completion on an mbarrier) while the current row is being transformed; the plain variant
//     gathers straight into registers (256 contiguous bytes per warp request when T >= 32);
//   * global output: the last stage scatters X[j + u*M/R] straight from registers, again in runs of
//     contiguous 8-byte elements per warp request.
//   * real transforms: the split (R2C) runs on mirror-paired butterflies in registers or through the exchange
//     buffer, the twist (C2R) in shared memory before stage 0 -- no extra pass over HBM, no tmpBuf traffic.
//
// Twiddle values come from the context's device table W_Nt^k, which the host fills with the
// reference's exact fp32 formula (src/ckfft/context.cpp:90-105), so every twiddle used here is
// bit-identical to the table entry the reference would have used for the same angle (or a product of at
// most log2(R) such entries).
#pragma once
#include <stdint.h>
#include "fft_regs.cuh"

namespace ckb {

enum Mode { MODE_C2C = 0, MODE_R2C = 1, MODE_C2R = 2 };

struct KernelParams {
    const cf* in;        // MODE_C2C/C2R: complex input; MODE_R2C: real input viewed as M complex
    cf* out;             // MODE_C2C/R2C: complex output; MODE_C2R: real output viewed as M complex
    const cf* table;     // W_Nt^k, k in [0, Nt)  (forward sign)
    int log2_nt;         // log2(Nt)
    long long batch;
    long long in_stride;   // distance between consecutive transforms, in 8-byte units
    long long out_stride;  // (AUDIO kernels: the output is a float array and this stride is in floats)
    const cf* window;      // AUDIO kernels: analysis window, n floats viewed as M complex, or nullptr
    // PLANAR kernels (split-complex layout): `in` / `out` are the arrays of real parts viewed as
    // float*, these the arrays of imaginary parts; both strides are in floats
    const float* in_im;
    float* out_im;
};

template <int LOGPAD> __device__ __forceinline__ int padidx(int p) { return p + (p >> LOGPAD); }
// padidx(a + c) == padidx(a) + padoff(c) whenever c is a multiple of the padding period 2^LOGPAD.  The stage functions
// use this to address a butterfly's R values as ONE run-time base plus compile-time offsets: ptxas does not find the
// identity by itself and computed every padded index separately (4 integer instructions and one live register per
// shared-memory access: 14-22 % of the instructions of the three-stage kernels, and the cause of their spills).
template <int LOGPAD> __host__ __device__ constexpr int padoff(int c) { return c + (c >> LOGPAD); }

__device__ __forceinline__ cf ld_stream(const cf* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(cf* p, cf v) { __stcs(p, v); }

__device__ __forceinline__ cf table_w(const cf* table, int idx, bool inverse)
{
    cf w = __ldg(table + idx);
    if (inverse) w.y = -w.y;
    return w;
}

// Table element base[c * 2^sh] for a compile-time c: the byte step 8 << sh goes through an opaque register so that the
// address is ONE multiply-add (IMAD.WIDE.U32 base + c * step) instead of a run-time shift plus a 64-bit add per look-up.
__device__ __forceinline__ unsigned table_step_bytes(int sh)
{
    unsigned r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(8u), "r"((unsigned) sh));
    return r;
}
__device__ __forceinline__ cf table_at(const cf* __restrict__ base, unsigned c, unsigned step_bytes)
{
    return __ldg(reinterpret_cast<const cf*>(reinterpret_cast<const char*>(base) + (unsigned long long) step_bytes * c));
}

template <int T>
__device__ __forceinline__ void group_sync(int g)
{
    if constexpr (T <= 32) {
        __syncwarp();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(T) : "memory");
    }
}

// ---- bulk asynchronous copies (cp.async.bulk = TMA's 1-D mode, SASS UBLKCP) and their mbarriers ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    while (!mbar_try_wait(bar, parity)) { }
}

__device__ __forceinline__ unsigned long long l2_evict_first_policy()
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is signalled on `bar`
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar, unsigned long long policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

enum Src { SRC_GLOBAL = 0, SRC_XBUF = 1, SRC_INBUF = 2, SRC_GLOBAL_KEEP = 3 };   // _KEEP: default cache policy (L2-resident input)
enum Dst { DST_GLOBAL = 0, DST_XCHG = 1, DST_XNAT = 2 };
enum Tw { TW_NONE = 0, TW_LUT = 1 };   // register-resident twiddles have their own stage functions (TWR, power bases)

// TWR: the stage twiddles W^(t*m), t = a*LO + b, are rebuilt from LO-1 + HI-1 table values that the
// thread keeps in registers for the whole kernel:  W^(t*m) = W^(a*LO*m) * W^(b*m)  (one extra rounding on
// the products, none on the bases).  Removes the per-transform LUT traffic from the shared-memory port.
template <int R> struct TwSplit {
    static constexpr int LO = R >= 32 ? 8 : (R >= 8 ? 4 : 2);
    static constexpr int HI = R / LO;
    static constexpr int NB = (LO - 1) + (HI - 1);
};

template <int R, bool INV>
__device__ __forceinline__ void load_tw_bases(cf (&twb)[TwSplit<R>::NB], const cf* __restrict__ table, int m, int tshift)
{
    constexpr int LO = TwSplit<R>::LO, HI = TwSplit<R>::HI;
    static_for<1, LO>([&](auto b_) { constexpr int b = decltype(b_)::value; twb[b - 1] = table_w(table, (m * b) << tshift, INV); });
    static_for<1, HI>([&](auto a_) { constexpr int a = decltype(a_)::value; twb[LO - 1 + a - 1] = table_w(table, (m * LO * a) << tshift, INV); });
}

template <int R, int TT>
__device__ __forceinline__ cf tw_from_bases(const cf (&twb)[TwSplit<R>::NB])
{
    constexpr int LO = TwSplit<R>::LO;
    constexpr int a = TT / LO, b = TT % LO;
    if constexpr (a == 0) return twb[b - 1];
    else if constexpr (b == 0) return twb[LO - 1 + a - 1];
    else return cmul(twb[LO - 1 + a - 1], twb[b - 1]);
}

// Which butterfly (0 .. M/R-1) register block q of thread j works on.  The usual map is j + q*T.  The
// PAIRED map (last stage of the real-forward kernel) hands a thread butterflies in mirror pairs
// (p, STR - p): butterfly p yields Z[p + u*STR] and its mirror images Z[M - (p + u*STR)] =
// Z[(STR - p) + (R-1-u)*STR] all come from butterfly STR - p, so the real split
// (fft_real_default.cpp:30-58) runs entirely in this thread's registers.  Butterflies 0 and STR/2 are their
// own mirrors and share the pair slot of thread 0.
template <int T, int STR, bool PAIRED>
__device__ __forceinline__ int bfly_index(int j, int q)
{
    if constexpr (!PAIRED) return j + q * T;
    else {
        const int p = j + (q >> 1) * T;
        if ((q & 1) == 0) return p;
        return p == 0 ? STR / 2 : STR - p;
    }
}

// One Stockham stage = gather + math + scatter over the E register values of a thread.  They are
// separate so the kernel can put a group barrier between a shared-memory gather and the scatter
// that reuses the same buffer.
template <int M, int T, int E, int R, int LOGPAD, int SRC, bool PAIRED = false, bool LIN = true>
__device__ __forceinline__ void stage_gather(cf (&v)[E], const cf* __restrict__ gsrc, const cf* xb, int j, bool valid)
{
    constexpr int B = E / R;      // butterflies per thread
    constexpr int STR = M / R;    // distance between the R inputs of one butterfly
    constexpr int PADW = LIN ? 1 << LOGPAD : (1 << 30);     // !LIN (Cfg::LIN, by measurement): every padded index on its own
    static_assert(B * R == E, "radix must divide the per-thread element count");
    static_for<0, B>([&](auto q_) {
        constexpr int q = decltype(q_)::value;
        const int jq = bfly_index<T, STR, PAIRED>(j, q);
        // padded position of the butterfly's first input; the others are compile-time offsets from it (see padoff)
        const cf* xq = xb;
        (void) xq;
        if constexpr (SRC == SRC_XBUF && STR % PADW == 0) {
            if constexpr (T % PADW != 0)  xq = xb + padidx<LOGPAD>(jq);
            else if constexpr (!PAIRED || (q & 1) == 0) xq = xb + padidx<LOGPAD>(j) + padoff<LOGPAD>((PAIRED ? q >> 1 : q) * T);
            else if constexpr (q == 1)    xq = xb + padidx<LOGPAD>(jq);                  // STR - j, or STR/2 for thread 0
            else                          xq = xb + padidx<LOGPAD>(STR - j) - padoff<LOGPAD>((q >> 1) * T);   // STR - (j + s*T), s >= 1
        }
        static_for<0, R>([&](auto t_) {
            constexpr int t = decltype(t_)::value;
            constexpr int slot = q * R + bitrev<R>(t);
            if constexpr (SRC == SRC_GLOBAL) {
                if (valid) v[slot] = ld_stream(gsrc + jq + t * STR);   // an invalid group computes on garbage and stores nothing
            } else if constexpr (SRC == SRC_GLOBAL_KEEP) {
                if (valid) v[slot] = __ldg(gsrc + jq + t * STR);
            } else if constexpr (SRC == SRC_INBUF) {
                v[slot] = xb[jq + t * STR];            // dense staging buffer filled by a bulk copy
            } else if constexpr (STR % PADW == 0) {
                v[slot] = xq[padoff<LOGPAD>(t * STR)];
            } else {
                v[slot] = xb[padidx<LOGPAD>(jq + t * STR)];
            }
        });
    });
}

// Last stage of the three-stage plans: per butterfly q the thread keeps W^(m*2^i), i < log2(R), in registers
// (exact table values; m = j + q*T is loop invariant) and forms W^(m*t) as W^(m*(t - msb)) * W^(m*msb):
// one complex multiply per non-power-of-two t, at most log2(R)-1 roundings deep.
template <int M, int T, int E, int R, bool INV, bool PAIRED>
__device__ __forceinline__ void load_pow_bases(cf (&pw)[(E / R) * ilog2(R)], const cf* __restrict__ table, int j, int tshift)
{
    constexpr int B = E / R, LG = ilog2(R);
    static_for<0, B>([&](auto q_) {
        constexpr int q = decltype(q_)::value;
        const int m = bfly_index<T, M / R, PAIRED>(j, q);
        static_for<0, LG>([&](auto i_) {
            constexpr int i = decltype(i_)::value;
            pw[q * LG + i] = table_w(table, (m << i) << tshift, INV);
        });
    });
}

template <int R, int TT, int LG, int NPW>
__device__ __forceinline__ cf pow_twiddle(const cf (&pw)[NPW], int qoff)
{
    constexpr int msb = 1 << (ilog2(TT + 1) - 1);       // highest set bit of TT
    if constexpr (TT == msb) return pw[qoff + ilog2(msb)];
    else return cmul(pow_twiddle<R, TT - msb, LG, NPW>(pw, qoff), pw[qoff + ilog2(msb)]);
}

template <int E, int R, bool INV>
__device__ __forceinline__ void stage_math_pow(cf (&v)[E], const cf (&pw)[(E / R) * ilog2(R)])
{
    constexpr int B = E / R, LG = ilog2(R);
    static_for<0, B>([&](auto q_) {
        constexpr int q = decltype(q_)::value;
        static_for<1, R>([&](auto t_) {
            constexpr int t = decltype(t_)::value;
            constexpr int slot = q * R + bitrev<R>(t);
            v[slot] = cmul(v[slot], pow_twiddle<R, t, LG, B * LG>(pw, q * LG));
        });
    });
    static_for<0, B>([&](auto q_) { fft_regs<R, decltype(q_)::value * R, INV>(v); });
}

// stage_math with register-resident twiddle bases (one butterfly per thread: m is loop invariant)
template <int E, int R, bool INV>
__device__ __forceinline__ void stage_math_regs(cf (&v)[E], const cf (&twb)[TwSplit<R>::NB])
{
    static_assert(E == R, "TWR needs one butterfly per thread");
    static_for<1, R>([&](auto t_) {
        constexpr int t = decltype(t_)::value;
        constexpr int slot = bitrev<R>(t);
        v[slot] = cmul(v[slot], tw_from_bases<R, t>(twb));
    });
    fft_regs<R, 0, INV>(v);
}

template <int T, int E, int R, int NS, bool INV, int TW, bool PAIRED = false>
__device__ __forceinline__ void stage_math(cf (&v)[E], const cf* lut, const cf* __restrict__ table, int tshift, int j)
{
    constexpr int B = E / R;
    static_assert(!PAIRED || NS * R == T * E, "the paired map is for the last stage (STR == NS)");
    if constexpr (TW != TW_NONE) {
        static_for<0, B>([&](auto q_) {
            constexpr int q = decltype(q_)::value;
            const int m = bfly_index<T, NS, PAIRED>(j, q) & (NS - 1);
            static_for<1, R>([&](auto t_) {
                constexpr int t = decltype(t_)::value;
                constexpr int slot = q * R + bitrev<R>(t);
                v[slot] = cmul(v[slot], lut[(t - 1) * NS + m]);
            });
        });
    }
    static_for<0, B>([&](auto q_) { fft_regs<R, decltype(q_)::value * R, INV>(v); });
}

template <int M, int T, int E, int R, int NS, int LOGPAD, int DST, bool LIN = true>
__device__ __forceinline__ void stage_scatter(const cf (&v)[E], cf* __restrict__ gdst, cf* xb, int j, bool valid)
{
    constexpr int B = E / R;
    constexpr int STR = M / R;
    constexpr int PADW = LIN ? 1 << LOGPAD : (1 << 30);
    static_assert(DST != DST_GLOBAL || NS * R == M, "only the last stage writes global memory");
    static_for<0, B>([&](auto q_) {
        constexpr int q = decltype(q_)::value;
        const int jq = j + q * T;
        (void) jq;
        // one run-time base per butterfly, compile-time offsets for its R outputs (see padoff)
        constexpr bool LIN_XCHG = DST == DST_XCHG && ((NS == 1 && R == PADW) || (NS > 1 && NS % PADW == 0));
        constexpr bool LIN_XNAT = DST == DST_XNAT && STR % PADW == 0;
        cf* xq = xb;
        (void) xq;
        if constexpr (LIN_XCHG && NS == 1) {
            xq = xb + j * (R + 1) + q * T * (R + 1);                   // padidx(jq * R + u) = jq * (R + 1) + u for u < R = 2^LOGPAD
        } else if constexpr (LIN_XCHG) {
            if constexpr (T % NS == 0) xq = xb + padidx<LOGPAD>((j / NS) * (NS * R) + (j & (NS - 1))) + padoff<LOGPAD>(q * T * R);
            else                       xq = xb + padidx<LOGPAD>((jq / NS) * (NS * R) + (jq & (NS - 1)));
        } else if constexpr (LIN_XNAT) {
            if constexpr (T % PADW == 0) xq = xb + padidx<LOGPAD>(j) + padoff<LOGPAD>(q * T);
            else                         xq = xb + padidx<LOGPAD>(jq);
        }
        static_for<0, R>([&](auto u_) {
            constexpr int u = decltype(u_)::value;
            constexpr int slot = q * R + u;
            if constexpr (DST == DST_GLOBAL) {
                if (valid) st_stream(gdst + jq + u * STR, v[slot]);
            } else if constexpr (LIN_XNAT) {
                xq[padoff<LOGPAD>(u * STR)] = v[slot];
            } else if constexpr (DST == DST_XNAT) {
                xb[padidx<LOGPAD>(jq + u * STR)] = v[slot];
            } else if constexpr (LIN_XCHG && NS == 1) {
                xq[u] = v[slot];
            } else if constexpr (LIN_XCHG) {
                xq[padoff<LOGPAD>(u * NS)] = v[slot];
            } else {
                const int p = (jq / NS) * (NS * R) + (jq & (NS - 1)) + u * NS;
                xb[padidx<LOGPAD>(p)] = v[slot];
            }
        });
    });
}

// Split-complex ("planar") rows: real and imaginary parts live in two float arrays.  A warp request is then a run of
// contiguous 4-byte elements (128 B per array for T >= 32), still whole sectors; same butterfly maps as above.
template <int M, int T, int E, int R>
__device__ __forceinline__ void gather_planar(cf (&v)[E], const float* __restrict__ re, const float* __restrict__ im, int j, bool valid)
{
    constexpr int B = E / R, STR = M / R;
    static_for<0, B>([&](auto q_) {
        constexpr int q = decltype(q_)::value;
        const int jq = j + q * T;
        static_for<0, R>([&](auto t_) {
            constexpr int t = decltype(t_)::value;
            if (valid) v[q * R + bitrev<R>(t)] = make_float2(__ldcs(re + jq + t * STR), __ldcs(im + jq + t * STR));
        });
    });
}

template <int M, int T, int E, int R>
__device__ __forceinline__ void scatter_planar(const cf (&v)[E], float* __restrict__ re, float* __restrict__ im, int j, bool valid)
{
    constexpr int B = E / R, STR = M / R;
    static_for<0, B>([&](auto q_) {
        constexpr int q = decltype(q_)::value;
        const int jq = j + q * T;
        static_for<0, R>([&](auto u_) {
            constexpr int u = decltype(u_)::value;
            if (valid) { __stcs(re + jq + u * STR, v[q * R + u].x); __stcs(im + jq + u * STR, v[q * R + u].y); }
        });
    });
}

// Where a real-forward kernel puts bin k: the complex value, or -- AUDIO kernels (what an audio caller
// of CkFftRealForward does next) -- its squared magnitude into a float array, which halves the bytes written.
template <bool AUDIO>
__device__ __forceinline__ void put_bin(cf* __restrict__ dst, int k, cf y)
{
    if constexpr (AUDIO) __stcs(reinterpret_cast<float*>(dst) + k, fmaf(y.x, y.x, y.y * y.y));
    else                 st_stream(dst + k, y);
}

// the same through a pointer to the bin itself (float for the power spectrum, complex otherwise)
template <bool AUDIO> struct BinType { typedef cf type; };
template <> struct BinType<true> { typedef float type; };
template <bool AUDIO>
__device__ __forceinline__ void put_at(typename BinType<AUDIO>::type* __restrict__ p, cf y)
{
    if constexpr (AUDIO) __stcs(p, fmaf(y.x, y.x, y.y * y.y));
    else                 st_stream(p, y);
}

// Real-forward split on a thread's mirror-paired butterflies (see bfly_index).  v holds, per pair slot s,
// Z[p + u*STR] in block 2s and Z[pbar + u*STR] in block 2s+1 (natural u order).  Writes Y[0 .. M].
//   Y[k] = (Z[k] + conj Z[M-k]) - i W_2M^k (Z[k] - conj Z[M-k]);  with c = i W^k * diff:
//   Y[k] = sum - c,  Y[M-k] = conj(sum + c)          (the two statements of fft_real_default.cpp:30-58)
// Split factor W_2M^k of bin k = j + s*T + u*STR: a table value (CONST_TW = false), or the thread's own W_2M^j (`wj`, an
// exact table value loaded once per kernel) times the compile-time constant W_64^(s*32/E + u*32/R) -- no table traffic
// per transform, one extra rounding like the register stage twiddles.  Which one is faster depends on the registers the
// plan has left (Cfg::RTWC, by measurement).
// Addresses: bin k = j + c lives at (row + j) + c and its mirror at (row + M - j) - c with c a compile-time constant, so
// the 2 E stores and E/2 table loads of a thread hang off three pointers computed once per row (ptxas rebuilt every
// 64-bit address from k: ~ 4 integer instructions per access).
template <int M, int T, int E, int R, bool AUDIO, bool CONST_TW>
__device__ __forceinline__ void r2c_paired_epilogue(const cf (&v)[E], cf* __restrict__ dst, const cf* __restrict__ table, int sh_real,
                                                    cf wj, int j, bool valid)
{
    constexpr int B = E / R, STR = M / R;
    typedef typename BinType<AUDIO>::type bin_t;
    static_assert(B % 2 == 0, "butterflies come in mirror pairs");
    static_assert(!CONST_TW || (32 % E == 0 && 16 % R == 0), "split factors are multiples of 1/64 turn");
    bin_t* __restrict__ lo = reinterpret_cast<bin_t*>(dst) + j;            // bin j + c      at lo[c]
    bin_t* __restrict__ hi = reinterpret_cast<bin_t*>(dst) + (M - j);      // bin M - j - c  at hi[-c]
    const cf* __restrict__ tj = table + ((long long) j << sh_real);        // W_2M^(j + c)   at tj[c << sh_real]
    const unsigned tstep = table_step_bytes(sh_real);
    (void) tj; (void) tstep;
    auto emit = [&](cf z0, cf z1, int c, cf w) {                // c = k - j; w = W_2M^k; f = i w = (-w.y, w.x)
        const cf sum = make_float2(z0.x + z1.x, z0.y - z1.y);
        const cf dif = make_float2(z0.x - z1.x, z0.y + z1.y);
        const cf cc = cmul(make_float2(-w.y, w.x), dif);
        put_at<AUDIO>(lo + c, make_float2(sum.x - cc.x, sum.y - cc.y));
        put_at<AUDIO>(hi - c, make_float2(sum.x + cc.x, -(sum.y + cc.y)));
    };
    auto factor = [&](auto k64_, int c) -> cf {
        if constexpr (CONST_TW) return cmul_w64<decltype(k64_)::value>(wj);
        else                    return table_at(tj, (unsigned) c, tstep);
    };
    if (!valid) return;
    static_for<0, B / 2>([&](auto s_) {
        constexpr int s = decltype(s_)::value;
        constexpr int A = 2 * s * R, Bk = (2 * s + 1) * R;
        if constexpr (s == 0) {
            // Thread 0's first pair holds butterflies 0 (block A) and STR/2 (block B), which mirror into
            // themselves: same arithmetic, different operands -- chosen with selects, no divergent branch.
            const bool self = (j == 0);
            auto pick = [&](cf a, cf b) { return self ? b : a; };
            static_for<0, R>([&](auto u_) {
                constexpr int u = decltype(u_)::value;
                if constexpr (u < R / 2) {
                    emit(v[A + u], pick(v[Bk + R - 1 - u], v[A + (R - u) % R]), u * STR,
                         factor(Int<u * (32 / R)>{}, u * STR));                                    // self: Y[u*STR] (u = 0: Y[0], Y[M])
                } else {
                    constexpr int w = u - R / 2;
                    const int c = self ? STR / 2 + w * STR : u * STR;                              // (self: j = 0, so c is the bin itself)
                    cf f;
                    if constexpr (CONST_TW) {
                        // self: bin STR/2 + w*STR, factor W_64^(16/R + w*32/R) (wj = 1 there: the product is the constant itself)
                        constexpr int ks = 16 / R + w * (32 / R), kn = u * (32 / R);
                        f = cmul(wj, self ? make_float2(cos64(ks), -sin64(ks)) : make_float2(cos64(kn), -sin64(kn)));
                    } else {
                        f = table_at(tj, (unsigned) c, tstep);
                    }
                    emit(pick(v[A + u], v[Bk + w]), pick(v[Bk + R - 1 - u], v[Bk + R - 1 - w]), c, f);
                }
            });
            if (self) {
                const cf mid = v[A + R / 2];                                         // Z[M/2]
                put_at<AUDIO>(lo + M / 2, make_float2(2.0f * mid.x, -2.0f * mid.y));
            }
        } else {
            static_for<0, R>([&](auto u_) {
                constexpr int u = decltype(u_)::value;
                emit(v[A + u], v[Bk + R - 1 - u], s * T + u * STR, factor(Int<s * (32 / E) + u * (32 / R)>{}, s * T + u * STR));
            });
        }
    });
}

// Row-strided global access for the multi-pass tile kernels: butterfly input t of a column lives `rowstride`
// elements apart (the column is one of C adjacent columns of a [L][ncols] array).
template <int M, int T, int E, int R>
__device__ __forceinline__ void gather_rows(cf (&v)[E], const cf* __restrict__ col, long long rowstride, int j, int stream)
{
    constexpr int B = E / R, STR = M / R;
    static_for<0, B>([&](auto q_) {
        constexpr int q = decltype(q_)::value;
        const int jq = j + q * T;
        static_for<0, R>([&](auto t_) {
            constexpr int t = decltype(t_)::value;
            const cf* a = col + (long long) (jq + t * STR) * rowstride;
            v[q * R + bitrev<R>(t)] = stream ? __ldcs(a) : __ldg(a);
        });
    });
}

// Compile-time description of one kernel variant.
// PF_ = true: every group owns a dense staging buffer that a bulk asynchronous copy (cp.async.bulk)
// refills with its NEXT transform while the current one is being computed: the load of item i+1 is in
// flight from the end of item i's first gather until item i+1 starts.
//
// PF_ = 2: same idea without the second buffer: the bulk copy of the next transform lands in the exchange
// buffer itself as soon as the last gather of the current transform has drained it, so it overlaps the
// last stage's arithmetic and the stores (no extra shared memory, occupancy unchanged).
//
// PF_ = 3 (PF_SPLIT), 16384 points: one transform fills an SM's register file, so its successor can only wait in
// shared memory, and a full second row (128 KiB) does not fit next to the exchange buffer (135 KiB).  The row is
// prefetched in two halves: the lower half into a staging buffer of its own as soon as stage 0 has gathered (in
// flight for the whole transform), the upper half into the exchange buffer once the last stage has gathered.
enum Prefetch { PF_NONE = 0, PF_DOUBLE = 1, PF_INPLACE = 2, PF_SPLIT = 3 };

template <int M_, int E_, int R0_, int R1_, int R2_, int G_, bool INV_, int MODE_, int MINB_ = 1, int PF_ = PF_NONE, bool TWR_ = false,
          bool AUDIO_ = false, bool PLANAR_ = false>
struct Cfg {
    static constexpr bool PLANAR = PLANAR_;   // split-complex input and output (complex transforms, plain loads)
    static_assert(!PLANAR_ || (MODE_ == MODE_C2C && (PF_ == PF_NONE || PF_ == PF_INPLACE || PF_ == PF_SPLIT)), "planar rows: complex transforms");
    static constexpr bool AUDIO = AUDIO_;   // real forward with a fused analysis window (load) and power spectrum (store)
    static_assert(!AUDIO_ || MODE_ == MODE_R2C, "the audio front end is a real-forward kernel");
    static constexpr int PF = PF_;
    static constexpr bool TWR = TWR_;       // stage-1 twiddles from register-resident bases (two-stage plans with R1 == E)
    static constexpr int M = M_, E = E_, R0 = R0_, R1 = R1_, R2 = R2_, G = G_, MODE = MODE_, MINB = MINB_;
    static constexpr bool INV = INV_;
    static constexpr int T = M / E;
    static constexpr int THREADS = G * T;
    static constexpr int NSTAGE = R2 > 1 ? 3 : 2;
    static constexpr int RLAST = R2 > 1 ? R2 : R1;
    // real forward: split in registers on mirror-paired butterflies when the last stage gives a thread >= 2 of them
    // (measured exception: M = 512 runs faster with the shared-memory split, .93 vs .78 of the copy peak)
    static constexpr bool PAIRED = MODE_ == MODE_R2C && ((E_ / RLAST) % 2 == 0) && M_ != 512;
    // real split / twist factors W_2M^k: from the table, or the thread's own W_2M^j times compile-time constants (see
    // r2c_paired_epilogue).  By measurement (real n = 2M): forward 128 .75 -> .79, 256 .85 -> .89, 16384 .72 -> .73;
    // inverse 64 .57 -> .59, 128 .73 -> .78, 256 .73 -> .75, 8192 .79 -> .80, 16384 .71 -> .72; slower elsewhere (spills).
#ifdef CKB_RTWC_ALL      // development A/B builds: constant factors wherever the plan allows them / nowhere
    static constexpr bool RTWC = CKB_RTWC_ALL != 0 && MODE_ != MODE_C2C && 32 % E_ == 0 && (MODE_ == MODE_C2R || 16 % (R2_ > 1 ? R2_ : R1_) == 0 || (E_ / (R2_ > 1 ? R2_ : R1_)) % 2 != 0 || M_ == 512);
#else
    // round 2 (leaner addressing, registers to spare): forward 4096 .84 -> .89, 16384 .70 -> .73; inverse 2048 .91 -> .92, 16384 .53 -> .70
    static constexpr bool RTWC = MODE_ == MODE_R2C ? (M_ == 64 || M_ == 128 || M_ == 4096 || M_ == 8192 || M_ == 16384)
                               : MODE_ == MODE_C2R ? (M_ == 32 || M_ == 64 || M_ == 128 || M_ == 2048 || M_ == 4096 || M_ == 8192 || M_ == 16384) : false;
#endif
    static constexpr int LOGPAD = ilog2(R0);
    // exchange-buffer addresses as one base + compile-time offsets (see padoff).  CKB_LEGACY_SIZES: plans that keep the
    // per-element padded indices (development A/B switch)
    // Measured exception: the complex 2048-point kernel (in-place prefetch, four groups of two warps) runs at 0.96 of the copy peak
    // with the per-element indices and at 0.84-0.86 with the lean ones -- ptxas then schedules the stage-0 loads just in time
    // (96 instead of 128 registers) and the groups spend 39 % instead of 26 % of their time waiting for their row.
#ifndef CKB_LEGACY_SIZES
#define CKB_LEGACY_SIZES (MODE_ == MODE_C2C && M_ == 2048)
#endif
    static constexpr bool LIN = !(CKB_LEGACY_SIZES);
    // complex slots per group (+ slot M for the real modes).  Plans with several groups per half-warp (T = 4, 8) need a
    // group pitch of 12 resp. 8 (mod 16) 8-byte words: with the raw pitch (6 or 10 mod 16) the groups of a half-warp land on
    // each other's banks -- 4 to 8 conflicting lanes per exchange request (tools/bank_model.py, tests/test_bank_model.py).
    static constexpr int XRAW = M + (M >> LOGPAD) + 2;
    // Measured on B200 (fraction of the copy peak, raw -> conflict-free pitch): real n = 128 forward 0.79 -> 0.86, inverse
    // 0.78 -> 0.88; real n = 256 inverse 0.75 -> 0.97 -- the split / twist passes go through the buffer, too.  The complex
    // 128-point kernel measured 0.91 -> 0.85 with the wider pitch and keeps the raw one.
    static constexpr bool WIDE_PITCH = MODE_ != MODE_C2C;
    static constexpr int XBUF = !WIDE_PITCH ? XRAW
                              : (M / E_) == 8 ? XRAW + (8 + 16 - XRAW % 16) % 16 : (M / E_) == 4 ? XRAW + (12 + 16 - XRAW % 16) % 16 : XRAW;
    static constexpr int LUT1 = TWR_ ? 0 : (R1 - 1) * R0;       // stage 1: Ns = R0
    static constexpr bool LUT2_SMEM = NSTAGE == 3 && (R2 - 1) * R0 * R1 <= 4096;
    static constexpr bool POW2 = NSTAGE == 3 && !LUT2_SMEM;        // last-stage twiddles from register power bases
    static constexpr int NPOW = POW2 ? (E / R2) * ilog2(R2) : 1;
    static constexpr int LUT2 = LUT2_SMEM ? (R2 - 1) * R0 * R1 : 0;   // stage 2: Ns = R0*R1
    static constexpr int XSLOTS = XBUF;                                      // complex slots of the exchange buffer
    static constexpr int GROUP_SLOTS = XSLOTS + (PF == PF_DOUBLE ? M : (PF == PF_SPLIT ? M / 2 : 0));   // exchange (+ staging) buffer
    // (Real inverse with split prefetch was built and measured in round 2 -- the twist evaluated straight from the two raw half
    // rows, every pair by both of its owners: correct, .87 -> .89 at 4096 points, .78 -> .54 / .78 at 8192, .72 -> .67 at 16384,
    // and no longer bit-identical to the pairwise kernels that take unaligned and last rows; removed again, see DESIGN.md.)
    static_assert(PF_ != PF_SPLIT || (R2_ > 1 && (MODE_ == MODE_C2C || (MODE_ == MODE_R2C && (E_ / R2_) % 2 == 0))),
                  "split prefetch: three-stage plans whose last stage leaves the exchange buffer idle");
    static constexpr int SMEM_BYTES = 8 * (LUT1 + LUT2 + G * GROUP_SLOTS) + (PF ? 16 * G : 0);   // + two mbarriers per group
    static_assert(PF != PF_DOUBLE || MODE_ != MODE_C2R, "C2R prefetches in place (see the C2R prologue)");
    static_assert(PF != PF_INPLACE || MODE_ != MODE_R2C || ((E_ / (R2_ > 1 ? R2_ : R1_)) % 2 == 0 && M_ != 512),
                  "in-place prefetch: the shared-memory real epilogue still owns the buffer");
    static_assert(!TWR_ || R1_ == E_, "TWR: stage 1 must be one butterfly per thread (m = j mod R0 is loop invariant)");
    static_assert(R0 * R1 * R2 == M, "radices must multiply to the transform length");
    static_assert(R0 <= E && R1 <= E && R2 <= E, "a butterfly must fit one thread");
    static_assert(T <= 32 || G <= 15, "named barriers 1..15");
    static_assert(THREADS <= 1024, "CTA too large");
};

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::MINB) fft_kernel(const KernelParams p)
{
    constexpr int M = C::M, E = C::E, T = C::T, G = C::G, R0 = C::R0, R1 = C::R1, R2 = C::R2;
    constexpr int LOGPAD = C::LOGPAD, MODE = C::MODE;
    constexpr bool INV = C::INV;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf* lut1 = reinterpret_cast<cf*>(smem_raw);
    cf* lut2 = lut1 + C::LUT1;
    const int tid = threadIdx.x;
    const int g = tid / T;
    const int j = tid % T;
    cf* xb = lut2 + C::LUT2 + g * C::GROUP_SLOTS;
    cf* inb = (C::PF == PF_DOUBLE || C::PF == PF_SPLIT) ? xb + C::XSLOTS : xb;     // staging buffer of the bulk copies
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(lut2 + C::LUT2 + G * C::GROUP_SLOTS) + 2 * g;
    unsigned long long* mbar_hi = mbar + 1;                             // PF_SPLIT: barrier of the upper half row
    unsigned long long l2pol = 0;
    if constexpr (C::PF) {
        if (j == 0) { mbar_init(mbar, 1); mbar_init(mbar_hi, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        l2pol = l2_evict_first_policy();
    }

    // stage twiddle LUTs, [t-1][m] so that a thread's R-1 loads are immediate offsets from one base
    {
        const int sh1 = p.log2_nt - ilog2(R0 * R1);
        for (int i = tid; i < C::LUT1; i += C::THREADS)
            lut1[i] = table_w(p.table, ((i / R0 + 1) * (i % R0)) << sh1, INV);
        if constexpr (C::LUT2 > 0) {
            constexpr int NS2 = R0 * R1;
            const int sh2 = p.log2_nt - ilog2(M);
            for (int i = tid; i < C::LUT2; i += C::THREADS)
                lut2[i] = table_w(p.table, ((i / NS2 + 1) * (i % NS2)) << sh2, INV);
        }
    }
    __syncthreads();

    const int sh_last = p.log2_nt - ilog2(M);        // W_M^k   = table[k << sh_last]
    const int sh_real = p.log2_nt - ilog2(2 * M);    // W_2M^k  = table[k << sh_real]  (real modes)

    // real modes: the thread's own split / twist factor W_2M^j; bin j + i*T takes it times W_64^(i*32/E)
    cf wj = make_float2(1.f, 0.f);
    if constexpr (C::RTWC) wj = __ldg(p.table + (j << sh_real));
    static_assert(!C::RTWC || 32 % E == 0, "split / twist factors are multiples of 1/64 turn");
    const cf* __restrict__ tab_j = p.table + ((long long) j << sh_real);
    const unsigned tab_step = table_step_bytes(sh_real);
    auto real_factor = [&](auto i_, int k) -> cf {                 // forward W_2M^k, k = j + i*T
        if constexpr (C::RTWC) return cmul_w64<decltype(i_)::value * (32 / E)>(wj);
        else                   return table_at(tab_j, (unsigned) (decltype(i_)::value * T), tab_step);
    };
    (void) tab_j; (void) tab_step;
    (void) real_factor;

    cf twb[TwSplit<R1>::NB];
    if constexpr (C::TWR) load_tw_bases<R1, INV>(twb, p.table, j & (R0 - 1), p.log2_nt - ilog2(R0 * R1));
    cf pw[C::NPOW];
    if constexpr (C::POW2) load_pow_bases<M, T, E, R2, INV, C::PAIRED>(pw, p.table, j, p.log2_nt - ilog2(M));

    unsigned phase = 0;
    // One thread per group refills the staging buffer with the group's next transform and every thread of the group
    // acquires the data by waiting on the group's own mbarrier.  When two groups share a warp (T = 16) the two halves
    // of the warp therefore wait on different barriers.  compute-sanitizer's racecheck reports that divergent wait as
    // a *warning* (no error; memcheck is clean): each thread does perform the acquiring try_wait on the barrier its
    // copy completes on, which is what the PTX memory model asks for.  A variant with one barrier per warp is
    // racecheck-silent but measured 5 % slower at 256 and 512 points (the half that gets its row first can no longer
    // start gathering), so the per-group barriers stay.
    // C2R rows hold M+1 complex values and start on 8-byte boundaries only: the copy starts at the 16-byte boundary
    // at or below the row and takes M+2 values, so Y[k] lands in slot k + pad (pad = 0 or 1).  The launcher keeps a
    // last row whose over-copy would leave the array out of this kernel.
    constexpr unsigned kCopyBytes = (MODE == MODE_C2R ? M + 2 : M) * 8;
    auto issue_row = [&](long long row) {
        if constexpr (C::PLANAR) {
            // split-complex rows: two copies, the real parts into floats [0, M) of the buffer, the imaginary parts behind them
            if (j == 0 && row < p.batch) {
                float* fb = reinterpret_cast<float*>(inb);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(mbar, M * 8);
                bulk_load(fb, reinterpret_cast<const float*>(p.in) + row * p.in_stride, M * 4, mbar, l2pol);
                bulk_load(fb + M, p.in_im + row * p.in_stride, M * 4, mbar, l2pol);
            }
        } else
        if (j == 0 && row < p.batch) {
            const cf* rsrc = p.in + row * p.in_stride;
            if constexpr (MODE == MODE_C2R) rsrc = reinterpret_cast<const cf*>(reinterpret_cast<uintptr_t>(rsrc) & ~uintptr_t(15));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(mbar, kCopyBytes);
            bulk_load(inb, rsrc, kCopyBytes, mbar, l2pol);
        }
    };
    auto issue_next = [&](long long item) { issue_row(item + (long long) gridDim.x * G); };
    (void) issue_next;
    // PF_SPLIT: half 0 = points [0, M/2) -> staging buffer, half 1 = points [M/2, M) -> exchange buffer
    auto issue_half = [&](long long row, int half) {
        if (j == 0 && row < p.batch) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(half ? mbar_hi : mbar, M * 4);
            if constexpr (C::PLANAR)      // split-complex rows: the two "halves" are the plane of real parts and the plane of imaginary parts
                bulk_load(half ? xb : inb, half ? p.in_im + row * p.in_stride : reinterpret_cast<const float*>(p.in) + row * p.in_stride,
                          M * 4, half ? mbar_hi : mbar, l2pol);
            else
            bulk_load(half ? xb : inb, p.in + row * p.in_stride + half * (M / 2), M * 4, half ? mbar_hi : mbar, l2pol);
        }
    };
    (void) issue_half;
    if constexpr (C::PF == PF_SPLIT) {
        issue_half((long long) blockIdx.x * G + g, 0);
        issue_half((long long) blockIdx.x * G + g, 1);
    } else if constexpr (C::PF != PF_NONE) issue_row((long long) blockIdx.x * G + g);

    for (long long base = (long long) blockIdx.x * G; base < p.batch; base += (long long) gridDim.x * G) {
        const long long item = base + g;
        const bool valid = item < p.batch;
        const cf* __restrict__ src = p.in + item * p.in_stride;
        cf* __restrict__ dst = C::AUDIO ? reinterpret_cast<cf*>(reinterpret_cast<float*>(p.out) + item * p.out_stride)
                                        : p.out + item * p.out_stride;
        cf v[E];

        // ---- stage 0 (Ns = 1, no twiddles) ----
        if constexpr (MODE == MODE_C2R && C::PF == PF_INPLACE) {
            // The row was bulk-copied into the (dense) buffer; twist it in place -- a thread owns both slots of a
            // mirror pair -- then gather stage 0 from the dense layout.
            if (valid) mbar_wait(mbar, phase);
            phase ^= 1u;
            cf* yb = xb + ((reinterpret_cast<uintptr_t>(src) >> 3) & 1);        // Y[k] sits in yb[k]
            static_for<0, E / 2>([&](auto i_) {
                const int k = j + decltype(i_)::value * T;       // 0 .. M/2-1
                const cf y0 = yb[k], y1 = yb[M - k];
                const cf w = real_factor(i_, k);                 // forward W_2M^k
                const cf sum = make_float2(y0.x + y1.x, y0.y - y1.y);
                const cf dif = make_float2(y0.x - y1.x, y0.y + y1.y);
                const cf c = cmul(make_float2(w.y, w.x), dif);
                yb[k] = make_float2(sum.x + c.x, sum.y + c.y);
                if (k != 0) yb[M - k] = make_float2(sum.x - c.x, -(sum.y - c.y));
            });
            if (j == 0) {
                const cf y = yb[M / 2];
                yb[M / 2] = make_float2(2.0f * y.x, -2.0f * y.y);
            }
            group_sync<T>(g);
            stage_gather<M, T, E, R0, LOGPAD, SRC_INBUF>(v, src, yb, j, valid);
            group_sync<T>(g);
        } else if constexpr (MODE == MODE_C2R) {
            // twist (fft_real_default.cpp:65-111): T[k] = (Y[k] + conj Y[M-k]) + i conj(W_2M^k) (Y[k] - conj Y[M-k]),
            // computed pairwise: with c = f * diff,  T[k] = sum + c,  T[M-k] = conj(sum - c).
            static_for<0, E / 2>([&](auto i_) {
                const int k = j + decltype(i_)::value * T;       // 0 .. M/2-1
                cf y0 = make_float2(0.f, 0.f), y1 = y0;
                if (valid) { y0 = ld_stream(src + k); y1 = ld_stream(src + (M - k)); }
                const cf w = real_factor(i_, k);                 // forward W_2M^k; e = conj(w); f = i e = (w.y, w.x)
                const cf sum = make_float2(y0.x + y1.x, y0.y - y1.y);
                const cf dif = make_float2(y0.x - y1.x, y0.y + y1.y);
                const cf c = cmul(make_float2(w.y, w.x), dif);
                xb[padidx<LOGPAD>(k)] = make_float2(sum.x + c.x, sum.y + c.y);
                if (k != 0) xb[padidx<LOGPAD>(M - k)] = make_float2(sum.x - c.x, -(sum.y - c.y));
            });
            if (j == 0) {
                cf y = valid ? ld_stream(src + M / 2) : make_float2(0.f, 0.f);
                xb[padidx<LOGPAD>(M / 2)] = make_float2(2.0f * y.x, -2.0f * y.y);
            }
            group_sync<T>(g);
            stage_gather<M, T, E, R0, LOGPAD, SRC_XBUF, false, C::LIN>(v, src, xb, j, valid);
            group_sync<T>(g);
        } else if constexpr (C::PF == PF_SPLIT) {
            static_assert(E == R0, "one stage-0 butterfly per thread");
            if (valid) { mbar_wait(mbar, phase); mbar_wait(mbar_hi, phase); }
            phase ^= 1u;
            static_for<0, R0>([&](auto t_) {
                constexpr int t = decltype(t_)::value;
                if constexpr (C::PLANAR)
                    v[bitrev<R0>(t)] = make_float2(reinterpret_cast<const float*>(inb)[j + t * T], reinterpret_cast<const float*>(xb)[j + t * T]);
                else
                    v[bitrev<R0>(t)] = t < R0 / 2 ? inb[j + t * T] : xb[j + (t - R0 / 2) * T];
            });
            group_sync<T>(g);                      // both halves are in registers
            issue_half(item + (long long) gridDim.x * G, 0);
        } else if constexpr (C::PF != PF_NONE) {
            if (valid) mbar_wait(mbar, phase);
            phase ^= 1u;
            if constexpr (C::PLANAR) {
                const float* fb = reinterpret_cast<const float*>(inb);
                static_for<0, E / R0>([&](auto q_) {
                    constexpr int q = decltype(q_)::value;
                    const int jq = j + q * T;
                    static_for<0, R0>([&](auto t_) {
                        constexpr int t = decltype(t_)::value;
                        v[q * R0 + bitrev<R0>(t)] = make_float2(fb[jq + t * (M / R0)], fb[M + jq + t * (M / R0)]);
                    });
                });
            } else {
                stage_gather<M, T, E, R0, LOGPAD, SRC_INBUF>(v, src, inb, j, valid);
            }
            group_sync<T>(g);                      // every thread of the group has drained the staging buffer
            if constexpr (C::PF == PF_DOUBLE) issue_next(item);
        } else if constexpr (C::PLANAR) {
            gather_planar<M, T, E, R0>(v, reinterpret_cast<const float*>(p.in) + item * p.in_stride, p.in_im + item * p.in_stride, j, valid);
        } else {
            stage_gather<M, T, E, R0, LOGPAD, SRC_GLOBAL>(v, src, xb, j, valid);
        }
        if constexpr (C::AUDIO) {
            // analysis window fused into the load: sample pair (2e, 2e+1) is complex element e of the row
            if (p.window != nullptr) {
                static_assert(E == R0, "one stage-0 butterfly per thread");
                static_for<0, R0>([&](auto t_) {
                    constexpr int t = decltype(t_)::value;
                    const cf w = __ldg(p.window + j + t * T);
                    v[bitrev<R0>(t)].x *= w.x;
                    v[bitrev<R0>(t)].y *= w.y;
                });
            }
        }
        stage_math<T, E, R0, 1, INV, TW_NONE>(v, nullptr, p.table, 0, j);
        constexpr bool PAIR1 = C::PAIRED && C::NSTAGE == 2;     // last stage of a two-stage real-forward plan
        stage_scatter<M, T, E, R0, 1, LOGPAD, DST_XCHG, C::LIN>(v, dst, xb, j, valid);
        group_sync<T>(g);
        // ---- stage 1 (Ns = R0) ----
        stage_gather<M, T, E, R1, LOGPAD, SRC_XBUF, PAIR1, C::LIN>(v, src, xb, j, valid);
        if constexpr (C::PF == PF_INPLACE && C::NSTAGE == 2) { group_sync<T>(g); issue_next(item); }
        if constexpr (C::TWR) stage_math_regs<E, R1, INV>(v, twb);
        else                  stage_math<T, E, R1, R0, INV, TW_LUT, PAIR1>(v, lut1, p.table, 0, j);
        if constexpr (C::NSTAGE == 2) {
            if constexpr (PAIR1) {
                r2c_paired_epilogue<M, T, E, R1, C::AUDIO, C::RTWC>(v, dst, p.table, sh_real, wj, j, valid);
            } else if constexpr (MODE == MODE_R2C) {
                group_sync<T>(g);
                stage_scatter<M, T, E, R1, R0, LOGPAD, DST_XNAT, C::LIN>(v, dst, xb, j, valid);
            } else if constexpr (C::PLANAR) {
                scatter_planar<M, T, E, R1>(v, reinterpret_cast<float*>(p.out) + item * p.out_stride, p.out_im + item * p.out_stride, j, valid);
            } else {
                stage_scatter<M, T, E, R1, R0, LOGPAD, DST_GLOBAL>(v, dst, xb, j, valid);
            }
        } else {
            group_sync<T>(g);
            stage_scatter<M, T, E, R1, R0, LOGPAD, DST_XCHG, C::LIN>(v, dst, xb, j, valid);
            group_sync<T>(g);
            // ---- stage 2 (Ns = R0*R1) ----
            stage_gather<M, T, E, R2, LOGPAD, SRC_XBUF, C::PAIRED, C::LIN>(v, src, xb, j, valid);
            if constexpr (C::PF == PF_INPLACE) { group_sync<T>(g); issue_next(item); }
            if constexpr (C::PF == PF_SPLIT) { group_sync<T>(g); issue_half(item + (long long) gridDim.x * G, 1); }
            if constexpr (C::POW2) stage_math_pow<E, R2, INV>(v, pw);
            else                   stage_math<T, E, R2, R0 * R1, INV, TW_LUT, C::PAIRED>(v, lut2, p.table, sh_last, j);
            if constexpr (C::PAIRED) {
                r2c_paired_epilogue<M, T, E, R2, C::AUDIO, C::RTWC>(v, dst, p.table, sh_real, wj, j, valid);
            } else if constexpr (MODE == MODE_R2C) {
                group_sync<T>(g);
                stage_scatter<M, T, E, R2, R0 * R1, LOGPAD, DST_XNAT, C::LIN>(v, dst, xb, j, valid);
            } else if constexpr (C::PLANAR) {
                scatter_planar<M, T, E, R2>(v, reinterpret_cast<float*>(p.out) + item * p.out_stride, p.out_im + item * p.out_stride, j, valid);
            } else {
                stage_scatter<M, T, E, R2, R0 * R1, LOGPAD, DST_GLOBAL>(v, dst, xb, j, valid);
            }
        }

        if constexpr (MODE == MODE_R2C && !C::PAIRED) {
            // split (fft_real_default.cpp:23-62): Y[k] = (Z[k] + conj Z[M-k]) - i W_2M^k (Z[k] - conj Z[M-k]);
            // pairwise with c = f * diff:  Y[k] = sum - c,  Y[M-k] = conj(sum + c);  Z[M] == Z[0].
            group_sync<T>(g);
            static_for<0, E / 2>([&](auto i_) {
                const int k = j + decltype(i_)::value * T;       // 0 .. M/2-1
                const cf z0 = xb[padidx<LOGPAD>(k)];
                const cf z1 = xb[padidx<LOGPAD>((M - k) & (M - 1))];
                const cf w = real_factor(i_, k);                 // e = W_2M^k; f = i e = (-w.y, w.x)
                const cf sum = make_float2(z0.x + z1.x, z0.y - z1.y);
                const cf dif = make_float2(z0.x - z1.x, z0.y + z1.y);
                const cf c = cmul(make_float2(-w.y, w.x), dif);
                if (valid) {
                    put_bin<C::AUDIO>(dst, k, make_float2(sum.x - c.x, sum.y - c.y));
                    put_bin<C::AUDIO>(dst, M - k, make_float2(sum.x + c.x, -(sum.y + c.y)));
                }
            });
            if (j == 0 && valid) {
                const cf z = xb[padidx<LOGPAD>(M / 2)];
                put_bin<C::AUDIO>(dst, M / 2, make_float2(2.0f * z.x, -2.0f * z.y));
            }
        }
        // the next iteration's first scatter must not overtake this iteration's last gather
        // (the in-place prefetch already put a group barrier behind that gather)
        if constexpr (C::PF != PF_INPLACE && C::PF != PF_SPLIT) group_sync<T>(g);
    }
}

}  // namespace ckb
