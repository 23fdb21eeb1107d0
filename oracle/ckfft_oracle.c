/*
 * ckfft_oracle.c -- CPU restatement of the ckfft transform hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/,
 * __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of
 * bench.py may load it.  The product (ckfft_b200/csrc) never links, loads or
 * calls anything in oracle/.
 *
 * What it restates (paths relative to the reference tree):
 *   - twiddle table         src/ckfft/context.cpp:90-105
 *   - complex driver        src/ckfft/fft.cpp:13-46
 *   - radix-4 DIT kernel    src/ckfft/fft_default.cpp:12-266
 *   - complex helpers       src/ckfft/math_util.h:17-33
 *   - real drivers          src/ckfft/fft_real.cpp:13-105
 *   - real split / twist    src/ckfft/fft_real_default.cpp:13-114
 *   - argument checks       src/ckfft/ckfft.cpp:36-114
 *
 * The reference is a recursive decimation-in-time radix-4 transform.  This
 * file states the same arithmetic ITERATIVELY: a base-4 digit-reversed leaf
 * pass (leaves of 4 or 8 points) followed by log4 combine passes.  Every
 * floating-point operation has the same operands, in the same order, as the
 * reference's scalar path, so when both are compiled without FMA contraction
 * (-ffp-contract=off) the outputs are BIT-IDENTICAL.  tests/test_oracle.py
 * pins that against oracle/_ref/libckfft_ref.so (the unmodified reference
 * compiled here) and against the committed golden vectors in tests/golden/.
 *
 * Parity status: PINNED (bit-exact vs. the compiled reference for N=1..2^20
 * and vs. tests/golden/ckfft_golden.npz produced from the reference's own
 * fixture src/test/input.txt).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct { float re, im; } cpx;

/* math_util.h:12-15 */
static int is_pow2(unsigned int x) { return x != 0 && !(x & (x - 1)); }

/* math_util.h:29-33: (a.r*b.r - a.i*b.i, a.i*b.r + a.r*b.i), a = data, b = twiddle */
static inline cpx cmul(cpx a, cpx b)
{
    cpx o;
    o.re = a.re * b.re - a.im * b.im;
    o.im = a.im * b.re + a.re * b.im;
    return o;
}
static inline cpx cadd(cpx a, cpx b) { cpx o = { a.re + b.re, a.im + b.im }; return o; }
static inline cpx csub(cpx a, cpx b) { cpx o = { a.re - b.re, a.im - b.im }; return o; }

/*
 * context.cpp:90-105.  theta is formed in float exactly as the reference does:
 * (-2.0f * (float)M_PI * i) / nmax, left-to-right, i and nmax converted to float.
 * sign = +1 gives the inverse table (c, -s).
 */
void ckfft_oracle_twiddles(int nmax, int inverse, float* table /* 2*nmax floats */)
{
    for (int i = 0; i < nmax; ++i) {
        float theta = -2.0f * (float) M_PI * i / nmax;
        float c = cosf(theta);
        float s = sinf(theta);
        table[2 * i] = c;
        table[2 * i + 1] = inverse ? -s : s;
    }
}

/* one radix-4 output quartet, fft_default.cpp:224-256 (and :53-66, :119-132, :152-166) */
static inline void quartet(cpx sum02, cpx diff02, cpx sum13, cpx diff13, int inverse,
                           cpx* o0, cpx* o1, cpx* o2, cpx* o3)
{
    *o0 = cadd(sum02, sum13);
    *o2 = csub(sum02, sum13);
    if (inverse) {
        o1->re = diff02.re - diff13.im;  o1->im = diff02.im + diff13.re;
        o3->re = diff02.re + diff13.im;  o3->im = diff02.im - diff13.re;
    } else {
        o1->re = diff02.re + diff13.im;  o1->im = diff02.im - diff13.re;
        o3->re = diff02.re - diff13.im;  o3->im = diff02.im + diff13.re;
    }
}

/* reverse the low `digits` base-4 digits of b */
static unsigned rev4(unsigned b, int digits)
{
    unsigned r = 0;
    for (int d = 0; d < digits; ++d) { r = (r << 2) | (b & 3u); b >>= 2; }
    return r;
}

/*
 * Complex transform of n >= 4 points; `table` is a full-circle table of tmax
 * entries (tmax >= n, both powers of two), i.e. the reference's expTable with
 * expTableStride = tmax / n (fft.cpp:34-35).
 *
 * Recursion unrolled: the reference's call tree for n = L * 4^m (L = 4 or 8)
 * visits leaf number b (output offset b*L) with input offset equal to the
 * base-4 digit reversal of b and input stride 4^m (fft_default.cpp:176-185
 * applied m times), then combines blocks of 4L, 16L, ..., n.
 */
static void fft_core(const cpx* in, cpx* out, int n, int inverse, const cpx* table, int tmax)
{
    int log2n = 0;
    while ((1 << log2n) < n) ++log2n;
    const int leaf = (log2n & 1) ? 8 : 4;
    int m = 0;                       /* number of combine levels */
    for (int c = leaf; c < n; c <<= 2) ++m;
    const int lstride = n / leaf;    /* 4^m */
    const int nleaf = n / leaf;

    for (int b = 0; b < nleaf; ++b) {
        const cpx* src = in + rev4((unsigned) b, m);
        cpx* o = out + (size_t) b * leaf;
        if (leaf == 4) {
            /* fft_default.cpp:22-67 */
            cpx x0 = src[0], x1 = src[lstride], x2 = src[2 * lstride], x3 = src[3 * lstride];
            quartet(cadd(x0, x2), csub(x0, x2), cadd(x1, x3), csub(x1, x3), inverse,
                    &o[0], &o[1], &o[2], &o[3]);
        } else {
            /* fft_default.cpp:68-167: four radix-2 on (q, q+4), then one radix-4 combine of n=2 */
            cpx t[8];
            for (int q = 0; q < 4; ++q) {
                cpx a = src[q * lstride], c = src[(q + 4) * lstride];
                t[2 * q] = cadd(a, c);
                t[2 * q + 1] = csub(a, c);
            }
            /* i = 0: no twiddles */
            quartet(cadd(t[0], t[4]), csub(t[0], t[4]), cadd(t[2], t[6]), csub(t[2], t[6]), inverse,
                    &o[0], &o[2], &o[4], &o[6]);
            /* i = 1: W8^1, W8^2, W8^3 = table[q * tmax/8] */
            {
                const int s8 = tmax / 8;
                cpx f1 = cmul(t[3], table[s8]);
                cpx f2 = cmul(t[5], table[2 * s8]);
                cpx f3 = cmul(t[7], table[3 * s8]);
                quartet(cadd(t[1], f2), csub(t[1], f2), cadd(f1, f3), csub(f1, f3), inverse,
                        &o[1], &o[3], &o[5], &o[7]);
            }
        }
    }

    /* combine passes, fft_default.cpp:187-266 */
    for (int c = leaf * 4; c <= n; c <<= 2) {
        const int q = c / 4;
        const int ts = tmax / c;     /* stride * expTableStride for a block of size c */
        for (int base = 0; base < n; base += c) {
            cpx* o0 = out + base;
            cpx* o1 = o0 + q;
            cpx* o2 = o1 + q;
            cpx* o3 = o2 + q;
            for (int i = 0; i < q; ++i) {
                cpx f1 = cmul(o1[i], table[(size_t) i * ts]);
                cpx f2 = cmul(o2[i], table[(size_t) i * ts * 2]);
                cpx f3 = cmul(o3[i], table[(size_t) i * ts * 3]);
                cpx f0 = o0[i];
                quartet(cadd(f0, f2), csub(f0, f2), cadd(f1, f3), csub(f1, f3), inverse,
                        &o0[i], &o1[i], &o2[i], &o3[i]);
            }
        }
    }
}

/* fft.cpp:13-46 */
static void fft_any(const cpx* in, cpx* out, int n, int inverse, const cpx* table, int tmax)
{
    if (n == 1) {
        out[0] = in[0];
    } else if (n == 2) {
        cpx a = in[0], b = in[1];
        out[0] = cadd(a, b);
        out[1] = csub(a, b);
    } else {
        fft_core(in, out, n, inverse, table, tmax);
    }
}

/* fft_real.cpp:13-60 and fft_real_default.cpp:13-63; out has n/2+1 entries */
static void real_forward(const float* in, cpx* out, int n, const cpx* table, int tmax)
{
    if (n == 1) {
        out[0].re = in[0] * 2.0f;  out[0].im = 0.0f;
    } else if (n == 2) {
        out[0].re = (in[0] + in[1]) * 2.0f;  out[0].im = 0.0f;
        out[1].re = (in[0] - in[1]) * 2.0f;  out[1].im = 0.0f;
    } else if (n == 4) {
        float sum02 = (in[0] + in[2]) * 2.0f, diff02 = (in[0] - in[2]) * 2.0f;
        float sum13 = (in[1] + in[3]) * 2.0f, diff13 = (in[1] - in[3]) * 2.0f;
        out[0].re = sum02 + sum13;  out[0].im = 0.0f;
        out[1].re = diff02;         out[1].im = -diff13;
        out[2].re = sum02 - sum13;  out[2].im = 0.0f;
        /* the reference also stores out[3] = (diff02, +diff13), one element past
         * the documented n/2+1 (fft_real.cpp:46-47); not reproduced. */
    } else {
        const int h = n / 2, qn = n / 4;
        const int ts = tmax / n;
        fft_core((const cpx*) in, out, h, 0, table, tmax);
        out[h] = out[0];
        for (int i = 0; i < qn; ++i) {
            cpx z0 = out[i], z1 = out[h - i];
            cpx sum, diff, f, c;
            sum.re = z0.re + z1.re;   sum.im = z0.im - z1.im;
            diff.re = z0.re - z1.re;  diff.im = z0.im + z1.im;
            f.re = -table[(size_t) i * ts].im;  f.im = table[(size_t) i * ts].re;
            c = cmul(f, diff);
            out[i] = csub(sum, c);

            diff.re = -diff.re;
            sum.im = -sum.im;
            f.re = -table[(size_t) (h - i) * ts].im;  f.im = table[(size_t) (h - i) * ts].re;
            c = cmul(f, diff);
            out[h - i] = csub(sum, c);
        }
        out[qn].re = out[qn].re * 2.0f;
        out[qn].im = -out[qn].im * 2.0f;
    }
}

/* fft_real.cpp:62-105 and fft_real_default.cpp:65-114; in has n/2+1 entries, tmp n/2+1 scratch */
static void real_inverse(const cpx* in, float* out, cpx* tmp, int n, const cpx* table, int tmax)
{
    if (n == 1) {
        out[0] = in[0].re;
    } else if (n == 2) {
        out[0] = in[0].re + in[1].re;
        out[1] = in[0].re - in[1].re;
    } else if (n == 4) {
        float sum02_r = in[0].re + in[2].re;
        float sum13_r = 2.0f * in[1].re;
        cpx diff02 = csub(in[0], in[2]);
        float diff13_i = 2.0f * in[1].im;
        out[0] = sum02_r + sum13_r;
        out[1] = diff02.re - diff13_i;
        out[2] = sum02_r - sum13_r;
        out[3] = diff02.re + diff13_i;
    } else {
        const int h = n / 2, qn = n / 4;
        const int ts = tmax / n;
        for (int i = 0; i < qn; ++i) {
            cpx z0 = in[i], z1 = in[h - i];
            cpx sum, diff, f, c;
            sum.re = z0.re + z1.re;   sum.im = z0.im - z1.im;
            diff.re = z0.re - z1.re;  diff.im = z0.im + z1.im;
            f.re = -table[(size_t) i * ts].im;  f.im = table[(size_t) i * ts].re;
            c = cmul(f, diff);
            tmp[i] = cadd(sum, c);

            diff.re = -diff.re;
            sum.im = -sum.im;
            f.re = -table[(size_t) (h - i) * ts].im;  f.im = table[(size_t) (h - i) * ts].re;
            c = cmul(f, diff);
            tmp[h - i] = cadd(sum, c);
        }
        tmp[qn].re = in[qn].re * 2.0f;
        tmp[qn].im = -in[qn].im * 2.0f;
        fft_core(tmp, (cpx*) out, h, 1, table, tmax);
    }
}

/* ------------------------------------------------------------------------ */
/* exported entry points (plain C, ctypes-friendly)                          */
/* ------------------------------------------------------------------------ */

typedef struct {
    int nmax;
    cpx* fwd;   /* NULL unless the Forward bit was requested (context.cpp:73-88) */
    cpx* inv;
} oracle_ctx;

/* ckfft.cpp:14-34 argument checks; direction bits as in inc/ckfft/ckfft.h:21-27 */
void* ckfft_oracle_init(int nmax, int direction)
{
    if (nmax <= 0 || !is_pow2((unsigned) nmax)) return NULL;
    if (direction != 1 && direction != 2 && direction != 3) return NULL;
    oracle_ctx* c = (oracle_ctx*) calloc(1, sizeof(oracle_ctx));
    if (!c) return NULL;
    c->nmax = nmax;
    if (direction & 1) {
        c->fwd = (cpx*) malloc(sizeof(cpx) * (size_t) nmax);
        ckfft_oracle_twiddles(nmax, 0, (float*) c->fwd);
    }
    if (direction & 2) {
        c->inv = (cpx*) malloc(sizeof(cpx) * (size_t) nmax);
        ckfft_oracle_twiddles(nmax, 1, (float*) c->inv);
    }
    return c;
}

void ckfft_oracle_shutdown(void* h)
{
    oracle_ctx* c = (oracle_ctx*) h;
    if (!c) return;
    free(c->fwd);
    free(c->inv);
    free(c);
}

/* ckfft.cpp:78-114.  n <= 0 is rejected (the reference's unsigned isPowerOfTwo lets INT_MIN through). */
int ckfft_oracle_complex(void* h, int n, const float* in, float* out, int inverse)
{
    oracle_ctx* c = (oracle_ctx*) h;
    const cpx* table = c ? (inverse ? c->inv : c->fwd) : NULL;
    if (!c || !table) return 0;
    if (n <= 0 || !is_pow2((unsigned) n) || n > c->nmax) return 0;
    if (!in || !out || in == out) return 0;
    fft_any((const cpx*) in, (cpx*) out, n, inverse, table, c->nmax);
    return 1;
}

/* ckfft.cpp:36-53 */
int ckfft_oracle_real_forward(void* h, int n, const float* in, float* out)
{
    oracle_ctx* c = (oracle_ctx*) h;
    if (!c || !c->fwd) return 0;
    if (n <= 0 || !is_pow2((unsigned) n) || n > c->nmax) return 0;
    if (!in || !out || (const void*) in == (const void*) out) return 0;
    real_forward(in, (cpx*) out, n, c->fwd, c->nmax);
    return 1;
}

/* ckfft.cpp:55-76 */
int ckfft_oracle_real_inverse(void* h, int n, const float* in, float* out, float* tmp)
{
    oracle_ctx* c = (oracle_ctx*) h;
    if (!tmp) return 0;
    if (!c || !c->inv) return 0;
    if (n <= 0 || !is_pow2((unsigned) n) || n > c->nmax) return 0;
    if (!in || !out || (const void*) in == (const void*) out) return 0;
    real_inverse((const cpx*) in, out, (cpx*) tmp, n, c->inv, c->nmax);
    return 1;
}

/*
 * Batched helpers (contiguous transforms, OpenMP over the batch when built
 * with -fopenmp).  Used by tests and by bench.py's cpu_baseline "port" leg.
 * One shared context is legal: contexts hold no state (inc/ckfft/ckfft.h:39-41).
 */
int ckfft_oracle_complex_batch(void* h, int n, const float* in, float* out, long batch, int inverse)
{
    int ok = 1;
#pragma omp parallel for schedule(static) reduction(&: ok)
    for (long b = 0; b < batch; ++b)
        ok &= ckfft_oracle_complex(h, n, in + (size_t) b * 2 * n, out + (size_t) b * 2 * n, inverse);
    return ok;
}

int ckfft_oracle_real_forward_batch(void* h, int n, const float* in, float* out, long batch)
{
    const size_t ostride = 2 * ((size_t) n / 2 + 1);
    int ok = 1;
#pragma omp parallel for schedule(static) reduction(&: ok)
    for (long b = 0; b < batch; ++b)
        ok &= ckfft_oracle_real_forward(h, n, in + (size_t) b * n, out + (size_t) b * ostride);
    return ok;
}

int ckfft_oracle_real_inverse_batch(void* h, int n, const float* in, float* out, long batch)
{
    const size_t istride = 2 * ((size_t) n / 2 + 1);
    int ok = 1;
#pragma omp parallel
    {
        float* tmp = (float*) malloc(sizeof(float) * istride);
#pragma omp for schedule(static) reduction(&: ok)
        for (long b = 0; b < batch; ++b)
            ok &= ckfft_oracle_real_inverse(h, n, in + (size_t) b * istride, out + (size_t) b * n, tmp);
        free(tmp);
    }
    return ok;
}
