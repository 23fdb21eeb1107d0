/* abi_smoke.c -- a plain C99 caller of the drop-in library, written the way the reference's example
 * (src/example/main.cpp:44-84) and manual use the API: size query + caller-owned context buffer, real forward,
 * real inverse with a tmp buffer, round-trip error; then the classic complex calls and one batched call.
 * Built and run by tests/test_c_abi_gpu.py with gcc against include/ckfft and libckfft_b200.so. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "ckfft/ckfft.h"
#include "ckfft/ckfft_b200.h"

#define N 1024
#define BATCH 33

static float frand(unsigned* s) { *s = *s * 1664525u + 1013904223u; return (float) ((*s >> 8) & 0xFFFF) / 32768.0f - 1.0f; }

int main(void)
{
    unsigned seed = 12345u;
    size_t need = 0;
    void* mem;
    CkFftContext* ctx;
    float* x = (float*) malloc(sizeof(float) * N);
    float* back = (float*) malloc(sizeof(float) * N);
    CkFftComplex* spec = (CkFftComplex*) malloc(sizeof(CkFftComplex) * (N / 2 + 1));
    CkFftComplex* tmp = (CkFftComplex*) malloc(sizeof(CkFftComplex) * (N / 2 + 1));
    CkFftComplex* cin = (CkFftComplex*) malloc(sizeof(CkFftComplex) * N * BATCH);
    CkFftComplex* cout = (CkFftComplex*) malloc(sizeof(CkFftComplex) * N * BATCH);
    CkFftComplex* cback = (CkFftComplex*) malloc(sizeof(CkFftComplex) * N * BATCH);
    double err = 0.0, cerr = 0.0, cnorm = 0.0;
    int i;

    /* the usage of inc/ckfft/ckfft.h:51-55 */
    if (CkFftInit(N, kCkFftDirection_Both, NULL, &need) != NULL || need == 0) { printf("size query failed\n"); return 2; }
    mem = malloc(need);
    ctx = CkFftInit(N, kCkFftDirection_Both, mem, &need);
    if (!ctx) { printf("CkFftInit failed: %s\n", CkFftB200LastError()); return 3; }

    for (i = 0; i < N; ++i) x[i] = frand(&seed);
    if (!CkFftRealForward(ctx, N, x, spec)) { printf("real forward: %s\n", CkFftB200LastError()); return 4; }
    if (!CkFftRealInverse(ctx, N, spec, back, tmp)) { printf("real inverse: %s\n", CkFftB200LastError()); return 5; }
    for (i = 0; i < N; ++i) { double d = back[i] / (2.0 * N) - x[i]; err += d * d; }
    printf("error: %f\n", err);                       /* the reference example prints "error: 0.000000" */
    if (err > 1e-8) return 6;

    for (i = 0; i < N * BATCH; ++i) { cin[i].real = frand(&seed); cin[i].imag = frand(&seed); }
    if (!CkFftComplexForward(ctx, N, cin, cout)) return 7;                       /* classic call, first transform */
    if (!CkFftComplexForwardBatch(ctx, N, cin, cout, BATCH)) return 8;           /* batched extension */
    if (!CkFftComplexInverseBatch(ctx, N, cout, cback, BATCH)) return 9;
    for (i = 0; i < N * BATCH; ++i) {
        double dr = cback[i].real / N - cin[i].real, di = cback[i].imag / N - cin[i].imag;
        cerr += dr * dr + di * di;
        cnorm += (double) cin[i].real * cin[i].real + (double) cin[i].imag * cin[i].imag;
    }
    printf("complex round trip relative rms: %.3e\n", sqrt(cerr / cnorm));
    if (sqrt(cerr / cnorm) > 1e-5) return 10;

    /* error returns as in src/ckfft/ckfft.cpp */
    if (CkFftComplexForward(ctx, N, cin, cin)) return 11;      /* in == out */
    if (CkFftComplexForward(ctx, 2 * N, cin, cout)) return 12; /* n > nMax */
    if (CkFftRealInverse(ctx, N, spec, back, NULL)) return 13; /* tmpBuf NULL */

    CkFftShutdown(ctx);   /* does not free `mem`: the caller owns it */
    free(mem);
    free(x); free(back); free(spec); free(tmp); free(cin); free(cout); free(cback);
    printf("ok\n");
    return 0;
}
