// ref_driver.cpp -- thin ctypes-friendly driver linked INTO oracle/_ref/libckfft_ref.so
// together with the unmodified reference sources (compiled from /root/reference where
// they lie; see oracle/build.py).  TEST INFRASTRUCTURE ONLY: loaded by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
//
// It calls the reference's public C API (inc/ckfft/ckfft.h:59-158) and nothing else,
// one shared context per batch (legal: the context holds no state, ckfft.h:39-41),
// OpenMP `parallel for` over independent transforms.  It also exposes KISS FFT 1.3.0
// (ext/kiss_fft130), the regression oracle of the reference's own harness
// (src/test/test.cpp:315-395,754-786).
#include <stdlib.h>
#include <string.h>
#include <omp.h>

#include "ckfft/ckfft.h"
extern "C" {
#include "kiss_fft.h"
}

extern "C" {

int ckref_max_threads(void) { return omp_get_max_threads(); }

void* ckref_init(int nmax, int direction)
{
    return CkFftInit(nmax, (CkFftDirection) direction, NULL, NULL);
}

void ckref_shutdown(void* ctx) { CkFftShutdown((CkFftContext*) ctx); }

int ckref_complex(void* ctx, int n, const float* in, float* out, int inverse)
{
    CkFftContext* c = (CkFftContext*) ctx;
    return inverse ? CkFftComplexInverse(c, n, (const CkFftComplex*) in, (CkFftComplex*) out)
                   : CkFftComplexForward(c, n, (const CkFftComplex*) in, (CkFftComplex*) out);
}

int ckref_real_forward(void* ctx, int n, const float* in, float* out)
{
    return CkFftRealForward((CkFftContext*) ctx, n, in, (CkFftComplex*) out);
}

int ckref_real_inverse(void* ctx, int n, const float* in, float* out, float* tmp)
{
    return CkFftRealInverse((CkFftContext*) ctx, n, (const CkFftComplex*) in, out, (CkFftComplex*) tmp);
}

// threads <= 0: all OpenMP threads
int ckref_complex_batch(void* ctx, int n, const float* in, float* out, long batch, int inverse, int threads)
{
    int ok = 1;
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads) reduction(&: ok)
    for (long b = 0; b < batch; ++b)
        ok &= ckref_complex(ctx, n, in + (size_t) b * 2 * n, out + (size_t) b * 2 * n, inverse);
    return ok;
}

int ckref_real_forward_batch(void* ctx, int n, const float* in, float* out, long batch, int threads)
{
    const size_t ostride = 2 * ((size_t) n / 2 + 1);
    int ok = 1;
    if (threads <= 0) threads = omp_get_max_threads();
    if (n == 4) {
        // the reference writes output[3] for n == 4 (src/ckfft/fft_real.cpp:46-47): give it room
#pragma omp parallel for schedule(static) num_threads(threads) reduction(&: ok)
        for (long b = 0; b < batch; ++b) {
            float tmp[8];
            ok &= ckref_real_forward(ctx, n, in + (size_t) b * n, tmp);
            memcpy(out + (size_t) b * ostride, tmp, sizeof(float) * ostride);
        }
        return ok;
    }
#pragma omp parallel for schedule(static) num_threads(threads) reduction(&: ok)
    for (long b = 0; b < batch; ++b)
        ok &= ckref_real_forward(ctx, n, in + (size_t) b * n, out + (size_t) b * ostride);
    return ok;
}

int ckref_real_inverse_batch(void* ctx, int n, const float* in, float* out, long batch, int threads)
{
    const size_t istride = 2 * ((size_t) n / 2 + 1);
    int ok = 1;
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads)
    {
        float* tmp = (float*) malloc(sizeof(float) * istride);
#pragma omp for schedule(static) reduction(&: ok)
        for (long b = 0; b < batch; ++b)
            ok &= ckref_real_inverse(ctx, n, in + (size_t) b * istride, out + (size_t) b * n, tmp);
        free(tmp);
    }
    return ok;
}

// KISS FFT complex transform, as the reference harness drives it (src/test/test.cpp:315-360)
int ckref_kiss_complex(int n, const float* in, float* out, int inverse)
{
    kiss_fft_cfg cfg = kiss_fft_alloc(n, inverse, NULL, NULL);
    if (!cfg) return 0;
    kiss_fft(cfg, (const kiss_fft_cpx*) in, (kiss_fft_cpx*) out);
    free(cfg);
    return 1;
}

}  // extern "C"
