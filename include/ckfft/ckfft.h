/*
 * ckfft.h -- the classic ckfft C ABI, served by the B200 library (libckfft_b200.so).
 *
 * This header declares exactly the surface of the reference's public header
 * (reference: inc/ckfft/ckfft.h) so that a caller of Cricket FFT can relink against
 * libckfft_b200.so without source changes:
 *
 *   type / function        reference declaration        reference implementation
 *   CkFftComplex           inc/ckfft/ckfft.h:10-15      -
 *   CkFftContext           inc/ckfft/ckfft.h:18         src/ckfft/context.h:4-19
 *   CkFftDirection         inc/ckfft/ckfft.h:21-27      -
 *   CkFftInit              inc/ckfft/ckfft.h:59         src/ckfft/ckfft.cpp:14-34, context.cpp:24-114
 *   CkFftRealForward       inc/ckfft/ckfft.h:85         src/ckfft/ckfft.cpp:36-53
 *   CkFftRealInverse       inc/ckfft/ckfft.h:108        src/ckfft/ckfft.cpp:55-76
 *   CkFftComplexForward    inc/ckfft/ckfft.h:129        src/ckfft/ckfft.cpp:78-95
 *   CkFftComplexInverse    inc/ckfft/ckfft.h:150        src/ckfft/ckfft.cpp:97-114
 *   CkFftShutdown          inc/ckfft/ckfft.h:158        src/ckfft/ckfft.cpp:116-119
 *
 * Semantics kept from the reference: power-of-two sizes only; transforms are out of place and
 * un-normalised (inverse(forward(x)) = n*x for complex data, 2n*x for real data; the real forward
 * transform returns 2*rfft(x) in n/2+1 bins); every transform call returns 1 on success and 0 on
 * any invalid argument; CkFftInit returns NULL on failure; a context is immutable after creation
 * and may be shared between threads.
 *
 * What differs: the arithmetic runs on an NVIDIA B200 (sm_100a).  The twiddle tables are built
 * with the reference's formula and uploaded to the GPU that is current when CkFftInit is called.
 * Data pointers may be host pointers (staged through the library's own device buffers) or device
 * pointers on that GPU (used in place; 8-byte alignment required).  A CUDA failure, including
 * "no GPU", is reported the way the reference reports bad arguments: NULL from CkFftInit, 0 from
 * the transforms; CkFftB200LastError() in ckfft_b200.h gives the reason.  There is no CPU fallback.
 * n <= 0 is rejected (the reference's unsigned power-of-two test lets INT_MIN through).
 */
#ifndef CKFFT_CKFFT_H
#define CKFFT_CKFFT_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* one complex sample: two packed floats, layout-compatible with float2 / cuFloatComplex */
typedef struct
{
    float real;
    float imag;
} CkFftComplex;

typedef struct _CkFftContext CkFftContext;

/* bit flags: which transform directions a context can serve */
typedef enum
{
    kCkFftDirection_Forward = 1,
    kCkFftDirection_Inverse = 2,
    kCkFftDirection_Both = 3
} CkFftDirection;

/*
 * Create a context for transforms of up to nMax points (power of two).
 * buf/bufSize: optional caller-provided storage for the host part of the context.
 *   - both NULL: the library allocates;
 *   - bufSize given and (buf == NULL or *bufSize too small): the required byte count is written
 *     to *bufSize and NULL is returned (size query);
 *   - buf given without bufSize: NULL.
 * Device-side tables are always owned by the library and released by CkFftShutdown.
 */
CkFftContext* CkFftInit(int nMax, CkFftDirection direction, void* buf, size_t* bufSize);

/* real input[n] -> complex output[n/2+1], scaled by 2 */
int CkFftRealForward(CkFftContext* context, int n, const float* input, CkFftComplex* output);

/* complex input[n/2+1] -> real output[n]; tmpBuf (n/2+1 complex) must be non-NULL as in the
 * reference, but is not touched: the twist is fused into the transform kernel */
int CkFftRealInverse(CkFftContext* context, int n, const CkFftComplex* input, float* output, CkFftComplex* tmpBuf);

/* complex input[n] -> complex output[n], forward sign exp(-2*pi*i*j*k/n) */
int CkFftComplexForward(CkFftContext* context, int n, const CkFftComplex* input, CkFftComplex* output);

/* complex input[n] -> complex output[n], inverse sign, not divided by n */
int CkFftComplexInverse(CkFftContext* context, int n, const CkFftComplex* input, CkFftComplex* output);

/* release a context (NULL is allowed); frees host storage only if the library allocated it */
void CkFftShutdown(CkFftContext* context);

#ifdef __cplusplus
}
#endif

#endif /* CKFFT_CKFFT_H */
