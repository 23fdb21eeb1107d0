/*
 * ckfft_b200.h -- B200 extensions to the ckfft C ABI: a batched variant of each transform call,
 * stream-ordered variants for device-resident data, and a few introspection helpers.
 *
 * Each batched call is "the classic call, `batch` times": transform b reads
 * input + b*in_stride and writes output + b*out_stride, with the argument checks, scaling and
 * return convention of the classic call it extends (reference: src/ckfft/ckfft.cpp:36-114).
 * Plain C, no CUDA or torch types: streams travel as void* (a cudaStream_t).
 */
#ifndef CKFFT_CKFFT_B200_H
#define CKFFT_CKFFT_B200_H

#include "ckfft.h"

#ifdef __cplusplus
extern "C" {
#endif

/*
 * Synchronous batched transforms over densely packed arrays.  Pointers may be host or device
 * memory (both arrays of one call on the same side).  Host arrays are streamed through the GPU in
 * chunks with copies and kernels overlapped; pinned host memory (CkFftB200HostAlloc) makes those
 * copies asynchronous.  On return the results are complete.
 *   complex:       input/output stride n complex
 *   real forward:  input stride n floats, output stride n/2+1 complex   (extends CkFftRealForward,  ckfft.h / ref :85)
 *   real inverse:  input stride n/2+1 complex, output stride n floats   (extends CkFftRealInverse,  ref :108);
 *                  tmpBuf is ignored and may be NULL
 */
int CkFftComplexForwardBatch(CkFftContext* context, int n, const CkFftComplex* input, CkFftComplex* output, size_t batch);
int CkFftComplexInverseBatch(CkFftContext* context, int n, const CkFftComplex* input, CkFftComplex* output, size_t batch);
int CkFftRealForwardBatch(CkFftContext* context, int n, const float* input, CkFftComplex* output, size_t batch);
int CkFftRealInverseBatch(CkFftContext* context, int n, const CkFftComplex* input, float* output, CkFftComplex* tmpBuf, size_t batch);

/*
 * Stream-ordered batched transforms on DEVICE memory of the context's GPU.  The call enqueues the
 * work on `stream` (NULL = the legacy default stream) and returns without synchronising.
 * Strides are in elements of the respective array (complex elements or floats); 0 selects the
 * dense default.  Real-array strides must be even and all pointers 8-byte aligned.
 * Returns 1 if the work was enqueued, 0 on invalid arguments or a CUDA error.
 */
int CkFftComplexForwardBatchAsync(CkFftContext* context, int n, const CkFftComplex* input, CkFftComplex* output,
                                  size_t batch, size_t inStride, size_t outStride, void* stream);
int CkFftComplexInverseBatchAsync(CkFftContext* context, int n, const CkFftComplex* input, CkFftComplex* output,
                                  size_t batch, size_t inStride, size_t outStride, void* stream);
int CkFftRealForwardBatchAsync(CkFftContext* context, int n, const float* input, CkFftComplex* output,
                               size_t batch, size_t inStride, size_t outStride, void* stream);
int CkFftRealInverseBatchAsync(CkFftContext* context, int n, const CkFftComplex* input, float* output,
                               size_t batch, size_t inStride, size_t outStride, void* stream);

/*
 * Audio front end: what a caller of CkFftRealForward on audio frames does next, fused into the transform kernel.
 *   power[b][k] = | CkFftRealForward(window .* input[b]) [k] |^2 ,  k = 0 .. n/2      (= 4 |rfft(w x)|^2)
 * `window` (n floats, device memory) may be NULL for a rectangular window.  The window is applied while the frame
 * is loaded and only the n/2+1 powers are written (floats), so a frame costs 4n + 4(n/2+1) bytes of HBM traffic
 * instead of the 3 x that of separate window / transform / magnitude passes.  32 <= n <= 32768, device pointers,
 * stream-ordered; strides in floats (0 = dense), the input stride must be even.
 */
int CkFftB200RealForwardPowerBatchAsync(CkFftContext* context, int n, const float* input, const float* window, float* power,
                                        size_t batch, size_t inStride, size_t outStride, void* stream);

/* How the library will run a transform of n points (host-side planner, no GPU needed). */
typedef struct
{
    int n;                    /* transform length as passed */
    int isReal;               /* 0 complex, 1 real */
    int complexPoints;        /* length of the complex transform actually computed (n or n/2) */
    int passes;               /* kernel launches (= HBM round trips) per batch: 1 single pass, 2 four-step */
    int radix[2][3];          /* radices of each pass, 0-terminated */
    int threadsPerTransform;  /* threads cooperating on one transform (first pass) */
    int elemsPerThread;       /* complex values held in registers per thread */
    int transformsPerCta;
    int sharedBytes;          /* dynamic shared memory per CTA */
} CkFftB200Plan;

/* returns 1 and fills *plan, or 0 if n is not a supported power of two */
int CkFftB200GetPlan(int n, int isReal, CkFftB200Plan* plan);

/* text of the last failure on the calling thread ("" if none) */
const char* CkFftB200LastError(void);

/* number of kernels this library has launched in this process (all threads) */
unsigned long long CkFftB200KernelLaunches(void);

/* pinned host memory for fast host<->device streaming; NULL on failure */
void* CkFftB200HostAlloc(size_t bytes);
void CkFftB200HostFree(void* p);

/* device the context is bound to, or -1 */
int CkFftB200ContextDevice(const CkFftContext* context);

/*
 * Local steps of the distributed six-step transform of ONE very large 1-D FFT (n = n1*n2 spread over P GPUs,
 * ckfft_b200/distributed.py).  The exchange between GPUs (all-to-all) is the caller's, these are the
 * stream-ordered device kernels around it.  Device pointers, out of place unless noted.
 *   PackColumns:      in[rows][parts][width]          -> out[parts][rows][width]      (slab q goes to peer q)
 *   UnpackTranspose:  in[parts][rowsPerPart][width]   -> out[width][parts*rowsPerPart]
 *   TwiddleRows:      data[i][k] *= exp(-+2*pi*i*(firstRow+i)*k/n) in place, i < rows, k < cols;
 *                     needs a context created with nMax >= n > 16384 and (firstRow+rows)*cols <= n
 */
int CkFftB200PackColumnsAsync(const CkFftComplex* in, CkFftComplex* out, size_t rows, int parts, size_t width, void* stream);
int CkFftB200UnpackTransposeAsync(const CkFftComplex* in, CkFftComplex* out, int parts, size_t rowsPerPart, size_t width,
                                  void* stream);
int CkFftB200TwiddleRowsAsync(CkFftContext* context, int n, CkFftComplex* data, size_t rows, size_t cols, size_t firstRow,
                              int inverse, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* CKFFT_CKFFT_B200_H */
