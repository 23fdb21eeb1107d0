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
 *
 * In place: unlike the classic calls (which keep the reference's `input == output -> 0`, src/ckfft/ckfft.cpp:46,69,
 * 88,107), these accept input == output when every transform reads and writes the same bytes: complex transforms
 * with inStride == outStride; real transforms with the real-array stride equal to twice the spectrum stride (rows
 * padded to n + 2 floats).  Real transforms longer than 32768 points and partially overlapping arrays are not
 * supported in place.
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
 * Split-complex ("planar") arrays: real and imaginary parts in two float arrays, as vDSP-style callers hold them
 * (the layout of the reference harness's Accelerate competitor, DSPSplitComplex, src/test/test.cpp:398-461).  Same
 * transform, scaling and checks as CkFftComplexForward / CkFftComplexInverse; n <= 16384, device pointers (4-byte
 * aligned), stream-ordered; strides in floats (0 = n).  In place (outRe == inRe and outIm == inIm, equal strides)
 * is accepted; any other aliasing between the four arrays returns 0.
 */
int CkFftB200ComplexForwardPlanarBatchAsync(CkFftContext* context, int n, const float* inRe, const float* inIm,
                                            float* outRe, float* outIm, size_t batch, size_t inStride, size_t outStride,
                                            void* stream);
int CkFftB200ComplexInversePlanarBatchAsync(CkFftContext* context, int n, const float* inRe, const float* inIm,
                                            float* outRe, float* outIm, size_t batch, size_t inStride, size_t outStride,
                                            void* stream);

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
    int transformsPerCta;     /* of the plain-load plan (any alignment / stride); the bulk-prefetch kernels that take 16-byte aligned */
    int sharedBytes;          /* rows run the same radices with their own grouping and staging buffers (csrc/plans.h) */
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
 * Batched scheduler: one host call spreads a batch of independent transforms over several GPUs of the box.
 * The transforms of a batch share no state -- the reference says so itself ("the context does not contain state,
 * so contexts can be used simultaneously on different threads", inc/ckfft/ckfft.h:39-41) -- so the batch is cut into
 * shards, every device runs the ordinary host-buffer pipeline of the CkFft*Batch calls (chunked H2D -> kernel -> D2H,
 * several chunks in flight) on its own streams from its own host thread, and the call returns when all of them are
 * done.  There is no collective and no traffic between the GPUs.  Small batches are cut into contiguous shards, one
 * per device (CkFftB200ShardRange); batches of 256 MiB or more are scheduled dynamically -- every device draws 32 MiB
 * chunks of the batch from a shared counter, so that devices behind a slower host link take fewer and all finish
 * together (CKFFT_B200_MULTI_STATIC=1 forces static shards).  Results do not depend on the schedule.
 *
 *   CkFftB200MultiInit(nMax, direction, devices, nDevices)
 *       one context replica (CkFftInit(nMax, direction) on that device) and one worker thread per entry of
 *       devices[0 .. nDevices).  devices == NULL: the first nDevices visible devices (nDevices <= 0: all of them).
 *       NULL on invalid arguments (the checks of CkFftInit, src/ckfft/ckfft.cpp:16-31), if a device does not exist or
 *       if a replica cannot be created; CkFftB200LastError() says why.
 *   CkFft{ComplexForward,ComplexInverse,RealForward,RealInverse}BatchMulti
 *       the CkFft*Batch call of the same name on HOST arrays (pageable or pinned; dense strides), same checks and
 *       1 / 0 return.  Pinned arrays (CkFftB200HostAlloc / cudaHostRegister) let the copies of all devices run
 *       asynchronously; pageable arrays work but are staged by the driver (CKFFT_B200_PIN=1 page-locks them for the
 *       duration of calls that move >= 64 MiB -- measured to cost as much as it saves).  One call at a time per handle
 *       (concurrent callers are serialised).
 *   CkFftB200MultiContext(m, i)  the replica on device i (for device-resident data: use it with the *BatchAsync calls)
 *   CkFftB200ShardRange(batch, part, parts, &first, &count)
 *       the shard of `part`: [first, first + count); the first batch % parts shards are one transform longer.
 *       Pure host arithmetic (no GPU needed); returns 0 on bad arguments.
 */
typedef struct CkFftB200Multi CkFftB200Multi;
CkFftB200Multi* CkFftB200MultiInit(int nMax, CkFftDirection direction, const int* devices, int nDevices);
void CkFftB200MultiShutdown(CkFftB200Multi* multi);
int CkFftB200MultiDeviceCount(const CkFftB200Multi* multi);
int CkFftB200MultiDevice(const CkFftB200Multi* multi, int index);           /* CUDA device ordinal of replica `index`, or -1 */
CkFftContext* CkFftB200MultiContext(const CkFftB200Multi* multi, int index);
int CkFftB200ShardRange(size_t batch, int part, int parts, size_t* first, size_t* count);
int CkFftComplexForwardBatchMulti(CkFftB200Multi* multi, int n, const CkFftComplex* input, CkFftComplex* output, size_t batch);
int CkFftComplexInverseBatchMulti(CkFftB200Multi* multi, int n, const CkFftComplex* input, CkFftComplex* output, size_t batch);
int CkFftRealForwardBatchMulti(CkFftB200Multi* multi, int n, const float* input, CkFftComplex* output, size_t batch);
int CkFftRealInverseBatchMulti(CkFftB200Multi* multi, int n, const CkFftComplex* input, float* output, size_t batch);

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


/*
 * Fused distributed transform of ONE very large 1-D complex FFT over the P <= 8 GPUs of one NVSwitch domain
 * (n = n1*n2 = 2^14 .. 2^30, one process per GPU, rank r owns the natural-order slice [r*n/P, (r+1)*n/P) of the
 * input and of the output).  No collective library on the data path: the FFT passes themselves store their results
 * into the memory of the GPU that needs them next (peer stores over NVLink, 128-byte row chunks), so each of
 * the two inner all-to-all transposes of the six-step algorithm is fused into the transform pass that feeds it:
 *
 *   exchange  x[n1][n2], rows on rank(n1)            -> work[n1][n2 in my column block]          (push copy)
 *   pass A    (n1 = a*lb + b, only if n1 > 1024)  la-point FFTs over a, * W_n1^(b*ka)             local
 *   pass B    lb-point FFTs over b, * W_n^(n2*k1), row k1 stored on rank(k1): mid[k1][n2]         NVLink stores
 *   pass C    (n2 = c*ld + d, only if n2 > 1024)  lc-point FFTs over c, * W_n2^(d*kc)             local
 *   pass D    ld-point FFTs over d, X[k1 + n1*k2] stored on rank(k2): out[k2][k1]                 NVLink stores
 *
 * with a flag barrier over peer memory after the exchange, after pass B and after pass D.  The reference has no
 * multi-device path; the index algebra is the six-step of ext/fftw-3.3.2/mpi/dft-rank1.c:58-79.
 *
 * Buffers: every rank allocates three arrays of n/P complex values (work, mid, out) and a 256-byte flag block with
 * CkFftB200PeerAlloc, exports them (CkFftB200PeerExport, 64-byte handles the caller ships to the other processes by
 * any means) and opens the peers' (CkFftB200PeerOpen).  The result of an execution is the rank's own `out` array; it
 * stays valid until the next execution on the plan.  All ranks must execute the same sequence of calls.
 */
typedef struct
{
    int log2n, world;
    int log2n1, log2n2;       /* n = n1 * n2 */
    int la, lb;               /* n1 = la * lb (la = 1: n1 is one pass) */
    int lc, ld;               /* n2 = lc * ld (lc = 1: n2 is one pass) */
    int passes;               /* FFT passes over the data (2 .. 4), not counting the exchange */
    int pull;                 /* 1: no exchange kernel -- the first pass fetches its tiles from the peers' input arrays (TMA) */
} CkFftB200DistLayout;

/* One FFT pass as the tile kernel sees it (the CPU tests replay these descriptors in numpy). */
typedef struct
{
    int kind;                 /* 0: column pass over a [L][ncols] array per problem, 1: last pass (contiguous columns) */
    int routed;               /* 1: output rows are distributed over the ranks (see below), 0: local geometry */
    int L;                    /* points per transform of this pass */
    long long nproblems;
    int ncols;
    int twLog2;               /* kind 0: outputs are multiplied by W_(2^twLog2)^(cc * kt) */
    int twColBase, twColShift;/*         cc = (twColBase + c) >> twColShift;  kt = k (local) or kk (routed) */
    int kProbMul, kMul;       /* routed: bin k of problem q is row kk = q*kProbMul + k*kMul ...                */
    int rankShift;            /*         ... of rank kk >> rankShift, local row kk & (2^rankShift - 1),        */
    long long outRowStride;   /*         stored at row*outRowStride + outColBase + c of that rank's dst array  */
    long long outColBase;
    long long inColStride;    /* routed kind 1: column c of problem q starts at c*inColStride + q*inProbStride */
    long long inProbStride;
    int src, dst;             /* 0 work, 1 mid, 2 out */
    int pull;                 /* 1 (first pass of a pull layout): row a of the [L][ncols] problem is row a % pullRows of   */
    int pullRows;             /*   rank a / pullRows's INPUT array viewed as [pullRows][pullRowLen], and column c is its    */
    long long pullRowLen;     /*   element (c / pullW) * pullN2 + pullCol0 + c % pullW                                       */
    long long pullN2, pullCol0;
    int pullW;
} CkFftB200DistPass;

/* Pure host arithmetic (no GPU needed).  preferPasses: 0 = default, 3 or 4 forces that pass count where possible.
 * Returns 1, or 0 if n is not a power of two in 2^14 .. 2^30 or cannot be spread over `world` ranks. */
int CkFftB200DistGetLayout(long long n, int world, int preferPasses, CkFftB200DistLayout* layout);
/* Fills passes[0 .. return value) for `rank`; at most 4. */
int CkFftB200DistDescribe(const CkFftB200DistLayout* layout, int rank, CkFftB200DistPass passes[4]);

/* Exportable device memory on the current device (zero-filled).  */
void* CkFftB200PeerAlloc(size_t bytes);
void CkFftB200PeerFree(void* p);
int CkFftB200PeerExport(void* p, unsigned char handle[64]);
void* CkFftB200PeerOpen(const unsigned char handle[64]);     /* in another process; maps the peer's allocation */
void CkFftB200PeerClose(void* p);

typedef struct CkFftB200DistPlan CkFftB200DistPlan;
/* work/mid/out/flags: arrays of `world` device pointers, entry q = rank q's buffer (own allocation at q == rank).
 * `in` (may be NULL): a fourth array of n/world complex per rank.  When given, the plan runs in PULL mode: there is no
 * exchange kernel, the first FFT pass fetches its tiles straight from the peers' `in` arrays with TMA; an execution
 * whose input is not the rank's own `in` array copies it there first.
 * The context must have been created with nMax >= n (it carries the twiddles of W_n).
 * Flag blocks may be reused by a later plan on the same ranks (the new plan continues from the epoch the block holds)
 * provided every execution of the earlier plan has completed on every rank; all ranks must rendezvous (application
 * barrier) between creating their plans and the first execution. */
CkFftB200DistPlan* CkFftB200DistPlanCreate(CkFftContext* context, long long n, int rank, int world, int preferPasses,
                                           void* const* work, void* const* mid, void* const* out, void* const* flags,
                                           void* const* in);
/* Enqueue one transform of this rank's slice `input` (n/world complex, device memory, not one of the plan's
 * buffers) on `stream`.  inverse != 0: un-normalised inverse.  Returns 1 if everything was enqueued. */
int CkFftB200DistExecAsync(CkFftB200DistPlan* plan, const CkFftComplex* input, int inverse, void* stream);
/* Synchronises the device and returns 1 if no barrier of this plan has timed out so far, else 0.
 * A barrier gives up after 20 s (CKFFT_B200_DIST_TIMEOUT_MS overrides) instead of hanging the GPU; the execution then
 * carries on with incomplete data, ExecAsync has long returned 1, and the rank's `out` array is overwritten with NaNs
 * by the execution's last barrier.  Call this after the executions whose results matter. */
int CkFftB200DistPlanStatus(CkFftB200DistPlan* plan);
/* Profiling: when on, every execution records an event after each of its kernels.  CkFftB200DistPlanPhases
 * synchronises and returns the number of phases of the LAST execution, their device times in ms[] (room for 12)
 * and their names, comma separated, in names[]. */
int CkFftB200DistPlanSetProfiling(CkFftB200DistPlan* plan, int on);
int CkFftB200DistPlanPhases(CkFftB200DistPlan* plan, float* ms, char* names, size_t namesBytes);
void CkFftB200DistPlanDestroy(CkFftB200DistPlan* plan);

#ifdef __cplusplus
}
#endif

#endif /* CKFFT_CKFFT_B200_H */
