// api.cu -- the C ABI of libckfft_b200.so: ckfft's six classic entry points plus the batched /
// stream-ordered variants declared in include/ckfft/ckfft_b200.h.
//
// Host-side mirror of the reference's API layer and context:
//   argument validation      src/ckfft/ckfft.cpp:14-119   (same checks, same 1/0/NULL returns)
//   context + twiddle tables src/ckfft/context.cpp:24-122 (same size-query / user-buffer protocol,
//                                                          same fp32 table formula)
//   size dispatch            src/ckfft/fft.cpp:13-46, src/ckfft/fft_real.cpp:13-105
// The arithmetic itself is in the CUDA kernels (fft_kernel.cuh, tiny_kernel.cuh, four_step.cu).
// There is no CPU path: without a usable GPU CkFftInit returns NULL and says why through
// CkFftB200LastError().
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "ckfft/ckfft.h"
#include "ckfft/ckfft_b200.h"
#include "launch.h"
#include "plans.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace ckb {                                             // host_copy.cpp
int stream_copy_level();
void stream_copy(void* dst, const void* src, size_t bytes, int level);
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
// The first five members mirror the reference's struct (src/ckfft/context.h:4-19) so that code
// which reaches behind the ABI the way the reference's own harness does (src/test/test.cpp:245-261
// writes context->neon) keeps working.  `neon` is always false here.
struct _CkFftContext
{
    bool neon;
    int maxCount;
    const CkFftComplex* fwdExpTable;   // host copy, NULL unless the Forward bit was requested
    const CkFftComplex* invExpTable;   // host copy, NULL unless the Inverse bit was requested
    bool ownBuf;

    // --- B200 part ---
    uint32_t magic;
    int device;
    int tableCount;          // Nt = min(maxCount, CKB_MAX_TABLE): entries in the device table
    int log2Table;
    float2* dTable;          // device: W_Nt^k = (cos, sin)(-2 pi k / Nt), forward sign
    // multi-pass transforms (nMax > CKB_MAX_SINGLE_PASS): two-level twiddles of W_Tmax, Tmax = nMax
    float2* dTwLo;           // W_Tmax^j,        j < 2^twH
    float2* dTwHi;           // W_Tmax^(i<<twH), i < Tmax >> twH
    int twH;
    int log2Tmax;
};

namespace {

constexpr uint32_t kMagic = 0x434b4232u;   // "CKB2"

thread_local char tl_error[256] = "";

void set_error(const char* what, cudaError_t e = cudaSuccess)
{
    if (e != cudaSuccess) snprintf(tl_error, sizeof(tl_error), "%s: %s", what, cudaGetErrorString(e));
    else                  snprintf(tl_error, sizeof(tl_error), "%s", what);
}

std::atomic<unsigned long long> g_launches{0};

inline bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }
inline int ilog2i(long long x) { int l = 0; while ((1LL << l) < x) ++l; return l; }

// RAII: make the context's device current for the duration of a call
struct DeviceGuard
{
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
        else if (prev == dev) prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

size_t context_bytes() { return (sizeof(_CkFftContext) + 7) & ~size_t(7); }

enum Kind { K_C2C_FWD, K_C2C_INV, K_R2C, K_C2R };

inline size_t in_elems(Kind k, int n)  { return k == K_C2R ? (size_t) n / 2 + 1 : (size_t) n; }
inline size_t out_elems(Kind k, int n) { return k == K_R2C ? (size_t) n / 2 + 1 : (size_t) n; }
inline size_t in_elem_bytes(Kind k)  { return k == K_R2C ? 4 : 8; }
inline size_t out_elem_bytes(Kind k) { return k == K_C2R ? 4 : 8; }

cudaError_t enqueue(const _CkFftContext* c, Kind kind, int n, const void* in, void* out, long long batch,
                    long long in_stride, long long out_stride, cudaStream_t s);

// ---- multi-pass lengths ---------------------------------------------------------------------
// Scratch is stream-ordered (ckb::scratch_alloc = the library's own memory pool / cudaFreeAsync): nothing mutable lives in the context.
constexpr size_t kScratchCapBytes = size_t(2) << 30;      // real-transform glue buffers

// Group size of the multi-pass transforms = size of the stream-ordered scratch array.  Measured on B200: groups
// small enough for the intermediate to stay in the 126 MB L2 (16-64 MiB) LOSE (.33 -> .22-.30 of the copy peak at
// 2^15..2^20): the per-group launches are too short to fill the machine.  Default 2 GiB.
// (CKFFT_B200_L2_GROUP_MB overrides it for measurements.)
size_t l2_group_bytes()
{
    static const size_t bytes = [] {
        const char* e = getenv("CKFFT_B200_L2_GROUP_MB");
        const long mb = e ? atol(e) : 0;
        return size_t(mb > 0 && mb <= 8192 ? mb : 2048) << 20;
    }();
    return bytes;
}

bool getenv_flag(const char* name, int dflt)
{
    const char* e = getenv(name);
    return (e && *e ? atoi(e) : dflt) != 0;
}

ckb::BigTwiddles big_tw(const _CkFftContext* c) { return ckb::BigTwiddles{ c->dTwLo, c->dTwHi, c->twH, c->log2Tmax }; }

cudaError_t enqueue_large_c2c(const _CkFftContext* c, bool inv, int n, const ckb::cf* in, ckb::cf* out, long long batch,
                              long long in_stride, long long out_stride, cudaStream_t s)
{
    if (!c->dTwLo) return cudaErrorNotSupported;
    if (in_stride != n || out_stride != n) {
        // Padded rows.  The multi-pass kernels describe a batch as ONE dense [batch * L0][L1] array (tensor
        // maps, ring slots), and a transform of >= 2^15 points fills the machine on its own: strided batches run one
        // transform per launch sequence, each dense in itself.
        cudaError_t e = cudaSuccess;
        for (long long i = 0; i < batch && e == cudaSuccess; ++i)
            e = enqueue_large_c2c(c, inv, n, in + i * in_stride, out + i * out_stride, 1, n, n, s);
        return e;
    }
    if (n <= (1 << 20) && ckb::pipe_enabled() && (((uintptr_t) in) & 15) == 0) {
        // two passes as one persistent kernel, intermediate ring resident in L2 (pipe_kernel.cuh)
        const long long chunk = 1LL << 20;                 // problems per launch (keeps the ticket in 32 bits)
        cudaError_t e = cudaSuccess;
        for (long long done = 0; done < batch && e == cudaSuccess; done += chunk) {
            const long long cnt = batch - done < chunk ? batch - done : chunk;
            e = ckb::launch_pipe(inv, ilog2i(n), in + done * n, out + done * n, cnt, c->dTable, c->log2Table, big_tw(c), s);
        }
        return e;
    }
    const size_t per = (size_t) n * sizeof(ckb::cf);
    long long sub = (long long) (l2_group_bytes() / per);
    if (sub < 1) sub = 1;
    if (sub > batch) sub = batch;
    ckb::cf* scratch = nullptr;
    cudaError_t e = ckb::scratch_alloc((void**) &scratch, per * (size_t) sub, s);
    if (e != cudaSuccess) return e;
    for (long long done = 0; done < batch && e == cudaSuccess; done += sub) {
        const long long cnt = batch - done < sub ? batch - done : sub;
        e = ckb::launch_four_step(inv, ilog2i(n), in + done * n, out + done * n, scratch, cnt, c->dTable, c->log2Table,
                                  big_tw(c), s);
    }
    cudaError_t e2 = cudaFreeAsync(scratch, s);
    return e != cudaSuccess ? e : e2;
}

// real n > CKB_MAX_TABLE: half-length complex transform (single- or multi-pass) + element-wise split / twist pass
cudaError_t enqueue_large_real(const _CkFftContext* c, bool inverse, int n, const void* in, void* out, long long batch,
                               long long in_stride, long long out_stride, cudaStream_t s)
{
    using ckb::cf;
    if (!c->dTwLo) return cudaErrorNotSupported;
    const int M = n / 2;
    if ((!inverse && (in_stride != n || out_stride != M + 1)) || (inverse && (in_stride != M + 1 || out_stride != n))) {
        // padded rows: one frame per launch sequence (see enqueue_large_c2c); strides are in elements of each array
        cudaError_t e = cudaSuccess;
        for (long long i = 0; i < batch && e == cudaSuccess; ++i) {
            const void* fi = inverse ? (const void*) ((const cf*) in + i * in_stride) : (const void*) ((const float*) in + i * in_stride);
            void* fo = inverse ? (void*) ((float*) out + i * out_stride) : (void*) ((cf*) out + i * out_stride);
            e = enqueue_large_real(c, inverse, n, fi, fo, 1, inverse ? M + 1 : n, inverse ? n : M + 1, s);
        }
        return e;
    }
    if (!inverse && M <= (1 << 20) && ckb::pipe_enabled() && getenv_flag("CKFFT_B200_PIPE_REAL", 1) && (((uintptr_t) in) & 15) == 0) {
        // real forward: half-length complex transform + split in ONE dataflow kernel (pipe_kernel.cuh, PipeCfg::REAL)
        const long long chunk = 1LL << 20;
        cudaError_t e = cudaSuccess;
        for (long long done = 0; done < batch && e == cudaSuccess; done += chunk) {
            const long long cnt = batch - done < chunk ? batch - done : chunk;
            e = ckb::launch_pipe_r2c(ilog2i(M), (const cf*) ((const float*) in + done * n), (cf*) out + done * (M + 1), cnt, M + 1,
                                     c->dTable, c->log2Table, big_tw(c), s);
        }
        return e;
    }
    if (inverse && M <= (1 << 20) && ckb::pipe_enabled() && getenv_flag("CKFFT_B200_PIPE_REAL", 1)) {
        // real inverse: twist + half-length complex transform in ONE dataflow kernel (pipe_kernel.cuh, PipeCfg::TWIST)
        const long long chunk = 1LL << 20;
        cudaError_t e = cudaSuccess;
        for (long long done = 0; done < batch && e == cudaSuccess; done += chunk) {
            const long long cnt = batch - done < chunk ? batch - done : chunk;
            e = ckb::launch_pipe_c2r(ilog2i(M), (const cf*) in + done * (M + 1), (cf*) ((float*) out + done * n), cnt, M + 1,
                                     c->dTable, c->log2Table, big_tw(c), s);
            if (e == cudaErrorNotSupported && done == 0) break;      // no tensor map for this array: separate twist pass below
        }
        if (e != cudaErrorNotSupported) return e;
        cudaGetLastError();
    }
    const size_t per = (size_t) M * sizeof(cf);
    long long sub = (long long) (kScratchCapBytes / per);
    if (sub < 1) sub = 1;
    if (sub > batch) sub = batch;
    cf* half = nullptr;      // Z (forward) or T (inverse): M complex per frame
    cudaError_t e = ckb::scratch_alloc((void**) &half, per * (size_t) sub, s);
    if (e != cudaSuccess) return e;
    for (long long done = 0; done < batch && e == cudaSuccess; done += sub) {
        const long long cnt = batch - done < sub ? batch - done : sub;
        if (!inverse) {
            const cf* x = (const cf*) ((const float*) in + done * n);      // n floats = M complex per frame
            cf* y = (cf*) out + done * (M + 1);
            e = enqueue(c, K_C2C_FWD, M, x, half, cnt, M, M, s);
            if (e == cudaSuccess) e = ckb::launch_real_split(half, y, n, cnt, M, M + 1, big_tw(c), s);
        } else {
            const cf* y = (const cf*) in + done * (M + 1);
            cf* x = (cf*) ((float*) out + done * n);
            e = ckb::launch_real_twist(y, half, n, cnt, M + 1, M, big_tw(c), s);
            if (e == cudaSuccess) e = enqueue(c, K_C2C_INV, M, half, x, cnt, M, M, s);
        }
    }
    cudaError_t e2 = cudaFreeAsync(half, s);
    return e != cudaSuccess ? e : e2;
}

// Enqueue `batch` transforms on DEVICE memory.  Strides are in elements of the respective array.
// Returns cudaSuccess or the launch error.
cudaError_t enqueue(const _CkFftContext* c, Kind kind, int n, const void* in, void* out, long long batch,
                    long long in_stride, long long out_stride, cudaStream_t s)
{
    using namespace ckb;
    if (batch <= 0) return cudaSuccess;
    const cf* table = c->dTable;
    if (kind == K_C2C_FWD || kind == K_C2C_INV) {
        KernelParams p{ (const cf*) in, (cf*) out, table, c->log2Table, batch, in_stride, out_stride };
        const bool inv = kind == K_C2C_INV;
        if (n >= 8 && n <= 64 && small_enabled()) return launch_small_c2c(n, inv, p, s);
        if (n <= 8) return launch_tiny_c2c(n, inv, p, s);
        if (n <= CKB_MAX_SINGLE_PASS) return inv ? launch_c2c_inv(n, p, s) : launch_c2c_fwd(n, p, s);
        return enqueue_large_c2c(c, inv, n, (const cf*) in, (cf*) out, batch, in_stride, out_stride, s);
    }
    if (n >= 16 && n <= 64 && small_enabled() && ((kind == K_R2C ? in_stride : out_stride) & 1) == 0) {
        // rows of the real array on 8-byte boundaries: both arrays addressed in 8-byte units
        if (kind == K_R2C) {
            KernelParams p{ (const cf*) in, (cf*) out, table, c->log2Table, batch, in_stride / 2, out_stride };
            return launch_small_r2c(n / 2, p, s);
        }
        KernelParams p{ (const cf*) in, (cf*) out, table, c->log2Table, batch, in_stride, out_stride / 2 };
        return launch_small_c2r(n / 2, p, s);
    }
    if (n <= 16) {
        return kind == K_R2C
            ? launch_tiny_r2c(n, (const float*) in, (cf*) out, table, c->log2Table, batch, in_stride, out_stride, s)
            : launch_tiny_c2r(n, (const cf*) in, (float*) out, table, c->log2Table, batch, in_stride, out_stride, s);
    }
    const int M = n / 2;
    if (M <= CKB_MAX_SINGLE_PASS && n <= c->tableCount) {
        // the cooperative kernel addresses both arrays in 8-byte units
        if (kind == K_R2C) {
            KernelParams p{ (const cf*) in, (cf*) out, table, c->log2Table, batch, in_stride / 2, out_stride };
            return launch_r2c(M, p, s);
        }
        KernelParams p{ (const cf*) in, (cf*) out, table, c->log2Table, batch, in_stride, out_stride / 2 };
        return launch_c2r(M, p, s);
    }
    return enqueue_large_real(c, kind == K_C2R, n, in, out, batch, in_stride, out_stride, s);
}

// the reference's checks for one transform call (src/ckfft/ckfft.cpp:36-114), plus count <= 0
bool check_call(const _CkFftContext* c, Kind kind, int n, const void* in, const void* out, bool allow_inplace = false)
{
    if (!c || c->magic != kMagic) { set_error("invalid context"); return false; }
    const bool needs_inv = (kind == K_C2C_INV || kind == K_C2R);
    if (needs_inv ? !c->invExpTable : !c->fwdExpTable) { set_error("context was not created for this direction"); return false; }
    if (!is_pow2(n) || n > c->maxCount) { set_error("n must be a power of two <= nMax"); return false; }
    if (!in || !out || (in == out && !allow_inplace)) { set_error("input/output must be distinct non-NULL buffers"); return false; }
    return true;
}

// per-thread staging for host-pointer calls: no mutable state lives in the (shared) context.
// Three ROLE streams -- all H2D copies on one, all kernels on the second, all D2H copies on the third, chained by
// events per buffer slot -- so that each DMA direction always has its next copy queued right behind the current one,
// exactly like a plain back-to-back copy loop.  (Round 1 gave every slot its own stream carrying H2D -> kernel -> D2H:
// a direction then idles whenever all slots happen to be in the other phases.  Fine on an idle link -- 92 of 98 GB/s on
// one GPU -- but with the 8 GPUs of the box sharing the host it reached only 63 % of what plain copies achieve.)
struct Staging
{
    int device = -1;
    static constexpr int kSlots = 4;
    void* d_in[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    void* d_out[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    size_t in_cap = 0, out_cap = 0;
    int in_slots = 0, out_slots = 0;           // buffers currently allocated (a single-chunk call allocates one pair)
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[kSlots] = {nullptr, nullptr, nullptr, nullptr};     // slot's H2D copy has landed
    cudaEvent_t ev_k[kSlots] = {nullptr, nullptr, nullptr, nullptr};      // slot's kernel(s) have finished
    cudaEvent_t ev_out[kSlots] = {nullptr, nullptr, nullptr, nullptr};    // slot's D2H copy has finished: the slot is free

    void release()
    {
        if (device < 0) return;
        int prev = -1;
        cudaGetDevice(&prev);
        cudaSetDevice(device);
        for (int i = 0; i < kSlots; ++i) {
            if (d_in[i]) cudaFree(d_in[i]);
            if (d_out[i]) cudaFree(d_out[i]);
            if (ev_in[i]) cudaEventDestroy(ev_in[i]);
            if (ev_k[i]) cudaEventDestroy(ev_k[i]);
            if (ev_out[i]) cudaEventDestroy(ev_out[i]);
            d_in[i] = d_out[i] = nullptr;
            ev_in[i] = ev_k[i] = ev_out[i] = nullptr;
        }
        for (cudaStream_t* st : { &s_in, &s_k, &s_out }) {
            if (*st) cudaStreamDestroy(*st);
            *st = nullptr;
        }
        in_cap = out_cap = 0;
        in_slots = out_slots = 0;
        if (prev >= 0) cudaSetDevice(prev);
        device = -1;
    }
    // `slots` = chunks that will be in flight (a call with one chunk needs one pair of buffers, not four)
    cudaError_t reserve(int dev, size_t in_bytes, size_t out_bytes, int slots)
    {
        if (device != dev) { release(); device = dev; }
        cudaError_t e;
        for (cudaStream_t* st : { &s_in, &s_k, &s_out })
            if (!*st && (e = cudaStreamCreateWithFlags(st, cudaStreamNonBlocking)) != cudaSuccess) return e;
        for (int i = 0; i < kSlots; ++i)
            for (cudaEvent_t* ev : { &ev_in[i], &ev_k[i], &ev_out[i] })
                if (!*ev && (e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = grow(d_in, in_cap, in_slots, in_bytes, slots)) != cudaSuccess) return e;
        if ((e = grow(d_out, out_cap, out_slots, out_bytes, slots)) != cudaSuccess) return e;
        return cudaSuccess;
    }
    // one side's buffers: `slots` buffers of at least `bytes` each (growing the capacity reallocates all of them)
    cudaError_t grow(void* (&buf)[kSlots], size_t& cap, int& have, size_t bytes, int slots)
    {
        if (bytes > cap) {
            for (int i = 0; i < kSlots; ++i) { if (buf[i]) cudaFree(buf[i]); buf[i] = nullptr; }
            cap = bytes;
            have = 0;
        }
        for (int i = have; i < slots; ++i) {
            cudaError_t e = cudaMalloc(&buf[i], cap);
            if (e != cudaSuccess) { buf[i] = nullptr; drop_buffers(); return e; }
            have = i + 1;
        }
        return cudaSuccess;
    }
    // free the device buffers but keep the streams (after a call whose chunks were unusually large, or a failed reserve)
    void drop_buffers()
    {
        for (int i = 0; i < kSlots; ++i) {
            if (d_in[i]) cudaFree(d_in[i]);
            if (d_out[i]) cudaFree(d_out[i]);
            d_in[i] = d_out[i] = nullptr;
        }
        in_cap = out_cap = 0;
        in_slots = out_slots = 0;
        cudaGetLastError();
    }
    cudaError_t drain()
    {
        cudaError_t e = cudaSuccess;
        for (cudaStream_t st : { s_in, s_k, s_out }) {
            const cudaError_t e1 = st ? cudaStreamSynchronize(st) : cudaSuccess;
            if (e == cudaSuccess) e = e1;
        }
        return e;
    }
    // A worker thread that exits gives its staging buffers back.  (At process teardown the runtime may already be
    // gone; the calls then fail harmlessly and the driver reclaims the memory.)
    ~Staging() { release(); cudaGetLastError(); }
};
thread_local Staging tl_staging;

constexpr size_t kStagingKeepBytes = size_t(256) << 20;   // larger per-slot staging buffers are freed at the end of the call

enum Side { SIDE_HOST, SIDE_DEVICE, SIDE_BAD };

Side classify(const void* p, int dev)
{
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return SIDE_HOST; }
    if (a.type == cudaMemoryTypeDevice) return a.device == dev ? SIDE_DEVICE : SIDE_BAD;
    if (a.type == cudaMemoryTypeManaged) return SIDE_DEVICE;
    return SIDE_HOST;   // unregistered or pinned host memory
}

// Small host calls (the classic single-transform use of the reference, src/test/test.cpp:254-301): the data goes
// through a per-thread pinned, device-mapped bounce buffer and the kernel works on that host memory directly over
// PCIe (the bulk copies and the stores address it like any global memory).  One launch and one stream
// synchronisation instead of two staged copies: 23 us -> ~10 us for one 1024-point transform.
constexpr size_t kSmallCallBytes = size_t(256) << 10;

struct Bounce
{
    void* in = nullptr;
    void* out = nullptr;
    volatile unsigned* flag = nullptr;     // completion word in mapped host memory, written by the GPU in stream order
    void* dflag = nullptr;                 // its device address
    unsigned seq = 0;
    bool tried = false;
    bool ensure()
    {
        if (!tried) {
            tried = true;
            void* f = nullptr;
            if (cudaHostAlloc(&in, kSmallCallBytes, cudaHostAllocMapped) != cudaSuccess) in = nullptr;
            if (cudaHostAlloc(&out, kSmallCallBytes, cudaHostAllocMapped) != cudaSuccess) out = nullptr;
            if (cudaHostAlloc(&f, 64, cudaHostAllocMapped) == cudaSuccess) {
                memset(f, 0, 64);
                if (cudaHostGetDevicePointer(&dflag, f, 0) == cudaSuccess) flag = (volatile unsigned*) f;
                else cudaFreeHost(f);
            }
            cudaGetLastError();
        }
        return in && out;
    }
    ~Bounce() { if (in) cudaFreeHost(in); if (out) cudaFreeHost(out); if (flag) cudaFreeHost((void*) flag); cudaGetLastError(); }
};
thread_local Bounce tl_bounce;

// cuStreamWriteValue32 (a stream-ordered 4-byte write, no kernel), fetched like the tensor-map encoder: no libcuda link
typedef int (*StreamWriteValue32Fn)(void* stream, unsigned long long dptr, unsigned value, unsigned flags);
StreamWriteValue32Fn stream_write_value32()
{
    static StreamWriteValue32Fn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (getenv_flag("CKFFT_B200_SPIN_SYNC", 0) == 0 ||
            cudaGetDriverEntryPoint("cuStreamWriteValue32", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            f = nullptr;
        cudaGetLastError();
        return (StreamWriteValue32Fn) f;
    }();
    return fn;
}

// Completion of a small call.  Experiment: instead of cudaStreamSynchronize the stream writes a
// sequence number into mapped host memory behind the kernel (cuStreamWriteValue32) and the calling thread spins on it.
// Measured on B200 (tools/latency_probe.py, one CkFftComplexForward on host buffers, median of 3000 calls through ctypes):
// N = 1024 16.25 us vs 16.80 us with cudaStreamSynchronize, N = 4096 23.5 vs 22.4 us -- the wait path of the driver is not
// where the time goes (launch ~ 5 us, the kernel's PCIe round trips ~ 4 us, the write-value operation costs what the
// cheaper wait saves), so it is OFF by default (CKFFT_B200_SPIN_SYNC=1 enables it).  Falls back to cudaStreamSynchronize if
// the word does not arrive within ~2 ms (a faulting kernel never writes it; the synchronisation then reports the error).
cudaError_t finish_small_call(Bounce& b, cudaStream_t s)
{
    StreamWriteValue32Fn wr = stream_write_value32();
    if (!wr || !b.flag) return cudaStreamSynchronize(s);
    const unsigned want = ++b.seq;
    if (wr((void*) s, (unsigned long long) (uintptr_t) b.dflag, want, 0) != 0) return cudaStreamSynchronize(s);
    for (int spins = 0; spins < 400000; ++spins) {
        if (*b.flag == want) {
            std::atomic_thread_fence(std::memory_order_acquire);
            return cudaSuccess;
        }
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
    return cudaStreamSynchronize(s);
}

// returns 1 done, 0 failed, -1 not applicable (caller takes the chunked path)
int run_host_small(const _CkFftContext* c, Kind kind, int n, const void* in, void* out, size_t batch)
{
    const size_t ib = in_elems(kind, n) * in_elem_bytes(kind) * batch;
    const size_t ob = out_elems(kind, n) * out_elem_bytes(kind) * batch;
    if (ib > kSmallCallBytes || ob > kSmallCallBytes || n > CKB_MAX_SINGLE_PASS) return -1;
    Bounce& b = tl_bounce;
    if (!b.ensure()) return -1;
    void *din = nullptr, *dout = nullptr;
    if (cudaHostGetDevicePointer(&din, b.in, 0) != cudaSuccess || cudaHostGetDevicePointer(&dout, b.out, 0) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    memcpy(b.in, in, ib);
    cudaError_t e = enqueue(c, kind, n, din, dout, (long long) batch, (long long) in_elems(kind, n),
                            (long long) out_elems(kind, n), cudaStreamPerThread);
    if (e == cudaSuccess) e = finish_small_call(b, cudaStreamPerThread);
    if (e != cudaSuccess) { set_error("transform failed", e); return 0; }
    memcpy(out, b.out, ob);
    return 1;
}

// Pageable host arrays: cudaMemcpyAsync on them is staged through the driver's bounce buffers and is synchronous, so
// nothing overlaps (measured: 7 - 13 GB/s end to end against 75 - 95 GB/s for pinned arrays).  Page-locking the caller's
// arrays for the duration of the call (cudaHostRegister) makes the copies asynchronous -- but registering and
// unregistering costs about as much as the staged copy it saves: measured on the B200 box, 2 GiB in + 2 GiB out,
// 12.7 GB/s with registration vs 12.9 without on one GPU, 15.2 vs 26.6 with two processes registering at once.  It is
// therefore OPT-IN (CKFFT_B200_PIN=1, calls that move at least kPinThresholdBytes); a caller who reuses its buffers
// should pin them once itself (CkFftB200HostAlloc or cudaHostRegister), which is what makes the host path fast.
constexpr size_t kPinThresholdBytes = size_t(64) << 20;

struct ScopedPin
{
    void* p = nullptr;
    ScopedPin(const void* ptr, size_t bytes, bool read_only)
    {
        if (bytes == 0) return;
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return; }
        if (a.type != cudaMemoryTypeUnregistered) return;                       // already pinned (or not plain host memory)
        cudaError_t e = cudaErrorUnknown;
        if (read_only) e = cudaHostRegister((void*) ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterReadOnly);
        if (e != cudaSuccess) { cudaGetLastError(); e = cudaHostRegister((void*) ptr, bytes, cudaHostRegisterPortable); }
        if (e == cudaSuccess) p = (void*) ptr;
        else cudaGetLastError();                                                // refused: the copies simply stay synchronous
    }
    ~ScopedPin() { if (p) { cudaHostUnregister(p); cudaGetLastError(); } }
};

// Host arrays: stream the batch through the GPU in chunks, up to four in flight
// (H2D of chunk i+1, kernel of chunk i and D2H of chunk i-1 overlap on the three role streams of Staging).
// `cursor` (multi-device scheduler, multi.cu): several threads -- one per device -- work on the SAME batch and draw chunk
// numbers from the shared counter, so a device behind a slower host link simply takes fewer chunks.
int run_host(const _CkFftContext* c, Kind kind, int n, const void* in, void* out, size_t batch, std::atomic<size_t>* cursor = nullptr)
{
    const size_t ib = in_elems(kind, n) * in_elem_bytes(kind);     // bytes per transform
    const size_t ob = out_elems(kind, n) * out_elem_bytes(kind);
    const bool pin = (ib + ob) * batch >= kPinThresholdBytes && getenv_flag("CKFFT_B200_PIN", 0);
    ScopedPin pin_in(in, pin ? ib * batch : 0, true), pin_out(out, pin ? ob * batch : 0, false);
    static const size_t target = [] {                              // input bytes per chunk (default 32 MiB)
        const char* e = getenv("CKFFT_B200_CHUNK_MB");
        const long mb = e ? atol(e) : 0;
        return size_t(mb > 0 && mb <= 4096 ? mb : 32) << 20;
    }();
    size_t per_chunk = target / ib;
    if (per_chunk < 1) per_chunk = 1;
    if (per_chunk > batch) per_chunk = batch;
    // device buffers are padded to 16 bytes so that every slot keeps 8-byte aligned transforms
    Staging& st = tl_staging;
    if (cursor) {
        // shared batch: keep the chunk size every thread derives identical, and small enough that all devices get work
        const size_t share = (batch + 63) / 64;
        if (per_chunk > share) per_chunk = share ? share : 1;
    }
    const size_t nchunks = (batch + per_chunk - 1) / per_chunk;
    const int slots = nchunks < (size_t) Staging::kSlots ? (int) nchunks : Staging::kSlots;
    cudaError_t e = st.reserve(c->device, per_chunk * ib + 16, per_chunk * ob + 16, slots);
    if (e != cudaSuccess) { set_error("staging allocation", e); return 0; }
    // Staging stays with the thread for the next call -- unless one transform is so long that a single chunk pins
    // hundreds of MiB of device memory (n >= 2^25): that is handed back when the call ends.
    struct GiveBack {
        Staging& st; bool on;
        ~GiveBack() { if (on) st.drop_buffers(); }
    } give_back{ st, per_chunk * ib > kStagingKeepBytes || per_chunk * ob > kStagingKeepBytes };

    size_t done = 0;
    for (size_t chunk = 0; done < batch; ++chunk) {
        if (cursor) {                                             // next unclaimed chunk of the shared batch
            const size_t ci = cursor->fetch_add(1, std::memory_order_relaxed);
            if (ci >= nchunks) break;
            done = ci * per_chunk;
        }
        const size_t cnt = (batch - done < per_chunk) ? batch - done : per_chunk;
        const int slot = (int) (chunk % (size_t) slots);
        // the slot's previous chunk must have left the device before its buffers are reused (host-side wait: the host
        // runs at most `slots` chunks ahead of the D2H stream)
        if (chunk >= (size_t) slots && (e = cudaEventSynchronize(st.ev_out[slot])) != cudaSuccess) { set_error("event sync", e); st.drain(); return 0; }
        if ((e = cudaMemcpyAsync(st.d_in[slot], (const char*) in + done * ib, cnt * ib, cudaMemcpyHostToDevice, st.s_in)) != cudaSuccess ||
            (e = cudaEventRecord(st.ev_in[slot], st.s_in)) != cudaSuccess) {
            set_error("H2D copy", e); st.drain(); return 0;
        }
        if ((e = cudaStreamWaitEvent(st.s_k, st.ev_in[slot], 0)) == cudaSuccess)
            e = enqueue(c, kind, n, st.d_in[slot], st.d_out[slot], (long long) cnt,
                        (long long) in_elems(kind, n), (long long) out_elems(kind, n), st.s_k);
        if (e == cudaSuccess) e = cudaEventRecord(st.ev_k[slot], st.s_k);
        if (e != cudaSuccess) { set_error("kernel launch", e); st.drain(); return 0; }
        if ((e = cudaStreamWaitEvent(st.s_out, st.ev_k[slot], 0)) != cudaSuccess ||
            (e = cudaMemcpyAsync((char*) out + done * ob, st.d_out[slot], cnt * ob, cudaMemcpyDeviceToHost, st.s_out)) != cudaSuccess ||
            (e = cudaEventRecord(st.ev_out[slot], st.s_out)) != cudaSuccess) {
            set_error("D2H copy", e); st.drain(); return 0;
        }
        done += cnt;
    }
    if ((e = st.drain()) != cudaSuccess) { set_error("transform failed", e); return 0; }
    return 1;
}

// ---- pageable host arrays: library-side staging ------------------------------------------------------------------------
// What a drop-in caller of the reference hands over is malloc'ed memory.  cudaMemcpyAsync on it goes through the driver's
// own bounce buffer, synchronously (measured: H2D 11 GB/s, D2H 21 GB/s, nothing overlaps: 13-15 GB/s end to end, less than
// the reference reaches on the 16 host cores of the same box), and page-locking the arrays per call costs what it saves
// (ScopedPin above).  So large calls on pageable arrays are staged by the library itself: a team of host threads copies
// chunk i + 1 from the caller's array into a pinned slot while the copy engines move chunk i and the GPU transforms it, and
// a second team behind an output thread copies finished chunks from their pinned slots into the caller's output array.
// Both DMA directions, the kernels and both host copies overlap; the pinned slots stay with the calling thread.
// The teams copy with non-temporal stores (host_copy.cpp): measured on the 16-core B200 box, 2 GiB in + 2 GiB out, memcpy
// 44 GB/s -> 56-58 GB/s end to end (the pipeline is bound by host-memory traffic, and memcpy read every destination line first).
class CopyTeam
{
public:
    explicit CopyTeam(int n) : n_(n < 1 ? 1 : n)
    {
        try {
            th_.reserve((size_t) n_);
            for (int i = 1; i < n_; ++i) th_.emplace_back([this, i] { loop(i); });
        } catch (...) { }                          // out of threads: the team is as large as it got (the caller always copies, too)
        n_ = 1 + (int) th_.size();
    }
    ~CopyTeam()
    {
        { std::lock_guard<std::mutex> l(m_); stop_ = true; ++gen_; }
        cv_.notify_all();
        for (std::thread& t : th_) t.join();
    }
    // blocking: the caller copies the first piece itself
    void copy(void* dst, const void* src, size_t bytes)
    {
        const int nt = ckb::stream_copy_level();          // non-temporal stores by default (host_copy.cpp)
        { std::lock_guard<std::mutex> l(m_); dst_ = (char*) dst; src_ = (const char*) src; bytes_ = bytes; nt_ = nt; pending_ = n_ - 1; ++gen_; }
        cv_.notify_all();
        piece(0);
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [&] { return pending_ == 0; });
    }

private:
    void piece(int i) const
    {
        const size_t per = ((bytes_ + (size_t) n_ - 1) / (size_t) n_ + 4095) & ~size_t(4095);     // whole pages per thread
        const size_t lo = (size_t) i * per < bytes_ ? (size_t) i * per : bytes_;
        const size_t hi = lo + per < bytes_ ? lo + per : bytes_;
        if (hi > lo) ckb::stream_copy(dst_ + lo, src_ + lo, hi - lo, nt_);
    }
    void loop(int i)
    {
        unsigned long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            piece(i);
            std::lock_guard<std::mutex> l(m_);
            if (--pending_ == 0) done_.notify_one();
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    unsigned long gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
    char* dst_ = nullptr;
    const char* src_ = nullptr;
    size_t bytes_ = 0;
    int nt_ = 1;
};

// pinned host staging slots of the calling thread (kept between calls: cudaHostAlloc costs milliseconds)
struct PinnedSlots
{
    static constexpr int kSlots = Staging::kSlots;
    void* h_in[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    void* h_out[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    size_t in_cap = 0, out_cap = 0;
    static cudaError_t grow(void* (&buf)[kSlots], size_t& cap, size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        for (int i = 0; i < kSlots; ++i) { if (buf[i]) cudaFreeHost(buf[i]); buf[i] = nullptr; }
        cap = 0;
        for (int i = 0; i < kSlots; ++i) {
            cudaError_t e = cudaHostAlloc(&buf[i], bytes, cudaHostAllocPortable);
            if (e != cudaSuccess) { buf[i] = nullptr; return e; }
        }
        cap = bytes;
        return cudaSuccess;
    }
    cudaError_t reserve(size_t in_bytes, size_t out_bytes)
    {
        cudaError_t e = grow(h_in, in_cap, in_bytes);
        return e != cudaSuccess ? e : grow(h_out, out_cap, out_bytes);
    }
    ~PinnedSlots()
    {
        for (int i = 0; i < kSlots; ++i) { if (h_in[i]) cudaFreeHost(h_in[i]); if (h_out[i]) cudaFreeHost(h_out[i]); }
        cudaGetLastError();
    }
};
thread_local PinnedSlots tl_pinned;

constexpr size_t kPageableStageBytes = size_t(64) << 20;     // smaller calls are not worth the helper threads
constexpr size_t kPageableMaxRowBytes = size_t(64) << 20;    // longer transforms (n > 2^23) would pin gigabytes of staging slots

bool is_pageable(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

int host_copy_threads()
{
    static const int n = [] {
        const char* e = getenv("CKFFT_B200_HOST_THREADS");
        long v = e ? atol(e) : 0;
        // per team (there are two): measured on the 16-core B200 box, 2 GiB in + 2 GiB out: 2 threads 22.9 GB/s, 4 35.8, 8 44.8
        // (12 threads 34, 16 threads 41: two teams of 8 are the 16 cores; driver-staged copies 12.9, the reference on all 16 cores
        // 37.5); the call is synchronous, the cores are the caller's
        if (v <= 0) { v = (long) std::thread::hardware_concurrency() / 2; if (v < 2) v = 2; if (v > 8) v = 8; }
        return (int) (v > 64 ? 64 : v);
    }();
    return n;
}

// returns 1 done, 0 failed, -1 not applicable (no helper thread could be started: the caller takes the plain path)
// Threads of ONE call that share the host with sibling calls of the same batch (the multi-device scheduler's workers, multi.cu)
// divide the copy teams among them.
thread_local int tl_host_sharers = 1;

// `cursor` (multi-device scheduler): this thread's device works on a batch it shares with other devices and draws chunk numbers
// from the shared counter, like run_host.
int run_host_pageable(const _CkFftContext* c, Kind kind, int n, const void* in, void* out, size_t batch, std::atomic<size_t>* cursor = nullptr)
{
    const size_t ib = in_elems(kind, n) * in_elem_bytes(kind);
    const size_t ob = out_elems(kind, n) * out_elem_bytes(kind);
    static const size_t target = [] {                              // input bytes per chunk (default 16 MiB)
        const char* e = getenv("CKFFT_B200_PAGEABLE_CHUNK_MB");
        const long mb = e ? atol(e) : 0;
        return size_t(mb > 0 && mb <= 1024 ? mb : 16) << 20;
    }();
    size_t per_chunk = target / ib;
    if (per_chunk < 1) per_chunk = 1;
    if (per_chunk > batch) per_chunk = batch;
    if (cursor) {                                   // shared batch: every thread derives the same chunk size, all devices get work
        const size_t share = (batch + 63) / 64;
        if (per_chunk > share) per_chunk = share ? share : 1;
    }
    const size_t nchunks = (batch + per_chunk - 1) / per_chunk;
    const int slots = nchunks < (size_t) Staging::kSlots ? (int) nchunks : Staging::kSlots;
    Staging& st = tl_staging;
    PinnedSlots& ps = tl_pinned;
    cudaError_t e = st.reserve(c->device, per_chunk * ib + 16, per_chunk * ob + 16, slots);
    if (e == cudaSuccess) e = ps.reserve(per_chunk * ib, per_chunk * ob);
    if (e != cudaSuccess) { set_error("staging allocation", e); cudaGetLastError(); return 0; }

    const int team = host_copy_threads() / tl_host_sharers > 1 ? host_copy_threads() / tl_host_sharers : 1;
    CopyTeam team_in(team), team_out(team);
    std::mutex m;
    std::condition_variable cv;
    size_t queued = 0, copied = 0;          // this thread's chunks enqueued on the device / copied out to the caller's array
    std::vector<size_t> drawn;              // their positions in the batch (local chunk k = chunk drawn[k] of the batch)
    drawn.reserve(nchunks);
    bool closed = false;                    // no more chunks will be queued
    bool failed = false;
    cudaError_t err = cudaSuccess;
    const char* where = "";
    auto fail = [&](const char* w, cudaError_t ce) {
        { std::lock_guard<std::mutex> l(m); if (!failed) { failed = true; err = ce; where = w; } }
        cv.notify_all();
    };
    const int device = c->device;
    std::thread out_thread;
    try {
        out_thread = std::thread([&] {
        cudaSetDevice(device);
        for (size_t i = 0;; ++i) {
            size_t ci;
            {
                std::unique_lock<std::mutex> l(m);
                cv.wait(l, [&] { return queued > i || closed || failed; });
                if (queued <= i) return;
                ci = drawn[i];
            }
            const int slot = (int) (i % (size_t) slots);
            const cudaError_t ce = cudaEventSynchronize(st.ev_out[slot]);
            if (ce != cudaSuccess) { fail("D2H copy", ce); return; }
            const size_t off = ci * per_chunk, cnt = batch - off < per_chunk ? batch - off : per_chunk;
            team_out.copy((char*) out + off * ob, ps.h_out[slot], cnt * ob);
            { std::lock_guard<std::mutex> l(m); copied = i + 1; }
            cv.notify_all();
        }
        });
    } catch (...) {
        return -1;                               // no thread to be had: the caller takes the plain path
    }
    for (size_t i = 0;; ++i) {
        const size_t ci = cursor ? cursor->fetch_add(1, std::memory_order_relaxed) : i;      // next (unclaimed) chunk of the batch
        if (ci >= nchunks) break;
        const int slot = (int) (i % (size_t) slots);
        if (i >= (size_t) slots) {               // the slot's previous chunk has left for the caller's array (so its H2D copy is long done)
            std::unique_lock<std::mutex> l(m);
            cv.wait(l, [&] { return copied + (size_t) slots > i || failed; });
            if (failed) break;
        }
        const size_t off = ci * per_chunk, cnt = batch - off < per_chunk ? batch - off : per_chunk;
        team_in.copy(ps.h_in[slot], (const char*) in + off * ib, cnt * ib);
        if ((e = cudaMemcpyAsync(st.d_in[slot], ps.h_in[slot], cnt * ib, cudaMemcpyHostToDevice, st.s_in)) != cudaSuccess ||
            (e = cudaEventRecord(st.ev_in[slot], st.s_in)) != cudaSuccess) { fail("H2D copy", e); break; }
        if ((e = cudaStreamWaitEvent(st.s_k, st.ev_in[slot], 0)) == cudaSuccess)
            e = enqueue(c, kind, n, st.d_in[slot], st.d_out[slot], (long long) cnt,
                        (long long) in_elems(kind, n), (long long) out_elems(kind, n), st.s_k);
        if (e == cudaSuccess) e = cudaEventRecord(st.ev_k[slot], st.s_k);
        if (e != cudaSuccess) { fail("kernel launch", e); break; }
        if ((e = cudaStreamWaitEvent(st.s_out, st.ev_k[slot], 0)) != cudaSuccess ||
            (e = cudaMemcpyAsync(ps.h_out[slot], st.d_out[slot], cnt * ob, cudaMemcpyDeviceToHost, st.s_out)) != cudaSuccess ||
            (e = cudaEventRecord(st.ev_out[slot], st.s_out)) != cudaSuccess) { fail("D2H copy", e); break; }
        { std::lock_guard<std::mutex> l(m); drawn.push_back(ci); queued = i + 1; }
        cv.notify_all();
    }
    { std::lock_guard<std::mutex> l(m); closed = true; }
    cv.notify_all();
    out_thread.join();
    e = st.drain();
    if (failed) { set_error(where, err); return 0; }
    if (e != cudaSuccess) { set_error("transform failed", e); return 0; }
    return 1;
}

bool supported_size(const _CkFftContext* c, Kind kind, int n)
{
    (void) c; (void) kind;
    return n <= (1 << 30);    // single pass up to 16384 complex / 32768 real points, multi-pass above
}

// large calls on two pageable arrays go through the library's own staging (run_host_pageable)
bool wants_pageable_staging(Kind kind, int n, const void* in, const void* out, size_t batch)
{
    const size_t ib = in_elems(kind, n) * in_elem_bytes(kind), ob = out_elems(kind, n) * out_elem_bytes(kind);
    return (ib + ob) * batch >= kPageableStageBytes && ib <= kPageableMaxRowBytes && getenv_flag("CKFFT_B200_PAGEABLE_PIPE", 1) &&
           !getenv_flag("CKFFT_B200_PIN", 0) && is_pageable(in) && is_pageable(out);
}

// shared body of the synchronous entry points
int run_sync(CkFftContext* c, Kind kind, int n, const void* in, void* out, size_t batch)
{
    if (!check_call(c, kind, n, in, out)) return 0;
    if (batch == 0) return 1;
    if (!supported_size(c, kind, n)) { set_error("transform length not supported by this build"); return 0; }
    DeviceGuard guard(c->device);
    if (!guard.ok) { set_error("cannot select the context's device"); return 0; }
    const Side si = classify(in, c->device), so = classify(out, c->device);
    if (si == SIDE_BAD || so == SIDE_BAD || si != so) {
        set_error("input and output must both be host memory or both be memory of the context's device");
        return 0;
    }
    if (si == SIDE_HOST) {
        const int small = run_host_small(c, kind, n, in, out, batch);
        if (small >= 0) return small;
        if (wants_pageable_staging(kind, n, in, out, batch)) {
            const int staged = run_host_pageable(c, kind, n, in, out, batch);
            if (staged >= 0) return staged;
        }
        return run_host(c, kind, n, in, out, batch);
    }
    if (((uintptr_t) in | (uintptr_t) out) & 7) { set_error("device pointers must be 8-byte aligned"); return 0; }
    cudaError_t e = enqueue(c, kind, n, in, out, (long long) batch, (long long) in_elems(kind, n),
                            (long long) out_elems(kind, n), cudaStreamPerThread);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamPerThread);
    if (e != cudaSuccess) { set_error("transform failed", e); return 0; }
    return 1;
}

// In-place rule of the stream-ordered calls: `input == output` is accepted when every transform
// reads and writes the same bytes -- complex: equal strides; real: the real-array stride is twice the spectrum
// stride (rows padded to n + 2 floats, the usual in-place real layout).  Every kernel reads a whole row (tile) into
// registers / shared memory before it stores any of it, and the multi-pass paths write `output` only after the
// pass that reads `input` has consumed that transform, so such calls are safe.  Partial overlap stays undefined.
bool inplace_strides_ok(Kind kind, size_t in_stride, size_t out_stride)
{
    if (kind == K_R2C) return in_stride == 2 * out_stride;
    if (kind == K_C2R) return out_stride == 2 * in_stride;
    return in_stride == out_stride;
}

int run_async(CkFftContext* c, Kind kind, int n, const void* in, void* out, size_t batch, size_t in_stride,
              size_t out_stride, void* stream)
{
    if (!check_call(c, kind, n, in, out, true)) return 0;
    if (batch == 0) return 1;
    if (!supported_size(c, kind, n)) { set_error("transform length not supported by this build"); return 0; }
    if (in_stride == 0) in_stride = in_elems(kind, n);
    if (out_stride == 0) out_stride = out_elems(kind, n);
    if (in_stride < in_elems(kind, n) || out_stride < out_elems(kind, n)) { set_error("stride shorter than one transform"); return 0; }
    if (in == out && !inplace_strides_ok(kind, in_stride, out_stride)) {
        set_error("in-place call: input and output rows must coincide (complex: equal strides; real: real stride = 2 x spectrum stride)");
        return 0;
    }
    if (((uintptr_t) in | (uintptr_t) out) & 7) { set_error("device pointers must be 8-byte aligned"); return 0; }
    if ((kind == K_R2C && n > 16 && (in_stride & 1)) || (kind == K_C2R && n > 16 && (out_stride & 1))) {
        set_error("real-array strides must be even");
        return 0;
    }
    DeviceGuard guard(c->device);
    if (!guard.ok) { set_error("cannot select the context's device"); return 0; }
    cudaError_t e = enqueue(c, kind, n, in, out, (long long) batch, (long long) in_stride, (long long) out_stride,
                            (cudaStream_t) stream);
    if (e != cudaSuccess) { set_error("kernel launch", e); return 0; }
    return 1;
}

}  // namespace

namespace ckb {
void set_last_error(const char* text) { set_error(text); }      // multi.cu reports through the same channel

// multi.cu: one device's share of a batch in HOST arrays that several devices work on together (chunks are drawn from
// `cursor`).  kind: 0 complex forward, 1 complex inverse, 2 real forward, 3 real inverse.  The caller has validated the call.
int run_host_shared(CkFftContext* c, int kind, int n, const void* in, void* out, size_t batch, std::atomic<size_t>* cursor)
{
    if (!c || c->magic != kMagic) { set_error("invalid context"); return 0; }
    DeviceGuard guard(c->device);
    if (!guard.ok) { set_error("cannot select the context's device"); return 0; }
    if (wants_pageable_staging((Kind) kind, n, in, out, batch)) {
        const int staged = run_host_pageable(c, (Kind) kind, n, in, out, batch, cursor);
        // (-1 = no helper thread to be had BEFORE any chunk was drawn: the plain path takes over)
        if (staged >= 0) return staged;
    }
    return run_host(c, (Kind) kind, n, in, out, batch, cursor);
}

// multi.cu: how many worker threads of one BatchMulti call share the host (divides the pageable-staging copy teams)
void set_host_sharers(int n) { tl_host_sharers = n < 1 ? 1 : n; }

// Stream-ordered scratch from the library's OWN memory pool (one per device).  The default pool of cudaMallocAsync
// gives unused memory back to the driver at every synchronisation (release threshold 0), so a caller who synchronises
// between calls -- every classic CkFft* call does -- paid a fresh cudaMalloc of the multi-pass scratch (64 MiB .. 2 GiB)
// on every call: measured 2 x the kernel time for real n = 2^17 .. 2^21.  The private pool keeps up to
// CKFFT_B200_POOL_KEEP_MB (default 2304) MiB between calls and touches nobody else's allocator settings.
cudaError_t scratch_alloc(void** ptr, size_t bytes, cudaStream_t s)
{
    static cudaMemPool_t pools[64] = {nullptr};
    static std::atomic<int> state[64];            // 0 none, 1 being created, 2 ready, 3 unavailable (fall back to the default pool)
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaMallocAsync(ptr, bytes, s);
    int st = state[dev].load(std::memory_order_acquire);
    if (st == 0) {
        int expected = 0;
        if (state[dev].compare_exchange_strong(expected, 1, std::memory_order_acq_rel)) {
            cudaMemPoolProps props;
            memset(&props, 0, sizeof(props));
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            cudaMemPool_t pool = nullptr;
            bool ok = cudaMemPoolCreate(&pool, &props) == cudaSuccess;
            if (ok) {
                const char* env = getenv("CKFFT_B200_POOL_KEEP_MB");
                const long long mb = env && *env ? atoll(env) : 2304;
                unsigned long long keep = (unsigned long long) (mb < 0 ? 0 : mb) << 20;
                ok = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess;
            }
            if (ok) pools[dev] = pool;
            else cudaGetLastError();
            state[dev].store(ok ? 2 : 3, std::memory_order_release);
        }
        while ((st = state[dev].load(std::memory_order_acquire)) == 1) { }
    } else {
        while (st == 1) st = state[dev].load(std::memory_order_acquire);
    }
    if (st != 2) return cudaMallocAsync(ptr, bytes, s);
    return cudaMallocFromPoolAsync(ptr, bytes, pools[dev], s);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count_of_current_device()
{
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 1;
    if (cache[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 1;
        cache[dev] = n;
    }
    return cache[dev];
}
}  // namespace ckb

// ---------------------------------------------------------------------------------------------
// exported C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

CkFftContext* CkFftInit(int maxCount, CkFftDirection direction, void* userBuf, size_t* userBufSize)
{
    // src/ckfft/ckfft.cpp:16-31
    if (maxCount <= 0 || !is_pow2(maxCount)) { set_error("nMax must be a positive power of two"); return NULL; }
    if (direction != kCkFftDirection_Forward && direction != kCkFftDirection_Inverse && direction != kCkFftDirection_Both) {
        set_error("invalid direction");
        return NULL;
    }
    if (userBuf && !userBufSize) { set_error("buf given without bufSize"); return NULL; }

    // src/ckfft/context.cpp:27-51.  The host tables keep the reference's meaning (non-NULL table ==
    // direction available) but only the Nt = min(nMax, CKB_MAX_TABLE) entries the kernels can use
    // are materialised; larger transforms derive their twiddles on the device.
    const int nt = maxCount < CKB_MAX_TABLE ? maxCount : CKB_MAX_TABLE;
    const size_t tableBytes = (size_t) nt * sizeof(CkFftComplex);
    size_t need = context_bytes();
    if (direction & kCkFftDirection_Forward) need += tableBytes;
    if (direction & kCkFftDirection_Inverse) need += tableBytes;
    if (userBufSize && (!userBuf || *userBufSize < need)) {
        *userBufSize = need;
        return NULL;
    }

    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { set_error("no usable CUDA device (this library has no CPU path)", e); return NULL; }

    void* buf = userBuf ? userBuf : malloc(need);
    if (!buf) { set_error("out of host memory"); return NULL; }
    _CkFftContext* c = new (buf) _CkFftContext();
    memset(c, 0, sizeof(*c));

    CkFftComplex* tab = (CkFftComplex*) ((char*) buf + context_bytes());
    CkFftComplex* fwd = NULL;
    CkFftComplex* inv = NULL;
    if (direction & kCkFftDirection_Forward) { fwd = tab; tab += nt; }
    if (direction & kCkFftDirection_Inverse) { inv = tab; }

    // src/ckfft/context.cpp:90-105: the angle is formed in fp32 exactly like the reference, so the
    // values are bit-identical to the reference table entries for the same angle
    // (for power-of-two sizes table_Nt[k] == table_nMax[k * nMax / Nt]).
    CkFftComplex* master = fwd ? fwd : inv;
    for (int i = 0; i < nt; ++i) {
        float theta = -2.0f * (float) M_PI * i / nt;
        float cs = cosf(theta);
        float sn = sinf(theta);
        if (fwd) { fwd[i].real = cs; fwd[i].imag = sn; }
        if (inv) { inv[i].real = cs; inv[i].imag = -sn; }
    }

    float2* dTable = NULL;
    e = cudaMalloc((void**) &dTable, tableBytes);
    if (e == cudaSuccess) {
        if (fwd) {
            e = cudaMemcpy(dTable, fwd, tableBytes, cudaMemcpyHostToDevice);
        } else {
            // device table always holds the forward sign; conjugate the inverse host table on the way
            CkFftComplex* tmp = (CkFftComplex*) malloc(tableBytes);
            if (!tmp) e = cudaErrorMemoryAllocation;
            else {
                for (int i = 0; i < nt; ++i) { tmp[i].real = master[i].real; tmp[i].imag = -master[i].imag; }
                e = cudaMemcpy(dTable, tmp, tableBytes, cudaMemcpyHostToDevice);
                free(tmp);
            }
        }
    }
    if (e != cudaSuccess) {
        set_error("device twiddle table", e);
        if (dTable) cudaFree(dTable);
        if (!userBuf) free(buf);
        return NULL;
    }

    // two-level inter-pass twiddles for multi-pass lengths, in double precision:
    // W_Tmax^e = hi[e >> h] * lo[e & (2^h - 1)]
    float2* dLo = NULL;
    float2* dHi = NULL;
    int twH = 0, log2Tmax = 0;
    if (maxCount > CKB_MAX_SINGLE_PASS) {
        log2Tmax = ilog2i(maxCount);
        twH = (log2Tmax + 1) / 2;
        const size_t nLo = size_t(1) << twH, nHi = size_t(1) << (log2Tmax - twH);
        float2* host = (float2*) malloc((nLo + nHi) * sizeof(float2));
        if (!host) e = cudaErrorMemoryAllocation;
        else {
            const double step = -2.0 * M_PI / (double) ((long long) 1 << log2Tmax);
            for (size_t j = 0; j < nLo; ++j) { host[j].x = (float) cos(step * (double) j); host[j].y = (float) sin(step * (double) j); }
            for (size_t i = 0; i < nHi; ++i) {
                const double a = step * (double) (i << twH);
                host[nLo + i].x = (float) cos(a);
                host[nLo + i].y = (float) sin(a);
            }
            e = cudaMalloc((void**) &dLo, (nLo + nHi) * sizeof(float2));
            if (e == cudaSuccess) e = cudaMemcpy(dLo, host, (nLo + nHi) * sizeof(float2), cudaMemcpyHostToDevice);
            dHi = dLo ? dLo + nLo : NULL;
            free(host);
        }
        if (e != cudaSuccess) {
            set_error("device twiddle tables (multi-pass)", e);
            if (dLo) cudaFree(dLo);
            cudaFree(dTable);
            if (!userBuf) free(buf);
            return NULL;
        }
    }

    c->neon = false;
    c->maxCount = maxCount;
    c->fwdExpTable = fwd;
    c->invExpTable = inv;
    c->ownBuf = (userBuf == NULL);
    c->magic = kMagic;
    c->device = dev;
    c->tableCount = nt;
    c->log2Table = ilog2i(nt);
    c->dTable = dTable;
    c->dTwLo = dLo;
    c->dTwHi = dHi;
    c->twH = twH;
    c->log2Tmax = log2Tmax;
    return c;
}

void CkFftShutdown(CkFftContext* c)
{
    // src/ckfft/context.cpp:116-122; device memory is ours whoever owns the host buffer
    if (!c || c->magic != kMagic) return;
    {
        DeviceGuard guard(c->device);
        if (c->dTable) cudaFree(c->dTable);
        if (c->dTwLo) cudaFree(c->dTwLo);     // dTwHi lives in the same allocation
    }
    c->dTable = NULL;
    c->magic = 0;
    if (c->ownBuf) free(c);
}

int CkFftComplexForward(CkFftContext* c, int n, const CkFftComplex* in, CkFftComplex* out)
{
    return run_sync(c, K_C2C_FWD, n, in, out, 1);
}

int CkFftComplexInverse(CkFftContext* c, int n, const CkFftComplex* in, CkFftComplex* out)
{
    return run_sync(c, K_C2C_INV, n, in, out, 1);
}

int CkFftRealForward(CkFftContext* c, int n, const float* in, CkFftComplex* out)
{
    return run_sync(c, K_R2C, n, in, out, 1);
}

int CkFftRealInverse(CkFftContext* c, int n, const CkFftComplex* in, float* out, CkFftComplex* tmpBuf)
{
    if (!tmpBuf) { set_error("tmpBuf must not be NULL"); return 0; }   // src/ckfft/ckfft.cpp:57-60
    return run_sync(c, K_C2R, n, in, out, 1);
}

int CkFftComplexForwardBatch(CkFftContext* c, int n, const CkFftComplex* in, CkFftComplex* out, size_t batch)
{
    return run_sync(c, K_C2C_FWD, n, in, out, batch);
}

int CkFftComplexInverseBatch(CkFftContext* c, int n, const CkFftComplex* in, CkFftComplex* out, size_t batch)
{
    return run_sync(c, K_C2C_INV, n, in, out, batch);
}

int CkFftRealForwardBatch(CkFftContext* c, int n, const float* in, CkFftComplex* out, size_t batch)
{
    return run_sync(c, K_R2C, n, in, out, batch);
}

int CkFftRealInverseBatch(CkFftContext* c, int n, const CkFftComplex* in, float* out, CkFftComplex* tmpBuf, size_t batch)
{
    (void) tmpBuf;
    return run_sync(c, K_C2R, n, in, out, batch);
}

int CkFftComplexForwardBatchAsync(CkFftContext* c, int n, const CkFftComplex* in, CkFftComplex* out, size_t batch,
                                  size_t inStride, size_t outStride, void* stream)
{
    return run_async(c, K_C2C_FWD, n, in, out, batch, inStride, outStride, stream);
}

int CkFftComplexInverseBatchAsync(CkFftContext* c, int n, const CkFftComplex* in, CkFftComplex* out, size_t batch,
                                  size_t inStride, size_t outStride, void* stream)
{
    return run_async(c, K_C2C_INV, n, in, out, batch, inStride, outStride, stream);
}

int CkFftRealForwardBatchAsync(CkFftContext* c, int n, const float* in, CkFftComplex* out, size_t batch,
                               size_t inStride, size_t outStride, void* stream)
{
    return run_async(c, K_R2C, n, in, out, batch, inStride, outStride, stream);
}

int CkFftRealInverseBatchAsync(CkFftContext* c, int n, const CkFftComplex* in, float* out, size_t batch,
                               size_t inStride, size_t outStride, void* stream)
{
    return run_async(c, K_C2R, n, in, out, batch, inStride, outStride, stream);
}

int CkFftB200GetPlan(int n, int isReal, CkFftB200Plan* plan)
{
    if (!plan || !is_pow2(n)) return 0;
    memset(plan, 0, sizeof(*plan));
    plan->n = n;
    plan->isReal = isReal ? 1 : 0;
    const int m = isReal ? (n >= 2 ? n / 2 : 1) : n;
    plan->complexPoints = m;
    const bool tiny = isReal ? n <= 64 : n <= 32;      // one thread per transform (tiny_kernel.cuh, small_kernel.cuh)
    if (!isReal && n == 64) {                           // two threads per row: radix 32 x 2 (small_kernel.cuh)
        plan->passes = 1;
        plan->radix[0][0] = 32; plan->radix[0][1] = 2;
        plan->threadsPerTransform = 2;
        plan->elemsPerThread = 32;
        plan->transformsPerCta = 64;
        return 1;
    }
    if (tiny) {
        plan->passes = 1;
        plan->radix[0][0] = m;
        plan->threadsPerTransform = 1;
        plan->elemsPerThread = m;
        plan->transformsPerCta = 128;
        return 1;
    }
    if (m > CKB_MAX_SINGLE_PASS) {
        if (n > (1 << 30)) return 0;
        int npass = 0, L[3];
        ckb::four_step_plan(ilog2i(m), &npass, L);
        plan->passes = npass + ((isReal && m > (1 << 20)) ? 1 : 0);   // real, three-pass lengths: + one element-wise split / twist pass (fused below that)
        for (int i = 0; i < npass && i < 2; ++i) {
            const ckb::PlanRow* r = ckb::find_plan(L[i]);
            plan->radix[i][0] = r->R0;
            plan->radix[i][1] = r->R1;
        }
        const ckb::PlanRow* r0 = ckb::find_plan(L[0]);
        plan->threadsPerTransform = r0->M / r0->E;    // threads per column of a tile
        plan->elemsPerThread = r0->E;
        plan->transformsPerCta = L[0] == 1024 ? 8 : 16;   // columns per tile
        return 1;
    }
    const ckb::PlanRow* r = ckb::find_plan(m);
    if (!r || (isReal && n > CKB_MAX_TABLE)) return 0;
    plan->passes = 1;
    plan->radix[0][0] = r->R0;
    plan->radix[0][1] = r->R1;
    plan->radix[0][2] = r->R2 > 1 ? r->R2 : 0;
    plan->threadsPerTransform = r->M / r->E;
    plan->elemsPerThread = r->E;
    plan->transformsPerCta = r->G;
    plan->sharedBytes = r->smem_bytes;
    return 1;
}

const char* CkFftB200LastError(void) { return tl_error; }

unsigned long long CkFftB200KernelLaunches(void) { return g_launches.load(std::memory_order_relaxed); }

void* CkFftB200HostAlloc(size_t bytes)
{
    void* p = NULL;
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) { set_error("pinned allocation", e); return NULL; }
    return p;
}

void CkFftB200HostFree(void* p)
{
    if (p) cudaFreeHost(p);
}

int CkFftB200ContextDevice(const CkFftContext* c) { return (c && c->magic == kMagic) ? c->device : -1; }

// ---- split-complex ("planar") arrays: re[] and im[] separate, strides in floats ----
static int run_planar(CkFftContext* c, bool inverse, int n, const float* inRe, const float* inIm, float* outRe, float* outIm,
                      size_t batch, size_t inStride, size_t outStride, void* stream)
{
    const Kind kind = inverse ? K_C2C_INV : K_C2C_FWD;
    if (!check_call(c, kind, n, inRe, outRe, true)) return 0;
    if (!inIm || !outIm) { set_error("planar: NULL imaginary array"); return 0; }
    if (n > CKB_MAX_SINGLE_PASS) { set_error("planar: n must be <= 16384"); return 0; }
    if (batch == 0) return 1;
    if (inStride == 0) inStride = (size_t) n;
    if (outStride == 0) outStride = (size_t) n;
    if (inStride < (size_t) n || outStride < (size_t) n) { set_error("stride shorter than one transform"); return 0; }
    if (((uintptr_t) inRe | (uintptr_t) inIm | (uintptr_t) outRe | (uintptr_t) outIm) & 3) { set_error("planar: misaligned pointer"); return 0; }
    // in place = both planes onto themselves with equal strides; any other aliasing between the four arrays is rejected
    const bool inplace = inRe == outRe && inIm == outIm;
    if (inplace ? inStride != outStride : (inRe == outRe || inIm == outIm || (const float*) outRe == inIm || (const float*) outIm == inRe)) {
        set_error("planar: arrays must be distinct, or both planes transformed in place with equal strides");
        return 0;
    }
    if (inRe == inIm || outRe == outIm) { set_error("planar: real and imaginary arrays must be distinct"); return 0; }
    DeviceGuard guard(c->device);
    if (!guard.ok) { set_error("cannot select the context's device"); return 0; }
    ckb::KernelParams p{ (const ckb::cf*) inRe, (ckb::cf*) outRe, c->dTable, c->log2Table, (long long) batch,
                         (long long) inStride, (long long) outStride, nullptr, inIm, outIm };
    cudaError_t e;
    if (n >= 8 && n <= 64 && ckb::small_enabled()) e = ckb::launch_small_c2c(n, inverse, p, (cudaStream_t) stream);
    else if (n <= 8) e = ckb::launch_tiny_c2c(n, inverse, p, (cudaStream_t) stream);
    else        e = inverse ? ckb::launch_c2c_inv_planar(n, p, (cudaStream_t) stream) : ckb::launch_c2c_fwd_planar(n, p, (cudaStream_t) stream);
    if (e != cudaSuccess) { set_error("kernel launch", e); return 0; }
    return 1;
}

int CkFftB200ComplexForwardPlanarBatchAsync(CkFftContext* c, int n, const float* inRe, const float* inIm, float* outRe, float* outIm,
                                            size_t batch, size_t inStride, size_t outStride, void* stream)
{
    return run_planar(c, false, n, inRe, inIm, outRe, outIm, batch, inStride, outStride, stream);
}

int CkFftB200ComplexInversePlanarBatchAsync(CkFftContext* c, int n, const float* inRe, const float* inIm, float* outRe, float* outIm,
                                            size_t batch, size_t inStride, size_t outStride, void* stream)
{
    return run_planar(c, true, n, inRe, inIm, outRe, outIm, batch, inStride, outStride, stream);
}

// ---- audio front end: window, real forward transform and power spectrum in one kernel ----
int CkFftB200RealForwardPowerBatchAsync(CkFftContext* c, int n, const float* in, const float* window, float* power,
                                        size_t batch, size_t inStride, size_t outStride, void* stream)
{
    if (!check_call(c, K_R2C, n, in, power)) return 0;
    if (n < 32 || n > c->tableCount) { set_error("power spectrum: n must be 32 .. 32768"); return 0; }
    if (batch == 0) return 1;
    if (inStride == 0) inStride = (size_t) n;
    if (outStride == 0) outStride = (size_t) n / 2 + 1;
    if (inStride < (size_t) n || outStride < (size_t) n / 2 + 1 || (inStride & 1)) {
        set_error("power spectrum: strides must cover a frame and the input stride must be even");
        return 0;
    }
    if ((((uintptr_t) in | (uintptr_t) window) & 7) || ((uintptr_t) power & 3)) { set_error("power spectrum: misaligned pointer"); return 0; }
    DeviceGuard guard(c->device);
    if (!guard.ok) { set_error("cannot select the context's device"); return 0; }
    ckb::KernelParams p{ (const ckb::cf*) in, (ckb::cf*) power, c->dTable, c->log2Table, (long long) batch,
                         (long long) inStride / 2, (long long) outStride, (const ckb::cf*) window };
    cudaError_t e = ckb::launch_r2c_audio(n / 2, p, (cudaStream_t) stream);
    if (e != cudaSuccess) { set_error("kernel launch", e); return 0; }
    return 1;
}

// ---- local steps of the distributed six-step transform (device pointers, stream-ordered) ----
int CkFftB200PackColumnsAsync(const CkFftComplex* in, CkFftComplex* out, size_t rows, int parts, size_t width, void* stream)
{
    if (!in || !out || in == out || parts <= 0) { set_error("pack: bad arguments"); return 0; }
    if (rows == 0 || width == 0) return 1;
    cudaError_t e = ckb::launch_pack_columns((const ckb::cf*) in, (ckb::cf*) out, (long long) rows, parts, (long long) width,
                                             (cudaStream_t) stream);
    if (e != cudaSuccess) { set_error("pack kernel", e); return 0; }
    return 1;
}

int CkFftB200UnpackTransposeAsync(const CkFftComplex* in, CkFftComplex* out, int parts, size_t rowsPerPart, size_t width,
                                  void* stream)
{
    if (!in || !out || in == out || parts <= 0) { set_error("unpack: bad arguments"); return 0; }
    if (rowsPerPart == 0 || width == 0) return 1;
    cudaError_t e = ckb::launch_unpack_transpose((const ckb::cf*) in, (ckb::cf*) out, parts, (long long) rowsPerPart,
                                                 (long long) width, (cudaStream_t) stream);
    if (e != cudaSuccess) { set_error("unpack kernel", e); return 0; }
    return 1;
}

int CkFftB200TwiddleRowsAsync(CkFftContext* c, int n, CkFftComplex* data, size_t rows, size_t cols, size_t firstRow,
                              int inverse, void* stream)
{
    if (!c || c->magic != kMagic) { set_error("invalid context"); return 0; }
    if (!is_pow2(n) || n > c->maxCount || !c->dTwLo) { set_error("twiddle: n must be a power of two <= nMax of a multi-pass context"); return 0; }
    if (!data) { set_error("twiddle: NULL data"); return 0; }
    if (rows == 0 || cols == 0) return 1;
    if ((unsigned long long) (firstRow + rows - 1) * (unsigned long long) (cols - 1) >= (unsigned long long) n) {
        set_error("twiddle: (firstRow + rows) * cols exceeds n");
        return 0;
    }
    DeviceGuard guard(c->device);
    if (!guard.ok) { set_error("cannot select the context's device"); return 0; }
    cudaError_t e = ckb::launch_twiddle_rows((ckb::cf*) data, (long long) rows, (long long) cols, (long long) firstRow, big_tw(c),
                                             ilog2i(n), inverse != 0, (cudaStream_t) stream);
    if (e != cudaSuccess) { set_error("twiddle kernel", e); return 0; }
    return 1;
}

// ---- fused distributed transform (dist_fused.cu): layout, peer memory, plan ----
int CkFftB200DistGetLayout(long long n, int world, int preferPasses, CkFftB200DistLayout* layout)
{
    if (!ckb::dist_layout(n, world, preferPasses, layout)) { set_error("distributed layout: n must be a power of two in 2^14..2^30 with whole 16-column tiles per rank"); return 0; }
    return 1;
}

int CkFftB200DistDescribe(const CkFftB200DistLayout* layout, int rank, CkFftB200DistPass passes[4])
{
    if (!layout || !passes || rank < 0 || rank >= layout->world) return 0;
    return ckb::dist_describe(*layout, rank, passes);
}

void* CkFftB200PeerAlloc(size_t bytes)
{
    void* p = NULL;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { set_error("peer allocation", e); if (p) cudaFree(p); return NULL; }
    return p;
}

void CkFftB200PeerFree(void* p) { if (p) cudaFree(p); }

int CkFftB200PeerExport(void* p, unsigned char handle[64])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    if (!p || !handle) { set_error("peer export: NULL argument"); return 0; }
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { set_error("cudaIpcGetMemHandle", e); return 0; }
    memcpy(handle, &h, 64);
    return 1;
}

void* CkFftB200PeerOpen(const unsigned char handle[64])
{
    if (!handle) { set_error("peer open: NULL handle"); return NULL; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* p = NULL;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { set_error("cudaIpcOpenMemHandle", e); return NULL; }
    return p;
}

void CkFftB200PeerClose(void* p) { if (p) cudaIpcCloseMemHandle(p); }

struct CkFftB200DistPlan
{
    uint32_t magic;
    CkFftContext* ctx;
    CkFftB200DistLayout layout;
    int rank;
    ckb::DistBuffers bufs;
    unsigned epoch;
    bool profiling;
    ckb::DistMarks marks;
};

CkFftB200DistPlan* CkFftB200DistPlanCreate(CkFftContext* c, long long n, int rank, int world, int preferPasses,
                                           void* const* work, void* const* mid, void* const* out, void* const* flags,
                                           void* const* in)
{
    if (!c || c->magic != kMagic) { set_error("invalid context"); return NULL; }
    if (!c->fwdExpTable && !c->invExpTable) { set_error("invalid context"); return NULL; }
    if (n > c->maxCount || !c->dTwLo) { set_error("distributed plan: the context must be created with nMax >= n"); return NULL; }
    CkFftB200DistLayout lay;
    if (!CkFftB200DistGetLayout(n, world, preferPasses, &lay)) return NULL;
    if (rank < 0 || rank >= world || !work || !mid || !out || !flags) { set_error("distributed plan: bad arguments"); return NULL; }
    CkFftB200DistPlan* p = (CkFftB200DistPlan*) calloc(1, sizeof(CkFftB200DistPlan));
    if (!p) { set_error("out of host memory"); return NULL; }
    for (int q = 0; q < world; ++q) {
        if (!work[q] || !mid[q] || !out[q] || !flags[q] || (((uintptr_t) work[q] | (uintptr_t) mid[q] | (uintptr_t) out[q]) & 127)) {
            set_error("distributed plan: every rank's buffers must be non-NULL and 128-byte aligned");
            free(p);
            return NULL;
        }
        p->bufs.buf[0][q] = (ckb::cf*) work[q];
        p->bufs.buf[1][q] = (ckb::cf*) mid[q];
        p->bufs.buf[2][q] = (ckb::cf*) out[q];
        p->bufs.flags[q] = (unsigned*) flags[q];
    }
    p->magic = kMagic;
    p->ctx = c;
    p->layout = lay;
    p->rank = rank;
    // The flag block is caller-owned and may have served an earlier plan: a barrier passes when the slot has REACHED
    // its epoch, so a new plan must continue from the value the block holds (every completed barrier leaves the same
    // epoch in every slot of every rank), not from 0 -- otherwise its first barriers would all pass at once on the
    // stale values and the passes would race.  The error word of an earlier time-out is cleared.
    {
        DeviceGuard guard(c->device);
        unsigned words[CKB_MAX_PEERS + 1] = {0};
        cudaError_t e = cudaDeviceSynchronize();
        if (e == cudaSuccess) e = cudaMemcpy(words, flags[rank], sizeof(words), cudaMemcpyDeviceToHost);
        unsigned zero = 0;
        if (e == cudaSuccess) e = cudaMemcpy((unsigned*) flags[rank] + CKB_MAX_PEERS, &zero, sizeof(zero), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { set_error("distributed plan: reading the flag block", e); free(p); return NULL; }
        unsigned start = words[0];
        for (int q = 1; q < world; ++q)
            if ((int) (words[q] - start) > 0) start = words[q];
        for (int q = 0; q < world; ++q)
            if (words[q] != start) {
                set_error("distributed plan: the flag block holds an unfinished barrier of an earlier plan (zero it on every rank, then rendezvous)");
                free(p);
                return NULL;
            }
        p->epoch = start;
    }
    if (in) {
        // pull mode: tensor maps over every rank's input array, built once
        DeviceGuard guard(c->device);
        for (int q = 0; q < world; ++q) {
            if (!in[q] || ((uintptr_t) in[q] & 127)) { set_error("distributed plan: input arrays must be non-NULL and 128-byte aligned"); free(p); return NULL; }
            p->bufs.in[q] = (ckb::cf*) in[q];
        }
        p->layout.pull = 1;
        void* dmaps = NULL;
        cudaError_t e = cudaMalloc(&dmaps, 128 * CKB_MAX_PEERS);
        if (e == cudaSuccess) e = ckb::dist_make_pull_maps(p->layout, rank, p->bufs.in, (::CUtensorMap_st*) dmaps, &p->bufs.pull_box_rows);
        if (e != cudaSuccess) { set_error("distributed plan: tensor maps of the input arrays", e); if (dmaps) cudaFree(dmaps); free(p); return NULL; }
        p->bufs.pull_maps = (const ::CUtensorMap_st*) dmaps;
    }
    return p;
}

int CkFftB200DistExecAsync(CkFftB200DistPlan* p, const CkFftComplex* input, int inverse, void* stream)
{
    if (!p || p->magic != kMagic || !p->ctx || p->ctx->magic != kMagic) { set_error("invalid distributed plan"); return 0; }
    const _CkFftContext* c = p->ctx;
    if (inverse ? !c->invExpTable : !c->fwdExpTable) { set_error("context was not created for this direction"); return 0; }
    if (!input || ((uintptr_t) input & 15)) { set_error("distributed exec: input must be a 16-byte aligned device pointer"); return 0; }
    for (int k = 0; k < 3; ++k)
        if ((const void*) input == (const void*) p->bufs.buf[k][p->rank]) { set_error("distributed exec: input must not be one of the plan's buffers"); return 0; }
    DeviceGuard guard(c->device);
    if (!guard.ok) { set_error("cannot select the context's device"); return 0; }
    cudaError_t e = ckb::dist_exec(p->layout, p->rank, p->bufs, &p->epoch, (const ckb::cf*) input, inverse != 0, c->dTable,
                                   c->log2Table, big_tw(c), (cudaStream_t) stream, p->profiling ? &p->marks : nullptr);
    if (e != cudaSuccess) { set_error("distributed exec", e); return 0; }
    return 1;
}

int CkFftB200DistPlanSetProfiling(CkFftB200DistPlan* p, int on)
{
    if (!p || p->magic != kMagic) { set_error("invalid distributed plan"); return 0; }
    DeviceGuard guard(p->ctx->device);
    if (on && !p->profiling) {
        for (int i = 0; i < ckb::DistMarks::kMax; ++i)
            if (cudaEventCreate(&p->marks.ev[i]) != cudaSuccess) { set_error("event creation"); return 0; }
        p->marks.count = 0;
        p->profiling = true;
    } else if (!on && p->profiling) {
        cudaDeviceSynchronize();
        for (int i = 0; i < ckb::DistMarks::kMax; ++i) cudaEventDestroy(p->marks.ev[i]);
        p->profiling = false;
    }
    return 1;
}

int CkFftB200DistPlanPhases(CkFftB200DistPlan* p, float* ms, char* names, size_t namesBytes)
{
    if (!p || p->magic != kMagic || !p->profiling || !ms || !names || namesBytes == 0) { set_error("phases: profiling is off or bad arguments"); return 0; }
    DeviceGuard guard(p->ctx->device);
    if (cudaDeviceSynchronize() != cudaSuccess) return 0;
    names[0] = 0;
    int n = 0;
    for (int i = 1; i < p->marks.count; ++i) {
        if (cudaEventElapsedTime(&ms[n], p->marks.ev[i - 1], p->marks.ev[i]) != cudaSuccess) { cudaGetLastError(); return 0; }
        const size_t used = strlen(names);
        snprintf(names + used, namesBytes - used, "%s%s", n ? "," : "", p->marks.name[i]);
        ++n;
    }
    return n;
}

int CkFftB200DistPlanStatus(CkFftB200DistPlan* p)
{
    if (!p || p->magic != kMagic) { set_error("invalid distributed plan"); return 0; }
    DeviceGuard guard(p->ctx->device);
    unsigned err = 0;
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(&err, p->bufs.flags[p->rank] + CKB_MAX_PEERS, sizeof(err), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("distributed status", e); return 0; }
    if (err) { set_error("distributed transform: a peer did not reach a barrier in time"); return 0; }
    return 1;
}

void CkFftB200DistPlanDestroy(CkFftB200DistPlan* p)
{
    if (!p || p->magic != kMagic) return;
    CkFftB200DistPlanSetProfiling(p, 0);
    if (p->bufs.pull_maps) { DeviceGuard guard(p->ctx->device); cudaFree((void*) p->bufs.pull_maps); }
    p->magic = 0;
    free(p);
}

}  // extern "C"
