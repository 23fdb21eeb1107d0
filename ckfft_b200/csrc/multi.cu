// multi.cu -- the batched scheduler: ONE host call spreads a batch of independent transforms over several GPUs.
//
// north_star (5) / SURVEY.md 8e "batched transforms": the transforms of a batch share no state (the reference says
// so itself: "the context does not contain state, so contexts can be used simultaneously on different threads",
// inc/ckfft/ckfft.h:39-41), so the batch is cut into contiguous shards, one per device, and every device runs the
// library's ordinary host-buffer pipeline (chunked H2D -> kernel -> D2H, three chunks in flight, api.cu run_host) on
// its own streams, driven by its own host thread.  No collective, no peer traffic: the only synchronisation is the
// join at the end of the call.
//
//   CkFftB200MultiInit      one context replica (twiddle tables) per device + one persistent worker thread per device
//   CkFft*BatchMulti        shard -> post to the workers -> join; returns 1 only if every shard returned 1
//   CkFftB200ShardRange     the shard arithmetic itself (pure host code; bench.py and the tests use the same function)
//
// Pageable host memory: every worker stages its chunks through pinned slots with its share of the host's copy threads
// (api.cu, run_host_pageable; set_host_sharers below).  With CKFFT_B200_PIN=1 instead a call that
// moves at least kPinThreshold bytes page-locks the caller's arrays for its duration (cudaHostRegister, portable across
// the devices); off by default because registration costs as much as it saves (api.cu, ScopedPin).  Arrays that are
// already pinned are left alone.
#include <stdarg.h>
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

#include <cuda_runtime.h>

#include "ckfft/ckfft.h"
#include "ckfft/ckfft_b200.h"

namespace ckb {                                             // api.cu
void set_last_error(const char* text);
int run_host_shared(CkFftContext* c, int kind, int n, const void* in, void* out, size_t batch, std::atomic<size_t>* cursor);
void set_host_sharers(int n);
}

namespace {

constexpr uint32_t kMultiMagic = 0x434b4d47u;              // "CKMG"
constexpr size_t kPinThreshold = size_t(64) << 20;

// failures are reported through the library's one error channel, CkFftB200LastError() (api.cu)
void fail(const char* fmt, ...)
{
    char buf[256];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    ckb::set_last_error(buf);
}

enum JobKind { JOB_NONE = 0, JOB_C2C_FWD, JOB_C2C_INV, JOB_R2C, JOB_C2R };

struct Job {
    int kind = JOB_NONE;
    int n = 0;
    const void* in = nullptr;
    void* out = nullptr;
    size_t batch = 0;
    std::atomic<size_t>* cursor = nullptr;     // dynamic schedule: the whole batch + a shared chunk counter; static: this worker's shard
    int parts = 1;                             // workers of this call
};

struct Worker {
    int device = -1;
    CkFftContext* ctx = nullptr;
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    Job job;
    bool has_job = false, done = false, quit = false, ready = false;
    int result = 0;
    char error[256] = "";
};

void worker_main(Worker* w, int nMax, CkFftDirection dir)
{
    // the context is created on the worker's own thread: its device stays current for the thread's lifetime, and the
    // per-thread staging buffers / streams of the host-buffer path (api.cu) belong to this device only
    int ok = cudaSetDevice(w->device) == cudaSuccess;
    if (ok) {
        w->ctx = CkFftInit(nMax, dir, NULL, NULL);
        ok = w->ctx != NULL;
        if (!ok) snprintf(w->error, sizeof(w->error), "device %d: %s", w->device, CkFftB200LastError());
    } else {
        snprintf(w->error, sizeof(w->error), "device %d: cannot be selected", w->device);
        cudaGetLastError();
    }
    {
        std::lock_guard<std::mutex> lk(w->m);
        w->ready = true;
        w->result = ok;
    }
    w->cv.notify_all();
    if (!ok) return;
    for (;;) {
        Job j;
        {
            std::unique_lock<std::mutex> lk(w->m);
            w->cv.wait(lk, [&] { return w->has_job || w->quit; });
            if (w->quit) break;
            j = w->job;
            w->has_job = false;
        }
        int r = 1;
        ckb::set_host_sharers(j.parts);                    // pageable arrays: the workers of one call divide the host's copy threads
        if (j.batch > 0 && j.cursor) {
            r = ckb::run_host_shared(w->ctx, j.kind - JOB_C2C_FWD, j.n, j.in, j.out, j.batch, j.cursor);
        } else if (j.batch > 0) {
            switch (j.kind) {
                case JOB_C2C_FWD: r = CkFftComplexForwardBatch(w->ctx, j.n, (const CkFftComplex*) j.in, (CkFftComplex*) j.out, j.batch); break;
                case JOB_C2C_INV: r = CkFftComplexInverseBatch(w->ctx, j.n, (const CkFftComplex*) j.in, (CkFftComplex*) j.out, j.batch); break;
                case JOB_R2C:     r = CkFftRealForwardBatch(w->ctx, j.n, (const float*) j.in, (CkFftComplex*) j.out, j.batch); break;
                case JOB_C2R:     r = CkFftRealInverseBatch(w->ctx, j.n, (const CkFftComplex*) j.in, (float*) j.out, NULL, j.batch); break;
                default:          r = 0;
            }
        }
        {
            std::lock_guard<std::mutex> lk(w->m);
            if (!r) snprintf(w->error, sizeof(w->error), "device %d: %s", w->device, CkFftB200LastError());
            w->result = r;
            w->done = true;
        }
        w->cv.notify_all();
    }
    CkFftShutdown(w->ctx);
    w->ctx = nullptr;
}

bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

// is this host range already page-locked (or not host memory at all)?
bool needs_pinning(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

bool pin_enabled()
{
    const char* e = getenv("CKFFT_B200_PIN");          // same opt-in switch as the single-device path (api.cu)
    return e && e[0] == '1';
}

}  // namespace

struct CkFftB200Multi {
    uint32_t magic;
    int nMax;
    CkFftDirection direction;
    std::vector<Worker*> workers;
    std::mutex call;                 // one batched call at a time per handle (the workers are the shared state)
};

extern "C" {

int CkFftB200ShardRange(size_t batch, int part, int parts, size_t* first, size_t* count)
{
    if (parts < 1 || part < 0 || part >= parts || !first || !count) return 0;
    const size_t base = batch / (size_t) parts, extra = batch % (size_t) parts;
    const size_t p = (size_t) part;
    *first = p * base + (p < extra ? p : extra);
    *count = base + (p < extra ? 1 : 0);
    return 1;
}

CkFftB200Multi* CkFftB200MultiInit(int nMax, CkFftDirection direction, const int* devices, int nDevices)
{
    if (!is_pow2(nMax)) { fail("nMax must be a positive power of two"); return NULL; }
    if (direction != kCkFftDirection_Forward && direction != kCkFftDirection_Inverse && direction != kCkFftDirection_Both) {
        fail("invalid direction");
        return NULL;
    }
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < 1) {
        cudaGetLastError();
        fail("no usable CUDA device (this library has no CPU path)");
        return NULL;
    }
    std::vector<int> devs;
    if (devices == NULL) {
        const int cnt = nDevices > 0 && nDevices < visible ? nDevices : visible;     // NULL: the first nDevices (<= 0: all) devices
        for (int i = 0; i < cnt; ++i) devs.push_back(i);
    } else {
        if (nDevices < 1 || nDevices > 64) { fail("nDevices must be 1 .. 64"); return NULL; }
        for (int i = 0; i < nDevices; ++i) {
            if (devices[i] < 0 || devices[i] >= visible) {
                fail("device %d does not exist (%d visible)", devices[i], visible);
                return NULL;
            }
            devs.push_back(devices[i]);          // a device may be listed more than once: each entry gets its own replica + thread
        }
    }
    CkFftB200Multi* m = new (std::nothrow) CkFftB200Multi();
    if (!m) { fail("out of host memory"); return NULL; }
    m->magic = kMultiMagic;
    m->nMax = nMax;
    m->direction = direction;
    bool ok = true;
    char why[256] = "";
    for (int d : devs) {
        Worker* w = new (std::nothrow) Worker();
        if (!w) { ok = false; break; }
        w->device = d;
        m->workers.push_back(w);
        w->th = std::thread(worker_main, w, nMax, direction);
    }
    for (Worker* w : m->workers) {
        std::unique_lock<std::mutex> lk(w->m);
        w->cv.wait(lk, [&] { return w->ready; });
        if (!w->result) {
            if (ok) snprintf(why, sizeof(why), "%s", w->error);
            ok = false;
        }
    }
    if (!ok) {
        char keep[256];
        snprintf(keep, sizeof(keep), "%s", why[0] ? why : "out of host memory");
        CkFftB200MultiShutdown(m);
        fail("%s", keep);
        return NULL;
    }
    return m;
}

void CkFftB200MultiShutdown(CkFftB200Multi* m)
{
    if (!m || m->magic != kMultiMagic) return;
    for (Worker* w : m->workers) {
        {
            std::lock_guard<std::mutex> lk(w->m);
            w->quit = true;
        }
        w->cv.notify_all();
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    m->workers.clear();
    m->magic = 0;
    delete m;
}

int CkFftB200MultiDeviceCount(const CkFftB200Multi* m) { return (m && m->magic == kMultiMagic) ? (int) m->workers.size() : 0; }

int CkFftB200MultiDevice(const CkFftB200Multi* m, int index)
{
    if (!m || m->magic != kMultiMagic || index < 0 || index >= (int) m->workers.size()) return -1;
    return m->workers[index]->device;
}

CkFftContext* CkFftB200MultiContext(const CkFftB200Multi* m, int index)
{
    if (!m || m->magic != kMultiMagic || index < 0 || index >= (int) m->workers.size()) return NULL;
    return m->workers[index]->ctx;
}

static int run_multi(CkFftB200Multi* m, int kind, int n, const void* in, void* out, size_t batch)
{
    if (!m || m->magic != kMultiMagic) { fail("invalid multi-device handle"); return 0; }
    // the classic checks (src/ckfft/ckfft.cpp:36-114), made once up front so that a bad call fails before any shard runs
    const bool inverse = kind == JOB_C2C_INV || kind == JOB_C2R;
    if (!(m->direction & (inverse ? kCkFftDirection_Inverse : kCkFftDirection_Forward))) {
        fail("handle was not created for this direction");
        return 0;
    }
    if (!is_pow2(n) || n > m->nMax) { fail("n must be a power of two <= nMax"); return 0; }
    if (!in || !out || in == out) { fail("input/output must be distinct non-NULL buffers"); return 0; }
    if (batch == 0) return 1;
    for (const void* p : { in, (const void*) out }) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) == cudaSuccess) {
            if (a.type == cudaMemoryTypeDevice) {
                fail("multi-device calls take HOST arrays (device arrays belong to one GPU: use the per-device contexts)");
                return 0;
            }
        } else cudaGetLastError();
    }
    const size_t in_elems = kind == JOB_C2R ? (size_t) n / 2 + 1 : (size_t) n, out_elems = kind == JOB_R2C ? (size_t) n / 2 + 1 : (size_t) n;
    const size_t ib = in_elems * (kind == JOB_R2C ? 4 : 8), ob = out_elems * (kind == JOB_C2R ? 4 : 8);

    std::lock_guard<std::mutex> call_lock(m->call);
    bool pinned_in = false, pinned_out = false;
    if (pin_enabled() && (ib + ob) * batch >= kPinThreshold) {
        if (needs_pinning(in)) {
            pinned_in = cudaHostRegister((void*) in, ib * batch, cudaHostRegisterPortable | cudaHostRegisterReadOnly) == cudaSuccess ||
                        (cudaGetLastError(), cudaHostRegister((void*) in, ib * batch, cudaHostRegisterPortable) == cudaSuccess);
        }
        if (needs_pinning(out)) pinned_out = cudaHostRegister(out, ob * batch, cudaHostRegisterPortable) == cudaSuccess;
        cudaGetLastError();          // a refused registration only means the copies stay synchronous
    }
    // Schedule.  Dynamic (default for batches of at least 256 MiB): every device works on the whole batch and draws 32 MiB
    // chunks from a shared counter, so the devices finish together even when their host links are not equally fast --
    // on the 8-GPU box four of the GPUs get ~22 GB/s and one ~90 GB/s when all copy at once, and with equal static shards
    // the call took as long as the slowest link needed (133 GB/s in total; plain concurrent copies: 212).  Static
    // (CKFFT_B200_MULTI_STATIC=1, and small batches): contiguous shards as CkFftB200ShardRange cuts them.
    const int parts = (int) m->workers.size();
    const char* st_env = getenv("CKFFT_B200_MULTI_STATIC");
    const bool dynamic = parts > 1 && (ib + ob) * batch >= (size_t(256) << 20) && !(st_env && st_env[0] == '1');
    std::atomic<size_t> cursor{0};
    for (int i = 0; i < parts; ++i) {
        Worker* w = m->workers[i];
        size_t first = 0, count = 0;
        CkFftB200ShardRange(batch, i, parts, &first, &count);
        {
            std::lock_guard<std::mutex> lk(w->m);
            w->job.kind = kind;
            w->job.n = n;
            w->job.in = dynamic ? in : (const void*) ((const char*) in + first * ib);
            w->job.out = dynamic ? out : (void*) ((char*) out + first * ob);
            w->job.batch = dynamic ? batch : count;
            w->job.cursor = dynamic ? &cursor : nullptr;
            w->job.parts = parts;
            w->done = false;
            w->has_job = true;
        }
        w->cv.notify_all();
    }
    int ok = 1;
    for (Worker* w : m->workers) {
        std::unique_lock<std::mutex> lk(w->m);
        w->cv.wait(lk, [&] { return w->done; });
        if (!w->result) {
            if (ok) fail("%s", w->error);
            ok = 0;
        }
    }
    if (pinned_in) cudaHostUnregister((void*) in);
    if (pinned_out) cudaHostUnregister(out);
    cudaGetLastError();
    return ok;
}

int CkFftComplexForwardBatchMulti(CkFftB200Multi* m, int n, const CkFftComplex* in, CkFftComplex* out, size_t batch)
{
    return run_multi(m, JOB_C2C_FWD, n, in, out, batch);
}

int CkFftComplexInverseBatchMulti(CkFftB200Multi* m, int n, const CkFftComplex* in, CkFftComplex* out, size_t batch)
{
    return run_multi(m, JOB_C2C_INV, n, in, out, batch);
}

int CkFftRealForwardBatchMulti(CkFftB200Multi* m, int n, const float* in, CkFftComplex* out, size_t batch)
{
    return run_multi(m, JOB_R2C, n, in, out, batch);
}

int CkFftRealInverseBatchMulti(CkFftB200Multi* m, int n, const CkFftComplex* in, float* out, size_t batch)
{
    return run_multi(m, JOB_C2R, n, in, out, batch);
}

}  // extern "C"
