// host_copy.cpp -- the host-side copy of the pageable staging pipeline (api.cu, run_host_pageable / CopyTeam).
//
// A drop-in caller of the reference hands over malloc'ed arrays (src/test/test.cpp:254-301 does).  The library stages them
// through pinned slots with teams of host threads, and that pipeline is bound by host-memory traffic: two host copies and two
// DMA passes per byte of payload.  A plain memcpy of a thread's 2 MiB piece stays under glibc's non-temporal threshold, so
// every destination line is first READ for ownership -- three units of traffic per unit copied.  The destination (a pinned slot
// the copy engine reads next, or the caller's output array) is not touched again by the copying core, so it is written with
// non-temporal stores instead: two units per unit copied, and nothing of it evicts the caller's cache.
// Widest vector unit the CPU has, chosen once at run time (plain host C++, compiled by the host compiler only).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace ckb {

#if defined(__x86_64__)
namespace {

// each kernel copies whole 256-byte blocks to a 64-byte aligned destination and returns the bytes it copied
size_t nt_sse2(char* dst, const char* src, size_t bytes)
{
    const size_t blocks = bytes / 256;
    for (size_t i = 0; i < blocks; ++i) {
        for (int h = 0; h < 4; ++h) {
            const __m128i a = _mm_loadu_si128((const __m128i*) src + 0), b = _mm_loadu_si128((const __m128i*) src + 1);
            const __m128i c = _mm_loadu_si128((const __m128i*) src + 2), d = _mm_loadu_si128((const __m128i*) src + 3);
            _mm_stream_si128((__m128i*) dst + 0, a);
            _mm_stream_si128((__m128i*) dst + 1, b);
            _mm_stream_si128((__m128i*) dst + 2, c);
            _mm_stream_si128((__m128i*) dst + 3, d);
            src += 64; dst += 64;
        }
    }
    return blocks * 256;
}

__attribute__((target("avx2"))) size_t nt_avx2(char* dst, const char* src, size_t bytes)
{
    const size_t blocks = bytes / 256;
    for (size_t i = 0; i < blocks; ++i) {
        for (int h = 0; h < 2; ++h) {
            const __m256i a = _mm256_loadu_si256((const __m256i*) src + 0), b = _mm256_loadu_si256((const __m256i*) src + 1);
            const __m256i c = _mm256_loadu_si256((const __m256i*) src + 2), d = _mm256_loadu_si256((const __m256i*) src + 3);
            _mm256_stream_si256((__m256i*) dst + 0, a);
            _mm256_stream_si256((__m256i*) dst + 1, b);
            _mm256_stream_si256((__m256i*) dst + 2, c);
            _mm256_stream_si256((__m256i*) dst + 3, d);
            src += 128; dst += 128;
        }
    }
    return blocks * 256;
}

__attribute__((target("avx512f"))) size_t nt_avx512(char* dst, const char* src, size_t bytes)
{
    const size_t blocks = bytes / 256;
    for (size_t i = 0; i < blocks; ++i) {
        const __m512i a = _mm512_loadu_si512(src), b = _mm512_loadu_si512(src + 64);
        const __m512i c = _mm512_loadu_si512(src + 128), d = _mm512_loadu_si512(src + 192);
        _mm512_stream_si512((__m512i*) dst, a);
        _mm512_stream_si512((__m512i*) (dst + 64), b);
        _mm512_stream_si512((__m512i*) (dst + 128), c);
        _mm512_stream_si512((__m512i*) (dst + 192), d);
        src += 256; dst += 256;
    }
    return blocks * 256;
}

typedef size_t (*NtKernel)(char*, const char*, size_t);

// level: 1 = widest the CPU supports, 2 = SSE2, 3 = AVX2, 4 = AVX-512 (A/B measurements; an unsupported level falls back)
NtKernel pick(int level)
{
    __builtin_cpu_init();
    const bool avx2 = __builtin_cpu_supports("avx2"), avx512 = __builtin_cpu_supports("avx512f");
    if ((level == 1 || level == 4) && avx512) return nt_avx512;
    if ((level == 1 || level == 3 || level == 4) && avx2) return nt_avx2;
    return nt_sse2;
}

}  // namespace
#endif

// CKFFT_B200_NT_COPY: 0 = memcpy, 1 (default) = non-temporal stores with the widest vectors of this CPU, 2 / 3 / 4 = SSE2 /
// AVX2 / AVX-512 forced.  Read per call (a call copies megabytes).
int stream_copy_level()
{
    const char* e = getenv("CKFFT_B200_NT_COPY");
    const int v = e && *e ? atoi(e) : 1;
    return v < 0 || v > 4 ? 1 : v;
}

const char* stream_copy_name(int level)
{
#if defined(__x86_64__)
    if (level == 0) return "memcpy";
    const NtKernel k = pick(level);
    return k == nt_avx512 ? "non-temporal avx512" : k == nt_avx2 ? "non-temporal avx2" : "non-temporal sse2";
#else
    (void) level;
    return "memcpy";
#endif
}

// copy `bytes` from src to dst (no overlap); on return the data is globally visible (the caller may start a DMA on it)
void stream_copy(void* dst_, const void* src_, size_t bytes, int level)
{
    char* dst = (char*) dst_;
    const char* src = (const char*) src_;
#if defined(__x86_64__)
    if (level > 0 && bytes >= 4096) {
        static const NtKernel auto_kernel = pick(1);
        const NtKernel k = level == 1 ? auto_kernel : pick(level);
        const size_t head = (size_t) (0 - (uintptr_t) dst) & 63;          // up to the destination's next 64-byte boundary
        if (head) { memcpy(dst, src, head); dst += head; src += head; bytes -= head; }
        const size_t done = k(dst, src, bytes);
        dst += done; src += done; bytes -= done;
        _mm_sfence();                                                      // the write-combining buffers drain before anyone is told
    }
#else
    (void) level;
#endif
    if (bytes) memcpy(dst, src, bytes);
}

}  // namespace ckb
