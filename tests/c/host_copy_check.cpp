// CPU check of ckfft_b200/csrc/host_copy.cpp (the staging copy of the pageable host path): every level, awkward sizes and
// alignments, guard bytes around the destination.  Built and run by tests/test_host_copy.py; prints "ok <levels...>".
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace ckb {
int stream_copy_level();
const char* stream_copy_name(int level);
void stream_copy(void* dst, const void* src, size_t bytes, int level);
}

int main(int argc, char** argv)
{
    if (argc > 1) { printf("ok default=%d\n", ckb::stream_copy_level()); return 0; }      // only the environment switch
    const size_t span = (size_t(1) << 21) + 8192;
    std::vector<unsigned char> a(span), b(span), c(span);
    for (size_t i = 0; i < span; ++i) a[i] = (unsigned char) ((i * 2654435761u) >> 13);
    const size_t sizes[] = { 0, 1, 63, 64, 255, 256, 4095, 4096, 4097, 4096 + 255, 65536 + 7, 1000003, size_t(1) << 21 };
    int bad = 0;
    for (int level = 0; level <= 4; ++level)
        for (int so = 0; so < 70; so += 23)
            for (int d0 = 0; d0 < 70; d0 += 5)
                for (size_t n : sizes) {
                    memset(b.data(), 0x55, span);
                    memset(c.data(), 0x55, span);
                    ckb::stream_copy(b.data() + d0, a.data() + so, n, level);
                    memcpy(c.data() + d0, a.data() + so, n);
                    if (memcmp(b.data(), c.data(), span)) {
                        ++bad;
                        if (bad < 10) printf("mismatch level=%d src+%d dst+%d n=%zu\n", level, so, d0, n);
                    }
                }
    if (bad) { printf("FAILED %d\n", bad); return 1; }
    printf("ok default=%d", ckb::stream_copy_level());
    for (int level = 0; level <= 4; ++level) printf(" | %d: %s", level, ckb::stream_copy_name(level));
    printf("\n");
    return 0;
}
