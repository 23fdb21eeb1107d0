// harness_compat.cpp -- lets the reference's UNMODIFIED test harness (src/test/test.cpp) link against
// libckfft_b200.so instead of the reference library.  TEST INFRASTRUCTURE ONLY (built by oracle/build.py
// into oracle/_ref/ckfft_test_b200, run by tests/test_harness_gpu.py).
//
// The harness reaches behind the C ABI in one place: it includes the private header ckfft/context.h and
// calls the static member CkFftContext::isNeonSupported() (src/test/test.cpp:245-249).  That symbol is not
// part of the public ABI, so the product library does not export it; it is supplied here, compiled against
// the reference's own header.  There is one code path on the GPU: the answer is "no NEON".
#include "ckfft/context.h"

bool _CkFftContext::isNeonSupported() { return false; }
