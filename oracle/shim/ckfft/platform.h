/* Shim that shadows the reference's src/ckfft/platform.h when the UNMODIFIED
 * reference sources are compiled on Linux for the oracle (oracle/build.py).
 * The real header #errors on Linux (src/ckfft/platform.h:27-29).  Declaring the
 * Android platform selects clock_gettime in the harness timer
 * (src/test/timer.h:55-60) and makes isNeonSupported() consult the stub
 * cpu-features.h next to this file, which reports "no NEON" -> scalar path.
 * Test infrastructure only. */
#pragma once
#define CKFFT_PLATFORM_ANDROID 1
namespace ckfft
{
typedef unsigned char uchar;   typedef unsigned short ushort;
typedef unsigned int uint;     typedef unsigned long ulong;
typedef signed char int8;      typedef unsigned char uint8;
typedef signed short int16;    typedef unsigned short uint16;
typedef signed int int32;      typedef unsigned int uint32;
typedef signed long long int64; typedef unsigned long long uint64;
}
