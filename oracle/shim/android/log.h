/* Stub of <android/log.h> for src/ckfft/debug.cpp:6-7,25: log to stdout. Test infrastructure only. */
#pragma once
#include <stdarg.h>
#include <stdio.h>
#define ANDROID_LOG_INFO 4
static inline int __android_log_vprint(int prio, const char* tag, const char* fmt, va_list ap)
{
    (void) prio; (void) tag;
    return vprintf(fmt, ap);
}
