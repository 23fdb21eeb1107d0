/* Stub of the Android NDK header used by src/ckfft/context.cpp:12-14,129:
 * report no CPU features so the reference takes its scalar path. Test infrastructure only. */
#pragma once
#define ANDROID_CPU_ARM_FEATURE_NEON 4
static inline unsigned long long android_getCpuFeatures(void) { return 0; }
