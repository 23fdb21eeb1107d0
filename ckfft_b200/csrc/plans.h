// plans.h -- the single-pass plan table: one row per complex transform length M.
//   X(M, E, R0, R1, R2, G, MINB)
//     E    complex values per thread (registers)      T = M/E threads per transform
//     R*   radices of the 2 or 3 Stockham stages (R2 = 1: two stages)
//     G    transforms per CTA                         CTA = G*T threads
//     MINB __launch_bounds__ min CTAs/SM (caps registers so that many CTAs stay resident)
// The launch planner (launch.cu) and CkFftB200GetPlan (api.cu) both expand this table, so the
// host-visible plan is by construction the kernel that runs.
#pragma once

#define CKB_SINGLE_PASS_PLANS(X) \
    X(16,     4,  4,  4,  1, 32, 8) \
    X(32,     8,  8,  4,  1, 32, 8) \
    X(64,     8,  8,  8,  1, 16, 8) \
    X(128,   16, 16,  8,  1, 16, 4) \
    X(256,   16, 16, 16,  1,  8, 4) \
    X(512,   32, 32, 16,  1,  8, 4) \
    X(1024,  32, 32, 32,  1,  4, 4) \
    X(2048,  32, 32, 32,  2,  4, 2) \
    X(4096,  16, 16, 16, 16,  2, 2) \
    X(8192,  32, 32, 16, 16,  1, 2) \
    X(16384, 32, 32, 32, 16,  1, 1)

#define CKB_MAX_SINGLE_PASS 16384   /* largest complex length done in one launch */
#define CKB_MAX_TABLE 32768         /* device twiddle table W_Nt^k covers real n up to this in one pass */
