// plans.h -- the single-pass plan table: one row per complex transform length M.
//   X(M, E, R0, R1, R2, G, MINB, TWR)
//     E    complex values per thread (registers)      T = M/E threads per transform
//     R*   radices of the 2 or 3 Stockham stages (R2 = 1: two stages)
//     G    transforms per CTA                         CTA = G*T threads
//     MINB __launch_bounds__ min CTAs/SM (caps registers so that many CTAs stay resident)
//     TWR  1: stage-1 twiddles live in registers for the whole kernel (two-stage plans with R1 == E)
// The launch planner (launch.cu) and CkFftB200GetPlan (api.cu) both expand this table, so the
// host-visible plan is by construction the kernel that runs.
#pragma once

#define CKB_SINGLE_PASS_PLANS(X) \
    X(16,     4,  4,  4,  1, 32, 8, 1) \
    X(32,     8,  8,  4,  1, 32, 8, 0) \
    X(64,     8,  8,  8,  1, 16, 8, 1) \
    X(128,   16, 16,  8,  1, 16, 4, 0) \
    X(256,   16, 16, 16,  1,  8, 4, 1) \
    X(512,   32, 32, 16,  1,  8, 4, 0) \
    X(1024,  32, 32, 32,  1,  4, 4, 1) \
    X(2048,  32, 32, 32,  2,  4, 2, 0) \
    X(4096,  32, 32, 32,  4,  2, 2, 0) \
    X(8192,  32, 32, 32,  8,  1, 2, 0) \
    X(16384, 32, 32, 32, 16,  1, 1, 1)

// Bulk-prefetch variants (Cfg::PF): each group's next transform is fetched by cp.async.bulk into a dense
// staging buffer while the current one is computed.  Used for complex and real-forward transforms whose
// rows are 16-byte aligned (CKFFT_B200_PREFETCH=0 disables them, for A/B measurements).
#define CKB_PREFETCH_PLANS(X) \
    X(1024,  32, 32, 32,  1,  4, 3, 1)

// Split prefetch (Cfg::PF == PF_SPLIT): the lower half of a group's NEXT row is fetched into a half-row staging buffer as
// soon as stage 0 has gathered the current one (in flight for the whole transform), the upper half into the exchange buffer
// once the last stage has gathered.  Introduced for 16384 points (one transform fills the register file of an SM, a whole
// second row does not fit next to the exchange buffer: 0.64 -> 0.70 of the copy peak; stage-1 twiddles from register bases,
// TWR, on top: .70 -> .77).
// Round 2: ONE CTA per SM with several groups (G x T = 384 .. 512 threads, MINB = 1) sharing one copy of the twiddle LUTs
// beats two CTAs with in-place prefetch, whose load is only in flight during the short last stage (ncu: 26-39 % of the
// stall samples on the mbarrier wait).  Measured (tools/exp_build.sh, fraction of the copy peak, in-place -> split):
//   C2C 4096 (G = 3, LUT)  .88-.95 -> .995     C2C 8192 (G = 2, TWR)  .88 -> .97      C2C 2048 (G = 6) .96 -> .88 (stays in place)
//   R2C M = 2048 (G = 6, LUT) .84-.90 -> .945 (config 3: .93 -> .96)      R2C M = 4096 (G = 3, TWR) .88 -> .93
#ifndef CKB_PLANAR_SPLIT_ALL      /* development A/B switch */
#define CKB_PLANAR_SPLIT_ALL 1
#endif
#if CKB_PLANAR_SPLIT_ALL          /* split-complex rows: the plane of real parts and the plane of imaginary parts are the two halves */
#define CKB_SPLIT_PREFETCH_PLANS_PLANAR(X) \
    X(4096,  32, 32, 32,  4,  3, 1, 0) \
    X(8192,  32, 32, 32,  8,  2, 1, 1) \
    X(16384, 32, 32, 32, 16,  1, 1, 1)
#else
#define CKB_SPLIT_PREFETCH_PLANS_PLANAR(X) \
    X(16384, 32, 32, 32, 16,  1, 1, 1)
#endif
#define CKB_SPLIT_PREFETCH_PLANS_C2C(X) \
    X(4096,  32, 32, 32,  4,  3, 1, 0) \
    X(8192,  32, 32, 32,  8,  2, 1, 1) \
    X(16384, 32, 32, 32, 16,  1, 1, 1)
// (16384 real forward: TWR measured .686 without and .672 with it and keeps the LUT)
#define CKB_SPLIT_PREFETCH_PLANS_R2C(X) \
    X(2048,  32, 32, 32,  2,  6, 1, 0) \
    X(4096,  32, 32, 32,  4,  3, 1, 1) \
    X(8192,  32, 32, 32,  8,  2, 1, 0) \
    X(16384, 32, 32, 32, 16,  1, 1, 0)
// (R2C M = 8192, real n = 16384: in place G = 1 x 2 CTAs, TWR .75; split G = 1 x 2 CTAs, TWR .74; split G = 2, TWR .61 -- the
//  register twiddle bases next to the constant split factors spill; split G = 2 with the LUT .815)
// The audio front end (window + real forward + power spectrum) is issue-bound and wants the 16 warps per SM of the two-CTA
// in-place plans: 2048 points in place .72, split G = 6 .66, split G = 8 .65, in place G = 5 x 2 CTAs (102 registers) .52.
#define CKB_INPLACE_PREFETCH_PLANS_AUDIO(X) \
    X(2048,  32, 32, 32,  2,  4, 2, 1) \
    X(4096,  32, 32, 32,  4,  2, 2, 1) \
    X(8192,  32, 32, 32,  8,  1, 2, 1)
#define CKB_SPLIT_PREFETCH_PLANS_AUDIO(X) \
    X(16384, 32, 32, 32, 16,  1, 1, 0)

// In-place prefetch variants (Cfg::PF == PF_INPLACE), complex transforms only.  Measured on B200 (fraction of
// the 6.55 TB/s copy peak, without -> with): 256 .89->.94, 512 .87->.95, 2048 .90->.95, 4096 .66->.95,
// 8192 .74->.87; rows shorter than 2 KiB lose (too many tiny bulk copies), 16384 loses (.64->.59).  Register stage
// twiddles (TWR) for 2048 / 4096 complex: .958 -> .922, .947 -> .944, so those keep the LUT.
#define CKB_INPLACE_PREFETCH_PLANS(X) \
    X(256,   16, 16, 16,  1,  8, 4, 1) \
    X(512,   32, 32, 16,  1,  8, 4, 0) \
    X(2048,  32, 32, 32,  2,  4, 2, 0) \
    X(4096,  32, 32, 32,  4,  2, 2, 0) \
    X(8192,  32, 32, 32,  8,  1, 2, 1)
// Split-complex ("planar") rows with in-place prefetch: two bulk copies per row (real parts, imaginary parts); rows must
// be 16-byte aligned (pointers and strides in multiples of 4 floats).  Measured (plain loads -> prefetch): 2048 .85 -> .94,
// 4096 .67 -> .89, 8192 .71 -> .87; shorter rows lose (256 .90 -> .83, 512 .91 -> .86, 1024 .90 -> .88) and keep the plain
// loads.  16384 points use the split prefetch with the two planes as the two halves.
#define CKB_INPLACE_PREFETCH_PLANS_PLANAR(X) \
    X(2048,  32, 32, 32,  2,  4, 2, 0) \
    X(4096,  32, 32, 32,  4,  2, 2, 0) \
    X(8192,  32, 32, 32,  8,  1, 2, 1)
// (Real-forward transforms of 4096 .. 16384 points used in-place prefetch with register stage twiddles in round 1 --
// 2048 .87 -> .89, 4096 .77 -> .80 -- and moved to the multi-group split prefetch above in round 2.)

// Real-inverse in-place prefetch (rows are bulk-copied from the 16-byte boundary below them, twisted in place).
// (16384: the row does not leave room for a second buffer and the split prefetch needs the halves at different times,
// but the twist needs the mirror pairs together; in-place prefetch measured .47 -> .55; with register twiddles on top .49.)
// (shorter rows lose: in-place prefetch at M = 128 / 256 / 512 / 1024 measured .75 -> .63, .95 -> .69, .85 -> .70, .95 -> .81)
#define CKB_INPLACE_PREFETCH_PLANS_C2R(X) \
    X(2048,  32, 32, 32,  2,  4, 2, 1) \
    X(4096,  32, 32, 32,  4,  2, 2, 1) \
    X(8192,  32, 32, 32,  8,  1, 2, 1)      /* round 2: G = 1 x 2 CTAs with the LUT .73, G = 2 x 1 CTA .77 / .78, table factors .73 */ \
    X(16384, 32, 32, 32, 16,  1, 1, 1)      /* round 2: register twiddles + constant twist factors .53 -> .72 */

#define CKB_MAX_SINGLE_PASS 16384   /* largest complex length done in one launch */
#define CKB_MAX_TABLE 32768         /* device twiddle table W_Nt^k covers real n up to this in one pass */
