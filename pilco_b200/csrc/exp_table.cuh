// exp_table.cuh -- the table-driven fp64 exp of the tile kernels (forward and backward moment match).
//
// Shared source for the device build and for a HOST build (tests/host_harness/exp_harness.cpp) that the CPU suite
// checks against libm: the bodies below are the code the kernels execute; only the bit-level helpers differ
// (device intrinsics / inline PTX vs. portable C).  Exponents are carried PRE-SCALED by EXP_SC = EXP_TAB / ln 2.
#pragma once
#include <math.h>
#include <stdint.h>

// Two table layouts (compile-time):
//   default        1024 entries, cubic expm1 remainder: 7 fp64-pipe instructions per exp; the data-dependent 8-byte gather
//                  from a 1024-entry shared-memory table costs ~6.4 LSU wavefronts per warp (bank conflicts)
//   PILCO_EXP256   256 entries x 16 interleaved copies (32 KB): lane l reads copy l & 15, i.e. ALWAYS its own bank pair:
//                  2 wavefronts per gather (the minimum) whatever the indices; the remainder is 4x wider, so the
//                  polynomial needs one more term: 8 fp64-pipe instructions per exp
#ifdef PILCO_EXP256
#define EXP_TAB 256
#define EXP_SHIFT 8
#define EXP_STRIDE_LOG2 4
#define EXP_SC 369.32993046757463           // EXP_TAB / ln 2
#define EXP_CLAMP_HI 0xC10F8000u            // scaled exponent floor -258048 (= -698.7 unscaled)
#else
#define EXP_TAB 1024
#define EXP_SHIFT 10
#define EXP_STRIDE_LOG2 0
#define EXP_SC 1477.3197218702985           // EXP_TAB / ln 2: exponents are carried PRE-SCALED by this factor
#define EXP_CLAMP_HI 0xC12F8000u            // scaled exponent floor -1032192 (= -698.7 unscaled)
#endif
#define EXP_STRIDE (1 << EXP_STRIDE_LOG2)   // interleaved copies of the table (entry j of copy c at [j * EXP_STRIDE + c])
#define EXP_TAB_DOUBLES (EXP_TAB * EXP_STRIDE)
// a lane's view of the table: its own copy (copy 0 when there is only one)
#define EXP_LANE_TAB(tab, lane) ((tab) + ((lane) & (EXP_STRIDE - 1)))

#if defined(__CUDACC__)
#define PILCO_EXP_FN __device__ __forceinline__
#define PX_LO(x) __double2loint(x)
#define PX_HI(x) __double2hiint(x)
#define PX_MK(hi, lo) __hiloint2double(hi, lo)
#define PX_FUNNEL_R(lo, hi, sh) ((int)__funnelshift_r((unsigned)(lo), (unsigned)(hi), sh))
#define PX_UMIN(a, b) min(a, b)
#define PX_IMAX(a, b) max(a, b)
#define PX_MADEXP(out, k, hi) asm("mad.lo.s32 %0, %1, 0x100000, %2;" : "=r"(out) : "r"(k), "r"(hi))
#else
#include <string.h>
#define PILCO_EXP_FN static inline
static inline int px_lo_(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(uint32_t)u; }
static inline int px_hi_(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(uint32_t)(u >> 32); }
static inline double px_mk_(int hi, int lo) { uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &u, 8); return x; }
#define PX_LO(x) px_lo_(x)
#define PX_HI(x) px_hi_(x)
#define PX_MK(hi, lo) px_mk_(hi, lo)
#define PX_FUNNEL_R(lo, hi, sh) ((int)(uint32_t)(((((uint64_t)(uint32_t)(hi)) << 32) | (uint32_t)(lo)) >> (sh)))
#define PX_UMIN(a, b) ((a) < (b) ? (a) : (b))
#define PX_IMAX(a, b) ((a) > (b) ? (a) : (b))
#define PX_MADEXP(out, k, hi) (out) = (int)((uint32_t)(k) * 0x100000u + (uint32_t)(hi))
#endif

// byte offset of table entry (i mod EXP_TAB) of the lane's copy;  expm1(r s)/r with s = ln2 / EXP_TAB, |r| <= 1/2
#define EXP_TAB_OFFSET(i) (((i) << (3 + EXP_STRIDE_LOG2)) & ((EXP_TAB - 1) << (3 + EXP_STRIDE_LOG2)))
#ifdef PILCO_EXP256
#define EXP_POLY(q, r) double q = 2.239395190875157e-12;                      /* s^4/24 : degree 4, next term s^5/120 r^4 < 4e-17 */ \
                       q = fma(q, r, 3.308302680541371e-09);                  /* s^3/6 */ \
                       q = fma(q, r, 3.665565596910106e-06);                  /* s^2/2 */ \
                       q = fma(q, r, 0.0027076061740622863)                   /* s     */
#else
#define EXP_POLY(q, r) double q = 5.169222938345892e-11;   /* cubic, x^4 term economised into the x^2 coefficient */ \
                       q = fma(q, r, 2.2909785199379098e-07);  /* (max abs error 1.4e-16) */ \
                       q = fma(q, r, 0.0006769015435155716)
#endif

// exp(x) for a PRE-SCALED argument xs = x * EXP_SC (the setup kernels fold EXP_SC into A', B, U', so the DMMA
// delivers xs directly):  xs = 1024 k + j + r',  exp(x) = 2^k T[j] exp(r' ln2/1024),  |r'| <= 1/2.
// 7 fp64-pipe instructions (3 add, 2 fma, 1 mul, 1 fma) + integer ops + one shared-memory table read.
PILCO_EXP_FN double exp_scaled(double xs, const double* __restrict__ tab) {
    // clamp xs >= EXP_CLAMP with ONE integer instruction: for negative doubles a larger magnitude is a larger
    // high word, positive values (high word < 0x80000000) pass unchanged, NaNs propagate
    xs = PX_MK((int)PX_UMIN((unsigned)PX_HI(xs), EXP_CLAMP_HI), PX_LO(xs));
    const double MAGIC = 6755399441055744.0;              // 1.5 * 2^52
    const double t  = xs + MAGIC;                         // round to integer in the low mantissa bits
    const int    ki = PX_LO(t);
    const double kd = t - MAGIC;
    const double r  = xs - kd;                            // exact, in [-1/2, 1/2]
    EXP_POLY(q, r);
    const double tj = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(tab) + EXP_TAB_OFFSET(ki));
    const double em1 = q * r;                             // expm1(r ln2/1024), |.| < 3.4e-4
    const double v = fma(tj, em1, tj);                    // T[j] * exp(.), in [1, 2.01)
    // scale by 2^k, k = ki >> 10 >= -1008 after the clamp (v normal): one shift + one integer multiply-add
    int hi;
    PX_MADEXP(hi, ki >> EXP_SHIFT, PX_HI(v));
    return PX_MK(hi, PX_LO(v));
}

// exp((c + Ai) / EXP_SC) for a pre-scaled exponent c plus an INTEGER row offset Ai that the caller folds into the
// rounding constant: am = EXP_MAGIC + Ai (exact for |Ai| < 2^50).  One DADD cheaper than forming c + A first:
//   t = c + am -> integer field round(c) + Ai;  kd = t - am = round(c);  r = c - kd in [-1/2, 1/2]  (all exact)
// The caller multiplies the row's accumulated sums by exp((A - Ai)/EXP_SC) once (exp_row_split below).
// The 2^k exponent is taken with a funnel shift from the 64-bit integer field, so it is right for
// |c + Ai| < 2^(31 + EXP_SHIFT) (|log-kernel value| < 1.4e9 in both layouts), and clamped below at 2^-1008
// (result ~1e-304, i.e. 0).
#define EXP_MAGIC 6755399441055744.0                      // 1.5 * 2^52
PILCO_EXP_FN double exp_shifted(double c, double am, const double* __restrict__ tab) {
    const double t  = c + am;
    const int    lo = PX_LO(t), hi = PX_HI(t);
    const double kd = t - am;
    const double r  = c - kd;
    EXP_POLY(q, r);
    const double tj = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(tab) + EXP_TAB_OFFSET(lo));
    const double em1 = q * r;
    const double v = fma(tj, em1, tj);
    int k = PX_FUNNEL_R(lo, hi, EXP_SHIFT);                                // bits 10..41 of the integer field
    k = PX_IMAX(k, -1008);
    int vh;
    PX_MADEXP(vh, k, PX_HI(v));
    return PX_MK(vh, PX_LO(v));
}
// split a pre-scaled row exponent A into am = EXP_MAGIC + rint(A) and the row factor exp((A - rint(A))/EXP_SC)
PILCO_EXP_FN void exp_row_split(double A, double& am, double& rowfac) {
    const double Ai = rint(A);
    am = EXP_MAGIC + Ai;
    const double y = (A - Ai) * (1.0 / EXP_SC);           // |y| <= 0.5/EXP_SC = 3.4e-4: degree-4 Taylor, error < 4e-20
    rowfac = fma(y, fma(y, fma(y, fma(y, 1.0 / 24.0, 1.0 / 6.0), 0.5), 1.0), 1.0);
}

