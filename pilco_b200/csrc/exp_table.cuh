// exp_table.cuh -- the table-driven fp64 exp of the tile kernels (forward and backward moment match).
//
// Shared source for the device build and for a HOST build (tests/host_harness/exp_harness.cpp) that the CPU suite
// checks against libm: the bodies below are the code the kernels execute; only the bit-level helpers differ
// (device intrinsics / inline PTX vs. portable C).  Exponents are carried PRE-SCALED by EXP_SC = EXP_TAB / ln 2.
#pragma once
#include <math.h>
#include <stdint.h>

#define EXP_TAB 1024
#define EXP_SHIFT 10
#define EXP_SC 1477.3197218702985           // EXP_TAB / ln 2: exponents are carried PRE-SCALED by this factor
#define EXP_CLAMP (-1032192.0)              // scaled exponent floor (= -698.7 unscaled), hi word 0xC12F8000

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

// exp(x) for a PRE-SCALED argument xs = x * EXP_SC (the setup kernels fold EXP_SC into A', B, U', so the DMMA
// delivers xs directly):  xs = 1024 k + j + r',  exp(x) = 2^k T[j] exp(r' ln2/1024),  |r'| <= 1/2.
// 7 fp64-pipe instructions (3 add, 2 fma, 1 mul, 1 fma) + integer ops + one shared-memory table read.
PILCO_EXP_FN double exp_scaled(double xs, const double* __restrict__ tab) {
    // clamp xs >= EXP_CLAMP with ONE integer instruction: for negative doubles a larger magnitude is a larger
    // high word, positive values (high word < 0x80000000) pass unchanged, NaNs propagate
    xs = PX_MK((int)PX_UMIN((unsigned)PX_HI(xs), 0xC12F8000u), PX_LO(xs));
    const double MAGIC = 6755399441055744.0;              // 1.5 * 2^52
    const double t  = xs + MAGIC;                         // round to integer in the low mantissa bits
    const int    ki = PX_LO(t);
    const double kd = t - MAGIC;
    const double r  = xs - kd;                            // exact, in [-1/2, 1/2]
    double q = 5.169222938345892e-11;                     // expm1(r s)/r, s = ln2/1024: cubic, x^4 term economised
    q = fma(q, r, 2.2909785199379098e-07);                // into the x^2 coefficient (max abs error 1.4e-16)
    q = fma(q, r, 0.0006769015435155716);
    const double tj = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(tab) + ((ki << 3) & ((EXP_TAB - 1) << 3)));
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
// |c + Ai| < 2^41 (|log-kernel value| < 1.4e9), and clamped below at 2^-1008 (result ~1e-304, i.e. 0).
#define EXP_MAGIC 6755399441055744.0                      // 1.5 * 2^52
PILCO_EXP_FN double exp_shifted(double c, double am, const double* __restrict__ tab) {
    const double t  = c + am;
    const int    lo = PX_LO(t), hi = PX_HI(t);
    const double kd = t - am;
    const double r  = c - kd;
    double q = 5.169222938345892e-11;                     // expm1(r s)/r, s = ln2/1024: cubic, x^4 term economised
    q = fma(q, r, 2.2909785199379098e-07);                // into the x^2 coefficient (max abs error 1.4e-16)
    q = fma(q, r, 0.0006769015435155716);
    const double tj = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(tab) + ((lo << 3) & ((EXP_TAB - 1) << 3)));
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

