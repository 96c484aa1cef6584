// Bit-exact restatement of the two libm functions the reference's hot path depends on.
//
//   jsde_log(x)   == glibc 2.39 log(x)   for normal x > 0        (reference `jitcsde/random_numbers.c:127`)
//   jsde_pow(x,y) == glibc 2.39 pow(x,y) for finite x > 0, |y| in [2^-65, 2^63)
//                                                                 (reference `jitcsde/_jitcsde.py:587,592`: Python `**`)
//
// glibc ≥ 2.28 implements both with the ARM optimized-routines algorithm; on x86-64 with FMA3 the ifunc selects
// the variant in which the marked operations are fused.  The op/FMA sequence below follows SURVEY.md Appendix B
// (decoded from this image's libm and checked there); tests/test_libm_exact.py re-checks THIS file, compiled by
// gcc for the host, against the running libm on millions of inputs, and tests/test_gpu_rng.py checks the device
// build against the host build.  Every `JSDE_FMA` is one fused multiply-add; every other + - * is rounded
// separately — compile with `-fmad=false` (nvcc) / `-ffp-contract=off` (gcc).
//
// The same source text serves the device (`__device__`, tables in global memory read through the read-only
// path) and the host self-check (plain C, `static const` tables).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define JSDE_FN static __device__ __forceinline__
#define JSDE_TABLE_QUAL static __device__ const
// pow runs for one or two lanes of a warp at a time (controller, after a rejection or when the step may grow), so its
// tables sit in constant memory: few distinct addresses per access, and the lookup does not queue in the load/store
// unit behind the generator's outstanding asynchronous copies.
#define JSDE_POWTABLE_QUAL static __constant__ const
#define JSDE_FMA(a, b, c) __fma_rn((a), (b), (c))
#define JSDE_TAB(arr, i) __ldg(&(arr)[(i)])
#define JSDE_PTAB(arr, i) ((arr)[(i)])
#define JSDE_ASU(x) ((uint64_t)__double_as_longlong(x))
#define JSDE_ASD(u) __longlong_as_double((long long)(u))
#else
#include <math.h>
#include <string.h>
#define JSDE_FN static inline
#define JSDE_TABLE_QUAL static const
#define JSDE_FMA(a, b, c) fma((a), (b), (c))
#define JSDE_TAB(arr, i) ((arr)[(i)])
#define JSDE_PTAB(arr, i) ((arr)[(i)])
static inline uint64_t jsde_asu_(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double jsde_asd_(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
#define JSDE_ASU(x) jsde_asu_(x)
#define JSDE_ASD(u) jsde_asd_(u)
#endif

#include "glibc_tables.h"

#if defined(__CUDACC__) && !defined(JSDE_LITERAL_COEFFICIENTS)
// Device build: the scalar coefficients come from constant memory, so that every DFMA/DMUL takes them as a c[bank][offset]
// operand.  As literals each 64-bit coefficient costs two move instructions per use (no FP64 immediates): 104 + 81 of
// the fused kernel's stepper instructions were such moves.
static __constant__ double JSDE_COEFFICIENTS[] = {
	JSDE_LN2HI, JSDE_LN2LO, JSDE_EXP_INVLN2N, JSDE_EXP_SHIFT, JSDE_EXP_NEGLN2HIN, JSDE_EXP_NEGLN2LON, JSDE_LOG_A0, JSDE_LOG_A1, JSDE_LOG_A2, JSDE_LOG_A3, JSDE_LOG_A4, JSDE_LOG_B0, JSDE_LOG_B1, JSDE_LOG_B2, JSDE_LOG_B3, JSDE_LOG_B4, JSDE_LOG_B5, JSDE_LOG_B6, JSDE_LOG_B7, JSDE_LOG_B8, JSDE_LOG_B9, JSDE_LOG_B10, JSDE_POW_A0, JSDE_POW_A1, JSDE_POW_A2, JSDE_POW_A3, JSDE_POW_A4, JSDE_POW_A5, JSDE_POW_A6, JSDE_EXP_C2, JSDE_EXP_C3, JSDE_EXP_C4, JSDE_EXP_C5
};
#undef JSDE_LN2HI
#define JSDE_LN2HI (JSDE_COEFFICIENTS[0])
#undef JSDE_LN2LO
#define JSDE_LN2LO (JSDE_COEFFICIENTS[1])
#undef JSDE_EXP_INVLN2N
#define JSDE_EXP_INVLN2N (JSDE_COEFFICIENTS[2])
#undef JSDE_EXP_SHIFT
#define JSDE_EXP_SHIFT (JSDE_COEFFICIENTS[3])
#undef JSDE_EXP_NEGLN2HIN
#define JSDE_EXP_NEGLN2HIN (JSDE_COEFFICIENTS[4])
#undef JSDE_EXP_NEGLN2LON
#define JSDE_EXP_NEGLN2LON (JSDE_COEFFICIENTS[5])
#undef JSDE_LOG_A0
#define JSDE_LOG_A0 (JSDE_COEFFICIENTS[6])
#undef JSDE_LOG_A1
#define JSDE_LOG_A1 (JSDE_COEFFICIENTS[7])
#undef JSDE_LOG_A2
#define JSDE_LOG_A2 (JSDE_COEFFICIENTS[8])
#undef JSDE_LOG_A3
#define JSDE_LOG_A3 (JSDE_COEFFICIENTS[9])
#undef JSDE_LOG_A4
#define JSDE_LOG_A4 (JSDE_COEFFICIENTS[10])
#undef JSDE_LOG_B0
#define JSDE_LOG_B0 (JSDE_COEFFICIENTS[11])
#undef JSDE_LOG_B1
#define JSDE_LOG_B1 (JSDE_COEFFICIENTS[12])
#undef JSDE_LOG_B2
#define JSDE_LOG_B2 (JSDE_COEFFICIENTS[13])
#undef JSDE_LOG_B3
#define JSDE_LOG_B3 (JSDE_COEFFICIENTS[14])
#undef JSDE_LOG_B4
#define JSDE_LOG_B4 (JSDE_COEFFICIENTS[15])
#undef JSDE_LOG_B5
#define JSDE_LOG_B5 (JSDE_COEFFICIENTS[16])
#undef JSDE_LOG_B6
#define JSDE_LOG_B6 (JSDE_COEFFICIENTS[17])
#undef JSDE_LOG_B7
#define JSDE_LOG_B7 (JSDE_COEFFICIENTS[18])
#undef JSDE_LOG_B8
#define JSDE_LOG_B8 (JSDE_COEFFICIENTS[19])
#undef JSDE_LOG_B9
#define JSDE_LOG_B9 (JSDE_COEFFICIENTS[20])
#undef JSDE_LOG_B10
#define JSDE_LOG_B10 (JSDE_COEFFICIENTS[21])
#undef JSDE_POW_A0
#define JSDE_POW_A0 (JSDE_COEFFICIENTS[22])
#undef JSDE_POW_A1
#define JSDE_POW_A1 (JSDE_COEFFICIENTS[23])
#undef JSDE_POW_A2
#define JSDE_POW_A2 (JSDE_COEFFICIENTS[24])
#undef JSDE_POW_A3
#define JSDE_POW_A3 (JSDE_COEFFICIENTS[25])
#undef JSDE_POW_A4
#define JSDE_POW_A4 (JSDE_COEFFICIENTS[26])
#undef JSDE_POW_A5
#define JSDE_POW_A5 (JSDE_COEFFICIENTS[27])
#undef JSDE_POW_A6
#define JSDE_POW_A6 (JSDE_COEFFICIENTS[28])
#undef JSDE_EXP_C2
#define JSDE_EXP_C2 (JSDE_COEFFICIENTS[29])
#undef JSDE_EXP_C3
#define JSDE_EXP_C3 (JSDE_COEFFICIENTS[30])
#undef JSDE_EXP_C4
#define JSDE_EXP_C4 (JSDE_COEFFICIENTS[31])
#undef JSDE_EXP_C5
#define JSDE_EXP_C5 (JSDE_COEFFICIENTS[32])
#endif

// ---- log --------------------------------------------------------------------------------------------------
// glibc splits on |x-1|: inputs in [1-2^-4, 1+0x1.09p-4) take a table-free polynomial, everything else the table
// path.  The two halves are exposed separately so that a SIMT caller can run the (long, rare: 6 % of polar radii)
// near-one half only for the lanes and components that need it.
JSDE_FN int jsde_log_is_near_one(double x)
{
	const uint64_t LO = 0x3fee000000000000ULL;  // asuint64(1 - 0x1p-4)
	const uint64_t HI = 0x3ff1090000000000ULL;  // asuint64(1 + 0x1.09p-4)
	return JSDE_ASU(x) - LO < HI - LO;
}

JSDE_FN double jsde_log_near_one(double x)
{
	if (JSDE_ASU(x) == 0x3ff0000000000000ULL) return 0.0;
	const double r = x - 1.0;
	const double r2 = r * r;
	const double r3 = r * r2;
	const double P1 = JSDE_FMA(r2, JSDE_LOG_B3, JSDE_FMA(r, JSDE_LOG_B2, JSDE_LOG_B1));
	const double P2 = JSDE_FMA(r2, JSDE_LOG_B6, JSDE_FMA(r, JSDE_LOG_B5, JSDE_LOG_B4));
	const double P3 = JSDE_FMA(r3, JSDE_LOG_B10, JSDE_FMA(r2, JSDE_LOG_B9, JSDE_FMA(r, JSDE_LOG_B8, JSDE_LOG_B7)));
	const double Q = JSDE_FMA(JSDE_FMA(P3, r3, P2), r3, P1);
	// split r into rhi + rlo so that rhi*rhi*B0 is exact enough
	const double w = JSDE_FMA(r, 0x1p27, r);
	const double rhi = JSDE_FMA(-0x1p27, r, w);
	const double rlo = r - rhi;
	const double s = rhi * rhi;
	const double hi = JSDE_FMA(s, JSDE_LOG_B0, r);
	const double lo = JSDE_FMA(s, JSDE_LOG_B0, r - hi);
	const double lo2 = JSDE_FMA(JSDE_LOG_B0 * rlo, rhi + r, lo);
	return JSDE_FMA(Q, r3, lo2) + hi;
}

// x = 2^k z, z in [0x1.69555p-1, 0x1.69555p0); table on the top 7 mantissa bits of z.  `tab` holds {invc, logc}
// pairs (128 entries); the kernels pass a shared-memory copy so that the lookup does not queue behind outstanding
// global-memory traffic.
JSDE_FN double jsde_log_table_from(double x, const double *tab)
{
	const uint64_t ix = JSDE_ASU(x);
	const uint64_t tmp = ix - 0x3fe6000000000000ULL;
	const int i = (int)((tmp >> 45) & 127);
	const int64_t k = (int64_t)tmp >> 52;
	const uint64_t iz = ix - (tmp & (0xfffULL << 52));
	const double z = JSDE_ASD(iz);
	const double kd = (double)k;
	const double invc = tab ? tab[2 * i] : JSDE_TAB(JSDE_LOG_INVC, i);
	const double logc = tab ? tab[2 * i + 1] : JSDE_TAB(JSDE_LOG_LOGC, i);
	const double w = JSDE_FMA(kd, JSDE_LN2HI, logc);
	const double r = JSDE_FMA(z, invc, -1.0);
	const double hi = w + r;
	const double lo = JSDE_FMA(kd, JSDE_LN2LO, (w - hi) + r);
	const double r2 = r * r;
	const double r3 = r * r2;
	const double q = JSDE_FMA(JSDE_FMA(r, JSDE_LOG_A4, JSDE_LOG_A3), r2, JSDE_FMA(r, JSDE_LOG_A2, JSDE_LOG_A1));
	return JSDE_FMA(r3, q, JSDE_FMA(r2, JSDE_LOG_A0, lo)) + hi;
}

JSDE_FN double jsde_log_table(double x) { return jsde_log_table_from(x, (const double *)0); }

JSDE_FN double jsde_log(double x)
{
	return jsde_log_is_near_one(x) ? jsde_log_near_one(x) : jsde_log_table(x);
}

// ---- pow --------------------------------------------------------------------------------------------------
// Restricted to what the controller can ask for: x finite and > 0 (subnormals included), y finite with
// 2^-65 <= |y| < 2^63.  x == +inf / 0 / NaN and the trivial y are the caller's business (see controller.cuh).
// Overflow/underflow of the result cannot occur for |y| < 1; the exp part therefore omits glibc's
// `specialcase` branch and jsde_pow_in_domain() tells the caller when that omission would matter.
JSDE_FN int jsde_pow_in_domain(double x, double y)
{
	const uint64_t ix = JSDE_ASU(x);
	const uint64_t ay = JSDE_ASU(y) & 0x7fffffffffffffffULL;
	return ix > 0 && ix < 0x7ff0000000000000ULL && ay >= 0x3be0000000000000ULL /*2^-65*/ && ay <= 0x3ff0000000000000ULL /*1*/;
}

// where the lookups go: the tables of glibc_tables.h (host; device: constant memory — fine for a warp-uniform caller)
// or copies of them staged in shared memory by the caller (the lane-per-trajectory controller looks up 32 different
// entries at once, which constant memory would serialise)
typedef struct {
	const double *invc, *logc, *logctail;
	const unsigned long long *exp_tab;
} jsde_pow_tables;

JSDE_FN double jsde_pow_with(double x, double y, const jsde_pow_tables T)
{
	uint64_t ix = JSDE_ASU(x);
	if ((ix >> 52) == 0) {
		// subnormal x: normalise so that the exponent becomes negative (glibc pow.c, "topx == 0")
		ix = JSDE_ASU(x * 0x1p52);
		ix &= 0x7fffffffffffffffULL;
		ix -= 52ULL << 52;
	}
	// log_inline: log(x) = hi + lo with ~68 bit accuracy
	const uint64_t tmp = ix - 0x3fe6955500000000ULL;
	const int i = (int)((tmp >> 45) & 127);
	const int64_t k = (int64_t)tmp >> 52;
	const uint64_t iz = ix - (tmp & (0xfffULL << 52));
	const double z = JSDE_ASD(iz);
	const double kd = (double)k;
	const double invc = T.invc[i];
	const double logc = T.logc[i];
	const double logctail = T.logctail[i];
	const double t1 = JSDE_FMA(kd, JSDE_LN2HI, logc);
	const double lo1 = JSDE_FMA(kd, JSDE_LN2LO, logctail);
	const double r = JSDE_FMA(z, invc, -1.0);
	const double ar = r * JSDE_POW_A0;
	const double p12 = JSDE_FMA(r, JSDE_POW_A2, JSDE_POW_A1);
	const double p34 = JSDE_FMA(r, JSDE_POW_A4, JSDE_POW_A3);
	const double t2 = r + t1;
	const double lo2 = (t1 - t2) + r;
	const double ar2 = r * ar;
	const double ar3 = r * ar2;
	const double lo3 = JSDE_FMA(ar, r, -ar2);
	const double hi = t2 + ar2;
	const double p56 = JSDE_FMA(r, JSDE_POW_A6, JSDE_POW_A5);
	const double lo4 = (t2 - hi) + ar2;
	const double br = JSDE_FMA(ar2, JSDE_FMA(p56, ar2, p34), p12);
	const double lo = JSDE_FMA(ar3, br, ((lo1 + lo2) + lo3) + lo4);
	const double yh = hi + lo;
	const double yl = (hi - yh) + lo;
	// y*log(x) = ehi + elo
	const double ehi = y * yh;
	const double elo = JSDE_FMA(y, yl, JSDE_FMA(yh, y, -ehi));
	// exp_inline
	const uint32_t abstop = (uint32_t)(JSDE_ASU(ehi) >> 52) & 0x7ff;
	if (abstop < 0x3c9)  // |y log x| < 2^-54: result rounds to 1 +- tiny
		return 1.0 + ehi;
	const double ks = JSDE_FMA(ehi, JSDE_EXP_INVLN2N, JSDE_EXP_SHIFT);
	const uint64_t ki = JSDE_ASU(ks);
	const double kd2 = ks - JSDE_EXP_SHIFT;
	double rr = JSDE_FMA(kd2, JSDE_EXP_NEGLN2LON, JSDE_FMA(kd2, JSDE_EXP_NEGLN2HIN, ehi));
	rr = elo + rr;
	const int idx = 2 * (int)(ki & 127);
	const double tail = JSDE_ASD(T.exp_tab[idx]);
	const uint64_t sbits = T.exp_tab[idx + 1] + (ki << 45);
	const double scale = JSDE_ASD(sbits);
	const double rr2 = rr * rr;
	const double q = JSDE_FMA(JSDE_FMA(rr, JSDE_EXP_C3, JSDE_EXP_C2), rr2, rr + tail);
	const double tm = JSDE_FMA(JSDE_FMA(rr, JSDE_EXP_C5, JSDE_EXP_C4), rr2 * rr2, q);
	return JSDE_FMA(tm, scale, scale);
}

JSDE_FN double jsde_pow(double x, double y)
{
	const jsde_pow_tables T = {JSDE_POW_INVC, JSDE_POW_LOGC, JSDE_POW_LOGCTAIL, JSDE_EXP_TAB};
	return jsde_pow_with(x, y, T);
}

// glibc 2.39 exp(x) (sysdeps/ieee754/dbl-64/e_exp.c, the FMA variant the x86-64 ifunc selects), for |x| < 512: generated
// drift/diffusion code that calls exp() then gives the same bits as the reference's compiled C (e.g. the diffusion
// 5*exp(-y^2+y-1.5)+3 of the reference's own validation scenario, tests/validation.py:21-32).  Same table and polynomial
// as the exp_inline at the end of pow above.  Outside that range (overflow/underflow handling) the caller's ordinary
// exp() is used — tolerance parity only there.
JSDE_FN int jsde_exp_in_range(double x)
{
	const uint32_t abstop = (uint32_t)(JSDE_ASU(x) >> 52) & 0x7ff;
	return abstop - 0x3c9u < 0x408u - 0x3c9u;  // 2^-54 <= |x| < 512
}

JSDE_FN double jsde_exp_core(double x)
{
	const double ks = JSDE_FMA(JSDE_EXP_INVLN2N, x, JSDE_EXP_SHIFT);
	const uint64_t ki = JSDE_ASU(ks);
	const double kd = ks - JSDE_EXP_SHIFT;
	const double r = JSDE_FMA(kd, JSDE_EXP_NEGLN2LON, JSDE_FMA(kd, JSDE_EXP_NEGLN2HIN, x));
	const int idx = 2 * (int)(ki & 127);
	const double tail = JSDE_ASD(JSDE_PTAB(JSDE_EXP_TAB, idx));
	const uint64_t sbits = JSDE_PTAB(JSDE_EXP_TAB, idx + 1) + (ki << 45);
	const double scale = JSDE_ASD(sbits);
	const double r2 = r * r;
	const double q = JSDE_FMA(JSDE_FMA(r, JSDE_EXP_C3, JSDE_EXP_C2), r2, tail + r);
	const double tm = JSDE_FMA(JSDE_FMA(r, JSDE_EXP_C5, JSDE_EXP_C4), r2 * r2, q);
	return JSDE_FMA(scale, tm, scale);
}

JSDE_FN double jsde_exp(double x)
{
	if (jsde_exp_in_range(x))
		return jsde_exp_core(x);
	const uint32_t abstop = (uint32_t)(JSDE_ASU(x) >> 52) & 0x7ff;
	if (abstop < 0x3c9u)
		return 1.0 + x;  // |x| < 2^-54 (glibc: "tiny")
	return exp(x);
}
