// Marsaglia polar method exactly as reference jitcsde/random_numbers.c:100-130 computes it.
//
//   rk_double:  a = w0>>5 (27 bit), b = w1>>6 (26 bit), u = (a*2^26 + b) / 2^53            — every step exact
//   candidate:  x2 = 2u-1 from words 0,1;  x1 = 2u-1 from words 2,3  ("x2 first", :120-121) — exact
//               r2 = x1*x1 + x2*x2   (three roundings, NOT fused)
//   accepted:   f = scale*sqrt(-2.0*log(r2)/r2);  DW = x1*f;  DZ = x2*f                      (:127-129)
//
// 2u-1 = (k - 2^52) / 2^52 with k = a*2^26+b < 2^53 is exactly representable, so it is formed from the signed integer
// directly (one conversion + one exact scaling instead of the reference's seven exact operations).
#pragma once

#include "glibc_math.h"

namespace jsde {

__device__ __forceinline__ double polar_coordinate(uint32_t w_hi, uint32_t w_lo)
{
	const uint64_t k = ((uint64_t)(w_hi >> 5) << 26) | (uint64_t)(w_lo >> 6);
	return __ll2double_rn((long long)k - (1LL << 52)) * 0x1p-52;
}

// The same value without the 64-bit integer-to-double conversion: with a = w_hi>>5 (27 bit), b = w_lo>>6 (26 bit),
// 2u-1 = a*2^-26 + b*2^-52 - 1.  The doubles 2^52+a and 2^52+b are formed by placing a, b in the low mantissa word
// under the exponent of 2^52; (2^52+a)*2^-26 - 2^26 = a*2^-26 and (2^52+b)*2^-52 - 2 = b*2^-52 - 1 are exact as fused
// operations (multiples of 2^-26 below 2, resp. of 2^-52 in (-1, 0)), and so is their sum (a multiple of 2^-52 of
// magnitude below 1).  Three FP64 operations, no conversion unit.
__device__ __forceinline__ double polar_coordinate_exact(uint32_t w_hi, uint32_t w_lo)
{
	const double A = __hiloint2double(0x43300000, (int)(w_hi >> 5));
	const double B = __hiloint2double(0x43300000, (int)(w_lo >> 6));
	const double ta = __fma_rn(A, 0x1p-26, -0x1p26);
	const double tb = __fma_rn(B, 0x1p-52, -2.0);
	return __dadd_rn(ta, tb);
}

// `w` = four tempered words in stream order.  Returns true iff the reference's do-while would leave the loop.
__device__ __forceinline__ bool polar_candidate(uint4 w, double &x1, double &x2, double &r2)
{
	x2 = polar_coordinate_exact(w.x, w.y);
	x1 = polar_coordinate_exact(w.z, w.w);
	r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));
	return r2 < 1.0 && r2 != 0.0;
}

// scale-independent part of the Box–Muller factor
__device__ __forceinline__ double polar_radius_factor(double r2)
{
	return sqrt(__ddiv_rn(__dmul_rn(-2.0, jsde_log(r2)), r2));
}

// the same with glibc's log table in shared memory (`log_tab`, may be null: global table)
__device__ __forceinline__ double polar_radius_factor_with(double r2, const double *log_tab)
{
	const double lg = jsde_log_is_near_one(r2) ? jsde_log_near_one(r2) : jsde_log_table_from(r2, log_tab);
	return sqrt(__ddiv_rn(__dmul_rn(-2.0, lg), r2));
}

}  // namespace jsde
