// Counter-based jump variates (SURVEY.md §8 f item 3): the waiting-time variate and the standard-normal amplitude variate
// of jump j of trajectory k are pure functions of (seed, k, j), so the device needs no tape and the host can reproduce
// every value bit for bit (this header compiles for both; jsde_generated_jump_tape() is the host side).
//
//   block(attempt)   = Philox4x32-10( counter = (k_lo, k_hi, j, attempt), key = (seed_lo, seed_hi) )   [Salmon et al. 2011]
//   tau_j (attempt 0) = -log(1 - u),  u = rk_double(words 0, 1)            — a unit exponential; exact 1-u, glibc's log
//   xi_j (attempts 1, 2, …) = Marsaglia polar method on (words 0..3) exactly as random_numbers.c:109-130 forms a pair:
//                       x2 from words 0,1, x1 from words 2,3, accepted when 0 < r2 < 1, xi = x1*sqrt(-2*log(r2)/r2)
//
// Arithmetic is IEEE + the bit-exact log restatement of glibc_math.h, so host and device agree to the last bit.
#pragma once

#include "glibc_math.h"

JSDE_FN void jsde_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
	for (int round = 0; round < 10; round++) {
		const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
		const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
		c0 = n0; c1 = n1; c2 = n2; c3 = n3;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// rk_double (random_numbers.c:100-106): 53 random bits in [0, 1), every step exact
JSDE_FN double jsde_unit_double(uint32_t w_hi, uint32_t w_lo)
{
	const double a = (double)(w_hi >> 5), b = (double)(w_lo >> 6);
	return (a * 67108864.0 + b) / 9007199254740992.0;
}

JSDE_FN double jsde_jump_tau(uint64_t seed, int64_t trajectory, int64_t jump)
{
	uint32_t w[4];
	jsde_philox4x32_10((uint32_t)trajectory, (uint32_t)((uint64_t)trajectory >> 32), (uint32_t)jump, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
	const double u = jsde_unit_double(w[0], w[1]);
	return -jsde_log(1.0 - u);
}

JSDE_FN double jsde_jump_xi(uint64_t seed, int64_t trajectory, int64_t jump)
{
	for (uint32_t attempt = 1;; attempt++) {
		uint32_t w[4];
		jsde_philox4x32_10((uint32_t)trajectory, (uint32_t)((uint64_t)trajectory >> 32), (uint32_t)jump, attempt, (uint32_t)seed, (uint32_t)(seed >> 32), w);
		const double x2 = 2.0 * jsde_unit_double(w[0], w[1]) - 1.0;
		const double x1 = 2.0 * jsde_unit_double(w[2], w[3]) - 1.0;
		const double s1 = x1 * x1, s2 = x2 * x2;
		const double r2 = s1 + s2;
		if (r2 < 1.0 && r2 != 0.0) {
			const double m = -2.0 * jsde_log(r2);
			return x1 * sqrt(m / r2);
		}
	}
}
