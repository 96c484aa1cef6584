/* TEST INFRASTRUCTURE (oracle): host evaluation of the counter-based jump variates of jitcsde_b200/csrc/philox_jumps.h,
 * compiled by gcc from the SAME header in its host mode (glibc_math.h's plain-C branch).  The GPU tests compare the
 * device's values with these bit for bit; tests/test_philox.py pins the Philox4x32-10 core against the published
 * known-answer vectors (Random123 kat_vectors). */
#include <math.h>
#include <stdint.h>
#include "../jitcsde_b200/csrc/philox_jumps.h"

void philox_block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
	jsde_philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

void philox_jump_tape(uint64_t seed, int64_t first, int64_t count, int64_t length, double *taus, double *xis)
{
	for (int64_t k = 0; k < count; k++)
		for (int64_t j = 0; j < length; j++) {
			taus[k * length + j] = jsde_jump_tau(seed, first + k, j);
			xis[k * length + j] = jsde_jump_xi(seed, first + k, j);
		}
}
