// Per-trajectory Mersenne Twister with an *incremental* twist.
//
// Follows reference jitcsde/random_numbers.c:41-97 (rk_seed, rk_random).  The reference regenerates all 624 words
// in one batch whenever `pos == 624`.  A lane cannot afford a 624-iteration detour while its 31 neighbours wait,
// so the state array S is kept in the equivalent "regenerate as you consume" form:
//
//     S[i] always holds the word that will be returned the next time index i is consumed;
//     consuming S[i] immediately replaces it by   S[(i+397)%624] ^ mix(S[i], S[(i+1)%624]).
//
// In consumption order this reads exactly the operands the batch loop (random_numbers.c:73-84) reads: for i < 227
// the tap i+397 has not been consumed yet (old block), for i >= 227 the tap i-227 has already been replaced (new
// block), S[i+1] is still the old block, and for i = 623 the neighbour S[0] is already the new block — the three
// cases of the reference's loop.  Seeding therefore performs rk_seed's Knuth recurrence followed by one full pass of
// the replacement rule, after which `chunk = 0` corresponds to the reference's state right after its first twist.
//
// Every Gaussian candidate consumes exactly four words (random_numbers.c:100-121), so the unit of consumption is a
// 16-byte chunk of four consecutive words; 624 = 4 * 156 and 397 = 4 * 99 + 1, hence the tap words of chunk c are
// words 1..3 of chunk c+99 and word 0 of chunk c+100 (indices mod 156).
#pragma once

#include <stdint.h>

namespace jsde {

constexpr int MT_WORDS = 624;
constexpr int MT_CHUNKS = 156;

__device__ __forceinline__ uint32_t mt_mix(uint32_t a, uint32_t b)
{
	const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
	return (y >> 1) ^ ((b & 1u) ? 0x9908b0dfu : 0u);
}

// tap ^ mt_mix(a, b) in five instructions (bit select, shift, bit test, multiply on the otherwise idle FMA pipe,
// three-input xor) instead of the seven the compiler makes of the expression above
__device__ __forceinline__ uint32_t mt_twist(uint32_t a, uint32_t b, uint32_t tap)
{
	uint32_t y;
	asm("lop3.b32 %0, %1, %2, 0x80000000, 0xE4;" : "=r"(y) : "r"(a), "r"(b));  // (mask & a) | (~mask & b)
	const uint32_t mag = (b & 1u) * 0x9908b0dfu;
	return (y >> 1) ^ mag ^ tap;
}


__device__ __forceinline__ uint32_t mt_temper(uint32_t y)
{
	y ^= y >> 11;
	y ^= (y << 7) & 0x9d2c5680u;
	y ^= (y << 15) & 0xefc60000u;
	y ^= y >> 18;
	return y;
}

// c mod 156 for c in [0, 312): unsigned minimum of c and c - 156 (which wraps around to a huge value when c < 156)
__device__ __forceinline__ int mt_wrap(int c) { return (int)min((unsigned)c, (unsigned)c - 156u); }

// Consume chunk `c` of the trajectory whose state starts at `S` (156 uint4): returns the four *untempered* words and
// writes their replacements.
__device__ __forceinline__ uint4 mt_consume_chunk(uint4 *S, int c)
{
	const uint4 a = S[c];
	const uint32_t next = reinterpret_cast<const uint32_t *>(S + mt_wrap(c + 1))[0];
	const uint4 tap = S[mt_wrap(c + 99)];
	const uint32_t tap_hi = reinterpret_cast<const uint32_t *>(S + mt_wrap(c + 100))[0];
	uint4 r;
	r.x = tap.y ^ mt_mix(a.x, a.y);
	r.y = tap.z ^ mt_mix(a.y, a.z);
	r.z = tap.w ^ mt_mix(a.z, a.w);
	r.w = tap_hi ^ mt_mix(a.w, next);
	S[c] = r;
	return a;
}

// rk_seed (random_numbers.c:41-52) followed by the first regeneration pass.
__device__ inline void mt_seed(uint4 *S, uint32_t seed)
{
	uint32_t *w = reinterpret_cast<uint32_t *>(S);
	uint32_t v = seed;
	for (int i = 0; i < MT_WORDS; i++) {
		w[i] = v;
		v = 1812433253u * (v ^ (v >> 30)) + (uint32_t)i + 1u;
	}
	for (int c = 0; c < MT_CHUNKS; c++)
		(void)mt_consume_chunk(S, c);
}

}  // namespace jsde
