// Warp-cooperative Gaussian source: the 32 lanes of a warp jointly advance the Mersenne Twisters of the warp's 32
// trajectories, eight polar candidates of four trajectories per step, and queue the accepted candidates in per-lane
// FIFOs in shared memory.  The stepper then consumes exactly the pairs the reference's sequential stream would give
// (reference jitcsde/random_numbers.c:109-130 via jitcsde/jitced_template.c:145-154).
//
// Why not one generator per lane (round-1 v1, profiles/r01_v1_*): the polar method rejects 21.5 % of candidates, so
// lanes' stream positions drift apart and (a) every 16-byte state access of a warp touches 32 different cache lines,
// (b) a per-lane rejection loop keeps the warp busy for the unluckiest lane (measured: 16 of 32 threads active on
// average).  Here all global accesses to the generator state are 128-byte contiguous per trajectory, every candidate
// is computed exactly once, and the expensive log/div/sqrt run later, convergent, when a lane consumes a pair.
//
// One refill step ("op"): lane l serves trajectory T = (g-th trajectory that has room, g = l/8), candidate j = l%8:
//   loads state chunk c+j and tap chunk c+99+j of T (c = T's stream position), exchanges the two edge words with its
//   neighbour by shuffle, writes the regenerated chunk back (mt19937.cuh explains the incremental twist), tempers,
//   forms the candidate, and — if accepted — stores (x1, x2, r2) at T's FIFO tail + rank (rank by ballot/popc keeps
//   stream order).  A trajectory's position always advances by 8 chunks; surplus pairs stay queued, also across
//   kernel launches (spilled to global memory at kernel end).
#pragma once

#include "gauss.cuh"
#include "mt19937.cuh"

namespace jsde {

constexpr int RNG_GROUP = 8;     // candidates per trajectory per op
constexpr int RNG_STRIDE = 33;   // row stride (doubles) of the FIFO arrays: conflict-free column writes

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <int N>
struct WarpRng {
	// room for the pairs left over below one draw's need (N-1) plus a full group
	static constexpr int CAP = N - 1 + RNG_GROUP;
	static constexpr int SMEM_DOUBLES_PER_WARP = 3 * CAP * RNG_STRIDE;

	double *fx1, *fx2, *fr2;  // this warp's FIFO arrays [CAP][RNG_STRIDE]
	uint4 *S0;                // generator record of the warp's lane-0 trajectory
	int lane;
	int rd, lvl, chunk;       // this lane's FIFO read slot / fill level, its trajectory's next chunk
	bool usable;              // lane owns a real trajectory

	__device__ __forceinline__ WarpRng(double *smem_warp, uint4 *mt, int64_t k_lane0, bool live)
		: fx1(smem_warp), fx2(smem_warp + CAP * RNG_STRIDE), fr2(smem_warp + 2 * CAP * RNG_STRIDE),
		  S0(mt + k_lane0 * MT_CHUNKS), lane(threadIdx.x & 31), rd(0), lvl(0), chunk(0), usable(live) {}

	// ---- persistence across kernel launches: spill [CAP][3][B] + level; read slot is normalised to 0 ---------------
	__device__ __forceinline__ void load(const double *spill, const int *levels, const int *chunks, int64_t B, int64_t k)
	{
		rd = 0;
		lvl = 0;
		if (usable) {
			lvl = levels[k];
			chunk = chunks[k];
			for (int i = 0; i < lvl; i++) {
				fx1[i * RNG_STRIDE + lane] = spill[((int64_t)i * 3 + 0) * B + k];
				fx2[i * RNG_STRIDE + lane] = spill[((int64_t)i * 3 + 1) * B + k];
				fr2[i * RNG_STRIDE + lane] = spill[((int64_t)i * 3 + 2) * B + k];
			}
		}
		__syncwarp();
	}

	__device__ __forceinline__ void store(double *spill, int *levels, int *chunks, int64_t B, int64_t k) const
	{
		__syncwarp();
		if (usable) {
			levels[k] = lvl;
			chunks[k] = chunk;
			for (int i = 0; i < lvl; i++) {
				int s = rd + i;
				s = s >= CAP ? s - CAP : s;
				spill[((int64_t)i * 3 + 0) * B + k] = fx1[s * RNG_STRIDE + lane];
				spill[((int64_t)i * 3 + 1) * B + k] = fx2[s * RNG_STRIDE + lane];
				spill[((int64_t)i * 3 + 2) * B + k] = fr2[s * RNG_STRIDE + lane];
			}
		}
	}

	// ---- cooperative refill ------------------------------------------------------------------------------------------
	// `room` = lanes whose FIFO can take a full group.  They are served four at a time ("slot": 4 trajectories x 8
	// candidates), each exactly once.  Pass 1 only prefetches every slot's state lines into L1 so that the memory
	// latency is paid once per refill, not once per slot; pass 2 does the work.  Both are run-time loops: unrolling
	// them blew the 32 KB instruction cache (profiles/r01_v3_*: 40 % of stall samples were instruction fetch).
	__device__ __forceinline__ void refill(unsigned room)
	{
		const unsigned full = 0xffffffffu;
		const int g = lane >> 3, j = lane & 7;
		const int slots = (__popc(room) + 3) >> 2;
		// rank-(4s+g) set bit of `room` for slot s: strip g lowest bits once, then 4 more per slot
		unsigned m0 = room;
		if (g >= 1) m0 &= m0 - 1;
		if (g >= 2) m0 &= m0 - 1;
		if (g >= 3) m0 &= m0 - 1;
		if (slots > 1) {
			unsigned m = m0;
#pragma unroll 1
			for (int s = 0; s < slots; s++) {
				const bool have = m != 0;
				const int T = have ? __ffs(m) - 1 : 0;
				m &= m - 1; m &= m - 1; m &= m - 1; m &= m - 1;
				const int c = mt_wrap(__shfl_sync(full, chunk, T) + j);
				if (have) {
					const uint4 *ST = S0 + (int64_t)T * MT_CHUNKS;
					prefetch_l1(ST + c);
					prefetch_l1(ST + mt_wrap(c + 99));
					if (j == RNG_GROUP - 1) {
						prefetch_l1(ST + mt_wrap(c + 1));
						prefetch_l1(ST + mt_wrap(c + 100));
					}
				}
			}
		}
		const int below = __popc(room & ((1u << lane) - 1u));
		int mine = 0;
		unsigned m = m0;
#pragma unroll 1
		for (int s = 0; s < slots; s++) {
			const bool have = m != 0;
			const int T = have ? __ffs(m) - 1 : 0;
			m &= m - 1; m &= m - 1; m &= m - 1; m &= m - 1;
			uint4 *ST = S0 + (int64_t)T * MT_CHUNKS;
			const int c = mt_wrap(__shfl_sync(full, chunk, T) + j);
			const int rdT = __shfl_sync(full, rd, T);
			const int lvT = __shfl_sync(full, lvl, T);
			uint4 a = make_uint4(0, 0, 0, 0), tap = make_uint4(0, 0, 0, 0);
			uint32_t edge_next = 0, edge_tap = 0;
			if (have) {
				a = ST[c];
				tap = ST[mt_wrap(c + 99)];
				if (j == RNG_GROUP - 1) {
					edge_next = reinterpret_cast<const uint32_t *>(ST + mt_wrap(c + 1))[0];
					edge_tap = reinterpret_cast<const uint32_t *>(ST + mt_wrap(c + 100))[0];
				}
			}
			uint32_t next = __shfl_down_sync(full, a.x, 1);
			uint32_t tap_hi = __shfl_down_sync(full, tap.x, 1);
			if (j == RNG_GROUP - 1) {
				next = edge_next;
				tap_hi = edge_tap;
			}
			bool ok = false;
			double x1 = 0.0, x2 = 0.0, r2 = 0.0;
			if (have) {
				uint4 r;
				r.x = tap.y ^ mt_mix(a.x, a.y);
				r.y = tap.z ^ mt_mix(a.y, a.z);
				r.z = tap.w ^ mt_mix(a.z, a.w);
				r.w = tap_hi ^ mt_mix(a.w, next);
				ST[c] = r;
				a.x = mt_temper(a.x);
				a.y = mt_temper(a.y);
				a.z = mt_temper(a.z);
				a.w = mt_temper(a.w);
				ok = polar_candidate(a, x1, x2, r2);
			}
			const unsigned okm = __ballot_sync(full, ok);
			const unsigned grp = (okm >> (g * RNG_GROUP)) & 0xffu;
			const int rank = __popc(grp & ((1u << j) - 1u));
			const int cnt = __popc(grp);
			if (ok) {
				int slot = rdT + lvT + rank;
				slot = slot >= CAP ? slot - CAP : slot;
				fx1[slot * RNG_STRIDE + T] = x1;
				fx2[slot * RNG_STRIDE + T] = x2;
				fr2[slot * RNG_STRIDE + T] = r2;
			}
			// owner: the lane whose trajectory has rank r among `room` is served by slot r/4, group r%4
			const int v = __shfl_sync(full, cnt, (below & 3) * RNG_GROUP);
			if ((below >> 2) == s)
				mine = v;
		}
		if ((room >> lane) & 1u) {
			lvl += mine;
			chunk = mt_wrap(chunk + RNG_GROUP);
		}
		__syncwarp();
	}

	// Make sure every lane with `wants` holds at least `need` pairs.  Whole warp must call (convergent).
	// Everybody with room is served, not only the needy: every slot does useful work as long as it is full.
	__device__ __forceinline__ void ensure(bool wants, int need)
	{
		const unsigned full = 0xffffffffu;
		while (__ballot_sync(full, wants && usable && lvl < need))
			refill(__ballot_sync(full, wants && usable && lvl <= CAP - RNG_GROUP));
	}

	// Pop this lane's next pair (caller guarantees lvl >= 1 via ensure()).
	__device__ __forceinline__ void pop(double &x1, double &x2, double &r2)
	{
		x1 = fx1[rd * RNG_STRIDE + lane];
		x2 = fx2[rd * RNG_STRIDE + lane];
		r2 = fr2[rd * RNG_STRIDE + lane];
		rd = rd + 1 >= CAP ? 0 : rd + 1;
		lvl--;
	}
};

}  // namespace jsde
