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

template <int N>
struct WarpRng {
	// room for the pairs left over below one draw's need (N-1) plus a full group, rounded up
	static constexpr int CAP = ((N - 1 + RNG_GROUP + 3) / 4) * 4;
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

	// ---- cooperative refill: K slots of (4 trajectories x 8 candidates) with all global loads issued up front ----------
	// `room` = lanes whose FIFO can take a full group; the first 4K of them (lane order) are served, each exactly once.
	template <int K>
	__device__ __forceinline__ void refill(unsigned room)
	{
		const unsigned full = 0xffffffffu;
		const int g = lane >> 3, j = lane & 7;
		// rank-(4s+g) set bit of `room` for slot s: strip g lowest bits once, then 4 more per slot
		unsigned m = room;
		if (g >= 1) m &= m - 1;
		if (g >= 2) m &= m - 1;
		if (g >= 3) m &= m - 1;
		int Ts[K], c[K];
		bool have[K];
		uint4 a[K], tap[K];
		uint4 *ST[K];
#pragma unroll
		for (int s = 0; s < K; s++) {
			have[s] = m != 0;
			Ts[s] = have[s] ? __ffs(m) - 1 : 0;
			m &= m - 1; m &= m - 1; m &= m - 1; m &= m - 1;
			ST[s] = S0 + (int64_t)Ts[s] * MT_CHUNKS;
			c[s] = mt_wrap(__shfl_sync(full, chunk, Ts[s]) + j);
			a[s] = make_uint4(0, 0, 0, 0);
			tap[s] = make_uint4(0, 0, 0, 0);
		}
#pragma unroll
		for (int s = 0; s < K; s++)
			if (have[s]) {
				a[s] = ST[s][c[s]];
				tap[s] = ST[s][mt_wrap(c[s] + 99)];
			}
		uint32_t edge_next[K], edge_tap[K];
#pragma unroll
		for (int s = 0; s < K; s++) {
			edge_next[s] = 0;
			edge_tap[s] = 0;
			if (have[s] && j == RNG_GROUP - 1) {
				edge_next[s] = reinterpret_cast<const uint32_t *>(ST[s] + mt_wrap(c[s] + 1))[0];
				edge_tap[s] = reinterpret_cast<const uint32_t *>(ST[s] + mt_wrap(c[s] + 100))[0];
			}
		}
		int cnt[K];
#pragma unroll
		for (int s = 0; s < K; s++) {
			const int rdT = __shfl_sync(full, rd, Ts[s]);
			const int lvT = __shfl_sync(full, lvl, Ts[s]);
			uint32_t next = __shfl_down_sync(full, a[s].x, 1);
			uint32_t tap_hi = __shfl_down_sync(full, tap[s].x, 1);
			if (j == RNG_GROUP - 1) {
				next = edge_next[s];
				tap_hi = edge_tap[s];
			}
			bool ok = false;
			double x1 = 0.0, x2 = 0.0, r2 = 0.0;
			if (have[s]) {
				uint4 r;
				r.x = tap[s].y ^ mt_mix(a[s].x, a[s].y);
				r.y = tap[s].z ^ mt_mix(a[s].y, a[s].z);
				r.z = tap[s].w ^ mt_mix(a[s].z, a[s].w);
				r.w = tap_hi ^ mt_mix(a[s].w, next);
				ST[s][c[s]] = r;
				uint4 w;
				w.x = mt_temper(a[s].x);
				w.y = mt_temper(a[s].y);
				w.z = mt_temper(a[s].z);
				w.w = mt_temper(a[s].w);
				ok = polar_candidate(w, x1, x2, r2);
			}
			const unsigned okm = __ballot_sync(full, ok);
			const unsigned grp = (okm >> (g * RNG_GROUP)) & 0xffu;
			const int rank = __popc(grp & ((1u << j) - 1u));
			cnt[s] = __popc(grp);
			if (ok) {
				int slot = rdT + lvT + rank;
				slot = slot >= CAP ? slot - CAP : slot;
				fx1[slot * RNG_STRIDE + Ts[s]] = x1;
				fx2[slot * RNG_STRIDE + Ts[s]] = x2;
				fr2[slot * RNG_STRIDE + Ts[s]] = r2;
			}
		}
		// owners: the lane whose trajectory has rank r among `room` was served by slot r/4, group r%4
		const int below = __popc(room & ((1u << lane) - 1u));
		const bool served = ((room >> lane) & 1u) && below < 4 * K;
		int mine = 0;
#pragma unroll
		for (int s = 0; s < K; s++) {
			const int v = __shfl_sync(full, cnt[s], (below & 3) * RNG_GROUP);
			if ((below >> 2) == s)
				mine = v;
		}
		if (served) {
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
		while (__ballot_sync(full, wants && usable && lvl < need)) {
			const unsigned room = __ballot_sync(full, wants && usable && lvl <= CAP - RNG_GROUP);
			const int n = __popc(room);
			if (n > 12)
				refill<4>(room);
			else if (n > 8)
				refill<3>(room);
			else if (n > 4)
				refill<2>(room);
			else
				refill<1>(room);
		}
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
