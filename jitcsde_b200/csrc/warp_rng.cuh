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
// This class serves the single-step kernels (get_next_step, pin_noise, the generator debug stream), where the warp
// that consumes the pairs also produces them.  The fused integrate kernel splits the two roles between warps
// (duo.cuh) and shares this file's vocabulary (slot layout, FIFO row stride, spill format).
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
// Row stride (doubles) of the FIFO arrays [slot][lane].  Element (s, T) lies in bank pair (s + T) mod 16: the producer's
// compacted writes of ONE trajectory (same T, consecutive slots) never collide; the writes of the four trajectories of a
// slot can, and so can the consumer's pops when the lanes' read slots differ (ncu, profiles/r02_headline_sra_sliced_*:
// pops 2.4 wavefronts per instruction instead of 2, FIFO writes 1.8 times the ideal).  A stride of 32 makes the pops
// conflict-free and the writes eight-way conflicted: measured slower (profiles/r02_fifo_stride_ab.log).  Shared-memory
// wavefronts are not what limits the kernel (LSU data pipe 53 % busy, issue slots 69 %).
#ifndef JSDE_RNG_STRIDE
#define JSDE_RNG_STRIDE 33
#endif
constexpr int RNG_STRIDE = JSDE_RNG_STRIDE;

// 16-byte asynchronous copy global -> shared (LDGSTS): no register, no scoreboard wait until cp.async.wait_group
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
	const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }

constexpr int RNG_STAGE_CHUNKS = 18;   // staged words of one trajectory and slot (duo.cuh): chunks c..c+8 and c+99..c+107

// TAIL: the queue's third value is the scale-independent Gaussian factor s = sqrt(-2 log(r2)/r2) instead of r2 — the
// generator side computes it, the stepper only multiplies by the scale (f = scale*s: the same bits as
// random_numbers.c:127-129).  Chosen per model (kernels.cuh: rng_tail): for multiplicative-noise models the stepper is
// the longer of the two roles in the fused kernel, and the producer warps have time to spare.
template <int N, bool TAIL_ = false>
struct WarpRng {
	static constexpr bool TAIL = TAIL_;
	// room for the pairs left over below two draws' need (2N-1) plus a full group: the producer warps of the fused
	// integrate kernel (duo.cuh) keep every trajectory one draw ahead of its consumer; the spill format is shared
	static constexpr int CAP = 2 * N - 1 + RNG_GROUP;
	static constexpr int FIFO_DOUBLES = (3 * CAP * RNG_STRIDE + 1) & ~1;
	// FIFOs | order_()[32] bytes
	static constexpr int SMEM_DOUBLES_PER_WARP = FIFO_DOUBLES + 4;

	double *fx1;              // this warp's shared memory: FIFO arrays x1, x2, r2, each [CAP][RNG_STRIDE], then ...
	const double *log_tab;    // block-shared copy of glibc's log table {invc, logc}[128]
	__device__ __forceinline__ double *fx2_() const { return fx1 + CAP * RNG_STRIDE; }
	__device__ __forceinline__ double *fr2_() const { return fx1 + 2 * CAP * RNG_STRIDE; }
	// ... order_()[32]: scratch list of lanes being served
	__device__ __forceinline__ unsigned char *order_() const { return reinterpret_cast<unsigned char *>(fx1 + FIFO_DOUBLES); }
	uint4 *S0;                // generator record of the warp's lane-0 trajectory
	int lane;
	int rd, lvl, chunk;       // this lane's FIFO read slot / fill level, its trajectory's next chunk
	bool usable;              // lane owns a real trajectory

	__device__ __forceinline__ WarpRng(double *smem_warp, const double *shared_log_tab, uint4 *mt, int64_t k_lane0, bool live)
		: fx1(smem_warp), log_tab(shared_log_tab), S0(mt + k_lane0 * MT_CHUNKS), lane(threadIdx.x & 31), rd(0), lvl(0), chunk(0), usable(live) {}

	// the single-step kernels never run the controller; constant-memory tables keep Lane<> instantiable with this source
	__device__ __forceinline__ jsde_pow_tables pow_tables() const { return jsde_pow_tables{JSDE_POW_INVC, JSDE_POW_LOGC, JSDE_POW_LOGCTAIL, JSDE_EXP_TAB}; }

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
				fx2_()[i * RNG_STRIDE + lane] = spill[((int64_t)i * 3 + 1) * B + k];
				fr2_()[i * RNG_STRIDE + lane] = spill[((int64_t)i * 3 + 2) * B + k];
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
				spill[((int64_t)i * 3 + 1) * B + k] = fx2_()[s * RNG_STRIDE + lane];
				spill[((int64_t)i * 3 + 2) * B + k] = fr2_()[s * RNG_STRIDE + lane];
			}
		}
	}

	// ---- cooperative refill ------------------------------------------------------------------------------------------
	// `room` = lanes whose FIFO can take a full group.  They are served four at a time ("slot": 4 trajectories x 8
	// candidates), each exactly once, in one run-time loop (unrolled variants blew the 32 KB instruction cache,
	// profiles/r01_v3_*).
	struct SlotWords {
		uint4 a, tap;
		uint32_t edge_next, edge_tap;
		int T, c;
		bool have;
	};

	__device__ __forceinline__ int served_lane(int rank) const { return order_()[rank]; }

	__device__ __forceinline__ SlotWords fetch(int s, int nserved, int g, int j) const
	{
		SlotWords w;
		const int rank = 4 * s + g;
		w.have = rank < nserved;
		w.T = w.have ? order_()[rank] : 0;
		w.c = mt_wrap(__shfl_sync(0xffffffffu, chunk, w.T) + j);
		w.a = make_uint4(0, 0, 0, 0);
		w.tap = make_uint4(0, 0, 0, 0);
		w.edge_next = 0;
		w.edge_tap = 0;
		if (w.have) {
			const uint4 *ST = S0 + w.T * MT_CHUNKS;
			w.a = ST[w.c];
			w.tap = ST[mt_wrap(w.c + 99)];
			if (j == RNG_GROUP - 1) {
				w.edge_next = reinterpret_cast<const uint32_t *>(ST + mt_wrap(w.c + 1))[0];
				w.edge_tap = reinterpret_cast<const uint32_t *>(ST + mt_wrap(w.c + 100))[0];
			}
		}
		return w;
	}

	__device__ __forceinline__ void refill(unsigned room)
	{
		const unsigned full = 0xffffffffu;
		const int g = lane >> 3, j = lane & 7;
		__syncwarp();
		const int nserved = __popc(room);
		const int slots = (nserved + 3) >> 2;
		const unsigned lt = (1u << lane) - 1u;
		const int below = __popc(room & lt);
		const bool mine_served = (room >> lane) & 1u;
		if (mine_served)
			order_()[below] = (unsigned char)lane;  // compacted list of the lanes to serve
		__syncwarp();
		int mine = 0;
#pragma unroll 1
		for (int s = 0; s < slots; s++) {
			const SlotWords w = fetch(s, nserved, g, j);
			const int rdT = __shfl_sync(full, rd, w.T);
			const int lvT = __shfl_sync(full, lvl, w.T);
			uint32_t next = __shfl_down_sync(full, w.a.x, 1);
			uint32_t tap_hi = __shfl_down_sync(full, w.tap.x, 1);
			if (j == RNG_GROUP - 1) {
				next = w.edge_next;
				tap_hi = w.edge_tap;
			}
			bool ok = false;
			double x1 = 0.0, x2 = 0.0, r2 = 0.0;
			if (w.have) {
				uint4 r;
				r.x = w.tap.y ^ mt_mix(w.a.x, w.a.y);
				r.y = w.tap.z ^ mt_mix(w.a.y, w.a.z);
				r.z = w.tap.w ^ mt_mix(w.a.z, w.a.w);
				r.w = tap_hi ^ mt_mix(w.a.w, next);
				S0[w.T * MT_CHUNKS + w.c] = r;
				uint4 v;
				v.x = mt_temper(w.a.x);
				v.y = mt_temper(w.a.y);
				v.z = mt_temper(w.a.z);
				v.w = mt_temper(w.a.w);
				ok = polar_candidate(v, x1, x2, r2);
				if (TAIL && ok)
					r2 = polar_radius_factor_with(r2, log_tab);
			}
			const unsigned okm = __ballot_sync(full, ok);
			const unsigned grp = (okm >> (g * RNG_GROUP)) & 0xffu;
			if (ok) {
				int slot = rdT + lvT + __popc(grp & ((1u << j) - 1u));
				slot = slot >= CAP ? slot - CAP : slot;
				const int at = slot * RNG_STRIDE + w.T;
				fx1[at] = x1;
				fx2_()[at] = x2;
				fr2_()[at] = r2;
			}
			// owner: the lane whose trajectory has rank r among `room` was served by slot r/4, group r%4
			if ((below >> 2) == s)
				mine = __popc((okm >> ((below & 3) * RNG_GROUP)) & 0xffu);
		}
		if (mine_served) {
			lvl += mine;
			chunk = mt_wrap(chunk + RNG_GROUP);
		}
		__syncwarp();
	}

	// Make sure every lane with `wants` holds at least `need` pairs.  Whole warp must call (convergent).
	// Everybody with room is served, not only the needy: every slot does useful work as long as it is full.
	// Returns whether a refill happened (warp-uniform).
	__device__ __forceinline__ bool ensure(bool wants, int need)
	{
		const unsigned full = 0xffffffffu;
		bool refilled = false;
		while (__ballot_sync(full, wants && usable && lvl < need)) {
			refill(__ballot_sync(full, wants && usable && lvl <= CAP - RNG_GROUP));
			refilled = true;
		}
		return refilled;
	}

	// Pop this lane's next pair (caller guarantees lvl >= 1 via ensure()).
	__device__ __forceinline__ void pop(double &x1, double &x2, double &r2)
	{
		x1 = fx1[rd * RNG_STRIDE + lane];
		x2 = fx2_()[rd * RNG_STRIDE + lane];
		r2 = fr2_()[rd * RNG_STRIDE + lane];
		rd = rd + 1 >= CAP ? 0 : rd + 1;
		lvl--;
	}
};

}  // namespace jsde
