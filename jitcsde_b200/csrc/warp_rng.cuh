// Warp-cooperative Gaussian source: the 32 lanes of a warp jointly advance the Mersenne Twisters of the warp's 32
// trajectories and queue the accepted polar candidates (x1, x2, r2) per trajectory.  The stepper then consumes exactly
// the pairs the reference's sequential stream would give (reference jitcsde/random_numbers.c:109-130 via
// jitcsde/jitced_template.c:145-154).
//
// Why not one generator per lane (round-1 v1, profiles/r01_v1_*): the polar method rejects 21.5 % of candidates, so
// lanes' stream positions drift apart and (a) every 16-byte state access of a warp touches 32 different cache lines,
// (b) a per-lane rejection loop keeps the warp busy for the unluckiest lane (measured: 16 of 32 threads active on
// average).  Here all global accesses to the generator state are contiguous per trajectory, every candidate is computed
// exactly once, and the expensive log/div/sqrt run later, convergent, when a lane consumes a pair.
//
// PairRing<N>: the queue of a trajectory is a ring of RING pairs in global memory, [3][RING] doubles (x1s, x2s, r2s),
// addressed by (rng_cons, rng_level) — it survives between launches without any copy, and during a launch only the
// rings of the resident trajectories are touched (they stay in L2).  Two users share it:
//   * the fused integrate kernel (duo.cuh): a producer warp regenerates 32 chunks of ONE trajectory per service (uniform
//     control flow, generator words staged by bulk copies) and appends the accepted pairs; it then hands every consumer
//     lane its next pairs through a small lane-private FIFO in shared memory;
//   * this file's WarpRng, for the single-step kernels (get_next_step, pin_noise, the generator debug stream), where the
//     warp that consumes the pairs also produces them: eight candidates of four trajectories per step ("slot"), pops
//     straight from the ring.  Not performance-critical.
#pragma once

#include "gauss.cuh"
#include "mt19937.cuh"

namespace jsde {

constexpr int RNG_GROUP = 8;        // WarpRng: candidates per trajectory and slot
constexpr int SERVICE_CHUNKS = 32;  // duo.cuh: chunks (= candidates) per service of one trajectory, one per lane

template <int N>
struct PairRing {
	static constexpr int L0 = 2 * N;   // duo.cuh: pairs per lane in the shared-memory FIFO (two draws)
	static constexpr int LOW = 2 * N;  // duo.cuh: a trajectory is served when fewer pairs than this wait in its ring
	// a service appends up to SERVICE_CHUNKS pairs while at most L0 (handed over, perhaps unread) + LOW - 1 are queued
	static constexpr int RING = (L0 + LOW - 1 + SERVICE_CHUNKS + 1) & ~1;
	static constexpr int DOUBLES = 3 * RING;  // per trajectory
	static __device__ __forceinline__ int wrap(int i) { return i >= RING ? i - RING : i; }
};

// loads that bypass L1: ring entries are written by other lanes of the same warp / block shortly before they are read
__device__ __forceinline__ double ld_ring(const double *p) { return __ldcg(p); }

template <int N>
struct WarpRng {
	using R = PairRing<N>;
	// shared memory of one warp: order_()[32] bytes (scratch list of the lanes being served)
	static constexpr int SMEM_DOUBLES_PER_WARP = 4;

	unsigned char *order;     // this warp's scratch
	const double *log_tab;    // block-shared copy of glibc's log table {invc, logc}[128]
	uint4 *S0;                // generator record of the warp's lane-0 trajectory
	double *ring0;            // pair ring of the warp's lane-0 trajectory
	int lane;
	int cons, lvl, chunk;     // this lane's ring read index / fill level, its trajectory's next chunk
	bool usable;              // lane owns a real trajectory

	__device__ __forceinline__ WarpRng(double *smem_warp, const double *shared_log_tab, uint4 *mt, double *rings, int64_t k_lane0, bool live)
		: order(reinterpret_cast<unsigned char *>(smem_warp)), log_tab(shared_log_tab), S0(mt + k_lane0 * MT_CHUNKS),
		  ring0(rings + k_lane0 * R::DOUBLES), lane(threadIdx.x & 31), cons(0), lvl(0), chunk(0), usable(live) {}

	// the single-step kernels never run the controller; constant-memory tables keep Lane<> instantiable with this source
	__device__ __forceinline__ jsde_pow_tables pow_tables() const { return jsde_pow_tables{JSDE_POW_INVC, JSDE_POW_LOGC, JSDE_POW_LOGCTAIL, JSDE_EXP_TAB}; }

	__device__ __forceinline__ void load(const int *cons_, const int *levels, const int *chunks, int64_t k)
	{
		cons = lvl = 0;
		if (usable) {
			cons = cons_[k];
			lvl = levels[k];
			chunk = chunks[k];
		}
		__syncwarp();
	}

	__device__ __forceinline__ void store(int *cons_, int *levels, int *chunks, int64_t k) const
	{
		__syncwarp();
		if (usable) {
			cons_[k] = cons;
			levels[k] = lvl;
			chunks[k] = chunk;
		}
	}

	// ---- cooperative refill ------------------------------------------------------------------------------------------
	// `room` = lanes whose ring can take a full group.  They are served four at a time ("slot": 4 trajectories x 8
	// candidates), each exactly once.  Lane (g, j) = (lane/8, lane%8) handles candidate j of the g-th trajectory of the
	// slot: state chunk c+j and tap chunk c+99+j (mt19937.cuh), edge words from the neighbouring lane.
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
			order[below] = (unsigned char)lane;  // compacted list of the lanes to serve
		__syncwarp();
		int mine = 0;
#pragma unroll 1
		for (int s = 0; s < slots; s++) {
			const int rank = 4 * s + g;
			const bool have = rank < nserved;
			const int T = have ? order[rank] : 0;
			const int c = mt_wrap(__shfl_sync(full, chunk, T) + j);
			const int tail = R::wrap(__shfl_sync(full, cons, T) + __shfl_sync(full, lvl, T));  // cons < RING, lvl <= RING - 8
			uint4 a = make_uint4(0, 0, 0, 0), tap = make_uint4(0, 0, 0, 0);
			uint32_t edge_next = 0, edge_tap = 0;
			uint4 *ST = S0 + (int64_t)T * MT_CHUNKS;
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
				uint4 v;
				v.x = mt_temper(a.x);
				v.y = mt_temper(a.y);
				v.z = mt_temper(a.z);
				v.w = mt_temper(a.w);
				ok = polar_candidate(v, x1, x2, r2);
			}
			const unsigned okm = __ballot_sync(full, ok);
			const unsigned grp = (okm >> (g * RNG_GROUP)) & 0xffu;
			if (ok) {
				const int at = R::wrap(tail + __popc(grp & ((1u << j) - 1u)));
				double *ring = ring0 + (int64_t)T * R::DOUBLES;
				ring[at] = x1;
				ring[R::RING + at] = x2;
				ring[2 * R::RING + at] = r2;
			}
			// owner: the lane whose trajectory has rank r among `room` was served by slot r/4, group r%4
			if ((below >> 2) == s)
				mine = __popc((okm >> ((below & 3) * RNG_GROUP)) & 0xffu);
		}
		if (mine_served) {
			lvl += mine;
			chunk = mt_wrap(chunk + RNG_GROUP);
		}
		__threadfence_block();
		__syncwarp();
	}

	// Make sure every lane with `wants` holds at least `need` pairs (need <= 2N).  Whole warp must call (convergent).
	// Everybody with room is served, not only the needy: every slot does useful work as long as it is full.
	// Returns whether a refill happened (warp-uniform).
	__device__ __forceinline__ bool ensure(bool wants, int need)
	{
		const unsigned full = 0xffffffffu;
		bool refilled = false;
		while (__ballot_sync(full, wants && usable && lvl < need)) {
			refill(__ballot_sync(full, wants && usable && lvl <= R::RING - RNG_GROUP));
			refilled = true;
		}
		return refilled;
	}

	// Pop this lane's next pair (caller guarantees lvl >= 1 via ensure()).
	__device__ __forceinline__ void pop(double &x1, double &x2, double &r2)
	{
		const double *ring = ring0 + (int64_t)lane * R::DOUBLES;
		x1 = ld_ring(ring + cons);
		x2 = ld_ring(ring + R::RING + cons);
		r2 = ld_ring(ring + 2 * R::RING + cons);
		cons = R::wrap(cons + 1);
		lvl--;
	}
};

}  // namespace jsde
