// Producer / consumer warp pairs for the fused integrate kernel.
//
// A 256-thread block holds four *consumer* warps (warpgroup 0: one trajectory per lane, runs the stepper of
// stepper.cuh) and four *producer* warps (warpgroup 1).  Producer warp p advances the Mersenne Twisters of the 32
// trajectories of consumer warp p with the cooperative scheme of warp_rng.cuh (eight polar candidates of four
// trajectories per step, 128-byte contiguous accesses to the generator state, accepted candidates compacted in
// stream order into per-trajectory FIFOs in shared memory) — but in its own warp, with its own (small) register
// budget, concurrently with the consumer's floating-point work:
//
//   * the generator's integer instructions and the stepper's FP64 dependency chains come from different warps, so the
//     scheduler interleaves them (round 1, profiles/r01_v12_*: a warp that did both issued on 60 % of cycles with
//     2.2 cycles of fixed-latency stall per instruction and only 16 warps per SM because of 128 registers/thread);
//   * `setmaxnreg` moves registers from the producers to the consumers: 24 warps per SM instead of 16;
//   * the generator's HBM latency is hidden by the other warps and by staging the state words of the next two slots
//     with cp.async while one slot is processed — no prediction of the next refill (measured: an additional L2
//     prefetch one round ahead, or a deeper staging pipeline, changes nothing).
//
// Protocol (per pair, one named barrier of 64 threads per round, control words double-buffered by round parity):
//
//   consumer, before barrier r:  popped[r&1][lane] = pairs popped so far;  mask[r&1] = lanes that still have work
//   producer, after  barrier r:  level = pushed - popped (a lower bound of the true level during round r);
//                                refill until every lane in mask has level >= 2N
//   consumer, during round r:    pops at most N pairs  ->  at barrier r+1 every lane still holds >= N  ->  round r+1 is
//                                safe while the producer tops up again.  Round 0 of the consumer is empty.
//
// The producer writes slots [wr, wr+added) of a trajectory's ring while the consumer reads [rd, rd+N); they are
// disjoint because `added` is bounded by the room computed from the *published* (older, smaller) popped count.
// Both loops leave on the same published mask == 0.  The producer owns the persistence of queued pairs between
// launches (same spill format as WarpRng, which the single-step kernels use).
//
// Trajectory recycling (kernels.cuh, RECYCLE): the grid is at most one wave of blocks; a lane whose trajectory is done
// takes the next one from a pool instead of idling until its warp's and block's slowest trajectory
// finishes (+26 % on SRI-Lorenz, whose step counts differ by tens of per cent between trajectories).  The consumer
// announces the change through `swap[lane]` together with the final popped count of the old trajectory; the producer
// then spills the old trajectory's queued pairs and stream position, loads the new one's, and refills — the lane sits
// out that one round.  Every lane's first trajectory arrives the same way (round 0).  `swap = -2` releases the current
// trajectory without taking another (first half of a yield: kernels.cuh).  The published mask holds the lanes that keep the
// warp pair alive: lanes that step, that have a trajectory to hand back, or that are in the middle of a yield.
#pragma once

#include "warp_rng.cuh"

namespace jsde {

constexpr int DUO_PAIRS = 4;
constexpr int DUO_BLOCK = 64 * DUO_PAIRS;
// Resident blocks per SM and the register split between the warpgroups (the launch allocation 65536/(256*blocks),
// rounded down to a multiple of 8, is redistributed by setmaxnreg: consumer + producer = 2 x allocation).  Up to three
// components the stepper fits 104 registers (SRA) and three blocks are resident; beyond that it would spill, so two
// blocks with 168/88.  The SRI stepper is 2.4 times the generator's work, so its producers wait most of the time and
// can live with spills: 112/48 (SRI-Lorenz +11 %, Duffing +1.5 %), and with three components 120/40 (SRI-Lorenz another
// +5 %; Duffing -1 %, so two components stay at 112/48) or, when the stepper also carries the jump code, 128/32 (jumpy
// Lorenz +9 %) — profiles/r02_register_split_ab.log; for SRA, where both roles are equally loaded, 112/48
// and the two-block splits are 5-8 % slower.  One-component SRI models (geometric Brownian motion) have a single serial
// FP64 chain per lane and fit 88 registers without spilling: four blocks with 88/40 (+10 %, profiles/r02_blocks4_ab.log;
// two components already spill there: Duffing -3.5 %).  The macros override.
template <int N, bool ADDITIVE, bool JUMP = false>
struct DuoConfig {
#ifdef JSDE_DUO_MIN_BLOCKS
	static constexpr int BLOCKS = JSDE_DUO_MIN_BLOCKS;
#else
	static constexpr int BLOCKS = (N == 1 && !ADDITIVE) ? 4 : N <= 3 ? 3 : 2;
#endif
#ifdef JSDE_DUO_CONSUMER_REGS
	static constexpr int CONSUMER_REGS = JSDE_DUO_CONSUMER_REGS;
#else
	static constexpr int CONSUMER_REGS = (N == 1 && !ADDITIVE) ? 88 : N <= 3 ? (ADDITIVE ? 104 : N == 3 ? (JUMP ? 128 : 120) : 112) : 168;
#endif
#ifdef JSDE_DUO_PRODUCER_REGS
	static constexpr int PRODUCER_REGS = JSDE_DUO_PRODUCER_REGS;
#else
	static constexpr int PRODUCER_REGS = (N == 1 && !ADDITIVE) ? 40 : N <= 3 ? (ADDITIVE ? 56 : N == 3 ? (JUMP ? 32 : 40) : 48) : 88;
#endif
};
#ifndef JSDE_DUO_STAGE_AHEAD
#define JSDE_DUO_STAGE_AHEAD 2  // slots whose generator words are in flight (cp.async) ahead of the one being processed
#endif

__device__ __forceinline__ void pair_barrier(int pair)
{
	asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory");  // barrier 0 is __syncthreads()
}

// glibc's log table {invc, logc}[128] and pow tables invc[128], logc[128], logctail[128], exp_tab[256] in shared memory:
// a lane-per-trajectory caller looks up 32 different entries at once.  Whole block must call.
constexpr int MATH_TAB_DOUBLES = 256 + 3 * 128 + 256;
__device__ __forceinline__ void stage_math_tables(double *dst)
{
	for (int i = threadIdx.x; i < 128; i += blockDim.x) {
		dst[2 * i] = JSDE_LOG_INVC[i];
		dst[2 * i + 1] = JSDE_LOG_LOGC[i];
		dst[256 + i] = JSDE_POW_INVC[i];
		dst[384 + i] = JSDE_POW_LOGC[i];
		dst[512 + i] = JSDE_POW_LOGCTAIL[i];
	}
	for (int i = threadIdx.x; i < 256; i += blockDim.x)
		dst[640 + i] = __longlong_as_double((long long)JSDE_EXP_TAB[i]);
	__syncthreads();
}

template <int N>
struct PairLayout {
	static constexpr int CAP = WarpRng<N>::CAP;
	static constexpr int FIFO_DOUBLES = 3 * CAP * RNG_STRIDE;
	// FIFOs x1, x2, r2 [CAP][RNG_STRIDE] | popped[2][32] u32 | mask[2] u32 | order[32] u8 | swap[2][32] i32 | stage[3][4][18] uint4
	static constexpr int STAGE_DEPTH = JSDE_DUO_STAGE_AHEAD + 1;  // slots whose generator words are in flight (cp.async) or being processed
	static constexpr int STAGE_OFFSET = (FIFO_DOUBLES + 32 + 1 + 4 + 32 + 1) & ~1;  // 16-byte aligned
	static constexpr int DOUBLES = STAGE_OFFSET + STAGE_DEPTH * 4 * RNG_STAGE_CHUNKS * 2;
	static __device__ __forceinline__ uint4 *stage(double *base, int buf, int g) { return reinterpret_cast<uint4 *>(base + STAGE_OFFSET) + (buf * 4 + g) * RNG_STAGE_CHUNKS; }
	static __device__ __forceinline__ unsigned *popped(double *base, int parity) { return reinterpret_cast<unsigned *>(base + FIFO_DOUBLES) + 32 * parity; }
	static __device__ __forceinline__ unsigned *mask(double *base, int parity) { return reinterpret_cast<unsigned *>(base + FIFO_DOUBLES + 32) + parity; }
	static __device__ __forceinline__ unsigned char *order(double *base) { return reinterpret_cast<unsigned char *>(base + FIFO_DOUBLES + 33); }
	// swap[2][32] i32: trajectory a lane switches to in this round (-1: none)
	static __device__ __forceinline__ int *swap(double *base, int parity) { return reinterpret_cast<int *>(base + FIFO_DOUBLES + 37) + 32 * parity; }
};

// ---- consumer side: what Lane<> needs from its pair source ---------------------------------------------------------
template <int N, bool TAIL_ = false>
struct PairConsumer {
	static constexpr bool TAIL = TAIL_;  // see WarpRng
	using L = PairLayout<N>;
	double *base;
	const double *log_tab;  // block-shared copy of glibc's log table {invc, logc}[128], followed by the pow tables
	int pair, lane, rd, parity;
	unsigned popped;

	__device__ __forceinline__ PairConsumer(double *pair_smem, const double *shared_log_tab, int pair_index)
		: base(pair_smem), log_tab(shared_log_tab), pair(pair_index), lane(threadIdx.x & 31), rd(0), parity(0), popped(0) {}

	// block-shared copies of the controller's pow tables (layout: stage_math_tables())
	__device__ __forceinline__ jsde_pow_tables pow_tables() const
	{
		const double *p = log_tab + 256;
		return jsde_pow_tables{p, p + 128, p + 256, reinterpret_cast<const unsigned long long *>(p + 384)};
	}

	__device__ __forceinline__ void pop(double &x1, double &x2, double &r2)
	{
		const double *p = base + rd * RNG_STRIDE + lane;
		x1 = p[0];
		x2 = p[L::CAP * RNG_STRIDE];
		r2 = p[2 * L::CAP * RNG_STRIDE];
		rd = rd + 1 >= L::CAP ? 0 : rd + 1;
		popped++;
	}

	// publish this lane's progress (and the trajectory it switches to, if any: `next_trajectory` >= 0), meet the
	// producer; returns the lanes that still have work (warp-uniform)
	__device__ __forceinline__ unsigned rendezvous(bool active, int next_trajectory)
	{
		const unsigned act = __ballot_sync(0xffffffffu, active);
		L::popped(base, parity)[lane] = popped;  // on a switch: the old trajectory's final count
		L::swap(base, parity)[lane] = next_trajectory;
		if (lane == 0)
			*L::mask(base, parity) = act;
		if (next_trajectory >= 0) {  // the new trajectory's queue starts at slot 0
			rd = 0;
			popped = 0;
		}
		parity ^= 1;
		__syncwarp();
		pair_barrier(pair);
		return act;
	}
};

// ---- producer side ------------------------------------------------------------------------------------------------------
template <int N, bool TAIL = false>
struct PairProducer {
	using L = PairLayout<N>;
	static constexpr int CAP = L::CAP;
	double *base;
	const double *log_tab;  // block-shared copy of glibc's log table (TAIL)
	uint4 *mt;  // generator records [B][156]
	int pair, lane;
	int traj;            // this lane's trajectory (-1: none yet)
	int wr, lvl, chunk;  // its ring write slot, (lower bound of the) fill level, next chunk
	unsigned pushed;

	__device__ __forceinline__ PairProducer(double *pair_smem, const double *shared_log_tab, int pair_index, uint4 *records)
		: base(pair_smem), log_tab(shared_log_tab), mt(records), pair(pair_index), lane(threadIdx.x & 31), traj(-1), wr(0), lvl(0), chunk(0), pushed(0) {}

	__device__ __forceinline__ double *fx1_() const { return base; }
	__device__ __forceinline__ double *fx2_() const { return base + CAP * RNG_STRIDE; }
	__device__ __forceinline__ double *fr2_() const { return base + 2 * CAP * RNG_STRIDE; }

	// queued pairs of the previous launch go to ring slots 0..lvl-1 (the consumer starts reading at slot 0)
	__device__ __forceinline__ void load(const double *spill, const int *levels, const int *chunks, int64_t B, int64_t k)
	{
		{
			lvl = levels[k];
			chunk = chunks[k];
			for (int i = 0; i < lvl; i++) {
				fx1_()[i * RNG_STRIDE + lane] = spill[((int64_t)i * 3 + 0) * B + k];
				fx2_()[i * RNG_STRIDE + lane] = spill[((int64_t)i * 3 + 1) * B + k];
				fr2_()[i * RNG_STRIDE + lane] = spill[((int64_t)i * 3 + 2) * B + k];
			}
			pushed = (unsigned)lvl;
			wr = lvl >= CAP ? lvl - CAP : lvl;
		}
	}

	__device__ __forceinline__ void store(double *spill, int *levels, int *chunks, int64_t B, int64_t k) const
	{
		{
			levels[k] = lvl;
			chunks[k] = chunk;
			int s = wr - lvl;
			s = s < 0 ? s + CAP : s;
			for (int i = 0; i < lvl; i++) {
				spill[((int64_t)i * 3 + 0) * B + k] = fx1_()[s * RNG_STRIDE + lane];
				spill[((int64_t)i * 3 + 1) * B + k] = fx2_()[s * RNG_STRIDE + lane];
				spill[((int64_t)i * 3 + 2) * B + k] = fr2_()[s * RNG_STRIDE + lane];
				s = s + 1 >= CAP ? 0 : s + 1;
			}
		}
	}

	// One "slot" = 4 trajectories x 8 candidates.  Lane (g, j) = (lane/8, lane%8) handles candidate j of the g-th
	// trajectory of the slot: state chunk c+j and tap chunk c+99+j (mt19937.cuh).  The words travel global -> shared
	// memory with cp.async, two slots ahead of the one being processed: no registers are tied up by loads in flight
	// (the producer lives on 56 registers) and lane j finds its neighbour's edge word next to its own chunk.
	// Stage entry of one trajectory: chunks c..c+8 at [0..8], chunks c+99..c+107 at [9..17].
	__device__ __forceinline__ void stage_slot(int s, int nserved, int g, int j) const
	{
		const int rank = (4 * s + g) & 31;
		const bool have = 4 * s + g < nserved;
		const int T = L::order(base)[rank] & 31;  // absent ranks: a stale but valid lane number, only ever used for shuffles
		const int c = __shfl_sync(0xffffffffu, chunk, T);
		const int kT = __shfl_sync(0xffffffffu, traj, T);
		if (have) {
			const uint4 *ST = mt + (int64_t)kT * MT_CHUNKS;
			uint4 *E = L::stage(base, s % L::STAGE_DEPTH, g);
			cp_async16(E + j, ST + mt_wrap(c + j));
			cp_async16(E + 9 + j, ST + mt_wrap(c + 99 + j));
			if (j == RNG_GROUP - 1) {
				cp_async16(E + 8, ST + mt_wrap(c + 8));
				cp_async16(E + 17, ST + mt_wrap(c + 107));
			}
		}
		cp_async_commit();  // one group per slot, empty or not: wait_group counts groups
	}

	// serve every lane in `room` exactly once: its trajectory advances by one group of 8 chunks
	__device__ __forceinline__ void refill(unsigned room)
	{
		const unsigned full = 0xffffffffu;
		const int g = lane >> 3, j = lane & 7;
		const int nserved = __popc(room);
		const int slots = (nserved + 3) >> 2;
		const int below = __popc(room & ((1u << lane) - 1u));
		const bool mine_served = (room >> lane) & 1u;
		__syncwarp();
		if (mine_served)
			L::order(base)[below] = (unsigned char)lane;  // compacted list of the lanes to serve
		__syncwarp();
		int mine = 0;
#pragma unroll
		for (int s = 0; s < JSDE_DUO_STAGE_AHEAD; s++)
			stage_slot(s, nserved, g, j);
#pragma unroll 1
		for (int s = 0; s < slots; s++) {
			stage_slot(s + JSDE_DUO_STAGE_AHEAD, nserved, g, j);  // reuses the buffer read in iteration s-1 (the ballot below ordered those reads)
			asm volatile("cp.async.wait_group %0;" ::"n"(JSDE_DUO_STAGE_AHEAD) : "memory");
			__syncwarp();
			const int rank = 4 * s + g;
			const bool have = rank < nserved;
			const int T = L::order(base)[rank] & 31;
			const int c = mt_wrap(__shfl_sync(full, chunk, T) + j);
			const int wrT = __shfl_sync(full, wr, T);
			const int kT = __shfl_sync(full, traj, T);
			bool ok = false;
			double x1 = 0.0, x2 = 0.0, r2 = 0.0;
			if (have) {
				const uint4 *E = L::stage(base, s % L::STAGE_DEPTH, g);
				const uint4 a = E[j], tap = E[9 + j];
				const uint32_t next = E[j + 1].x, tap_hi = E[10 + j].x;
				uint4 r;
				r.x = mt_twist(a.x, a.y, tap.y);
				r.y = mt_twist(a.y, a.z, tap.z);
				r.z = mt_twist(a.z, a.w, tap.w);
				r.w = mt_twist(a.w, next, tap_hi);
				mt[(int64_t)kT * MT_CHUNKS + c] = r;
				uint4 v;
				v.x = mt_temper(a.x);
				v.y = mt_temper(a.y);
				v.z = mt_temper(a.z);
				v.w = mt_temper(a.w);
				ok = polar_candidate(v, x1, x2, r2);
				if (TAIL && ok)
					r2 = polar_radius_factor_with(r2, log_tab);
			}
			const unsigned okm = __ballot_sync(full, ok);
			const unsigned grp = (okm >> (g * RNG_GROUP)) & 0xffu;
			if (ok) {
				int slot = wrT + __popc(grp & ((1u << j) - 1u));
				slot = slot >= CAP ? slot - CAP : slot;
				const int at = slot * RNG_STRIDE + T;
				fx1_()[at] = x1;
				fx2_()[at] = x2;
				fr2_()[at] = r2;
			}
			// owner: the lane whose trajectory has rank r among `room` was served by slot r/4, group r%4
			if ((below >> 2) == s)
				mine = __popc((okm >> ((below & 3) * RNG_GROUP)) & 0xffu);
		}
		if (mine_served) {
			lvl += mine;
			pushed += (unsigned)mine;
			wr += mine;
			wr = wr >= CAP ? wr - CAP : wr;
			chunk = mt_wrap(chunk + RNG_GROUP);
		}
	}

	// the producer warp's whole life
	__device__ __forceinline__ void run(double *spill, int *levels, int *chunks, int64_t B)
	{
		const unsigned full = 0xffffffffu;
		int parity = 0;
		while (true) {
			pair_barrier(pair);
			const unsigned act = *L::mask(base, parity);
			lvl = (int)(pushed - L::popped(base, parity)[lane]);
			const int next = L::swap(base, parity)[lane];
			parity ^= 1;
			if (next >= 0 || next == -2) {  // the consumer lane has moved on to another trajectory, or gives this one up (-2)
				if (traj >= 0)
					store(spill, levels, chunks, B, traj);
				// a trajectory can change hands between SMs within a launch (kernels.cuh, yield): what this warp wrote for the
				// old one must be visible before the consumer publishes it, and what another SM wrote for the new one before
				// this warp reads it
				__threadfence();
				traj = next >= 0 ? next : -1;
				if (traj >= 0)
					load(spill, levels, chunks, B, traj);
			}
			__syncwarp();
			if (!act)
				break;
			const bool wants = traj >= 0 && ((act >> lane) & 1u);
			while (__ballot_sync(full, wants && lvl < 2 * N))
				refill(__ballot_sync(full, wants && lvl <= CAP - RNG_GROUP));
		}
		if (traj >= 0)
			store(spill, levels, chunks, B, traj);  // queued pairs and stream position survive until the next launch
	}
};

template <int REGS>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }

}  // namespace jsde
