// Producer / consumer warp pairs for the fused integrate kernel.
//
// A 256-thread block holds four *consumer* warps (warpgroup 0: one trajectory per lane, runs the stepper of
// stepper.cuh) and four *producer* warps (warpgroup 1).  Producer warp p runs the Mersenne Twisters of the 32
// trajectories of consumer warp p — in its own warp, with its own (small) register budget, concurrently with the
// consumer's floating-point work:
//
//   * the generator's integer instructions and the stepper's FP64 dependency chains come from different warps, so the
//     scheduler interleaves them;
//   * `setmaxnreg` moves registers from the producers to the consumers: 24 warps per SM instead of 16.
//
// Round 2: the producer's data path.  Round 1 advanced four trajectories by eight chunks each per step (shared-memory
// FIFOs of 13 pairs per lane left no room for more), and two thirds of the producer's ~1050 instructions per round of 32
// attempts were bookkeeping around the ~70 that matter per candidate: per-lane shuffles of the served trajectory's
// counters, 16-byte cp.async copies with per-lane wrap-around arithmetic, compacted writes to four different FIFOs
// (profiles/r01_duo_final_*: duo.cuh = 30 % of all instructions executed).  Now
//
//   * a *service* regenerates SERVICE_CHUNKS = 32 consecutive chunks of ONE trajectory, one chunk per lane: everything
//     about the trajectory is warp-uniform, the 33 + 33 chunks it reads (own words and the tap words 397 positions
//     ahead, mt19937.cuh) arrive by `cp.async.bulk` (TMA, one or two copies per range, issued by one lane, completion
//     on an mbarrier) in a stage buffer that was requested one round earlier, and the ~25 accepted candidates go, in
//     stream order (ballot/popc), to the trajectory's pair ring in global memory (warp_rng.cuh: PairRing — 1 KB per
//     trajectory; only the rings of resident trajectories are hot and they stay in L2);
//   * every round, producer lane l copies the next pairs of ITS trajectory from the ring into a lane-private FIFO of
//     L0 = 2N pairs in shared memory (column l of [L0][32]: conflict-free for both sides), from which consumer lane l
//     pops.  The consumer never waits for global memory.
//
// Protocol (per pair, one named barrier of 64 threads per round, control words double-buffered by round parity):
//
//   consumer, before barrier r:  popped[r&1][lane] = pairs popped so far;  mask[r&1] = lanes that still have work
//   producer, after  barrier r:  finishes the services requested in round r-1; tops every lane's FIFO up to L0 pairs
//                                as judged by the published (older, smaller) popped count — which also makes the slots
//                                it writes disjoint from the ones the consumer may still read; requests services
//                                (bulk copies) for the trajectories whose ring holds fewer than LOW pairs
//   consumer, during round r:    pops at most N pairs  ->  at barrier r+1 every lane still holds >= N  ->  round r+1 is
//                                safe while the producer tops up again.  Round 0 of the consumer is empty.
//
// If a ring cannot fill its lane's FIFO (first round of a launch, or a service that yields fewer pairs than needed —
// every candidate may be rejected), the trajectory is served synchronously: request, wait, regenerate, repeat.
// Both loops leave on the same published mask == 0.  Queued pairs and the stream position persist in global memory
// (rng_cons, rng_level, mt_chunk): repeated integrate() calls and the single-step kernels continue the stream.
//
// Trajectory recycling (kernels.cuh, RECYCLE): the grid is at most one wave of blocks; a lane whose trajectory is done
// takes the next one from a pool (atomic counter) instead of idling until its warp's and block's slowest trajectory
// finishes.  The consumer announces the change through `swap[lane]` together with the final popped count of the old
// trajectory; the producer writes the old trajectory's counters back, loads the new one's and fills the FIFO — the lane
// sits out that one round.  Every lane's first trajectory arrives the same way (round 0).
#pragma once

#include "warp_rng.cuh"

namespace jsde {

constexpr int DUO_PAIRS = 4;
constexpr int DUO_BLOCK = 64 * DUO_PAIRS;
// Resident blocks per SM and the register split between the warpgroups (the launch allocation 65536/(256*blocks),
// rounded down to a multiple of 8, is redistributed by setmaxnreg: consumer + producer = 2 x allocation).  Up to three
// components the stepper fits 104 registers (SRA) and three blocks are resident; beyond that it would spill, so two
// blocks with 168/88.  The macros override.
template <int N, bool ADDITIVE>
struct DuoConfig {
#ifdef JSDE_DUO_MIN_BLOCKS
	static constexpr int BLOCKS = JSDE_DUO_MIN_BLOCKS;
#else
	static constexpr int BLOCKS = N <= 3 ? 3 : 2;
#endif
#ifdef JSDE_DUO_CONSUMER_REGS
	static constexpr int CONSUMER_REGS = JSDE_DUO_CONSUMER_REGS;
#else
	static constexpr int CONSUMER_REGS = N <= 3 ? (ADDITIVE ? 104 : 112) : 168;
#endif
#ifdef JSDE_DUO_PRODUCER_REGS
	static constexpr int PRODUCER_REGS = JSDE_DUO_PRODUCER_REGS;
#else
	static constexpr int PRODUCER_REGS = N <= 3 ? (ADDITIVE ? 56 : 48) : 88;
#endif
};
#ifndef JSDE_DUO_STAGE_BUFFERS
#define JSDE_DUO_STAGE_BUFFERS 8  // services per producer warp whose generator words can be in flight at once
#endif

__device__ __forceinline__ void pair_barrier(int pair)
{
	asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory");  // barrier 0 is __syncthreads()
}

// ---- mbarrier + bulk copy (TMA) primitives ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity)
{
	unsigned done;
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
		"selp.u32 %0, 1, 0, p;\n"
		"}"
		: "=r"(done)
		: "r"(smem_u32(bar)), "r"(parity)
		: "memory");
	return done != 0;
}
// try_wait suspends the thread in hardware for a while; a copy that never completes (a bug) traps instead of hanging
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
	if (mbar_try_wait(bar, parity))
		return;
	const long long start = clock64();
	while (!mbar_try_wait(bar, parity))
		if (clock64() - start > 8000000000ll)
			__trap();
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on `bar` (complete_tx)
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
		"l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
		: "memory");
}
// order this thread's earlier global stores before bulk copies (async proxy) issued after the next warp synchronisation
__device__ __forceinline__ void fence_stores_before_bulk() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

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

// shared memory of one producer/consumer pair
template <int N>
struct PairLayout {
	using R = PairRing<N>;
	static constexpr int L0 = R::L0;
	static constexpr int NBUF = JSDE_DUO_STAGE_BUFFERS;
	static constexpr int STAGE_CHUNKS = 2 * (SERVICE_CHUNKS + 1);  // own chunks c..c+32 at [0..32], tap chunks c+99..c+131 at [33..65]
	static constexpr int FIFO_DOUBLES = 3 * L0 * 32;               // x1, x2, r2: [L0][32] each, column = lane
	// FIFOs | popped[2][32] u32 | swap[2][32] i32 | mask[2] u32 | mbarriers[NBUF] u64 | stage[NBUF][STAGE_CHUNKS] uint4
	static constexpr int BAR_OFFSET = FIFO_DOUBLES + 32 + 32 + 1;
	static constexpr int STAGE_OFFSET = (BAR_OFFSET + NBUF + 1) & ~1;  // 16-byte aligned
	static constexpr int DOUBLES = STAGE_OFFSET + NBUF * STAGE_CHUNKS * 2;
	static __device__ __forceinline__ unsigned *popped(double *base, int parity) { return reinterpret_cast<unsigned *>(base + FIFO_DOUBLES) + 32 * parity; }
	// swap[2][32] i32: trajectory a lane switches to in this round (-1: none)
	static __device__ __forceinline__ int *swap(double *base, int parity) { return reinterpret_cast<int *>(base + FIFO_DOUBLES + 32) + 32 * parity; }
	static __device__ __forceinline__ unsigned *mask(double *base, int parity) { return reinterpret_cast<unsigned *>(base + FIFO_DOUBLES + 64) + parity; }
	static __device__ __forceinline__ unsigned long long *bar(double *base, int buf) { return reinterpret_cast<unsigned long long *>(base + BAR_OFFSET) + buf; }
	static __device__ __forceinline__ uint4 *stage(double *base, int buf) { return reinterpret_cast<uint4 *>(base + STAGE_OFFSET) + buf * STAGE_CHUNKS; }
};

// ---- consumer side: what Lane<> needs from its pair source ---------------------------------------------------------
template <int N>
struct PairConsumer {
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
		const double *p = base + rd * 32 + lane;
		x1 = p[0];
		x2 = p[L::L0 * 32];
		r2 = p[2 * L::L0 * 32];
		rd = rd + 1 >= L::L0 ? 0 : rd + 1;
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
template <int N>
struct PairProducer {
	using L = PairLayout<N>;
	using R = PairRing<N>;
	static constexpr int L0 = L::L0;
	static constexpr int NBUF = L::NBUF;
	static constexpr unsigned FULL = 0xffffffffu;
	double *base;
	uint4 *mt;       // generator records [B][156]
	double *rings;   // pair rings [B][3][RING]
	int pair, lane;
	int traj;        // this lane's trajectory (-1: none yet)
	int chunk;       // its next chunk to regenerate
	int rpos, rlvl;  // its ring: index of the next pair to hand over, pairs waiting behind it
	int wr0, lvl0;   // its FIFO: write slot, (upper bound of the) fill level
	unsigned pushed; // pairs handed over so far
	unsigned pend;   // lanes whose trajectories have a service in flight: the i-th set bit uses stage buffer i (warp-uniform)
	unsigned phases; // current phase parity of every stage buffer's mbarrier (warp-uniform)
	bool dirty;      // generator words stored since the last proxy fence (warp-uniform)

	__device__ __forceinline__ PairProducer(double *pair_smem, int pair_index, uint4 *records, double *pair_rings)
		: base(pair_smem), mt(records), rings(pair_rings), pair(pair_index), lane(threadIdx.x & 31), traj(-1), chunk(0), rpos(0), rlvl(0),
		  wr0(0), lvl0(0), pushed(0), pend(0), phases(0), dirty(false) {}

	__device__ __forceinline__ void load(const int *cons, const int *levels, const int *chunks, int64_t k)
	{
		rpos = cons[k];
		rlvl = levels[k];
		chunk = chunks[k];
		pushed = 0;  // the consumer lane restarts its FIFO at slot 0 with popped = 0
		wr0 = 0;
		lvl0 = 0;
	}

	// `popped_final`: what the consumer lane popped of this trajectory; pairs handed over but not popped stay queued
	__device__ __forceinline__ void store(int *cons, int *levels, int *chunks, int64_t k, unsigned popped_final) const
	{
		const int unread = (int)(pushed - popped_final);
		const int c = rpos - unread;
		cons[k] = c < 0 ? c + R::RING : c;
		levels[k] = unread + rlvl;
		chunks[k] = chunk;
	}

	// ---- services -------------------------------------------------------------------------------------------------------
	// Request the generator words of the trajectories of the lanes in `want` (at most NBUF, lowest lanes first; the rest
	// waits for the next call).  Lane T issues its own copies: chunks c..c+32 and c+99..c+131 (mod 156) of its record.
	__device__ __forceinline__ void request(unsigned want)
	{
		if (__popc(want) > NBUF)
			want &= (2u << __fns(want, 0, NBUF)) - 1u;  // keep the NBUF lowest set bits
#if !defined(JSDE_DUO_NO_FENCE) && !defined(JSDE_DUO_STAGE_LDG)
		if (dirty) {  // the regenerated words of earlier services may be the tap words of these
			fence_stores_before_bulk();
			dirty = false;
		}
#endif
		__syncwarp();  // ... and every lane is done reading the stage buffers
		pend = want;
#ifdef JSDE_DUO_STAGE_LDG
		if ((want >> lane) & 1u) {  // experiment: no staging, only an L2 prefetch of the lines the service will read
			const uint4 *ST = mt + (int64_t)traj * MT_CHUNKS;
			for (int q = 0; q < SERVICE_CHUNKS + 1; q += 8) {
				asm volatile("prefetch.global.L2 [%0];" ::"l"(ST + mt_wrap(chunk + q)));
				asm volatile("prefetch.global.L2 [%0];" ::"l"(ST + mt_wrap(chunk + 99 + q)));
			}
		}
		return;
#endif
		if ((want >> lane) & 1u) {
			const int buf = __popc(want & ((1u << lane) - 1u));
			unsigned long long *bar = L::bar(base, buf);
			uint4 *E = L::stage(base, buf);
			const uint4 *ST = mt + (int64_t)traj * MT_CHUNKS;
			mbar_expect_tx(bar, L::STAGE_CHUNKS * 16);
#pragma unroll
			for (int part = 0; part < 2; part++) {
				const int first = part ? mt_wrap(chunk + 99) : chunk;
				const int head = min(SERVICE_CHUNKS + 1, MT_CHUNKS - first);  // chunks before the record's end
				uint4 *dst = E + part * (SERVICE_CHUNKS + 1);
				bulk_load(dst, ST + first, head * 16, bar);
				if (head < SERVICE_CHUNKS + 1)
					bulk_load(dst + head, ST, (SERVICE_CHUNKS + 1 - head) * 16, bar);
			}
		}
	}

	// regenerate the 32 chunks staged in buffer `buf` for lane T's trajectory; append the accepted candidates to its ring
	__device__ __forceinline__ void serve(int T, int buf)
	{
		const int cT = __shfl_sync(FULL, chunk, T);
		const int kT = __shfl_sync(FULL, traj, T);
		const int tail = R::wrap(__shfl_sync(FULL, rpos, T) + __shfl_sync(FULL, rlvl, T));  // rpos < RING, rlvl < RING
#ifdef JSDE_DUO_STAGE_LDG
		const uint4 *ST = mt + (int64_t)kT * MT_CHUNKS;
		const uint4 a = ST[mt_wrap(cT + lane)], tap = ST[mt_wrap(cT + 99 + lane)];
		uint32_t edge_next = 0, edge_tap = 0;
		if (lane == 31) {
			edge_next = ST[mt_wrap(cT + SERVICE_CHUNKS)].x;
			edge_tap = ST[mt_wrap(cT + 99 + SERVICE_CHUNKS)].x;
		}
		uint32_t next = __shfl_down_sync(FULL, a.x, 1);
		uint32_t tap_hi = __shfl_down_sync(FULL, tap.x, 1);
		if (lane == 31) {
			next = edge_next;
			tap_hi = edge_tap;
		}
		(void)buf;
#else
		mbar_wait(L::bar(base, buf), (phases >> buf) & 1u);
		phases ^= 1u << buf;
		const uint4 *E = L::stage(base, buf);
		const uint4 a = E[lane], tap = E[SERVICE_CHUNKS + 1 + lane];
		uint32_t next = __shfl_down_sync(FULL, a.x, 1);
		uint32_t tap_hi = __shfl_down_sync(FULL, tap.x, 1);
		if (lane == 31) {
			next = E[SERVICE_CHUNKS].x;
			tap_hi = E[2 * SERVICE_CHUNKS + 1].x;
		}
#endif
		uint4 r;
		r.x = mt_twist(a.x, a.y, tap.y);
		r.y = mt_twist(a.y, a.z, tap.z);
		r.z = mt_twist(a.z, a.w, tap.w);
		r.w = mt_twist(a.w, next, tap_hi);
		mt[(int64_t)kT * MT_CHUNKS + mt_wrap(cT + lane)] = r;
		uint4 v;
		v.x = mt_temper(a.x);
		v.y = mt_temper(a.y);
		v.z = mt_temper(a.z);
		v.w = mt_temper(a.w);
		double x1, x2, r2;
		const bool ok = polar_candidate(v, x1, x2, r2);
		const unsigned okm = __ballot_sync(FULL, ok);
		if (ok) {
			const int at = R::wrap(tail + __popc(okm & ((1u << lane) - 1u)));
			double *ring = rings + (int64_t)kT * R::DOUBLES + at;
			ring[0] = x1;
			ring[R::RING] = x2;
			ring[2 * R::RING] = r2;
		}
		if (lane == T) {
			rlvl += __popc(okm);
			chunk = mt_wrap(chunk + SERVICE_CHUNKS);
		}
		dirty = true;
	}

	__device__ __forceinline__ void finish_services()
	{
		int buf = 0;
#pragma unroll 1
		for (unsigned m = pend; m; m &= m - 1u, buf++)
			serve(__ffs(m) - 1, buf);
		if (pend)
			__syncwarp();  // ring entries written by other lanes are read below
		pend = 0;
	}

	// hand this lane's next pairs over: ring -> FIFO column `lane`, up to L0 (judged by the published popped count)
	__device__ __forceinline__ void top_up()
	{
		int n = min(L0 - lvl0, rlvl);
		const double *ring = rings + (int64_t)(traj < 0 ? 0 : traj) * R::DOUBLES;
		double *fifo = base + lane;
#pragma unroll 1
		while (__any_sync(FULL, n > 0)) {
			double v[N][3];
#pragma unroll
			for (int i = 0; i < N; i++) {
				if (i < n) {
					const int at = R::wrap(rpos + i);
					v[i][0] = ld_ring(ring + at);
					v[i][1] = ld_ring(ring + R::RING + at);
					v[i][2] = ld_ring(ring + 2 * R::RING + at);
				}
			}
#pragma unroll
			for (int i = 0; i < N; i++) {
				if (i < n) {
					int s = wr0 + i;
					s = s >= L0 ? s - L0 : s;
					fifo[s * 32] = v[i][0];
					fifo[(L0 + s) * 32] = v[i][1];
					fifo[(2 * L0 + s) * 32] = v[i][2];
				}
			}
			const int m = min(n, N);
			if (m > 0) {
				rpos = R::wrap(rpos + m);
				wr0 = wr0 + m >= L0 ? wr0 + m - L0 : wr0 + m;
				rlvl -= m;
				lvl0 += m;
				pushed += (unsigned)m;
				n -= m;
			}
		}
	}

	// the producer warp's whole life
	__device__ __forceinline__ void run(int *cons, int *levels, int *chunks)
	{
		if (lane == 0) {
			for (int b = 0; b < NBUF; b++)
				mbar_init(L::bar(base, b), 1);
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		__syncwarp();
		int parity = 0;
		unsigned popped_final = 0;
		while (true) {
			pair_barrier(pair);
			const unsigned act = *L::mask(base, parity);
			const unsigned popped = L::popped(base, parity)[lane];
			const int next = L::swap(base, parity)[lane];
			parity ^= 1;
			finish_services();  // requested in the previous round, for the trajectories the lanes had then
			lvl0 = (int)(pushed - popped);
			popped_final = popped;
			if (next >= 0) {  // the consumer lane has moved on to another trajectory
				if (traj >= 0)
					store(cons, levels, chunks, traj, popped);
				traj = next;
				load(cons, levels, chunks, traj);
				popped_final = 0;
			}
			if (!act)
				break;
			const bool wants = traj >= 0 && ((act >> lane) & 1u);
			// a ring that cannot fill its lane's FIFO is served here and now (start of a launch; an unlucky service)
#pragma unroll 1
			while (const unsigned starved = __ballot_sync(FULL, wants && lvl0 + rlvl < L0)) {
				request(starved);
				finish_services();
			}
			top_up();
			request(__ballot_sync(FULL, wants && rlvl < R::LOW && lvl0 + rlvl + SERVICE_CHUNKS <= R::RING));
		}
		if (traj >= 0)
			store(cons, levels, chunks, traj, popped_final);  // queued pairs and stream position survive until the next launch
	}
};

template <int REGS>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }

}  // namespace jsde
