// Kernels: the fused integrate loop (one trajectory per lane, whole call in one launch), the single-step surface that
// mirrors the reference core call by call, and the small utilities (seeding, layout transposes, moments, FP64 peak).
#pragma once

#include "stepper.cuh"
#include "duo.cuh"
#include "kernels_common.cuh"
#include "philox_jumps.h"

namespace jsde {

constexpr int BLOCK = 128;  // single-step kernels: one trajectory per thread

// Who computes the scale-independent Gaussian factor s = sqrt(-2 log(r2)/r2) (warp_rng.cuh: TAIL): the stepper (default) or
// the generator side.  One choice per model library — queue contents must mean the same to every kernel.  Round 2 measured
// the generator-side variant (the idea: for multiplicative-noise models the producer warps idle more than half of the time):
// additive Lorenz -17 %, and with the final register splits SRI-Lorenz -4 %, jumpy Lorenz -11 %, geometric Brownian motion
// -6.5 %, stochastic Duffing +3.6 % (profiles/r02_tail_in_producer_ab.log) — bit-identical results either way (the parity
// suite passes in both settings).  It is on where it pays: two-component SRI models, whose producers keep 48 registers.
#ifndef JSDE_RNG_TAIL
#define JSDE_RNG_TAIL(additive, n) (!(additive) && (n) == 2)
#endif
template <class Model>
constexpr bool rng_tail = JSDE_RNG_TAIL(Model::ADDITIVE, Model::N);
template <class Model>
using ModelRng = WarpRng<Model::N, rng_tail<Model>>;

// ---- shared memory of the single-step kernels: one set of generator FIFOs per warp (warp_rng.cuh) ---------------------
constexpr int LOG_TAB_DOUBLES = 256;  // {invc, logc}[128], shared by the block

template <class Model>
constexpr size_t rng_smem_bytes() { return ((size_t)(BLOCK / 32) * WarpRng<Model::N>::SMEM_DOUBLES_PER_WARP + LOG_TAB_DOUBLES) * sizeof(double); }

// Whole block must call (contains a __syncthreads).
template <class Model>
__device__ __forceinline__ ModelRng<Model> make_rng(const EnsembleView &e, int64_t k, bool live)
{
	extern __shared__ __align__(16) double jsde_smem[];
	for (int i = threadIdx.x; i < 128; i += BLOCK) {
		jsde_smem[2 * i] = JSDE_LOG_INVC[i];
		jsde_smem[2 * i + 1] = JSDE_LOG_LOGC[i];
	}
	__syncthreads();
	const int warp = threadIdx.x >> 5;
	return ModelRng<Model>(jsde_smem + LOG_TAB_DOUBLES + warp * WarpRng<Model::N>::SMEM_DOUBLES_PER_WARP, jsde_smem, e.mt, k - (threadIdx.x & 31), live);
}

// ---- the hot path -------------------------------------------------------------------------------------------------
// jitcsde.integrate (_jitcsde.py:597-630) and, with a tape, jitcsde_jump.integrate (:693-700) for every trajectory.
// Warp-specialised (duo.cuh): consumer warp w steps trajectories [128*block + 32*w, +32), one per lane, one attempted
// step per round; producer warp w keeps their polar-pair FIFOs filled.  Choosing the next target and applying a jump
// are folded into the consumer's loop so that lanes which jump do not serialise the warp.
template <class Model>
constexpr size_t duo_smem_bytes() { return ((size_t)DUO_PAIRS * PairLayout<Model::N>::DOUBLES + MATH_TAB_DOUBLES) * sizeof(double); }

// RECYCLE: the grid is one wave of blocks and the trajectories beyond it wait in a pool (e.pool: a compact array of
// trajectory numbers, CallTotals::waiting of them).  A lane whose trajectory is done pops the last waiting one; and every
// `e.slice_rounds` rounds a warp's lanes *yield*: each exchanges its unfinished trajectory with a waiting one (atomicExch at
// a rotating position), so that all trajectories advance at the same average pace and finish together — without that, an
// ensemble of 2.3 times the resident lanes (2^20 trajectories split over 8 GPUs) takes three trajectory lengths instead of
// 2.3.  Measured (profiles/r02_time_slicing_ab.log; same box, bit-identical results in every setting): additive Lorenz,
// 2^17 trajectories to t = 100: 1.07e10 (plain grid) / 1.04e10 (pool only) -> 1.15e10 accepted steps/s; 2^20 trajectories:
// 1.16e10 / 1.11e10 -> 1.17e10; SRI-Lorenz 2^20 to t = 10: 5.57e9 / 6.58e9 -> 6.78e9.  The pool alone pays when the
// trajectories' step counts differ a lot (multiplicative noise) and costs a few per cent when they do not (fuller warps,
// lower clock under the power cap); with time slicing it is the default for every model (JSDE_RECYCLE=0/1, JSDE_SLICE and
// jsde_set_option("recycle" | "slice") override).
//
// A yield takes three passes of the (cold) outer loop, because the trajectory's generator-side state is written by the
// producer warp: pass 1 hands the stepper state back and announces the release (swap = -2); the producer spills queue
// and stream position during the following round; pass 3, two rendezvous later, publishes the trajectory number with
// the exchange and takes the other one.  __threadfence on both sides of the exchange makes the state written by one
// SM's consumer and producer visible to the SM that continues the trajectory.  Exchanges conserve trajectory numbers
// whatever the interleaving: a lane that finds a slot emptied by a concurrent pop (-1) takes its deposit back.
template <class Model, bool GRID, bool RECYCLE>
__global__ void __launch_bounds__(DUO_BLOCK, DuoConfig<Model::N, Model::ADDITIVE, Model::HAS_JUMP>::BLOCKS) k_integrate(EnsembleView e, Controller c, double target, int use_tape,
	const double *__restrict__ grid_times_, int grid_count, double *__restrict__ grid_out)
{
	constexpr int N = Model::N;
	constexpr bool SLICED = RECYCLE && !GRID;  // a yielded trajectory would have to carry its sample index along
	extern __shared__ __align__(16) double jsde_smem[];
	stage_math_tables(jsde_smem);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int pair = warp & (DUO_PAIRS - 1);
	double *pair_smem = jsde_smem + MATH_TAB_DOUBLES + pair * PairLayout<N>::DOUBLES;

	if (warp >= DUO_PAIRS) {  // ---- producer warpgroup ----
		setmaxnreg_dec<DuoConfig<N, Model::ADDITIVE, Model::HAS_JUMP>::PRODUCER_REGS>();
		PairProducer<N, rng_tail<Model>> P(pair_smem, jsde_smem, pair, e.mt);
		P.run(e.rng_spill, e.rng_level, e.mt_chunk, e.B);
		return;
	}

	// ---- consumer warpgroup ----
	setmaxnreg_inc<DuoConfig<N, Model::ADDITIVE, Model::HAS_JUMP>::CONSUMER_REGS>();
	const double *grid_times = GRID ? grid_times_ : nullptr;  // compile-time nullptr keeps the plain kernel free of grid code
	// Sampling-grid mode (grid_times != nullptr): the lane integrates to grid_times[0], records its state in
	// grid_out[0][·][k], goes on to grid_times[1], … — exactly the sequence of integrate() calls of
	// examples/noisy_lorenz.py:95-96 (truncated last step + Brownian bridge at every sample time), in one launch.
	PairConsumer<N, rng_tail<Model>> rng(pair_smem, jsde_smem, pair);
	Lane<Model, PairConsumer<N, rng_tail<Model>>> L(e, 0, rng);
	// This lane starts with the trajectory of its global index and (RECYCLE) goes on with trajectories from the pool.
	// Results do not depend on which lane integrates a trajectory, nor on how often it changes hands.
	int k = -1;  // B < 2^31 (jsde_create)
	int upcoming = (int)min((int64_t)0x7fffffff, ((int64_t)blockIdx.x * DUO_PAIRS + pair) * 32 + lane);
	if (upcoming >= e.B)
		upcoming = -1;
	bool seeking = false;   // RECYCLE: no trajectory, but the pool may still have one
	int held = -1;          // SLICED: the trajectory this lane is giving up (not yet published)
	int yield_phase = 0;    // SLICED: 0 none, 1 released (the producer is spilling), 2 ready for the exchange
	bool yielding = false;
	int rounds_left = SLICED && e.slice_rounds > 0 ? e.slice_rounds : 0x7fffffff;  // warp-uniform
	unsigned acc = 0, rej = 0, jumps = 0;  // per call and trajectory: far below 2^32 (the noise memory would overflow first)
	int grid_index = 0;
	double cur_target = target;  // differs from `target` only in sampling-grid mode
	double next_jump = 0.0;
	bool jump_set = false, exhausted = false;
	long long tape_pos = 0;
	bool active = false, sit_out = false;
	bool need_target = true, to_jump = false;
	double tgt = target;
	bool primed = false;  // round 0 is empty: the producer fills the FIFOs for the first time
	// Two nested loops: the inner one is the hot round loop; the outer one (cold) hands trajectories back and takes new
	// ones, so that its register needs do not shape the allocation of the hot loop.
	bool finished = false;
	while (!finished) {
		int next_trajectory = -1;  // what the producer is told: >= 0 switch to that trajectory, -2 release the current one
		CallTotals *tot = e.totals;
		if (SLICED && yield_phase == 2) {  // the producer has spilled `held`: publish it and take another
			int got = held;
			__threadfence();
			const int waiting = *(volatile int *)&tot->waiting;
			if (waiting > 0) {
				const int p = (int)(atomicAdd(&tot->cursor, 1ull) % (unsigned long long)waiting);
				got = atomicExch(&e.pool[p], held);
				if (got < 0)  // a concurrent pop has just emptied this slot: take the deposit (or a racer's) back
					got = atomicExch(&e.pool[p], -1);
			}
			__threadfence();
			held = -1;
			yield_phase = 0;
			upcoming = got;
			seeking = got < 0;
		} else if (SLICED && yield_phase == 1)
			yield_phase = 2;
		else if (!active && k >= 0) {  // hand back: the trajectory is done (or failed), or this lane yields it
			L.store();
			if (Model::HAS_JUMP && use_tape) {
				e.next_jump[k] = next_jump;
				e.jump_set[k] = jump_set ? 1 : 0;
				e.tape_pos[k] = tape_pos;
			}
			e.accepted[k] += acc;
			e.rejected[k] += rej;
			// once per trajectory (and slice): cheaper than carrying running sums in registers
			if (acc) atomicAdd(&tot->accepted, (unsigned long long)acc);
			if (rej) atomicAdd(&tot->rejected, (unsigned long long)rej);
			if (L.fresh) atomicAdd(&tot->fresh_draws, (unsigned long long)L.fresh);
			if (jumps) atomicAdd(&tot->jumps, (unsigned long long)jumps);
			if (Model::HAS_JUMP && exhausted) atomicAdd(&tot->tape_exhausted, 1ull);
			if (L.status == TRAJ_MIN_STEP) atomicAdd((unsigned long long *)&tot->n_min_step, 1ull);
			if (L.status == TRAJ_NOISE_OVERFLOW) atomicAdd((unsigned long long *)&tot->n_overflow, 1ull);
			if (L.max_depth) atomicMax(&tot->max_depth, L.max_depth);
			if (SLICED && yielding) {
				held = k;
				yield_phase = 1;
				next_trajectory = -2;
				yielding = false;
			} else
				seeking = RECYCLE;
			k = -1;
		}
		if (RECYCLE && seeking) {  // pop the last waiting trajectory
			seeking = false;
			const int w = atomicSub(&tot->waiting, 1) - 1;
			if (w >= 0) {
				upcoming = atomicExch(&e.pool[w], -1);  // never empty: slots below `waiting` hold trajectories (see the exchange above)
				__threadfence();
			} else
				atomicAdd(&tot->waiting, 1);  // the pool is empty: this lane is done
		}
		if (upcoming >= 0) {  // take the trajectory; this lane sits out the round in which its queue is rebuilt
			k = upcoming;
			upcoming = -1;
			next_trajectory = k;
			L.rebind(k);
			L.load();
			if (L.status == TRAJ_MIN_STEP)
				L.status = TRAJ_OK;  // the reference raises once per call and lets the caller carry on (:569-579)
			L.cursor = 0;
			L.fresh = 0;
			L.max_depth = 0;
			acc = rej = jumps = 0;
			exhausted = false;
			if (Model::HAS_JUMP && use_tape) {
				next_jump = e.next_jump[k];
				jump_set = e.jump_set[k] != 0;
				tape_pos = e.tape_pos[k];
			}
			grid_index = 0;
			cur_target = grid_times ? grid_times[0] : target;
			need_target = true;
			to_jump = false;
			active = L.status == TRAJ_OK;  // a trajectory that cannot run is handed back in the next round
			sit_out = true;
		}
		while (true) {
		// the warp goes on while any lane steps, holds a trajectory to hand back, or is in the middle of a yield
		if (!rng.rendezvous(active || k >= 0 || (SLICED && yield_phase != 0), next_trajectory)) {
			finished = true;
			break;
		}
		next_trajectory = -1;
		const bool resting = sit_out;  // the producer is rebuilding this lane's queue
		sit_out = false;
		if (!primed) {
			primed = true;
			continue;
		}
		bool attempt = false, last = false;
		double actual_dt = 0.0;
		if (active && !resting) {
			attempt = true;
			if (need_target) {
				tgt = GRID ? cur_target : target;
				to_jump = false;
				if (Model::HAS_JUMP && use_tape) {
					if (!jump_set) {  // next_jump property (:682-687): t + IJI(t, y)
						double wait = __longlong_as_double(0x7ff0000000000000LL);  // past the end of a tape: no further jumps
						if (e.generated_jumps)
							wait = Model::waiting_time(L.t, L.y, jsde_jump_tau(e.jump_seed, e.jump_first + k, tape_pos), L.par);
						else if (tape_pos < e.tape_len)
							wait = Model::waiting_time(L.t, L.y, e.taus[(int64_t)k * e.tape_len + tape_pos], L.par);
						else
							exhausted = true;  // reported (jsde_stats.n_tape_exhausted): this trajectory's jump process ends here
						next_jump = L.t + wait;
						jump_set = true;
					}
					if (next_jump < (GRID ? cur_target : target)) {
						tgt = next_jump;
						to_jump = true;
					}
				}
				need_target = false;
				if (L.t >= tgt) {  // integrate() with nothing to do (:614)
					attempt = false;
					if (!to_jump) {
						active = false;
						if (grid_times) {  // next sample time (cold path: once per sample, not per step)
#pragma unroll
							for (int i = 0; i < N; i++)
								grid_out[((int64_t)grid_index * N + i) * e.B + k] = L.y[i];
							if (++grid_index < grid_count) {
								cur_target = grid_times[grid_index];
								need_target = true;
								active = true;
							}
						}
					} else {
						double change[N];
						Model::jump(L.t, L.y, e.generated_jumps ? jsde_jump_xi(e.jump_seed, e.jump_first + k, tape_pos) : e.xis[(int64_t)k * e.tape_len + tape_pos], L.par, change);
#pragma unroll
						for (int i = 0; i < N; i++)
							L.y[i] += change[i];
						tape_pos++;
						jumps++;
						jump_set = false;
						need_target = true;
					}
				}
			}
		}
		if (attempt) {
			if (L.t + L.dt < tgt)
				actual_dt = L.dt;
			else {
				actual_dt = tgt - L.t;
				last = true;
			}
			if (!L.propose(actual_dt))
				active = false;  // noise memory overflow
			else {
				const int verdict = L.adjust_step(c, actual_dt);
				if (verdict == 1) {
					L.commit();
					acc++;
					if (last) {
						if (!to_jump) {
							active = false;
							if (grid_times) {
#pragma unroll
								for (int i = 0; i < N; i++)
									grid_out[((int64_t)grid_index * N + i) * e.B + k] = L.y[i];
								if (++grid_index < grid_count) {
									cur_target = grid_times[grid_index];
									need_target = true;
									active = true;
								}
							}
						} else {
							double change[N];  // apply_jump(amp(time, state)) (:696-698)
							Model::jump(L.t, L.y, e.generated_jumps ? jsde_jump_xi(e.jump_seed, e.jump_first + k, tape_pos) : e.xis[(int64_t)k * e.tape_len + tape_pos], L.par, change);
#pragma unroll
							for (int i = 0; i < N; i++)
								L.y[i] += change[i];
							tape_pos++;
							jumps++;
							jump_set = false;
							need_target = true;
						}
					}
				} else {
					rej++;
					if (verdict < 0) {
						L.status = TRAJ_MIN_STEP;
						active = false;
					}
				}
			}
		}
		if (SLICED && --rounds_left == 0) {  // the warp's lanes give their trajectories up for waiting ones
			rounds_left = e.slice_rounds;
			if (active && !resting && *(volatile int *)&e.totals->waiting > 0) {
				yielding = true;
				active = false;
			}
		}
		if (__any_sync(0xffffffffu, (!active && k >= 0) || (SLICED && yield_phase != 0)))
			break;  // somebody has a trajectory to hand back (and perhaps one to take), or a yield to complete
		}
	}
	// every lane has handed back its last trajectory before the loop can end: a lane with a trajectory (k >= 0) or a yield
	// in progress keeps the warp alive (rendezvous), and the cold pass above is entered whenever such a lane is not stepping
}

// ---- single-step surface (reference core methods, one launch each) ---------------------------------------------
template <class Model>
__global__ void __launch_bounds__(BLOCK) k_get_next_step(EnsembleView e, const double *h)
{
	const int64_t k = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
	const bool live = k < e.B;
	const int64_t kk = live ? k : 0;
	ModelRng<Model> rng = make_rng<Model>(e, k, live);
	rng.load(e.rng_spill, e.rng_level, e.mt_chunk, e.B, kk);
	Lane<Model, ModelRng<Model>> L(e, kk, rng);
	bool go = false;
	if (live) {
		L.load();
		go = L.status != TRAJ_NOISE_OVERFLOW;
	}
	rng.ensure(go, Model::N);
	if (go) {
		if (L.propose(h[k]))
			L.store_proposal();
		L.store();
	}
	rng.store(e.rng_spill, e.rng_level, e.mt_chunk, e.B, kk);
}

template <class Model>
__global__ void __launch_bounds__(BLOCK) k_get_p(EnsembleView e, double atol, double rtol, double *p)
{
	const int64_t k = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
	if (k >= e.B)
		return;
	ModelRng<Model> rng(nullptr, nullptr, e.mt, 0, false);
	Lane<Model, ModelRng<Model>> L(e, k, rng);
	L.load_proposal();
	p[k] = L.error_ratio(atol, rtol);
}

template <class Model>
__global__ void __launch_bounds__(BLOCK) k_accept_step(EnsembleView e, const unsigned char *mask)
{
	const int64_t k = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
	if (k >= e.B || (mask && !mask[k]))
		return;
	ModelRng<Model> rng(nullptr, nullptr, e.mt, 0, false);
	Lane<Model, ModelRng<Model>> L(e, k, rng);
	L.load();
	L.load_proposal();
	L.commit();
	L.store();
}

template <class Model>
__global__ void __launch_bounds__(BLOCK) k_pin_noise(EnsembleView e, unsigned number, double step)
{
	const int64_t k = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
	const bool live = k < e.B;
	const int64_t kk = live ? k : 0;
	ModelRng<Model> rng = make_rng<Model>(e, k, live);
	rng.load(e.rng_spill, e.rng_level, e.mt_chunk, e.B, kk);
	Lane<Model, ModelRng<Model>> L(e, kk, rng);
	if (live)
		L.load();
	for (unsigned j = 0; j < number; j++) {
		const bool go = live && L.status == TRAJ_OK;
		rng.ensure(go, Model::N);
		if (go)
			L.pin_one(step);
	}
	if (live)
		L.store();
	rng.store(e.rng_spill, e.rng_level, e.mt_chunk, e.B, kk);
}

// pin_noise with caller-supplied Wiener increments: item j of trajectory k has length steps[j], DW = dw[k][j][·],
// DZ = dz[k][j][·] (no generator involved)
template <class Model>
__global__ void __launch_bounds__(BLOCK) k_pin_path(EnsembleView e, unsigned number, const double *steps, const double *dw, const double *dz)
{
	constexpr int N = Model::N;
	const int64_t k = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
	if (k >= e.B)
		return;
	ModelRng<Model> rng(nullptr, nullptr, e.mt, 0, false);
	Lane<Model, ModelRng<Model>> L(e, k, rng);
	L.load();
	for (unsigned j = 0; j < number && L.status == TRAJ_OK; j++) {
		double W[N], Z[N];
#pragma unroll
		for (int i = 0; i < N; i++) {
			W[i] = dw[((int64_t)k * number + j) * N + i];
			Z[i] = dz[((int64_t)k * number + j) * N + i];
		}
		L.pin_given(steps[j], W, Z);
	}
	L.store();
}

// rk_gauss stream of every trajectory: out [B][pairs][2]
template <class Model>
__global__ void __launch_bounds__(BLOCK) k_draw_gauss(EnsembleView e, double scale, int pairs, double *out)
{
	const int64_t k = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
	const bool live = k < e.B;
	const int64_t kk = live ? k : 0;
	ModelRng<Model> rng = make_rng<Model>(e, k, live);
	rng.load(e.rng_spill, e.rng_level, e.mt_chunk, e.B, kk);
	for (int j = 0; j < pairs; j++) {
		rng.ensure(live, 1);
		if (live) {
			double x1, x2, r2;
			rng.pop(x1, x2, r2);
			double f;
			if constexpr (rng_tail<Model>)
				f = scale * r2;  // the queue holds the factor itself
			else {
				const double lg = jsde_log_is_near_one(r2) ? jsde_log_near_one(r2) : jsde_log_table_from(r2, rng.log_tab);
				f = scale * sqrt(__ddiv_rn(__dmul_rn(-2.0, lg), r2));
			}
			out[(k * pairs + j) * 2] = x1 * f;
			out[(k * pairs + j) * 2 + 1] = x2 * f;
		}
	}
	rng.store(e.rng_spill, e.rng_level, e.mt_chunk, e.B, kk);
}

}  // namespace jsde
