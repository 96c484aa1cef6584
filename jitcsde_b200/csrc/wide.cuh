// "Wide" path for large state vectors (SURVEY.md §8 d config 5: noisy FitzHugh–Nagumo network, n = 1000, dense
// 500x500 coupling): one 256-thread block integrates T (<= 8) trajectories in lock-step rounds, one attempted step per
// trajectory and round.  The lane-per-trajectory kernels keep a whole state in one thread's registers, which stops
// making sense beyond a handful of components; here per-trajectory vectors live in global memory (they are small
// next to the coupling matrix and stay in L2) and the work is split by phase:
//
//   * per-trajectory, sequential-stream work — noise memory lookup, the Gaussian draw from the trajectory's ONE
//     Mersenne Twister stream, Brownian bridge, controller — is done by one WARP per trajectory, concurrently for the
//     T trajectories, without block-wide barriers;
//   * the dense coupling  c_i = sum_j A_ij (y_j - y_i)  — 2*units^2 subtract/multiply/add per evaluation, > 95 % of the
//     arithmetic — is evaluated for ALL T trajectories at once: every thread owns <= 2 rows i and keeps 2 x T partial
//     sums in registers, so one load of A_ij (L2 -> register, coalesced over i) and one 64-byte broadcast of
//     y_j[0..T) from shared memory feed 6 T FP64 instructions.  The sum over j runs in ascending order with separately
//     rounded multiply and add, exactly like the oracle's loop, so results stay bit-identical (this is also why the
//     FP64 tensor-core MMA is not used: it fuses and re-associates);
//   * elementwise stage algebra is done by the whole block, trajectory after trajectory.
//
// Same algorithm and the same bits as the reference, restated for this shape:
//   generator ........ jitcsde/random_numbers.c:41-130: batch twist (:73-84) by one warp on a state held in shared
//                      memory; candidates are formed 32 at a time, the accepted ones compacted in stream order by
//                      ballot/popc; pair k of a draw belongs to component k (get_gauss, jitced_template.c:145-154).
//                      The stream position `pos` is kept like the reference's rk_state.pos (no queue of spare pairs).
//   noise memory ..... jitced_template.c:170-250; items are addressed through a small permutation (logical position ->
//                      physical slot) so that a Brownian-bridge insertion or a pop never moves 16 KB items.
//   SRA1 step ........ jitced_template.c:466-494;   error ratio / commit ... :502-536;   controller ... _jitcsde.py:569-630
// SRA1 for additive and SRIW1 for multiplicative noise (template:408-464); jumps from a tape or a counter-based generator.
#pragma once

#include <cstdio>

#include "ensemble.cuh"
#include "exact_div.cuh"
#include "gauss.cuh"
#include "mt19937.cuh"
#include "philox_jumps.h"

namespace jsde {
namespace wide {

// Block shape: 256 threads / up to 8 trajectories / two blocks per SM (default), or 512 threads / up to 16 trajectories /
// one block per SM, where every A_ij fetched from L2 serves twice as many trajectories.  Round 2 measured the second shape
// on the n = 1000 network (2^11 trajectories, t = 0..1): exact coupling 4.43e6 vs 4.77e6 attempts/s, tensor-core coupling
// 5.53e6 vs 5.71e6 (profiles/r02_wide_block_shape_ab.log) — two independent blocks per SM overlap one block's
// per-trajectory phases with the other's coupling phase, which is worth more than the halved L2 traffic.  Both shapes
// pass the parity suite; the larger one stays a build option (-DJSDE_WIDE_BLOCK=512).
#ifndef JSDE_WIDE_BLOCK
#define JSDE_WIDE_BLOCK 256
#endif
constexpr int WBLOCK = JSDE_WIDE_BLOCK;
static_assert(WBLOCK == 256 || WBLOCK == 512, "wide block shape");
constexpr int WWARPS = WBLOCK / 32;
constexpr int TMAX = WWARPS;  // trajectories per block: one warp each in the per-trajectory phases
constexpr int TPAD = TMAX;    // row stride of the coupling stage buffer [units][TPAD]
constexpr int WBLOCKS_PER_SM = WBLOCK == 256 ? 2 : 1;
constexpr int CANDIDATES = MT_WORDS / 4;  // 156 per regenerated block
constexpr int MAX_DEPTH = 64;

struct WideView {
	int64_t B;
	int D;
	double *y;         // [B][n]   (same layout as the host's)
	double *t, *dt;    // [B]
	double *qh;        // [B][D]        length of the item in physical slot d
	double *qw;        // [B][D][2][n]  its DW, DZ
	int *order;        // [B][D]        logical position -> physical slot (a permutation; first q_count entries in use)
	int *q_count;      // [B]
	uint32_t *key;     // [B][624]      raw generator state (rk_state.key)
	int *pos;          // [B]           next unread word of the generator block (rk_state.pos; 624: regenerate first)
	int *status;
	unsigned long long *accepted, *rejected;
	const double *par;
	int64_t par_stride;  // 0: one parameter set for all trajectories; NPAR: one row per trajectory ([B][NPAR])
	CallTotals *totals;
	// scratch vectors [B][n], L2-resident: stage argument, Wiener increments DW/DZ of the current attempt, first drift
	// (times h), proposal
	double *arg, *sW, *sZ, *sF1, *sNY;
	// multiplicative noise (SRI) only: the three further stage arguments and the diffusion values g2, g3
	double *argA, *argB, *argC, *sG2, *sG3;
	// jumps (jitcsde_jump.integrate, _jitcsde.py:682-700): scheduling state per trajectory and the source of the variates
	double *next_jump;        // [B]
	unsigned char *jump_set;  // [B]
	long long *tape_pos;      // [B]
	const double *taus, *xis; // [B][tape_len] (tape mode)
	long long tape_len;
	int use_jumps, generated_jumps;
	unsigned long long jump_seed;
	long long jump_first;
	// opt-in: the coupling sums as FP64 tensor-core products (coupling_phase_dmma) instead of the exact register tile
	int dmma;
	// single-step surface (the reference core's get_next_step / get_p / accept_step, jitced_template.c:396-536): the last
	// proposal lives in sNY (new state), sErr (error estimate per component), h_last (its step) and cursor (noise items used)
	double *sErr;    // [B][n]
	double *h_last;  // [B]
	int *cursor;     // [B]
};

// what one launch of k_wide_integrate does
enum { WIDE_INTEGRATE = 0, WIDE_PROPOSE = 1, WIDE_PIN = 2 };
struct WideMode {
	int mode;
	const double *h;      // WIDE_PROPOSE: step size per trajectory (get_next_step, jitced_template.c:396-406)
	unsigned pin_number;  // WIDE_PIN: items to append (pin_noise, jitced_template.c:252-267)
	double pin_step;
};

__global__ void k_wide_seed(uint32_t *key, int *pos, const uint32_t *seeds, int64_t B)
{
	const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= B)
		return;
	uint32_t *w = key + k * MT_WORDS;  // rk_seed (random_numbers.c:41-52)
	uint32_t v = seeds[k];
	for (int i = 0; i < MT_WORDS; i++) {
		w[i] = v;
		v = 1812433253u * (v ^ (v >> 30)) + (uint32_t)i + 1u;
	}
	pos[k] = MT_WORDS;
}

// per-trajectory working set in shared memory
struct alignas(16) TrajShared {
	alignas(16) uint32_t key[MT_WORDS];
	double qh[MAX_DEPTH];
	int order[MAX_DEPTH];
	double warp_max[WWARPS];
	double t, dt, h, target;
	int pos, count, cursor, status, running, last, accept, max_depth;
	unsigned acc, rej, fresh;
	int grid_index, sample_slot;  // sampling-grid mode: next sample time, and (>= 0) the sample the commit phase records
	// jumps: time of the next one (valid if jump_set), jumps applied so far (index into the tape / counter of the
	// generator), whether the step under way ends at a jump, whether a jump is due on the state P7 has just committed
	double next_jump;
	long long tape_pos;
	int jump_set, to_jump, pending_jump, tape_exhausted;
	unsigned jumps, pins;
};

// The reference's batch twist (random_numbers.c:73-84) by ONE warp on a state in shared memory.  Within a phase,
// word i needs the old words i and i+1 (and a word the phase does not touch), so every 32-word strip is read, the warp
// synchronised, and only then written.
// Words [LO, HI) of one phase: every lane reads the operands of its (up to eight) words, the warp synchronises, then the
// words are written — eight independent loads per lane in flight instead of one strip of 32 words at a time (the strip
// version spent ~3000 cycles per regeneration in shared-memory latency and warp barriers; eight regenerations per draw
// of 1000 pairs were half of the per-trajectory phase).
template <int LO, int HI, int TAP>
__device__ __forceinline__ void twist_phase(uint32_t *key, int lane)
{
	constexpr int PER_LANE = (HI - LO + 31) / 32;
	uint32_t v[PER_LANE];
#pragma unroll
	for (int u = 0; u < PER_LANE; u++) {
		const int i = LO + lane + 32 * u;
		v[u] = i < HI ? mt_twist(key[i], key[i + 1], key[i + TAP]) : 0u;
	}
	__syncwarp();
#pragma unroll
	for (int u = 0; u < PER_LANE; u++) {
		const int i = LO + lane + 32 * u;
		if (i < HI)
			key[i] = v[u];
	}
	__syncwarp();
}

__device__ __forceinline__ void twist_warp(uint32_t *key, int lane)
{
	twist_phase<0, 227, 397>(key, lane);     // kk < N-M: key[kk] = key[kk+M] ^ mix(key[kk], key[kk+1]); taps are old words
	twist_phase<227, 454, -227>(key, lane);  // kk < N-1: key[kk] = key[kk+(M-N)] ^ mix(key[kk], key[kk+1]); taps 0..226: new
	twist_phase<454, 623, -227>(key, lane);  // ... taps 227..395: new, written by the phase before
	if (lane == 0)
		key[623] = mt_twist(key[623], key[0], key[396]);
	__syncwarp();
}

// c_i = sum_j AT[j][i] * (y_j - y_i) for the block's T trajectories; in: SC[j][tr] = y_j of trajectory tr, out (same
// buffer): SC[i][tr] = c_i.  Whole block must call.
template <int T, int UNITS>
__device__ __forceinline__ void coupling_phase(const double *__restrict__ AT, double *SC, int tid)
{
	constexpr int IPT = (UNITS + WBLOCK - 1) / WBLOCK;  // rows per thread
	constexpr int TP2 = (T + 1) / 2;
	double yi[IPT][T], acc[IPT][T];
	int row[IPT];
#pragma unroll
	for (int q = 0; q < IPT; q++) {
		const int i = tid + q * WBLOCK;
		row[q] = i < UNITS ? i : UNITS - 1;  // surplus threads recompute the last row and discard it
#pragma unroll
		for (int tr = 0; tr < T; tr++) {
			yi[q][tr] = SC[row[q] * TPAD + tr];
			acc[q][tr] = 0.0;
		}
	}
	// few live trajectories: little arithmetic per loaded A_ij, the loop waits for L2 — keep more loads in flight
	constexpr int J_UNROLL = T <= 2 ? 16 : T <= 4 ? 8 : 4;
#pragma unroll J_UNROLL
	for (int j = 0; j < UNITS; j++) {
		double a[IPT];
#pragma unroll
		for (int q = 0; q < IPT; q++)
			a[q] = __ldg(AT + (size_t)j * UNITS + row[q]);
		double yj[2 * TP2];
		const double2 *src = reinterpret_cast<const double2 *>(SC + j * TPAD);
#pragma unroll
		for (int u = 0; u < TP2; u++) {
			const double2 v = src[u];  // same address for the whole warp: a broadcast
			yj[2 * u] = v.x;
			yj[2 * u + 1] = v.y;
		}
#pragma unroll
		for (int q = 0; q < IPT; q++)
#pragma unroll
			for (int tr = 0; tr < T; tr++)
				acc[q][tr] = acc[q][tr] + a[q] * (yj[tr] - yi[q][tr]);  // s += A*(Y[j]-Y[i]): three roundings, ascending j
	}
	__syncthreads();  // everybody has read the last y_j
#pragma unroll
	for (int q = 0; q < IPT; q++) {
		const int i = tid + q * WBLOCK;
		if (i < UNITS) {
#pragma unroll
			for (int tr = 0; tr < T; tr++)
				SC[i * TPAD + tr] = acc[q][tr];
		}
	}
	__syncthreads();
}

// ---- opt-in: the same sums on the FP64 tensor cores ("coupling = dmma") --------------------------------------------------
// c_i = sum_j A_ij (y_j - y_i) = (A V)_i - y_i * rowsum_i(A), with V = [y of the block's live trajectories | 0 ... | 1]:
// the product's last column is the row sum.  One warp-wide  mma.sync.aligned.m8n8k4.f64  multiplies an 8x4 tile of A
// (rows i, columns j) with a 4x8 tile of V (rows j, one column per trajectory) — 1 load of A per thread and 256 fused
// multiply-adds per instruction, where the exact path issues 3 x T unfused FP64 instructions per loaded A_ij.  The k-sum is
// fused and associated the hardware's way, so results differ from the reference's separately rounded loop in the last bits:
// this path is validated at BASELINE.json's TOLERANCE tier (1e-9 relative per trajectory where the step sequence matches,
// moments within 3 standard errors; tests/test_gpu_wide.py) and is never the default.  Needs a free column: active <= 7.
__device__ __forceinline__ void dmma_m8n8k4(double (&d)[2], double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

template <int UNITS>
__device__ __forceinline__ void coupling_phase_dmma(const double *__restrict__ AT, double *SC, int tid, int active)
{
	constexpr int ROW_TILES = (UNITS + 7) / 8;
	constexpr int TPW = (ROW_TILES + WWARPS - 1) / WWARPS;  // row tiles per warp
	constexpr int NT = TPAD / 8;                             // column tiles of the stage buffer
	const int lane = tid & 31, warp = tid >> 5;
	const int g = lane >> 2, q = lane & 3;  // fragment coordinates: A[g][q], B[q][g], D[g][2q], D[g][2q+1]
	const int ones = active;                // the first free column yields the row sums (active < TPAD)
	const int ntiles = ones / 8 + 1;        // column tiles that hold something
	for (int j = tid; j < UNITS; j += WBLOCK)
		SC[j * TPAD + ones] = 1.0;
	__syncthreads();
	double acc[TPW][NT][2];
#pragma unroll
	for (int m = 0; m < TPW; m++)
#pragma unroll
		for (int nt = 0; nt < NT; nt++)
			acc[m][nt][0] = acc[m][nt][1] = 0.0;
#pragma unroll 2
	for (int k0 = 0; k0 < UNITS; k0 += 4) {
		const int j = k0 + q;
		const bool jin = j < UNITS;
		double b[NT];
#pragma unroll
		for (int nt = 0; nt < NT; nt++)
			b[nt] = (jin && nt < ntiles) ? SC[j * TPAD + nt * 8 + g] : 0.0;
		const double *Aj = AT + (size_t)(jin ? j : 0) * UNITS;
#pragma unroll
		for (int m = 0; m < TPW; m++) {
			const int i = (warp + m * WWARPS) * 8 + g;
			const double a = (jin && i < UNITS) ? __ldg(Aj + i) : 0.0;
#pragma unroll
			for (int nt = 0; nt < NT; nt++)
				if (nt < ntiles)
					dmma_m8n8k4(acc[m][nt], a, b[nt]);
		}
	}
	double yi[TPW][NT][2];
#pragma unroll
	for (int m = 0; m < TPW; m++) {
		const int i = (warp + m * WWARPS) * 8 + g;
#pragma unroll
		for (int nt = 0; nt < NT; nt++) {
			yi[m][nt][0] = i < UNITS ? SC[i * TPAD + nt * 8 + 2 * q] : 0.0;
			yi[m][nt][1] = i < UNITS ? SC[i * TPAD + nt * 8 + 2 * q + 1] : 0.0;
		}
	}
	__syncthreads();  // everybody has read the last y_j
	const int ones_tile = ones >> 3, ones_lane = (lane & ~3) | ((ones & 7) >> 1), ones_odd = ones & 1;
#pragma unroll
	for (int m = 0; m < TPW; m++) {
		const int i = (warp + m * WWARPS) * 8 + g;
		double r0 = acc[m][0][0], r1 = acc[m][0][1];
#pragma unroll
		for (int nt = 1; nt < NT; nt++)
			if (nt == ones_tile) {
				r0 = acc[m][nt][0];
				r1 = acc[m][nt][1];
			}
		const double rowsum = __shfl_sync(0xffffffffu, ones_odd ? r1 : r0, ones_lane);  // D[g][ones]
		if (i < UNITS) {
#pragma unroll
			for (int nt = 0; nt < NT; nt++) {
				const int col = nt * 8 + 2 * q;
				if (col < active)
					SC[i * TPAD + col] = acc[m][nt][0] - yi[m][nt][0] * rowsum;
				if (col + 1 < active)
					SC[i * TPAD + col + 1] = acc[m][nt][1] - yi[m][nt][1] * rowsum;
			}
			if (q == 0)
				SC[i * TPAD + ones] = 0.0;  // columns of absent trajectories stay zero
		}
	}
	__syncthreads();
}

// the block's live trajectories occupy the first `active` columns: run the variant unrolled for exactly that many
template <int T, int UNITS, int K = 1>
__device__ __forceinline__ void coupling_dispatch(int active, const double *__restrict__ AT, double *SC, int tid)
{
	if constexpr (K <= T) {
		if (active == K)
			coupling_phase<K, UNITS>(AT, SC, tid);
		else
			coupling_dispatch<T, UNITS, K + 1>(active, AT, SC, tid);
	}
}

template <class Model, int T>
constexpr size_t wide_smem_bytes()
{
	return (size_t)(Model::UNITS > 0 ? Model::UNITS : 1) * TPAD * sizeof(double) + (size_t)T * sizeof(TrajShared);
}

// Sampling-grid mode (grid_times != nullptr): every trajectory integrates to grid_times[0], records its state in
// grid_out[0][b][·], goes on to grid_times[1], … — the sequence of integrate() calls of examples/noisy_lorenz.py:95-96
// (truncated last step + Brownian bridge at every sample time) in one launch.
template <class Model, int T>
__global__ void __launch_bounds__(WBLOCK, WBLOCKS_PER_SM) k_wide_integrate(WideView e, Controller c, double target,
	const double *__restrict__ grid_times, int grid_count, double *__restrict__ grid_out, WideMode wm)
{
	static_assert(T >= 1 && T <= TMAX && T <= WWARPS, "one warp per trajectory");
	constexpr int N = Model::N;
	constexpr int UNITS = Model::UNITS;  // the first UNITS components are densely coupled (0: no coupling term)
	constexpr int CPT = (N + WBLOCK - 1) / WBLOCK;
	extern __shared__ __align__(16) unsigned char wide_smem[];
	double *SC = reinterpret_cast<double *>(wide_smem);
	TrajShared *ts = reinterpret_cast<TrajShared *>(wide_smem + (size_t)(UNITS > 0 ? UNITS : 1) * TPAD * sizeof(double));

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int64_t b0 = (int64_t)blockIdx.x * T;
	const int live = (int)min((int64_t)T, e.B - b0);  // trajectories of this block
	const int D = e.D;
	const unsigned lt = (1u << lane) - 1u;
#ifdef JSDE_WIDE_TIMING  // development aid: where a block's time goes, phase by phase (printed by block 0)
	long long ph[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, ph_last = clock64(), sub_last = 0;
	int ph_rounds = 0;
#define JSDE_MARK(i) do { if (tid == 0) { const long long now_ = clock64(); ph[i] += now_ - ph_last; ph_last = now_; } } while (0)
#define JSDE_SUBMARK_START() do { if (tid == 0) sub_last = clock64(); } while (0)
#define JSDE_SUBMARK(i) do { if (tid == 0) { const long long now_ = clock64(); ph[i] += now_ - sub_last; sub_last = now_; } } while (0)
#else
#define JSDE_MARK(i) do { } while (0)
#define JSDE_SUBMARK_START() do { } while (0)
#define JSDE_SUBMARK(i) do { } while (0)
#endif

	// ---- load: warp w owns trajectory w --------------------------------------------------------------------------
	if (warp < T) {
		TrajShared &S = ts[warp];
		if (warp < live) {
			const int64_t b = b0 + warp;
			for (int i = lane; i < MT_WORDS; i += 32)
				S.key[i] = e.key[b * MT_WORDS + i];
			for (int d = lane; d < D; d += 32) {
				S.qh[d] = e.qh[b * D + d];
				S.order[d] = e.order[b * D + d];
			}
			if (lane == 0) {
				S.t = e.t[b];
				S.dt = e.dt[b];
				S.pos = e.pos[b];
				S.count = e.q_count[b];
				int status = e.status[b];
				if (status == TRAJ_MIN_STEP)
					status = TRAJ_OK;  // the reference raises once per call and lets the caller carry on (_jitcsde.py:569-579)
				S.status = status;
				S.target = grid_times ? grid_times[0] : target;
				S.grid_index = 0;
				S.sample_slot = -1;
				// integrate() with nothing to do (:614), sample times already reached and due jumps are handled in P1
				S.running = status == TRAJ_OK;
				S.next_jump = 0.0;
				S.tape_pos = 0;
				S.jump_set = S.to_jump = S.pending_jump = S.tape_exhausted = 0;
				S.jumps = 0;
				if (Model::HAS_JUMP && e.use_jumps) {
					S.next_jump = e.next_jump[b];
					S.jump_set = e.jump_set[b] != 0;
					S.tape_pos = e.tape_pos[b];
				}
				S.acc = S.rej = S.fresh = 0;
				S.pins = 0;
				S.max_depth = 0;
				S.cursor = 0;
				S.accept = 0;
			}
		} else if (lane == 0) {
			S.running = 0;
			S.accept = 0;
			S.sample_slot = -1;
			S.jumps = 0;
			S.pending_jump = S.tape_exhausted = 0;
			S.status = TRAJ_OK;
			S.acc = S.rej = S.fresh = 0;
			S.max_depth = 0;
		}
	}
	if (UNITS > 0)
		for (int i = tid; i < UNITS * TPAD; i += WBLOCK)
			SC[i] = 0.0;  // columns of absent trajectories stay zero
	__syncthreads();

	while (true) {
		int any = 0;
#pragma unroll
		for (int tr = 0; tr < T; tr++)
			any |= ts[tr].running;
		if (!any)
			break;

		// ---- P1: per trajectory, one warp each: step size, noise memory (get_noise, template:204-250), draw ----------
		if (warp < T && ts[warp].running) {
			TrajShared &S = ts[warp];
			const int64_t b = b0 + warp;
			double *const qw = e.qw + b * D * 2 * N;
			double *const sW = e.sW + b * N, *const sZ = e.sZ + b * N;
			const double t = S.t, dt = S.dt;
			// ---- what to integrate to (_jitcsde.py:614, :682-700; the sampling loop of examples/noisy_lorenz.py:95-96):
			// apply a jump that is due, schedule the next one, record sample times that are already reached — until an
			// attempt is due or the trajectory is done.  Every lane runs this on identical inputs; lane 0 publishes.
			double tgt = 0.0;
			bool still_running = true;
			if (wm.mode == WIDE_INTEGRATE) {
			double cur_target = S.target, next_jump = S.next_jump;
			tgt = cur_target;
			int gi = S.grid_index, jump_set = S.jump_set;
			long long tape_pos = S.tape_pos;
			unsigned jumped = 0;
			bool to_jump = false, apply = S.pending_jump != 0, exhausted = false;
			double *const Ycur = e.y + b * N;
			while (true) {
				if (Model::HAS_JUMP && apply) {  // apply_jump(amp(time, state)) (:696-698) on the pre-jump state
					const double xi = e.generated_jumps ? jsde_jump_xi(e.jump_seed, e.jump_first + b, tape_pos) : e.xis[b * e.tape_len + tape_pos];
					double *const change = e.sNY + b * N;
					for (int i = lane; i < N; i += 32)
						change[i] = Model::jump_component(i, t, Ycur, xi, e.par + b * e.par_stride);
					__syncwarp();
					for (int i = lane; i < N; i += 32)
						Ycur[i] += change[i];
					__syncwarp();
					tape_pos++;
					jumped++;
					jump_set = 0;
					apply = false;
				}
				tgt = cur_target;
				to_jump = false;
				if (Model::HAS_JUMP && e.use_jumps) {
					if (!jump_set) {  // next_jump property (:682-687): t + IJI(t, y)
						double wait = __longlong_as_double(0x7ff0000000000000LL);  // past the end of a tape: no further jumps
						if (e.generated_jumps)
							wait = Model::waiting_time(t, Ycur, jsde_jump_tau(e.jump_seed, e.jump_first + b, tape_pos), e.par + b * e.par_stride);
						else if (tape_pos < e.tape_len)
							wait = Model::waiting_time(t, Ycur, e.taus[b * e.tape_len + tape_pos], e.par + b * e.par_stride);
						else
							exhausted = true;  // reported (jsde_stats.n_tape_exhausted): this trajectory's jump process ends here
						next_jump = t + wait;
						jump_set = 1;
					}
					if (next_jump < cur_target) {
						tgt = next_jump;
						to_jump = true;
					}
				}
				if (!(t >= tgt))
					break;  // an attempt is due
				if (to_jump) {
					apply = true;
					continue;
				}
				if (grid_times) {  // this sample time is reached without a step: integrate() is a no-op, record the state
					double *out = grid_out + ((int64_t)gi * e.B + b) * N;
					for (int i = lane; i < N; i += 32)
						out[i] = Ycur[i];
					gi++;
					if (gi < grid_count) {
						cur_target = grid_times[gi];
						continue;
					}
				}
				still_running = false;
				break;
			}
			__syncwarp();
			if (lane == 0) {
				S.grid_index = gi;
				S.target = cur_target;
				S.next_jump = next_jump;
				S.jump_set = jump_set;
				S.tape_pos = tape_pos;
				S.to_jump = to_jump;
				S.pending_jump = 0;
				S.jumps += jumped;
				if (exhausted)
					S.tape_exhausted = 1;
				if (!still_running)
					S.running = 0;
			}
			}  // WIDE_INTEGRATE
			__syncwarp();
			if (still_running) {
			double h;
			bool last = false;
			if (wm.mode == WIDE_PROPOSE)
				h = wm.h[b];  // get_next_step(h): the caller's step, whatever the time is
			else if (wm.mode == WIDE_PIN)
				h = wm.pin_step;
			else if (t + dt < tgt)
				h = dt;
			else {
				h = tgt - t;
				last = true;
			}
			JSDE_SUBMARK_START();
			int count = S.count, pos = S.pos;
			bool started = false, overflow = false;
			double need = h, hq = 0.0;
			int cur = wm.mode == WIDE_PIN ? count : 0;  // pin_noise appends behind everything that is stored
			while (need != 0.0 && cur < count) {
				const int s = S.order[cur];
				hq = S.qh[s];
				if (!(hq <= need))
					break;
				// eight independent loads per array in flight, then the stores: written the plain way (load, add, store per
				// component) every iteration waited for its own L2 round trip — the compiler may not move a load above a
				// store that could alias it — and this loop alone took ~20 000 cycles per consumed item
				constexpr int BATCH = 8;
				const double *const itemW = qw + (s * 2 + 0) * N, *const itemZ = qw + (s * 2 + 1) * N;
				for (int i0 = lane; i0 < N; i0 += 32 * BATCH) {
					double w[BATCH], z[BATCH];
#pragma unroll
					for (int u = 0; u < BATCH; u++) {
						const int i = i0 + 32 * u;
						w[u] = i < N ? itemW[i] : 0.0;
						z[u] = i < N ? itemZ[i] : 0.0;
						if (started && i < N) {
							w[u] = sW[i] + w[u];
							z[u] = sZ[i] + z[u];
						}
					}
#pragma unroll
					for (int u = 0; u < BATCH; u++) {
						const int i = i0 + 32 * u;
						if (i < N) {
							sW[i] = w[u];
							sZ[i] = z[u];
						}
					}
				}
				started = true;
				need -= hq;
				cur++;
			}
			__syncwarp();  // the draw below hands component k to an arbitrary lane: make the partial sums visible
			JSDE_SUBMARK(8);
			if (need != 0.0) {
				if (count >= D)
					overflow = true;
				else {
					const bool bridging = cur < count;
					double h_exc = 0.0, factor = 0.0, variance = need;  // append_noise (template:170-176)
					if (bridging) {  // Brownian_bridge (template:178-202); hq is the too-long item's length
						h_exc = hq - need;
						factor = h_exc / hq;
						variance = factor * need;
					}
					const double scale = sqrt(variance);
					const int snew = S.order[count];  // a free physical slot
					const int s1 = bridging ? S.order[cur] : snew;
					// n pairs in stream order (get_gauss, template:145-154): pair k belongs to component k.  Two passes, so that the
					// sequential part (which candidate is the k-th accepted one) stays light and the heavy part (log, division,
					// square root, bridge: ~100 dependent FP64 instructions per pair) runs four pairs per lane at a time:
					// pass 1 parks the raw pairs (x1, x2) in the new item's own storage and r2 in the stage-argument scratch
					// (free until P3), pass 2 turns them into increments.  Round 1 did both per 32-candidate strip and spent
					// about a third of the n = 1000 kernel's time with seven warps waiting for the slowest draw.
					double *const raw1 = qw + (snew * 2 + 0) * N, *const raw2 = qw + (snew * 2 + 1) * N, *const rawr = e.arg + b * N;
					int produced = 0;
					const uint4 *const chunks = reinterpret_cast<const uint4 *>(S.key);
					while (produced < N) {
						if (pos == MT_WORDS) {
							twist_warp(S.key, lane);
							pos = 0;
						}
						// two strips of 32 candidates per turn (two independent temper/polar chains per lane); the second one
						// counts only if the first is used up entirely
						const int c0 = pos >> 2;
						bool ok[2];
						double x1[2], x2[2], r2[2];
#pragma unroll
						for (int u = 0; u < 2; u++) {
							const int cand = c0 + 32 * u + lane;
							ok[u] = false;
							x1[u] = x2[u] = r2[u] = 0.0;
							if (cand < CANDIDATES) {
								const uint4 kw = chunks[cand];
								uint4 w;
								w.x = mt_temper(kw.x);
								w.y = mt_temper(kw.y);
								w.z = mt_temper(kw.z);
								w.w = mt_temper(kw.w);
								ok[u] = polar_candidate(w, x1[u], x2[u], r2[u]);
							}
						}
#pragma unroll
						for (int u = 0; u < 2; u++) {
							const int cu = c0 + 32 * u;
							if (u == 1 && (cu >= CANDIDATES || produced >= N || pos != 4 * cu))
								break;  // nothing left in this block, enough pairs, or the first strip was not used up
							const unsigned okm = __ballot_sync(0xffffffffu, ok[u]);
							const int navail = __popc(okm);
							const int take = min(navail, N - produced);
							const int rank = __popc(okm & lt);
							if (ok[u] && rank < take) {
								const int k = produced + rank;
								raw1[k] = x1[u];
								raw2[k] = x2[u];
								rawr[k] = r2[u];
							}
							// candidates up to the last pair taken are consumed; rejected ones behind it would be rejected
							// again by the next draw, so a fully used strip may be skipped as a whole
							if (take == navail)
								pos += 4 * min(32, CANDIDATES - cu);
							else
								pos += 4 * ((int)__fns(okm, 0, take) + 1);
							produced += take;
						}
					}
					__syncwarp();
					JSDE_SUBMARK(9);
					// pass 2: f = scale*sqrt(-2*log(r2)/r2), DW = x1*f, DZ = x2*f (random_numbers.c:127-129), bridge, sums
					constexpr int GROUP = 4;
					for (int k0 = lane; k0 < N; k0 += 32 * GROUP) {
						double x1[GROUP], x2[GROUP], r2[GROUP], lg[GROUP], f[GROUP];
						unsigned pending = 0;
#pragma unroll
						for (int u = 0; u < GROUP; u++) {
							const int k = k0 + 32 * u;
							const bool in = k < N;
							x1[u] = in ? raw1[k] : 0.0;
							x2[u] = in ? raw2[k] : 0.0;
							r2[u] = in ? rawr[k] : 0.5;
							lg[u] = jsde_log_table_from(r2[u], nullptr);
							if (jsde_log_is_near_one(r2[u]))
								pending |= 1u << u;
						}
						while (pending) {  // glibc's near-one branch: rare, once per pair that needs it
							const int which = __ffs(pending) - 1;
							pending &= pending - 1;
							double x = r2[0];
#pragma unroll
							for (int u = 1; u < GROUP; u++)
								if (u == which)
									x = r2[u];
							const double v = jsde_log_near_one(x);
#pragma unroll
							for (int u = 0; u < GROUP; u++)
								if (u == which)
									lg[u] = v;
						}
						bool tame = true;
#pragma unroll
						for (int u = 0; u < GROUP; u++)
							f[u] = scale * inline_sqrt(inline_div(__dmul_rn(-2.0, lg[u]), r2[u], tame), tame);
						if (!tame) {  // cold: operands near the ends of the exponent range
#pragma unroll
							for (int u = 0; u < GROUP; u++)
								f[u] = scale * sqrt(__ddiv_rn(__dmul_rn(-2.0, lg[u]), r2[u]));
						}
						// all loads of the group (the item being bridged, the partial sums) before any store, see the scan above
						double w1[GROUP], z1[GROUP], aw[GROUP], az[GROUP];
#pragma unroll
						for (int u = 0; u < GROUP; u++) {
							const int k = k0 + 32 * u;
							const bool in = k < N;
							w1[u] = (in && bridging) ? qw[(s1 * 2 + 0) * N + k] : 0.0;
							z1[u] = (in && bridging) ? qw[(s1 * 2 + 1) * N + k] : 0.0;
							aw[u] = (in && started) ? sW[k] : 0.0;
							az[u] = (in && started) ? sZ[k] : 0.0;
						}
#pragma unroll
						for (int u = 0; u < GROUP; u++) {
							const int k = k0 + 32 * u;
							if (k < N) {
								double gw = x1[u] * f[u], gz = x2[u] * f[u];
								if (bridging) {
									const double w2 = gw + w1[u] * factor, z2 = gz + z1[u] * factor;
									gw = w1[u] - w2;
									gz = z1[u] - z2;
									raw1[k] = w2;
									raw2[k] = z2;
									qw[(s1 * 2 + 0) * N + k] = gw;
									qw[(s1 * 2 + 1) * N + k] = gz;
								} else {
									raw1[k] = gw;
									raw2[k] = gz;
								}
								sW[k] = started ? aw[u] + gw : gw;
								sZ[k] = started ? az[u] + gz : gz;
							}
						}
					}
					__syncwarp();
					JSDE_SUBMARK(10);
					if (lane == 0) {
						if (bridging) {  // relink: the new second part goes right behind the cursor
							for (int d = count; d > cur + 1; d--)
								S.order[d] = S.order[d - 1];
							S.order[cur + 1] = snew;
							S.qh[s1] = need;
							S.qh[snew] = h_exc;
						} else
							S.qh[snew] = need;
					}
					count++;
					started = true;
					cur++;
				}
			}
			if (!started && !overflow)
				for (int i = lane; i < N; i += 32)
					sW[i] = sZ[i] = 0.0;
			if (lane == 0) {
				S.h = h;
				S.last = last;
				S.pos = pos;
				if (overflow) {
					S.status = TRAJ_NOISE_OVERFLOW;
					S.running = 0;
				} else {
					S.fresh += (count != S.count);
					S.count = count;
					S.cursor = cur;
					S.max_depth = count > S.max_depth ? count : S.max_depth;
				}
			}
			}
		}
		__syncthreads();
		JSDE_MARK(0);
		if (wm.mode == WIDE_PIN) {  // pin_noise (template:252-267): the appended item becomes the current one, nothing else happens
			if (tid < T && ts[tid].running) {
				TrajShared &S = ts[tid];
				S.cursor = S.count - 1;
				if (++S.pins >= wm.pin_number)
					S.running = 0;
			}
			__syncthreads();
			continue;
		}

		// Trajectories that attempt a step in this round, as a bit mask (uniform over the block).  Their coupled
		// components occupy the FIRST popc(mask) columns of the stage buffer, so that a block whose trajectories have
		// partly finished (step counts differ by a factor of three across an ensemble) only pays for the live ones.
		unsigned mask = 0;
#pragma unroll
		for (int tr = 0; tr < T; tr++)
			mask |= ts[tr].running ? 1u << tr : 0u;
		const int active = __popc(mask);

		// ---- P0: stage the coupled part of every running trajectory's state -----------------------------------------
		if (UNITS > 0) {
			int col = 0;
			for (unsigned m = mask; m; m &= m - 1, col++) {
				const double *Y = e.y + (b0 + (__ffs(m) - 1)) * N;
				for (int i = tid; i < UNITS; i += WBLOCK)
					SC[i * TPAD + col] = Y[i];
			}
			__syncthreads();
		}
		JSDE_MARK(1);

		for (int stage = 0; stage < 2; stage++) {
			if (UNITS > 0) {
				if (e.dmma && active < TPAD)
					coupling_phase_dmma<(UNITS > 0 ? UNITS : 8)>(Model::coupling_table(), SC, tid, active);
				else
					coupling_dispatch<T, UNITS>(active, Model::coupling_table(), SC, tid);
			}
			JSDE_MARK(2 + 2 * stage);
			if (stage == 0) {
				// ---- P3: first stage (template:468-481) ----------------------------------------------------------------
				int col = 0;
				for (unsigned m = mask; m; m &= m - 1, col++) {
					const int tr = __ffs(m) - 1;
					const int64_t b = b0 + tr;
					const double *Y = e.y + b * N;
					const double t = ts[tr].t, h = ts[tr].h;
					// results of the thread's CPT components first, stores afterwards: nothing that could alias stands between
					// the components' loads, so they are in flight together.  (A component-major order — one component for all
					// trajectories of the block at a time — was measured too: 20 % slower in these phases.)
					double fh1v[CPT], av[CPT], aAv[CPT], aBv[CPT];
#pragma unroll
					for (int cc = 0; cc < CPT; cc++) {
						const int i = cc * WBLOCK + tid;
						fh1v[cc] = av[cc] = aAv[cc] = aBv[cc] = 0.0;
						if (i < N) {
							const double cs = (UNITS > 0 && i < UNITS) ? SC[i * TPAD + col] : 0.0;
							const double I10 = (e.sW[b * N + i] + e.sZ[b * N + i] * (1. / 1.7320508075688772)) * (h / 2.);
							const double fh1 = Model::drift_local(i, t, Y, e.par + b * e.par_stride, cs) * h;
							if constexpr (Model::ADDITIVE) {
								const double g1 = Model::diffusion_component(i, t + h, Y, e.par + b * e.par_stride);
								av[cc] = Y[i] + 0.75 * fh1 + 0.5 * g1 * I10 / h;
							} else {  // SRIW1 (template:417-438): the arguments of fh2, g2 and g3 only need fh1 and g1
								const double g1 = Model::diffusion_component(i, t, Y, e.par + b * e.par_stride);
								const double sqh = sqrt(h);
								av[cc] = Y[i] + 0.75 * fh1 + 1.5 * g1 * I10 / h;
								aAv[cc] = Y[i] + 0.25 * fh1 + 0.5 * g1 * sqh;
								aBv[cc] = Y[i] + fh1 - g1 * sqh;
							}
							fh1v[cc] = fh1;
						}
					}
#pragma unroll
					for (int cc = 0; cc < CPT; cc++) {
						const int i = cc * WBLOCK + tid;
						if (i < N) {
							if constexpr (!Model::ADDITIVE) {
								e.argA[b * N + i] = aAv[cc];
								e.argB[b * N + i] = aBv[cc];
							}
							e.sF1[b * N + i] = fh1v[cc];
							e.arg[b * N + i] = av[cc];
							if (UNITS > 0 && i < UNITS)
								SC[i * TPAD + col] = av[cc];  // this thread's own cell: read above, rewritten here
						}
					}
				}
			} else {
				if constexpr (!Model::ADDITIVE) {
					// ---- P4 (SRI only): g2, g3 and the argument of g4 (template:439-445) ---------------------------------
					for (unsigned m = mask; m; m &= m - 1) {
						const int tr = __ffs(m) - 1;
						const int64_t b = b0 + tr;
						const double *Y = e.y + b * N;
						const double t = ts[tr].t, h = ts[tr].h;
						const double sqh = sqrt(h);
						double g2v[CPT], g3v[CPT], aCv[CPT];
#pragma unroll
						for (int cc = 0; cc < CPT; cc++) {
							const int i = cc * WBLOCK + tid;
							g2v[cc] = g3v[cc] = aCv[cc] = 0.0;
							if (i < N) {
								const double g1 = Model::diffusion_component(i, t, Y, e.par + b * e.par_stride);
								g2v[cc] = Model::diffusion_component(i, t + 0.25 * h, e.argA + b * N, e.par + b * e.par_stride);
								g3v[cc] = Model::diffusion_component(i, t + h, e.argB + b * N, e.par + b * e.par_stride);
								aCv[cc] = Y[i] + 0.25 * e.sF1[b * N + i] + sqh * (-5 * g1 + 3 * g2v[cc] + 0.5 * g3v[cc]);
							}
						}
#pragma unroll
						for (int cc = 0; cc < CPT; cc++) {
							const int i = cc * WBLOCK + tid;
							if (i < N) {
								e.sG2[b * N + i] = g2v[cc];
								e.sG3[b * N + i] = g3v[cc];
								e.argC[b * N + i] = aCv[cc];
							}
						}
					}
					__syncthreads();
				}
				// ---- P5: second stage, proposal, error ratio (template:482-526) -----------------------------------------
				int col = 0;
				for (unsigned m = mask; m; m &= m - 1, col++) {
					const int tr = __ffs(m) - 1;
					const int64_t b = b0 + tr;
					const double *Y = e.y + b * N, *A = e.arg + b * N;
					const double t = ts[tr].t, h = ts[tr].h;
					const double t_mid = t + 0.75 * h;
					double worst = 0.0;
					double nyv[CPT], erv[CPT];
#pragma unroll
					for (int cc = 0; cc < CPT; cc++) {
						const int i = cc * WBLOCK + tid;
						nyv[cc] = erv[cc] = 0.0;
						if (i < N) {
							const double cs = (UNITS > 0 && i < UNITS) ? SC[i * TPAD + col] : 0.0;
							const double W = e.sW[b * N + i];
							const double I10 = (W + e.sZ[b * N + i] * (1. / 1.7320508075688772)) * (h / 2.);
							const double fh1 = e.sF1[b * N + i];
							const double fh2 = Model::drift_local(i, t_mid, A, e.par + b * e.par_stride, cs) * h;
							double E_N, ny;
							if constexpr (Model::ADDITIVE) {
								const double g1 = Model::diffusion_component(i, t + h, Y, e.par + b * e.par_stride);
								const double g2 = Model::diffusion_component(i, t, Y, e.par + b * e.par_stride);
								E_N = I10 / h * (g2 - g1);
								ny = Y[i] + E_N + (1. / 3.) * (fh1 + 2 * fh2) + W * g1;
							} else {  // SRIW1 (template:447-464), expression shapes of stepper.cuh
								const double g1 = Model::diffusion_component(i, t, Y, e.par + b * e.par_stride);
								const double g2 = e.sG2[b * N + i], g3 = e.sG3[b * N + i];
								const double g4 = Model::diffusion_component(i, t + 0.25 * h, e.argC + b * N, e.par + b * e.par_stride);
								const double sqh = sqrt(h);
								const double third_over_h = 1. / h / 3.;
								const double I11s = (W * W - h) * 0.5 / sqh;
								const double I111 = (W * W * W - 3 * h * W) * (1. / 6.);
								E_N = third_over_h * (
									+ (6 * I10 - 6 * I111) * g1
									+ (-4 * I10 + 5 * I111) * g2
									+ (-2 * I10 - 2 * I111) * g3
									+ (3 * I111) * g4);
								ny = Y[i] + E_N + (1. / 3.) * (
									fh1 + 2 * fh2
									+ (-3 * W - 3 * I11s) * g1
									+ (4 * W + 4 * I11s) * g2
									+ (2 * W - I11s) * g3);
							}
							const double E_D = (fh2 - fh1) / 6;
							const double er = fabs(E_D) + fabs(E_N);
							nyv[cc] = ny;
							erv[cc] = er;
							const double tol = c.atol + c.rtol * fabs(ny);  // get_p (template:502-526)
							if (er != 0.0 || tol != 0.0) {
								const double x = er / tol;
								if (x > worst)
									worst = x;
							}
						}
					}
#pragma unroll
					for (int cc = 0; cc < CPT; cc++) {
						const int i = cc * WBLOCK + tid;
						if (i < N) {
							e.sNY[b * N + i] = nyv[cc];
							if (wm.mode == WIDE_PROPOSE)
								e.sErr[b * N + i] = erv[cc];  // get_p is a separate call there (template:502-526)
						}
					}
#pragma unroll
					for (int o = 16; o > 0; o >>= 1) {  // maximum: exact and order-independent
						const double other = __shfl_xor_sync(0xffffffffu, worst, o);
						worst = other > worst ? other : worst;
					}
					if (lane == 0)
						ts[tr].warp_max[warp] = worst;
				}
			}
			__syncthreads();
			JSDE_MARK(3 + 2 * stage);
		}

		// ---- P6: controller (_jitcsde.py:581-595) and commit bookkeeping, one warp per trajectory ----------------------
		if (wm.mode == WIDE_PROPOSE) {  // get_next_step ends here: the proposal waits for get_p / accept_step
			if (tid < T)
				ts[tid].running = 0;
		} else if (warp < T && ts[warp].running) {
			TrajShared &S = ts[warp];
			double p = 0.0;
#pragma unroll
			for (int w = 0; w < WWARPS; w++)
				p = S.warp_max[w] > p ? S.warp_max[w] : p;
			const double h = S.h;
			double dt = S.dt;
			const bool rejecting = p > c.dec_thr;
			const bool growing = !rejecting && p <= c.inc_thr;
			double pw = 0.0;
			if ((rejecting && p <= 1.7976931348623157e308) || (growing && p != 0.0))
				pw = jsde_pow(p, rejecting ? c.shrink_exp : c.grow_exp);
			__syncwarp();
			if (lane == 0) {
				if (rejecting) {
					const double shrink = c.safety * pw;
					dt = h * (c.min_factor > shrink ? c.min_factor : shrink);
					S.rej++;
					S.accept = 0;
					if (dt < c.min_step) {
						S.status = TRAJ_MIN_STEP;
						S.running = 0;
					}
				} else {
					if (growing) {
						const double factor = (p != 0.0) ? c.safety * pw : c.max_factor;
						const double sug = h * (c.max_factor < factor ? c.max_factor : factor);
						const double capped = c.max_step < sug ? c.max_step : sug;
						dt = capped > dt ? capped : dt;
					}
					// accept_step (template:528-536): drop the consumed items — rotate their slots behind the ones in use
					const int cursor = S.cursor, count = S.count;
					if (cursor > 0) {
						int freed[MAX_DEPTH];
						for (int d = 0; d < cursor; d++)
							freed[d] = S.order[d];
						for (int d = cursor; d < count; d++)
							S.order[d - cursor] = S.order[d];
						for (int d = 0; d < cursor; d++)
							S.order[count - cursor + d] = freed[d];
					}
					S.count = count - cursor;
					S.t = S.t + h;
					S.acc++;
					S.accept = 1;
					if (S.last && S.to_jump)
						S.pending_jump = 1;  // the step ended at a jump time: applied on the committed state at the next P1
					else if (S.last) {
						if (grid_times) {  // the commit phase records this sample; then on to the next sample time
							S.sample_slot = S.grid_index;
							S.grid_index++;
							if (S.grid_index < grid_count)
								S.target = grid_times[S.grid_index];
							else
								S.running = 0;
						} else
							S.running = 0;
					}
				}
				S.dt = dt;
			}
		}
		__syncthreads();

		JSDE_MARK(6);
		// ---- P7: commit the accepted proposals ---------------------------------------------------------------------------
		for (int tr = 0; tr < T; tr++) {
			if (!ts[tr].accept)
				continue;
			const int64_t b = b0 + tr;
			const int slot = ts[tr].sample_slot;
			double v[CPT];
#pragma unroll
			for (int cc = 0; cc < CPT; cc++) {
				const int i = cc * WBLOCK + tid;
				v[cc] = i < N ? e.sNY[b * N + i] : 0.0;
			}
#pragma unroll
			for (int cc = 0; cc < CPT; cc++) {
				const int i = cc * WBLOCK + tid;
				if (i < N) {
					e.y[b * N + i] = v[cc];
					if (slot >= 0)
						grid_out[((int64_t)slot * e.B + b) * N + i] = v[cc];
				}
			}
		}
		__syncthreads();
		if (tid < T) {
			ts[tid].accept = 0;
			ts[tid].sample_slot = -1;
		}
		__syncthreads();
		JSDE_MARK(7);
#ifdef JSDE_WIDE_TIMING
		ph_rounds++;
#endif
	}
#ifdef JSDE_WIDE_TIMING
	if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == 100))
		printf("block %d: %d rounds; cycles per round: P1 %lld (trajectory 0: scan %lld, draw pass 1 %lld, pass 2 %lld) | stage-in %lld | coupling1 %lld | P3 %lld | coupling2 %lld | P5 %lld | P6 %lld | P7 %lld\n",
			blockIdx.x, ph_rounds, ph[0] / max(ph_rounds, 1), ph[8] / max(ph_rounds, 1), ph[9] / max(ph_rounds, 1), ph[10] / max(ph_rounds, 1),
			ph[1] / max(ph_rounds, 1), ph[2] / max(ph_rounds, 1), ph[3] / max(ph_rounds, 1),
			ph[4] / max(ph_rounds, 1), ph[5] / max(ph_rounds, 1), ph[6] / max(ph_rounds, 1), ph[7] / max(ph_rounds, 1));
#endif

	// ---- store ----------------------------------------------------------------------------------------------------
	if (warp < live) {
		TrajShared &S = ts[warp];
		const int64_t b = b0 + warp;
		for (int i = lane; i < MT_WORDS; i += 32)
			e.key[b * MT_WORDS + i] = S.key[i];
		for (int d = lane; d < D; d += 32) {
			e.qh[b * D + d] = S.qh[d];
			e.order[b * D + d] = S.order[d];
		}
		if (lane == 0) {
			e.t[b] = S.t;
			e.dt[b] = S.dt;
			e.pos[b] = S.pos;
			e.q_count[b] = S.count;
			e.cursor[b] = wm.mode == WIDE_INTEGRATE ? 0 : S.cursor;
			if (wm.mode == WIDE_PROPOSE)
				e.h_last[b] = S.h;
			e.status[b] = S.status;
			e.accepted[b] += S.acc;
			e.rejected[b] += S.rej;
			if (Model::HAS_JUMP && e.use_jumps) {
				e.next_jump[b] = S.next_jump;
				e.jump_set[b] = S.jump_set ? 1 : 0;
				e.tape_pos[b] = S.tape_pos;
			}
			if (S.jumps) atomicAdd(&e.totals->jumps, (unsigned long long)S.jumps);
			if (S.tape_exhausted) atomicAdd(&e.totals->tape_exhausted, 1ull);
			if (S.acc) atomicAdd(&e.totals->accepted, (unsigned long long)S.acc);
			if (S.rej) atomicAdd(&e.totals->rejected, (unsigned long long)S.rej);
			if (S.fresh) atomicAdd(&e.totals->fresh_draws, (unsigned long long)S.fresh);
			if (S.status == TRAJ_MIN_STEP) atomicAdd((unsigned long long *)&e.totals->n_min_step, 1ull);
			if (S.status == TRAJ_NOISE_OVERFLOW) atomicAdd((unsigned long long *)&e.totals->n_overflow, 1ull);
			atomicMax(&e.totals->max_depth, S.max_depth);
		}
	}
}

// get_p (template:502-526) on the stored proposal: one block per trajectory
template <int N>
__global__ void __launch_bounds__(256) k_wide_get_p(WideView e, double atol, double rtol, double *p)
{
	__shared__ double part[8];
	const int64_t b = blockIdx.x;
	double worst = 0.0;
	for (int i = threadIdx.x; i < N; i += 256) {
		const double er = e.sErr[b * N + i];
		const double tol = atol + rtol * fabs(e.sNY[b * N + i]);
		if (er != 0.0 || tol != 0.0) {  // a 0/0 component is skipped; a NaN never wins the comparison
			const double x = er / tol;
			if (x > worst)
				worst = x;
		}
	}
	for (int o = 16; o > 0; o >>= 1) {
		const double other = __shfl_xor_sync(0xffffffffu, worst, o);
		worst = other > worst ? other : worst;
	}
	if ((threadIdx.x & 31) == 0)
		part[threadIdx.x >> 5] = worst;
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int w = 1; w < 8; w++)
			worst = part[w] > worst ? part[w] : worst;
		p[b] = worst;
	}
}

// accept_step (template:528-536): commit the proposal, drop the consumed noise items (their slots rotate behind the ones in use)
template <int N>
__global__ void __launch_bounds__(256) k_wide_accept(WideView e, const unsigned char *mask)
{
	const int64_t b = blockIdx.x;
	if (mask && !mask[b])
		return;
	for (int i = threadIdx.x; i < N; i += 256)
		e.y[b * N + i] = e.sNY[b * N + i];
	if (threadIdx.x == 0) {
		e.t[b] = e.t[b] + e.h_last[b];
		const int cursor = e.cursor[b], count = e.q_count[b];
		int *order = e.order + b * e.D;
		if (cursor > 0) {
			int freed[MAX_DEPTH];
			for (int d = 0; d < cursor; d++)
				freed[d] = order[d];
			for (int d = cursor; d < count; d++)
				order[d - cursor] = order[d];
			for (int d = 0; d < cursor; d++)
				order[count - cursor + d] = freed[d];
		}
		e.q_count[b] = count - cursor;
		e.cursor[b] = 0;
	}
}

// pin_noise with caller-supplied increments (jsde_pin_noise_path): item j has length steps[j], DW = dw[b][j][.], DZ = dz[b][j][.]
template <int N>
__global__ void __launch_bounds__(256) k_wide_pin_path(WideView e, unsigned number, const double *steps, const double *dw, const double *dz)
{
	const int64_t b = blockIdx.x;
	if (e.status[b] != TRAJ_OK)
		return;
	__shared__ int slot_s, stop_s;
	for (unsigned j = 0; j < number; j++) {
		if (threadIdx.x == 0) {
			const int count = e.q_count[b];
			stop_s = count >= e.D;
			if (stop_s)
				e.status[b] = TRAJ_NOISE_OVERFLOW;
			else {
				slot_s = e.order[b * e.D + count];
				e.qh[b * e.D + slot_s] = steps[j];
				e.q_count[b] = count + 1;
				e.cursor[b] = count;  // the newest item is the current one (template:173)
			}
		}
		__syncthreads();
		if (stop_s)
			return;
		double *item = e.qw + (b * e.D + slot_s) * 2 * N;
		for (int i = threadIdx.x; i < N; i += 256) {
			item[i] = dw[(b * number + j) * N + i];
			item[N + i] = dz[(b * number + j) * N + i];
		}
		__syncthreads();
	}
}

__global__ void k_wide_add(double *y, const double *change, int64_t count)
{
	const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx < count)
		y[idx] += change[idx];
}

__global__ void k_iota_rows(int *order, int64_t B, int D)
{
	const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx < B * D)
		order[idx] = (int)(idx % D);
}

}  // namespace wide
}  // namespace jsde
