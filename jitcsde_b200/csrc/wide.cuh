// "Wide" path: one thread block per trajectory, the n components spread over the threads (SURVEY.md §8 d config 5:
// noisy FitzHugh–Nagumo network, n = 1000, dense 500x500 coupling).  The lane-per-trajectory kernels keep a whole
// state in one thread's registers, which stops making sense beyond a handful of components; here the state lives in
// shared memory, every thread owns components tid, tid+256, ..., and the parts of the algorithm that are scalar per
// trajectory (noise-memory bookkeeping, controller) are evaluated redundantly by all threads on identical inputs, so
// that no broadcast is needed and control flow stays uniform.
//
// Same algorithm and the same bits as the reference, restated for this shape:
//   generator ........ jitcsde/random_numbers.c:41-130.  One trajectory needs n pairs per draw from ONE sequential
//                      stream, so the block regenerates the whole 624-word state at once (the reference's batch twist,
//                      :73-84, in three data-parallel phases), forms the block's 156 polar candidates in parallel,
//                      compacts the accepted ones in stream order (ballot + prefix sum) and hands pair k of a draw to
//                      the thread that owns component k.  Accepted pairs beyond a draw's need stay queued.
//   noise memory ..... jitced_template.c:170-250; items are addressed through a small permutation (logical position ->
//                      physical slot) so that a Brownian-bridge insertion or a pop never moves 16 KB items.
//   SRA1 step ........ jitced_template.c:466-494;   error ratio / commit ... :502-536;   controller ... _jitcsde.py:569-630
// Additive noise only (the network of config 5); no jumps, no sampling grid yet.
#pragma once

#include "ensemble.cuh"
#include "exact_div.cuh"
#include "gauss.cuh"
#include "mt19937.cuh"

namespace jsde {
namespace wide {

constexpr int WBLOCK = 256;
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
	double *fifo;      // [B][3][156]   accepted pairs (x1, x2, r2) of the current block not yet consumed
	int *fifo_rd, *fifo_cnt;  // [B]
	int *status;
	unsigned long long *accepted, *rejected;
	const double *par;
	CallTotals *totals;
};

__global__ void k_wide_seed(uint32_t *key, int *fifo_rd, int *fifo_cnt, const uint32_t *seeds, int64_t B)
{
	const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= B)
		return;
	uint32_t *w = key + k * MT_WORDS;  // rk_seed (random_numbers.c:41-52); pos = 624 <=> nothing queued
	uint32_t v = seeds[k];
	for (int i = 0; i < MT_WORDS; i++) {
		w[i] = v;
		v = 1812433253u * (v ^ (v >> 30)) + (uint32_t)i + 1u;
	}
	fifo_rd[k] = 0;
	fifo_cnt[k] = 0;
}

// The reference's batch twist (random_numbers.c:73-84) on a state held in shared memory; whole block must call.
__device__ __forceinline__ void twist_block(uint32_t *key, int tid)
{
	uint32_t v = 0;
	if (tid < 227)
		v = key[tid + 397] ^ mt_mix(key[tid], key[tid + 1]);
	__syncthreads();
	if (tid < 227)
		key[tid] = v;
	__syncthreads();
	if (tid < 227)
		v = key[tid] ^ mt_mix(key[tid + 227], key[tid + 228]);
	__syncthreads();
	if (tid < 227)
		key[tid + 227] = v;
	__syncthreads();
	if (tid < 169)
		v = key[tid + 227] ^ mt_mix(key[tid + 454], key[tid + 455]);
	__syncthreads();
	if (tid < 169)
		key[tid + 454] = v;
	__syncthreads();
	if (tid == 0)
		key[623] = key[396] ^ mt_mix(key[623], key[0]);
	__syncthreads();
}

template <class Model>
__global__ void __launch_bounds__(WBLOCK) k_wide_integrate(WideView e, Controller c, double target)
{
	constexpr int N = Model::N;
	constexpr int CPT = (N + WBLOCK - 1) / WBLOCK;  // components per thread
	__shared__ double Ysh[N], Ash[N];
	__shared__ uint32_t key[MT_WORDS];
	__shared__ double fx1[CANDIDATES], fx2[CANDIDATES], fr2[CANDIDATES];
	__shared__ double s_qh[MAX_DEPTH];
	__shared__ int s_order[MAX_DEPTH];
	__shared__ int s_warp_count[WBLOCK / 32];
	__shared__ double s_warp_max[WBLOCK / 32];
	__shared__ double s_par[Model::NPAR > 0 ? Model::NPAR : 1];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int64_t b = blockIdx.x;
	const int D = e.D;
	double *const qw = e.qw + b * D * 2 * N;

	// ---- load -----------------------------------------------------------------------------------------------------
	double yv[CPT];
#pragma unroll
	for (int cc = 0; cc < CPT; cc++) {
		const int i = cc * WBLOCK + tid;
		yv[cc] = i < N ? e.y[b * N + i] : 0.0;
		if (i < N)
			Ysh[i] = yv[cc];
	}
	for (int i = tid; i < MT_WORDS; i += WBLOCK)
		key[i] = e.key[b * MT_WORDS + i];
	int f_rd = e.fifo_rd[b], f_cnt = e.fifo_cnt[b];
	for (int i = tid; i < f_cnt; i += WBLOCK) {
		fx1[i] = e.fifo[(b * 3 + 0) * CANDIDATES + i];
		fx2[i] = e.fifo[(b * 3 + 1) * CANDIDATES + i];
		fr2[i] = e.fifo[(b * 3 + 2) * CANDIDATES + i];
	}
	if (tid < D) {
		s_qh[tid] = e.qh[b * D + tid];
		s_order[tid] = e.order[b * D + tid];
	}
	if (tid < Model::NPAR)
		s_par[tid] = e.par[tid];
	double t = e.t[b], dt = e.dt[b];
	int count = e.q_count[b];
	int status = e.status[b];
	if (status == TRAJ_MIN_STEP)
		status = TRAJ_OK;  // the reference raises once per call and lets the caller carry on (_jitcsde.py:569-579)
	unsigned acc = 0, rej = 0, fresh = 0;
	int max_depth = 0;
	__syncthreads();

	bool running = status == TRAJ_OK && !(t >= target);  // integrate() with nothing to do (:614)
	while (running) {
		// ---- one attempted step (identical control flow in every thread) ------------------------------------------
		double h;
		bool last = false;
		if (t + dt < target)
			h = dt;
		else {
			h = target - t;
			last = true;
		}
		// noise memory: get_noise (template:204-250)
		double W[CPT], Z[CPT];
		bool started = false;
		double need = h;
		int cur = 0;
		double hq = 0.0;
		while (need != 0.0 && cur < count) {
			const int s = s_order[cur];
			hq = s_qh[s];
			if (!(hq <= need))
				break;
#pragma unroll
			for (int cc = 0; cc < CPT; cc++) {
				const int i = cc * WBLOCK + tid;
				if (i < N) {
					const double w = qw[(s * 2 + 0) * N + i], z = qw[(s * 2 + 1) * N + i];
					W[cc] = started ? W[cc] + w : w;
					Z[cc] = started ? Z[cc] + z : z;
				}
			}
			started = true;
			need -= hq;
			cur++;
		}
		if (need != 0.0) {
			if (count >= D) {
				status = TRAJ_NOISE_OVERFLOW;
				break;
			}
			const bool bridging = cur < count;
			double h_exc = 0.0, factor = 0.0, scale;
			if (bridging) {  // Brownian_bridge (template:178-202); hq is the too-long item's length
				h_exc = hq - need;
				factor = h_exc / hq;
				scale = sqrt(factor * need);
			} else  // append_noise (template:170-176)
				scale = sqrt(need);
			// ---- draw n pairs in stream order (get_gauss, template:145-154) ----------------------------------------
			double gW[CPT], gZ[CPT];
			int produced = 0;
			while (produced < N) {
				if (f_rd == f_cnt) {  // nothing queued: regenerate the generator block and form its 156 candidates
					__syncthreads();  // everybody is done reading the old queue
					twist_block(key, tid);
					bool ok = false;
					double x1 = 0.0, x2 = 0.0, r2 = 0.0;
					if (tid < CANDIDATES) {
						uint4 w;
						w.x = mt_temper(key[4 * tid]);
						w.y = mt_temper(key[4 * tid + 1]);
						w.z = mt_temper(key[4 * tid + 2]);
						w.w = mt_temper(key[4 * tid + 3]);
						ok = polar_candidate(w, x1, x2, r2);
					}
					const unsigned okm = __ballot_sync(0xffffffffu, ok);
					if (lane == 0)
						s_warp_count[warp] = __popc(okm);
					__syncthreads();
					int base = 0, total = 0;
#pragma unroll
					for (int w = 0; w < WBLOCK / 32; w++) {
						const int cw = s_warp_count[w];
						base += w < warp ? cw : 0;
						total += cw;
					}
					if (ok) {
						const int at = base + __popc(okm & ((1u << lane) - 1u));
						fx1[at] = x1;
						fx2[at] = x2;
						fr2[at] = r2;
					}
					f_rd = 0;
					f_cnt = total;
					__syncthreads();
				}
				const int take = min(f_cnt - f_rd, N - produced);
#pragma unroll
				for (int cc = 0; cc < CPT; cc++) {
					const int i = cc * WBLOCK + tid;
					if (i >= produced && i < produced + take) {  // pair (i - produced) of this batch is component i's
						const int at = f_rd + (i - produced);
						const double r2 = fr2[at];
						const double f = scale * polar_radius_factor(r2);
						gW[cc] = fx1[at] * f;
						gZ[cc] = fx2[at] * f;
					}
				}
				f_rd += take;
				produced += take;
			}
			fresh++;
			const int snew = s_order[count];  // a free physical slot
			if (bridging) {
				const int s1 = s_order[cur];
				__syncthreads();
				if (tid == 0) {  // relink: the new second part goes right behind the cursor
					for (int d = count; d > cur + 1; d--)
						s_order[d] = s_order[d - 1];
					s_order[cur + 1] = snew;
					s_qh[s1] = need;
					s_qh[snew] = h_exc;
				}
#pragma unroll
				for (int cc = 0; cc < CPT; cc++) {
					const int i = cc * WBLOCK + tid;
					if (i < N) {
						const double w1 = qw[(s1 * 2 + 0) * N + i], z1 = qw[(s1 * 2 + 1) * N + i];
						const double w2 = gW[cc] + w1 * factor, z2 = gZ[cc] + z1 * factor;
						gW[cc] = w1 - w2;
						gZ[cc] = z1 - z2;
						qw[(snew * 2 + 0) * N + i] = w2;
						qw[(snew * 2 + 1) * N + i] = z2;
						qw[(s1 * 2 + 0) * N + i] = gW[cc];
						qw[(s1 * 2 + 1) * N + i] = gZ[cc];
					}
				}
				__syncthreads();
			} else {
				if (tid == 0)
					s_qh[snew] = need;
#pragma unroll
				for (int cc = 0; cc < CPT; cc++) {
					const int i = cc * WBLOCK + tid;
					if (i < N) {
						qw[(snew * 2 + 0) * N + i] = gW[cc];
						qw[(snew * 2 + 1) * N + i] = gZ[cc];
					}
				}
				__syncthreads();
			}
			count++;
#pragma unroll
			for (int cc = 0; cc < CPT; cc++) {
				W[cc] = started ? W[cc] + gW[cc] : gW[cc];
				Z[cc] = started ? Z[cc] + gZ[cc] : gZ[cc];
			}
			started = true;
			cur++;
		}
		if (!started) {
#pragma unroll
			for (int cc = 0; cc < CPT; cc++)
				W[cc] = Z[cc] = 0.0;
		}
		const int cursor = cur;
		max_depth = count > max_depth ? count : max_depth;

		// ---- SRA1 (template:468-494) -------------------------------------------------------------------------------
		double fh1[CPT], fh2[CPT], g1[CPT], ny[CPT], er[CPT], I10[CPT];
#pragma unroll
		for (int cc = 0; cc < CPT; cc++) {
			const int i = cc * WBLOCK + tid;
			if (i < N) {
				I10[cc] = (W[cc] + Z[cc] * (1. / 1.7320508075688772)) * (h / 2.);
				fh1[cc] = Model::drift_component(i, t, Ysh, s_par) * h;
				g1[cc] = Model::diffusion_component(i, t + h, Ysh, s_par);
				Ash[i] = yv[cc] + 0.75 * fh1[cc] + 0.5 * g1[cc] * I10[cc] / h;
			}
		}
		__syncthreads();
		double worst = 0.0;
		const double t_mid = t + 0.75 * h;
#pragma unroll
		for (int cc = 0; cc < CPT; cc++) {
			const int i = cc * WBLOCK + tid;
			if (i < N) {
				fh2[cc] = Model::drift_component(i, t_mid, Ash, s_par) * h;
				const double g2 = Model::diffusion_component(i, t, Ysh, s_par);
				const double E_N = I10[cc] / h * (g2 - g1[cc]);
				ny[cc] = yv[cc] + E_N + (1. / 3.) * (fh1[cc] + 2 * fh2[cc]) + W[cc] * g1[cc];
				const double E_D = (fh2[cc] - fh1[cc]) / 6;
				er[cc] = fabs(E_D) + fabs(E_N);
				const double tol = c.atol + c.rtol * fabs(ny[cc]);  // get_p (template:502-526)
				if (er[cc] != 0.0 || tol != 0.0) {
					const double x = er[cc] / tol;
					if (x > worst)
						worst = x;
				}
			}
		}
		// block maximum (exact, order-independent)
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			const double other = __shfl_xor_sync(0xffffffffu, worst, o);
			worst = other > worst ? other : worst;
		}
		if (lane == 0)
			s_warp_max[warp] = worst;
		__syncthreads();
		double p = 0.0;
#pragma unroll
		for (int w = 0; w < WBLOCK / 32; w++)
			p = s_warp_max[w] > p ? s_warp_max[w] : p;

		// ---- controller (_jitcsde.py:581-595), evaluated identically by every thread -------------------------------
		const bool rejecting = p > c.dec_thr;
		const bool growing = !rejecting && p <= c.inc_thr;
		double pw = 0.0;
		if ((rejecting && p <= 1.7976931348623157e308) || (growing && p != 0.0))
			pw = jsde_pow(p, rejecting ? c.shrink_exp : c.grow_exp);
		if (rejecting) {
			const double shrink = c.safety * pw;
			dt = h * (c.min_factor > shrink ? c.min_factor : shrink);
			rej++;
			if (dt < c.min_step) {
				status = TRAJ_MIN_STEP;
				running = false;
			}
		} else {
			if (growing) {
				const double factor = (p != 0.0) ? c.safety * pw : c.max_factor;
				const double sug = h * (c.max_factor < factor ? c.max_factor : factor);
				const double capped = c.max_step < sug ? c.max_step : sug;
				dt = capped > dt ? capped : dt;
			}
			// accept_step (template:528-536)
#pragma unroll
			for (int cc = 0; cc < CPT; cc++) {
				const int i = cc * WBLOCK + tid;
				if (i < N) {
					yv[cc] = ny[cc];
					Ysh[i] = ny[cc];
				}
			}
			t = t + h;
			if (cursor > 0 && tid == 0) {  // drop the consumed items: rotate their slots behind the ones still in use
				int freed[MAX_DEPTH];
				for (int d = 0; d < cursor; d++)
					freed[d] = s_order[d];
				for (int d = cursor; d < count; d++)
					s_order[d - cursor] = s_order[d];
				for (int d = 0; d < cursor; d++)
					s_order[count - cursor + d] = freed[d];
			}
			count -= cursor;
			acc++;
			if (last)
				running = false;
		}
		__syncthreads();
	}

	// ---- store ----------------------------------------------------------------------------------------------------
	__syncthreads();
#pragma unroll
	for (int cc = 0; cc < CPT; cc++) {
		const int i = cc * WBLOCK + tid;
		if (i < N)
			e.y[b * N + i] = yv[cc];
	}
	for (int i = tid; i < MT_WORDS; i += WBLOCK)
		e.key[b * MT_WORDS + i] = key[i];
	for (int i = tid; i < f_cnt - f_rd; i += WBLOCK) {  // queue normalised to start at 0
		e.fifo[(b * 3 + 0) * CANDIDATES + i] = fx1[f_rd + i];
		e.fifo[(b * 3 + 1) * CANDIDATES + i] = fx2[f_rd + i];
		e.fifo[(b * 3 + 2) * CANDIDATES + i] = fr2[f_rd + i];
	}
	if (tid < D) {
		e.qh[b * D + tid] = s_qh[tid];
		e.order[b * D + tid] = s_order[tid];
	}
	if (tid == 0) {
		e.t[b] = t;
		e.dt[b] = dt;
		e.q_count[b] = count;
		e.fifo_rd[b] = 0;
		e.fifo_cnt[b] = f_cnt - f_rd;
		e.status[b] = status;
		e.accepted[b] += acc;
		e.rejected[b] += rej;
		if (acc) atomicAdd(&e.totals->accepted, (unsigned long long)acc);
		if (rej) atomicAdd(&e.totals->rejected, (unsigned long long)rej);
		if (fresh) atomicAdd(&e.totals->fresh_draws, (unsigned long long)fresh);
		if (status == TRAJ_MIN_STEP) atomicAdd((unsigned long long *)&e.totals->n_min_step, 1ull);
		if (status == TRAJ_NOISE_OVERFLOW) atomicAdd((unsigned long long *)&e.totals->n_overflow, 1ull);
		atomicMax(&e.totals->max_depth, max_depth);
	}
}

__global__ void k_iota_rows(int *order, int64_t B, int D)
{
	const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx < B * D)
		order[idx] = (int)(idx % D);
}

}  // namespace wide
}  // namespace jsde
