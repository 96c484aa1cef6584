// One trajectory per lane: noise memory, Rößler SRA1 / SRIW1 proposal, error ratio, commit and the step-size
// controller, with the trajectory's state held in registers for the whole call.
//
// Arithmetic keeps the reference's expression shapes (operation order, true divisions, no fused multiply-adds
// outside the libm restatements) so that the result equals the strict-IEEE build of the reference bit for bit:
//   noise memory ........ reference jitcsde/jitced_template.c:170-250  (append_noise, Brownian_bridge, get_noise)
//   iterated integrals .. jitced_template.c:269-291                     (get_I)
//   SRI / SRA ........... jitced_template.c:408-464 / :466-494          (get_next_step)
//   error ratio ......... jitced_template.c:502-526                     (get_p)
//   commit .............. jitced_template.c:528-536                     (accept_step)
//   controller .......... reference jitcsde/_jitcsde.py:569-595         (_control_for_min_step, _adjust_step_size)
// This translation unit must be compiled with -fmad=false.
#pragma once

#include "ensemble.cuh"
#include "exact_div.cuh"
#include "warp_rng.cuh"

namespace jsde {

template <class Model, class Rng = WarpRng<Model::N>>
struct Lane {
	static constexpr int N = Model::N;
	static constexpr bool ADDITIVE = Model::ADDITIVE;
	static constexpr int NP = Model::NPAR;

	const EnsembleView &e;
	int64_t k;
	static constexpr int IT = noise_item_doubles(N);  // one noise item: {h, DW[N], DZ[N]} padded to 16-byte pairs

	Rng &rng;  // source of accepted polar pairs: pop(x1, x2, r2) and the shared-memory log table
	double *q0;  // slot 0 of this trajectory's noise memory; slot s is at q0 + s*B*IT

	double y[N], t, dt;
	double ny[N], er[N], nt;  // last proposal
	double par[NP > 0 ? NP : 1];
	int count, cursor, head, status;  // noise memory: depth, items used by the last proposal, ring slot of item 0
	unsigned fresh;
	int max_depth;

	__device__ __forceinline__ Lane(const EnsembleView &view, int64_t index, Rng &source)
		: e(view), k(index), rng(source), q0(view.q + index * IT), fresh(0), max_depth(0) {}

	// the fused kernel's lanes integrate one trajectory after the other
	__device__ __forceinline__ void rebind(int64_t index)
	{
		k = index;
		q0 = e.q + index * IT;
	}

	// ---- global-memory accessors ---------------------------------------------------------------------------
	// Noise memory: items of IT doubles in a ring of slots, [slot][B][IT]; a warp's 32 items of one slot are 32*IT*8
	// contiguous bytes and an item moves as IT/2 16-byte vectors.  Logical item d sits in slot (head + d) mod ring size.
	__device__ __forceinline__ double *item_ptr(int d) const { return q0 + (int64_t)((head + d) & e.Dmask) * e.B * IT; }

	__device__ __forceinline__ void load_item(int d, double &h, double (&W)[N], double (&Z)[N]) const
	{
		const double2 *p = reinterpret_cast<const double2 *>(item_ptr(d));
		double v[IT];
#pragma unroll
		for (int i = 0; i < IT / 2; i++) {
			const double2 t = p[i];
			v[2 * i] = t.x;
			v[2 * i + 1] = t.y;
		}
		h = v[0];
#pragma unroll
		for (int i = 0; i < N; i++) {
			W[i] = v[1 + i];
			Z[i] = v[1 + N + i];
		}
	}

	__device__ __forceinline__ void store_item(int d, double h, const double (&W)[N], const double (&Z)[N]) const
	{
		double2 *p = reinterpret_cast<double2 *>(item_ptr(d));
		double v[IT];
		v[0] = h;
#pragma unroll
		for (int i = 0; i < N; i++) {
			v[1 + i] = W[i];
			v[1 + N + i] = Z[i];
		}
		if (IT > 1 + 2 * N)
			v[IT - 1] = 0.0;
#pragma unroll
		for (int i = 0; i < IT / 2; i++)
			p[i] = make_double2(v[2 * i], v[2 * i + 1]);
	}

	__device__ __forceinline__ void move_item(int to, int from) const
	{
		const double2 *src = reinterpret_cast<const double2 *>(item_ptr(from));
		double2 *dst = reinterpret_cast<double2 *>(item_ptr(to));
#pragma unroll 1
		for (int i = 0; i < IT / 2; i++)
			dst[i] = src[i];
	}

	__device__ __forceinline__ void load()
	{
#pragma unroll
		for (int i = 0; i < N; i++)
			y[i] = e.y[(int64_t)i * e.B + k];
		t = e.t[k];
		dt = e.dt[k];
		count = e.q_count[k];
		cursor = e.q_cursor[k];
		head = e.q_head[k];
		status = e.status[k];
#pragma unroll
		for (int i = 0; i < NP; i++)
			par[i] = e.par[e.par_stride ? (int64_t)i * e.B + k : i];
	}

	__device__ __forceinline__ void store() const
	{
#pragma unroll
		for (int i = 0; i < N; i++)
			e.y[(int64_t)i * e.B + k] = y[i];
		e.t[k] = t;
		e.dt[k] = dt;
		e.q_count[k] = count;
		e.q_cursor[k] = cursor;
		e.q_head[k] = head;
		e.status[k] = status;
	}

	__device__ __forceinline__ void load_proposal()
	{
#pragma unroll
		for (int i = 0; i < N; i++) {
			ny[i] = e.new_y[(int64_t)i * e.B + k];
			er[i] = e.err[(int64_t)i * e.B + k];
		}
		nt = e.new_t[k];
	}

	__device__ __forceinline__ void store_proposal() const
	{
#pragma unroll
		for (int i = 0; i < N; i++) {
			e.new_y[(int64_t)i * e.B + k] = ny[i];
			e.err[(int64_t)i * e.B + k] = er[i];
		}
		e.new_t[k] = nt;
	}

	// ---- generator ---------------------------------------------------------------------------------------------
	// get_gauss (template:145-154): components in order, one accepted pair each, popped from this lane's FIFO (the
	// caller has run rng.ensure(…, N) with the whole warp).  Only the scale-dependent tail of rk_gauss is left:
	// f = scale*sqrt(-2*log(r2)/r2), DW = x1*f, DZ = x2*f (random_numbers.c:127-129).
	__device__ __forceinline__ void draw(double scale, double (&gW)[N], double (&gZ)[N])
	{
		if constexpr (Rng::TAIL) {  // the generator side has computed s = sqrt(-2*log(r2)/r2) already
#pragma unroll
			for (int i = 0; i < N; i++) {
				double s;
				rng.pop(gW[i], gZ[i], s);
				const double f = scale * s;
				gW[i] *= f;
				gZ[i] *= f;
			}
			return;
		}
		double r2[N], lg[N];
		unsigned pending = 0;  // components whose radius needs glibc's near-one log branch
#pragma unroll
		for (int i = 0; i < N; i++) {
			rng.pop(gW[i], gZ[i], r2[i]);
			lg[i] = jsde_log_table_from(r2[i], rng.log_tab);  // convergent; overwritten below for the ~6 % near-one radii
			if (jsde_log_is_near_one(r2[i]))
				pending |= 1u << i;
		}
		// the long near-one branch runs once per lane-with-such-a-component, not once per component for the whole warp
		while (pending) {
			const int which = __ffs(pending) - 1;
			pending &= pending - 1;
			double x = r2[0];
#pragma unroll
			for (int i = 1; i < N; i++)
				if (i == which)
					x = r2[i];
			const double v = jsde_log_near_one(x);
#pragma unroll
			for (int i = 0; i < N; i++)
				if (i == which)
					lg[i] = v;
		}
		// f = scale*sqrt(-2*log(r2)/r2) for all components in straight-line code (exact_div.cuh)
		double f[N];
		bool tame = true;
#pragma unroll
		for (int i = 0; i < N; i++)
			f[i] = scale * inline_sqrt(inline_div(__dmul_rn(-2.0, lg[i]), r2[i], tame), tame);
		if (!tame) {  // cold: operands near the ends of the exponent range
#pragma unroll
			for (int i = 0; i < N; i++)
				f[i] = scale * sqrt(__ddiv_rn(__dmul_rn(-2.0, lg[i]), r2[i]));
		}
#pragma unroll
		for (int i = 0; i < N; i++) {
			gW[i] *= f[i];
			gZ[i] *= f[i];
		}
	}

	// ---- noise memory -------------------------------------------------------------------------------------------
	// get_noise (template:204-250), in two halves so that the kernel can put the generator's asynchronous staging
	// between the noise memory's loads and everything that follows.  At most one fresh item is drawn per call: after
	// a bridge or an append the item at the cursor has exactly the missing length, so the reference's loop consumes
	// it and stops.
	struct NoiseScan {
		double need, hq;
		double qW[N], qZ[N];  // the too-long item at the cursor (valid when cur < count after the scan)
		int cur;
		bool started;
	};

	// first half: consume whole items from the head of the memory
	__device__ __forceinline__ void scan_noise(double h, NoiseScan &s, double (&W)[N], double (&Z)[N])
	{
		s.started = false;
		s.need = h;
		s.cur = 0;
		s.hq = 0.0;
		while (s.need != 0.0 && s.cur < count) {
			load_item(s.cur, s.hq, s.qW, s.qZ);
			if (!(s.hq <= s.need))
				break;
			if (s.started) {
#pragma unroll
				for (int i = 0; i < N; i++) {
					W[i] += s.qW[i];
					Z[i] += s.qZ[i];
				}
			} else {
				s.started = true;
#pragma unroll
				for (int i = 0; i < N; i++) {
					W[i] = s.qW[i];
					Z[i] = s.qZ[i];
				}
			}
			s.need -= s.hq;
			s.cur++;
		}
	}

	// second half: bridge or append with one fresh draw if something is still missing
	__device__ __forceinline__ bool finish_noise(NoiseScan &s, double (&W)[N], double (&Z)[N])
	{
		if (s.need != 0.0) {
			if (count >= e.D) {
				status = TRAJ_NOISE_OVERFLOW;
				return false;
			}
			const bool bridging = s.cur < count;  // then the scan left the too-long item in (hq, qW, qZ)
			double h_exc = 0.0, factor = 0.0, variance = s.need;  // append_noise (template:170-176): scale = sqrt(h_need)
			if (bridging) {  // Brownian_bridge (template:178-202): scale = sqrt(factor*h_need)
				h_exc = s.hq - s.need;
				factor = h_exc / s.hq;
				variance = factor * s.need;
			}
			const double scale = sqrt(variance);  // one square root for both kinds of lanes
			double gW[N], gZ[N];
			draw(scale, gW, gZ);
			fresh++;
			if (bridging) {
				// make room for one more item at the cursor (the reference relinks its list): shift whichever side is
				// shorter — the items behind the cursor up, or the already consumed ones in front of it down (the ring
				// has a free slot before its head).  The usual case, a bridge over the first or the last item, moves nothing.
				if (s.cur <= count - 1 - s.cur) {
					head = (head - 1) & e.Dmask;
#pragma unroll 1
					for (int d = 0; d < s.cur; d++)
						move_item(d, d + 1);
				} else {
#pragma unroll 1
					for (int d = count - 1; d > s.cur; d--)
						move_item(d + 1, d);
				}
				double w2[N], z2[N];
#pragma unroll
				for (int i = 0; i < N; i++) {
					w2[i] = gW[i] + s.qW[i] * factor;
					z2[i] = gZ[i] + s.qZ[i] * factor;
					gW[i] = s.qW[i] - w2[i];
					gZ[i] = s.qZ[i] - z2[i];
				}
				store_item(s.cur + 1, h_exc, w2, z2);
			}
			store_item(s.cur, s.need, gW, gZ);
			count++;
			if (s.started) {
#pragma unroll
				for (int i = 0; i < N; i++) {
					W[i] += gW[i];
					Z[i] += gZ[i];
				}
			} else {
				s.started = true;
#pragma unroll
				for (int i = 0; i < N; i++) {
					W[i] = gW[i];
					Z[i] = gZ[i];
				}
			}
			s.cur++;
		}
		if (!s.started) {
#pragma unroll
			for (int i = 0; i < N; i++)
				W[i] = Z[i] = 0.0;
		}
		cursor = s.cur;
		max_depth = count > max_depth ? count : max_depth;
		return true;
	}

	// pin_noise (template:252-267): append one fresh item of length `step` (the kernel loops `number` times so that the
	// whole warp can refill its generator FIFOs between items)
	__device__ __forceinline__ void pin_one(double step)
	{
		if (count >= e.D) {
			status = TRAJ_NOISE_OVERFLOW;
			return;
		}
		double gW[N], gZ[N];
		draw(sqrt(step), gW, gZ);
		store_item(count, step, gW, gZ);
		count++;
		cursor = count - 1;  // template:173: current_noise = the newest item, so an accept_step right after keeps only it
		max_depth = count > max_depth ? count : max_depth;
	}

	// the same with caller-supplied increments instead of a fresh draw (jsde_pin_noise_path)
	__device__ __forceinline__ void pin_given(double step, const double (&W)[N], const double (&Z)[N])
	{
		if (count >= e.D) {
			status = TRAJ_NOISE_OVERFLOW;
			return;
		}
		store_item(count, step, W, Z);
		count++;
		cursor = count - 1;
		max_depth = count > max_depth ? count : max_depth;
	}

	// ---- proposal (get_next_step) ---------------------------------------------------------------------------------
	__device__ __forceinline__ bool propose(double h)
	{
		double I1[N], DZ[N];
		NoiseScan s;
		scan_noise(h, s, I1, DZ);
		if (!finish_noise(s, I1, DZ))
			return false;
		stages(h, I1, DZ);
		return true;
	}

	// the Runge–Kutta stages proper, from the Wiener increments I1 = DW and DZ of this step.  The quotients by h,
	// sqrt(h) and 6 go through exact_div.cuh (same bits as `/`, straight-line code); if an operand is outside the range
	// for which that is proven, the attempt is redone with ordinary divisions.
	__device__ __forceinline__ void stages(double h, const double (&I1)[N], const double (&DZ)[N])
	{
		if (!stages_impl<true>(h, I1, DZ))
			stages_impl<false>(h, I1, DZ);  // cold: inlined here, never executed for sane step sizes and states
	}

	template <bool FAST>
	__device__ __forceinline__ bool stages_impl(double h, const double (&I1)[N], const double (&DZ)[N])
	{
		bool tame = true;
		const SharedDivisor by_h(h);
		const SharedDivisor by_6(6.0, 1.0 / 6.0);
		auto div_h = [&](double a) { return FAST ? by_h.divide(a, tame) : a / h; };
		auto div_6 = [&](double a) { return FAST ? by_6.divide(a, tame) : a / 6; };
		double I10[N];
		double arg[N], fh1[N], fh2[N], g1[N], g2[N];
		if constexpr (!ADDITIVE) {
			double g3[N], g4[N];
			const double sqh = sqrt(h);
			const SharedDivisor by_sqh(sqh);
			auto div_sqh = [&](double a) { return FAST ? by_sqh.divide(a, tame) : a / sqh; };
#pragma unroll
			for (int i = 0; i < N; i++)  // template:289; I_11/sqrt(h) and I_111 (:286-287) are formed where they are used
				I10[i] = (I1[i] + DZ[i] * (1. / 1.7320508075688772)) * (h / 2.);
			Model::drift(t, y, h, par, fh1);
			Model::diffusion(t, y, par, g1);
#pragma unroll
			for (int i = 0; i < N; i++)
				arg[i] = y[i] + 0.75 * fh1[i] + div_h(1.5 * g1[i] * I10[i]);
			Model::drift(t + 0.75 * h, arg, h, par, fh2);
#pragma unroll
			for (int i = 0; i < N; i++)
				arg[i] = y[i] + 0.25 * fh1[i] + 0.5 * g1[i] * sqh;
			Model::diffusion(t + 0.25 * h, arg, par, g2);
#pragma unroll
			for (int i = 0; i < N; i++)
				arg[i] = y[i] + fh1[i] - g1[i] * sqh;
			Model::diffusion(t + h, arg, par, g3);
#pragma unroll
			for (int i = 0; i < N; i++)
				arg[i] = y[i] + 0.25 * fh1[i] + sqh * (-5 * g1[i] + 3 * g2[i] + 0.5 * g3[i]);
			Model::diffusion(t + 0.25 * h, arg, par, g4);
			const double third_over_h = 1. / h / 3.;
#pragma unroll
			for (int i = 0; i < N; i++) {  // template:447-464
				const double I11s = div_sqh((I1[i] * I1[i] - h) * 0.5);
				const double I111 = (I1[i] * I1[i] * I1[i] - 3 * h * I1[i]) * (1. / 6.);
				const double E_N = third_over_h * (
					+ (6 * I10[i] - 6 * I111) * g1[i]
					+ (-4 * I10[i] + 5 * I111) * g2[i]
					+ (-2 * I10[i] - 2 * I111) * g3[i]
					+ (3 * I111) * g4[i]);
				ny[i] = y[i] + E_N + (1. / 3.) * (
					fh1[i] + 2 * fh2[i]
					+ (-3 * I1[i] - 3 * I11s) * g1[i]
					+ (4 * I1[i] + 4 * I11s) * g2[i]
					+ (2 * I1[i] - I11s) * g3[i]);
				const double E_D = div_6(fh2[i] - fh1[i]);
				er[i] = fabs(E_D) + fabs(E_N);
			}
		} else {
#pragma unroll
			for (int i = 0; i < N; i++)
				I10[i] = (I1[i] + DZ[i] * (1. / 1.7320508075688772)) * (h / 2.);
			Model::drift(t, y, h, par, fh1);
			Model::diffusion(t + h, y, par, g1);
#pragma unroll
			for (int i = 0; i < N; i++)
				arg[i] = y[i] + 0.75 * fh1[i] + div_h(0.5 * g1[i] * I10[i]);
			Model::drift(t + 0.75 * h, arg, h, par, fh2);
			Model::diffusion(t, y, par, g2);
#pragma unroll
			for (int i = 0; i < N; i++) {  // template:487-494
				const double E_N = div_h(I10[i]) * (g2[i] - g1[i]);
				ny[i] = y[i] + E_N + (1. / 3.) * (fh1[i] + 2 * fh2[i]) + I1[i] * g1[i];
				const double E_D = div_6(fh2[i] - fh1[i]);
				er[i] = fabs(E_D) + fabs(E_N);
			}
		}
		nt = t + h;
		return tame;
	}

	// get_p (template:502-526): a 0/0 component is skipped, a NaN never wins the comparison
	__device__ __forceinline__ double error_ratio(double atol, double rtol) const
	{
		// the N quotients in straight-line code (exact_div.cuh); a zero error needs no division: 0/tol is +0 or — for a
		// NaN tolerance — a NaN, and neither can win the comparison below
		double tol[N], x[N];
		bool tame = true;
#pragma unroll
		for (int i = 0; i < N; i++) {
			tol[i] = atol + rtol * fabs(ny[i]);
			bool mine = true;
			const double q = inline_div(er[i], tol[i], mine);
			x[i] = er[i] == 0.0 ? 0.0 : q;
			tame = tame && (mine || er[i] == 0.0);
		}
		if (!tame) {  // cold
#pragma unroll
			for (int i = 0; i < N; i++)
				x[i] = er[i] / tol[i];
		}
		double worst = 0.0;
#pragma unroll
		for (int i = 0; i < N; i++) {
			if (er[i] != 0.0 || tol[i] != 0.0) {
				if (x[i] > worst)
					worst = x[i];
			}
		}
		return worst;
	}

	// accept_step (template:528-536): commit the proposal, drop the consumed items
	__device__ __forceinline__ void commit()
	{
#pragma unroll
		for (int i = 0; i < N; i++)
			y[i] = ny[i];
		t = nt;
		count -= cursor;
		// an emptied memory restarts at slot 0: the usual append-accept cycle then rewrites the same 64 bytes, which stay
		// in L2 (a head that keeps advancing walks through the whole ring — 2 GB per 2^20 trajectories — and turns
		// every item into DRAM write traffic: measured +17 GB per 3*10^8 attempts)
		head = count == 0 ? 0 : (head + cursor) & e.Dmask;
		cursor = 0;
	}

	// _adjust_step_size (_jitcsde.py:581-595).  Returns 1 accepted, 0 rejected, -1 rejected with dt < min_step.
	// Python's max(a,b)/min(a,b) return b only if it compares strictly greater/smaller; kept literally.
	__device__ __forceinline__ int adjust_step(const Controller &c, double actual_dt)
	{
		const double p = error_ratio(c.atol, c.rtol);
		const bool rejecting = p > c.dec_thr;
		const bool growing = !rejecting && p <= c.inc_thr;
		// one call site for both uses of Python's `**` (:587, :592) so that rejecting and growing lanes share it
		double pw = 0.0;  // inf ** negative = 0.0
		if ((rejecting && p <= 1.7976931348623157e308) || (growing && p != 0.0))
			pw = jsde_pow_with(p, rejecting ? c.shrink_exp : c.grow_exp, rng.pow_tables());
		if (rejecting) {
			const double shrink = c.safety * pw;
			dt = actual_dt * (c.min_factor > shrink ? c.min_factor : shrink);
			return dt < c.min_step ? -1 : 0;
		}
		if (growing) {
			const double factor = (p != 0.0) ? c.safety * pw : c.max_factor;
			const double sug = actual_dt * (c.max_factor < factor ? c.max_factor : factor);
			const double capped = c.max_step < sug ? c.max_step : sug;
			dt = capped > dt ? capped : dt;
		}
		return 1;
	}
};

}  // namespace jsde
