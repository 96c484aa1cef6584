// extern "C" boundary (include/jsde.h) for WIDE models (one thread block per group of trajectories, wide.cuh).  Same ABI
// as jsde_abi.cu (shared plumbing: abi_common.cuh), including the single-step surface and pin_noise; what the wide path
// does not implement (the generator debug stream) returns JSDE_E_STATE with a message.
// Must be compiled with -fmad=false.
#include "glibc_math.h"
#include "wide.cuh"

#ifndef JSDE_MODEL_HEADER
#error "compile with -DJSDE_MODEL_HEADER=\"model.cuh\""
#endif
#include JSDE_MODEL_HEADER

#include "abi_common.cuh"

using jsde_model::Model;
using namespace jsde;
using namespace jsde::abi;
using namespace jsde::wide;

static_assert(Model::WIDE, "jsde_abi_wide.cu is for wide models");

struct jsde_handle : HandleBase {
	int tpb = 1;  // trajectories per thread block (wide.cuh)
	WideView v{};
	double *staging = nullptr;  // [B][n]: host-supplied per-trajectory vectors (step sizes, jump amplitudes, masks)
	double *par_rows = nullptr;  // [B][NPAR] per-trajectory control parameters (allocated on first use)
};

namespace {

int unsupported(const char *what)
{
	return fail(JSDE_E_STATE, std::string(what) + " is not implemented for wide (block-per-trajectory) models");
}

// drop the noise memory, restart the generator from the stored seeds (lazily), restart jump scheduling
int reset_integrator(jsde_handle *h)
{
	const int64_t B = h->B;
	CU(cudaMemsetAsync(h->v.q_count, 0, B * sizeof(int), h->stream));
	CU(cudaMemsetAsync(h->v.status, 0, B * sizeof(int), h->stream));
	CU(cudaMemsetAsync(h->v.accepted, 0, B * sizeof(unsigned long long), h->stream));
	CU(cudaMemsetAsync(h->v.rejected, 0, B * sizeof(unsigned long long), h->stream));
	CU(cudaMemsetAsync(h->v.jump_set, 0, B, h->stream));
	CU(cudaMemsetAsync(h->v.tape_pos, 0, B * sizeof(long long), h->stream));
	CU(cudaMemsetAsync(h->v.cursor, 0, B * sizeof(int), h->stream));
	CU(cudaMemsetAsync(h->v.h_last, 0, B * sizeof(double), h->stream));
	CU(cudaMemsetAsync(h->v.sNY, 0, (size_t)B * Model::N * sizeof(double), h->stream));
	CU(cudaMemsetAsync(h->v.sErr, 0, (size_t)B * Model::N * sizeof(double), h->stream));
	k_iota_rows<<<grid_for(B * h->D, 256), 256, 0, h->stream>>>(h->v.order, B, h->D);
	CU(cudaGetLastError());
	h->seed_dirty = true;
	return JSDE_OK;
}

int ensure_seeded(jsde_handle *h)
{
	if (h->seed_dirty) {
		k_wide_seed<<<grid_for(h->B, 64), 64, 0, h->stream>>>(h->v.key, h->v.pos, h->seeds, h->B);
		CU(cudaGetLastError());
		h->seed_dirty = false;
	}
	return JSDE_OK;
}

// One block integrates `tpb` trajectories (wide.cuh).  tpb is chosen so that the ensemble fills the GPU's block slots
// (2 blocks per SM) in one wave where possible: 2^11 trajectories on 148 SMs -> 7 per block, 293 blocks on 296 slots.
template <int T>
int launch_one(jsde_handle *h, double target_time, const double *grid_times, int grid_count, double *grid_out, const WideMode &wm)
{
	const size_t smem = wide_smem_bytes<Model, T>();
	CU(cudaFuncSetAttribute(k_wide_integrate<Model, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	k_wide_integrate<Model, T><<<grid_for(h->B, T), WBLOCK, smem, h->stream>>>(h->v, h->ctrl, target_time, grid_times, grid_count, grid_out, wm);
	CU(cudaGetLastError());
	return JSDE_OK;
}

int launch_integrate(jsde_handle *h, double target_time, const double *grid_times = nullptr, int grid_count = 0, double *grid_out = nullptr,
	const WideMode &wm = WideMode{WIDE_INTEGRATE, nullptr, 0u, 0.0})
{
	switch (h->tpb) {
	case 1: return launch_one<1>(h, target_time, grid_times, grid_count, grid_out, wm);
	case 2: return launch_one<2>(h, target_time, grid_times, grid_count, grid_out, wm);
	case 4: return launch_one<4>(h, target_time, grid_times, grid_count, grid_out, wm);
	case 7: return launch_one<7>(h, target_time, grid_times, grid_count, grid_out, wm);
	case 8: return launch_one<8>(h, target_time, grid_times, grid_count, grid_out, wm);
	default:
		if constexpr (TMAX >= 16) {
			if (h->tpb == 14) return launch_one<14>(h, target_time, grid_times, grid_count, grid_out, wm);
			return launch_one<16>(h, target_time, grid_times, grid_count, grid_out, wm);
		}
		return launch_one<8>(h, target_time, grid_times, grid_count, grid_out, wm);
	}
}

int choose_tpb(int device, int64_t B)
{
	static const int allowed[] = {1, 2, 4, 7, 8, 14, 16};
	if (const char *env = getenv("JSDE_WIDE_TPB")) {  // development / test knob
		const int want = atoi(env);
		for (int a : allowed)
			if (a == want && a <= TMAX) return a;
	}
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
		return 8;
	const int64_t slots = WBLOCKS_PER_SM * (int64_t)prop.multiProcessorCount;
	const int64_t need = (B + slots - 1) / slots;  // trajectories per block for a single wave
	for (int a : allowed)
		if (a >= need && a <= TMAX) return a;
	return TMAX;
}

}  // namespace

JSDE_DEFINE_COMMON_ENTRY_POINTS(Model)

extern "C" {

const char *jsde_version(void) { return "jsde-b200 0.1 wide (sm_100a, fp64, -fmad=false)"; }
int jsde_destroy(jsde_t *h)
{
	if (!h)
		return JSDE_OK;
	destroy_base(h);
	delete h;
	return JSDE_OK;
}

int jsde_create(jsde_t **out, int device, int64_t B, int noise_capacity)
{
	if (!out || B <= 0 || noise_capacity < 0 || noise_capacity > MAX_DEPTH)
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc = check_device(device);
	if (rc) return rc;
	cudaError_t err = cudaSuccess;
	jsde_handle *h = new jsde_handle();
	h->device = device;
	h->B = B;
	h->D = noise_capacity ? noise_capacity : 32;
	if (h->D < 2) h->D = 2;
	h->tpb = choose_tpb(device, B);
	constexpr int n = Model::N;
	auto bail = [&](int code) { jsde_destroy(h); return code; };
	if ((rc = bind(h)) || (rc = create_stream_and_events(h))) return bail(rc);
	WideView &v = h->v;
	v.B = B;
	v.D = h->D;
#define ALLOC(ptr, cnt) if ((rc = dev_alloc(h, &(ptr), (size_t)(cnt)))) return bail(rc)
	ALLOC(v.y, B * n); ALLOC(v.t, B); ALLOC(v.dt, B);
	ALLOC(v.qh, B * h->D); ALLOC(v.qw, B * h->D * 2 * n); ALLOC(v.order, B * h->D); ALLOC(v.q_count, B);
	ALLOC(v.key, B * MT_WORDS); ALLOC(v.pos, B);
	ALLOC(v.arg, B * n); ALLOC(v.sW, B * n); ALLOC(v.sZ, B * n); ALLOC(v.sF1, B * n); ALLOC(v.sNY, B * n);
	ALLOC(v.next_jump, B); ALLOC(v.jump_set, B); ALLOC(v.tape_pos, B);
	ALLOC(v.sErr, B * n); ALLOC(v.h_last, B); ALLOC(v.cursor, B); ALLOC(h->staging, B * n);
	if (!Model::ADDITIVE) {
		ALLOC(v.argA, B * n); ALLOC(v.argB, B * n); ALLOC(v.argC, B * n); ALLOC(v.sG2, B * n); ALLOC(v.sG3, B * n);
	}
	ALLOC(v.status, B); ALLOC(v.accepted, B); ALLOC(v.rejected, B); ALLOC(v.totals, 1);
	ALLOC(h->seeds, B); ALLOC(h->par, Model::NPAR > 0 ? Model::NPAR : 1); ALLOC(h->shift, n);
#undef ALLOC
	v.par = h->par;
	v.par_stride = 0;
	v.dmma = 0;
	if (const char *env = getenv("JSDE_WIDE_COUPLING"))  // development knob; the API is jsde_set_option
		v.dmma = Model::UNITS > 0 && !strcmp(env, "dmma");
	default_controller(h);
	if ((err = cudaMemsetAsync(h->seeds, 0, B * sizeof(uint32_t), h->stream)) != cudaSuccess ||
		(err = cudaMemsetAsync(v.y, 0, (size_t)B * n * sizeof(double), h->stream)) != cudaSuccess ||
		(err = cudaMemsetAsync(v.t, 0, B * sizeof(double), h->stream)) != cudaSuccess ||
		(err = cudaMemsetAsync(v.qh, 0, (size_t)B * h->D * sizeof(double), h->stream)) != cudaSuccess ||
		(err = cudaMemsetAsync(h->par, 0, (Model::NPAR > 0 ? Model::NPAR : 1) * sizeof(double), h->stream)) != cudaSuccess)
		return bail(fail(JSDE_E_CUDA, cudaGetErrorString(err)));
	k_fill<<<grid_for(B, 256), 256, 0, h->stream>>>(v.dt, h->first_step, B);
	if ((rc = reset_integrator(h))) return bail(rc);
	if ((err = cudaStreamSynchronize(h->stream)) != cudaSuccess)
		return bail(fail(JSDE_E_CUDA, cudaGetErrorString(err)));
	*out = h;
	return JSDE_OK;
}

int jsde_set_option(jsde_t *h, const char *name, const char *value)
{
	if (!h || !name || !value)
		return fail(JSDE_E_ARG, "Wrong input.");
	if (!strcmp(name, "coupling")) {
		if (!strcmp(value, "exact")) { h->v.dmma = 0; return JSDE_OK; }
		if (!strcmp(value, "dmma")) {
			if (Model::UNITS <= 0)
				return fail(JSDE_E_STATE, "this model has no dense coupling term");
			h->v.dmma = 1;
			return JSDE_OK;
		}
	}
	return fail(JSDE_E_ARG, std::string("Wrong input. (unknown option ") + name + "=" + value + ")");
}

int jsde_set_state(jsde_t *h, const double *y0_host, const double *t0_host, double t0)
{
	if (!h || !y0_host)
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc = bind(h);
	if (rc) return rc;
	CU(cudaMemcpyAsync(h->v.y, y0_host, (size_t)h->B * Model::N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
	if (t0_host)
		CU(cudaMemcpyAsync(h->v.t, t0_host, h->B * sizeof(double), cudaMemcpyHostToDevice, h->stream));
	else
		k_fill<<<grid_for(h->B, 256), 256, 0, h->stream>>>(h->v.t, t0, h->B);
	if ((rc = reset_integrator(h))) return rc;
	CU(cudaStreamSynchronize(h->stream));
	return JSDE_OK;
}

int jsde_seed(jsde_t *h, const uint32_t *seeds_host)
{
	if (!h || !seeds_host)
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc = bind(h);
	if (rc) return rc;
	CU(cudaMemcpyAsync(h->seeds, seeds_host, h->B * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
	if ((rc = reset_integrator(h))) return rc;
	CU(cudaStreamSynchronize(h->stream));
	return JSDE_OK;
}

int jsde_set_integration_parameters(jsde_t *h, const jsde_ctrl *p)
{
	if (!h || !p)
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc = apply_controller(h, p);
	if (rc) return rc;
	if ((rc = bind(h))) return rc;
	k_fill<<<grid_for(h->B, 256), 256, 0, h->stream>>>(h->v.dt, h->first_step, h->B);
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(h->stream));
	return JSDE_OK;
}

int jsde_set_parameters(jsde_t *h, const double *params_host, int per_trajectory)
{
	if (!h || (Model::NPAR > 0 && !params_host))
		return fail(JSDE_E_ARG, "Wrong input.");
	if (Model::NPAR == 0)
		return JSDE_OK;
	int rc = bind(h);
	if (rc) return rc;
	if (!per_trajectory) {
		CU(cudaMemcpyAsync(h->par, params_host, Model::NPAR * sizeof(double), cudaMemcpyHostToDevice, h->stream));
		h->v.par = h->par;
		h->v.par_stride = 0;
	} else {  // one row per trajectory, used in place: a block reads the rows of its own trajectories
		if (!h->par_rows && (rc = dev_alloc(h, &h->par_rows, (size_t)h->B * Model::NPAR)))  // allocated once, reused by later calls
			return rc;
		CU(cudaMemcpyAsync(h->par_rows, params_host, (size_t)h->B * Model::NPAR * sizeof(double), cudaMemcpyHostToDevice, h->stream));
		h->v.par = h->par_rows;
		h->v.par_stride = Model::NPAR;
	}
	CU(cudaStreamSynchronize(h->stream));
	return JSDE_OK;
}

namespace {

// jump_mode: 0 none, 1 tape (set before), 2 generated variates
int select_jump_source(jsde_handle *h, int jump_mode, uint64_t seed, int64_t first_index)
{
	h->v.use_jumps = 0;
	if (jump_mode == 0)
		return JSDE_OK;
	if (!Model::HAS_JUMP)
		return fail(JSDE_E_STATE, "this model was generated without a jump amplitude");
	if (jump_mode == 1) {
		if (!h->tape_set)
			return fail(JSDE_E_STATE, "no jump tape has been set (jsde_set_jump_tape)");
		h->v.taus = h->taus;
		h->v.xis = h->xis;
		h->v.tape_len = h->tape_len;
		h->v.generated_jumps = 0;
	} else if (jump_mode == 2) {
		if (first_index < 0)
			return fail(JSDE_E_ARG, "Wrong input.");
		h->v.generated_jumps = 1;
		h->v.jump_seed = seed;
		h->v.jump_first = first_index;
	} else
		return fail(JSDE_E_ARG, "Wrong input. (jump_mode)");
	h->v.use_jumps = 1;
	return JSDE_OK;
}

int integrate_common(jsde_handle *h, double target_time, int jump_mode, uint64_t seed, int64_t first_index, jsde_stats *stats_host)
{
	if (!h || !(target_time == target_time))
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc = bind(h);
	if (rc) return rc;
	if ((rc = select_jump_source(h, jump_mode, seed, first_index))) return rc;
	if ((rc = ensure_seeded(h))) return rc;
	CU(cudaMemsetAsync(h->v.totals, 0, sizeof(CallTotals), h->stream));
	CU(cudaEventRecord(h->ev0, h->stream));
	if ((rc = launch_integrate(h, target_time))) return rc;
	return finish_call(h, h->v.totals, stats_host, 1);
}

}  // namespace

int jsde_integrate(jsde_t *h, double target_time, jsde_stats *stats_host) { return integrate_common(h, target_time, 0, 0, 0, stats_host); }

int jsde_set_jump_tape(jsde_t *h, const jsde_jump_tape *tape)
{
	if (!h)
		return fail(JSDE_E_ARG, "Wrong input.");
	if (!Model::HAS_JUMP)
		return fail(JSDE_E_STATE, "this model was generated without a jump amplitude");
	int rc = bind(h);
	if (rc) return rc;
	return upload_tape(h, tape);
}

// jitcsde_jump.integrate (_jitcsde.py:693-700) with the variates from a tape …
int jsde_integrate_jumps(jsde_t *h, double target_time, const jsde_jump_tape *tape, jsde_stats *stats_host)
{
	if (!h)
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc;
	if (tape && (rc = jsde_set_jump_tape(h, tape))) return rc;
	return integrate_common(h, target_time, 1, 0, 0, stats_host);
}

// … or from the counter-based generator (csrc/philox_jumps.h)
int jsde_integrate_jumps_generated(jsde_t *h, double target_time, uint64_t seed, int64_t first_index, jsde_stats *stats_host)
{
	return integrate_common(h, target_time, 2, seed, first_index, stats_host);
}

// the caller's sampling loop `for time in times: integrate(time)` (examples/noisy_lorenz.py:95-96) in one launch;
// y_host [n_times][B][n]
int jsde_integrate_grid_jumps(jsde_t *h, const double *times_host, int64_t n_times, int jump_mode, uint64_t seed, int64_t first_index,
	double *y_host, jsde_stats *stats_host)
{
	if (!h || !times_host || n_times <= 0 || n_times > 0x7fffffff || !y_host)
		return fail(JSDE_E_ARG, "Wrong input.");
	for (int64_t j = 0; j < n_times; j++)
		if (!(times_host[j] == times_host[j]))
			return fail(JSDE_E_ARG, "Wrong input. (sample time is NaN)");
	int rc = bind(h);
	if (rc) return rc;
	if ((rc = select_jump_source(h, jump_mode, seed, first_index))) return rc;
	if ((rc = ensure_seeded(h))) return rc;
	const size_t count = (size_t)n_times * (size_t)h->B * Model::N;
	Scratch scratch;
	double *times = nullptr, *out = nullptr;
	CU(scratch.get(&times, n_times));
	if (scratch.get(&out, count) != cudaSuccess)
		return fail(JSDE_E_NOMEM, "sampling grid does not fit in device memory: n_times*B*n doubles");
	CU(cudaMemcpyAsync(times, times_host, n_times * sizeof(double), cudaMemcpyHostToDevice, h->stream));
	CU(cudaMemsetAsync(h->v.totals, 0, sizeof(CallTotals), h->stream));
	CU(cudaEventRecord(h->ev0, h->stream));
	if ((rc = launch_integrate(h, times_host[n_times - 1], times, (int)n_times, out)) || (rc = finish_call(h, h->v.totals, stats_host, 1))) return rc;
	CU(cudaMemcpy(y_host, out, count * sizeof(double), cudaMemcpyDeviceToHost));
	return JSDE_OK;
}

int jsde_integrate_grid(jsde_t *h, const double *times_host, int64_t n_times, const jsde_jump_tape *tape, double *y_host, jsde_stats *stats_host)
{
	if (!h)
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc;
	if (tape && (rc = jsde_set_jump_tape(h, tape))) return rc;
	return jsde_integrate_grid_jumps(h, times_host, n_times, tape ? 1 : 0, 0, 0, y_host, stats_host);
}

int jsde_get_state(jsde_t *h, double *y_host, double *t_host, int32_t *status_host)
{
	if (!h)
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc = bind(h);
	if (rc) return rc;
	if (y_host) CU(cudaMemcpyAsync(y_host, h->v.y, (size_t)h->B * Model::N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
	if (t_host) CU(cudaMemcpyAsync(t_host, h->v.t, h->B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
	if (status_host) CU(cudaMemcpyAsync(status_host, h->v.status, h->B * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	return JSDE_OK;
}

int jsde_get_controller_state(jsde_t *h, double *dt_host, uint64_t *accepted_host, uint64_t *rejected_host)
{
	if (!h)
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc = bind(h);
	if (rc) return rc;
	if (dt_host) CU(cudaMemcpyAsync(dt_host, h->v.dt, h->B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
	if (accepted_host) CU(cudaMemcpyAsync(accepted_host, h->v.accepted, h->B * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
	if (rejected_host) CU(cudaMemcpyAsync(rejected_host, h->v.rejected, h->B * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	return JSDE_OK;
}

int jsde_moments(jsde_t *h, double *acc_dev, const double *shift_host)
{
	if (!h || !acc_dev)
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc = bind(h);
	if (rc) return rc;
	constexpr int n = Model::N;
	if (shift_host)
		CU(cudaMemcpyAsync(h->shift, shift_host, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
	else
		CU(cudaMemsetAsync(h->shift, 0, n * sizeof(double), h->stream));
	CU(cudaMemsetAsync(acc_dev, 0, (1 + 2 * n) * sizeof(double), h->stream));
	const int threads = 128;
	unsigned chunks = (unsigned)((h->B + 31) / 32);  // >= 32 trajectories per block: one atomic per component and 32 rows
	if (chunks > 64) chunks = 64;
	k_moments_wide<<<dim3(grid_for(n, threads), chunks), threads, 0, h->stream>>>(h->v.y, h->v.status, h->B, n, h->shift, acc_dev);
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(h->stream));
	return JSDE_OK;
}

int jsde_set_time(jsde_t *h, const double *t_host)
{
	if (!h || !t_host)
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc = bind(h);
	if (rc) return rc;
	CU(cudaMemcpyAsync(h->v.t, t_host, h->B * sizeof(double), cudaMemcpyHostToDevice, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	return JSDE_OK;
}

// ---- single-step surface (reference core methods, jitced_template.c:252-267, 396-545), one launch each ----------------------
int jsde_pin_noise(jsde_t *h, unsigned number, double step)
{
	if (!h || !(step > 0.0))
		return fail(JSDE_E_ARG, "Wrong input.");
	if (number == 0)
		return JSDE_OK;
	int rc = bind(h);
	if (rc) return rc;
	if ((rc = ensure_seeded(h))) return rc;
	h->v.use_jumps = 0;
	CU(cudaMemsetAsync(h->v.totals, 0, sizeof(CallTotals), h->stream));
	if ((rc = launch_integrate(h, 0.0, nullptr, 0, nullptr, WideMode{WIDE_PIN, nullptr, number, step}))) return rc;
	CU(cudaStreamSynchronize(h->stream));
	return JSDE_OK;
}

int jsde_pin_noise_path(jsde_t *h, unsigned number, const double *steps_host, const double *dw_host, const double *dz_host)
{
	if (!h || (number && (!steps_host || !dw_host || !dz_host)))
		return fail(JSDE_E_ARG, "Wrong input.");
	for (unsigned j = 0; j < number; j++)
		if (!(steps_host[j] > 0.0))
			return fail(JSDE_E_ARG, "Wrong input. (step sizes must be positive)");
	if (number == 0)
		return JSDE_OK;
	int rc = bind(h);
	if (rc) return rc;
	const size_t values = (size_t)h->B * number * Model::N;
	Scratch scratch;
	double *steps = nullptr, *dw = nullptr, *dz = nullptr;
	if (scratch.get(&steps, number) != cudaSuccess || scratch.get(&dw, values) != cudaSuccess || scratch.get(&dz, values) != cudaSuccess)
		return fail(JSDE_E_NOMEM, "pin_noise_path: the increments do not fit in device memory");
	CU(cudaMemcpyAsync(steps, steps_host, number * sizeof(double), cudaMemcpyHostToDevice, h->stream));
	CU(cudaMemcpyAsync(dw, dw_host, values * sizeof(double), cudaMemcpyHostToDevice, h->stream));
	CU(cudaMemcpyAsync(dz, dz_host, values * sizeof(double), cudaMemcpyHostToDevice, h->stream));
	k_wide_pin_path<Model::N><<<(unsigned)h->B, 256, 0, h->stream>>>(h->v, number, steps, dw, dz);
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(h->stream));
	return JSDE_OK;
}

int jsde_get_next_step(jsde_t *h, const double *h_host)
{
	if (!h || !h_host)
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc = bind(h);
	if (rc) return rc;
	if ((rc = ensure_seeded(h))) return rc;
	h->v.use_jumps = 0;
	CU(cudaMemcpyAsync(h->staging, h_host, h->B * sizeof(double), cudaMemcpyHostToDevice, h->stream));
	CU(cudaMemsetAsync(h->v.totals, 0, sizeof(CallTotals), h->stream));
	if ((rc = launch_integrate(h, 0.0, nullptr, 0, nullptr, WideMode{WIDE_PROPOSE, h->staging, 0u, 0.0}))) return rc;
	CU(cudaStreamSynchronize(h->stream));
	return JSDE_OK;
}

int jsde_get_p(jsde_t *h, double atol, double rtol, double *p_host)
{
	if (!h || !p_host)
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc = bind(h);
	if (rc) return rc;
	k_wide_get_p<Model::N><<<(unsigned)h->B, 256, 0, h->stream>>>(h->v, atol, rtol, h->staging);
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(p_host, h->staging, h->B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	return JSDE_OK;
}

int jsde_accept_step(jsde_t *h, const uint8_t *mask_host)
{
	if (!h)
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc = bind(h);
	if (rc) return rc;
	unsigned char *mask = nullptr;
	if (mask_host) {
		mask = reinterpret_cast<unsigned char *>(h->staging);
		CU(cudaMemcpyAsync(mask, mask_host, h->B, cudaMemcpyHostToDevice, h->stream));
	}
	k_wide_accept<Model::N><<<(unsigned)h->B, 256, 0, h->stream>>>(h->v, mask);
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(h->stream));
	return JSDE_OK;
}

int jsde_apply_jump(jsde_t *h, const double *change_host)
{
	if (!h || !change_host)
		return fail(JSDE_E_ARG, "Wrong input. Note that the jump amplitudes must be an array of B*n doubles.");
	int rc = bind(h);
	if (rc) return rc;
	const int64_t count = h->B * Model::N;
	CU(cudaMemcpyAsync(h->staging, change_host, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
	k_wide_add<<<grid_for(count, 256), 256, 0, h->stream>>>(h->v.y, h->staging, count);
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(h->stream));
	return JSDE_OK;
}

int jsde_dump_noise(jsde_t *h, int64_t b, double *out_host, int max_items, int *count)
{
	if (!h || b < 0 || b >= h->B || !out_host || !count || max_items < 0)
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc = bind(h);
	if (rc) return rc;
	constexpr int n = Model::N;
	CU(cudaStreamSynchronize(h->stream));
	CU(cudaMemcpy(count, h->v.q_count + b, sizeof(int), cudaMemcpyDeviceToHost));
	std::vector<int> order(h->D);
	CU(cudaMemcpy(order.data(), h->v.order + b * h->D, h->D * sizeof(int), cudaMemcpyDeviceToHost));
	const int width = 1 + 2 * n;
	for (int d = 0; d < *count && d < max_items; d++) {
		const int slot = order[d];
		CU(cudaMemcpy(out_host + (size_t)d * width, h->v.qh + b * h->D + slot, sizeof(double), cudaMemcpyDeviceToHost));
		CU(cudaMemcpy(out_host + (size_t)d * width + 1, h->v.qw + (b * h->D + slot) * 2 * n, 2 * n * sizeof(double), cudaMemcpyDeviceToHost));
	}
	return JSDE_OK;
}

int jsde_draw_gauss_debug(jsde_t *, double, int, double *) { return unsupported("the generator debug stream"); }
// ---- checkpoint / restore: same contract as the lane library (include/jsde.h) ----------------------------------------
namespace {
struct WideCheckpointHeader {
	uint64_t magic;
	int32_t version, n, D, npar;
	int64_t B;
	Controller ctrl;
	double first_step;
};
constexpr uint64_t WIDE_CHECKPOINT_MAGIC = 0x4a53444557434b31ull;  // "JSDEWCK1"

Sections checkpoint_sections(jsde_handle *h)
{
	const size_t B = (size_t)h->B, n = Model::N, D = (size_t)h->D;
	WideView &v = h->v;
	return {
		{v.y, B * n * 8}, {v.t, B * 8}, {v.dt, B * 8}, {v.qh, B * D * 8}, {v.qw, B * D * 2 * n * 8}, {v.order, B * D * 4},
		{v.q_count, B * 4}, {v.key, B * MT_WORDS * 4}, {v.pos, B * 4}, {v.status, B * 4}, {v.accepted, B * 8}, {v.rejected, B * 8},
		{v.next_jump, B * 8}, {v.jump_set, B}, {v.tape_pos, B * 8},
		{v.sNY, B * n * 8}, {v.sErr, B * n * 8}, {v.h_last, B * 8}, {v.cursor, B * 4},
		{h->seeds, B * 4}, {h->par, (size_t)(Model::NPAR > 0 ? Model::NPAR : 1) * 8},
	};
}
}  // namespace

int64_t jsde_checkpoint_bytes(jsde_t *h)
{
	if (!h)
		return fail(JSDE_E_ARG, "Wrong input.");
	return (int64_t)(sizeof(WideCheckpointHeader) + sections_bytes(checkpoint_sections(h)));
}

int jsde_checkpoint_save(jsde_t *h, void *blob_host, int64_t bytes)
{
	if (!h || !blob_host || bytes != jsde_checkpoint_bytes(h))
		return fail(JSDE_E_ARG, "Wrong input. (checkpoint buffer size)");
	int rc = bind(h);
	if (rc) return rc;
	if ((rc = ensure_seeded(h))) return rc;
	WideCheckpointHeader hd{};
	hd.magic = WIDE_CHECKPOINT_MAGIC; hd.version = 2; hd.n = Model::N; hd.D = h->D; hd.npar = Model::NPAR; hd.B = h->B;
	hd.ctrl = h->ctrl; hd.first_step = h->first_step;
	char *out = static_cast<char *>(blob_host);
	memcpy(out, &hd, sizeof(hd));
	return save_sections(h, checkpoint_sections(h), out + sizeof(hd));
}

int jsde_checkpoint_restore(jsde_t *h, const void *blob_host, int64_t bytes)
{
	if (!h || !blob_host || bytes != jsde_checkpoint_bytes(h))
		return fail(JSDE_E_ARG, "Wrong input. (checkpoint buffer size)");
	WideCheckpointHeader hd;
	const char *in = static_cast<const char *>(blob_host);
	memcpy(&hd, in, sizeof(hd));
	if (hd.magic != WIDE_CHECKPOINT_MAGIC || hd.version != 2 || hd.n != Model::N || hd.B != h->B || hd.D != h->D || hd.npar != Model::NPAR)
		return fail(JSDE_E_ARG, "Wrong input. (checkpoint belongs to a different model or ensemble shape)");
	int rc = bind(h);
	if (rc) return rc;
	h->ctrl = hd.ctrl;
	h->first_step = hd.first_step;
	h->seed_dirty = false;
	h->v.par = h->par;  // shared parameters are part of the checkpoint; per-trajectory ones are set again by the caller
	h->v.par_stride = 0;
	return restore_sections(h, checkpoint_sections(h), in + sizeof(hd));
}

}  // extern "C"
