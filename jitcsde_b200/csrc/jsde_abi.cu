// extern "C" boundary (include/jsde.h).  One shared library per generated model: the build passes
// -DJSDE_MODEL_HEADER="<generated model header>" which defines jsde_model::Model.  Must be compiled with -fmad=false.
#include <cmath>

#include "kernels.cuh"

#ifndef JSDE_MODEL_HEADER
#error "compile with -DJSDE_MODEL_HEADER=\"model.cuh\""
#endif
#include JSDE_MODEL_HEADER

#include "abi_common.cuh"

using jsde_model::Model;
using namespace jsde;
using namespace jsde::abi;

struct jsde_handle : HandleBase {
	int resident_blocks = 0;  // blocks of k_integrate that fit on the device at once
	bool recycle = true;      // lanes take further trajectories from a pool (grid = one wave): see kernels.cuh
	EnsembleView v{};
	double *staging = nullptr;  // [B][max(n,1)] doubles: host-layout buffer for transposes and per-trajectory vectors
	double *par_rows = nullptr, *par_cols = nullptr;  // per-trajectory control parameters (allocated on first use)
};

namespace {

// drop the noise memory, restart the generator from the stored seeds (lazily: ensure_seeded), restart jump scheduling.
// dt is NOT touched: in the reference it lives on the Python object and survives reset_integrator (_jitcsde.py:242-246).
int reset_integrator(jsde_handle *h)
{
	const int64_t B = h->B;
	CU(cudaMemsetAsync(h->v.q_count, 0, B * sizeof(int), h->stream));
	CU(cudaMemsetAsync(h->v.q_cursor, 0, B * sizeof(int), h->stream));
	CU(cudaMemsetAsync(h->v.q_head, 0, B * sizeof(int), h->stream));
	CU(cudaMemsetAsync(h->v.status, 0, B * sizeof(int), h->stream));
	CU(cudaMemsetAsync(h->v.accepted, 0, B * sizeof(unsigned long long), h->stream));
	CU(cudaMemsetAsync(h->v.rejected, 0, B * sizeof(unsigned long long), h->stream));
	CU(cudaMemsetAsync(h->v.jump_set, 0, B, h->stream));
	CU(cudaMemsetAsync(h->v.tape_pos, 0, B * sizeof(long long), h->stream));
	CU(cudaMemsetAsync(h->v.new_y, 0, (size_t)B * Model::N * sizeof(double), h->stream));
	CU(cudaMemsetAsync(h->v.err, 0, (size_t)B * Model::N * sizeof(double), h->stream));
	CU(cudaMemsetAsync(h->v.rng_level, 0, B * sizeof(int), h->stream));
	h->seed_dirty = true;
	return JSDE_OK;
}

// rk_seed + first regeneration (2.6 GB written per 2^20 trajectories): once per (re)initialisation, before the first draw
int ensure_seeded(jsde_handle *h)
{
	if (h->seed_dirty) {
		k_seed<<<grid_for(h->B, 64), 64, 0, h->stream>>>(h->v.mt, h->v.mt_chunk, h->seeds, h->B);
		CU(cudaGetLastError());
		h->seed_dirty = false;
	}
	return JSDE_OK;
}

void bind_tape(jsde_handle *h)
{
	h->v.taus = h->taus;
	h->v.xis = h->xis;
	h->v.tape_len = h->tape_len;
}

}  // namespace

JSDE_DEFINE_COMMON_ENTRY_POINTS(Model)

extern "C" {

const char *jsde_version(void) { return "jsde-b200 0.1 (sm_100a, fp64, -fmad=false)"; }
int jsde_create(jsde_t **out, int device, int64_t B, int noise_capacity)
{
	if (!out || B <= 0 || B >= 0x7fffffff || noise_capacity < 0)
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc = check_device(device);
	if (rc) return rc;
	cudaError_t err = cudaSuccess;
	jsde_handle *h = new jsde_handle();
	h->device = device;
	h->B = B;
	h->D = noise_capacity ? noise_capacity : 32;
	if (h->D < 2) h->D = 2;
	constexpr int n = Model::N;
	auto bail = [&](int code) { jsde_destroy(h); return code; };
	if ((rc = bind(h)) || (rc = create_stream_and_events(h))) return bail(rc);
	EnsembleView &v = h->v;
	v.B = B;
	v.D = h->D;
	int ring = 2;
	while (ring < h->D) ring <<= 1;
	v.Dmask = ring - 1;
#define ALLOC(ptr, cnt) if ((rc = dev_alloc(h, &(ptr), (size_t)(cnt)))) return bail(rc)
	ALLOC(v.y, B * n); ALLOC(v.t, B); ALLOC(v.dt, B);
	ALLOC(v.new_y, B * n); ALLOC(v.err, B * n); ALLOC(v.new_t, B);
	ALLOC(v.q, (int64_t)ring * B * noise_item_doubles(n));
	ALLOC(v.q_count, B); ALLOC(v.q_cursor, B); ALLOC(v.q_head, B);
	ALLOC(v.mt, B * MT_CHUNKS); ALLOC(v.mt_chunk, B);
	ALLOC(v.rng_spill, (int64_t)WarpRng<Model::N>::CAP * 3 * B); ALLOC(v.rng_level, B);
	ALLOC(v.status, B); ALLOC(v.accepted, B); ALLOC(v.rejected, B);
	ALLOC(v.next_jump, B); ALLOC(v.jump_set, B); ALLOC(v.tape_pos, B);
	ALLOC(v.totals, 1);
	ALLOC(v.pool, B);
	ALLOC(h->seeds, B);
	ALLOC(h->par, Model::NPAR > 0 ? Model::NPAR : 1);
	ALLOC(h->staging, B * (n > 1 ? n : 1));
	ALLOC(h->shift, n);
#undef ALLOC
	v.par = h->par;
	v.par_stride = 0;
	bind_tape(h);
	v.generated_jumps = 0;
	v.jump_seed = 0;
	v.jump_first = 0;
	default_controller(h);
	if ((err = cudaMemsetAsync(h->seeds, 0, B * sizeof(uint32_t), h->stream)) != cudaSuccess ||
		(err = cudaMemsetAsync(v.y, 0, (size_t)B * n * sizeof(double), h->stream)) != cudaSuccess ||
		(err = cudaMemsetAsync(v.t, 0, B * sizeof(double), h->stream)) != cudaSuccess ||
		(err = cudaMemsetAsync(v.new_t, 0, B * sizeof(double), h->stream)) != cudaSuccess ||
		(err = cudaMemsetAsync(h->par, 0, (Model::NPAR > 0 ? Model::NPAR : 1) * sizeof(double), h->stream)) != cudaSuccess)
		return bail(fail(JSDE_E_CUDA, cudaGetErrorString(err)));
	k_fill<<<grid_for(B, 256), 256, 0, h->stream>>>(v.dt, h->first_step, B);
	if ((rc = reset_integrator(h))) return bail(rc);
	if ((err = cudaStreamSynchronize(h->stream)) != cudaSuccess)
		return bail(fail(JSDE_E_CUDA, cudaGetErrorString(err)));
	// the generator FIFOs + staging area exceed the 48 KB default for dynamic shared memory
	const int smem = (int)rng_smem_bytes<Model>();
	const int duo_smem = (int)duo_smem_bytes<Model>();
	if ((err = cudaFuncSetAttribute(k_integrate<Model, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, duo_smem)) != cudaSuccess ||
		(err = cudaFuncSetAttribute(k_integrate<Model, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, duo_smem)) != cudaSuccess ||
		(err = cudaFuncSetAttribute(k_integrate<Model, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, duo_smem)) != cudaSuccess ||
		(err = cudaFuncSetAttribute(k_integrate<Model, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, duo_smem)) != cudaSuccess ||
		(err = cudaFuncSetAttribute(k_get_next_step<Model>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess ||
		(err = cudaFuncSetAttribute(k_pin_noise<Model>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess ||
		(err = cudaFuncSetAttribute(k_draw_gauss<Model>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess)
		return bail(fail(JSDE_E_CUDA, cudaGetErrorString(err)));
	{
		int per_sm = 0;
		cudaDeviceProp prop;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_integrate<Model, false, true>, DUO_BLOCK, duo_smem) == cudaSuccess &&
			cudaGetDeviceProperties(&prop, device) == cudaSuccess && per_sm > 0)
			h->resident_blocks = per_sm * prop.multiProcessorCount;
		h->recycle = true;
		if (const char *env = getenv("JSDE_RECYCLE"))
			h->recycle = atoi(env) != 0;
		if (const char *env = getenv("JSDE_RESIDENT_BLOCKS"))  // test knob: a tiny grid makes small ensembles recycle lanes
			if (atoi(env) > 0) h->resident_blocks = atoi(env);
		// rounds (attempted steps) between two yields of a warp's lanes; ~18 ms of a full GPU's time at the default
		v.slice_rounds = 4096;
		if (const char *env = getenv("JSDE_SLICE"))
			if (atoi(env) >= 0) v.slice_rounds = atoi(env);
	}
	*out = h;
	return JSDE_OK;
}

int jsde_destroy(jsde_t *h)
{
	if (!h)
		return JSDE_OK;
	destroy_base(h);
	delete h;
	return JSDE_OK;
}

int jsde_set_option(jsde_t *h, const char *name, const char *value)
{
	if (!h || !name || !value)
		return fail(JSDE_E_ARG, "Wrong input.");
	if (!strcmp(name, "recycle") && (!strcmp(value, "0") || !strcmp(value, "1"))) {
		h->recycle = value[0] == '1';
		return JSDE_OK;
	}
	if (!strcmp(name, "slice")) {  // rounds between yields of recycled lanes; 0: trajectories never change hands unfinished
		char *end = nullptr;
		const long rounds = strtol(value, &end, 10);
		if (end == value || *end || rounds < 0 || rounds > 0x7fffffffL)
			return fail(JSDE_E_ARG, "Wrong input. (slice: a number of rounds >= 0)");
		h->v.slice_rounds = (int)rounds;
		return JSDE_OK;
	}
	if (!strcmp(name, "coupling") && !strcmp(value, "exact"))
		return JSDE_OK;
	if (!strcmp(name, "coupling"))
		return fail(JSDE_E_STATE, "coupling options exist for wide models with a dense coupling term only");
	return fail(JSDE_E_ARG, std::string("Wrong input. (unknown option ") + name + "=" + value + ")");
}

int jsde_set_state(jsde_t *h, const double *y0_host, const double *t0_host, double t0)
{
	if (!h || !y0_host)
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc = bind(h);
	if (rc) return rc;
	const int64_t B = h->B;
	constexpr int n = Model::N;
	CU(cudaMemcpyAsync(h->staging, y0_host, (size_t)B * n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
	k_rows_to_cols<<<grid_for(B * n, 256), 256, 0, h->stream>>>(h->staging, h->v.y, B, n, 0);
	if (t0_host)
		CU(cudaMemcpyAsync(h->v.t, t0_host, B * sizeof(double), cudaMemcpyHostToDevice, h->stream));
	else
		k_fill<<<grid_for(B, 256), 256, 0, h->stream>>>(h->v.t, t0, B);
	CU(cudaMemcpyAsync(h->v.new_t, h->v.t, B * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
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
	k_fill<<<grid_for(h->B, 256), 256, 0, h->stream>>>(h->v.dt, h->first_step, h->B);  // self.dt = first_step (:553)
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
	} else {
		if (!h->par_cols) {  // allocated once, reused by every later call (a parameter sweep must not grow the footprint)
			if ((rc = dev_alloc(h, &h->par_rows, (size_t)h->B * Model::NPAR))) return rc;
			if ((rc = dev_alloc(h, &h->par_cols, (size_t)h->B * Model::NPAR))) return rc;
		}
		CU(cudaMemcpyAsync(h->par_rows, params_host, (size_t)h->B * Model::NPAR * sizeof(double), cudaMemcpyHostToDevice, h->stream));
		k_rows_to_cols<<<grid_for(h->B * Model::NPAR, 256), 256, 0, h->stream>>>(h->par_rows, h->par_cols, h->B, Model::NPAR, 0);
		h->v.par = h->par_cols;
		h->v.par_stride = h->B;
	}
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(h->stream));
	return JSDE_OK;
}

int jsde_pin_noise(jsde_t *h, unsigned number, double step)
{
	if (!h || !(step > 0.0))
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc = bind(h);
	if (rc) return rc;
	if ((rc = ensure_seeded(h))) return rc;
	k_pin_noise<Model><<<grid_for(h->B, BLOCK), BLOCK, rng_smem_bytes<Model>(), h->stream>>>(h->v, number, step);
	CU(cudaGetLastError());
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
	k_pin_path<Model><<<grid_for(h->B, BLOCK), BLOCK, 0, h->stream>>>(h->v, number, steps, dw, dz);
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(h->stream));
	return JSDE_OK;
}

static int integrate_common(jsde_t *h, double target_time, int use_tape, jsde_stats *stats_host,
	const double *grid_times_dev = nullptr, int grid_count = 0, double *grid_out_dev = nullptr)
{
	if (!(target_time == target_time))
		return fail(JSDE_E_ARG, "Wrong input. (target time is NaN)");
	int rc = ensure_seeded(h);
	if (rc) return rc;
	CU(cudaMemsetAsync(h->v.totals, 0, sizeof(CallTotals), h->stream));
	CU(cudaEventRecord(h->ev0, h->stream));
	const unsigned all_blocks = grid_for(h->B, 32 * DUO_PAIRS);
	if (h->recycle && h->resident_blocks > 0 && all_blocks > (unsigned)h->resident_blocks) {
		// one wave of blocks: lanes take further trajectories from a pool as theirs finish (kernels.cuh)
		const unsigned grid = (unsigned)h->resident_blocks;
		const int64_t first_wave = (int64_t)grid * 32 * DUO_PAIRS;
		k_pool_init<<<grid_for(h->B - first_wave, 256), 256, 0, h->stream>>>(h->v.pool, (int)first_wave, (int)(h->B - first_wave), h->v.totals);
		if (grid_times_dev)
			k_integrate<Model, true, true><<<grid, DUO_BLOCK, duo_smem_bytes<Model>(), h->stream>>>(h->v, h->ctrl, target_time, use_tape,
				grid_times_dev, grid_count, grid_out_dev);
		else
			k_integrate<Model, false, true><<<grid, DUO_BLOCK, duo_smem_bytes<Model>(), h->stream>>>(h->v, h->ctrl, target_time, use_tape,
				nullptr, 0, nullptr);
	} else if (grid_times_dev)
		k_integrate<Model, true, false><<<all_blocks, DUO_BLOCK, duo_smem_bytes<Model>(), h->stream>>>(h->v, h->ctrl, target_time, use_tape,
			grid_times_dev, grid_count, grid_out_dev);
	else
		k_integrate<Model, false, false><<<all_blocks, DUO_BLOCK, duo_smem_bytes<Model>(), h->stream>>>(h->v, h->ctrl, target_time, use_tape,
			nullptr, 0, nullptr);
	CU(cudaGetLastError());
	return finish_call(h, h->v.totals, stats_host, 1);
}

// jump_mode: 0 none, 1 tape (set before), 2 generated variates
static int select_jump_source(jsde_t *h, int jump_mode, uint64_t seed, int64_t first_index, int *use_tape)
{
	*use_tape = 0;
	if (jump_mode == 0)
		return JSDE_OK;
	if (!Model::HAS_JUMP)
		return fail(JSDE_E_STATE, "this model was generated without a jump amplitude");
	if (jump_mode == 1) {
		if (!h->tape_set)
			return fail(JSDE_E_STATE, "no jump tape has been set (jsde_set_jump_tape)");
		bind_tape(h);
		h->v.generated_jumps = 0;
	} else if (jump_mode == 2) {
		if (first_index < 0)
			return fail(JSDE_E_ARG, "Wrong input.");
		h->v.generated_jumps = 1;
		h->v.jump_seed = seed;
		h->v.jump_first = first_index;
	} else
		return fail(JSDE_E_ARG, "Wrong input. (jump_mode)");
	*use_tape = 1;
	return JSDE_OK;
}

int jsde_set_jump_tape(jsde_t *h, const jsde_jump_tape *tape)
{
	if (!h)
		return fail(JSDE_E_ARG, "Wrong input.");
	if (!Model::HAS_JUMP)
		return fail(JSDE_E_STATE, "this model was generated without a jump amplitude");
	int rc = bind(h);
	if (rc) return rc;
	if ((rc = upload_tape(h, tape))) return rc;
	bind_tape(h);
	return JSDE_OK;
}

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
	int use_tape = 0;
	if ((rc = select_jump_source(h, jump_mode, seed, first_index, &use_tape))) return rc;
	constexpr int n = Model::N;
	const size_t count = (size_t)n_times * n * (size_t)h->B;
	Scratch scratch;
	double *times = nullptr, *cols = nullptr, *rows = nullptr;
	CU(scratch.get(&times, n_times));
	if (scratch.get(&cols, count) != cudaSuccess || scratch.get(&rows, (size_t)n * h->B) != cudaSuccess)
		return fail(JSDE_E_NOMEM, "sampling grid does not fit in device memory: n_times*n*B doubles");
	CU(cudaMemcpyAsync(times, times_host, n_times * sizeof(double), cudaMemcpyHostToDevice, h->stream));
	if ((rc = integrate_common(h, times_host[0], use_tape, stats_host, times, (int)n_times, cols))) return rc;
	// device layout [n_times][n][B] -> host layout [n_times][B][n], one sample time at a time
	for (int64_t j = 0; j < n_times; j++) {
		k_cols_to_rows<<<grid_for(h->B * n, 256), 256, 0, h->stream>>>(cols + (size_t)j * n * h->B, rows, h->B, n);
		CU(cudaMemcpyAsync(y_host + (size_t)j * n * h->B, rows, (size_t)n * h->B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
	}
	CU(cudaStreamSynchronize(h->stream));
	return JSDE_OK;
}

int jsde_integrate_grid(jsde_t *h, const double *times_host, int64_t n_times, const jsde_jump_tape *tape, double *y_host, jsde_stats *stats_host)
{
	if (!h)
		return fail(JSDE_E_ARG, "Wrong input.");
	if (tape) {
		int rc = jsde_set_jump_tape(h, tape);
		if (rc) return rc;
	}
	return jsde_integrate_grid_jumps(h, times_host, n_times, tape ? 1 : 0, 0, 0, y_host, stats_host);
}

int jsde_integrate(jsde_t *h, double target_time, jsde_stats *stats_host)
{
	if (!h)
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc = bind(h);
	if (rc) return rc;
	return integrate_common(h, target_time, 0, stats_host);
}

int jsde_integrate_jumps(jsde_t *h, double target_time, const jsde_jump_tape *tape, jsde_stats *stats_host)
{
	if (!h)
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc = bind(h);
	if (rc) return rc;
	if (tape && (rc = jsde_set_jump_tape(h, tape))) return rc;
	int use_tape = 0;
	if ((rc = select_jump_source(h, 1, 0, 0, &use_tape))) return rc;
	return integrate_common(h, target_time, use_tape, stats_host);
}

int jsde_integrate_jumps_generated(jsde_t *h, double target_time, uint64_t seed, int64_t first_index, jsde_stats *stats_host)
{
	if (!h)
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc = bind(h);
	if (rc) return rc;
	int use_tape = 0;
	if ((rc = select_jump_source(h, 2, seed, first_index, &use_tape))) return rc;
	return integrate_common(h, target_time, use_tape, stats_host);
}

int jsde_get_state(jsde_t *h, double *y_host, double *t_host, int32_t *status_host)
{
	if (!h)
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc = bind(h);
	if (rc) return rc;
	const int64_t B = h->B;
	constexpr int n = Model::N;
	if (y_host) {
		k_cols_to_rows<<<grid_for(B * n, 256), 256, 0, h->stream>>>(h->v.y, h->staging, B, n);
		CU(cudaGetLastError());
		CU(cudaMemcpyAsync(y_host, h->staging, (size_t)B * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
	}
	if (t_host)
		CU(cudaMemcpyAsync(t_host, h->v.t, B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
	if (status_host)
		CU(cudaMemcpyAsync(status_host, h->v.status, B * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
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
	const int threads = 256;
	unsigned blocks = grid_for(h->B, threads);
	if (blocks > 148 * 4) blocks = 148 * 4;
	k_moments<<<blocks, threads, (threads / 32) * (1 + 2 * n) * sizeof(double), h->stream>>>(h->v.y, h->v.status, h->B, n, h->shift, acc_dev, h->B, 1);
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
	CU(cudaMemcpyAsync(h->staging, h_host, h->B * sizeof(double), cudaMemcpyHostToDevice, h->stream));
	k_get_next_step<Model><<<grid_for(h->B, BLOCK), BLOCK, rng_smem_bytes<Model>(), h->stream>>>(h->v, h->staging);
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(h->stream));
	return JSDE_OK;
}

int jsde_get_p(jsde_t *h, double atol, double rtol, double *p_host)
{
	if (!h || !p_host)
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc = bind(h);
	if (rc) return rc;
	k_get_p<Model><<<grid_for(h->B, BLOCK), BLOCK, 0, h->stream>>>(h->v, atol, rtol, h->staging);
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
	k_accept_step<Model><<<grid_for(h->B, BLOCK), BLOCK, 0, h->stream>>>(h->v, mask);
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
	constexpr int n = Model::N;
	CU(cudaMemcpyAsync(h->staging, change_host, (size_t)h->B * n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
	k_rows_to_cols<<<grid_for(h->B * n, 256), 256, 0, h->stream>>>(h->staging, h->v.y, h->B, n, 1);
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

int jsde_draw_gauss_debug(jsde_t *h, double scale, int pairs, double *out_host)
{
	if (!h || pairs < 0 || !out_host)
		return fail(JSDE_E_ARG, "Wrong input.");
	int rc = bind(h);
	if (rc) return rc;
	if ((rc = ensure_seeded(h))) return rc;
	Scratch scratch;
	double *out = nullptr;
	const size_t count = (size_t)h->B * pairs * 2;
	CU(scratch.get(&out, count));
	k_draw_gauss<Model><<<grid_for(h->B, BLOCK), BLOCK, rng_smem_bytes<Model>(), h->stream>>>(h->v, scale, pairs, out);
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(out_host, out, count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
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
	int head = 0;
	CU(cudaMemcpy(&head, h->v.q_head + b, sizeof(int), cudaMemcpyDeviceToHost));
	const int width = 1 + 2 * n;
	for (int d = 0; d < *count && d < max_items; d++) {
		const int64_t slot = (head + d) & h->v.Dmask;
		CU(cudaMemcpy(out_host + (size_t)d * width, h->v.q + (slot * h->B + b) * noise_item_doubles(n), width * sizeof(double), cudaMemcpyDeviceToHost));
	}
	return JSDE_OK;
}

// ---- checkpoint / restore (SURVEY.md §8 f item 4) ---------------------------------------------------------------------
// Everything a later jsde_integrate* needs to continue bit for bit: states, times, step sizes, the last proposal, the
// noise memory, every trajectory's generator state with its queued pairs, statuses, counters, jump scheduling, the
// controller, seeds and shared control parameters.  Per-trajectory parameters and jump tapes are inputs the caller
// passes again.  Layout: header {magic, version, n, B, D, ring, CAP, additive} then the arrays in the order below.
namespace {

struct CheckpointHeader {
	uint64_t magic;
	int32_t version, n, additive, D, ring, cap, npar, reserved;
	int64_t B;
	Controller ctrl;
	double first_step;
};
constexpr uint64_t CHECKPOINT_MAGIC = 0x4a53444543484b31ull;  // "JSDECHK1"

Sections checkpoint_sections(jsde_handle *h)
{
	const size_t B = (size_t)h->B, n = Model::N, ring = (size_t)h->v.Dmask + 1;
	EnsembleView &v = h->v;
	return {
		{v.y, B * n * 8}, {v.t, B * 8}, {v.dt, B * 8}, {v.new_y, B * n * 8}, {v.err, B * n * 8}, {v.new_t, B * 8},
		{v.q, ring * B * noise_item_doubles((int)n) * 8}, {v.q_count, B * 4}, {v.q_cursor, B * 4}, {v.q_head, B * 4},
		{v.mt, B * MT_CHUNKS * sizeof(uint4)}, {v.mt_chunk, B * 4},
		{v.rng_spill, (size_t)WarpRng<Model::N>::CAP * 3 * B * 8}, {v.rng_level, B * 4},
		{v.status, B * 4}, {v.accepted, B * 8}, {v.rejected, B * 8},
		{v.next_jump, B * 8}, {v.jump_set, B}, {v.tape_pos, B * 8},
		{h->seeds, B * 4}, {h->par, (size_t)(Model::NPAR > 0 ? Model::NPAR : 1) * 8},
	};
}

}  // namespace

int64_t jsde_checkpoint_bytes(jsde_t *h)
{
	if (!h)
		return fail(JSDE_E_ARG, "Wrong input.");
	return (int64_t)(sizeof(CheckpointHeader) + sections_bytes(checkpoint_sections(h)));
}

int jsde_checkpoint_save(jsde_t *h, void *blob_host, int64_t bytes)
{
	if (!h || !blob_host || bytes != jsde_checkpoint_bytes(h))
		return fail(JSDE_E_ARG, "Wrong input. (checkpoint buffer size)");
	int rc = bind(h);
	if (rc) return rc;
	if ((rc = ensure_seeded(h))) return rc;  // the blob must hold the generator state the next draw would use
	CheckpointHeader hd{};
	hd.magic = CHECKPOINT_MAGIC; hd.version = 2; hd.n = Model::N; hd.additive = Model::ADDITIVE; hd.D = h->D; hd.ring = h->v.Dmask + 1;
	hd.cap = WarpRng<Model::N>::CAP; hd.npar = Model::NPAR; hd.B = h->B; hd.ctrl = h->ctrl; hd.first_step = h->first_step;
	char *out = static_cast<char *>(blob_host);
	memcpy(out, &hd, sizeof(hd));
	return save_sections(h, checkpoint_sections(h), out + sizeof(hd));
}

int jsde_checkpoint_restore(jsde_t *h, const void *blob_host, int64_t bytes)
{
	if (!h || !blob_host || bytes != jsde_checkpoint_bytes(h))
		return fail(JSDE_E_ARG, "Wrong input. (checkpoint buffer size)");
	CheckpointHeader hd;
	const char *in = static_cast<const char *>(blob_host);
	memcpy(&hd, in, sizeof(hd));
	if (hd.magic != CHECKPOINT_MAGIC || hd.version != 2 || hd.n != Model::N || hd.additive != (int)Model::ADDITIVE || hd.B != h->B ||
		hd.D != h->D || hd.ring != h->v.Dmask + 1 || hd.cap != WarpRng<Model::N>::CAP || hd.npar != Model::NPAR)
		return fail(JSDE_E_ARG, "Wrong input. (checkpoint belongs to a different model or ensemble shape)");
	int rc = bind(h);
	if (rc) return rc;
	h->ctrl = hd.ctrl;
	h->first_step = hd.first_step;
	h->seed_dirty = false;  // the generator state comes from the blob
	h->v.par = h->par;      // shared parameters are part of the checkpoint; per-trajectory ones are set again by the caller
	h->v.par_stride = 0;
	return restore_sections(h, checkpoint_sections(h), in + sizeof(hd));
}

}  // extern "C"
