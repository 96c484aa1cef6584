// extern "C" boundary (include/jsde.h).  One shared library per generated model: the build passes
// -DJSDE_MODEL_HEADER="<generated model header>" which defines jsde_model::Model.  Must be compiled with -fmad=false.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/jsde.h"
#include "kernels.cuh"

#ifndef JSDE_MODEL_HEADER
#error "compile with -DJSDE_MODEL_HEADER=\"model.cuh\""
#endif
#include JSDE_MODEL_HEADER

using jsde_model::Model;
using namespace jsde;

namespace {

thread_local std::string g_error;

int fail(int code, const std::string &msg)
{
	g_error = msg;
	return code;
}

#define CU(call)                                                                                          \
	do {                                                                                                  \
		cudaError_t err__ = (call);                                                                       \
		if (err__ != cudaSuccess)                                                                         \
			return fail(err__ == cudaErrorMemoryAllocation ? JSDE_E_NOMEM : JSDE_E_CUDA,                  \
				std::string(#call) + ": " + cudaGetErrorString(err__));                                   \
	} while (0)

inline unsigned grid_for(int64_t count, int block) { return (unsigned)((count + block - 1) / block); }

}  // namespace

struct jsde_handle {
	int device = 0;
	int64_t B = 0;
	int D = 0;
	int resident_blocks = 0;  // blocks of k_integrate that fit on the device at once
	bool recycle = false;     // lanes take further trajectories from a pool (grid = one wave): see kernels.cuh
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	EnsembleView v{};
	Controller ctrl{};
	double first_step = 1.0;
	uint32_t *seeds = nullptr;
	double *par = nullptr;
	double *staging = nullptr;  // [B][max(n,1)] doubles: host-layout buffer for transposes and per-trajectory vectors
	double *taus = nullptr, *xis = nullptr;
	const double *tape_src_taus = nullptr, *tape_src_xis = nullptr;
	long long tape_len = 0;
	double *shift = nullptr;
	std::vector<void *> allocations;
};

namespace {

template <class T>
int dev_alloc(jsde_handle *h, T **ptr, size_t count)
{
	void *p = nullptr;
	CU(cudaMalloc(&p, count * sizeof(T) > 0 ? count * sizeof(T) : 16));
	h->allocations.push_back(p);
	*ptr = static_cast<T *>(p);
	return JSDE_OK;
}

int bind(jsde_handle *h)
{
	CU(cudaSetDevice(h->device));
	return JSDE_OK;
}

void default_controller(jsde_handle *h)
{
	// defaults of set_integration_parameters (_jitcsde.py:491-502), q = 1.5 (:564)
	Controller &c = h->ctrl;
	c.atol = 1e-10; c.rtol = 1e-5; c.min_step = 1e-10; c.max_step = 10.0; c.dec_thr = 1.1; c.inc_thr = 0.5;
	c.safety = 0.9; c.max_factor = 5.0; c.min_factor = 0.2;
	c.shrink_exp = -1 / 1.5;
	c.grow_exp = -1 / (1.5 + 1);
	h->first_step = 1.0;
}

// drop the noise memory, restart the generator from the stored seeds, restart jump scheduling.  dt is NOT touched:
// in the reference it lives on the Python object and survives reset_integrator (_jitcsde.py:242-246).
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
	k_seed<<<grid_for(B, 64), 64, 0, h->stream>>>(h->v.mt, h->v.mt_chunk, h->seeds, B);
	CU(cudaGetLastError());
	return JSDE_OK;
}

int finish_call(jsde_handle *h, jsde_stats *stats, int launches)
{
	CallTotals tot;
	CU(cudaEventRecord(h->ev1, h->stream));
	CU(cudaMemcpyAsync(&tot, h->v.totals, sizeof(tot), cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	if (stats) {
		float ms = 0.f;
		CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
		stats->accepted = tot.accepted;
		stats->rejected = tot.rejected;
		stats->fresh_draws = tot.fresh_draws;
		stats->jumps = tot.jumps;
		stats->n_min_step = tot.n_min_step;
		stats->n_overflow = tot.n_overflow;
		stats->max_noise_depth = tot.max_depth;
		stats->kernel_launches = launches;
		stats->kernel_ms = ms;
	}
	return JSDE_OK;
}

}  // namespace

extern "C" {

const char *jsde_last_error(void) { return g_error.c_str(); }
const char *jsde_version(void) { return "jsde-b200 0.1 (sm_100a, fp64, -fmad=false)"; }
const char *jsde_model_name(void) { return JSDE_MODEL_NAME; }

int jsde_model_info(int *n, int *additive, int *n_params, int *has_jump)
{
	if (n) *n = Model::N;
	if (additive) *additive = Model::ADDITIVE ? 1 : 0;
	if (n_params) *n_params = Model::NPAR;
	if (has_jump) *has_jump = Model::HAS_JUMP ? 1 : 0;
	return JSDE_OK;
}

int jsde_create(jsde_t **out, int device, int64_t B, int noise_capacity)
{
	if (!out || B <= 0 || B >= 0x7fffffff || noise_capacity < 0)
		return fail(JSDE_E_ARG, "Wrong input.");
	int count = 0;
	cudaError_t err = cudaGetDeviceCount(&count);
	if (err != cudaSuccess || count == 0)
		return fail(JSDE_E_CUDA, std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(err));
	if (device < 0 || device >= count)
		return fail(JSDE_E_ARG, "Wrong input. (device index out of range)");
	jsde_handle *h = new jsde_handle();
	h->device = device;
	h->B = B;
	h->D = noise_capacity ? noise_capacity : 32;
	if (h->D < 2) h->D = 2;
	constexpr int n = Model::N;
	int rc = bind(h);
	auto bail = [&](int code) { jsde_destroy(h); return code; };
	if (rc) return bail(rc);
	if ((err = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess ||
		(err = cudaEventCreate(&h->ev0)) != cudaSuccess || (err = cudaEventCreate(&h->ev1)) != cudaSuccess)
		return bail(fail(JSDE_E_CUDA, cudaGetErrorString(err)));
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
	ALLOC(h->seeds, B);
	ALLOC(h->par, Model::NPAR > 0 ? Model::NPAR : 1);
	ALLOC(h->staging, B * (n > 1 ? n : 1));
	ALLOC(h->shift, n);
#undef ALLOC
	v.par = h->par;
	v.par_stride = 0;
	v.taus = v.xis = nullptr;
	v.tape_len = 0;
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
		h->recycle = !Model::ADDITIVE;
		if (const char *env = getenv("JSDE_RECYCLE"))
			h->recycle = atoi(env) != 0;
		if (const char *env = getenv("JSDE_RESIDENT_BLOCKS"))  // test knob: a tiny grid makes small ensembles recycle lanes
			if (atoi(env) > 0) h->resident_blocks = atoi(env);
	}
	*out = h;
	return JSDE_OK;
}

int jsde_destroy(jsde_t *h)
{
	if (!h)
		return JSDE_OK;
	cudaSetDevice(h->device);
	if (h->stream) cudaStreamSynchronize(h->stream);
	for (void *p : h->allocations) cudaFree(p);
	if (h->ev0) cudaEventDestroy(h->ev0);
	if (h->ev1) cudaEventDestroy(h->ev1);
	if (h->stream) cudaStreamDestroy(h->stream);
	delete h;
	return JSDE_OK;
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
	// the assertions of _jitcsde.py:543-549
	if (!(p->decrease_threshold >= 1.0) || !(p->increase_threshold <= 1.0) || !(p->max_factor >= 1.0) || !(p->min_factor <= 1.0) ||
		!(p->safety_factor <= 1.0) || !(p->atol >= 0.0) || !(p->rtol >= 0.0) || !(p->q > 0.0))
		return fail(JSDE_E_ARG, "Wrong input. (integration parameter out of range)");
	int rc = bind(h);
	if (rc) return rc;
	double first = p->first_step, min_step = p->min_step;
	if (first > p->max_step) first = p->max_step;  // :536-541
	if (min_step > first) min_step = first;
	Controller &c = h->ctrl;
	c.atol = p->atol; c.rtol = p->rtol; c.min_step = min_step; c.max_step = p->max_step;
	c.dec_thr = p->decrease_threshold; c.inc_thr = p->increase_threshold; c.safety = p->safety_factor;
	c.max_factor = p->max_factor; c.min_factor = p->min_factor;
	c.shrink_exp = -1 / p->q;
	c.grow_exp = -1 / (p->q + 1);
	h->first_step = first;
	k_fill<<<grid_for(h->B, 256), 256, 0, h->stream>>>(h->v.dt, first, h->B);  // self.dt = first_step (:553)
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
		double *rows = nullptr, *cols = nullptr;
		if ((rc = dev_alloc(h, &rows, (size_t)h->B * Model::NPAR))) return rc;
		if ((rc = dev_alloc(h, &cols, (size_t)h->B * Model::NPAR))) return rc;
		CU(cudaMemcpyAsync(rows, params_host, (size_t)h->B * Model::NPAR * sizeof(double), cudaMemcpyHostToDevice, h->stream));
		k_rows_to_cols<<<grid_for(h->B * Model::NPAR, 256), 256, 0, h->stream>>>(rows, cols, h->B, Model::NPAR, 0);
		h->v.par = cols;
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
	double *steps = nullptr, *dw = nullptr, *dz = nullptr;
	cudaError_t err = cudaMalloc(&steps, number * sizeof(double));
	if (err == cudaSuccess) err = cudaMalloc(&dw, values * sizeof(double));
	if (err == cudaSuccess) err = cudaMalloc(&dz, values * sizeof(double));
	auto cleanup = [&]() { cudaFree(steps); cudaFree(dw); cudaFree(dz); };
	if (err != cudaSuccess) { cleanup(); return fail(JSDE_E_NOMEM, cudaGetErrorString(err)); }
	err = cudaMemcpyAsync(steps, steps_host, number * sizeof(double), cudaMemcpyHostToDevice, h->stream);
	if (err == cudaSuccess) err = cudaMemcpyAsync(dw, dw_host, values * sizeof(double), cudaMemcpyHostToDevice, h->stream);
	if (err == cudaSuccess) err = cudaMemcpyAsync(dz, dz_host, values * sizeof(double), cudaMemcpyHostToDevice, h->stream);
	if (err == cudaSuccess) {
		k_pin_path<Model><<<grid_for(h->B, BLOCK), BLOCK, 0, h->stream>>>(h->v, number, steps, dw, dz);
		err = cudaGetLastError();
	}
	if (err == cudaSuccess) err = cudaStreamSynchronize(h->stream);
	cleanup();
	if (err != cudaSuccess)
		return fail(JSDE_E_CUDA, cudaGetErrorString(err));
	return JSDE_OK;
}

static int integrate_common(jsde_t *h, double target_time, int use_tape, jsde_stats *stats_host,
	const double *grid_times_dev = nullptr, int grid_count = 0, double *grid_out_dev = nullptr)
{
	if (!(target_time == target_time))
		return fail(JSDE_E_ARG, "Wrong input. (target time is NaN)");
	CU(cudaMemsetAsync(h->v.totals, 0, sizeof(CallTotals), h->stream));
	CU(cudaEventRecord(h->ev0, h->stream));
	const unsigned all_blocks = grid_for(h->B, 32 * DUO_PAIRS);
	if (h->recycle && h->resident_blocks > 0 && all_blocks > (unsigned)h->resident_blocks) {
		// one wave of blocks: lanes take further trajectories from a pool as theirs finish (kernels.cuh)
		const unsigned grid = (unsigned)h->resident_blocks;
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
	return finish_call(h, stats_host, 1);
}

int jsde_integrate_grid(jsde_t *h, const double *times_host, int64_t n_times, const jsde_jump_tape *tape, double *y_host, jsde_stats *stats_host)
{
	if (!h || !times_host || n_times <= 0 || n_times > 0x7fffffff || !y_host)
		return fail(JSDE_E_ARG, "Wrong input.");
	for (int64_t j = 0; j < n_times; j++)
		if (!(times_host[j] == times_host[j]))
			return fail(JSDE_E_ARG, "Wrong input. (sample time is NaN)");
	if (tape && !Model::HAS_JUMP)
		return fail(JSDE_E_STATE, "this model was generated without a jump amplitude");
	int rc = bind(h);
	if (rc) return rc;
	constexpr int n = Model::N;
	const size_t count = (size_t)n_times * n * (size_t)h->B;
	double *times = nullptr, *cols = nullptr, *rows = nullptr;
	CU(cudaMalloc(&times, n_times * sizeof(double)));
	cudaError_t err = cudaMalloc(&cols, count * sizeof(double));
	if (err == cudaSuccess) err = cudaMalloc(&rows, (size_t)n * h->B * sizeof(double));
	if (err != cudaSuccess) {
		cudaFree(times); cudaFree(cols);
		return fail(JSDE_E_NOMEM, "sampling grid does not fit in device memory: n_times*n*B doubles");
	}
	auto cleanup = [&]() { cudaFree(times); cudaFree(cols); cudaFree(rows); };
	if ((err = cudaMemcpyAsync(times, times_host, n_times * sizeof(double), cudaMemcpyHostToDevice, h->stream)) != cudaSuccess) {
		cleanup();
		return fail(JSDE_E_CUDA, cudaGetErrorString(err));
	}
	int use_tape = 0;
	if (tape) {
		jsde_stats dummy;
		// upload / reuse the tape exactly like jsde_integrate_jumps does, without integrating: target = -inf is a no-op
		(void)dummy;
		if (tape->taus_host != h->tape_src_taus || tape->xis_host != h->tape_src_xis || tape->length != h->tape_len) {
			const size_t tcount = (size_t)h->B * (size_t)tape->length;
			if ((rc = dev_alloc(h, &h->taus, tcount)) || (rc = dev_alloc(h, &h->xis, tcount))) { cleanup(); return rc; }
			cudaMemcpyAsync(h->taus, tape->taus_host, tcount * sizeof(double), cudaMemcpyHostToDevice, h->stream);
			cudaMemcpyAsync(h->xis, tape->xis_host, tcount * sizeof(double), cudaMemcpyHostToDevice, h->stream);
			h->tape_src_taus = tape->taus_host;
			h->tape_src_xis = tape->xis_host;
			h->tape_len = tape->length;
			h->v.taus = h->taus;
			h->v.xis = h->xis;
			h->v.tape_len = tape->length;
		}
		use_tape = 1;
	}
	rc = integrate_common(h, times_host[0], use_tape, stats_host, times, (int)n_times, cols);
	if (rc == JSDE_OK) {
		// device layout [n_times][n][B] -> host layout [n_times][B][n], one sample time at a time
		for (int64_t j = 0; j < n_times && err == cudaSuccess; j++) {
			k_cols_to_rows<<<grid_for(h->B * n, 256), 256, 0, h->stream>>>(cols + (size_t)j * n * h->B, rows, h->B, n);
			err = cudaMemcpyAsync(y_host + (size_t)j * n * h->B, rows, (size_t)n * h->B * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
		}
		if (err == cudaSuccess) err = cudaStreamSynchronize(h->stream);
		if (err != cudaSuccess) rc = fail(JSDE_E_CUDA, cudaGetErrorString(err));
	}
	cleanup();
	return rc;
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
	if (!h || !tape || tape->length < 0 || (tape->length > 0 && (!tape->taus_host || !tape->xis_host)))
		return fail(JSDE_E_ARG, "Wrong input.");
	if (!Model::HAS_JUMP)
		return fail(JSDE_E_STATE, "this model was generated without a jump amplitude");
	int rc = bind(h);
	if (rc) return rc;
	if (tape->taus_host != h->tape_src_taus || tape->xis_host != h->tape_src_xis || tape->length != h->tape_len) {
		const size_t count = (size_t)h->B * (size_t)tape->length;
		if ((rc = dev_alloc(h, &h->taus, count))) return rc;
		if ((rc = dev_alloc(h, &h->xis, count))) return rc;
		CU(cudaMemcpyAsync(h->taus, tape->taus_host, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
		CU(cudaMemcpyAsync(h->xis, tape->xis_host, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
		h->tape_src_taus = tape->taus_host;
		h->tape_src_xis = tape->xis_host;
		h->tape_len = tape->length;
		h->v.taus = h->taus;
		h->v.xis = h->xis;
		h->v.tape_len = tape->length;
	}
	h->v.generated_jumps = 0;
	return integrate_common(h, target_time, 1, stats_host);
}

int jsde_integrate_jumps_generated(jsde_t *h, double target_time, uint64_t seed, int64_t first_index, jsde_stats *stats_host)
{
	if (!h || first_index < 0)
		return fail(JSDE_E_ARG, "Wrong input.");
	if (!Model::HAS_JUMP)
		return fail(JSDE_E_STATE, "this model was generated without a jump amplitude");
	int rc = bind(h);
	if (rc) return rc;
	h->v.generated_jumps = 1;
	h->v.jump_seed = seed;
	h->v.jump_first = first_index;
	return integrate_common(h, target_time, 1, stats_host);
}

int jsde_generated_jump_tape(int device, uint64_t seed, int64_t first_index, int64_t count, int64_t length, double *taus_host, double *xis_host)
{
	if (first_index < 0 || count < 0 || length < 0 || !taus_host || !xis_host)
		return fail(JSDE_E_ARG, "Wrong input.");
	if (count * length == 0)
		return JSDE_OK;
	CU(cudaSetDevice(device));
	const size_t bytes = (size_t)count * (size_t)length * sizeof(double);
	double *taus = nullptr, *xis = nullptr;
	cudaError_t err = cudaMalloc(&taus, bytes);
	if (err == cudaSuccess) err = cudaMalloc(&xis, bytes);
	if (err == cudaSuccess) {
		k_jump_tape<<<grid_for(count * length, 256), 256>>>(seed, first_index, count, length, taus, xis);
		err = cudaGetLastError();
	}
	if (err == cudaSuccess) err = cudaMemcpy(taus_host, taus, bytes, cudaMemcpyDeviceToHost);
	if (err == cudaSuccess) err = cudaMemcpy(xis_host, xis, bytes, cudaMemcpyDeviceToHost);
	cudaFree(taus); cudaFree(xis);
	if (err != cudaSuccess)
		return fail(JSDE_E_CUDA, cudaGetErrorString(err));
	return JSDE_OK;
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
	double *out = nullptr;
	const size_t count = (size_t)h->B * pairs * 2;
	CU(cudaMalloc(&out, count * sizeof(double) + 16));
	k_draw_gauss<Model><<<grid_for(h->B, BLOCK), BLOCK, rng_smem_bytes<Model>(), h->stream>>>(h->v, scale, pairs, out);
	cudaError_t err = cudaGetLastError();
	if (err == cudaSuccess) err = cudaMemcpyAsync(out_host, out, count * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
	if (err == cudaSuccess) err = cudaStreamSynchronize(h->stream);
	cudaFree(out);
	if (err != cudaSuccess)
		return fail(JSDE_E_CUDA, cudaGetErrorString(err));
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

std::vector<std::pair<void *, size_t>> checkpoint_sections(jsde_handle *h)
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
	size_t total = sizeof(CheckpointHeader);
	for (auto &s : checkpoint_sections(h))
		total += s.second;
	return (int64_t)total;
}

int jsde_checkpoint_save(jsde_t *h, void *blob_host, int64_t bytes)
{
	if (!h || !blob_host || bytes != jsde_checkpoint_bytes(h))
		return fail(JSDE_E_ARG, "Wrong input. (checkpoint buffer size)");
	int rc = bind(h);
	if (rc) return rc;
	CheckpointHeader hd{};
	hd.magic = CHECKPOINT_MAGIC; hd.version = 1; hd.n = Model::N; hd.additive = Model::ADDITIVE; hd.D = h->D; hd.ring = h->v.Dmask + 1;
	hd.cap = WarpRng<Model::N>::CAP; hd.npar = Model::NPAR; hd.B = h->B; hd.ctrl = h->ctrl; hd.first_step = h->first_step;
	char *out = static_cast<char *>(blob_host);
	memcpy(out, &hd, sizeof(hd));
	out += sizeof(hd);
	CU(cudaStreamSynchronize(h->stream));
	for (auto &s : checkpoint_sections(h)) {
		CU(cudaMemcpyAsync(out, s.first, s.second, cudaMemcpyDeviceToHost, h->stream));
		out += s.second;
	}
	CU(cudaStreamSynchronize(h->stream));
	return JSDE_OK;
}

int jsde_checkpoint_restore(jsde_t *h, const void *blob_host, int64_t bytes)
{
	if (!h || !blob_host || bytes != jsde_checkpoint_bytes(h))
		return fail(JSDE_E_ARG, "Wrong input. (checkpoint buffer size)");
	CheckpointHeader hd;
	const char *in = static_cast<const char *>(blob_host);
	memcpy(&hd, in, sizeof(hd));
	in += sizeof(hd);
	if (hd.magic != CHECKPOINT_MAGIC || hd.version != 1 || hd.n != Model::N || hd.additive != (int)Model::ADDITIVE || hd.B != h->B ||
		hd.D != h->D || hd.ring != h->v.Dmask + 1 || hd.cap != WarpRng<Model::N>::CAP || hd.npar != Model::NPAR)
		return fail(JSDE_E_ARG, "Wrong input. (checkpoint belongs to a different model or ensemble shape)");
	int rc = bind(h);
	if (rc) return rc;
	h->ctrl = hd.ctrl;
	h->first_step = hd.first_step;
	for (auto &s : checkpoint_sections(h)) {
		CU(cudaMemcpyAsync(s.first, in, s.second, cudaMemcpyHostToDevice, h->stream));
		in += s.second;
	}
	h->v.par = h->par;  // shared parameters are part of the checkpoint; per-trajectory ones are set again by the caller
	h->v.par_stride = 0;
	CU(cudaStreamSynchronize(h->stream));
	return JSDE_OK;
}

int jsde_measure_fp64_peak(int device, double *tflops)
{
	if (!tflops)
		return fail(JSDE_E_ARG, "Wrong input.");
	CU(cudaSetDevice(device));
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, device));
	double *sink = nullptr;
	CU(cudaMalloc(&sink, 64));
	cudaEvent_t a, b;
	CU(cudaEventCreate(&a));
	CU(cudaEventCreate(&b));
	const int iters = 1 << 15, threads = 256, blocks = prop.multiProcessorCount * 8;
	double best = 0.0;
	for (int rep = 0; rep < 5; rep++) {
		CU(cudaEventRecord(a));
		k_dfma_peak<<<blocks, threads>>>(sink, iters);
		CU(cudaEventRecord(b));
		CU(cudaEventSynchronize(b));
		float ms = 0.f;
		CU(cudaEventElapsedTime(&ms, a, b));
		const double tf = 2.0 * 8.0 * iters * (double)threads * blocks / (ms * 1e-3) / 1e12;
		if (rep > 0 && tf > best) best = tf;
	}
	cudaEventDestroy(a);
	cudaEventDestroy(b);
	cudaFree(sink);
	*tflops = best;
	return JSDE_OK;
}

static int debug_unary(int device, const double *x_host, double *out_host, int64_t count, int which, double yexp)
{
	if (!x_host || !out_host || count < 0)
		return fail(JSDE_E_ARG, "Wrong input.");
	CU(cudaSetDevice(device));
	double *x = nullptr, *o = nullptr;
	CU(cudaMalloc(&x, count * sizeof(double) + 16));
	CU(cudaMalloc(&o, count * sizeof(double) + 16));
	CU(cudaMemcpy(x, x_host, count * sizeof(double), cudaMemcpyHostToDevice));
	if (which == 0)
		k_debug_log<<<grid_for(count, 256), 256>>>(x, o, count);
	else
		k_debug_pow<<<grid_for(count, 256), 256>>>(x, yexp, o, count);
	cudaError_t err = cudaGetLastError();
	if (err == cudaSuccess) err = cudaMemcpy(out_host, o, count * sizeof(double), cudaMemcpyDeviceToHost);
	cudaFree(x);
	cudaFree(o);
	if (err != cudaSuccess)
		return fail(JSDE_E_CUDA, cudaGetErrorString(err));
	return JSDE_OK;
}

int jsde_debug_shared_div(int device, const double *a_host, const double *b_host, double *out_host, unsigned char *tame_host, int64_t count)
{
	if (!a_host || !b_host || !out_host || !tame_host || count < 0)
		return fail(JSDE_E_ARG, "Wrong input.");
	CU(cudaSetDevice(device));
	double *a = nullptr, *b = nullptr, *o = nullptr;
	unsigned char *t = nullptr;
	CU(cudaMalloc(&a, count * sizeof(double) + 16));
	CU(cudaMalloc(&b, count * sizeof(double) + 16));
	CU(cudaMalloc(&o, count * sizeof(double) + 16));
	CU(cudaMalloc(&t, count + 16));
	cudaError_t err = cudaMemcpy(a, a_host, count * sizeof(double), cudaMemcpyHostToDevice);
	if (err == cudaSuccess) err = cudaMemcpy(b, b_host, count * sizeof(double), cudaMemcpyHostToDevice);
	if (err == cudaSuccess) {
		k_debug_shared_div<<<grid_for(count, 256), 256>>>(a, b, o, t, count);
		err = cudaGetLastError();
	}
	if (err == cudaSuccess) err = cudaMemcpy(out_host, o, count * sizeof(double), cudaMemcpyDeviceToHost);
	if (err == cudaSuccess) err = cudaMemcpy(tame_host, t, count, cudaMemcpyDeviceToHost);
	cudaFree(a); cudaFree(b); cudaFree(o); cudaFree(t);
	if (err != cudaSuccess)
		return fail(JSDE_E_CUDA, cudaGetErrorString(err));
	return JSDE_OK;
}

int jsde_debug_inline_math(int device, const double *a_host, const double *b_host, double *div_host, unsigned char *div_ok_host,
	double *sqrt_host, unsigned char *sqrt_ok_host, int64_t count)
{
	if (!a_host || !b_host || !div_host || !div_ok_host || !sqrt_host || !sqrt_ok_host || count < 0)
		return fail(JSDE_E_ARG, "Wrong input.");
	CU(cudaSetDevice(device));
	double *a = nullptr, *b = nullptr, *od = nullptr, *os = nullptr;
	unsigned char *td = nullptr, *tsq = nullptr;
	CU(cudaMalloc(&a, count * sizeof(double) + 16));
	CU(cudaMalloc(&b, count * sizeof(double) + 16));
	CU(cudaMalloc(&od, count * sizeof(double) + 16));
	CU(cudaMalloc(&os, count * sizeof(double) + 16));
	CU(cudaMalloc(&td, count + 16));
	CU(cudaMalloc(&tsq, count + 16));
	cudaError_t err = cudaMemcpy(a, a_host, count * sizeof(double), cudaMemcpyHostToDevice);
	if (err == cudaSuccess) err = cudaMemcpy(b, b_host, count * sizeof(double), cudaMemcpyHostToDevice);
	if (err == cudaSuccess) {
		k_debug_inline_math<<<grid_for(count, 256), 256>>>(a, b, od, td, os, tsq, count);
		err = cudaGetLastError();
	}
	if (err == cudaSuccess) err = cudaMemcpy(div_host, od, count * sizeof(double), cudaMemcpyDeviceToHost);
	if (err == cudaSuccess) err = cudaMemcpy(sqrt_host, os, count * sizeof(double), cudaMemcpyDeviceToHost);
	if (err == cudaSuccess) err = cudaMemcpy(div_ok_host, td, count, cudaMemcpyDeviceToHost);
	if (err == cudaSuccess) err = cudaMemcpy(sqrt_ok_host, tsq, count, cudaMemcpyDeviceToHost);
	cudaFree(a); cudaFree(b); cudaFree(od); cudaFree(os); cudaFree(td); cudaFree(tsq);
	if (err != cudaSuccess)
		return fail(JSDE_E_CUDA, cudaGetErrorString(err));
	return JSDE_OK;
}


int jsde_debug_log(int device, const double *x_host, double *out_host, int64_t count) { return debug_unary(device, x_host, out_host, count, 0, 0.0); }
int jsde_debug_pow(int device, const double *x_host, double y, double *out_host, int64_t count) { return debug_unary(device, x_host, out_host, count, 1, y); }

}  // extern "C"
