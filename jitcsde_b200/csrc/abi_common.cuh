// Host-side plumbing shared by the two extern "C" translation units (jsde_abi.cu: one trajectory per lane;
// jsde_abi_wide.cu: large state vectors): error reporting, the part of the handle both have, device allocations, the
// controller parameters' validation and clamps, jump-tape upload, call statistics, checkpoint I/O over a list of
// sections, and the model-independent utility entry points (FP64 peak probe, device self-tests, generated jump tape).
// Included exactly once per library, after the model header.
#pragma once

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/jsde.h"
#include "kernels_common.cuh"

namespace jsde {
namespace abi {

inline thread_local std::string g_error;

inline int fail(int code, const std::string &msg)
{
	g_error = msg;
	return code;
}

#define CU(call)                                                                                          \
	do {                                                                                                  \
		cudaError_t err__ = (call);                                                                       \
		if (err__ != cudaSuccess)                                                                         \
			return ::jsde::abi::fail(err__ == cudaErrorMemoryAllocation ? JSDE_E_NOMEM : JSDE_E_CUDA,     \
				std::string(#call) + ": " + cudaGetErrorString(err__));                                   \
	} while (0)

inline unsigned grid_for(int64_t count, int block) { return (unsigned)((count + block - 1) / block); }

// what both kinds of handle carry
struct HandleBase {
	int device = 0;
	int64_t B = 0;
	int D = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	Controller ctrl{};
	double first_step = 1.0;
	uint32_t *seeds = nullptr;
	double *par = nullptr, *shift = nullptr;
	// jump tape on the device: owned buffers, replaced only by an explicit upload (jsde_set_jump_tape or a non-NULL tape
	// argument).  Round 1 cached uploads by the identity of the caller's host pointers, which silently reused a stale
	// tape when a new array landed at a freed address and never saw in-place edits.
	double *taus = nullptr, *xis = nullptr;
	long long tape_len = 0;
	size_t tape_capacity = 0;  // doubles per buffer
	bool tape_set = false;
	// the generator is re-seeded lazily: set_state and seed both restart it (reference _jitcsde.py:242-246, :270-274), and
	// a front end calls them back to back — the 2.6 GB seeding pass runs once, before the first call that draws
	bool seed_dirty = true;
	std::vector<void *> allocations;
};

template <class T>
int dev_alloc(HandleBase *h, T **ptr, size_t count)
{
	void *p = nullptr;
	CU(cudaMalloc(&p, count * sizeof(T) > 0 ? count * sizeof(T) : 16));
	h->allocations.push_back(p);
	*ptr = static_cast<T *>(p);
	return JSDE_OK;
}

inline int bind(HandleBase *h)
{
	CU(cudaSetDevice(h->device));
	return JSDE_OK;
}

inline int check_device(int device)
{
	int count = 0;
	cudaError_t err = cudaGetDeviceCount(&count);
	if (err != cudaSuccess || count == 0)
		return fail(JSDE_E_CUDA, std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(err));
	if (device < 0 || device >= count)
		return fail(JSDE_E_ARG, "Wrong input. (device index out of range)");
	return JSDE_OK;
}

inline int create_stream_and_events(HandleBase *h)
{
	CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
	CU(cudaEventCreate(&h->ev0));
	CU(cudaEventCreate(&h->ev1));
	return JSDE_OK;
}

inline void destroy_base(HandleBase *h)
{
	cudaSetDevice(h->device);
	if (h->stream) cudaStreamSynchronize(h->stream);
	for (void *p : h->allocations) cudaFree(p);
	if (h->taus) cudaFree(h->taus);
	if (h->xis) cudaFree(h->xis);
	if (h->ev0) cudaEventDestroy(h->ev0);
	if (h->ev1) cudaEventDestroy(h->ev1);
	if (h->stream) cudaStreamDestroy(h->stream);
}

// defaults of set_integration_parameters (_jitcsde.py:491-502), q = 1.5 (:564)
inline void default_controller(HandleBase *h)
{
	Controller &c = h->ctrl;
	c.atol = 1e-10; c.rtol = 1e-5; c.min_step = 1e-10; c.max_step = 10.0; c.dec_thr = 1.1; c.inc_thr = 0.5;
	c.safety = 0.9; c.max_factor = 5.0; c.min_factor = 0.2;
	c.shrink_exp = -1 / 1.5;
	c.grow_exp = -1 / (1.5 + 1);
	h->first_step = 1.0;
}

// the assertions of _jitcsde.py:543-549 and the clamps of :536-541; the caller then sets every trajectory's dt
inline int apply_controller(HandleBase *h, const jsde_ctrl *p)
{
	if (!(p->decrease_threshold >= 1.0) || !(p->increase_threshold <= 1.0) || !(p->max_factor >= 1.0) || !(p->min_factor <= 1.0) ||
		!(p->safety_factor <= 1.0) || !(p->atol >= 0.0) || !(p->rtol >= 0.0) || !(p->q > 0.0))
		return fail(JSDE_E_ARG, "Wrong input. (integration parameter out of range)");
	double first = p->first_step, min_step = p->min_step;
	if (first > p->max_step) first = p->max_step;
	if (min_step > first) min_step = first;
	Controller &c = h->ctrl;
	c.atol = p->atol; c.rtol = p->rtol; c.min_step = min_step; c.max_step = p->max_step;
	c.dec_thr = p->decrease_threshold; c.inc_thr = p->increase_threshold; c.safety = p->safety_factor;
	c.max_factor = p->max_factor; c.min_factor = p->min_factor;
	c.shrink_exp = -1 / p->q;
	c.grow_exp = -1 / (p->q + 1);
	h->first_step = first;
	return JSDE_OK;
}

// copy the caller's tape to the device (always: the caller says when the tape changes); buffers are reused when they fit
inline int upload_tape(HandleBase *h, const jsde_jump_tape *tape)
{
	if (!tape || tape->length < 0 || (tape->length > 0 && (!tape->taus_host || !tape->xis_host)))
		return fail(JSDE_E_ARG, "Wrong input.");
	const size_t count = (size_t)h->B * (size_t)tape->length;
	if (count > h->tape_capacity) {
		CU(cudaStreamSynchronize(h->stream));
		if (h->taus) cudaFree(h->taus);
		if (h->xis) cudaFree(h->xis);
		h->taus = h->xis = nullptr;
		h->tape_capacity = 0;
		CU(cudaMalloc(&h->taus, count * sizeof(double)));
		CU(cudaMalloc(&h->xis, count * sizeof(double)));
		h->tape_capacity = count;
	}
	if (count) {
		CU(cudaMemcpyAsync(h->taus, tape->taus_host, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
		CU(cudaMemcpyAsync(h->xis, tape->xis_host, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
		CU(cudaStreamSynchronize(h->stream));  // the caller may free or edit its arrays as soon as the call returns
	}
	h->tape_len = tape->length;
	h->tape_set = true;
	return JSDE_OK;
}

// end of an integrate call: stop the clock, fetch the totals
inline int finish_call(HandleBase *h, const CallTotals *totals_dev, jsde_stats *stats, int launches)
{
	CallTotals tot;
	CU(cudaEventRecord(h->ev1, h->stream));
	CU(cudaMemcpyAsync(&tot, totals_dev, sizeof(tot), cudaMemcpyDeviceToHost, h->stream));
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
		stats->n_tape_exhausted = (long long)tot.tape_exhausted;
	}
	return JSDE_OK;
}

// ---- checkpoint blobs: {header, sections in a fixed order} -------------------------------------------------------------
using Sections = std::vector<std::pair<void *, size_t>>;

inline size_t sections_bytes(const Sections &sections)
{
	size_t total = 0;
	for (auto &s : sections)
		total += s.second;
	return total;
}

inline int save_sections(HandleBase *h, const Sections &sections, char *out)
{
	CU(cudaStreamSynchronize(h->stream));
	for (auto &s : sections) {
		CU(cudaMemcpyAsync(out, s.first, s.second, cudaMemcpyDeviceToHost, h->stream));
		out += s.second;
	}
	CU(cudaStreamSynchronize(h->stream));
	return JSDE_OK;
}

inline int restore_sections(HandleBase *h, const Sections &sections, const char *in)
{
	for (auto &s : sections) {
		CU(cudaMemcpyAsync(s.first, in, s.second, cudaMemcpyHostToDevice, h->stream));
		in += s.second;
	}
	CU(cudaStreamSynchronize(h->stream));
	return JSDE_OK;
}

// ---- model-independent entry points ----------------------------------------------------------------------------------
inline int generated_jump_tape(int device, uint64_t seed, int64_t first_index, int64_t count, int64_t length, double *taus_host, double *xis_host)
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

inline int measure_fp64_peak(int device, double *tflops)
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

// device buffers for the self-tests: freed on every path
struct Scratch {
	std::vector<void *> ptrs;
	~Scratch() { for (void *p : ptrs) cudaFree(p); }
	template <class T>
	cudaError_t get(T **ptr, size_t count)
	{
		void *p = nullptr;
		cudaError_t err = cudaMalloc(&p, count * sizeof(T) + 16);
		if (err == cudaSuccess) ptrs.push_back(p);
		*ptr = static_cast<T *>(p);
		return err;
	}
};

inline int debug_unary(int device, const double *x_host, double *out_host, int64_t count, int which, double yexp)
{
	if (!x_host || !out_host || count < 0)
		return fail(JSDE_E_ARG, "Wrong input.");
	CU(cudaSetDevice(device));
	Scratch s;
	double *x = nullptr, *o = nullptr;
	CU(s.get(&x, count));
	CU(s.get(&o, count));
	CU(cudaMemcpy(x, x_host, count * sizeof(double), cudaMemcpyHostToDevice));
	if (which == 0)
		k_debug_log<<<grid_for(count, 256), 256>>>(x, o, count);
	else if (which == 2)
		k_debug_exp<<<grid_for(count, 256), 256>>>(x, o, count);
	else
		k_debug_pow<<<grid_for(count, 256), 256>>>(x, yexp, o, count);
	CU(cudaGetLastError());
	CU(cudaMemcpy(out_host, o, count * sizeof(double), cudaMemcpyDeviceToHost));
	return JSDE_OK;
}

inline int debug_shared_div(int device, const double *a_host, const double *b_host, double *out_host, unsigned char *tame_host, int64_t count)
{
	if (!a_host || !b_host || !out_host || !tame_host || count < 0)
		return fail(JSDE_E_ARG, "Wrong input.");
	CU(cudaSetDevice(device));
	Scratch s;
	double *a = nullptr, *b = nullptr, *o = nullptr;
	unsigned char *t = nullptr;
	CU(s.get(&a, count)); CU(s.get(&b, count)); CU(s.get(&o, count)); CU(s.get(&t, count));
	CU(cudaMemcpy(a, a_host, count * sizeof(double), cudaMemcpyHostToDevice));
	CU(cudaMemcpy(b, b_host, count * sizeof(double), cudaMemcpyHostToDevice));
	k_debug_shared_div<<<grid_for(count, 256), 256>>>(a, b, o, t, count);
	CU(cudaGetLastError());
	CU(cudaMemcpy(out_host, o, count * sizeof(double), cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(tame_host, t, count, cudaMemcpyDeviceToHost));
	return JSDE_OK;
}

inline int debug_inline_math(int device, const double *a_host, const double *b_host, double *div_host, unsigned char *div_ok_host,
	double *sqrt_host, unsigned char *sqrt_ok_host, int64_t count)
{
	if (!a_host || !b_host || !div_host || !div_ok_host || !sqrt_host || !sqrt_ok_host || count < 0)
		return fail(JSDE_E_ARG, "Wrong input.");
	CU(cudaSetDevice(device));
	Scratch s;
	double *a = nullptr, *b = nullptr, *od = nullptr, *os = nullptr;
	unsigned char *td = nullptr, *tsq = nullptr;
	CU(s.get(&a, count)); CU(s.get(&b, count)); CU(s.get(&od, count)); CU(s.get(&os, count)); CU(s.get(&td, count)); CU(s.get(&tsq, count));
	CU(cudaMemcpy(a, a_host, count * sizeof(double), cudaMemcpyHostToDevice));
	CU(cudaMemcpy(b, b_host, count * sizeof(double), cudaMemcpyHostToDevice));
	k_debug_inline_math<<<grid_for(count, 256), 256>>>(a, b, od, td, os, tsq, count);
	CU(cudaGetLastError());
	CU(cudaMemcpy(div_host, od, count * sizeof(double), cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(sqrt_host, os, count * sizeof(double), cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(div_ok_host, td, count, cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(sqrt_ok_host, tsq, count, cudaMemcpyDeviceToHost));
	return JSDE_OK;
}

}  // namespace abi
}  // namespace jsde

// ---- extern "C" entry points that are the same in both libraries --------------------------------------------------------
#define JSDE_DEFINE_COMMON_ENTRY_POINTS(MODEL)                                                                                          \
	extern "C" {                                                                                                                        \
	const char *jsde_last_error(void) { return ::jsde::abi::g_error.c_str(); }                                                          \
	const char *jsde_model_name(void) { return JSDE_MODEL_NAME; }                                                                       \
	int jsde_model_info(int *n, int *additive, int *n_params, int *has_jump)                                                            \
	{                                                                                                                                   \
		if (n) *n = MODEL::N;                                                                                                           \
		if (additive) *additive = MODEL::ADDITIVE ? 1 : 0;                                                                              \
		if (n_params) *n_params = MODEL::NPAR;                                                                                          \
		if (has_jump) *has_jump = MODEL::HAS_JUMP ? 1 : 0;                                                                              \
		return JSDE_OK;                                                                                                                 \
	}                                                                                                                                   \
	int jsde_generated_jump_tape(int device, uint64_t seed, int64_t first_index, int64_t count, int64_t length, double *taus_host,      \
		double *xis_host)                                                                                                               \
	{                                                                                                                                   \
		return ::jsde::abi::generated_jump_tape(device, seed, first_index, count, length, taus_host, xis_host);                         \
	}                                                                                                                                   \
	int jsde_measure_fp64_peak(int device, double *tflops) { return ::jsde::abi::measure_fp64_peak(device, tflops); }                   \
	int jsde_debug_log(int device, const double *x_host, double *out_host, int64_t count)                                               \
	{                                                                                                                                   \
		return ::jsde::abi::debug_unary(device, x_host, out_host, count, 0, 0.0);                                                       \
	}                                                                                                                                   \
	int jsde_debug_exp(int device, const double *x_host, double *out_host, int64_t count)                                               \
	{                                                                                                                                   \
		return ::jsde::abi::debug_unary(device, x_host, out_host, count, 2, 0.0);                                                       \
	}                                                                                                                                   \
	int jsde_debug_pow(int device, const double *x_host, double y, double *out_host, int64_t count)                                     \
	{                                                                                                                                   \
		return ::jsde::abi::debug_unary(device, x_host, out_host, count, 1, y);                                                         \
	}                                                                                                                                   \
	int jsde_debug_shared_div(int device, const double *a_host, const double *b_host, double *out_host, unsigned char *tame_host,       \
		int64_t count)                                                                                                                  \
	{                                                                                                                                   \
		return ::jsde::abi::debug_shared_div(device, a_host, b_host, out_host, tame_host, count);                                       \
	}                                                                                                                                   \
	int jsde_debug_inline_math(int device, const double *a_host, const double *b_host, double *div_host, unsigned char *div_ok_host,    \
		double *sqrt_host, unsigned char *sqrt_ok_host, int64_t count)                                                                  \
	{                                                                                                                                   \
		return ::jsde::abi::debug_inline_math(device, a_host, b_host, div_host, div_ok_host, sqrt_host, sqrt_ok_host, count);           \
	}                                                                                                                                   \
	}
