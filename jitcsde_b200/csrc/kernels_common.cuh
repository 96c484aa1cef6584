// Model-independent kernels shared by the lane-per-trajectory and the wide ABI: layout transposes, fills, ensemble
// moments, the FP64 peak probe and the device self-tests of the arithmetic restatements.
#pragma once

#include "ensemble.cuh"
#include "exact_div.cuh"
#include "glibc_math.h"
#include "mt19937.cuh"
#include "philox_jumps.h"

namespace jsde {

// ---- model-independent helpers ------------------------------------------------------------------------------------
__global__ void k_seed(uint4 *mt, int *mt_chunk, const uint32_t *seeds, int64_t B)
{
	const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= B)
		return;
	mt_seed(mt + k * MT_CHUNKS, seeds[k]);
	mt_chunk[k] = 0;
}

// rows [B][n] (trajectory-major, host layout)  <->  cols [n][B] (device layout)
__global__ void k_rows_to_cols(const double *rows, double *cols, int64_t B, int n, int accumulate)
{
	const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= B * n)
		return;
	const int64_t k = idx / n;
	const int i = (int)(idx - k * n);
	if (accumulate)
		cols[(int64_t)i * B + k] += rows[idx];
	else
		cols[(int64_t)i * B + k] = rows[idx];
}

__global__ void k_cols_to_rows(const double *cols, double *rows, int64_t B, int n)
{
	const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= B * n)
		return;
	const int64_t k = idx / n;
	const int i = (int)(idx - k * n);
	rows[idx] = cols[(int64_t)i * B + k];
}

// lane recycling (kernels.cuh): the trajectories beyond the grid's first wave, in ascending order
__global__ void k_pool_init(int *pool, int first, int count, CallTotals *totals)
{
	const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx < count)
		pool[idx] = first + (int)idx;
	if (idx == 0) {
		totals->waiting = count;
		totals->cursor = 0;
	}
}

__global__ void k_fill(double *dst, double value, int64_t count)
{
	const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx < count)
		dst[idx] = value;
}

// acc[0] += #ok, acc[1+i] += sum (y_i - shift_i), acc[1+n+i] += sum (y_i - shift_i)^2 over trajectories with status OK
// y(i, k) = y[i * stride_i + k * stride_k]: [n][B] for the lane kernels (stride_i = B, stride_k = 1), [B][n] for the wide ones
__global__ void k_moments(const double *y, const int *status, int64_t B, int n, const double *shift, double *acc,
	int64_t stride_i, int64_t stride_k)
{
	extern __shared__ double sh[];  // [blockDim.x / 32][1 + 2n]
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
	const int width = 1 + 2 * n;
	for (int f = 0; f < width; f++) {
		double part = 0.0;
		for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < B; k += (int64_t)gridDim.x * blockDim.x) {
			if (status[k] != TRAJ_OK)
				continue;
			if (f == 0)
				part += 1.0;
			else {
				const int i = (f - 1) % n;
				const double d = y[(int64_t)i * stride_i + k * stride_k] - shift[i];
				part += (f <= n) ? d : d * d;
			}
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
			part += __shfl_xor_sync(0xffffffffu, part, o);
		if (lane == 0)
			sh[warp * width + f] = part;
	}
	__syncthreads();
	for (int f = threadIdx.x; f < width; f += blockDim.x) {
		double s = 0.0;
		for (int w = 0; w < nw; w++)
			s += sh[w * width + f];
		atomicAdd(&acc[f], s);
	}
}

// The same accumulators for state vectors stored [B][n] with large n (wide models): thread = component (coalesced over
// i), the block's share of the trajectories in a loop, one atomic per component and block — no shared memory that grows
// with n (the kernel above needs (1+2n) doubles per warp: 128 KB at n = 1000).  grid = (ceil(n/blockDim), chunks).
__global__ void k_moments_wide(const double *y, const int *status, int64_t B, int n, const double *shift, double *acc)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int64_t per = (B + gridDim.y - 1) / gridDim.y;
	const int64_t k0 = (int64_t)blockIdx.y * per, k1 = k0 + per < B ? k0 + per : B;
	double s1 = 0.0, s2 = 0.0, cnt = 0.0;
	const double c = i < n ? shift[i] : 0.0;
	for (int64_t k = k0; k < k1; k++) {
		if (status[k] != TRAJ_OK)
			continue;
		cnt += 1.0;
		if (i < n) {
			const double d = y[k * n + i] - c;
			s1 += d;
			s2 += d * d;
		}
	}
	if (i < n) {
		atomicAdd(&acc[1 + i], s1);
		atomicAdd(&acc[1 + n + i], s2);
	}
	if (i == 0)
		atomicAdd(&acc[0], cnt);
}

// the counter-based jump variates as a tape: taus, xis [count][length] for trajectories first .. first+count-1
__global__ void k_jump_tape(unsigned long long seed, long long first, long long count, long long length, double *taus, double *xis)
{
	const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= count * length)
		return;
	const long long k = idx / length, j = idx - k * length;
	taus[idx] = jsde_jump_tau(seed, first + k, j);
	xis[idx] = jsde_jump_xi(seed, first + k, j);
}

// FP64 peak: 8 independent dependent-chains of DFMA per thread
__global__ void k_dfma_peak(double *sink, int iters)
{
	double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
	const double m = 0.999999, b = 1e-7;
	for (int i = 0; i < iters; i++) {
		a0 = __fma_rn(a0, m, b); a1 = __fma_rn(a1, m, b); a2 = __fma_rn(a2, m, b); a3 = __fma_rn(a3, m, b);
		a4 = __fma_rn(a4, m, b); a5 = __fma_rn(a5, m, b); a6 = __fma_rn(a6, m, b); a7 = __fma_rn(a7, m, b);
	}
	const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
	if (s == 123.456)
		sink[0] = s;
}

__global__ void k_debug_log(const double *x, double *out, int64_t count)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count)
		out[i] = jsde_log(x[i]);
}

// out[i] = a[i]/b[i] through the shared-divisor path (exact_div.cuh); tame[i] tells whether that path vouches for it
__global__ void k_debug_shared_div(const double *a, const double *b, double *out, unsigned char *tame, int64_t count)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count) {
		bool ok = true;
		const SharedDivisor d(b[i]);
		out[i] = d.divide(a[i], ok);
		tame[i] = ok ? 1 : 0;
	}
}

// out_div[i] = a[i]/b[i] and out_sqrt[i] = sqrt(a[i]) through the straight-line restatements of exact_div.cuh, with their
// validity flags
__global__ void k_debug_inline_math(const double *a, const double *b, double *out_div, unsigned char *ok_div, double *out_sqrt,
	unsigned char *ok_sqrt, int64_t count)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count) {
		bool okd = true, oks = true;
		out_div[i] = inline_div(a[i], b[i], okd);
		out_sqrt[i] = inline_sqrt(a[i], oks);
		ok_div[i] = okd ? 1 : 0;
		ok_sqrt[i] = oks ? 1 : 0;
	}
}

__global__ void k_debug_exp(const double *x, double *out, int64_t count)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count)
		out[i] = jsde_exp(x[i]);
}

__global__ void k_debug_pow(const double *x, double yexp, double *out, int64_t count)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count)
		out[i] = jsde_pow(x[i], yexp);
}


}  // namespace jsde
