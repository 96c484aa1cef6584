// Microbenchmark: sustained FP64 rate of mma.sync.aligned.m8n8k4.f64 (DMMA) vs a DFMA chain on one B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak dmma_peak.cu ; run without arguments.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&d)[2], double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void k_dmma(double *out, int iters, double a, double b)
{
	double acc[CHAINS][2];
	for (int c = 0; c < CHAINS; c++)
		acc[c][0] = acc[c][1] = threadIdx.x + c;
	for (int i = 0; i < iters; i++)
#pragma unroll
		for (int c = 0; c < CHAINS; c++)
			dmma(acc[c], a, b);
	double s = 0;
	for (int c = 0; c < CHAINS; c++)
		s += acc[c][0] + acc[c][1];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void k_dfma(double *out, int iters, double a, double b)
{
	double acc[CHAINS];
	for (int c = 0; c < CHAINS; c++)
		acc[c] = threadIdx.x + c;
	for (int i = 0; i < iters; i++)
#pragma unroll
		for (int c = 0; c < CHAINS; c++)
			acc[c] = fma(acc[c], a, b);
	double s = 0;
	for (int c = 0; c < CHAINS; c++)
		s += acc[c];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_ms(F launch)
{
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	launch();
	cudaDeviceSynchronize();
	cudaEventRecord(e0);
	launch();
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	return ms;
}

int main()
{
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	const int blocks = p.multiProcessorCount * 4, threads = 256, iters = 1 << 14;
	double *out;
	cudaMalloc(&out, sizeof(double) * blocks * threads);
	constexpr int C = 8;
	float ms = time_ms([&] { k_dmma<C><<<blocks, threads>>>(out, iters, 1.0000001, 0.9999999); });
	double flop = 2.0 * 8 * 8 * 4 * C * (double)iters * (threads / 32) * blocks;
	printf("DMMA m8n8k4: %.2f TFLOP/s (%d blocks x %d threads, %d chains)\n", flop / ms * 1e-9, blocks, threads, C);
	ms = time_ms([&] { k_dfma<C><<<blocks, threads>>>(out, iters, 1.0000001, 0.9999999); });
	flop = 2.0 * C * (double)iters * threads * blocks;
	printf("DFMA       : %.2f TFLOP/s\n", flop / ms * 1e-9);
	printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
	return 0;
}
