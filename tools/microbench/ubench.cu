// ubench.cu -- B200 micro-measurements that drive the PCG / J+H kernel design (clock64 based).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ubench ubench.cu && ./ubench
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_fp64(double* out, long long* cyc, int mode, int iters)
{
	double a = threadIdx.x * 1e-3 + 1.0, b = 1.000001, c = 0.5, d = 0.25, e = 0.125, f = 0.0625, g = 2.0, h = 3.0;
	__syncthreads();
	const long long t0 = clock64();
	if (mode == 0) { for (int i = 0; i < iters; i++) { a = fma(a, b, c); } }                       // dependent DFMA chain -> latency
	else if (mode == 1) { for (int i = 0; i < iters; i++) { a = fma(a, b, c); d = fma(d, b, c); e = fma(e, b, c); f = fma(f, b, c); g = fma(g, b, c); h = fma(h, b, c); } }
	else if (mode == 2) { for (int i = 0; i < iters; i++) { a = a + b; } }                           // dependent DADD
	else if (mode == 3) { for (int i = 0; i < iters; i++) { a = b / (a + 1.5); } }                    // dependent DDIV
	else if (mode == 4) { for (int i = 0; i < iters; i++) { a = sqrt(a + 1.5); } }
	else if (mode == 5) { for (int i = 0; i < iters; i++) { a += __shfl_xor_sync(0xffffffffu, a, 1); } }
	else if (mode == 6) { float x = (float)a, y = 1.000001f, z = 0.5f; for (int i = 0; i < iters; i++) { x = fmaf(x, y, z); } a = x; }
	const long long t1 = clock64();
	__syncthreads();
	out[blockIdx.x * blockDim.x + threadIdx.x] = a + d + e + f + g + h;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_sync(long long* cyc, int iters)
{
	__syncthreads();
	const long long t0 = clock64();
	for (int i = 0; i < iters; i++) __syncthreads();
	const long long t1 = clock64();
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ping-pong between CTA 0 and CTA 1 through L2 with tagged 16-byte words (the k_pcg3 protocol)
__device__ __forceinline__ void ll_store(unsigned long long* slot, unsigned long long lo, unsigned long long hi)
{
	asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" :: "l"(slot), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ void ll_load(const unsigned long long* slot, unsigned long long& lo, unsigned long long& hi)
{
	asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(slot) : "memory");
}
__global__ void k_pingpong(unsigned long long* buf, long long* cyc, int iters, int other)
{
	// CTA 0 and CTA `other` bounce a tag; all other CTAs idle
	if (blockIdx.x != 0 && blockIdx.x != other) return;
	if (threadIdx.x != 0) return;
	const bool first = blockIdx.x == 0;
	unsigned long long* mine = buf + (first ? 0 : 32);
	unsigned long long* theirs = buf + (first ? 32 : 0);
	const long long t0 = clock64();
	for (unsigned long long i = 1; i <= (unsigned long long)iters; i++) {
		if (first) ll_store(theirs, i, i);
		unsigned long long lo = 0, hi = 0;
		do { ll_load(mine, lo, hi); } while (lo != i || hi != i);
		if (!first) ll_store(theirs, i, i);
	}
	const long long t1 = clock64();
	cyc[blockIdx.x] = t1 - t0;
}

// all-to-all: every CTA publishes one tagged value per round and waits for everybody's (the partial-product exchange)
__global__ void k_allgather(unsigned long long* buf, long long* cyc, int iters)
{
	const int G = gridDim.x;
	__shared__ int dummy;
	const long long t0 = clock64();
	for (unsigned long long i = 1; i <= (unsigned long long)iters; i++) {
		unsigned long long* base = buf + (i & 1) * 2 * (size_t)G;
		if (threadIdx.x == 0) ll_store(base + 2 * blockIdx.x, i, i);
		if (threadIdx.x < G) {
			unsigned long long lo = 0, hi = 0;
			do { ll_load(base + 2 * threadIdx.x, lo, hi); } while (lo != i || hi != i);
		}
		__syncthreads();
	}
	const long long t1 = clock64();
	if (threadIdx.x == 0) { cyc[blockIdx.x] = t1 - t0; dummy = 0; }
}

int main()
{
	double* out; long long* cyc; unsigned long long* buf;
	cudaMalloc(&out, sizeof(double) * 1024 * 256); cudaMalloc(&cyc, sizeof(long long) * 1024); cudaMalloc(&buf, 1 << 20);
	long long h[1024];
	const char* names[] = { "DFMA dependent chain", "DFMA 6 chains", "DADD dependent", "DDIV dependent", "DSQRT dependent", "SHFL+DADD dependent", "FFMA dependent" };
	const int iters = 2000;
	for (int mode = 0; mode < 7; mode++) {
		for (int threads : { 32, 128, 512, 1024 }) {
			k_fp64<<<1, threads>>>(out, cyc, mode, iters);
			cudaMemcpy(h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
			const int ops = mode == 1 ? 6 : 1;
			printf("%-22s %4d threads: %7.2f cycles/iter  (%6.2f warp-instr/clk/SM)\n", names[mode], threads, (double)h[0] / iters,
				(double)ops * iters * (threads / 32) / (double)h[0]);
		}
	}
	for (int threads : { 128, 512, 1024 }) {
		k_sync<<<1, threads>>>(cyc, 1000);
		cudaMemcpy(h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
		printf("__syncthreads %4d threads: %.1f cycles\n", threads, (double)h[0] / 1000);
	}
	int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
	for (int other : { 1, 2, 37, 74, 100, 147 }) {
		cudaMemset(buf, 0, 1 << 20);
		k_pingpong<<<148, 32>>>(buf, cyc, 2000, other);
		cudaMemcpy(h, cyc, sizeof(long long) * 148, cudaMemcpyDeviceToHost);
		printf("ping-pong CTA0<->CTA%-3d: %.0f cycles per round trip (%.2f us at %d MHz)\n", other, (double)h[0] / 2000, (double)h[0] / 2000 / (clk / 1e3), clk / 1000);
	}
	for (int G : { 2, 16, 31, 148 }) {
		cudaMemset(buf, 0, 1 << 20);
		void* args[] = { &buf, &cyc, (void*)&iters };
		int it2 = 2000; args[2] = &it2;
		cudaLaunchCooperativeKernel((void*)k_allgather, dim3(G), dim3(256), args, 0, 0);
		cudaError_t e = cudaDeviceSynchronize();
		cudaMemcpy(h, cyc, sizeof(long long) * G, cudaMemcpyDeviceToHost);
		printf("all-gather of one tagged word, %3d CTAs: %.0f cycles per round (%s)\n", G, (double)h[0] / 2000, cudaGetErrorString(e));
	}
	return 0;
}
