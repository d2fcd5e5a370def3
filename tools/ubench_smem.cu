// ubench_smem.cu -- what does one random shared-memory counter increment cost on sm_100a?
// (decides how k_rank counts postings).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_smem ubench_smem.cu
//   mode 0  atomicAdd on packed u8 counters, random targets (k_rank today)
//   mode 1  atomicAdd, conflict-free banks (bank == lane)
//   mode 2  LDS.U8 + STS.U8 (non atomic), random targets
//   mode 3  atomicAdd random, u16-packed
//   mode 4  LDS.32 + STS.32 random (non atomic RMW on the word)
//   mode 5  STS.U8 only (random)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define NT 100000u

template <int MODE> __global__ void k(uint32_t iters, uint32_t *out)
{
	extern __shared__ __align__(16) uint8_t U[];
	uint32_t *U32 = (uint32_t *)U;
	for (uint32_t i = threadIdx.x; i < (NT + 3) / 4; i += blockDim.x)
		U32[i] = 0;
	__syncthreads();
	uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
	const uint32_t lane = threadIdx.x & 31;
	// 16 fixed targets per thread, computed once: the loop below is nothing but the memory ops
	uint32_t t[16];
#pragma unroll
	for (int r = 0; r < 16; ++r) {
		s = s * 1664525u + 1013904223u;
		t[r] = (uint32_t)(((uint64_t)(s >> 4) * NT) >> 28);
		if (MODE == 1) // conflict-free: bank == lane
			t[r] = ((t[r] >> 7) << 7) | (lane << 2) | (t[r] & 3);
		if (MODE == 6) // 2-way conflicts: bank == lane / 2
			t[r] = ((t[r] >> 7) << 7) | ((lane >> 1) << 2) | (t[r] & 3) | ((lane & 1) << 13);
	}
	for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
		for (int r = 0; r < 16; ++r) {
			const uint32_t tt = t[r];
			if (MODE == 0 || MODE == 1 || MODE == 6)
				atomicAdd(&U32[tt >> 2], 1u << ((tt & 3) * 8));
			else if (MODE == 2)
				((volatile uint8_t *)U)[tt] = ((volatile uint8_t *)U)[tt] + 1;
			else if (MODE == 3)
				atomicAdd(&U32[tt >> 2], 1u << ((tt & 1) * 16));
			else if (MODE == 4)
				((volatile uint32_t *)U32)[tt >> 2] = ((volatile uint32_t *)U32)[tt >> 2] + (1u << ((tt & 3) * 8));
			else if (MODE == 5)
				((volatile uint8_t *)U)[tt] = (uint8_t)it;
		}
	}
	__syncthreads();
	uint32_t acc = 0;
	for (uint32_t i = threadIdx.x; i < (NT + 3) / 4; i += blockDim.x)
		acc += U32[i];
	if (acc == 0xdeadbeef)
		out[0] = acc;
}

template <int MODE> void run(const char *name, int threads, int ctas_per_sm, uint32_t iters)
{
	uint32_t *out;
	cudaMalloc(&out, 4);
	size_t smem = NT + 16;
	cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	int grid = 148 * ctas_per_sm;
	k<MODE><<<grid, threads, smem>>>(iters / 8, out);
	cudaEventRecord(e0);
	k<MODE><<<grid, threads, smem>>>(iters, out);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	double warp_inst = (double)grid * (threads / 32) * iters * 16.0;
	double per_sm = warp_inst / 148.0;
	int clk;
	cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
	printf("%-28s threads=%4d cta/sm=%d  %8.3f ms  %.2f cycles/warp-op/SM (at %d MHz)  %.1f G lane-ops/s  err=%s\n", name, threads,
	  ctas_per_sm, ms, ms * 1e-3 * clk * 1e3 / per_sm, clk / 1000, warp_inst * 32 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
	cudaFree(out);
}

int main()
{
	const uint32_t it = 8000;
	run<0>("atomicAdd u8x4 random", 1024, 1, it);
	run<0>("atomicAdd u8x4 random", 512, 2, it);
	run<0>("atomicAdd u8x4 random", 512, 1, it);
	run<0>("atomicAdd u8x4 random", 256, 1, it);
	run<1>("atomicAdd conflict-free", 1024, 1, it);
	run<2>("LDS.U8+STS.U8 random", 1024, 1, it);
	run<3>("atomicAdd u16x2 random", 1024, 1, it);
	run<4>("LDS.32+STS.32 random", 1024, 1, it);
	run<5>("STS.U8 random", 1024, 1, it);
	run<6>("atomicAdd 2-way conflicts", 1024, 1, it);
	return 0;
}
