// ubench_hbm.cu -- what HBM bandwidth does a B200 deliver for k_rank's access pattern: contiguous
// segments of S bytes at random 16-byte-aligned offsets of a 532 MB array (the posting rows of
// config 4 are 8.6 KB on average), read by one warp each with 128-bit loads, V loads per lane
// in flight?  Compared with a sequential sweep of the same array by the same code.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_hbm ubench_hbm.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int V, bool NOALLOC>
__global__ void k(const uint4 *buf, uint64_t n_vec, uint32_t seg_vec, uint32_t segs_per_warp, int sequential, uint32_t *out)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
	uint32_t s = gw * 2654435761u + 777u;
	uint32_t acc = 0;
	for (uint32_t it = 0; it < segs_per_warp; ++it) {
		uint64_t start;
		if (sequential)
			start = ((uint64_t)it * n_warps + gw) * seg_vec % (n_vec - seg_vec);
		else {
			s = s * 1664525u + 1013904223u;
			uint32_t s2 = s * 22695477u + 1u;
			start = (((uint64_t)(s >> 8) << 24) | (s2 >> 8)) % (n_vec - seg_vec);
		}
		const uint4 *p = buf + start;
		for (uint32_t i = lane; i < seg_vec; i += 32 * V) {
			uint4 x[V];
#pragma unroll
			for (int j = 0; j < V; ++j)
				if (i + 32 * j < seg_vec) {
					if (NOALLOC)
						asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
						             : "=r"(x[j].x), "=r"(x[j].y), "=r"(x[j].z), "=r"(x[j].w)
						             : "l"(p + i + 32 * j));
					else
						x[j] = __ldg(p + i + 32 * j);
				} else
					x[j] = make_uint4(0, 0, 0, 0);
#pragma unroll
			for (int j = 0; j < V; ++j)
				acc ^= x[j].x ^ x[j].y ^ x[j].z ^ x[j].w;
		}
	}
	if (acc == 0x12345678)
		out[0] = acc;
}

template <int V, bool NOALLOC>
void run(const uint4 *buf, uint64_t n_vec, uint32_t seg_bytes, int threads, int ctas_per_sm, int sequential, uint32_t *out)
{
	const uint32_t seg_vec = seg_bytes / 16;
	const int grid = 148 * ctas_per_sm;
	const uint64_t warps = (uint64_t)grid * threads / 32;
	const double target_bytes = 40e9;
	uint32_t segs_per_warp = (uint32_t)(target_bytes / ((double)warps * seg_bytes));
	if (segs_per_warp < 4)
		segs_per_warp = 4;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	k<V, NOALLOC><<<grid, threads>>>(buf, n_vec, seg_vec, segs_per_warp / 8 + 1, sequential, out);
	cudaEventRecord(e0);
	k<V, NOALLOC><<<grid, threads>>>(buf, n_vec, seg_vec, segs_per_warp, sequential, out);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	const double bytes = (double)warps * segs_per_warp * seg_bytes;
	printf("%s seg=%7u B  V=%d %s threads=%4d x%d  %8.2f ms  %7.1f GB/s  (%s)\n", sequential ? "sequential" : "random    ", seg_bytes, V,
	  NOALLOC ? "noalloc" : "ldg    ", threads, ctas_per_sm, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
	const uint64_t bytes = 532ull << 20;
	uint4 *buf;
	uint32_t *out;
	cudaMalloc(&buf, bytes);
	cudaMalloc(&out, 4);
	cudaMemset(buf, 1, bytes);
	const uint64_t n_vec = bytes / 16;
	run<4, false>(buf, n_vec, 1 << 20, 1024, 1, 1, out);
	run<4, false>(buf, n_vec, 1 << 20, 1024, 2, 1, out);
	const uint32_t segs[] = {512, 2048, 8192, 8608, 32768, 131072};
	for (uint32_t sb : segs) {
		run<4, false>(buf, n_vec, sb, 1024, 1, 0, out);
		run<4, false>(buf, n_vec, sb, 1024, 2, 0, out);
		run<8, false>(buf, n_vec, sb, 1024, 1, 0, out);
		run<4, true>(buf, n_vec, sb, 1024, 2, 0, out);
	}
	// carve-out effect: prefer max shared memory (small L1)
	cudaFuncSetAttribute(k<4, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
	cudaFuncSetAttribute(k<4, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
	run<4, false>(buf, n_vec, 8608, 1024, 2, 0, out);
	run<4, true>(buf, n_vec, 8608, 1024, 2, 0, out);
	return 0;
}
