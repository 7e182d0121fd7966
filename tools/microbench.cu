// microbench.cu -- measured ceilings for the access patterns of the count/correct path on this GPU:
// random gathers / atomics over working sets from L2-resident to Bloom-sized (16 GiB) and beyond.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench tools/microbench.cu
// Output: one line per (pattern, working set): Gop/s and GB/s of useful bytes.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint64_t mix(uint64_t x)
{
	x += 0x9E3779B97F4A7C15ULL;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
	return x ^ (x >> 31);
}

// pattern 0: 4 random 32-bit words inside one random 64-byte block (Bloom probe, H=4)
// pattern 1: one random 32-byte sector as 2 x 16-byte loads (table bucket)
// pattern 2: one random 32-bit atomicOr (Bloom bit set)
// pattern 3: one random 64-bit atomicCAS (table slot update)
// pattern 4: one random 8-byte load
// pattern 5: one full random 64-byte block as 4 x 16-byte loads
template <int PAT>
__global__ void k_rand(uint32_t *buf, uint64_t n_blocks64, uint64_t n_ops, uint64_t seed, unsigned long long *sink)
{
	uint64_t acc = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_ops; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t h = mix(seed + i);
		const uint64_t blk = h % n_blocks64;
		uint32_t *w = buf + (blk << 4);
		if (PAT == 0) {
			const uint32_t a = (h >> 40) & 15, b = (h >> 44) & 15, c = (h >> 48) & 15, d = (h >> 52) & 15;
			acc += w[a] + w[b] + w[c] + w[d];
		} else if (PAT == 1) {
			const uint4 *q = (const uint4*)(w + ((h >> 40) & 1) * 8);
			const uint4 u = __ldg(q), v = __ldg(q + 1);
			acc += u.x + u.w + v.y + v.z;
		} else if (PAT == 2) {
			atomicOr(w + ((h >> 40) & 15), 1u << ((h >> 50) & 31));
		} else if (PAT == 3) {
			unsigned long long *s = (unsigned long long*)w + ((h >> 40) & 7);
			const unsigned long long old = __ldcg(s);
			acc += atomicCAS(s, old, old + 1);
		} else if (PAT == 4) {
			acc += __ldg((const unsigned long long*)w + ((h >> 40) & 7));
		} else {
			const uint4 *q = (const uint4*)w;
			const uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
			acc += a.x + b.y + c.z + d.w;
		}
	}
	if (acc == 0x123456789ULL) *sink = acc;
}

template <int PAT>
static void run(const char *name, int useful_bytes, uint32_t *buf, uint64_t bytes, uint64_t n_ops, unsigned long long *sink, int sm)
{
	cudaEvent_t a, b;
	cudaEventCreate(&a); cudaEventCreate(&b);
	for (int threads = 256; threads <= 256; threads *= 2) {
		const int grid = sm * (2048 / threads);
		k_rand<PAT><<<grid, threads>>>(buf, bytes >> 6, n_ops / 8, 1, sink);
		float best = 1e30f;
		for (int rep = 0; rep < 3; ++rep) {
			cudaEventRecord(a);
			k_rand<PAT><<<grid, threads>>>(buf, bytes >> 6, n_ops, 7 + rep, sink);
			cudaEventRecord(b);
			cudaEventSynchronize(b);
			float ms;
			cudaEventElapsedTime(&ms, a, b);
			if (ms < best) best = ms;
		}
		printf("%-28s ws=%8.2f GiB  %7.2f Gop/s  %8.1f GB/s useful  (%.2f ms)\n", name, bytes / 1073741824.0,
		       n_ops / best / 1e6, n_ops * (double)useful_bytes / best / 1e6, best);
	}
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
}

int main(int argc, char **argv)
{
	const uint64_t max_bytes = (argc > 1 ? strtoull(argv[1], 0, 10) : 64ULL) << 30;
	uint32_t *buf;
	unsigned long long *sink;
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, 0);
	if (argc > 2) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, atoi(argv[2]));
	size_t gran = 0;
	cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity);
	printf("device: %s, %d SMs, L2 fetch granularity limit %zu\n", prop.name, prop.multiProcessorCount, gran);
	if (cudaMalloc(&buf, max_bytes) != cudaSuccess) { printf("cudaMalloc failed\n"); return 1; }
	cudaMalloc(&sink, 8);
	cudaMemset(buf, 0, max_bytes);
	const uint64_t n_ops = 1ULL << 28;
	for (uint64_t bytes = 64ULL << 20; bytes <= max_bytes; bytes <<= 2) {
		run<0>("bloom probe 4 words/64B", 64, buf, bytes, n_ops, sink, prop.multiProcessorCount);
		run<5>("full 64B block 4xLDG.128", 64, buf, bytes, n_ops, sink, prop.multiProcessorCount);
		run<1>("table bucket 32B sector", 32, buf, bytes, n_ops, sink, prop.multiProcessorCount);
		run<4>("single 8B load", 8, buf, bytes, n_ops, sink, prop.multiProcessorCount);
		run<2>("atomicOr 32-bit", 4, buf, bytes, n_ops, sink, prop.multiProcessorCount);
		run<3>("ld + atomicCAS 64-bit", 8, buf, bytes, n_ops, sink, prop.multiProcessorCount);
	}
	return 0;
}
