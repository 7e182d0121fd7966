// synth.cu -- device-side synthetic read generator for HBM-resident benchmarks.
// Not part of the reference's path: it only produces input (SURVEY.md 8(d) recipe:
// uniform random genome, uniform start/strand, substitution errors with low quality,
// rare N).  A counter-based generator (splitmix64 of seed and index), so the same
// reads can be regenerated anywhere; bfc_b200/synth.py holds the identical numpy
// formulation used by the CPU reference arm of bench.py.
#include "common.cuh"

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
	x += 0x9E3779B97F4A7C15ULL;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
	return x ^ (x >> 31);
}

__global__ void k_synth_genome(uint8_t *g, uint64_t G, uint64_t seed)
{
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < G; i += (uint64_t)gridDim.x * blockDim.x)
		g[i] = (uint8_t)(splitmix64(seed * 0x632BE59BD9B4E019ULL + i) >> 62);
}

struct SynthParams {
	const uint8_t *genome;
	uint64_t G, seed, first_read;
	int64_t n_reads;
	int L;
	uint32_t err16, nrate20; // error rate * 2^16, N rate * 2^20
	uint8_t *seq, *qual;
	uint64_t *off;
};

__global__ void k_synth_reads(SynthParams p)
{
	const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r > p.n_reads) return;
	const uint64_t o = (uint64_t)r * (p.L + 1);
	p.off[r] = o;
	if (r == p.n_reads) return;
	const uint64_t h = splitmix64(p.seed ^ ((p.first_read + (uint64_t)r) * 0xD6E8FEB86659FD93ULL));
	const uint64_t start = h % (p.G - p.L + 1);
	const int strand = (int)(splitmix64(h) & 1);
	for (int j = 0; j < p.L; ++j) {
		int c = strand ? 3 - p.genome[start + p.L - 1 - j] : p.genome[start + j];
		const uint64_t u = splitmix64(h ^ ((uint64_t)(j + 1) * 0xA24BAED4963EE407ULL));
		const bool is_err = (uint32_t)(u & 0xFFFF) < p.err16;
		if (is_err) c = (c + 1 + (int)((u >> 16 & 0xFF) % 3)) & 3;
		const int q = is_err ? 2 + (int)((u >> 28 & 0xFF) % 18) : 25 + (int)(u >> 24 & 15);
		const bool is_n = (uint32_t)(u >> 40 & 0xFFFFF) < p.nrate20;
		p.seq[o + j] = is_n ? 'N' : "ACGT"[c];
		p.qual[o + j] = (uint8_t)(33 + q);
	}
	p.seq[o + p.L] = 0, p.qual[o + p.L] = 0;
}

extern "C" int bfcg_synth_genome(uint8_t *d_genome, uint64_t G, uint64_t seed)
{
	int r;
	if ((r = bfcg_rt_init()) != BFCG_OK) return r;
	BfcgRuntime &rt = bfcg_rt();
	k_synth_genome<<<rt.sm_count * 16, 256, 0, rt.stream>>>(d_genome, G, seed);
	BFCG_LAUNCH_CHECK();
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	return BFCG_OK;
}

// seq/qual: n_reads * (L + 1) bytes each; off: n_reads + 1 entries (all device memory)
extern "C" int bfcg_synth_reads(const uint8_t *d_genome, uint64_t G, uint64_t seed, uint64_t first_read, int64_t n_reads, int L,
                                double err, double n_rate, uint8_t *d_seq, uint8_t *d_qual, uint64_t *d_off)
{
	int r;
	if ((r = bfcg_rt_init()) != BFCG_OK) return r;
	BfcgRuntime &rt = bfcg_rt();
	SynthParams p;
	p.genome = d_genome, p.G = G, p.seed = seed, p.first_read = first_read, p.n_reads = n_reads, p.L = L;
	p.err16 = (uint32_t)(err * 65536.0 + 0.5), p.nrate20 = (uint32_t)(n_rate * 1048576.0 + 0.5);
	p.seq = d_seq, p.qual = d_qual, p.off = d_off;
	k_synth_reads<<<(unsigned)((n_reads + 1 + 255) / 256), 256, 0, rt.stream>>>(p);
	BFCG_LAUNCH_CHECK();
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	return BFCG_OK;
}
