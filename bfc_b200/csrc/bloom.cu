// bloom.cu -- the reference's bfc_bf_* entry points (bbf.h:14-17) over a device-resident
// filter.  The single-element insert/get exist for API compatibility (one-thread
// kernels); the throughput path is bfcg_count_batch / bfcg_trim_batch.
#include "common.cuh"
#include <math.h>

__global__ void k_bf_insert1(BloomView bf, uint64_t hash, int *ret)
{
	// sequential test-then-set per probe, exactly bbf.c:35-42
	const BloomProbe p = bloom_locate(hash, bf.n_shift);
	uint32_t *w = bf.w + (p.blk << 4);
	int z = p.h1, done = 0, cnt = 0;
	while (done < bf.n_hashes) {
		if (z >= 8) {
			const uint32_t bit = 1u << (z & 31);
			const uint32_t old = atomicOr(w + (z >> 5), bit);
			cnt += (old & bit) != 0;
			++done;
		}
		z = (z + p.h2) & BFC_BLK_MASK;
	}
	*ret = cnt;
}

__global__ void k_bf_get1(BloomView bf, uint64_t hash, int *ret)
{
	const BloomProbe p = bloom_locate(hash, bf.n_shift);
	*ret = bloom_count_set<true>(bf.w + (p.blk << 4), p, bf.n_hashes);
}

// occupancy: set bits and blocks with at least one set bit (one streaming pass)
__global__ void __launch_bounds__(256) k_bf_load(const uint4 *w, uint64_t n_quads /* 16-byte pieces */, unsigned long long *out)
{
	unsigned long long bits = 0, blocks = 0;
	for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x; i0 < n_quads; i0 += (uint64_t)gridDim.x * blockDim.x) { // (whole warps stay in)
		const uint64_t i = i0 + threadIdx.x;
		const uint4 v = i < n_quads ? __ldcs(w + i) : make_uint4(0, 0, 0, 0);
		const int c = __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
		bits += c;
		// a block = 4 consecutive pieces = 4 consecutive lanes
		const unsigned any = __ballot_sync(0xffffffffu, c != 0);
		if ((threadIdx.x & 3) == 0 && (any >> (threadIdx.x & 31) & 0xF)) ++blocks;
	}
	block_add(out, bits);
	block_add(out + 1, blocks);
}

static int bf_call1(const bfc_bf_t *b, uint64_t hash, bool insert)
{
	if (bfcg_rt_init() != BFCG_OK) return -1;
	BfcgRuntime &rt = bfcg_rt();
	int *d_ret = (int*)bfcg_arena(256), ret = -1;
	if (!d_ret) return -1;
	if (insert) k_bf_insert1<<<1, 1, 0, rt.stream>>>(bloom_view(b), hash, d_ret);
	else k_bf_get1<<<1, 1, 0, rt.stream>>>(bloom_view(b), hash, d_ret);
	++rt.n_launches;
	if (cudaMemcpyAsync(&ret, d_ret, sizeof(int), cudaMemcpyDeviceToHost, rt.stream) != cudaSuccess ||
		cudaStreamSynchronize(rt.stream) != cudaSuccess) {
		bfcg_fail(__func__, "one-element Bloom kernel", cudaGetLastError());
		return -1;
	}
	return ret;
}

extern "C" {

// reference bbf.c:5-17
bfc_bf_t *bfc_bf_init(int n_shift, int n_hashes)
{
	if (n_shift + BFC_BLK_SHIFT > 64 || n_shift < BFC_BLK_SHIFT) return 0;
	if (bfcg_rt_init() != BFCG_OK) return 0;
	bfc_bf_t *b = (bfc_bf_t*)calloc(1, sizeof(bfc_bf_t));
	b->n_shift = n_shift, b->n_hashes = n_hashes;
	const size_t bytes = (size_t)1 << (n_shift - 3);
	if (cudaMalloc(&b->b, bytes) != cudaSuccess) {
		bfcg_fail(__func__, "cudaMalloc(Bloom filter)", cudaErrorMemoryAllocation);
		free(b);
		return 0;
	}
	cudaMemsetAsync(b->b, 0, bytes, bfcg_rt().stream);
	cudaStreamSynchronize(bfcg_rt().stream);
	return b;
}

// one shard of a filter split over n_owners ranks by the top bits of the block index: n_shift stays the
// GLOBAL one (the probe positions derive from it, bbf.c:27-33), the allocation is 1/n_owners of the bytes
bfc_bf_t *bfcg_bf_init_shard(int n_shift, int n_hashes, int n_owners)
{
	int bits = 0;
	while ((1 << bits) < n_owners) ++bits;
	if (n_shift + BFC_BLK_SHIFT > 64 || n_shift < BFC_BLK_SHIFT || (1 << bits) != n_owners || n_shift - BFC_BLK_SHIFT < bits) return 0;
	if (bfcg_rt_init() != BFCG_OK) return 0;
	bfc_bf_t *b = (bfc_bf_t*)calloc(1, sizeof(bfc_bf_t));
	b->n_shift = n_shift, b->n_hashes = n_hashes;
	const size_t bytes = ((size_t)1 << (n_shift - 3)) >> bits;
	if (cudaMalloc(&b->b, bytes) != cudaSuccess) {
		bfcg_fail(__func__, "cudaMalloc(Bloom filter shard)", cudaErrorMemoryAllocation);
		free(b);
		return 0;
	}
	cudaMemsetAsync(b->b, 0, bytes, bfcg_rt().stream);
	cudaStreamSynchronize(bfcg_rt().stream);
	return b;
}

// reference bbf.c:19-23
void bfc_bf_destroy(bfc_bf_t *b)
{
	if (b == 0) return;
	cudaFree(b->b);
	free(b);
}

int bfc_bf_insert(bfc_bf_t *b, uint64_t hash) { return bf_call1(b, hash, true); }
int bfc_bf_get(const bfc_bf_t *b, uint64_t hash) { return bf_call1(b, hash, false); }

// Occupancy telemetry -- what the reference's bfc_bf_load was meant to give (bbf.c:65-79: unused there, not in its
// header, and its loop bound is wrong): the fraction of the filter's data bits that are set (`n_owners` > 1: of this
// shard), the fraction of blocks touched, and the false-positive rate a blocked filter with that load has for a k-mer
// never inserted: load^n_hashes (each probe lands on a set bit with probability = load, inside one block).
int bfcg_bf_load(const bfc_bf_t *bf, int n_owners, double *bit_load, double *block_load, double *fp_rate)
{
	int r;
	if ((r = bfcg_rt_init()) != BFCG_OK) return r;
	BfcgRuntime &rt = bfcg_rt();
	if (!bf || n_owners < 1) return BFCG_ERR_ARG;
	const uint64_t bytes = ((uint64_t)1 << (bf->n_shift - 3)) / (uint64_t)n_owners, n_blocks = bytes >> 6;
	unsigned long long *d = (unsigned long long*)bfcg_arena(256), h[2] = {0, 0};
	if (!d) return BFCG_ERR_NOMEM;
	BFCG_CUDA(cudaMemsetAsync(d, 0, 16, rt.stream));
	if (n_blocks) k_bf_load<<<rt.sm_count * 8, 256, 0, rt.stream>>>((const uint4*)bf->b, bytes >> 4, d);
	BFCG_LAUNCH_CHECK();
	BFCG_CUDA(cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, rt.stream));
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	const double data_bits = (double)n_blocks * (512 - 8); // bits 0-7 of a block are the reference's lock byte: never data
	const double load = data_bits > 0 ? (double)h[0] / data_bits : 0.0;
	if (bit_load) *bit_load = load;
	if (block_load) *block_load = n_blocks ? (double)h[1] / (double)n_blocks : 0.0;
	if (fp_rate) { double f = 1.0; for (int i = 0; i < bf->n_hashes; ++i) f *= load; *fp_rate = f; }
	return BFCG_OK;
}

// The `-b` a filter needs so that, after `n_distinct` different k-mers, a k-mer seen for the first time passes it with
// probability <= target_fp (the reference sizes by genome size alone: b = log2(size) + 8, bfc.c:50-52).  Load after n
// insertions of H bits each into m bits: 1 - exp(-H n / m); false-positive rate = load^H.
int bfcg_bf_suggest_shift(uint64_t n_distinct, int n_hashes, double target_fp)
{
	if (n_hashes < 1 || target_fp <= 0.0 || target_fp >= 1.0) return -1;
	const double load = pow(target_fp, 1.0 / n_hashes);
	const double m = -(double)n_hashes * (double)n_distinct / log(1.0 - load) * (512.0 / 504.0);
	int b = BFC_BLK_SHIFT;
	while (b < BFC_MAX_BF_SHIFT && (double)((uint64_t)1 << b) < m) ++b;
	return b;
}

int bfcg_bf_download(const bfc_bf_t *bf, uint8_t *dst)
{
	return bfcg_d2h(dst, bf->b, (uint64_t)1 << (bf->n_shift - 3));
}

int bfcg_bf_upload(bfc_bf_t *bf, const uint8_t *src)
{
	return bfcg_h2d(bf->b, src, (uint64_t)1 << (bf->n_shift - 3));
}

int bfcg_bf_clear(bfc_bf_t *bf)
{
	if (bfcg_rt_init() != BFCG_OK) return BFCG_ERR_CUDA;
	BFCG_CUDA(cudaMemsetAsync(bf->b, 0, (size_t)1 << (bf->n_shift - 3), bfcg_rt().stream));
	BFCG_CUDA(cudaStreamSynchronize(bfcg_rt().stream));
	return BFCG_OK;
}

} // extern "C"
