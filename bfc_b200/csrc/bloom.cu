// bloom.cu -- the reference's bfc_bf_* entry points (bbf.h:14-17) over a device-resident
// filter.  The single-element insert/get exist for API compatibility (one-thread
// kernels); the throughput path is bfcg_count_batch / bfcg_trim_batch.
#include "common.cuh"

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
