// trim.cu -- the `-1` trim lookup: max_streak (reference correct.c:478-497) and the keep
// rule of worker_ec's filter_mode branch (correct.c:555-569), one read per thread.
#include "common.cuh"
#include <algorithm>
#include <vector>

struct TrimParams {
	const uint64_t *off;
	const uint8_t *seq;
	int64_t n_reads;
	BloomView bf;
	int k;
	float min_frac;
	uint8_t *keep;
	int32_t *tstart, *tend;
	unsigned long long *ctr; // [1] n_lookups
};

// reference correct.c:478-497 (max_streak) + the keep rule of worker_ec (correct.c:555-569)
__global__ void __launch_bounds__(256) k_trim(TrimParams P)
{
	unsigned long long n_lookups = 0;
	for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < P.n_reads; r += (int64_t)gridDim.x * blockDim.x) {
		const uint64_t o = P.off[r];
		const int n = (int)(P.off[r + 1] - o - 1), k = P.k;
		const uint8_t *seq = P.seq + o;
		uint64_t x[4] = {0, 0, 0, 0}, max = 0, t = 0;
		int l = 0;
		for (int i = 0; i < n; ++i) {
			const int c = base_code(seq[i]);
			if (c < 4) {
				bfc_kmer_append(k, x, c);
				if (++l >= k) {
					uint64_t y[2];
					const BloomProbe pr = bloom_locate(bfc_kmer_hash(k, x, y), P.bf.n_shift);
					++n_lookups;
					if (bloom_count_set<false>(P.bf.w + (pr.blk << 4), pr, P.bf.n_hashes) == P.bf.n_hashes) t += 1ULL << 32;
					else t = i + 1;
				} else t = i + 1;
			} else l = 0, x[0] = x[1] = x[2] = x[3] = 0, t = i + 1;
			max = max > t ? max : t;
		}
		uint8_t keep = 0;
		int32_t ts = 0, te = 0;
		if (max >> 32 && (double)((max >> 32) + k) / n > P.min_frac) { // float min_frac promoted, as in C
			const int start = (int)(uint32_t)max;
			te = start + (int)(max >> 32), ts = start - (k - 1), keep = 1;
		}
		P.keep[r] = keep, P.tstart[r] = ts, P.tend[r] = te;
	}
	block_add(P.ctr + 1, n_lookups);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static uint64_t batch_bytes_limit()
{
	const char *e = getenv("BFC_B200_EC_BATCH");
	return e && atoll(e) >= 4096 ? (uint64_t)atoll(e) : 1ULL << 27;
}

extern "C" int bfcg_trim_batch(const bfc_opt_t *opt, const bfc_bf_t *bf_high, const bfcg_batch_t *batch,
                               uint8_t *keep, int32_t *tstart, int32_t *tend, bfcg_stats_t *stats)
{
	int r;
	if ((r = bfcg_rt_init()) != BFCG_OK) return r;
	BfcgRuntime &rt = bfcg_rt();
	if (!opt || !bf_high || !batch || !batch->off || !keep || !tstart || !tend || opt->k < 1 || opt->k > BFC_MAX_KMER)
		return bfcg_fail(__func__, "invalid arguments", cudaSuccess), BFCG_ERR_ARG;
	if (batch->n_reads == 0) return BFCG_OK;
	const bool host = batch->where == BFCG_HOST;
	const int64_t n = batch->n_reads;
	const uint64_t limit = batch_bytes_limit();

	BfcgTimer timer(stats);
	if (!host) {
		unsigned long long *ctr = (unsigned long long*)bfcg_arena(256), c[2];
		if (!ctr) return BFCG_ERR_NOMEM;
		TrimParams P;
		P.off = batch->off, P.seq = batch->seq, P.n_reads = n, P.bf = bloom_view(bf_high), P.k = opt->k, P.min_frac = opt->min_frac;
		P.keep = keep, P.tstart = tstart, P.tend = tend, P.ctr = ctr;
		BFCG_CUDA(cudaMemsetAsync(ctr, 0, 64, rt.stream));
		{ KTime kt(KT_TRIM); k_trim<<<(unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)rt.sm_count * 8), 256, 0, rt.stream>>>(P); }
		BFCG_LAUNCH_CHECK();
		BFCG_CUDA(cudaMemcpyAsync(c, ctr, sizeof(c), cudaMemcpyDeviceToHost, rt.stream));
		timer.stop();
		BFCG_CUDA(cudaStreamSynchronize(rt.stream));
		if (stats) stats->n_lookups += c[1];
		return BFCG_OK;
	}
	for (int64_t r0 = 0; r0 < n;) {
		int64_t r1 = r0 + 1;
		while (r1 < n && batch->off[r1 + 1] - batch->off[r0] <= limit) ++r1;
		const int64_t nr = r1 - r0;
		const uint64_t b0 = batch->off[r0], nb = batch->off[r1] - b0;
		size_t tot = 0, o_seq, o_off, o_keep, o_ts, o_te, o_ctr;
		o_seq = tot; tot = align_up(tot + nb, 256);
		o_off = tot; tot = align_up(tot + (nr + 1) * 8, 256);
		o_keep = tot; tot = align_up(tot + nr, 256);
		o_ts = tot; tot = align_up(tot + nr * 4, 256);
		o_te = tot; tot = align_up(tot + nr * 4, 256);
		o_ctr = tot; tot += 256;
		uint8_t *a = (uint8_t*)bfcg_arena(tot);
		if (!a) return BFCG_ERR_NOMEM;
		std::vector<uint64_t> rel(nr + 1);
		for (int64_t i = 0; i <= nr; ++i) rel[i] = batch->off[r0 + i] - b0;
		BFCG_CUDA(cudaMemcpyAsync(a + o_seq, batch->seq + b0, nb, cudaMemcpyHostToDevice, rt.stream));
		BFCG_CUDA(cudaMemcpyAsync(a + o_off, rel.data(), (nr + 1) * 8, cudaMemcpyHostToDevice, rt.stream));
		TrimParams P;
		P.off = (const uint64_t*)(a + o_off), P.seq = a + o_seq, P.n_reads = nr, P.bf = bloom_view(bf_high), P.k = opt->k, P.min_frac = opt->min_frac;
		P.keep = a + o_keep, P.tstart = (int32_t*)(a + o_ts), P.tend = (int32_t*)(a + o_te), P.ctr = (unsigned long long*)(a + o_ctr);
		BFCG_CUDA(cudaMemsetAsync(P.ctr, 0, 64, rt.stream));
		{ KTime kt(KT_TRIM); k_trim<<<(unsigned)std::min<int64_t>((nr + 255) / 256, (int64_t)rt.sm_count * 8), 256, 0, rt.stream>>>(P); }
		BFCG_LAUNCH_CHECK();
		unsigned long long c[2];
		BFCG_CUDA(cudaMemcpyAsync(c, P.ctr, sizeof(c), cudaMemcpyDeviceToHost, rt.stream));
		BFCG_CUDA(cudaMemcpyAsync(keep + r0, a + o_keep, nr, cudaMemcpyDeviceToHost, rt.stream));
		BFCG_CUDA(cudaMemcpyAsync(tstart + r0, a + o_ts, nr * 4, cudaMemcpyDeviceToHost, rt.stream));
		BFCG_CUDA(cudaMemcpyAsync(tend + r0, a + o_te, nr * 4, cudaMemcpyDeviceToHost, rt.stream));
		BFCG_CUDA(cudaStreamSynchronize(rt.stream));
		if (stats) stats->n_lookups += c[1];
		r0 = r1;
	}
	timer.stop();
	return BFCG_OK;
}
