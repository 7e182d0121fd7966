// trim.cu -- the `-1` trim lookup: max_streak (reference correct.c:478-497) and the keep
// rule of worker_ec's filter_mode branch (correct.c:555-569), one read per thread.
#include "common.cuh"
#include <algorithm>
#include <vector>
#include <cstring>

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
	uint64_t limit = batch_bytes_limit();
	if (host && getenv("BFC_B200_EC_BATCH") == 0) // at least three windows (of at least 16 MB), so that a small host batch overlaps its copies too
		limit = std::min<uint64_t>(limit, std::max<uint64_t>((uint64_t)16 << 20, batch->n_bytes / 3 + 1));

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
	// host batch: windows of at most `limit` bytes, two buffer sets.  Window w+1 is copied in on its own stream while
	// window w is looked up; the results travel back through a pinned ring on a third stream and are handed to the
	// caller one window late (a copy straight into pageable caller memory would block the host, and the next launch).
	std::vector<int64_t> cut;
	for (int64_t r0 = 0; r0 < n;) {
		int64_t r1 = r0 + 1;
		while (r1 < n && batch->off[r1 + 1] - batch->off[r0] <= limit) ++r1;
		cut.push_back(r0);
		r0 = r1;
	}
	cut.push_back(n);
	const size_t n_win = cut.size() - 1;
	int64_t nr_max = 0;
	uint64_t nb_max = 0;
	for (size_t w = 0; w < n_win; ++w) {
		nr_max = std::max<int64_t>(nr_max, cut[w + 1] - cut[w]);
		nb_max = std::max<uint64_t>(nb_max, batch->off[cut[w + 1]] - batch->off[cut[w]]);
	}
	enum { NBUF = 2 };
	size_t tot = 0, o_seq[NBUF], o_off[NBUF], o_keep[NBUF], o_ts[NBUF], o_te[NBUF], o_ctr[NBUF];
	for (int b = 0; b < NBUF; ++b) {
		o_seq[b] = tot; tot = align_up(tot + nb_max, 256);
		o_off[b] = tot; tot = align_up(tot + (nr_max + 1) * 8, 256);
		o_keep[b] = tot; tot = align_up(tot + nr_max, 256);
		o_ts[b] = tot; tot = align_up(tot + nr_max * 4, 256);
		o_te[b] = tot; tot = align_up(tot + nr_max * 4, 256);
		o_ctr[b] = tot; tot += 256;
	}
	uint8_t *a = (uint8_t*)bfcg_arena(tot);
	if (!a) return BFCG_ERR_NOMEM;
	const size_t out_stride = align_up((size_t)nr_max * 9 + 64, 256), off_stride = (size_t)(nr_max + 1) * 8;
	uint8_t *pin_out = bfcg_pinned(0, NBUF * out_stride), *pin_off = bfcg_pinned(1, NBUF * off_stride);
	if (!pin_out || !pin_off) return BFCG_ERR_NOMEM;
	auto copy_in = [&](size_t w) -> cudaError_t {
		const int b = (int)(w % NBUF);
		const int64_t r0 = cut[w], nr = cut[w + 1] - r0;
		const uint64_t b0 = batch->off[r0], nb = batch->off[r0 + nr] - b0;
		uint64_t *rel = (uint64_t*)(pin_off + b * off_stride);
		cudaError_t ce;
		if (w >= NBUF) {
			if ((ce = cudaEventSynchronize(rt.ev_in[b])) != cudaSuccess) return ce;               // the ring slot has left the host
			if ((ce = cudaStreamWaitEvent(rt.copy_in, rt.ev_done[b], 0)) != cudaSuccess) return ce; // window w - NBUF has been looked up
		}
		for (int64_t i = 0; i <= nr; ++i) rel[i] = batch->off[r0 + i] - b0;
		if ((ce = cudaMemcpyAsync(a + o_seq[b], batch->seq + b0, nb, cudaMemcpyHostToDevice, rt.copy_in)) != cudaSuccess) return ce;
		if ((ce = cudaMemcpyAsync(a + o_off[b], rel, (size_t)(nr + 1) * 8, cudaMemcpyHostToDevice, rt.copy_in)) != cudaSuccess) return ce;
		return cudaEventRecord(rt.ev_in[b], rt.copy_in);
	};
	auto drain = [&](size_t w) -> cudaError_t { // window w's results: pinned ring -> caller
		const int b = (int)(w % NBUF);
		const int64_t r0 = cut[w], nr = cut[w + 1] - r0;
		const uint8_t *o = pin_out + b * out_stride;
		const cudaError_t ce = cudaEventSynchronize(rt.ev_out[b]);
		if (ce != cudaSuccess) return ce;
		memcpy(keep + r0, o, (size_t)nr);
		memcpy(tstart + r0, o + align_up((size_t)nr_max, 8), (size_t)nr * 4);
		memcpy(tend + r0, o + align_up((size_t)nr_max, 8) + (size_t)nr_max * 4, (size_t)nr * 4);
		if (stats) {
			unsigned long long c1;
			memcpy(&c1, o + align_up((size_t)nr_max, 8) + (size_t)nr_max * 8 + 8, 8);
			stats->n_lookups += c1;
		}
		return cudaSuccess;
	};
	BFCG_CUDA(copy_in(0));
	for (size_t w = 0; w < n_win; ++w) {
		const int b = (int)(w % NBUF);
		const int64_t nr = cut[w + 1] - cut[w];
		if (w + 1 < n_win) BFCG_CUDA(copy_in(w + 1));
		BFCG_CUDA(cudaStreamWaitEvent(rt.stream, rt.ev_in[b], 0));
		if (w >= NBUF) BFCG_CUDA(cudaStreamWaitEvent(rt.stream, rt.ev_out[b], 0)); // the result buffers of window w - NBUF have been read out
		TrimParams P;
		P.off = (const uint64_t*)(a + o_off[b]), P.seq = a + o_seq[b], P.n_reads = nr, P.bf = bloom_view(bf_high), P.k = opt->k, P.min_frac = opt->min_frac;
		P.keep = a + o_keep[b], P.tstart = (int32_t*)(a + o_ts[b]), P.tend = (int32_t*)(a + o_te[b]), P.ctr = (unsigned long long*)(a + o_ctr[b]);
		BFCG_CUDA(cudaMemsetAsync(P.ctr, 0, 64, rt.stream));
		{ KTime kt(KT_TRIM); k_trim<<<(unsigned)std::min<int64_t>((nr + 255) / 256, (int64_t)rt.sm_count * 8), 256, 0, rt.stream>>>(P); }
		BFCG_LAUNCH_CHECK();
		BFCG_CUDA(cudaEventRecord(rt.ev_done[b], rt.stream));
		if (w >= NBUF) BFCG_CUDA(drain(w - NBUF)); // before ring slot b is written (and its event re-recorded) for this window
		uint8_t *o = pin_out + b * out_stride;
		BFCG_CUDA(cudaStreamWaitEvent(rt.copy_out, rt.ev_done[b], 0));
		BFCG_CUDA(cudaMemcpyAsync(o, a + o_keep[b], (size_t)nr, cudaMemcpyDeviceToHost, rt.copy_out));
		BFCG_CUDA(cudaMemcpyAsync(o + align_up((size_t)nr_max, 8), a + o_ts[b], (size_t)nr * 4, cudaMemcpyDeviceToHost, rt.copy_out));
		BFCG_CUDA(cudaMemcpyAsync(o + align_up((size_t)nr_max, 8) + (size_t)nr_max * 4, a + o_te[b], (size_t)nr * 4, cudaMemcpyDeviceToHost, rt.copy_out));
		BFCG_CUDA(cudaMemcpyAsync(o + align_up((size_t)nr_max, 8) + (size_t)nr_max * 8, a + o_ctr[b], 16, cudaMemcpyDeviceToHost, rt.copy_out));
		BFCG_CUDA(cudaEventRecord(rt.ev_out[b], rt.copy_out));
	}
	for (size_t w = n_win > NBUF ? n_win - NBUF : 0; w < n_win; ++w) BFCG_CUDA(drain(w));
	timer.stop();
	return BFCG_OK;
}
