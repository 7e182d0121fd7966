// count.cu -- the count phase: what the reference runs inside
// kt_for(..., worker_count, ...) (count.c:106, :72-89) with the Bloom -> table cascade
// of bfc_kmer_insert (count.c:54-70), reproduced with the SEQUENTIAL (-t1) semantics.
//
// Order dependence (count.c:9-18): occurrence t of a k-mer "passes" iff all n_hashes
// Bloom bits were set by occurrences < t.  All ordering constraints are local to one
// 64-byte Bloom block, so a sub-batch of the read stream is processed as
//
//   K0 k_enum           rolling canonical k-mer + hash per stream position (enum.cuh):
//                       a dense array of 16-byte records (y0 | is_high << 63, y1).
//   K1 k_count_probe    one thread per record: test its bits against the filter as it
//                       was BEFORE the sub-batch (no data bit is written in K1).  All
//                       bits set -> it passes regardless of order: table upsert (or
//                       bf_high insert in trim mode) right away.  Otherwise it is
//                       "pending": its index is appended to a list and it marks its
//                       Bloom block in the two spare bits of the block's lock byte
//                       (bit 0: one pending occurrence, bit 1: more than one).
//   K2 k_count_resolve  a pending occurrence alone in its block sets a bit nobody else
//                       touches in this sub-batch: it cannot pass; its bits are OR-ed in.
//                       Pending occurrences that share a block go to the conflict list.
//   K3 radix sort       conflict list by (block, stream position)
//   K4 k_count_replay   one thread per conflicting block replays its occurrences in
//                       stream order with the reference's test-then-set, exactly.
//
// The lock byte is 0 again when the sub-batch ends, as in the reference (bbf.c:43).
#include "common.cuh"
#include "enum.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>

struct CountParams {
	const unsigned long long *rec_y0, *rec_y1; // records of the window, in k_enum's blocked order
	uint64_t n_rec;              // records (= padded window positions)
	int k;
	BloomView bf, bf_high;       // bf_high.w == 0 in normal mode
	TabView tab;                 // tab.slots == 0 in trim mode
	uint32_t *pend;              // record indices of the pending occurrences
	unsigned long long *ctr;     // [0] n_pending [1] n_kmers [2] n_pass [3] n_conflict
	unsigned long long *conf_key;
	uint32_t *conf_val;
	int linear;                  // records are in stream order (received from other ranks); else k_enum's blocked order
};

__global__ void __launch_bounds__(256) k_count_probe(CountParams p)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned lane = threadIdx.x & 31;
	bool pend = false, pass = false;
	uint64_t y0f = 0, y1 = ~0ULL;
	BloomProbe pr;
	pr.blk = 0, pr.h1 = pr.h2 = 0;
	if (i < p.n_rec) y1 = __ldg(p.rec_y1 + i);
	const bool valid = y1 != ~0ULL;
	if (valid) {
		y0f = __ldg(p.rec_y0 + i);
		pr = bloom_locate(hash_from_y(p.k, y0f & ~(1ULL << 63), y1), p.bf.n_shift);
		pass = bloom_count_set<false>(bloom_block(p.bf, pr.blk), pr, p.bf.n_hashes) == p.bf.n_hashes;
		pend = !pass;
	}
	// warp-aggregated append of the pending occurrences
	const unsigned pm = __ballot_sync(0xffffffffu, pend);
	if (pm) {
		unsigned long long at = 0;
		if (lane == (unsigned)(__ffs(pm) - 1)) at = atomicAdd(p.ctr, (unsigned long long)__popc(pm));
		at = __shfl_sync(0xffffffffu, at, __ffs(pm) - 1);
		if (pend) {
			p.pend[at + __popc(pm & ((1u << lane) - 1))] = (uint32_t)i;
			uint32_t *w0 = bloom_block(p.bf, pr.blk);
			if (atomicOr(w0, 1u) & 1u) atomicOr(w0, 2u);
		}
	}
	unsigned long long n_new = 0;
	if (pass) {
		const uint64_t y0 = y0f & ~(1ULL << 63);
		if (p.tab.slots) n_new = tab_upsert(p.tab, y0, y1, (int)(y0f >> 63)) == 1;
		else {
			const BloomProbe ph = bloom_locate(hash_from_y(p.k, y0, y1), p.bf_high.n_shift);
			bloom_set_atomic(bloom_block(p.bf_high, ph.blk), ph, p.bf_high.n_hashes);
		}
	}
	block_add(p.ctr + 1, valid ? 1ULL : 0ULL);
	block_add(p.ctr + 2, pass ? 1ULL : 0ULL);
	if (p.tab.slots) block_add(p.tab.counters, n_new);
}

__global__ void __launch_bounds__(256) k_count_resolve(CountParams p, uint64_t n_pending)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned lane = threadIdx.x & 31;
	bool conflict = false;
	uint64_t blk = 0;
	uint32_t r = 0;
	if (i < n_pending) {
		r = p.pend[i];
		const uint64_t y0 = p.rec_y0[r] & ~(1ULL << 63), y1 = p.rec_y1[r];
		const BloomProbe pr = bloom_locate(hash_from_y(p.k, y0, y1), p.bf.n_shift);
		uint32_t *w = bloom_block(p.bf, pr.blk);
		blk = pr.blk;
		if (__ldcg(w) & 2u) conflict = true;
		else { // the only pending occurrence of this block: sets at least one new bit => does not pass
			bloom_set_atomic(w, pr, p.bf.n_hashes);
			atomicAnd(w, ~1u);
		}
	}
	const unsigned cm = __ballot_sync(0xffffffffu, conflict);
	if (cm) {
		unsigned long long at = 0;
		if (lane == (unsigned)(__ffs(cm) - 1)) at = atomicAdd(p.ctr + 3, (unsigned long long)__popc(cm));
		at = __shfl_sync(0xffffffffu, at, __ffs(cm) - 1);
		if (conflict) {
			at += __popc(cm & ((1u << lane) - 1));
			p.conf_key[at] = blk << 32 | (p.linear ? r : enum_pos_of_record(r)); // order inside a block = stream order
			p.conf_val[at] = r;
		}
	}
}

__global__ void __launch_bounds__(256) k_count_replay(CountParams p, const unsigned long long *key, const uint32_t *val, uint64_t n)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long n_pass = 0, n_new = 0;
	if (i < n) {
		const uint64_t blk = key[i] >> 32;
		if (i == 0 || (key[i - 1] >> 32) != blk) { // head of the block's run: replay it in stream order
			volatile uint32_t *w = bloom_block(p.bf, blk);
			const int H = p.bf.n_hashes;
			for (uint64_t j = i; j < n && (key[j] >> 32) == blk; ++j) {
				const uint32_t r = val[j];
				const uint64_t y0f = p.rec_y0[r], y0 = y0f & ~(1ULL << 63), y1 = p.rec_y1[r];
				const int is_high = (int)(y0f >> 63);
				const uint64_t hash = hash_from_y(p.k, y0, y1);
				const BloomProbe pr = bloom_locate(hash, p.bf.n_shift);
				// reference bbf.c:35-42: test-then-set each probe
				int z = pr.h1, done = 0, cnt = 0;
				while (done < H) {
					if (z >= 8) {
						const uint32_t bit = 1u << (z & 31), v = w[z >> 5];
						cnt += (v & bit) != 0;
						w[z >> 5] = v | bit;
						++done;
					}
					z = (z + pr.h2) & BFC_BLK_MASK;
				}
				if (cnt == H) { // reference count.c:60
					++n_pass;
					if (p.tab.slots) n_new += tab_upsert(p.tab, y0, y1, is_high) == 1;
					else {
						const BloomProbe ph = bloom_locate(hash, p.bf_high.n_shift);
						bloom_set_atomic(bloom_block(p.bf_high, ph.blk), ph, p.bf_high.n_hashes);
					}
				}
			}
			w[0] = w[0] & ~3u; // release the block's marker bits
		}
	}
	block_add(p.ctr + 2, n_pass);
	if (p.tab.slots) block_add(p.tab.counters, n_new);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static uint64_t sub_batch_positions(const bfc_opt_t *opt)
{
	const char *e = getenv("BFC_B200_SUBBATCH");
	if (e && atoll(e) >= 1024) return (uint64_t)atoll(e);
	// conflicts cost a sort: keep the expected number of pending occurrences per Bloom
	// block well below 1 (2^(b-9) blocks), within [2^20, 2^26] positions per sub-batch
	int lg = opt->bf_shift - BFC_BLK_SHIFT - 2;
	if (lg < 20) lg = 20;
	if (lg > 26) lg = 26;
	return 1ULL << lg;
}

// scratch of one count call, carved from the arena for the worst case (every record pending and conflicting)
struct CountScratch {
	size_t temp_bytes;
	uint8_t *tmp;
	unsigned long long *ck1;
	uint32_t *cv1;
	size_t bytes(uint64_t rec_max)
	{
		temp_bytes = 0;
		cub::DeviceRadixSort::SortPairs((void*)0, temp_bytes, (const unsigned long long*)0, (unsigned long long*)0,
		                                (const uint32_t*)0, (uint32_t*)0, (size_t)rec_max, 0, 64, bfcg_rt().stream);
		return align_up(rec_max * 4, 256) + 2 * align_up(rec_max * 8, 256) + 2 * align_up(rec_max * 4, 256) + align_up(temp_bytes, 256) + 256;
	}
	void carve(uint8_t *a, uint64_t rec_max, CountParams &p)
	{
		size_t o = 0;
		p.pend = (uint32_t*)(a + o); o += align_up(rec_max * 4, 256);
		p.conf_key = (unsigned long long*)(a + o); o += align_up(rec_max * 8, 256);
		ck1 = (unsigned long long*)(a + o); o += align_up(rec_max * 8, 256);
		p.conf_val = (uint32_t*)(a + o); o += align_up(rec_max * 4, 256);
		cv1 = (uint32_t*)(a + o); o += align_up(rec_max * 4, 256);
		tmp = a + o; o += align_up(temp_bytes, 256);
		p.ctr = (unsigned long long*)(a + o);
	}
};

// K1-K4 over p.n_rec records (p.rec_y0 / p.rec_y1 set by the caller); grows the table first
static int count_window(CountParams &p, CountScratch &sc, bfc_bf_t *bf, bfc_ch_t *ch, uint64_t max_new_keys, bfcg_stats_t *stats)
{
	BfcgRuntime &rt = bfcg_rt();
	int r;
	if (ch) {
		if ((r = bfcg_tab_reserve(ch, max_new_keys)) != BFCG_OK) return r;
		p.tab = tab_view(ch);
	}
	BFCG_CUDA(cudaMemsetAsync(p.ctr, 0, 64, rt.stream));
	{ KTime kt(KT_COUNT_PROBE); k_count_probe<<<(unsigned)((p.n_rec + 255) / 256), 256, 0, rt.stream>>>(p); }
	BFCG_LAUNCH_CHECK();
	unsigned long long c[4];
	BFCG_CUDA(cudaMemcpyAsync(c, p.ctr, sizeof(c), cudaMemcpyDeviceToHost, rt.stream));
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	const uint64_t n_pending = c[0];
	uint64_t n_conflict = 0;
	if (n_pending) {
		{ KTime kt(KT_COUNT_RESOLVE); k_count_resolve<<<(unsigned)((n_pending + 255) / 256), 256, 0, rt.stream>>>(p, n_pending); }
		BFCG_LAUNCH_CHECK();
		BFCG_CUDA(cudaMemcpyAsync(c, p.ctr, sizeof(c), cudaMemcpyDeviceToHost, rt.stream));
		BFCG_CUDA(cudaStreamSynchronize(rt.stream));
		n_conflict = c[3];
	}
	if (n_conflict) {
		size_t tb = sc.temp_bytes;
		KTime *kts = new KTime(KT_COUNT_SORT);
		const cudaError_t se = cub::DeviceRadixSort::SortPairs(sc.tmp, tb, (const unsigned long long*)p.conf_key, sc.ck1,
		                                                       (const uint32_t*)p.conf_val, sc.cv1, (size_t)n_conflict, 0,
		                                                       32 + (bf->n_shift - BFC_BLK_SHIFT), rt.stream);
		delete kts;
		BFCG_CUDA(se);
		rt.n_launches += 1 + (32 + bf->n_shift - BFC_BLK_SHIFT + 7) / 8; // histogram + one onesweep pass per 8 bits
		{ KTime kt(KT_COUNT_REPLAY); k_count_replay<<<(unsigned)((n_conflict + 255) / 256), 256, 0, rt.stream>>>(p, sc.ck1, sc.cv1, n_conflict); }
		BFCG_LAUNCH_CHECK();
		BFCG_CUDA(cudaMemcpyAsync(c, p.ctr, sizeof(c), cudaMemcpyDeviceToHost, rt.stream));
		BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	}
	if (ch && (r = bfcg_tab_drain_deferred(ch)) != BFCG_OK) return r;
	if (stats) {
		stats->n_kmers += c[1], stats->n_pass += c[2];
		stats->n_pending += n_pending, stats->n_conflict += n_conflict;
	}
	return BFCG_OK;
}

static bool count_args_ok(const bfc_opt_t *opt, const bfc_bf_t *bf, const bfc_bf_t *bf_high, const bfc_ch_t *ch)
{
	return opt && bf && (ch == 0) != (bf_high == 0) && opt->k >= 1 && opt->k <= BFC_MAX_KMER &&
	       (!ch || bfc_ch_get_k(ch) == opt->k) && bf->n_hashes >= 1 &&
	       (!bf_high || (bf_high->n_shift == bf->n_shift && bf_high->n_hashes == bf->n_hashes));
}

extern "C" int bfcg_count_batch(const bfc_opt_t *opt, bfc_bf_t *bf, bfc_bf_t *bf_high, bfc_ch_t *ch,
                                const bfcg_batch_t *batch, bfcg_stats_t *stats)
{
	int r;
	if ((r = bfcg_rt_init()) != BFCG_OK) return r;
	BfcgRuntime &rt = bfcg_rt();
	if (!batch || !count_args_ok(opt, bf, bf_high, ch))
		return bfcg_fail(__func__, "invalid arguments", cudaSuccess), BFCG_ERR_ARG;
	if (batch->n_bytes == 0) return BFCG_OK;
	if (bfcg_count_part_usable(opt, bf->n_shift, 0)) return bfcg_count_part_batch(opt, bf, bf_high, ch, batch, stats);

	const uint64_t sub = sub_batch_positions(opt);
	const uint64_t halo = opt->k - 1;
	const uint64_t win_max = std::min<uint64_t>(sub, batch->n_bytes) + halo;
	const uint64_t rec_max = enum_padded(win_max);
	const bool host = batch->where == BFCG_HOST;

	CountScratch sc;
	size_t o_seq = 0, o_qual = 0, o_y0, o_y1, o_sc, tot = 0;
	if (host) { o_seq = tot; tot = align_up(tot + win_max, 256); o_qual = tot; tot = align_up(tot + win_max, 256); }
	o_y0 = tot; tot = align_up(tot + rec_max * 8, 256);
	o_y1 = tot; tot = align_up(tot + rec_max * 8, 256);
	o_sc = tot; tot += sc.bytes(rec_max);
	uint8_t *a = (uint8_t*)bfcg_arena(tot);
	if (!a) return BFCG_ERR_NOMEM;

	EnumParams ep;
	memset(&ep, 0, sizeof(ep));
	ep.k = opt->k, ep.q = opt->q;
	ep.rec_y0 = (unsigned long long*)(a + o_y0), ep.rec_y1 = (unsigned long long*)(a + o_y1);
	CountParams p;
	memset(&p, 0, sizeof(p));
	p.k = opt->k;
	p.rec_y0 = ep.rec_y0, p.rec_y1 = ep.rec_y1;
	p.bf = bloom_view(bf);
	if (bf_high) p.bf_high = bloom_view(bf_high);
	sc.carve(a + o_sc, rec_max, p);

	BfcgTimer timer(stats);
	for (uint64_t s = 0; s < batch->n_bytes; s += sub) {
		const uint64_t e = std::min(batch->n_bytes, s + sub);
		const uint64_t w0 = s >= halo ? s - halo : 0; // window start: k-1 bases of warm-up before s
		const uint64_t len = e - w0;
		if (host) {
			BFCG_CUDA(cudaMemcpyAsync(a + o_seq, batch->seq + w0, len, cudaMemcpyHostToDevice, rt.stream));
			if (batch->qual) BFCG_CUDA(cudaMemcpyAsync(a + o_qual, batch->qual + w0, len, cudaMemcpyHostToDevice, rt.stream));
			ep.seq = a + o_seq, ep.qual = batch->qual ? a + o_qual : 0;
		} else ep.seq = batch->seq + w0, ep.qual = batch->qual ? batch->qual + w0 : 0;
		ep.len = len, ep.emit_from = s - w0;
		p.n_rec = enum_padded(len - ep.emit_from);
		{ KTime kt(KT_ENUM); k_enum<<<(unsigned)(p.n_rec / ENUM_SEG), ENUM_THREADS, 0, rt.stream>>>(ep); }
		BFCG_LAUNCH_CHECK();
		if ((r = count_window(p, sc, bf, ch, e - s, stats)) != BFCG_OK) return r;
	}
	timer.stop();
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	return BFCG_OK;
}

// ------------------------------------------------------------------ sharded counting (one rank of N)

// owner of a k-mer = top log2(n_owners) bits of its Bloom block index (bbf.c:27-28)
__device__ __forceinline__ int record_owner(int k, int n_shift, int owner_bits, uint64_t y0, uint64_t y1)
{
	const int x = n_shift - BFC_BLK_SHIFT;
	const uint64_t blk = hash_from_y(k, y0, y1) & ((1ULL << x) - 1);
	return owner_bits ? (int)(blk >> (x - owner_bits)) : 0;
}

#define BK_MAX_OWNERS 8

struct BucketParams {
	const unsigned long long *rec_y0, *rec_y1; // k_enum's blocked order
	int k, n_shift, owner_bits, n_owners;
	uint32_t *seg_cnt;           // [n_seg][BK_MAX_OWNERS] records per (segment, owner); after the scan: output offsets
	unsigned long long *out_y0, *out_y1;
	unsigned long long *totals;  // [BK_MAX_OWNERS]
};

// One CTA per k_enum segment; thread t owns stream positions t*36 .. t*36+35 of the segment (records j*256 + t),
// so "stable" = thread-major.  PASS 0 counts, PASS 1 scatters to seg_cnt (now offsets) + rank inside the segment.
template <int PASS>
__global__ void __launch_bounds__(ENUM_THREADS) k_bucket(BucketParams p)
{
	__shared__ uint32_t s_warp[ENUM_THREADS / 32][BK_MAX_OWNERS];
	__shared__ uint32_t s_base[BK_MAX_OWNERS];
	const uint64_t seg = blockIdx.x;
	const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t cnt[BK_MAX_OWNERS];
#pragma unroll
	for (int o = 0; o < BK_MAX_OWNERS; ++o) cnt[o] = 0;
	for (int j = 0; j < ENUM_CHUNK; ++j) {
		const uint64_t i = seg * ENUM_SEG + (uint64_t)j * ENUM_THREADS + threadIdx.x;
		const unsigned long long y1 = __ldg(p.rec_y1 + i);
		if (y1 != ~0ULL) {
			const int o = record_owner(p.k, p.n_shift, p.owner_bits, __ldg(p.rec_y0 + i) & ~(1ULL << 63), y1);
#pragma unroll
			for (int q = 0; q < BK_MAX_OWNERS; ++q) cnt[q] += q == o;
		}
	}
	// exclusive prefix over the threads of the CTA, per owner
	uint32_t pre[BK_MAX_OWNERS];
#pragma unroll
	for (int o = 0; o < BK_MAX_OWNERS; ++o) {
		uint32_t v = cnt[o];
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t u = __shfl_up_sync(0xffffffffu, v, d);
			if (lane >= (unsigned)d) v += u;
		}
		pre[o] = v - cnt[o];
		if (lane == 31) s_warp[warp][o] = v;
	}
	__syncthreads();
	if (threadIdx.x < BK_MAX_OWNERS) {
		uint32_t run = 0;
		for (int w = 0; w < ENUM_THREADS / 32; ++w) { const uint32_t v = s_warp[w][threadIdx.x]; s_warp[w][threadIdx.x] = run; run += v; }
		if (PASS == 0) p.seg_cnt[seg * BK_MAX_OWNERS + threadIdx.x] = run;
		else s_base[threadIdx.x] = p.seg_cnt[seg * BK_MAX_OWNERS + threadIdx.x];
	}
	if (PASS == 0) return;
	__syncthreads();
	uint32_t at[BK_MAX_OWNERS];
#pragma unroll
	for (int o = 0; o < BK_MAX_OWNERS; ++o) at[o] = pre[o] + s_warp[warp][o];
	for (int j = 0; j < ENUM_CHUNK; ++j) {
		const uint64_t i = seg * ENUM_SEG + (uint64_t)j * ENUM_THREADS + threadIdx.x;
		const unsigned long long y1 = __ldg(p.rec_y1 + i);
		if (y1 != ~0ULL) {
			const unsigned long long y0f = __ldg(p.rec_y0 + i);
			const int o = record_owner(p.k, p.n_shift, p.owner_bits, y0f & ~(1ULL << 63), y1);
			uint32_t slot = 0;
#pragma unroll
			for (int q = 0; q < BK_MAX_OWNERS; ++q) if (q == o) slot = at[q]++;
			const uint64_t dst = p.totals[o] + s_base[o] + slot; // totals[] holds the bucket starts in pass 1
			p.out_y0[dst] = y0f, p.out_y1[dst] = y1;
		}
	}
}

// exclusive scan of seg_cnt over the segments, per owner; totals[o] = bucket size
__global__ void k_bucket_scan(uint32_t *seg_cnt, uint64_t n_seg, unsigned long long *totals)
{
	const int o = threadIdx.x;
	if (o >= BK_MAX_OWNERS) return;
	unsigned long long run = 0;
	for (uint64_t s = 0; s < n_seg; ++s) {
		const uint32_t v = seg_cnt[s * BK_MAX_OWNERS + o];
		seg_cnt[s * BK_MAX_OWNERS + o] = (uint32_t)run;
		run += v;
	}
	totals[o] = run;
}

static int log2_exact(int n)
{
	int b = 0;
	while ((1 << b) < n) ++b;
	return (1 << b) == n ? b : -1;
}

extern "C" int bfcg_enum_records(const bfc_opt_t *opt, const bfcg_batch_t *batch, int n_owners,
                                 uint64_t *d_y0, uint64_t *d_y1, uint64_t *counts)
{
	int r;
	if ((r = bfcg_rt_init()) != BFCG_OK) return r;
	BfcgRuntime &rt = bfcg_rt();
	const int owner_bits = log2_exact(n_owners);
	if (!opt || !batch || !d_y0 || !d_y1 || !counts || owner_bits < 0 || n_owners > BK_MAX_OWNERS ||
		opt->bf_shift - BFC_BLK_SHIFT < owner_bits || batch->n_bytes > (1ULL << 31) || opt->k < 1 || opt->k > BFC_MAX_KMER)
		return bfcg_fail(__func__, "invalid arguments (owners must be 1, 2, 4 or 8; at most 2^31 bytes per call)", cudaSuccess), BFCG_ERR_ARG;
	for (int o = 0; o < n_owners; ++o) counts[o] = 0;
	if (batch->n_bytes == 0) return BFCG_OK;
	if (bfcg_count_part_usable(opt, opt->bf_shift, owner_bits)) return bfcg_enum_part_records(opt, batch, owner_bits, d_y0, d_y1, counts);
	const bool host = batch->where == BFCG_HOST;
	const uint64_t nb = batch->n_bytes, n_rec = enum_padded(nb), n_seg = n_rec / ENUM_SEG;
	size_t o_seq = 0, o_qual = 0, o_y0, o_y1, o_cnt, o_tot, tot = 0;
	if (host) { o_seq = tot; tot = align_up(tot + nb, 256); o_qual = tot; tot = align_up(tot + nb, 256); }
	o_y0 = tot; tot = align_up(tot + n_rec * 8, 256);
	o_y1 = tot; tot = align_up(tot + n_rec * 8, 256);
	o_cnt = tot; tot = align_up(tot + n_seg * BK_MAX_OWNERS * 4, 256);
	o_tot = tot; tot += 256;
	uint8_t *a = (uint8_t*)bfcg_arena(tot);
	if (!a) return BFCG_ERR_NOMEM;
	EnumParams ep;
	memset(&ep, 0, sizeof(ep));
	ep.k = opt->k, ep.q = opt->q, ep.len = nb, ep.emit_from = 0;
	ep.rec_y0 = (unsigned long long*)(a + o_y0), ep.rec_y1 = (unsigned long long*)(a + o_y1);
	if (host) {
		BFCG_CUDA(cudaMemcpyAsync(a + o_seq, batch->seq, nb, cudaMemcpyHostToDevice, rt.stream));
		if (batch->qual) BFCG_CUDA(cudaMemcpyAsync(a + o_qual, batch->qual, nb, cudaMemcpyHostToDevice, rt.stream));
		ep.seq = a + o_seq, ep.qual = batch->qual ? a + o_qual : 0;
	} else ep.seq = batch->seq, ep.qual = batch->qual;
	{ KTime kt(KT_ENUM); k_enum<<<(unsigned)n_seg, ENUM_THREADS, 0, rt.stream>>>(ep); }
	BFCG_LAUNCH_CHECK();
	BucketParams bp;
	bp.rec_y0 = ep.rec_y0, bp.rec_y1 = ep.rec_y1, bp.k = opt->k, bp.n_shift = opt->bf_shift, bp.owner_bits = owner_bits, bp.n_owners = n_owners;
	bp.seg_cnt = (uint32_t*)(a + o_cnt), bp.out_y0 = (unsigned long long*)d_y0, bp.out_y1 = (unsigned long long*)d_y1;
	bp.totals = (unsigned long long*)(a + o_tot);
	unsigned long long h_tot[BK_MAX_OWNERS], h_start[BK_MAX_OWNERS];
	{
		KTime kt(KT_BUCKET);
		k_bucket<0><<<(unsigned)n_seg, ENUM_THREADS, 0, rt.stream>>>(bp);
		k_bucket_scan<<<1, 32, 0, rt.stream>>>(bp.seg_cnt, n_seg, bp.totals);
	}
	BFCG_LAUNCH_CHECK();
	++rt.n_launches;
	BFCG_CUDA(cudaMemcpyAsync(h_tot, bp.totals, sizeof(h_tot), cudaMemcpyDeviceToHost, rt.stream));
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	unsigned long long run = 0;
	for (int o = 0; o < BK_MAX_OWNERS; ++o) { h_start[o] = run; run += h_tot[o]; }
	BFCG_CUDA(cudaMemcpyAsync(bp.totals, h_start, sizeof(h_start), cudaMemcpyHostToDevice, rt.stream));
	{ KTime kt(KT_BUCKET); k_bucket<1><<<(unsigned)n_seg, ENUM_THREADS, 0, rt.stream>>>(bp); }
	BFCG_LAUNCH_CHECK();
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	for (int o = 0; o < n_owners; ++o) counts[o] = h_tot[o];
	return BFCG_OK;
}

extern "C" int bfcg_count_records(const bfc_opt_t *opt, bfc_bf_t *bf, bfc_bf_t *bf_high, bfc_ch_t *ch, uint64_t n_rec,
                                  const uint64_t *d_y0, const uint64_t *d_y1, int n_owners, bfcg_stats_t *stats)
{
	int r;
	if ((r = bfcg_rt_init()) != BFCG_OK) return r;
	BfcgRuntime &rt = bfcg_rt();
	const int owner_bits = log2_exact(n_owners);
	if (!count_args_ok(opt, bf, bf_high, ch) || owner_bits < 0 || bf->n_shift - BFC_BLK_SHIFT < owner_bits || (n_rec && (!d_y0 || !d_y1)))
		return bfcg_fail(__func__, "invalid arguments", cudaSuccess), BFCG_ERR_ARG;
	if (n_rec == 0) return BFCG_OK;
	if (bfcg_count_part_usable(opt, bf->n_shift, owner_bits)) return bfcg_count_part_records(opt, bf, bf_high, ch, n_rec, d_y0, d_y1, owner_bits, stats);
	if (ch && (r = bfcg_tab_align_to_filter(ch, bf->n_shift - BFC_BLK_SHIFT)) != BFCG_OK) return r; // (settles a shard request)
	const uint64_t sub = std::min<uint64_t>(sub_batch_positions(opt), 1ULL << 31);
	const uint64_t rec_max = std::min<uint64_t>(sub, n_rec);
	CountScratch sc;
	uint8_t *a = (uint8_t*)bfcg_arena(sc.bytes(rec_max));
	if (!a) return BFCG_ERR_NOMEM;
	CountParams p;
	memset(&p, 0, sizeof(p));
	p.k = opt->k, p.linear = 1;
	p.bf = bloom_view(bf);
	p.bf.blk_mask = (1ULL << (bf->n_shift - BFC_BLK_SHIFT - owner_bits)) - 1; // this rank holds 1/n_owners of the blocks
	if (bf_high) { p.bf_high = bloom_view(bf_high); p.bf_high.blk_mask = p.bf.blk_mask; }
	sc.carve(a, rec_max, p);
	BfcgTimer timer(stats);
	for (uint64_t s = 0; s < n_rec; s += sub) {
		p.n_rec = std::min(sub, n_rec - s);
		p.rec_y0 = (const unsigned long long*)d_y0 + s, p.rec_y1 = (const unsigned long long*)d_y1 + s;
		if ((r = count_window(p, sc, bf, ch, p.n_rec, stats)) != BFCG_OK) return r;
	}
	timer.stop();
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	return BFCG_OK;
}

extern "C" int bfcg_count_record_runs(const bfc_opt_t *opt, bfc_bf_t *bf, bfc_bf_t *bf_high, bfc_ch_t *ch, int n_runs, const uint64_t *run_counts,
                                      uint64_t *d_y0, uint64_t *d_y1, int n_owners, bfcg_stats_t *stats)
{
	int r;
	if ((r = bfcg_rt_init()) != BFCG_OK) return r;
	const int owner_bits = log2_exact(n_owners);
	if (!count_args_ok(opt, bf, bf_high, ch) || owner_bits < 0 || bf->n_shift - BFC_BLK_SHIFT < owner_bits || n_runs < 1 || n_runs > BK_MAX_OWNERS || !run_counts)
		return bfcg_fail(__func__, "invalid arguments", cudaSuccess), BFCG_ERR_ARG;
	uint64_t n = 0;
	for (int i = 0; i < n_runs; ++i) n += run_counts[i];
	if (n && (!d_y0 || !d_y1)) return bfcg_fail(__func__, "invalid arguments", cudaSuccess), BFCG_ERR_ARG;
	if (bfcg_count_part_usable(opt, bf->n_shift, owner_bits))
		return bfcg_count_part_runs(opt, bf, bf_high, ch, n_runs, run_counts, d_y0, d_y1, owner_bits, stats);
	return bfcg_count_records(opt, bf, bf_high, ch, n, d_y0, d_y1, n_owners, stats); // pieces in stream order
}
