// dist.cu -- the multi-GPU count phase inside the library: one rank per GPU, k-mers sharded by Bloom-block prefix, the
// exchange done with grouped ncclSend / ncclRecv over NVLink on a stream of its own.
//
// What the reference does in one address space -- every thread inserts into one filter and one table (count.c:54-70,
// kt_for over reads, count.c:106) -- is split by the only thing its ordering depends on, the 64-byte Bloom block
// (bbf.c:25-45): rank r owns the blocks whose index has the top log2(N) bits equal to r, and the table entries of
// exactly those k-mers.  Reads are consumed in global chunks; every rank holds one piece of every chunk.  Per chunk c
//
//   E(c)  enumerate the piece's k-mers and sort them by (owner, count partition): packed records, 9 bytes at k <= 35,
//         the owners' buckets contiguous and each already partitioned                       [engine stream]
//   X(c)  all-to-all of the buckets                                                          [exchange stream, NCCL]
//   C(c)  the Bloom -> table cascade over the N pieces received, in source-rank order = global read order, without
//         sorting again (count_part.cu)                                                      [engine stream]
//
// and the calls are arranged so that X(c) runs while the engine stream does C(c-1) and then E(c+1): two sets of send /
// receive buffers, events between the two streams, no device-wide synchronisation.  The result is that of the
// reference's `-t1` run over the whole input, bit for bit.
//
// After the last chunk the shards are replicated: table entries by an all-gather of (sub-table, slot) pairs, bf_high
// (trim mode) in place.  NCCL is bound at run time (dlopen): the library loads without it for single-GPU use, and
// inside a process that already carries an NCCL (torch) the same one is used.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <algorithm>
#include <vector>

struct NcclApi {
	void *h;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*);
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
	ncclResult_t (*CommDestroy)(ncclComm_t);
	ncclResult_t (*GroupStart)();
	ncclResult_t (*GroupEnd)();
	ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
	ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	const char *(*GetErrorString)(ncclResult_t);
};

static NcclApi g_nccl;

static int nccl_load()
{
	if (g_nccl.h) return BFCG_OK;
	const char *names[] = { getenv("BFC_B200_NCCL"), "libnccl.so.2", "libnccl.so" };
	void *h = 0;
	for (int i = 0; i < 3 && !h; ++i)
		if (names[i] && *names[i]) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
	if (!h) return bfcg_fail(__func__, "libnccl.so.2 not found (needed for multi-GPU counting)", cudaSuccess);
#define NCCL_SYM(field, name) \
	if (!(*(void**)(&g_nccl.field) = dlsym(h, name))) return bfcg_fail(__func__, "symbol missing in libnccl: " name, cudaSuccess)
	NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
	NCCL_SYM(CommInitRank, "ncclCommInitRank");
	NCCL_SYM(CommDestroy, "ncclCommDestroy");
	NCCL_SYM(GroupStart, "ncclGroupStart");
	NCCL_SYM(GroupEnd, "ncclGroupEnd");
	NCCL_SYM(Send, "ncclSend");
	NCCL_SYM(Recv, "ncclRecv");
	NCCL_SYM(AllGather, "ncclAllGather");
	NCCL_SYM(AllReduce, "ncclAllReduce");
	NCCL_SYM(Broadcast, "ncclBroadcast");
	NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef NCCL_SYM
	g_nccl.h = h;
	return BFCG_OK;
}

#define BFCG_NCCL(call)                                                                      \
	do {                                                                                     \
		ncclResult_t r_ = (call);                                                            \
		if (r_ != ncclSuccess) {                                                             \
			snprintf(bfcg_rt().err, sizeof(bfcg_rt().err), "%s: %s: %s", __func__, #call, g_nccl.GetErrorString(r_)); \
			fprintf(stderr, "[E::%s] %s: %s\n", __func__, #call, g_nccl.GetErrorString(r_)); \
			return BFCG_ERR_CUDA;                                                            \
		}                                                                                    \
	} while (0)

#define DIST_MAX 8

struct DistBuf { // one of the two buffer sets of the pipeline
	unsigned long long *skey, *rkey;   // sorted records to send / records received (keys)
	uint8_t *sval, *rval;              // ... and their values (vb bytes each)
	uint64_t scap, rcap;               // capacities in records
	uint8_t *stage_seq, *stage_qual;   // a HOST piece lands here, copied on the copy stream while the engine stream still
	uint64_t stage_cap;                //   runs the cascade of the chunk before (the scratch arena is busy with that)
	cudaEvent_t sorted, received, staged; // engine stream: send buffer complete; exchange stream: receive buffer complete; copy stream: piece on the device
	uint64_t run_counts[DIST_MAX];     // records received from every rank
	bool pending;                      // received (or being received), not counted yet
	int vb;
};

struct BfcgDist {
	int rank, world, owner_bits;
	ncclComm_t comm;
	cudaStream_t xs;                   // the exchange stream
	DistBuf buf[2];
	uint64_t chunk;                    // chunks counted so far
	unsigned long long *d_cnt;         // device: DIST_MAX counts of this rank, then world * DIST_MAX gathered
	unsigned long long *h_cnt;         // pinned host copy
	uint64_t sent_records, recv_records;
	double xchg_ms;                    // device time of the exchanges (events on the exchange stream), when timing is on
	std::vector<std::pair<cudaEvent_t, cudaEvent_t> > xspans;
};

static BfcgDist *g_dist = 0;

static int buf_reserve(DistBuf &b, uint64_t n_send, uint64_t n_recv, int vb)
{
	if (b.vb != vb) { // (another record format: start over)
		cudaFree(b.skey); cudaFree(b.sval); cudaFree(b.rkey); cudaFree(b.rval);
		b.skey = b.rkey = 0, b.sval = b.rval = 0, b.scap = b.rcap = 0, b.vb = vb;
	}
	if (n_send > b.scap) {
		cudaFree(b.skey); cudaFree(b.sval);
		b.scap = n_send + n_send / 8 + 1024;
		if (cudaMalloc(&b.skey, b.scap * 8) != cudaSuccess || cudaMalloc(&b.sval, b.scap * (size_t)vb) != cudaSuccess)
			return bfcg_fail(__func__, "cudaMalloc(send buffer)", cudaErrorMemoryAllocation);
	}
	if (n_recv > b.rcap) {
		cudaFree(b.rkey); cudaFree(b.rval);
		b.rcap = n_recv + n_recv / 8 + 1024;
		if (cudaMalloc(&b.rkey, b.rcap * 8) != cudaSuccess || cudaMalloc(&b.rval, b.rcap * (size_t)vb) != cudaSuccess)
			return bfcg_fail(__func__, "cudaMalloc(receive buffer)", cudaErrorMemoryAllocation);
	}
	return BFCG_OK;
}

// C(c) of a buffer set whose exchange has been issued
static int count_pending(BfcgDist *d, DistBuf &b, const bfc_opt_t *opt, bfc_bf_t *bf, bfc_bf_t *bf_high, bfc_ch_t *ch, bfcg_stats_t *stats)
{
	if (!b.pending) return BFCG_OK;
	BfcgRuntime &rt = bfcg_rt();
	BFCG_CUDA(cudaStreamWaitEvent(rt.stream, b.received, 0));
	b.pending = false;
	return bfcg_count_part_runs_fmt(opt, bf, bf_high, ch, d->world, b.run_counts, b.vb, (const uint64_t*)b.rkey, b.rval, d->owner_bits, stats);
}

extern "C" {

int bfcg_dist_unique_id(void *id)
{
	int r;
	if (!id) return BFCG_ERR_ARG;
	if ((r = nccl_load()) != BFCG_OK) return r;
	ncclUniqueId u;
	BFCG_NCCL(g_nccl.GetUniqueId(&u));
	memset(id, 0, BFCG_DIST_ID_BYTES);
	memcpy(id, &u, sizeof(u) < BFCG_DIST_ID_BYTES ? sizeof(u) : BFCG_DIST_ID_BYTES);
	return BFCG_OK;
}

int bfcg_dist_init(int rank, int world, const void *id)
{
	int r;
	if ((r = bfcg_rt_init()) != BFCG_OK) return r;
	int ob = 0;
	while ((1 << ob) < world) ++ob;
	if (world < 1 || world > DIST_MAX || (1 << ob) != world || rank < 0 || rank >= world || !id)
		return bfcg_fail(__func__, "invalid arguments (ranks must be 1, 2, 4 or 8)", cudaSuccess), BFCG_ERR_ARG;
	if (g_dist) return bfcg_fail(__func__, "already initialised", cudaSuccess), BFCG_ERR_ARG;
	if ((r = nccl_load()) != BFCG_OK) return r;
	BfcgDist *d = new BfcgDist();
	memset(d->buf, 0, sizeof(d->buf));
	d->rank = rank, d->world = world, d->owner_bits = ob, d->chunk = 0, d->sent_records = d->recv_records = 0, d->xchg_ms = 0;
	ncclUniqueId u;
	memcpy(&u, id, sizeof(u) < BFCG_DIST_ID_BYTES ? sizeof(u) : BFCG_DIST_ID_BYTES);
	BFCG_NCCL(g_nccl.CommInitRank(&d->comm, world, u, rank));
	{ // the exchange's kernels are small and must not wait behind the million-CTA grids of the cascade: highest priority
		int least = 0, greatest = 0;
		BFCG_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
		BFCG_CUDA(cudaStreamCreateWithPriority(&d->xs, cudaStreamNonBlocking, greatest));
	}
	for (int i = 0; i < 2; ++i) {
		BFCG_CUDA(cudaEventCreateWithFlags(&d->buf[i].sorted, cudaEventDisableTiming));
		BFCG_CUDA(cudaEventCreateWithFlags(&d->buf[i].received, cudaEventDisableTiming));
		BFCG_CUDA(cudaEventCreateWithFlags(&d->buf[i].staged, cudaEventDisableTiming));
		d->buf[i].vb = -1;
	}
	BFCG_CUDA(cudaMalloc(&d->d_cnt, (DIST_MAX + DIST_MAX * DIST_MAX) * 8));
	BFCG_CUDA(cudaMallocHost(&d->h_cnt, (DIST_MAX + DIST_MAX * DIST_MAX) * 8));
	g_dist = d;
	return BFCG_OK;
}

void bfcg_dist_finalize(void)
{
	BfcgDist *d = g_dist;
	if (!d) return;
	cudaStreamSynchronize(d->xs);
	for (int i = 0; i < 2; ++i) {
		DistBuf &b = d->buf[i];
		cudaFree(b.skey); cudaFree(b.sval); cudaFree(b.rkey); cudaFree(b.rval);
		cudaFree(b.stage_seq); cudaFree(b.stage_qual);
		cudaEventDestroy(b.sorted); cudaEventDestroy(b.received); cudaEventDestroy(b.staged);
	}
	cudaFree(d->d_cnt); cudaFreeHost(d->h_cnt);
	g_nccl.CommDestroy(d->comm);
	cudaStreamDestroy(d->xs);
	delete d;
	g_dist = 0;
}

int bfcg_dist_rank(void) { return g_dist ? g_dist->rank : 0; }
int bfcg_dist_world(void) { return g_dist ? g_dist->world : 1; }

// exchange statistics since initialisation: records sent to other ranks, records received (all ranks), device time of
// the exchanges in ms (needs bfcg_set_timing(1))
int bfcg_dist_stats(uint64_t *sent, uint64_t *received, double *exchange_ms)
{
	BfcgDist *d = g_dist;
	if (!d) return BFCG_ERR_ARG;
	cudaStreamSynchronize(d->xs);
	for (size_t i = 0; i < d->xspans.size(); ++i) {
		float ms = 0;
		if (cudaEventElapsedTime(&ms, d->xspans[i].first, d->xspans[i].second) == cudaSuccess) d->xchg_ms += ms;
		cudaEventDestroy(d->xspans[i].first); cudaEventDestroy(d->xspans[i].second);
	}
	d->xspans.clear();
	if (sent) *sent = d->sent_records;
	if (received) *received = d->recv_records;
	if (exchange_ms) *exchange_ms = d->xchg_ms;
	return BFCG_OK;
}

// Count this rank's piece of the next global chunk: a collective -- every rank calls it once per chunk, with an empty
// piece if it has no reads left.  On return the piece has been enumerated and its exchange is under way; the cascade
// over what this rank received runs during the next call (or in bfcg_dist_count_finish).
int bfcg_dist_count_piece(const bfc_opt_t *opt, bfc_bf_t *bf, bfc_bf_t *bf_high, bfc_ch_t *ch, const bfcg_batch_t *piece, bfcg_stats_t *stats)
{
	int r;
	if ((r = bfcg_rt_init()) != BFCG_OK) return r;
	BfcgDist *d = g_dist;
	BfcgRuntime &rt = bfcg_rt();
	if (!d) return bfcg_fail(__func__, "bfcg_dist_init has not been called", cudaSuccess), BFCG_ERR_ARG;
	if (!opt || !bf || !piece || (ch == 0) == (bf_high == 0) || piece->n_bytes > (1ULL << 31))
		return bfcg_fail(__func__, "invalid arguments (at most 2^31 bytes per piece)", cudaSuccess), BFCG_ERR_ARG;
	if (!bfcg_count_part_usable(opt, opt->bf_shift, d->owner_bits))
		return bfcg_fail(__func__, "this (k, filter size) needs the caller-side exchange (bfcg_enum_records / bfcg_count_record_runs)", cudaSuccess), BFCG_ERR_ARG;
	const int W = d->world, vb = bfcg_part_record_value_bytes(opt->k);
	DistBuf &cur = d->buf[d->chunk & 1], &prev = d->buf[(d->chunk & 1) ^ 1];
	// this set's previous receive buffer must have been counted (it was, one call ago) before X(c) may overwrite it
	if ((r = count_pending(d, cur, opt, bf, bf_high, ch, stats)) != BFCG_OK) return r;

	// E(c): sorted records of the piece into the send buffer; bucket sizes to the host
	uint64_t counts[DIST_MAX];
	memset(counts, 0, sizeof(counts));
	if ((r = buf_reserve(cur, std::max<uint64_t>(piece->n_bytes, 1), 0, vb)) != BFCG_OK) return r;
	bfcg_batch_t staged_piece;
	if (piece->where == BFCG_HOST && piece->n_bytes) {
		// (this set's staging buffers were last read by E(c-2), which the host has waited for)
		if (piece->n_bytes > cur.stage_cap) {
			cudaFree(cur.stage_seq); cudaFree(cur.stage_qual);
			cur.stage_seq = cur.stage_qual = 0;
			cur.stage_cap = piece->n_bytes + piece->n_bytes / 8;
			if (cudaMalloc(&cur.stage_seq, cur.stage_cap) != cudaSuccess || cudaMalloc(&cur.stage_qual, cur.stage_cap) != cudaSuccess) {
				cur.stage_cap = 0;
				return bfcg_fail(__func__, "cudaMalloc(piece staging)", cudaErrorMemoryAllocation);
			}
		}
		BFCG_CUDA(cudaMemcpyAsync(cur.stage_seq, piece->seq, piece->n_bytes, cudaMemcpyHostToDevice, rt.copy_in));
		if (piece->qual) BFCG_CUDA(cudaMemcpyAsync(cur.stage_qual, piece->qual, piece->n_bytes, cudaMemcpyHostToDevice, rt.copy_in));
		BFCG_CUDA(cudaEventRecord(cur.staged, rt.copy_in));
		BFCG_CUDA(cudaStreamWaitEvent(rt.stream, cur.staged, 0));
		staged_piece = *piece;
		staged_piece.where = BFCG_DEVICE, staged_piece.seq = cur.stage_seq, staged_piece.qual = piece->qual ? cur.stage_qual : 0;
		staged_piece.off = 0; // (the enumeration goes by the terminators, not by offsets)
		piece = &staged_piece;
	}
	if (piece->n_bytes && (r = bfcg_enum_part_records_fmt(opt, piece, d->owner_bits, vb, (uint64_t*)cur.skey, cur.sval, counts)) != BFCG_OK) return r;
	BFCG_CUDA(cudaEventRecord(cur.sorted, rt.stream));

	// bucket sizes of every rank (tiny all-gather, queued behind X(c-1) on the exchange stream)
	for (int i = 0; i < DIST_MAX; ++i) d->h_cnt[i] = i < W ? counts[i] : 0;
	BFCG_CUDA(cudaMemcpyAsync(d->d_cnt, d->h_cnt, DIST_MAX * 8, cudaMemcpyHostToDevice, d->xs));
	BFCG_NCCL(g_nccl.AllGather(d->d_cnt, d->d_cnt + DIST_MAX, DIST_MAX, ncclUint64, d->comm, d->xs));
	BFCG_CUDA(cudaMemcpyAsync(d->h_cnt + DIST_MAX, d->d_cnt + DIST_MAX, (size_t)W * DIST_MAX * 8, cudaMemcpyDeviceToHost, d->xs));
	BFCG_CUDA(cudaStreamSynchronize(d->xs));
	uint64_t n_recv = 0, n_send = 0;
	for (int s = 0; s < W; ++s) cur.run_counts[s] = d->h_cnt[DIST_MAX + s * DIST_MAX + d->rank], n_recv += cur.run_counts[s];
	for (int t = 0; t < W; ++t) n_send += counts[t];
	if (n_recv >= (1ULL << 32)) return bfcg_fail(__func__, "more than 2^32 records for one rank in one chunk: use smaller chunks", cudaSuccess), BFCG_ERR_ARG;
	if ((r = buf_reserve(cur, 0, std::max<uint64_t>(n_recv, 1), vb)) != BFCG_OK) return r;

	// X(c): every bucket to its owner, the pieces arriving in source-rank order
	BFCG_CUDA(cudaStreamWaitEvent(d->xs, cur.sorted, 0));
	cudaEvent_t t0 = 0, t1 = 0;
	if (rt.timing) { cudaEventCreate(&t0); cudaEventCreate(&t1); cudaEventRecord(t0, d->xs); }
	BFCG_NCCL(g_nccl.GroupStart());
	uint64_t so = 0, ro = 0;
	for (int p = 0; p < W; ++p) {
		if (counts[p]) {
			BFCG_NCCL(g_nccl.Send(cur.skey + so, counts[p] * 8, ncclUint8, p, d->comm, d->xs));
			BFCG_NCCL(g_nccl.Send(cur.sval + so * vb, counts[p] * (size_t)vb, ncclUint8, p, d->comm, d->xs));
		}
		if (cur.run_counts[p]) {
			BFCG_NCCL(g_nccl.Recv(cur.rkey + ro, cur.run_counts[p] * 8, ncclUint8, p, d->comm, d->xs));
			BFCG_NCCL(g_nccl.Recv(cur.rval + ro * vb, cur.run_counts[p] * (size_t)vb, ncclUint8, p, d->comm, d->xs));
		}
		so += counts[p], ro += cur.run_counts[p];
	}
	BFCG_NCCL(g_nccl.GroupEnd());
	if (rt.timing) { cudaEventRecord(t1, d->xs); d->xspans.push_back(std::make_pair(t0, t1)); }
	BFCG_CUDA(cudaEventRecord(cur.received, d->xs));
	cur.pending = true;
	d->sent_records += n_send - counts[d->rank], d->recv_records += n_recv;
	++d->chunk;

	// C(c-1) on the engine stream while X(c) travels
	return count_pending(d, prev, opt, bf, bf_high, ch, stats);
}

// the cascade over whatever has been received and not counted yet; afterwards the shards are complete
int bfcg_dist_count_finish(const bfc_opt_t *opt, bfc_bf_t *bf, bfc_bf_t *bf_high, bfc_ch_t *ch, bfcg_stats_t *stats)
{
	int r;
	if ((r = bfcg_rt_init()) != BFCG_OK) return r;
	BfcgDist *d = g_dist;
	if (!d) return bfcg_fail(__func__, "bfcg_dist_init has not been called", cudaSuccess), BFCG_ERR_ARG;
	// in chunk order: the set used by the older chunk first
	DistBuf &older = d->buf[d->chunk & 1], &newer = d->buf[(d->chunk & 1) ^ 1];
	if ((r = count_pending(d, older, opt, bf, bf_high, ch, stats)) != BFCG_OK) return r;
	if ((r = count_pending(d, newer, opt, bf, bf_high, ch, stats)) != BFCG_OK) return r;
	BFCG_CUDA(cudaStreamSynchronize(bfcg_rt().stream));
	return BFCG_OK;
}

// Replicate the table: every rank's entries as (sub-table, slot) pairs, broadcast in turn, imported into `full` (an
// ordinary table, empty on entry or holding an earlier gather's keys cleared by the caller).
int bfcg_dist_gather_table(const bfc_ch_t *shard, bfc_ch_t *full)
{
	int r;
	if ((r = bfcg_rt_init()) != BFCG_OK) return r;
	BfcgDist *d = g_dist;
	BfcgRuntime &rt = bfcg_rt();
	if (!d || !shard || !full) return bfcg_fail(__func__, "invalid arguments", cudaSuccess), BFCG_ERR_ARG;
	const int W = d->world;
	if (shard->own_bits == d->owner_bits && full->own_bits == 0 && full != shard && !getenv("BFC_B200_GATHER_ENTRIES")) {
		// The shards' slot arrays, concatenated in owner order, ARE the whole table (bfcg_tab_shape_like_shards) once every
		// shard has the same region size: one all-gather over NVLink, no export, no re-insertion.
		bfc_ch_s *mine = (bfc_ch_s*)shard;
		uint64_t rb = (uint64_t)mine->rbits;
		{ // the largest region size of any shard
			double v = (double)rb;
			if ((r = bfcg_dist_allreduce_max_f64(&v, 1)) != BFCG_OK) return r;
			rb = (uint64_t)v;
		}
		if ((r = bfcg_tab_set_rbits(mine, (int)rb)) != BFCG_OK) return r;
		if ((r = bfcg_tab_shape_like_shards(full, mine)) != BFCG_OK) return r;
		uint64_t cnt = bfc_ch_count(mine);
		if ((r = bfcg_dist_allreduce_sum_u64(&cnt, 1)) != BFCG_OK) return r;
		BFCG_CUDA(cudaStreamSynchronize(rt.stream));
		BFCG_NCCL(g_nccl.AllGather(mine->slots, full->slots, bfcg_tab_capacity(mine) * 8, ncclUint8, d->comm, d->xs));
		BFCG_CUDA(cudaStreamSynchronize(d->xs));
		return bfcg_tab_set_count(full, cnt);
	}
	const uint64_t n_mine = bfcg_ch_export_device(shard, 0, 0);
	// sizes of every shard
	d->h_cnt[0] = n_mine;
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	BFCG_CUDA(cudaMemcpyAsync(d->d_cnt, d->h_cnt, 8, cudaMemcpyHostToDevice, d->xs));
	BFCG_NCCL(g_nccl.AllGather(d->d_cnt, d->d_cnt + DIST_MAX, 1, ncclUint64, d->comm, d->xs));
	BFCG_CUDA(cudaMemcpyAsync(d->h_cnt + DIST_MAX, d->d_cnt + DIST_MAX, (size_t)W * 8, cudaMemcpyDeviceToHost, d->xs));
	BFCG_CUDA(cudaStreamSynchronize(d->xs));
	uint64_t total = 0, off[DIST_MAX + 1];
	for (int p = 0; p < W; ++p) off[p] = total, total += d->h_cnt[DIST_MAX + p];
	off[W] = total;
	if (total == 0) return BFCG_OK;
	uint32_t *g_sub = 0;
	unsigned long long *g_key = 0;
	if (cudaMalloc(&g_sub, total * 4) != cudaSuccess || cudaMalloc(&g_key, total * 8) != cudaSuccess) {
		cudaFree(g_sub); cudaFree(g_key);
		return bfcg_fail(__func__, "cudaMalloc(gathered table entries)", cudaErrorMemoryAllocation);
	}
	if (n_mine && bfcg_ch_export_device(shard, g_sub + off[d->rank], (uint64_t*)g_key + off[d->rank]) != n_mine) { cudaFree(g_sub); cudaFree(g_key); return BFCG_ERR_CUDA; }
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	BFCG_NCCL(g_nccl.GroupStart());
	for (int p = 0; p < W; ++p) {
		const uint64_t n = off[p + 1] - off[p];
		if (n == 0) continue;
		BFCG_NCCL(g_nccl.Broadcast(g_sub + off[p], g_sub + off[p], n * 4, ncclUint8, p, d->comm, d->xs));
		BFCG_NCCL(g_nccl.Broadcast(g_key + off[p], g_key + off[p], n * 8, ncclUint8, p, d->comm, d->xs));
	}
	BFCG_NCCL(g_nccl.GroupEnd());
	BFCG_CUDA(cudaStreamSynchronize(d->xs));
	if ((r = bfcg_ch_reserve(full, total)) == BFCG_OK) r = bfcg_ch_import_device(full, total, g_sub, (const uint64_t*)g_key);
	cudaFree(g_sub); cudaFree(g_key);
	return r;
}

// Replicate bf_high (trim mode): `full` is an ordinary filter of the whole size; every rank's shard lands at its place
int bfcg_dist_gather_filter(const bfc_bf_t *shard, bfc_bf_t *full)
{
	int r;
	if ((r = bfcg_rt_init()) != BFCG_OK) return r;
	BfcgDist *d = g_dist;
	if (!d || !shard || !full || shard->n_shift != full->n_shift) return bfcg_fail(__func__, "invalid arguments", cudaSuccess), BFCG_ERR_ARG;
	const size_t bytes = ((size_t)1 << (full->n_shift - 3)) / d->world;
	BFCG_CUDA(cudaStreamSynchronize(bfcg_rt().stream));
	BFCG_NCCL(g_nccl.AllGather(shard->b, full->b, bytes, ncclUint8, d->comm, d->xs));
	BFCG_CUDA(cudaStreamSynchronize(d->xs));
	return BFCG_OK;
}

int bfcg_dist_allreduce_sum_u64(uint64_t *v, int n)
{
	BfcgDist *d = g_dist;
	if (!d || n < 0 || n > DIST_MAX * DIST_MAX) return BFCG_ERR_ARG;
	BFCG_CUDA(cudaStreamSynchronize(bfcg_rt().stream));
	memcpy(d->h_cnt, v, (size_t)n * 8);
	BFCG_CUDA(cudaMemcpyAsync(d->d_cnt, d->h_cnt, (size_t)n * 8, cudaMemcpyHostToDevice, d->xs));
	BFCG_NCCL(g_nccl.AllReduce(d->d_cnt, d->d_cnt, n, ncclUint64, ncclSum, d->comm, d->xs));
	BFCG_CUDA(cudaMemcpyAsync(d->h_cnt, d->d_cnt, (size_t)n * 8, cudaMemcpyDeviceToHost, d->xs));
	BFCG_CUDA(cudaStreamSynchronize(d->xs));
	memcpy(v, d->h_cnt, (size_t)n * 8);
	return BFCG_OK;
}

int bfcg_dist_allreduce_max_f64(double *v, int n)
{
	BfcgDist *d = g_dist;
	if (!d || n < 0 || n > DIST_MAX * DIST_MAX) return BFCG_ERR_ARG;
	BFCG_CUDA(cudaStreamSynchronize(bfcg_rt().stream));
	memcpy(d->h_cnt, v, (size_t)n * 8);
	BFCG_CUDA(cudaMemcpyAsync(d->d_cnt, d->h_cnt, (size_t)n * 8, cudaMemcpyHostToDevice, d->xs));
	BFCG_NCCL(g_nccl.AllReduce(d->d_cnt, d->d_cnt, n, ncclDouble, ncclMax, d->comm, d->xs));
	BFCG_CUDA(cudaMemcpyAsync(d->h_cnt, d->d_cnt, (size_t)n * 8, cudaMemcpyDeviceToHost, d->xs));
	BFCG_CUDA(cudaStreamSynchronize(d->xs));
	memcpy(v, d->h_cnt, (size_t)n * 8);
	return BFCG_OK;
}

int bfcg_dist_barrier(void)
{
	uint64_t one = 1;
	return bfcg_dist_allreduce_sum_u64(&one, 1);
}

} // extern "C"
