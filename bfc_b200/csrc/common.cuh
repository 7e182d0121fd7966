// common.cuh -- shared by every translation unit of libbfc_b200 (sm_100a only).
//
// Device-side views of the two HBM-resident structures of the count/correct path
// (blocked Bloom filter, counting table) with the inline device functions that
// touch them, plus the host-side runtime singleton (stream, scratch arena, error
// string, kernel timing).  No -rdc: everything device-side here is inline.
#pragma once

#include <cuda_runtime.h>
#include <vector>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/bfc_b200.h"

// ------------------------------------------------------------------ host runtime

struct BfcgRuntime {
	int dev;
	bool ready;
	cudaStream_t stream;
	cudaEvent_t ev0, ev1;
	bool timing;
	uint64_t n_launches;
	char err[512];
	// grow-only device scratch arena (one allocation, carved per call)
	uint8_t *arena;
	size_t arena_bytes;
	// grow-only pinned staging buffers for host batches
	uint8_t *pin[2];
	size_t pin_bytes[2];
	int sm_count;
	// per-kernel device timing (only when `timing`): event pairs resolved lazily
	struct Span { int id; cudaEvent_t a, b; };
	std::vector<Span> spans;
	std::vector<cudaEvent_t> ev_pool;
	double kt_ms[24];
	uint64_t kt_n[24];
	cudaEvent_t user_ev[8];
	// host <-> device copies of host batches run on their own streams so that they overlap the kernels
	cudaStream_t copy_in, copy_out;
	cudaEvent_t ev_in[3], ev_free[3], ev_done[3], ev_out[3];
};

enum { KT_COUNT_PROBE = 0, KT_COUNT_RESOLVE, KT_COUNT_SORT, KT_COUNT_REPLAY, KT_CORRECT, KT_CORRECT_REDO,
       KT_TRIM, KT_TAB_REHASH, KT_TAB_HIST, KT_TAB_APPLY, KT_ENUM, KT_EC_LOOKUP, KT_EC_SETUP, KT_EC_MERGE, KT_BUCKET,
       KT_COUNT_PART, KT_COUNT_BOUNDS, KT_ENUM_LIN, KT_EC_EXT, KT_N };

int  bfcg_kt_begin(int id);   // records a start event on the stream when timing is on; returns a span index or -1
void bfcg_kt_end(int idx);
struct KTime { // RAII: time the launches enqueued while it is alive
	int idx;
	explicit KTime(int id) : idx(bfcg_kt_begin(id)) {}
	~KTime() { bfcg_kt_end(idx); }
};

BfcgRuntime &bfcg_rt();
int bfcg_rt_init();                       // lazily selects the device, creates the stream; BFCG_OK or error
int bfcg_fail(const char *func, const char *what, cudaError_t e);
void *bfcg_arena(size_t bytes);           // returns device scratch of at least `bytes` (contents undefined)
uint8_t *bfcg_pinned(int which, size_t bytes);

#define BFCG_CUDA(call)                                                        \
	do {                                                                       \
		cudaError_t e_ = (call);                                               \
		if (e_ != cudaSuccess) return bfcg_fail(__func__, #call, e_);          \
	} while (0)

#define BFCG_LAUNCH_CHECK()                                                    \
	do {                                                                       \
		++bfcg_rt().n_launches;                                                \
		cudaError_t e_ = cudaGetLastError();                                   \
		if (e_ != cudaSuccess) return bfcg_fail(__func__, "kernel launch", e_);\
	} while (0)

struct BfcgTimer { // accumulates device time of the enclosed stream work into stats->kernel_ms
	bfcg_stats_t *st;
	bool on;
	uint64_t launches0;
	explicit BfcgTimer(bfcg_stats_t *s) : st(s), on(bfcg_rt().timing && s), launches0(bfcg_rt().n_launches) {
		if (on) cudaEventRecord(bfcg_rt().ev0, bfcg_rt().stream);
	}
	void stop() { // call after the work is enqueued; synchronises the stream
		if (on) {
			float ms = 0;
			cudaEventRecord(bfcg_rt().ev1, bfcg_rt().stream);
			cudaEventSynchronize(bfcg_rt().ev1);
			cudaEventElapsedTime(&ms, bfcg_rt().ev0, bfcg_rt().ev1);
			st->kernel_ms += ms;
		}
		if (st) st->n_launches += bfcg_rt().n_launches - launches0;
	}
};

// ------------------------------------------------------------------ host structs

struct bfc_ch_s {
	int k, l_pre;             // l_pre after the adjustment of reference htab.c:24-26
	int rbits;                // log2(slots per region); capacity = 2^(l_pre + rbits) slots
	int rot;                  // region of sub-table s = s rotated right by rot inside l_pre bits (tab_region)
	unsigned long long *slots;
	unsigned long long *counters; // device, 8 words: [0] n_entries, [1] n_deferred, [2] rehash failures
	unsigned long long *deferred; // device: 2 words per deferred insert (y0 | is_high<<63, y1)
	uint64_t def_cap;
	uint64_t prev_new;        // distinct keys added by the previous count window (growth estimate of the next one)
	int have_prev;
	// One shard of N (sharded counting): the owner bits of a k-mer are the top bits of its Bloom
	// block index, which tab_region turns into the TOP own_bits bits of the region index -- so a shard only ever touches
	// the regions whose top bits equal own_val, and only those 2^(l_pre - own_bits) regions are allocated.  When the
	// geometry does not allow that (block index not a bit field of the sub-table index) the shard keeps every region and
	// `skew` = N tells the growth policy that 1/N of them take all the keys.
	int req_owners, req_owner; // what bfcg_ch_set_shard asked for (applied by bfcg_tab_align_to_filter, table empty)
	int own_bits;
	uint32_t own_val;
	int skew;
};

// ------------------------------------------------------------------ device views

struct BloomView {
	uint32_t *w;      // 16 words per 64-byte block
	int n_shift, n_hashes;
	uint64_t blk_mask; // block index -> index inside this allocation (all ones unless the filter is one shard of N)
};

struct TabView {
	unsigned long long *slots;
	unsigned long long *counters;
	unsigned long long *deferred;
	unsigned long long def_cap;
	int k, l_pre, rbits, rot;
	uint32_t rmask;   // region index bits kept by this table (all l_pre of them unless it is a shard)
};

static inline BloomView bloom_view(const bfc_bf_t *b)
{
	BloomView v;
	v.w = (uint32_t*)b->b, v.n_shift = b->n_shift, v.n_hashes = b->n_hashes, v.blk_mask = ~0ULL;
	return v;
}

static inline TabView tab_view(const bfc_ch_s *c)
{
	TabView v;
	v.slots = c->slots, v.counters = c->counters, v.deferred = c->deferred, v.def_cap = c->def_cap;
	v.k = c->k, v.l_pre = c->l_pre, v.rbits = c->rbits, v.rot = c->rot;
	v.rmask = (1u << (c->l_pre - c->own_bits)) - 1;
	return v;
}

// table growth policy (htab.cu): make room for `extra` more distinct keys at load <= 1/2
int bfcg_tab_reserve(bfc_ch_s *ch, uint64_t extra);
// place the sub-tables so that k-mers of neighbouring Bloom blocks (2^x blocks) sit in neighbouring regions
int bfcg_tab_align_to_filter(bfc_ch_s *ch, int x);
// re-apply inserts that found their region full (after growing); htab.cu
int bfcg_tab_drain_deferred(bfc_ch_s *ch);
// double the slots of every region
int bfcg_tab_grow(bfc_ch_s *ch);
// table replication by concatenating shard slot arrays (dist.cu)
int bfcg_tab_set_rbits(bfc_ch_s *ch, int rbits);
int bfcg_tab_shape_like_shards(bfc_ch_s *full, const bfc_ch_s *shard);
int bfcg_tab_set_count(bfc_ch_s *ch, uint64_t n);
// slots of the table (a shard holds 2^(l_pre - own_bits) regions)
static inline uint64_t bfcg_tab_capacity(const bfc_ch_s *ch) { return 1ULL << (ch->l_pre - ch->own_bits + ch->rbits); }

// the partitioned count path (count_part.cu); count.cu dispatches to it when it applies
bool bfcg_count_part_usable(const bfc_opt_t *opt, int n_shift, int owner_bits);
int bfcg_count_part_batch(const bfc_opt_t *opt, bfc_bf_t *bf, bfc_bf_t *bf_high, bfc_ch_t *ch, const bfcg_batch_t *batch, bfcg_stats_t *stats);
int bfcg_count_part_records(const bfc_opt_t *opt, bfc_bf_t *bf, bfc_bf_t *bf_high, bfc_ch_t *ch, uint64_t n_rec,
                            const uint64_t *d_y0, const uint64_t *d_y1, int owner_bits, bfcg_stats_t *stats);
int bfcg_enum_part_records(const bfc_opt_t *opt, const bfcg_batch_t *batch, int owner_bits, uint64_t *d_y0, uint64_t *d_y1, uint64_t *counts);
int bfcg_count_part_runs(const bfc_opt_t *opt, bfc_bf_t *bf, bfc_bf_t *bf_high, bfc_ch_t *ch, int n_runs, const uint64_t *run_counts,
                         const uint64_t *d_y0, uint64_t *d_y1, int owner_bits, bfcg_stats_t *stats);
// the same two with the record format given (0 = 16-byte wire records, else bytes of a packed record's value)
int bfcg_part_record_value_bytes(int k);
int bfcg_enum_part_records_fmt(const bfc_opt_t *opt, const bfcg_batch_t *batch, int owner_bits, int vb, uint64_t *d_y0, void *d_y1, uint64_t *counts);
int bfcg_count_part_runs_fmt(const bfc_opt_t *opt, bfc_bf_t *bf, bfc_bf_t *bf_high, bfc_ch_t *ch, int n_runs, const uint64_t *run_counts,
                             int vb, const uint64_t *d_y0, void *d_y1, int owner_bits, bfcg_stats_t *stats);

#ifdef __CUDACC__

// ------------------------------------------------------------------ base codes

// reference bseq.c:9-26 minus one: A/a 0, C/c 1, G/g 2, T/t 3, anything else 4
__device__ __forceinline__ int base_code(uint8_t ch)
{
	const uint32_t u = ch & 0xDFu; // fold case
	return u == 'A' ? 0 : u == 'C' ? 1 : u == 'G' ? 2 : u == 'T' ? 3 : 4;
}

// the 64-bit Bloom hash of a record (y0, y1): inverse of the last two lines of
// bfc_kmer_hash (kmer.h:85-86): h1 = y1, h0 = y0 - y1
__device__ __forceinline__ uint64_t hash_from_y(int k, uint64_t y0, uint64_t y1)
{
	const uint64_t m = (1ULL << k) - 1, h0 = (y0 - y1) & m;
	return ((h0 ^ y1) << k) | y0;
}

// ------------------------------------------------------------------ Bloom probes

struct BloomProbe {
	uint64_t blk; // block index
	int h1, h2;
};

// reference bbf.c:27-33
__device__ __forceinline__ BloomProbe bloom_locate(uint64_t hash, int n_shift)
{
	BloomProbe p;
	const int x = n_shift - BFC_BLK_SHIFT;
	p.blk = hash & ((1ULL << x) - 1);
	p.h1 = (int)(hash >> x) & BFC_BLK_MASK;
	p.h2 = (int)(hash >> n_shift) & BFC_BLK_MASK;
	if ((p.h2 & 31) == 0) p.h2 = (p.h2 + 1) & BFC_BLK_MASK;
	return p;
}

// the 16 words of a block inside this view's allocation
__device__ __forceinline__ uint32_t *bloom_block(const BloomView &v, uint64_t blk) { return v.w + ((blk & v.blk_mask) << 4); }

// number of probe bits set in the block at `w` (16 words); the reference's probe walk
// (bbf.c:35-42 / 54-61): positions < 8 belong to the lock byte and do not count.
template <bool CG>
__device__ __forceinline__ int bloom_count_set(const uint32_t *w, const BloomProbe &p, int n_hashes)
{
	int z = p.h1, done = 0, cnt = 0;
	while (done < n_hashes) {
		if (z >= 8) {
			const uint32_t v = CG ? __ldcg(w + (z >> 5)) : w[z >> 5];
			cnt += (v >> (z & 31)) & 1;
			++done;
		}
		z = (z + p.h2) & BFC_BLK_MASK;
	}
	return cnt;
}

// set all probe bits with L2 atomics (order-free: used where only the union matters)
__device__ __forceinline__ void bloom_set_atomic(uint32_t *w, const BloomProbe &p, int n_hashes)
{
	int z = p.h1, done = 0;
	while (done < n_hashes) {
		if (z >= 8) {
			atomicOr(w + (z >> 5), 1u << (z & 31));
			++done;
		}
		z = (z + p.h2) & BFC_BLK_MASK;
	}
}

// ------------------------------------------------------------------ counting table

// reference htab.c:45-58: (sub-table index, 50-bit key) of a hashed k-mer
__device__ __forceinline__ void tab_subkey(int k, int l_pre, uint64_t y0, uint64_t y1, uint32_t &sub, uint64_t &key)
{
	if (k <= 32) {
		const int t = 2 * k - l_pre;
		const uint64_t z = y0 << k | y1;
		key = z & ((1ULL << t) - 1);
		sub = (uint32_t)(z >> t);
	} else {
		const int t = k - l_pre;
		const int shift = t + k < BFC_CH_KEYBITS ? k : BFC_CH_KEYBITS - t;
		key = (((y0 & ((1ULL << t) - 1)) << shift) ^ y1) & ((1ULL << BFC_CH_KEYBITS) - 1);
		sub = (uint32_t)(y0 >> t);
	}
}

// Where sub-table `sub` lives: its index rotated right by t.rot inside l_pre bits.  With rot = (number of Bloom
// block-index bits that are also sub-table bits) the low y0 bits a count partition fixes become the TOP bits of
// the region index, so one partition's upserts fall into one contiguous stretch of the table (count_part.cu).
__host__ __device__ __forceinline__ uint32_t tab_region(int l_pre, int rot, uint32_t sub)
{
	return rot ? ((sub >> rot) | (sub << (l_pre - rot))) & ((1u << l_pre) - 1) : sub;
}
__host__ __device__ __forceinline__ uint32_t tab_region_inv(int l_pre, int rot, uint32_t reg)
{
	return rot ? ((reg << rot) | (reg >> (l_pre - rot))) & ((1u << l_pre) - 1) : reg;
}

// first slot of the region of sub-table `sub` inside this table's allocation
__device__ __forceinline__ unsigned long long *tab_region_ptr(const TabView &t, uint32_t sub)
{
	return t.slots + ((uint64_t)(tab_region(t.l_pre, t.rot, sub) & t.rmask) << t.rbits);
}

__device__ __forceinline__ uint64_t tab_mix(uint64_t key)
{
	key ^= key >> 29;
	key *= 0xBF58476D1CE4E5B9ULL;
	key ^= key >> 32;
	key *= 0x94D049BB133111EBULL;
	key ^= key >> 29;
	return key;
}

// Slot = key50 << 14 | high6 << 8 | cnt8; 0 = empty (cnt8 >= 1 for every stored key).
// Probing: 4-slot (32-byte, one DRAM sector) buckets, linear over buckets inside the
// key's region.  Returns 1 = new key stored, 0 = existing key updated, -1 = region
// full (insert parked in the deferred list and re-applied after the table has grown).
template <bool PARK>
__device__ __forceinline__ int tab_upsert_t(const TabView &t, uint64_t y0, uint64_t y1, int is_high)
{
	uint32_t sub; uint64_t key;
	tab_subkey(t.k, t.l_pre, y0, y1, sub, key);
	const uint64_t R = 1ULL << t.rbits;
	unsigned long long *reg = tab_region_ptr(t, sub);
	uint64_t h = tab_mix(key) & (R - 1) & ~3ULL;
	const unsigned long long fresh = key << 14 | (unsigned long long)(is_high ? 1 : 0) << 8 | 1ULL;
	for (uint64_t n = 0; n < R; n += 4, h = (h + 4) & (R - 1)) {
		unsigned long long *b = reg + h;
		const ulonglong2 v01 = __ldcg((const ulonglong2*)b);
		const ulonglong2 v23 = __ldcg((const ulonglong2*)b + 1);
		unsigned long long v[4] = { v01.x, v01.y, v23.x, v23.y };
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			unsigned long long s = v[j];
			if (s == 0) {
				const unsigned long long prev = atomicCAS(b + j, 0ULL, fresh);
				if (prev == 0) return 1;
				s = prev;
			}
			if ((s >> 14) == key) {
				for (;;) { // saturating increments (reference htab.c:76-79)
					unsigned long long ns = s;
					if ((s & 0xff) != 0xff) ++ns;
					if (is_high && ((s >> 8) & 0x3f) != 0x3f) ns += 1 << 8;
					if (ns == s) return 0;
					const unsigned long long prev = atomicCAS(b + j, s, ns);
					if (prev == s) return 0;
					s = prev;
				}
			}
		}
	}
	if (PARK) {
		const unsigned long long idx = atomicAdd(t.counters + 1, 1ULL);
		if (idx < t.def_cap) {
			t.deferred[2 * idx] = y0 | (unsigned long long)(is_high ? 1 : 0) << 63;
			t.deferred[2 * idx + 1] = y1;
		}
	}
	return -1;
}
__device__ __forceinline__ int tab_upsert(const TabView &t, uint64_t y0, uint64_t y1, int is_high) { return tab_upsert_t<true>(t, y0, y1, is_high); }

// store a ready-made slot value (restore / rehash); the key must not be present yet
__device__ __forceinline__ bool tab_put_raw(const TabView &t, uint32_t sub, unsigned long long slot)
{
	const uint64_t R = 1ULL << t.rbits;
	unsigned long long *reg = tab_region_ptr(t, sub);
	uint64_t h = tab_mix(slot >> 14) & (R - 1) & ~3ULL;
	for (uint64_t n = 0; n < R; ++n) {
		const uint64_t i = (h + n) & (R - 1);
		if (__ldcg(reg + i) == 0 && atomicCAS(reg + i, 0ULL, slot) == 0) return true;
	}
	return false;
}

// reference htab.c:84-92: -1 = absent, else the low 14 bits.  Read-only phase.
__device__ __forceinline__ int tab_get(const TabView &t, uint64_t y0, uint64_t y1)
{
	uint32_t sub; uint64_t key;
	tab_subkey(t.k, t.l_pre, y0, y1, sub, key);
	const uint64_t R = 1ULL << t.rbits;
	const unsigned long long *reg = tab_region_ptr(t, sub);
	uint64_t h = tab_mix(key) & (R - 1) & ~3ULL;
	for (uint64_t n = 0; n < R; n += 4, h = (h + 4) & (R - 1)) {
		const ulonglong2 v01 = __ldg((const ulonglong2*)(reg + h));
		const ulonglong2 v23 = __ldg((const ulonglong2*)(reg + h) + 1);
		const unsigned long long v[4] = { v01.x, v01.y, v23.x, v23.y };
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			if (v[j] == 0) return -1;
			if ((v[j] >> 14) == key) return (int)(v[j] & 0x3fff);
		}
	}
	return -1;
}

// ---- the same lookup in two halves, for callers that keep several of them in flight: tab_locate (where the key's
// first bucket is), the caller's own loads of that bucket, tab_match_bucket (-2 = bucket full of other keys: go on
// with tab_get_rest from the next bucket)
__device__ __forceinline__ const unsigned long long *tab_locate(const TabView &t, uint64_t y0, uint64_t y1, uint64_t &key, uint32_t &h)
{
	uint32_t sub;
	tab_subkey(t.k, t.l_pre, y0, y1, sub, key);
	h = (uint32_t)(tab_mix(key) & ((1ULL << t.rbits) - 1) & ~3ULL);
	return tab_region_ptr(t, sub);
}

__device__ __forceinline__ int tab_match_bucket(const ulonglong2 &v01, const ulonglong2 &v23, uint64_t key)
{
	const unsigned long long v[4] = { v01.x, v01.y, v23.x, v23.y };
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		if (v[j] == 0) return -1;
		if ((v[j] >> 14) == key) return (int)(v[j] & 0x3fff);
	}
	return -2;
}

static __device__ __noinline__ int tab_get_rest(const TabView &t, const unsigned long long *reg, uint32_t h0, uint64_t key)
{
	const uint64_t R = 1ULL << t.rbits;
	uint64_t h = ((uint64_t)h0 + 4) & (R - 1);
	for (uint64_t n = 4; n < R; n += 4, h = (h + 4) & (R - 1)) {
		const int r = tab_match_bucket(__ldg((const ulonglong2*)(reg + h)), __ldg((const ulonglong2*)(reg + h) + 1), key);
		if (r != -2) return r;
	}
	return -1;
}

// reference htab.c:94-99
__device__ __forceinline__ int tab_kmer_occ(const TabView &t, const uint64_t x[4])
{
	uint64_t y[2];
	bfc_kmer_hash(t.k, x, y);
	return tab_get(t, y[0], y[1]);
}

// block-wide sum of a per-thread counter, one atomic per CTA
__device__ __forceinline__ void block_add(unsigned long long *dst, unsigned long long v)
{
	__shared__ unsigned long long s_acc;
	if (threadIdx.x == 0) s_acc = 0;
	__syncthreads();
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_acc, v);
	__syncthreads();
	if (threadIdx.x == 0 && s_acc) atomicAdd(dst, s_acc);
	__syncthreads(); // calls follow each other: the (speculated, predicated-off) reads of s_acc above must not meet the next call's reset
}

#endif // __CUDACC__
