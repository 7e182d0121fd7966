// count_part.cu -- the count phase as a PARTITIONED pass: the Bloom -> table cascade of
// bfc_kmer_insert (reference count.c:54-70) over the k-mers of worker_count (count.c:72-89)
// with the sequential (-t1) semantics, arranged so that every random access lands in
// shared memory or in a small window of L2 instead of HBM.
//
// Every ordering constraint of the cascade is local to one 64-byte Bloom block
// (bbf.c:35-44: occurrence t passes iff occurrences < t of the same block set all its
// bits).  A window of up to 2^30 stream positions is therefore processed as
//
//   K0' k_enum_lin     canonical k-mer hash of every stream position, from bit planes of
//                      the window staged in shared memory (no rolling state, no warm-up):
//                      16-byte records (y0 | is_high << 63, y1) in STREAM order
//   K1' radix sort     stable partition of the records by the top bits of their Bloom
//                      block index (cub onesweep over bits [10, x) of y0): partition =
//                      1024 consecutive blocks = 64 KB of the filter
//   K2' k_part_bounds  first / last record of every partition (binary searches in the sorted keys)
//   K3' k_count_part   one CTA per partition: the 64 KB slice of the filter is loaded into
//                      shared memory, the partition's records are replayed in stream order
//                      256 at a time -- records whose bits are all set pass (order-free);
//                      the others take turns per block, earliest first (shared-memory
//                      atomicMin claim), each applying the reference's test-then-set -- and
//                      the slice is written back once.  Passing occurrences upsert the
//                      table (or set bf_high in trim mode) straight from the CTA.
//
// HBM traffic per window: records 16 B written + 104 B through the sort + 16 B read, the
// filter once in and once out (sequential), the table through L2.  Requires the block
// index to be a bit field of y0 (n_shift - 9 <= k); otherwise count.cu's probe/resolve/
// replay path (random access, same results) is used.
#include "common.cuh"
#include "enum.cuh"
#include "partition.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <functional>

#define CP_SLOG2   8                             // log2(Bloom blocks per partition): 16 KB slices
#define CP_THREADS 128                           // 12 CTAs per SM (measured best; 64 KB / 256 threads / 3 CTAs: +40 %)
#define CP_MIN_CTAS 12
#define CP_BLK_BYTES 64                           // one Bloom block (bbf.h: 512 bits)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ------------------------------------------------------------------ record formats
//
// A record is (y0, y1, is_high): 2k + 1 bits.  Two layouts, both "a 64-bit key whose low bits are y0" (what the
// partition sort keys on) plus a value:
//   wire (PACKED = false, VT = u64): key = y0 | is_high << 63, value = y1; ~0 in the value = no k-mer / does not
//       pass.  The format of the multi-GPU exchange and of count.cu.
//   packed (PACKED = true): key = y0 | is_high << k | (low 63-k bits of y1) << (k+1), value = the remaining 2k-63
//       bits of y1 in the smallest of u8/u16/u32/u64 that leaves the top bit free; top bit set = no k-mer / does
//       not pass.  9 bytes per record at k <= 35 instead of 16: the sort, the replay and the table apply all
//       stream less.
template <typename VT, bool PACKED> struct Rec {
	static __device__ __forceinline__ VT mark() { return PACKED ? (VT)((VT)1 << (8 * sizeof(VT) - 1)) : (VT)~(VT)0; }
	static __device__ __forceinline__ void pack(int k, uint64_t y0, uint64_t y1, int high, unsigned long long &key, VT &val)
	{
		if (PACKED) {
			const int lb = 63 - k;
			key = y0 | (unsigned long long)high << k | (lb ? (y1 & ((1ULL << lb) - 1)) << (k + 1) : 0ULL);
			val = (VT)(y1 >> lb);
		} else key = y0 | (unsigned long long)high << 63, val = (VT)y1;
	}
	static __device__ __forceinline__ bool valid(VT val) { return PACKED ? (val & mark()) == 0 : val != mark(); }
	// y0f = y0 | is_high << 63
	static __device__ __forceinline__ void unpack(int k, unsigned long long key, VT val, unsigned long long &y0f, unsigned long long &y1)
	{
		if (PACKED) {
			const int lb = 63 - k;
			y0f = (key & ((1ULL << k) - 1)) | (key >> k & 1) << 63;
			y1 = (lb ? key >> (k + 1) : 0ULL) | (unsigned long long)val << lb;
		} else y0f = key, y1 = (unsigned long long)val;
	}
};

static int packed_value_bytes(int k) { return k <= 35 ? 1 : k <= 39 ? 2 : k <= 47 ? 4 : 8; }

// run `call` with the record format: vb = bytes of a packed value, or 0 = wire format
#define REC_DISPATCH(vb, call)                                                             \
	do {                                                                                   \
		if ((vb) == 0) { typedef unsigned long long VT; const bool PK = false; (void)PK; call; }   \
		else if ((vb) == 1) { typedef uint8_t VT; const bool PK = true; (void)PK; call; }        \
		else if ((vb) == 2) { typedef uint16_t VT; const bool PK = true; (void)PK; call; }       \
		else if ((vb) == 4) { typedef uint32_t VT; const bool PK = true; (void)PK; call; }       \
		else { typedef unsigned long long VT; const bool PK = true; (void)PK; call; }            \
	} while (0)

// ------------------------------------------------------------------ K0': enumeration in stream order

struct EnumLinParams {
	const uint8_t *seq, *qual;   // stream window (device); qual may be 0 (= every base has high quality)
	uint64_t len;                // bytes in the window
	uint64_t emit_from;          // records are emitted for window positions >= emit_from
	int k, q;
	unsigned long long *key;
	void *val;
	uint32_t *seg_cnt;           // k_enum_count: records per segment
	const uint32_t *seg_off;     // k_enum_lin: first record of every segment (exclusive scan of seg_cnt)
};

// how many k-mers end in every segment (the records are written back to back, in stream order)
__global__ void __launch_bounds__(EL_THREADS) k_enum_count(EnumLinParams p)
{
	__shared__ uint32_t s_nb[EL_WORDS], s_warp[EL_THREADS / 32 + 1];
	el_stage_nb(s_nb, p.seq, p.len, (int64_t)p.emit_from + (int64_t)blockIdx.x * EL_SEG);
	uint32_t total;
	el_block_scan((uint32_t)__popc(el_valid_word(s_nb, threadIdx.x, p.k)), s_warp, &total);
	if (threadIdx.x == 0) p.seg_cnt[blockIdx.x] = total;
}

// What worker_count does per base (count.c:76-88): map the character (bseq.c:9-26), restart on
// a non-ACGT one, and once k bases are in, hash the canonical k-mer (kmer.h:79-88) with
// is_high = all k bases have Q >= q.  Instead of rolling, the k-mer ending at a position is cut
// out of the base bit planes of the segment (the same construction as correct.cu's extract_kmer).
template <typename VT, bool PACKED>
__global__ void __launch_bounds__(EL_THREADS) k_enum_lin(EnumLinParams p)
{
	__shared__ uint32_t s_pl[4][EL_WORDS]; // B0, B1 (base code bits), NB (not ACGT / outside), Q (Q >= q)
	__shared__ uint32_t s_v[EL_THREADS], s_vpre[EL_THREADS], s_warp[EL_THREADS / 32 + 1];
	const int64_t seg0 = (int64_t)p.emit_from + (int64_t)blockIdx.x * EL_SEG;
	el_stage_planes(s_pl, p.seq, p.qual, p.len, seg0, p.q);
	const int k = p.k;
	const uint64_t kmask = (1ULL << k) - 1;
	{ // where k-mers end, and how many end before every plane word: the records leave compacted
		const uint32_t v = el_valid_word(s_pl[2], threadIdx.x, k);
		uint32_t total;
		s_v[threadIdx.x] = v;
		s_vpre[threadIdx.x] = el_block_scan((uint32_t)__popc(v), s_warp, &total);
		__syncthreads();
	}
	const unsigned lane = threadIdx.x & 31;
	unsigned long long *const ok = p.key + __ldg(p.seg_off + blockIdx.x);
	VT *const ov = (VT*)p.val + __ldg(p.seg_off + blockIdx.x);
#pragma unroll 4
	for (int j = 0; j < EL_ITERS; ++j) {
		const uint32_t pp = (uint32_t)(j * EL_THREADS + threadIdx.x);   // position inside the segment
		const uint32_t v = s_v[pp >> 5];
		if (v >> lane & 1) {
			const uint32_t bit = pp + EL_LEAD - (uint32_t)(k - 1);      // oldest base of the k-mer ending at pp
			const uint32_t at = s_vpre[pp >> 5] + __popc(v & ((1u << lane) - 1));
			unsigned long long key;
			VT val;
			uint64_t y[2];
			el_kmer_hash_at(s_pl, bit, k, kmask, y);
			Rec<VT, PACKED>::pack(k, y[0], y[1], (win64(s_pl[3], bit) & kmask) == kmask, key, val);
			ok[at] = key;
			ov[at] = val;
		}
	}
}

// ------------------------------------------------------------------ K2': partition bounds

// The records are sorted by partition: the range of partition p is found by two binary searches (a few million probes
// per window, against streaming every key past a comparison).  y0 points at record `base` of the array the bounds refer to.
__global__ void __launch_bounds__(256) k_part_bounds(const unsigned long long *y0, uint64_t n, uint64_t base, int shift, uint32_t pmask, uint32_t *start, uint32_t *end)
{
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p > pmask) return;
	uint64_t lo = 0, hi = n; // first record whose partition is >= p
	while (lo < hi) {
		const uint64_t mid = (lo + hi) >> 1;
		if (((uint32_t)(__ldg(y0 + mid) >> shift) & pmask) < p) lo = mid + 1; else hi = mid;
	}
	const uint64_t first = lo;
	hi = n;              // first record whose partition is > p
	while (lo < hi) {
		const uint64_t mid = (lo + hi) >> 1;
		if (((uint32_t)(__ldg(y0 + mid) >> shift) & pmask) <= p) lo = mid + 1; else hi = mid;
	}
	start[p] = (uint32_t)(base + first), end[p] = (uint32_t)(base + lo);
}

// ------------------------------------------------------------------ K3': one CTA per partition

struct PartParams {
	const unsigned long long *key;     // records, stably partitioned
	void *val;                         //   the value of a record that does not pass is overwritten with the mark (normal mode)
	const uint32_t *start, *end;       // record range per (run, partition): [run * n_parts + partition]
	uint32_t n_runs, n_parts;          // runs = record arrays partitioned separately, to be replayed one after the other
	uint32_t blocks_per_part;          // Bloom blocks per partition (power of two, <= 2^CP_SLOG2)
	int k;
	BloomView bf, bf_high;             // bf_high.w == 0 in normal mode
	unsigned long long *ctr;           // [1] n_kmers [2] n_pass [3] occurrences that had to wait for an earlier one
};

// reference bbf.c:35-42 on a block held in shared memory: test-then-set each probe, returns the number already set
__device__ __forceinline__ int bloom_test_set_smem(uint32_t *w, int h1, int h2, int H)
{
	int z = h1, done = 0, cnt = 0;
	while (done < H) {
		if (z >= 8) {
			const uint32_t bit = 1u << (z & 31), v = w[z >> 5];
			cnt += (v & bit) != 0;
			w[z >> 5] = v | bit;
			++done;
		}
		z = (z + h2) & BFC_BLK_MASK;
	}
	return cnt;
}

// ---- the slice travels as ONE bulk copy each way (TMA engine, SASS UBLKCP): no per-thread LDG / STG loop, the LSU and
// the registers stay with the replay
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_load_slice(uint32_t *s_dst, const uint32_t *g_src, uint32_t bytes, uint64_t *mbar)
{
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(mbar)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(mbar)), "r"(bytes) : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		             :: "r"(smem_u32(s_dst)), "l"(g_src), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
	}
}

__device__ __forceinline__ void bulk_load_wait(uint64_t *mbar)
{
	asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
	             :: "r"(smem_u32(mbar)) : "memory");
}

// after a __syncthreads(): the CTA's writes to the slice become visible to the async proxy, one thread sends it
__device__ __forceinline__ void bulk_store_slice(uint32_t *g_dst, const uint32_t *s_src, uint32_t bytes)
{
	if (threadIdx.x == 0) {
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(g_dst), "r"(smem_u32(s_src)), "r"(bytes) : "memory");
		asm volatile("cp.async.bulk.commit_group;" ::: "memory");
		asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // shared memory must outlive the read
	}
}

// One round = CP_THREADS consecutive records of the partition.  Records whose bits are all set pass whatever the
// order and write nothing.  Of the others, the EARLIEST of every block (shared-memory atomicMin claim) applies the
// reference's test-then-set at once -- different blocks, disjoint words.  The few that lost a claim (a later
// occurrence of the same block in the same round) are replayed afterwards by warp 0 in stream order: lane l walks
// the losers' bitmap in ascending order and handles the blocks with index = l mod 32, so one block is always
// replayed sequentially and different blocks in parallel.
template <typename VT, bool PACKED, bool BULK>
__global__ void __launch_bounds__(CP_THREADS, CP_MIN_CTAS) k_count_part(PartParams p)
{
	extern __shared__ __align__(128) uint32_t s_w[]; // the partition's slice of the filter: blocks_per_part x 16 words
	__shared__ __align__(8) uint64_t s_mbar;
	__shared__ uint32_t s_claim[1 << CP_SLOG2];      // per block: lowest thread with a pending occurrence this round
	__shared__ uint32_t s_info[CP_THREADS];          // losers: block << 18 | h1 << 9 | h2
	__shared__ uint32_t s_lose[CP_THREADS / 32], s_res[CP_THREADS / 32]; // bitmaps over the round's threads
	typedef Rec<VT, PACKED> R;
	const uint32_t part = blockIdx.x, nb = p.blocks_per_part, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	{ // nothing for this slice in this window?
		bool any = false;
		for (uint32_t r = 0; r < p.n_runs; ++r) any |= __ldg(p.start + (uint64_t)r * p.n_parts + part) < __ldg(p.end + (uint64_t)r * p.n_parts + part);
		if (!any) return;
	}
	uint32_t *const g_slice = p.bf.w + ((uint64_t)part * nb << 4);
	if (BULK) bulk_load_slice(s_w, g_slice, nb * CP_BLK_BYTES, &s_mbar);
	else for (uint32_t i = tid; i < nb * 4; i += CP_THREADS) ((uint4*)s_w)[i] = __ldcs((const uint4*)g_slice + i);
	for (uint32_t i = tid; i < nb; i += CP_THREADS) s_claim[i] = ~0u;
	if (tid < CP_THREADS / 32) s_lose[tid] = 0, s_res[tid] = 0;
	const int H = p.bf.n_hashes;
	const bool mark = p.bf_high.w == 0;
	VT *const vals = (VT*)p.val;
	unsigned long long n_k = 0, n_pass = 0, n_wait = 0;
	__syncthreads();          // (the barrier's initialisation is visible to every thread)
	bool loaded = false;
	for (uint32_t run = 0; run < p.n_runs; ++run) {
		const uint64_t beg = __ldg(p.start + (uint64_t)run * p.n_parts + part), end = __ldg(p.end + (uint64_t)run * p.n_parts + part);
		unsigned long long key = 0;
		VT val = R::mark();
		if (beg + tid < end) val = __ldg(vals + beg + tid), key = __ldg(p.key + beg + tid);
		if (BULK && !loaded) { bulk_load_wait(&s_mbar); loaded = true; } // the first records were requested while the slice travelled
		for (uint64_t base = beg; base < end; base += CP_THREADS) {
			const unsigned long long ckey = key;
			const VT cval = val;
			{ // next round's records are in flight while this round is replayed
				const uint64_t ni = base + CP_THREADS + tid;
				val = R::mark();
				if (ni < end) val = __ldg(vals + ni), key = __ldg(p.key + ni);
			}
			const bool valid = R::valid(cval);
			unsigned long long c0f = 0, c1 = 0;
			R::unpack(p.k, ckey, cval, c0f, c1);
			const uint64_t c0 = c0f & ~(1ULL << 63);
			bool pass = false, pend = false;
			BloomProbe pr;
			pr.blk = 0, pr.h1 = pr.h2 = 0;
			uint32_t lb = 0;
			if (valid) {
				pr = bloom_locate(hash_from_y(p.k, c0, c1), p.bf.n_shift);
				lb = (uint32_t)pr.blk & (nb - 1);
				pass = bloom_count_set<false>(s_w + (lb << 4), pr, H) == H; // all set: passes, writes nothing
				pend = !pass;
			}
			if (__syncthreads_or(pend)) {
				if (pend) atomicMin(&s_claim[lb], tid);
				__syncthreads();
				bool lost = false;
				if (pend) {
					if (s_claim[lb] == tid) pass = bloom_test_set_smem(s_w + (lb << 4), pr.h1, pr.h2, H) == H; // reference count.c:60
					else {
						lost = true;
						s_info[tid] = lb << 18 | (uint32_t)pr.h1 << 9 | (uint32_t)pr.h2;
						atomicOr(&s_lose[warp], 1u << lane);
					}
				}
				if (__syncthreads_or(lost)) {
					if (warp == 0) {
						for (int wd = 0; wd < CP_THREADS / 32; ++wd)
							for (uint32_t m = s_lose[wd]; m; m &= m - 1) {
								const int t = wd * 32 + __ffs(m) - 1;
								const uint32_t info = s_info[t];
								if (((info >> 18) & 31) == lane && bloom_test_set_smem(s_w + ((info >> 18) << 4), info >> 9 & 511, info & 511, H) == H)
									atomicOr(&s_res[wd], 1u << (t & 31));
							}
					}
					__syncthreads();
					if (lost) {
						pass = s_res[warp] >> lane & 1;
						atomicAnd(&s_lose[warp], ~(1u << lane));
						atomicAnd(&s_res[warp], ~(1u << lane));
						++n_wait;
					}
				}
				if (pend && !lost) s_claim[lb] = ~0u; // the winner's claim; a loser's block was claimed by a winner
			}
			if (valid) {
				if (mark) { if (!pass) vals[base + tid] = R::mark(); } // the table upserts of the passing records follow in k_tab_apply_marked
				else if (pass) {
					const BloomProbe ph = bloom_locate(hash_from_y(p.k, c0, c1), p.bf_high.n_shift);
					bloom_set_atomic(bloom_block(p.bf_high, ph.blk), ph, p.bf_high.n_hashes);
				}
			}
			n_k += valid, n_pass += pass;
		}
	}
	__syncthreads();
	if (BULK) bulk_store_slice(g_slice, s_w, nb * CP_BLK_BYTES);
	else for (uint32_t i = tid; i < nb * 4; i += CP_THREADS) __stcs((uint4*)g_slice + i, ((const uint4*)s_w)[i]);
	block_add(p.ctr + 1, n_k);
	block_add(p.ctr + 2, n_pass);
	block_add(p.ctr + 3, n_wait);
}

// bfc_ch_insert (htab.c:60-82) for every record that passed, in partition order: a warp's 32 records belong to one
// partition, whose sub-tables are neighbours in the table (tab_region), so the probes stay in L2.  Latency-bound (a
// load, then a CAS, per record): as many threads as the register file takes.
template <typename VT, bool PACKED>
__global__ void __launch_bounds__(256, 8) k_tab_apply_marked(TabView t, const unsigned long long *key, const VT *val, uint64_t n)
{
	unsigned long long added = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		const VT v = __ldg(val + i);
		if (Rec<VT, PACKED>::valid(v)) {
			unsigned long long y0f, y1;
			Rec<VT, PACKED>::unpack(t.k, __ldg(key + i), v, y0f, y1);
			added += tab_upsert(t, y0f & ~(1ULL << 63), y1, (int)(y0f >> 63)) == 1;
		}
	}
	block_add(t.counters, added);
}

// ------------------------------------------------------------------ host side

struct PartGeom {
	int x;              // log2(Bloom blocks this rank holds)
	int pshift;         // log2(Bloom blocks per partition) = min(CP_SLOG2, x) = first partition bit of y0
	int pbits;          // log2(partitions)
	uint32_t n_parts, blocks_per_part;
};

static PartGeom part_geom(int n_shift, int owner_bits)
{
	PartGeom g;
	g.x = n_shift - BFC_BLK_SHIFT - owner_bits;
	g.pshift = g.x < CP_SLOG2 ? g.x : CP_SLOG2;
	g.pbits = g.x - g.pshift;
	g.n_parts = 1u << g.pbits;
	g.blocks_per_part = 1u << g.pshift;
	return g;
}

bool bfcg_count_part_usable(const bfc_opt_t *opt, int n_shift, int owner_bits)
{
	const char *e = getenv("BFC_B200_COUNT");
	if (e && strcmp(e, "probe") == 0) return false;
	const int x = n_shift - BFC_BLK_SHIFT;
	// the block index (bbf.c:27-28) must be a bit field of y0
	return x <= opt->k && x - owner_bits >= 0 && x - owner_bits <= 36;
}

static inline int value_bytes(int vb);

// Stable partition of n records by bits [begin, end) of the key; vb = record format (REC_DISPATCH).  The work is done
// by partition.cuh (two passes of 10 bits for the usual 20 partition bits); BFC_B200_CUB_SORT=1 switches back to the
// library radix sort of round 1 (three passes of 8 bits) for comparison.  Called with tmp == 0 it only reports the
// scratch it needs.
static bool use_cub_sort() { static int v = -1; if (v < 0) v = getenv("BFC_B200_CUB_SORT") != 0; return v != 0; }

static cudaError_t sort_records(void *tmp, size_t &tmp_bytes, int vb, const unsigned long long *k_in, unsigned long long *k_out,
                                const void *v_in, void *v_out, uint64_t n, int begin, int end)
{
	cudaError_t e = cudaSuccess;
	if (use_cub_sort()) {
		REC_DISPATCH(vb, e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, (const VT*)v_in, (VT*)v_out, (int64_t)n, begin, end, bfcg_rt().stream));
		if (tmp) bfcg_rt().n_launches += 1 + (end - begin + 7) / 8; // histogram + one onesweep pass per 8 bits
		return e;
	}
	const size_t o_val = align_up(n * 8, 256), o_scr = o_val + align_up(n * (size_t)value_bytes(vb), 256);
	if (tmp == 0) { tmp_bytes = o_scr + rp_scratch_bytes(n, begin, end); return cudaSuccess; }
	uint8_t *t = (uint8_t*)tmp;
	REC_DISPATCH(vb, e = rp_partition<VT>(bfcg_rt().stream, bfcg_rt().sm_count, k_in, (const VT*)v_in, k_out, (VT*)v_out, (unsigned long long*)t, (VT*)(t + o_val),
	                                      t + o_scr, n, begin, end, &bfcg_rt().n_launches));
	return e;
}

static size_t sort_temp_bytes(int vb, uint64_t n, int begin, int end)
{
	size_t t = 0;
	if (end > begin) sort_records(0, t, vb, 0, 0, 0, 0, n, begin, end);
	return t;
}

static inline int value_bytes(int vb) { return vb ? vb : 8; }

// the partition on its own (tests): n records with vb-byte values (1, 2, 4 or 8) in device arrays
extern "C" int bfcg_partition_records(const uint64_t *d_key_in, const void *d_val_in, uint64_t *d_key_out, void *d_val_out, uint64_t n, int vb, int begin, int end)
{
	int r;
	if ((r = bfcg_rt_init()) != BFCG_OK) return r;
	BfcgRuntime &rt = bfcg_rt();
	if ((vb != 1 && vb != 2 && vb != 4 && vb != 8) || begin < 0 || end > 64 || end <= begin || n >= (1ULL << 30)) return BFCG_ERR_ARG;
	size_t tb = sort_temp_bytes(vb, n, begin, end);
	uint8_t *a = (uint8_t*)bfcg_arena(tb + 256);
	if (!a) return BFCG_ERR_NOMEM;
	BFCG_CUDA(sort_records(a, tb, vb, (const unsigned long long*)d_key_in, (unsigned long long*)d_key_out, d_val_in, d_val_out, n, begin, end));
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	return BFCG_OK;
}


struct PartScratch {
	unsigned long long *s_key; // sorted records
	void *s_val;
	uint8_t *tmp;
	size_t tmp_bytes;
	uint32_t *bounds;          // start[n_parts], end[n_parts]
	unsigned long long *ctr;
	static size_t bytes(uint64_t n, const PartGeom &g, int vb)
	{
		return align_up(n * 8, 256) + align_up(n * value_bytes(vb), 256) + align_up(sort_temp_bytes(vb, n, g.pshift, g.pshift + g.pbits), 256) +
		       align_up((size_t)g.n_parts * 8, 256) + 256;
	}
	void carve(uint8_t *a, uint64_t n, const PartGeom &g, int vb)
	{
		size_t o = 0;
		s_key = (unsigned long long*)(a + o); o += align_up(n * 8, 256);
		s_val = a + o; o += align_up(n * value_bytes(vb), 256);
		tmp_bytes = sort_temp_bytes(vb, n, g.pshift, g.pshift + g.pbits);
		tmp = a + o; o += align_up(tmp_bytes, 256);
		bounds = (uint32_t*)(a + o); o += align_up((size_t)g.n_parts * 8, 256);
		ctr = (unsigned long long*)(a + o);
	}
};

// room for the keys a window of n records can add, at load <= 1/2: at most every second occurrence is new unless
// the filter misfires; after the first window the previous one's growth is the estimate.  A region that fills up
// anyway parks its inserts (tab_upsert) and the load is put right after the window (tab_after_window).
static int tab_before_window(bfc_ch_t *ch, uint64_t n, unsigned long long *before)
{
	*before = bfc_ch_count(ch);
	uint64_t extra = n / 2;
	if (ch->have_prev) extra = std::min<uint64_t>(extra, std::max<uint64_t>(2 * ch->prev_new, n / 16));
	return bfcg_tab_reserve(ch, std::max<uint64_t>(extra, 1024));
}

static int tab_after_window(bfc_ch_t *ch, unsigned long long before)
{
	int r;
	if ((r = bfcg_tab_drain_deferred(ch)) != BFCG_OK) return r;
	ch->prev_new = bfc_ch_count(ch) - before, ch->have_prev = 1;
	return bfcg_tab_reserve(ch, 0);
}

// K2'-K3' + the table upserts over records that are already partitioned: n_runs arrays back to back in key / val (run r
// = records [run_off[r], run_off[r+1])), each stably sorted by partition; the runs are replayed one after the other.
// The values are overwritten (marks).  `bounds` has room for 2 * n_runs * n_parts words.
static int count_part_sorted(const bfc_opt_t *opt, bfc_bf_t *bf, bfc_bf_t *bf_high, bfc_ch_t *ch, const PartGeom &g, uint64_t blk_mask, int vb,
                             const unsigned long long *key, void *val, int n_runs, const uint64_t *run_off,
                             uint32_t *bounds, unsigned long long *ctr, bfcg_stats_t *stats, const std::function<int()> *launched,
                             unsigned long long n_before /* keys in the table now (no kernel that adds any is in flight) */)
{
	BfcgRuntime &rt = bfcg_rt();
	int r;
	const uint64_t n = run_off[n_runs];
	PartParams p;
	memset(&p, 0, sizeof(p));
	p.k = opt->k, p.blocks_per_part = g.blocks_per_part, p.n_runs = n_runs, p.n_parts = g.n_parts;
	p.bf = bloom_view(bf), p.bf.blk_mask = blk_mask;
	if (bf_high) p.bf_high = bloom_view(bf_high), p.bf_high.blk_mask = blk_mask;
	p.ctr = ctr, p.key = key, p.val = val;
	uint32_t *start = bounds, *end = bounds + (size_t)n_runs * g.n_parts;
	BFCG_CUDA(cudaMemsetAsync(bounds, 0, (size_t)n_runs * g.n_parts * 8, rt.stream));
	{
		KTime kt(KT_COUNT_BOUNDS);
		for (int i = 0; i < n_runs; ++i) {
			const uint64_t m = run_off[i + 1] - run_off[i];
			if (m) k_part_bounds<<<(g.n_parts + 255) / 256, 256, 0, rt.stream>>>(key + run_off[i], m, run_off[i], g.pshift, g.n_parts - 1,
			                                                                    start + (size_t)i * g.n_parts, end + (size_t)i * g.n_parts);
		}
	}
	BFCG_LAUNCH_CHECK();
	p.start = start, p.end = end;
	BFCG_CUDA(cudaMemsetAsync(p.ctr, 0, 64, rt.stream));
	const size_t smem = (size_t)g.blocks_per_part * CP_BLK_BYTES;
	{
		KTime kt(KT_COUNT_PART);
		if (getenv("BFC_B200_NO_BULK")) REC_DISPATCH(vb, (k_count_part<VT, PK, false><<<g.n_parts, CP_THREADS, smem, rt.stream>>>(p)));
		else REC_DISPATCH(vb, (k_count_part<VT, PK, true><<<g.n_parts, CP_THREADS, smem, rt.stream>>>(p)));
	}
	BFCG_LAUNCH_CHECK();
	// bfc_ch_insert for the records that passed.  The growth estimate (tab_before_window) sizes the table for the keys a
	// window usually adds; should a window add far more (a saturated first filter passes everything), no region may
	// fill up beyond what the parking list takes.  So the records are applied in stretches that cannot lift the load
	// above 3/4 even if every one of them is a new key -- normally the whole window is one stretch -- and the table
	// grows between stretches when it has to.
	unsigned long long c[4];
	for (uint64_t pos = 0; ch && pos < n;) {
		const uint64_t cap = bfcg_tab_capacity(ch), skew = ch->skew > 1 ? (uint64_t)ch->skew : 1;
		const unsigned long long c0 = pos == 0 ? n_before : bfc_ch_count(ch); // (the first stretch goes out without a host sync)
		const uint64_t lim = cap / 4 * 3 / skew;
		const uint64_t room = lim > c0 ? lim - c0 : 0;
		if (room < std::min<uint64_t>(n - pos, std::max<uint64_t>(cap / 16 / skew, 1))) {
			if ((r = bfcg_tab_grow(ch)) != BFCG_OK) return r;
			continue;
		}
		const uint64_t m = std::min<uint64_t>(n - pos, room);
		{
			KTime kt(KT_TAB_APPLY);
			const unsigned grid = (unsigned)std::min<uint64_t>((m + 255) / 256, (uint64_t)rt.sm_count * 8);
			REC_DISPATCH(vb, (k_tab_apply_marked<VT, PK><<<grid, 256, 0, rt.stream>>>(tab_view(ch), key + pos, (const VT*)val + pos, m)));
		}
		BFCG_LAUNCH_CHECK();
		pos += m;
		if (pos == m && launched) { // host work that should overlap the kernels just enqueued (the next window's copy)
			if ((r = (*launched)()) != BFCG_OK) return r;
			launched = 0;
		}
		if (pos < n && (r = bfcg_tab_drain_deferred(ch)) != BFCG_OK) return r; // (the last stretch is drained by the caller)
	}
	if (launched && (r = (*launched)()) != BFCG_OK) return r;
	BFCG_CUDA(cudaMemcpyAsync(c, p.ctr, sizeof(c), cudaMemcpyDeviceToHost, rt.stream));
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	if (stats) {
		stats->n_kmers += c[1], stats->n_pass += c[2];
		stats->n_pending += c[1] - c[2], stats->n_conflict += c[3];
	}
	return BFCG_OK;
}

// K1'-K3' over n records in stream order (in_key / in_val, device)
static int count_part_window(const bfc_opt_t *opt, bfc_bf_t *bf, bfc_bf_t *bf_high, bfc_ch_t *ch, const PartGeom &g, uint64_t blk_mask, int vb,
                             const unsigned long long *in_key, const void *in_val, uint64_t n, PartScratch &sc, bfcg_stats_t *stats,
                             const std::function<int()> *launched = 0)
{
	BfcgRuntime &rt = bfcg_rt();
	int r;
	unsigned long long before = 0;
	if (ch && (r = tab_before_window(ch, n, &before)) != BFCG_OK) return r;
	const unsigned long long *key = in_key;
	if (g.pbits > 0) {
		size_t tb = sc.tmp_bytes;
		cudaError_t se;
		{
			KTime kt(KT_COUNT_SORT);
			se = sort_records(sc.tmp, tb, vb, in_key, sc.s_key, in_val, sc.s_val, n, g.pshift, g.pshift + g.pbits);
		}
		BFCG_CUDA(se);
		key = sc.s_key;
	} else BFCG_CUDA(cudaMemcpyAsync(sc.s_val, in_val, n * value_bytes(vb), cudaMemcpyDeviceToDevice, rt.stream)); // the values get marked in place
	const uint64_t run_off[2] = { 0, n };
	if ((r = count_part_sorted(opt, bf, bf_high, ch, g, blk_mask, vb, key, sc.s_val, 1, run_off, sc.bounds, sc.ctr, stats, launched, before)) != BFCG_OK) return r;
	return ch ? tab_after_window(ch, before) : BFCG_OK;
}

static int part_kernel_setup()
{
	static bool done = false;
	if (!done) {
		const int bytes = (1 << CP_SLOG2) * CP_BLK_BYTES;
#define CP_ATTR(VT, PK) \
		BFCG_CUDA(cudaFuncSetAttribute(k_count_part<VT, PK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)); \
		BFCG_CUDA(cudaFuncSetAttribute(k_count_part<VT, PK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))
		CP_ATTR(unsigned long long, false);
		CP_ATTR(uint8_t, true);
		CP_ATTR(uint16_t, true);
		CP_ATTR(uint32_t, true);
		CP_ATTR(unsigned long long, true);
#undef CP_ATTR
		done = true;
	}
	return BFCG_OK;
}

// stream positions per window: as many as the free memory takes, at most 2^30
static uint64_t window_positions(uint64_t n_positions, bool host, const PartGeom &g, int vb)
{
	const char *e = getenv("BFC_B200_COUNT_WINDOW");
	uint64_t P = 1ULL << 30;
	if (e && atoll(e) >= EL_SEG) P = el_padded((uint64_t)atoll(e));
	else {
		size_t fr = 0, tot = 0;
		cudaMemGetInfo(&fr, &tot);
		const double avail = 0.55 * (double)(fr + bfcg_rt().arena_bytes);
		while (P > (1ULL << 22) && (double)(P * (2 * (8 + value_bytes(vb)) + (host ? 4 : 0))) + (double)sort_temp_bytes(vb, P, g.pshift, g.pshift + g.pbits) > avail) P >>= 1;
	}
	return std::min(P, el_padded(n_positions));
}

#define ENUM_SCAN_TMP 256

// exclusive prefix sum of the per-segment record counts (at most 2^17 + 1 of them): one CTA, every thread sums a
// contiguous stretch, the stretch totals are scanned across the CTA, the stretches are written back
__global__ void __launch_bounds__(1024) k_seg_scan(const uint32_t *cnt, uint32_t *off, uint32_t n)
{
	__shared__ uint32_t s_w[33];
	const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t per = (n + 1023) / 1024, lo = tid * per < n ? tid * per : n, hi = lo + per < n ? lo + per : n;
	uint32_t sum = 0;
	for (uint32_t i = lo; i < hi; ++i) sum += cnt[i];
	uint32_t inc = sum;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (unsigned)d) inc += u; }
	if (lane == 31) s_w[warp] = inc;
	__syncthreads();
	if (warp == 0) {
		const uint32_t t = s_w[lane];
		uint32_t ti = t;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, ti, d); if (lane >= (unsigned)d) ti += u; }
		s_w[lane] = ti - t;
	}
	__syncthreads();
	uint32_t run = s_w[warp] + inc - sum;
	for (uint32_t i = lo; i < hi; ++i) { const uint32_t c = cnt[i]; off[i] = run; run += c; }
}

// bytes of scratch enumerate_window needs for n_pos (padded) positions: per-segment counts + offsets + scan scratch
static size_t enum_scratch_bytes(uint64_t n_pos) { return 2 * align_up((n_pos / EL_SEG + 1) * 4, 256) + ENUM_SCAN_TMP; }

// K0': the records of the n_pos (padded) positions described by ep, compacted, into ep.key / ep.val; *n_valid = how many
static int enumerate_window(int vb, EnumLinParams ep, uint64_t n_pos, uint8_t *scratch, uint64_t *n_valid)
{
	BfcgRuntime &rt = bfcg_rt();
	const uint64_t n_seg = n_pos / EL_SEG;
	uint32_t *seg_cnt = (uint32_t*)scratch, *seg_off = (uint32_t*)(scratch + align_up((n_seg + 1) * 4, 256));
	uint32_t last[2];
	ep.seg_cnt = seg_cnt, ep.seg_off = seg_off;
	KTime kt(KT_ENUM_LIN);
	k_enum_count<<<(unsigned)n_seg, EL_THREADS, 0, rt.stream>>>(ep);
	BFCG_LAUNCH_CHECK();
	k_seg_scan<<<1, 1024, 0, rt.stream>>>(seg_cnt, seg_off, (uint32_t)n_seg);
	BFCG_LAUNCH_CHECK();
	REC_DISPATCH(vb, (k_enum_lin<VT, PK><<<(unsigned)n_seg, EL_THREADS, 0, rt.stream>>>(ep)));
	BFCG_LAUNCH_CHECK();
	BFCG_CUDA(cudaMemcpyAsync(&last[0], seg_cnt + n_seg - 1, 4, cudaMemcpyDeviceToHost, rt.stream));
	BFCG_CUDA(cudaMemcpyAsync(&last[1], seg_off + n_seg - 1, 4, cudaMemcpyDeviceToHost, rt.stream));
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	*n_valid = (uint64_t)last[0] + last[1];
	return BFCG_OK;
}

int bfcg_count_part_batch(const bfc_opt_t *opt, bfc_bf_t *bf, bfc_bf_t *bf_high, bfc_ch_t *ch, const bfcg_batch_t *batch, bfcg_stats_t *stats)
{
	BfcgRuntime &rt = bfcg_rt();
	int r;
	if ((r = part_kernel_setup()) != BFCG_OK) return r;
	const PartGeom g = part_geom(bf->n_shift, 0);
	if (ch && (r = bfcg_tab_align_to_filter(ch, bf->n_shift - BFC_BLK_SHIFT)) != BFCG_OK) return r;
	const int vb = getenv("BFC_B200_COUNT_WIRE") ? 0 : packed_value_bytes(opt->k); // records never leave the device: packed
	const bool host = batch->where == BFCG_HOST;
	const uint64_t nbytes = batch->n_bytes, halo = opt->k - 1;
	const uint64_t P = window_positions(nbytes, host, g, vb);
	const uint64_t W = align_up(P + halo, 256);
	PartScratch sc;
	size_t o_stage[2][2] = {{0, 0}, {0, 0}}, o_key, o_val, o_sc, tot = 0;
	if (host)
		for (int b = 0; b < 2; ++b)
			for (int j = 0; j < 2; ++j) { o_stage[b][j] = tot; tot += W; }
	size_t o_enum;
	o_key = tot; tot += align_up(P * 8, 256);
	o_val = tot; tot += align_up(P * value_bytes(vb), 256);
	o_enum = tot; tot += enum_scratch_bytes(P);
	o_sc = tot; tot += PartScratch::bytes(P, g, vb);
	uint8_t *a = (uint8_t*)bfcg_arena(tot);
	if (!a) return BFCG_ERR_NOMEM;
	sc.carve(a + o_sc, P, g, vb);
	unsigned long long *rec_key = (unsigned long long*)(a + o_key);
	void *rec_val = a + o_val;

	const uint64_t n_win = (nbytes + P - 1) / P;
	// host batches: the copy of window i+1 runs on its own stream while window i is counted
	auto issue_copy = [&](uint64_t wi) -> cudaError_t {
		const uint64_t s = wi * P, e = std::min(nbytes, s + P), w0 = s >= halo ? s - halo : 0;
		const int b = (int)(wi & 1);
		cudaError_t ce;
		if (wi >= 2 && (ce = cudaStreamWaitEvent(rt.copy_in, rt.ev_free[b], 0)) != cudaSuccess) return ce;
		if ((ce = cudaMemcpyAsync(a + o_stage[b][0], batch->seq + w0, e - w0, cudaMemcpyHostToDevice, rt.copy_in)) != cudaSuccess) return ce;
		if (batch->qual && (ce = cudaMemcpyAsync(a + o_stage[b][1], batch->qual + w0, e - w0, cudaMemcpyHostToDevice, rt.copy_in)) != cudaSuccess) return ce;
		return cudaEventRecord(rt.ev_in[b], rt.copy_in);
	};
	BfcgTimer timer(stats);
	r = BFCG_OK;
	if (host) {
		BFCG_CUDA(cudaStreamSynchronize(rt.stream)); // the arena may still be in use by earlier work on the engine's stream
		BFCG_CUDA(issue_copy(0));
	}
	for (uint64_t wi = 0; wi < n_win && r == BFCG_OK; ++wi) {
		const uint64_t s = wi * P, e = std::min(nbytes, s + P), w0 = s >= halo ? s - halo : 0;
		const int b = (int)(wi & 1);
		EnumLinParams ep;
		memset(&ep, 0, sizeof(ep));
		ep.k = opt->k, ep.q = opt->q, ep.key = rec_key, ep.val = rec_val;
		ep.len = e - w0, ep.emit_from = s - w0;
		if (host) {
			const cudaError_t ce = cudaStreamWaitEvent(rt.stream, rt.ev_in[b], 0);
			if (ce != cudaSuccess) { r = bfcg_fail(__func__, "staging copy", ce); break; }
			ep.seq = a + o_stage[b][0], ep.qual = batch->qual ? a + o_stage[b][1] : 0;
		} else ep.seq = batch->seq + w0, ep.qual = batch->qual ? batch->qual + w0 : 0;
		uint64_t n_rec = 0;
		if ((r = enumerate_window(vb, ep, el_padded(e - s), a + o_enum, &n_rec)) != BFCG_OK) break;
		if (host) cudaEventRecord(rt.ev_free[b], rt.stream);
		// the next window travels while this one is counted (issued after this window's launches: a copy from
		// pageable memory blocks the host, and should not hold the kernels back)
		const std::function<int()> next_copy = [&]() -> int {
			if (!host || wi + 1 >= n_win) return BFCG_OK;
			const cudaError_t ce = issue_copy(wi + 1);
			return ce == cudaSuccess ? BFCG_OK : bfcg_fail("bfcg_count_part_batch", "staging copy", ce);
		};
		if (n_rec) r = count_part_window(opt, bf, bf_high, ch, g, ~0ULL, vb, rec_key, rec_val, n_rec, sc, stats, &next_copy);
		else r = next_copy();
	}
	if (host) cudaStreamSynchronize(rt.copy_in);
	if (r != BFCG_OK) { cudaStreamSynchronize(rt.stream); return r; }
	timer.stop();
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	return BFCG_OK;
}

// records in stream order in the wire format (count.cu's bucket exchange): partitioned here
int bfcg_count_part_records(const bfc_opt_t *opt, bfc_bf_t *bf, bfc_bf_t *bf_high, bfc_ch_t *ch, uint64_t n_rec,
                            const uint64_t *d_y0, const uint64_t *d_y1, int owner_bits, bfcg_stats_t *stats)
{
	BfcgRuntime &rt = bfcg_rt();
	int r;
	if ((r = part_kernel_setup()) != BFCG_OK) return r;
	const PartGeom g = part_geom(bf->n_shift, owner_bits);
	if (ch && (r = bfcg_tab_align_to_filter(ch, bf->n_shift - BFC_BLK_SHIFT)) != BFCG_OK) return r;
	const uint64_t P = std::min<uint64_t>(window_positions(n_rec, false, g, 0), n_rec);
	PartScratch sc;
	uint8_t *a = (uint8_t*)bfcg_arena(PartScratch::bytes(P, g, 0));
	if (!a) return BFCG_ERR_NOMEM;
	sc.carve(a, P, g, 0);
	const uint64_t blk_mask = (1ULL << g.x) - 1; // this rank holds 1/n_owners of the blocks
	BfcgTimer timer(stats);
	for (uint64_t s = 0; s < n_rec; s += P) {
		const uint64_t n = std::min(P, n_rec - s);
		if ((r = count_part_window(opt, bf, bf_high, ch, g, blk_mask, 0, (const unsigned long long*)d_y0 + s, d_y1 + s, n, sc, stats)) != BFCG_OK) return r;
	}
	timer.stop();
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	return BFCG_OK;
}

// ------------------------------------------------------------------ sharded counting with the partition done by the sender

// bfcg_enum_records when the partitioned path applies: enumerate in stream order, then ONE stable radix sort over the
// whole block prefix -- owner bits on top, partition bits below -- so that the buckets of the owners are contiguous for
// the all-to-all AND every bucket arrives already partitioned: the receiver replays the pieces of the ranks one after
// the other (count_part_sorted with n_runs = ranks) and never sorts.
int bfcg_enum_part_records(const bfc_opt_t *opt, const bfcg_batch_t *batch, int owner_bits, uint64_t *d_y0, uint64_t *d_y1, uint64_t *counts)
{
	return bfcg_enum_part_records_fmt(opt, batch, owner_bits, 0, d_y0, d_y1, counts);
}

int bfcg_part_record_value_bytes(int k) { return packed_value_bytes(k); }

// vb = 0: 16-byte wire records (key = y0 | is_high << 63, value = y1); vb = 1, 2, 4, 8: packed records with a vb-byte
// value (Rec<VT, true>), what stays inside one job and what the library's own exchange (dist.cu) sends
int bfcg_enum_part_records_fmt(const bfc_opt_t *opt, const bfcg_batch_t *batch, int owner_bits, int vb, uint64_t *d_y0, void *d_y1, uint64_t *counts)
{
	BfcgRuntime &rt = bfcg_rt();
	const int n_owners = 1 << owner_bits;
	const PartGeom g = part_geom(opt->bf_shift, owner_bits);
	const bool host = batch->where == BFCG_HOST;
	const uint64_t nb = batch->n_bytes, n_rec = el_padded(nb);
	const int sort_bits = g.pbits + owner_bits;
	size_t temp = sort_temp_bytes(vb, n_rec, g.pshift, g.pshift + sort_bits);
	size_t o_seq = 0, o_qual = 0, o_y0, o_y1, o_tmp, o_bnd, tot = 0;
	if (host) { o_seq = tot; tot = align_up(tot + nb, 256); o_qual = tot; tot = align_up(tot + nb, 256); }
	o_y0 = tot; tot = align_up(tot + n_rec * 8, 256);
	o_y1 = tot; tot = align_up(tot + n_rec * value_bytes(vb), 256);
	o_tmp = tot; tot = align_up(tot + temp, 256);
	size_t o_enum = tot; tot += enum_scratch_bytes(n_rec);
	o_bnd = tot; tot += 256;
	uint8_t *a = (uint8_t*)bfcg_arena(tot);
	if (!a) return BFCG_ERR_NOMEM;
	EnumLinParams ep;
	memset(&ep, 0, sizeof(ep));
	ep.k = opt->k, ep.q = opt->q, ep.len = nb, ep.emit_from = 0;
	ep.key = (unsigned long long*)(a + o_y0), ep.val = a + o_y1; // n_rec (padded) records
	if (host) {
		BFCG_CUDA(cudaMemcpyAsync(a + o_seq, batch->seq, nb, cudaMemcpyHostToDevice, rt.stream));
		if (batch->qual) BFCG_CUDA(cudaMemcpyAsync(a + o_qual, batch->qual, nb, cudaMemcpyHostToDevice, rt.stream));
		ep.seq = a + o_seq, ep.qual = batch->qual ? a + o_qual : 0;
	} else ep.seq = batch->seq, ep.qual = batch->qual;
	uint64_t nv = 0; // records = positions where a k-mer ends (<= batch->n_bytes, what the caller's arrays hold)
	{ int er = enumerate_window(vb, ep, n_rec, a + o_enum, &nv); if (er != BFCG_OK) return er; }
	if (nv == 0) return BFCG_OK;
	if (sort_bits > 0) {
		cudaError_t se;
		{
			KTime kt(KT_COUNT_SORT);
			se = sort_records(a + o_tmp, temp, vb, ep.key, (unsigned long long*)d_y0, ep.val, d_y1, nv, g.pshift, g.pshift + sort_bits);
		}
		BFCG_CUDA(se);
	} else {
		BFCG_CUDA(cudaMemcpyAsync(d_y0, ep.key, nv * 8, cudaMemcpyDeviceToDevice, rt.stream));
		BFCG_CUDA(cudaMemcpyAsync(d_y1, ep.val, nv * value_bytes(vb), cudaMemcpyDeviceToDevice, rt.stream));
	}
	uint32_t h_bnd[2 * 8];
	memset(h_bnd, 0, sizeof(h_bnd));
	if (owner_bits > 0) { // bucket sizes = runs of the owner bits in the sorted keys
		uint32_t *bnd = (uint32_t*)(a + o_bnd);
		BFCG_CUDA(cudaMemsetAsync(bnd, 0, 64, rt.stream));
		{ KTime kt(KT_BUCKET); k_part_bounds<<<1, 256, 0, rt.stream>>>((const unsigned long long*)d_y0, nv, 0, g.pshift + g.pbits, n_owners - 1, bnd, bnd + 8); }
		BFCG_LAUNCH_CHECK();
		BFCG_CUDA(cudaMemcpyAsync(h_bnd, bnd, 64, cudaMemcpyDeviceToHost, rt.stream));
	} else h_bnd[8] = (uint32_t)nv;
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	for (int o = 0; o < n_owners; ++o) counts[o] = h_bnd[8 + o] - h_bnd[o];
	return BFCG_OK;
}

// The cascade over the pieces an all-to-all delivered, back to back in source-rank order (= global read order), each
// piece partitioned by its sender (bfcg_enum_part_records).  d_y1 is overwritten.
int bfcg_count_part_runs(const bfc_opt_t *opt, bfc_bf_t *bf, bfc_bf_t *bf_high, bfc_ch_t *ch, int n_runs, const uint64_t *run_counts,
                         const uint64_t *d_y0, uint64_t *d_y1, int owner_bits, bfcg_stats_t *stats)
{
	return bfcg_count_part_runs_fmt(opt, bf, bf_high, ch, n_runs, run_counts, 0, d_y0, d_y1, owner_bits, stats);
}

int bfcg_count_part_runs_fmt(const bfc_opt_t *opt, bfc_bf_t *bf, bfc_bf_t *bf_high, bfc_ch_t *ch, int n_runs, const uint64_t *run_counts,
                             int vb, const uint64_t *d_y0, void *d_y1, int owner_bits, bfcg_stats_t *stats)
{
	BfcgRuntime &rt = bfcg_rt();
	int r;
	if ((r = part_kernel_setup()) != BFCG_OK) return r;
	const PartGeom g = part_geom(bf->n_shift, owner_bits);
	if (ch && (r = bfcg_tab_align_to_filter(ch, bf->n_shift - BFC_BLK_SHIFT)) != BFCG_OK) return r;
	uint64_t run_off[9];
	run_off[0] = 0;
	for (int i = 0; i < n_runs; ++i) run_off[i + 1] = run_off[i] + run_counts[i];
	const uint64_t n = run_off[n_runs];
	if (n == 0) return BFCG_OK;
	if (n >= (1ULL << 32)) return bfcg_fail(__func__, "more than 2^32 records in one exchange", cudaSuccess), BFCG_ERR_ARG;
	const size_t o_ctr = align_up((size_t)n_runs * g.n_parts * 8, 256);
	uint8_t *a = (uint8_t*)bfcg_arena(o_ctr + 256);
	if (!a) return BFCG_ERR_NOMEM;
	BfcgTimer timer(stats);
	unsigned long long before = 0;
	if (ch && (r = tab_before_window(ch, n, &before)) != BFCG_OK) return r;
	if ((r = count_part_sorted(opt, bf, bf_high, ch, g, (1ULL << g.x) - 1, vb, (const unsigned long long*)d_y0, d_y1, n_runs, run_off,
	                           (uint32_t*)a, (unsigned long long*)(a + o_ctr), stats, 0, before)) != BFCG_OK) return r;
	if (ch && (r = tab_after_window(ch, before)) != BFCG_OK) return r;
	timer.stop();
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	return BFCG_OK;
}
