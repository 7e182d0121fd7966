// correct.cu -- the correction phase: what the reference runs inside
// kt_for(..., worker_ec, ...) (correct.c:587): bfc_ec1 per read (correct.c:388-472)
// with its two bfc_ec1dir heap searches (correct.c:249-386), and the `-1` trim
// lookup (max_streak, correct.c:478-497, keep rule :555-569).
//
// One read per thread: the search is a chain of dependent table lookups, so
// throughput comes from reads in flight, not from per-read speed.  Each thread owns a
// fixed heap (max_heap + 4 entries are enough: the reference stops growing it at
// max_heap, correct.c:349) and a fixed stack; a read whose stack overflows is re-run
// by the same kernel with a larger stack (never on the CPU).
//
// Pop/push order, klib heap tie-breaking (ksort.h:125-146) and every threshold follow
// the reference exactly: the `ec:Z:` tag exposes max_heap / n_absent, so the search
// internals are part of the byte-parity contract.
#include "common.cuh"
#include <algorithm>
#include <climits>
#include <vector>

#define EC_OVERFLOW (-100)

struct HeapEnt {                 // reference correct.c:153-160 (echeap1_t)
	int tot_pen, i, k;
	int ecpos_high[BFC_EC_HIST_HIGH];
	int ecpos[BFC_EC_HIST];
	uint64_t x[4];
};

struct StackEnt {                // reference correct.c:162-167 (ecstack1_t), cnt dropped (only logged)
	int parent, i, tot_pen;
	uint32_t info;               // b | ec << 4 | ec_high << 5 | absent << 6 | absent_high << 7
};

struct Pen { int ec, ec_high, absent, absent_high, b; };

struct EcParams {
	const uint64_t *off;
	uint8_t *seq, *qual;
	int64_t n_reads;
	uint32_t *aux;
	uint8_t *fb, *p0, *p1;       // per-base scratch, same offsets as seq
	TabView tab;
	int k, q, min_cov, win_multi_ec, max_end_ext;
	int w_ec, w_ec_high, w_absent, w_absent_high, max_path_diff, max_heap, mode;
	HeapEnt *heap;               // heap_cap entries per thread slot
	StackEnt *stack;             // stack_cap entries per thread slot
	int heap_cap, stack_cap;
	const uint32_t *redo;        // when set: thread slot s handles read redo[s]; n_reads = #redo
	uint32_t *overflow;          // read indices whose stack overflowed
	unsigned long long *ctr;     // [0] n_overflow, [1] n_lookups
};

// per-base scratch byte
#define FB_B(v)        ((v) & 7)
#define FB_Q           8
#define FB_LCOV        16   // lcov >= min_cov + 1
#define FB_HCOV        32   // hcov > 0.75 k
#define FB_SOLID       64   // solid_end
#define FB_HSOLID      128  // solid_end && high_end

__device__ __forceinline__ int comp_b(int b) { return b < 4 ? 3 - b : 4; }

// view of the read in search coordinates (dir 1 = reverse complement, correct.c:39-57)
struct RView {
	const uint8_t *fb;
	int n, dir;
	__device__ __forceinline__ uint8_t raw(int i) const { return fb[dir ? n - 1 - i : i]; }
	__device__ __forceinline__ int b(int i) const { const int v = FB_B(raw(i)); return dir ? comp_b(v) : v; }
};

// klib heap with "less" = larger tot_pen: root = smallest penalty (correct.c:179, ksort.h:125-146)
__device__ __forceinline__ void heap_down(HeapEnt *l, int n)
{
	int i = 0, c;
	const HeapEnt tmp = l[0];
	while ((c = 2 * i + 1) < n) {
		if (c != n - 1 && l[c].tot_pen > l[c + 1].tot_pen) ++c;
		if (l[c].tot_pen > tmp.tot_pen) break;
		l[i] = l[c]; i = c;
	}
	l[i] = tmp;
}

__device__ __forceinline__ void heap_up(HeapEnt *l, int n)
{
	int c = n - 1;
	const HeapEnt tmp = l[c];
	while (c) {
		const int par = (c - 1) >> 1;
		if (tmp.tot_pen > l[par].tot_pen) break;
		l[c] = l[par]; c = par;
	}
	l[c] = tmp;
}

__device__ __forceinline__ int pen_weight(const EcParams &P, const Pen &p)
{
	return P.w_ec * p.ec + P.w_ec_high * p.ec_high + P.w_absent * p.absent + P.w_absent_high * p.absent_high;
}

// reference correct.c:198-230 (buf_update); false when the stack is full
__device__ __forceinline__ bool push_state(const EcParams &P, HeapEnt *heap, int &heap_n, StackEnt *stack, int &stack_n,
                                           int stack_cap, const HeapEnt &prev, const Pen &pen)
{
	if (stack_n >= stack_cap || heap_n >= P.heap_cap) return false;
	StackEnt q;
	q.parent = prev.k, q.i = prev.i;
	q.info = (uint32_t)pen.b | pen.ec << 4 | pen.ec_high << 5 | pen.absent << 6 | pen.absent_high << 7;
	q.tot_pen = prev.tot_pen + pen_weight(P, pen);
	stack[stack_n++] = q;
	HeapEnt r;
	r.i = prev.i + 1;
	r.k = stack_n - 1;
	r.x[0] = prev.x[0], r.x[1] = prev.x[1], r.x[2] = prev.x[2], r.x[3] = prev.x[3];
	if (pen.ec_high) r.ecpos_high[0] = prev.i, r.ecpos_high[1] = prev.ecpos_high[0];
	else r.ecpos_high[0] = prev.ecpos_high[0], r.ecpos_high[1] = prev.ecpos_high[1];
	if (pen.ec) {
		r.ecpos[0] = prev.i;
#pragma unroll
		for (int t = 1; t < BFC_EC_HIST; ++t) r.ecpos[t] = prev.ecpos[t - 1];
	} else {
#pragma unroll
		for (int t = 0; t < BFC_EC_HIST; ++t) r.ecpos[t] = prev.ecpos[t];
	}
	r.tot_pen = q.tot_pen;
	bfc_kmer_append(P.k, r.x, pen.b);
	heap[heap_n++] = r;
	heap_up(heap, heap_n);
	return true;
}

// reference correct.c:249-386 (bfc_ec1dir).  `path` receives ec[].b in FORWARD read
// coordinates (for dir 1 that is the result after the reference's final revcomp).
__device__ int ec_search(const EcParams &P, const RView &rv, int start, int end, HeapEnt *heap, StackEnt *stack,
                         int stack_cap, uint8_t *path, int &max_heap, unsigned long long &n_lookups)
{
	const int k = P.k, n = rv.n;
	HeapEnt z;
	int heap_n = 0, stack_n = 0, rvl = -1, n_paths = 0, best = -1, best_pen = INT_MAX, n_fail = 0, run = 0;
	int paths[BFC_MAX_PATHS];
	max_heap = 0;
	z.tot_pen = 0, z.k = -1;
	z.x[0] = z.x[1] = z.x[2] = z.x[3] = 0;
	for (z.i = start; z.i < end; ++z.i) { // seed: k-1 bases (correct.c:260-267)
		const int c = rv.b(z.i);
		if (c < 4) {
			if (++run == k) break;
			bfc_kmer_append(k, z.x, c);
		} else run = 0, z.x[0] = z.x[1] = z.x[2] = z.x[3] = 0;
	}
	if (z.i >= end) return -1; // the reference asserts; cannot happen after an island / rescue was found
#pragma unroll
	for (int t = 0; t < BFC_EC_HIST; ++t) z.ecpos[t] = -1;
#pragma unroll
	for (int t = 0; t < BFC_EC_HIST_HIGH; ++t) z.ecpos_high[t] = -1;
	heap[heap_n++] = z;

	for (;;) {
		bool stop = false;
		max_heap = max_heap > 255 ? 255 : max_heap > heap_n ? max_heap : heap_n;
		if (heap_n == 0) { rvl = -2; break; }
		z = heap[0];
		heap[0] = heap[--heap_n];
		heap_down(heap, heap_n);
		if (best >= 0 && z.tot_pen > best_pen + P.max_path_diff) break;
		if (z.i - end > P.max_end_ext) stop = true;
		if (!stop) {
			const bool has_c = z.i < n;
			const uint8_t craw = has_c ? rv.raw(z.i) : 0;
			const int cb = has_c ? (rv.dir ? comp_b(FB_B(craw)) : FB_B(craw)) : -1;
			const int cq = (craw & FB_Q) != 0;
			int os = -1, other_ext = 0, n_added = 0;
			bool fixed = z.i > end;
			Pen added[4];
			if (has_c && cb < 4) {
				uint64_t x[4] = { z.x[0], z.x[1], z.x[2], z.x[3] };
				bfc_kmer_append(k, x, cb);
				os = tab_kmer_occ(P.tab, x);
				++n_lookups;
				if (cq && (os & 0xff) >= P.min_cov + 1 && (craw & FB_LCOV)) fixed = true;
				else if (craw & FB_HCOV) fixed = true;
			}
			for (int b = 0; b < 4; ++b) {
				Pen pen;
				if (fixed && has_c && b != cb) continue;
				if (!has_c || b != cb) {
					if (has_c) {
						if (cq && z.ecpos_high[BFC_EC_HIST_HIGH - 1] >= 0 && z.i - z.ecpos_high[BFC_EC_HIST_HIGH - 1] < P.win_multi_ec) continue;
						if (z.ecpos[BFC_EC_HIST - 1] >= 0 && z.i - z.ecpos[BFC_EC_HIST - 1] < P.win_multi_ec) continue;
					}
					uint64_t x[4] = { z.x[0], z.x[1], z.x[2], z.x[3] };
					bfc_kmer_append(k, x, b);
					const int s = tab_kmer_occ(P.tab, x);
					++n_lookups;
					if (s < 0 || (s & 0xff) < P.min_cov) continue;
					pen.ec = has_c && cb < 4 ? 1 : 0;
					pen.ec_high = pen.ec ? cq : 0; // oq == q for every base (correct.c:32-33)
					pen.absent = 0;
					pen.absent_high = ((s >> 8 & 0xff) < P.min_cov);
					pen.b = b;
					added[n_added++] = pen;
					++other_ext;
				} else {
					pen.ec = pen.ec_high = 0;
					pen.absent = (os < 0 || (os & 0xff) < P.min_cov);
					pen.absent_high = (os < 0 || (os >> 8 & 0xff) < P.min_cov);
					pen.b = b;
					added[n_added++] = pen;
				}
			}
			if (!fixed && other_ext == 0) ++n_fail;
			if (n_fail > n * 2) { rvl = -3; break; }
			if (has_c || n_added == 1) {
				if (n_added > 1 && heap_n > P.max_heap) { // keep only the cheapest extension (first on ties)
					int min_b = -1, min = INT_MAX;
					for (int b = 0; b < n_added; ++b) {
						const int t = pen_weight(P, added[b]);
						if (min > t) min = t, min_b = b;
					}
					if (!push_state(P, heap, heap_n, stack, stack_n, stack_cap, z, added[min_b])) return EC_OVERFLOW;
				} else {
					for (int b = 0; b < n_added; ++b)
						if (!push_state(P, heap, heap_n, stack, stack_n, stack_cap, z, added[b])) return EC_OVERFLOW;
				}
			} else {
				if (n_added == 0) stack[z.k].tot_pen += P.w_absent * (P.max_end_ext - (z.i - end));
				stop = true;
			}
		}
		if (stop) {
			if (stack[z.k].tot_pen < best_pen) best_pen = stack[z.k].tot_pen, best = n_paths;
			paths[n_paths++] = z.k;
			if (n_paths == BFC_MAX_PATHS) break;
		}
	}
	if (n_paths == 0) return rvl;
	// ec[].b := read bases, then the best path (buf_backtrack, correct.c:232-247), then the mask (correct.c:378-379)
	for (int j = 0; j < n; ++j) path[j] = FB_B(rv.fb[j]);
	int n_absent = 0;
	for (int e = paths[best]; e >= 0; e = stack[e].parent) {
		const int i = stack[e].i;
		if (i < n) {
			const int b = stack[e].info & 15;
			path[rv.dir ? n - 1 - i : i] = (uint8_t)(rv.dir ? comp_b(b) : b);
			n_absent += stack[e].info >> 6 & 1;
		}
	}
	for (int i = 0; i < n; ++i)
		if (i < start + k || i >= end) path[rv.dir ? n - 1 - i : i] = 4;
	return n_absent;
}

// reference correct.c:63-80
__device__ int ec_greedy_k(const EcParams &P, const uint64_t x[4], unsigned long long &n_lookups)
{
	const int k = P.k;
	int max = 0, max_ec = -1, max2 = 0;
	for (int i = 0; i < k; ++i) {
		const int c = (int)((x[1] >> i & 1) << 1 | (x[0] >> i & 1));
		for (int j = 0; j < 4; ++j) {
			if (j == c) continue;
			uint64_t y[4] = { x[0], x[1], x[2], x[3] };
			bfc_kmer_change(k, y, i, j);
			const int ret = tab_kmer_occ(P.tab, y);
			++n_lookups;
			if (ret < 0) continue;
			if ((max & 0xff) < (ret & 0xff)) max2 = max, max = ret, max_ec = i << 2 | j;
			else if ((max2 & 0xff) < (ret & 0xff)) max2 = ret;
		}
	}
	return (max & 0xff) * 3 > P.mode && (max2 & 0xff) < 3 ? max_ec : -1;
}

// reference correct.c:82-94
__device__ int ec_first_kmer(int k, const uint8_t *fb, int n, int start, uint64_t x[4])
{
	int i, l = 0;
	x[0] = x[1] = x[2] = x[3] = 0;
	for (i = start; i < n; ++i) {
		const int b = FB_B(fb[i]);
		if (b < 4) {
			bfc_kmer_append(k, x, b);
			if (++l == k) break;
		} else l = 0, x[0] = x[1] = x[2] = x[3] = 0;
	}
	return i;
}

// reference correct.c:388-472 (bfc_ec1) + the packing of worker_ec (correct.c:552-553)
__device__ int ec_read(const EcParams &P, int64_t r, HeapEnt *heap, StackEnt *stack, unsigned long long &n_lookups)
{
	const uint64_t o = P.off[r];
	const int n = (int)(P.off[r + 1] - o - 1), k = P.k;
	uint8_t *seq = P.seq + o, *fb = P.fb + o;
	uint8_t *qual = P.qual && n > 0 && P.qual[o] != 0xFF ? P.qual + o : 0;
	const bool has_q = qual != 0;
	uint32_t ec_code = 1, brute = 0, n_ec = 0, n_ec_high = 0, n_absent = 0, mh = 0;
	int start = 0, end = 0, n_n = 0;

	// bfc_seq_conv (correct.c:23-37)
	for (int i = 0; i < n; ++i) {
		const int b = base_code(seq[i]);
		const int q = b > 3 ? 0 : !has_q ? 1 : (int)qual[i] - 33 >= P.q;
		fb[i] = (uint8_t)(b | (q ? FB_Q : 0));
		n_n += b > 3;
	}
	do {
		if (n_n > n * .05) { ec_code = 2; break; }
		// bfc_ec_kcov (correct.c:96-117): one lookup per k-mer, then the per-base coverage
		{
			uint64_t x[4] = {0, 0, 0, 0};
			int l = 0;
			for (int i = 0; i < n; ++i) {
				const int b = FB_B(fb[i]);
				if (b >= 4) { l = 0, x[0] = x[1] = x[2] = x[3] = 0; continue; }
				bfc_kmer_append(k, x, b);
				if (++l < k) continue;
				const int occ = tab_kmer_occ(P.tab, x);
				++n_lookups;
				if (occ >= 0 && (occ & 0xff) >= P.min_cov)
					fb[i] |= FB_SOLID | ((occ >> 8 & 0x3f) >= P.min_cov + 1 ? FB_HSOLID : 0);
			}
			// lcov[j] = #solid k-mers ending in [j, j+k-1]; hcov likewise for solid && high_end
			int lc = 0, hc = 0;
			for (int j = n - 1; j >= 0; --j) {
				lc += (fb[j] & FB_SOLID) != 0, hc += (fb[j] & FB_HSOLID) != 0;
				if (j + k < n) lc -= (fb[j + k] & FB_SOLID) != 0, hc -= (fb[j + k] & FB_HSOLID) != 0;
				if (lc >= P.min_cov + 1) fb[j] |= FB_LCOV;
				if (4 * hc > 3 * k) fb[j] |= FB_HCOV; // hcov > k * .75
			}
		}
		// bfc_ec_best_island (correct.c:119-130)
		{
			int l = 0, max = 0, max_i = -1, i;
			for (i = k - 1; i < n; ++i) {
				if (!(fb[i] & FB_SOLID)) {
					if (l > max) max = l, max_i = i;
					l = 0;
				} else ++l;
			}
			if (l > max) max = l, max_i = i;
			if (max > 0) start = max_i - max - k + 1, end = max_i;
		}
		if (end == 0 && start == 0) { // no solid k-mer: single-edit rescue (correct.c:405-421)
			// NB: the reference tests the packed (start<<32|end) == 0; start == end == 0 is the only such island
			uint64_t x[4];
			int ec = -1;
			while ((end = ec_first_kmer(k, fb, n, start, x)) < n) {
				ec = ec_greedy_k(P, x, n_lookups);
				if (ec >= 0) break;
				if (end + (k >> 1) >= n) break;
				start = end - (k >> 1);
			}
			if (ec >= 0) {
				fb[end - (ec >> 2)] = (uint8_t)((fb[end - (ec >> 2)] & ~7) | (ec & 3));
				++end; start = end - k;
				brute = 1;
			} else { ec_code = 3; break; }
		}
		RView rv;
		rv.fb = fb, rv.n = n;
		int mh0 = 0, mh1 = 0, rv0, rv1;
		rv.dir = 0;
		rv0 = ec_search(P, rv, start, n, heap, stack, P.stack_cap, P.p0 + o, mh0, n_lookups);
		if (rv0 == EC_OVERFLOW) return EC_OVERFLOW;
		if (rv0 < 0) { ec_code = rv0 == -2 ? 4 : rv0 == -3 ? 5 : 1; break; }
		rv.dir = 1;
		rv1 = ec_search(P, rv, n - end, n, heap, stack, P.stack_cap, P.p1 + o, mh1, n_lookups);
		if (rv1 == EC_OVERFLOW) return EC_OVERFLOW;
		if (rv1 < 0) { ec_code = rv1 == -2 ? 4 : rv1 == -3 ? 5 : 1; break; }
		mh = mh0 > mh1 ? mh0 : mh1;
		ec_code = 0, n_absent = rv0 + rv1;
		// merge the two directions and rewrite the read (correct.c:443-459)
		const uint8_t *p0 = P.p0 + o, *p1 = P.p1 + o;
		for (int i = 0; i < n; ++i) {
			const int f = p0[i], g = p1[i], cur = FB_B(fb[i]), ob = base_code(seq[i]);
			int nb;
			if (f == g) nb = f > 3 ? cur : f;
			else if (g > 3) nb = f;
			else if (f > 3) nb = g;
			else nb = ob;
			const bool diff = nb != ob;
			const int q = (fb[i] & FB_Q) != 0;
			if (diff) { ++n_ec; n_ec_high += q; }
			seq[i] = (uint8_t)((diff ? "acgtn" : "ACGTN")[nb]);
			if (has_q) qual[i] = (uint8_t)(diff ? 34 + ob : (q ? '?' : '+'));
		}
	} while (0);
	P.aux[2 * r] = (n_ec & 0x3fff) << 18 | (n_ec_high & 0x3fff) << 4 | brute << 3 | ec_code;
	P.aux[2 * r + 1] = (n_absent & 0x3fffff) << 10 | 0u << 8 | (mh & 0xff);
	return 0;
}

__global__ void __launch_bounds__(128) k_correct(EcParams P)
{
	const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, n_slots = (int64_t)gridDim.x * blockDim.x;
	HeapEnt *heap = P.heap + slot * P.heap_cap;
	StackEnt *stack = P.stack + slot * P.stack_cap;
	unsigned long long n_lookups = 0;
	for (int64_t s = slot; s < P.n_reads; s += n_slots) {
		const int64_t r = P.redo ? (int64_t)P.redo[s] : s;
		if (ec_read(P, r, heap, stack, n_lookups) == EC_OVERFLOW)
			P.overflow[atomicAdd(P.ctr, 1ULL)] = (uint32_t)r;
	}
	block_add(P.ctr + 1, n_lookups);
}

// ------------------------------------------------------------------ trim

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

// ------------------------------------------------------------------ host side

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static uint64_t batch_bytes_limit()
{
	const char *e = getenv("BFC_B200_EC_BATCH");
	return e && atoll(e) >= 4096 ? (uint64_t)atoll(e) : 1ULL << 27;
}

extern "C" int bfcg_correct_batch(const bfc_opt_t *opt, const bfc_ch_t *ch, int mode, bfcg_batch_t *batch,
                                  uint32_t *aux, bfcg_stats_t *stats)
{
	int r;
	if ((r = bfcg_rt_init()) != BFCG_OK) return r;
	BfcgRuntime &rt = bfcg_rt();
	if (!opt || !ch || !batch || !aux || !batch->off || bfc_ch_get_k(ch) != opt->k || opt->refine_ec || opt->max_heap < 1 || opt->max_heap > 4000)
		return bfcg_fail(__func__, "invalid arguments (refine mode -R is not supported)", cudaSuccess), BFCG_ERR_ARG;
	if (batch->n_reads == 0) return BFCG_OK;
	const bool host = batch->where == BFCG_HOST;
	const int64_t n = batch->n_reads;

	// read offsets on the host (needed to cut launches at read boundaries)
	std::vector<uint64_t> off_copy;
	const uint64_t *h_off = batch->off;
	if (!host) {
		off_copy.resize(n + 1);
		BFCG_CUDA(cudaMemcpyAsync(off_copy.data(), batch->off, (n + 1) * 8, cudaMemcpyDeviceToHost, rt.stream));
		BFCG_CUDA(cudaStreamSynchronize(rt.stream));
		h_off = off_copy.data();
	}
	const uint64_t limit = batch_bytes_limit();
	const int threads = 128;
	const int64_t max_slots = (int64_t)rt.sm_count * 1024;
	const int heap_cap = opt->max_heap + 5; // the search never holds more than max_heap + 4 states
	const int stack_cap0 = 512;

	BfcgTimer timer(stats);
	for (int64_t r0 = 0; r0 < n;) {
		int64_t r1 = r0 + 1;
		while (r1 < n && h_off[r1 + 1] - h_off[r0] <= limit) ++r1;
		const int64_t nr = r1 - r0;
		const uint64_t b0 = h_off[r0], nb = h_off[r1] - b0;
		const int64_t slots = std::min<int64_t>(max_slots, (nr + threads - 1) / threads * threads);
		size_t tot = 0, o_seq = 0, o_qual = 0, o_off = 0, o_aux = 0, o_fb, o_p0, o_p1, o_heap, o_stack, o_ovf, o_ctr;
		if (host) {
			o_seq = tot; tot = align_up(tot + nb, 256);
			o_qual = tot; tot = align_up(tot + nb, 256);
			o_off = tot; tot = align_up(tot + (nr + 1) * 8, 256);
			o_aux = tot; tot = align_up(tot + nr * 8, 256);
		}
		o_fb = tot; tot = align_up(tot + nb, 256);
		o_p0 = tot; tot = align_up(tot + nb, 256);
		o_p1 = tot; tot = align_up(tot + nb, 256);
		o_heap = tot; tot = align_up(tot + (size_t)slots * heap_cap * sizeof(HeapEnt), 256);
		o_stack = tot; tot = align_up(tot + (size_t)slots * stack_cap0 * sizeof(StackEnt), 256);
		o_ovf = tot; tot = align_up(tot + nr * 4, 256);
		o_ctr = tot; tot += 256;
		uint8_t *a = (uint8_t*)bfcg_arena(tot);
		if (!a) return BFCG_ERR_NOMEM;

		EcParams P;
		memset(&P, 0, sizeof(P));
		std::vector<uint64_t> rel;
		if (host) {
			rel.resize(nr + 1);
			for (int64_t i = 0; i <= nr; ++i) rel[i] = h_off[r0 + i] - b0;
			BFCG_CUDA(cudaMemcpyAsync(a + o_seq, batch->seq + b0, nb, cudaMemcpyHostToDevice, rt.stream));
			if (batch->qual) BFCG_CUDA(cudaMemcpyAsync(a + o_qual, batch->qual + b0, nb, cudaMemcpyHostToDevice, rt.stream));
			BFCG_CUDA(cudaMemcpyAsync(a + o_off, rel.data(), (nr + 1) * 8, cudaMemcpyHostToDevice, rt.stream));
			P.off = (const uint64_t*)(a + o_off), P.seq = a + o_seq, P.qual = batch->qual ? a + o_qual : 0;
			P.aux = (uint32_t*)(a + o_aux);
			P.fb = a + o_fb, P.p0 = a + o_p0, P.p1 = a + o_p1;
		} else { // device batch: offsets are absolute, scratch is indexed relative to b0
			P.off = batch->off + r0, P.seq = batch->seq, P.qual = batch->qual;
			P.aux = aux + 2 * r0;
			P.fb = a + o_fb - b0, P.p0 = a + o_p0 - b0, P.p1 = a + o_p1 - b0;
		}
		P.n_reads = nr;
		P.tab = tab_view(ch);
		P.k = opt->k, P.q = opt->q, P.min_cov = opt->min_cov, P.win_multi_ec = opt->win_multi_ec, P.max_end_ext = opt->max_end_ext;
		P.w_ec = opt->w_ec, P.w_ec_high = opt->w_ec_high, P.w_absent = opt->w_absent, P.w_absent_high = opt->w_absent_high;
		P.max_path_diff = opt->max_path_diff, P.max_heap = opt->max_heap, P.mode = mode;
		P.heap = (HeapEnt*)(a + o_heap), P.stack = (StackEnt*)(a + o_stack);
		P.heap_cap = heap_cap, P.stack_cap = stack_cap0;
		P.overflow = (uint32_t*)(a + o_ovf), P.ctr = (unsigned long long*)(a + o_ctr);
		BFCG_CUDA(cudaMemsetAsync(P.ctr, 0, 64, rt.stream));
		{ KTime kt(KT_CORRECT); k_correct<<<(unsigned)(slots / threads), threads, 0, rt.stream>>>(P); }
		BFCG_LAUNCH_CHECK();
		unsigned long long c[2];
		BFCG_CUDA(cudaMemcpyAsync(c, P.ctr, sizeof(c), cudaMemcpyDeviceToHost, rt.stream));
		BFCG_CUDA(cudaStreamSynchronize(rt.stream));
		// reads whose search outgrew the stack: same kernel, fewer threads, larger stacks
		uint32_t *redo = 0;
		StackEnt *big = 0;
		for (int cap = stack_cap0 * 16; c[0] > 0; cap *= 16) {
			const uint64_t n_redo = c[0];
			if (stats) stats->n_redo += n_redo;
			if (cap > (1 << 25)) { cudaFree(redo); cudaFree(big); return bfcg_fail(__func__, "search stack overflow beyond 2^25 entries", cudaSuccess), BFCG_ERR_OVERFLOW; }
			const int64_t rs = std::min<int64_t>(std::min<int64_t>(slots, (int64_t)((n_redo + 31) / 32 * 32)), std::max<int64_t>(32, (int64_t)((8ULL << 30) / ((uint64_t)cap * sizeof(StackEnt))) / 32 * 32));
			cudaFree(redo); cudaFree(big);
			redo = 0, big = 0;
			BFCG_CUDA(cudaMalloc(&redo, n_redo * 4));
			BFCG_CUDA(cudaMalloc(&big, (size_t)rs * cap * sizeof(StackEnt)));
			BFCG_CUDA(cudaMemcpyAsync(redo, P.overflow, n_redo * 4, cudaMemcpyDeviceToDevice, rt.stream));
			BFCG_CUDA(cudaMemsetAsync(P.ctr, 0, 8, rt.stream));
			EcParams Q = P;
			Q.redo = redo, Q.n_reads = (int64_t)n_redo, Q.stack = big, Q.stack_cap = cap;
			{ KTime kt(KT_CORRECT_REDO); k_correct<<<(unsigned)(rs / 32), 32, 0, rt.stream>>>(Q); }
			BFCG_LAUNCH_CHECK();
			BFCG_CUDA(cudaMemcpyAsync(c, P.ctr, sizeof(c), cudaMemcpyDeviceToHost, rt.stream));
			BFCG_CUDA(cudaStreamSynchronize(rt.stream));
		}
		cudaFree(redo); cudaFree(big);
		if (stats) stats->n_lookups += c[1];
		if (host) {
			BFCG_CUDA(cudaMemcpyAsync(batch->seq + b0, a + o_seq, nb, cudaMemcpyDeviceToHost, rt.stream));
			if (batch->qual) BFCG_CUDA(cudaMemcpyAsync(batch->qual + b0, a + o_qual, nb, cudaMemcpyDeviceToHost, rt.stream));
			BFCG_CUDA(cudaMemcpyAsync(aux + 2 * r0, a + o_aux, nr * 8, cudaMemcpyDeviceToHost, rt.stream));
			BFCG_CUDA(cudaStreamSynchronize(rt.stream));
		}
		r0 = r1;
	}
	timer.stop();
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	return BFCG_OK;
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
