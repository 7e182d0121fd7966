// correct.cu -- the correction phase: what the reference runs inside
// kt_for(..., worker_ec, ...) (correct.c:587): bfc_ec1 per read (correct.c:388-472)
// with its two bfc_ec1dir heap searches (correct.c:249-386).
//
// Three kernels per batch:
//   K0 k_enum         (enum.cuh) canonical k-mer hash of every stream position
//   K5 k_ec_lookup    bfc_ec_kcov's one-lookup-per-k-mer (correct.c:106) for the whole
//                     batch at once: a thread does 36 independent table probes, results
//                     go to occ16[position] through shared memory (coalesced both ways)
//   K6 k_ec_read      one read per thread: per-base coverage flags, longest solid
//                     island, rescue, the two heap searches, merge, rewrite.
//
// The search is a chain of dependent table lookups, so throughput comes from reads in
// flight.  Two things keep a search step cheap without changing its result:
//   * the heap lives in global memory, but while it holds a single state (the common,
//     unbranched walk) that state stays in registers -- states enter the memory heap
//     in exactly the reference's push order, so ties pop identically (ksort.h:125-146);
//   * the lookup of the READ's own k-mer at a position (`os`, correct.c:299) is the
//     value K5 already fetched whenever the path has matched the original read for the
//     last k-1 bases; only k-mers that contain an edit are hashed and probed again.
// Each thread owns a fixed heap (max_heap + 4 states suffice: growth stops at
// max_heap, correct.c:349) and a fixed stack; a read whose stack overflows is re-run by
// the same kernel with a larger stack (never on the CPU).
//
// Pop/push order and every threshold follow the reference exactly: the `ec:Z:` tag
// exposes max_heap / n_absent, so the search internals are part of byte parity.
#include "common.cuh"
#include "enum.cuh"
#include <algorithm>
#include <climits>
#include <vector>

#define EC_OVERFLOW (-100)
#define OCC_NONE 0xFFFFu

struct HeapEnt {                 // reference correct.c:153-160 (echeap1_t) + `clean`
	int tot_pen, i, k;
	int ecpos_high[BFC_EC_HIST_HIGH];
	int ecpos[BFC_EC_HIST];
	int clean;                   // trailing path bases equal to the original read (memoised lookups)
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
	uint64_t base0;              // stream offset of the window: occ16 / scratch index = off - base0
	uint32_t *aux;
	const uint16_t *occ16;       // per window position: bfc_ch_kmer_occ of the k-mer ending there, OCC_NONE = absent / no k-mer
	uint8_t *fb, *p0, *p1;       // per-base scratch (window-relative)
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

// ------------------------------------------------------------------ K5: batched k-mer lookups

__global__ void __launch_bounds__(ENUM_THREADS) k_ec_lookup(TabView tab, const unsigned long long *rec_y0, const unsigned long long *rec_y1,
                                                              uint16_t *occ16, uint64_t n_pos, unsigned long long *ctr)
{
	__shared__ uint16_t s_occ[ENUM_SEG];
	const uint64_t seg = blockIdx.x;
	unsigned long long n_lookups = 0;
	for (int j = 0; j < ENUM_CHUNK; ++j) {
		const uint64_t i = seg * ENUM_SEG + (uint64_t)j * ENUM_THREADS + threadIdx.x;
		const unsigned long long y1 = __ldg(rec_y1 + i);
		uint32_t v = OCC_NONE;
		if (y1 != ~0ULL) {
			const int r = tab_get(tab, __ldg(rec_y0 + i) & ~(1ULL << 63), y1);
			if (r >= 0) v = (uint32_t)r;
			++n_lookups;
		}
		s_occ[threadIdx.x * ENUM_CHUNK + j] = (uint16_t)v;
	}
	__syncthreads();
	for (int idx = threadIdx.x; idx < ENUM_SEG; idx += ENUM_THREADS) {
		const uint64_t pos = seg * ENUM_SEG + idx;
		if (pos < n_pos) occ16[pos] = s_occ[idx];
	}
	block_add(ctr + 1, n_lookups);
}

// ------------------------------------------------------------------ K6: per-read search

// view of the read in search coordinates (dir 1 = reverse complement, correct.c:39-57)
struct RView {
	const uint8_t *fb, *seq;
	const uint16_t *occ;
	int n, dir, k;
	__device__ __forceinline__ uint8_t raw(int i) const { return fb[dir ? n - 1 - i : i]; }
	__device__ __forceinline__ int b(int i) const { const int v = FB_B(raw(i)); return dir ? comp_b(v) : v; }
	// the ORIGINAL read base (what K5's k-mers were made of)
	__device__ __forceinline__ int ob(int i) const { const int v = base_code(seq[dir ? n - 1 - i : i]); return dir ? comp_b(v) : v; }
	// K5's value for the read k-mer that ends at search position i
	__device__ __forceinline__ int occ_end(int i) const
	{
		const uint32_t v = occ[dir ? n - 1 - i + k - 1 : i];
		return v == OCC_NONE ? -1 : (int)v;
	}
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

// Logical heap = memory heap[0..heap_n) plus, when `top_valid`, one state held in
// registers (`top`).  Invariant: top_valid implies heap_n == 0, so the logical heap is
// either {top} or the memory heap; states reach the memory heap in push order.
struct SearchState {
	HeapEnt *heap;
	StackEnt *stack;
	int heap_n, stack_n, stack_cap, heap_cap;
	bool top_valid;
	HeapEnt top;
	__device__ __forceinline__ int size() const { return heap_n + (top_valid ? 1 : 0); }
};

// reference correct.c:198-230 (buf_update); false when the stack is full
__device__ __forceinline__ bool push_state(const EcParams &P, SearchState &S, const HeapEnt &prev, const Pen &pen, int ob_prev)
{
	if (S.stack_n >= S.stack_cap || S.heap_n + 2 > S.heap_cap) return false;
	StackEnt q;
	q.parent = prev.k, q.i = prev.i;
	q.info = (uint32_t)pen.b | pen.ec << 4 | pen.ec_high << 5 | pen.absent << 6 | pen.absent_high << 7;
	q.tot_pen = prev.tot_pen + pen_weight(P, pen);
	S.stack[S.stack_n++] = q;
	HeapEnt r;
	r.i = prev.i + 1;
	r.k = S.stack_n - 1;
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
	r.clean = pen.b == ob_prev ? prev.clean + 1 : 0;
	bfc_kmer_append(P.k, r.x, pen.b);
	if (S.heap_n == 0 && !S.top_valid) { S.top = r; S.top_valid = true; return true; }
	if (S.top_valid) { // a second state arrives: the register state goes to memory first (same order as the reference)
		S.heap[S.heap_n++] = S.top;
		heap_up(S.heap, S.heap_n);
		S.top_valid = false;
	}
	S.heap[S.heap_n++] = r;
	heap_up(S.heap, S.heap_n);
	return true;
}

// reference correct.c:249-386 (bfc_ec1dir).  `path` receives ec[].b in FORWARD read
// coordinates (for dir 1 that is the result after the reference's final revcomp).
__device__ int ec_search(const EcParams &P, const RView &rv, int start, int end, SearchState &S, uint8_t *path,
                         int &max_heap, bool memo, unsigned long long &n_lookups)
{
	const int k = P.k, n = rv.n;
	HeapEnt z;
	int rvl = -1, n_paths = 0, best = -1, best_pen = INT_MAX, n_fail = 0, run = 0;
	int paths[BFC_MAX_PATHS];
	S.heap_n = S.stack_n = 0, S.top_valid = false;
	max_heap = 0;
	z.tot_pen = 0, z.k = -1, z.clean = 0;
	z.x[0] = z.x[1] = z.x[2] = z.x[3] = 0;
	for (z.i = start; z.i < end; ++z.i) { // seed: k-1 bases (correct.c:260-267)
		const int c = rv.b(z.i);
		if (c < 4) {
			if (++run == k) break;
			bfc_kmer_append(k, z.x, c);
			z.clean = c == rv.ob(z.i) ? z.clean + 1 : 0;
		} else run = 0, z.clean = 0, z.x[0] = z.x[1] = z.x[2] = z.x[3] = 0;
	}
	if (z.i >= end) return -1; // the reference asserts; cannot happen after an island / rescue was found
#pragma unroll
	for (int t = 0; t < BFC_EC_HIST; ++t) z.ecpos[t] = -1;
#pragma unroll
	for (int t = 0; t < BFC_EC_HIST_HIGH; ++t) z.ecpos_high[t] = -1;
	S.top = z, S.top_valid = true;

	for (;;) {
		bool stop = false;
		{
			const int hs = S.size();
			max_heap = max_heap > 255 ? 255 : max_heap > hs ? max_heap : hs;
			if (hs == 0) { rvl = -2; break; }
		}
		if (S.top_valid) { z = S.top; S.top_valid = false; }
		else {
			z = S.heap[0];
			S.heap[0] = S.heap[--S.heap_n];
			heap_down(S.heap, S.heap_n);
		}
		if (best >= 0 && z.tot_pen > best_pen + P.max_path_diff) break;
		if (z.i - end > P.max_end_ext) stop = true;
		if (!stop) {
			const bool has_c = z.i < n;
			const uint8_t craw = has_c ? rv.raw(z.i) : 0;
			const int cb = has_c ? (rv.dir ? comp_b(FB_B(craw)) : FB_B(craw)) : -1;
			const int cq = (craw & FB_Q) != 0;
			const int cob = has_c ? rv.ob(z.i) : -1;
			int os = -1, other_ext = 0, n_added = 0;
			bool fixed = z.i > end;
			Pen added[4];
			if (has_c && cb < 4) {
				if (memo && z.clean >= k - 1 && cb == cob) os = rv.occ_end(z.i); // the read's own k-mer: fetched by K5
				else {
					uint64_t x[4] = { z.x[0], z.x[1], z.x[2], z.x[3] };
					bfc_kmer_append(k, x, cb);
					os = tab_kmer_occ(P.tab, x);
					++n_lookups;
				}
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
				if (n_added > 1 && S.size() > P.max_heap) { // keep only the cheapest extension (first on ties)
					int min_b = -1, min = INT_MAX;
					for (int b = 0; b < n_added; ++b) {
						const int t = pen_weight(P, added[b]);
						if (min > t) min = t, min_b = b;
					}
					if (!push_state(P, S, z, added[min_b], cob)) return EC_OVERFLOW;
				} else {
					for (int b = 0; b < n_added; ++b)
						if (!push_state(P, S, z, added[b], cob)) return EC_OVERFLOW;
				}
			} else {
				if (n_added == 0) S.stack[z.k].tot_pen += P.w_absent * (P.max_end_ext - (z.i - end));
				stop = true;
			}
		}
		if (stop) {
			if (S.stack[z.k].tot_pen < best_pen) best_pen = S.stack[z.k].tot_pen, best = n_paths;
			paths[n_paths++] = z.k;
			if (n_paths == BFC_MAX_PATHS) break;
		}
	}
	if (n_paths == 0) return rvl;
	// ec[].b := read bases, then the best path (buf_backtrack, correct.c:232-247), then the mask (correct.c:378-379)
	for (int j = 0; j < n; ++j) path[j] = FB_B(rv.fb[j]);
	int n_absent = 0;
	for (int e = paths[best]; e >= 0; e = S.stack[e].parent) {
		const int i = S.stack[e].i;
		if (i < n) {
			const int b = S.stack[e].info & 15;
			path[rv.dir ? n - 1 - i : i] = (uint8_t)(rv.dir ? comp_b(b) : b);
			n_absent += S.stack[e].info >> 6 & 1;
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
__device__ int ec_read(const EcParams &P, int64_t r, SearchState &S, unsigned long long &n_lookups)
{
	const uint64_t o = P.off[r], ow = o - P.base0;
	const int n = (int)(P.off[r + 1] - o - 1), k = P.k;
	uint8_t *seq = P.seq + o, *fb = P.fb + ow;
	uint8_t *qual = P.qual && n > 0 && P.qual[o] != 0xFF ? P.qual + o : 0;
	const uint16_t *occ = P.occ16 + ow;
	const bool has_q = qual != 0;
	uint32_t ec_code = 1, brute = 0, n_ec = 0, n_ec_high = 0, n_absent = 0, mh = 0;
	int start = 0, end = 0, n_n = 0;
	bool island = false;

	// bfc_seq_conv (correct.c:23-37) + the solid / high flags of bfc_ec_kcov (correct.c:106-108)
	for (int i = 0; i < n; ++i) {
		const int b = base_code(seq[i]);
		const int q = b > 3 ? 0 : !has_q ? 1 : (int)qual[i] - 33 >= P.q;
		uint32_t f = (uint32_t)b | (q ? FB_Q : 0);
		const uint32_t v = occ[i];
		if (v != OCC_NONE && (int)(v & 0xff) >= P.min_cov)
			f |= FB_SOLID | ((int)(v >> 8 & 0x3f) >= P.min_cov + 1 ? FB_HSOLID : 0);
		fb[i] = (uint8_t)f;
		n_n += b > 3;
	}
	do {
		if (n_n > n * .05) { ec_code = 2; break; }
		{ // lcov[j] = #solid k-mers ending in [j, j+k-1]; hcov likewise for solid && high_end (correct.c:109-112)
			int lc = 0, hc = 0;
			for (int j = n - 1; j >= 0; --j) {
				lc += (fb[j] & FB_SOLID) != 0, hc += (fb[j] & FB_HSOLID) != 0;
				if (j + k < n) lc -= (fb[j + k] & FB_SOLID) != 0, hc -= (fb[j + k] & FB_HSOLID) != 0;
				if (lc >= P.min_cov + 1) fb[j] |= FB_LCOV;
				if (4 * hc > 3 * k) fb[j] |= FB_HCOV; // hcov > k * .75
			}
		}
		{ // bfc_ec_best_island (correct.c:119-130)
			int l = 0, max = 0, max_i = -1, i;
			for (i = k - 1; i < n; ++i) {
				if (!(fb[i] & FB_SOLID)) {
					if (l > max) max = l, max_i = i;
					l = 0;
				} else ++l;
			}
			if (l > max) max = l, max_i = i;
			if (max > 0) start = max_i - max - k + 1, end = max_i, island = true;
		}
		if (!island) { // no solid k-mer: single-edit rescue (correct.c:405-421)
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
		rv.fb = fb, rv.seq = seq, rv.occ = occ, rv.n = n, rv.k = k;
		int mh0 = 0, mh1 = 0, rv0, rv1;
		rv.dir = 0;
		rv0 = ec_search(P, rv, start, n, S, P.p0 + ow, mh0, true, n_lookups);
		if (rv0 == EC_OVERFLOW) return EC_OVERFLOW;
		if (rv0 < 0) { ec_code = rv0 == -2 ? 4 : rv0 == -3 ? 5 : 1; break; }
		rv.dir = 1;
		// the reverse-complement k-mer hashes like the forward one only for odd k (kmer.h:81)
		rv1 = ec_search(P, rv, n - end, n, S, P.p1 + ow, mh1, (k & 1) != 0, n_lookups);
		if (rv1 == EC_OVERFLOW) return EC_OVERFLOW;
		if (rv1 < 0) { ec_code = rv1 == -2 ? 4 : rv1 == -3 ? 5 : 1; break; }
		mh = mh0 > mh1 ? mh0 : mh1;
		ec_code = 0, n_absent = rv0 + rv1;
		// merge the two directions and rewrite the read (correct.c:443-459)
		const uint8_t *p0 = P.p0 + ow, *p1 = P.p1 + ow;
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

__global__ void __launch_bounds__(128) k_ec_read(EcParams P)
{
	const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, n_slots = (int64_t)gridDim.x * blockDim.x;
	SearchState S;
	S.heap = P.heap + slot * P.heap_cap, S.stack = P.stack + slot * P.stack_cap;
	S.heap_cap = P.heap_cap, S.stack_cap = P.stack_cap;
	S.heap_n = S.stack_n = 0, S.top_valid = false;
	unsigned long long n_lookups = 0;
	for (int64_t s = slot; s < P.n_reads; s += n_slots) {
		const int64_t r = P.redo ? (int64_t)P.redo[s] : s;
		if (ec_read(P, r, S, n_lookups) == EC_OVERFLOW)
			P.overflow[atomicAdd(P.ctr, 1ULL)] = (uint32_t)r;
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
	const int heap_cap = opt->max_heap + 6; // the search never holds more than max_heap + 4 states
	const int stack_cap0 = 512;

	BfcgTimer timer(stats);
	for (int64_t r0 = 0; r0 < n;) {
		int64_t r1 = r0 + 1;
		while (r1 < n && h_off[r1 + 1] - h_off[r0] <= limit) ++r1;
		const int64_t nr = r1 - r0;
		const uint64_t b0 = h_off[r0], nb = h_off[r1] - b0;
		const uint64_t n_rec = enum_padded(nb);
		const int64_t slots = std::min<int64_t>(max_slots, (nr + threads - 1) / threads * threads);
		size_t tot = 0, o_seq = 0, o_qual = 0, o_off = 0, o_aux = 0, o_fb, o_p0, o_p1, o_y0, o_y1, o_occ, o_heap, o_stack, o_ovf, o_ctr;
		if (host) {
			o_seq = tot; tot = align_up(tot + nb, 256);
			o_qual = tot; tot = align_up(tot + nb, 256);
			o_off = tot; tot = align_up(tot + (nr + 1) * 8, 256);
			o_aux = tot; tot = align_up(tot + nr * 8, 256);
		}
		o_fb = tot; tot = align_up(tot + nb, 256);
		o_p0 = tot; tot = align_up(tot + nb, 256);
		o_p1 = tot; tot = align_up(tot + nb, 256);
		o_y0 = tot; tot = align_up(tot + n_rec * 8, 256);
		o_y1 = tot; tot = align_up(tot + n_rec * 8, 256);
		o_occ = tot; tot = align_up(tot + n_rec * 2, 256);
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
			P.base0 = 0;
		} else { // device batch: offsets are absolute, the window starts at b0
			P.off = batch->off + r0, P.seq = batch->seq, P.qual = batch->qual;
			P.aux = aux + 2 * r0;
			P.base0 = b0;
		}
		P.fb = a + o_fb, P.p0 = a + o_p0, P.p1 = a + o_p1;
		P.occ16 = (const uint16_t*)(a + o_occ);
		P.n_reads = nr;
		P.tab = tab_view(ch);
		P.k = opt->k, P.q = opt->q, P.min_cov = opt->min_cov, P.win_multi_ec = opt->win_multi_ec, P.max_end_ext = opt->max_end_ext;
		P.w_ec = opt->w_ec, P.w_ec_high = opt->w_ec_high, P.w_absent = opt->w_absent, P.w_absent_high = opt->w_absent_high;
		P.max_path_diff = opt->max_path_diff, P.max_heap = opt->max_heap, P.mode = mode;
		P.heap = (HeapEnt*)(a + o_heap), P.stack = (StackEnt*)(a + o_stack);
		P.heap_cap = heap_cap, P.stack_cap = stack_cap0;
		P.overflow = (uint32_t*)(a + o_ovf), P.ctr = (unsigned long long*)(a + o_ctr);
		BFCG_CUDA(cudaMemsetAsync(P.ctr, 0, 64, rt.stream));

		EnumParams ep;
		memset(&ep, 0, sizeof(ep));
		ep.seq = (host ? a + o_seq : batch->seq + b0), ep.qual = 0; // the quality flag of a record is not used here
		ep.len = nb, ep.emit_from = 0, ep.k = opt->k, ep.q = opt->q;
		ep.rec_y0 = (unsigned long long*)(a + o_y0), ep.rec_y1 = (unsigned long long*)(a + o_y1);
		{ KTime kt(KT_ENUM); k_enum<<<(unsigned)(n_rec / ENUM_SEG), ENUM_THREADS, 0, rt.stream>>>(ep); }
		BFCG_LAUNCH_CHECK();
		{ KTime kt(KT_EC_LOOKUP); k_ec_lookup<<<(unsigned)(n_rec / ENUM_SEG), ENUM_THREADS, 0, rt.stream>>>(P.tab, ep.rec_y0, ep.rec_y1, (uint16_t*)(a + o_occ), nb, P.ctr); }
		BFCG_LAUNCH_CHECK();
		{ KTime kt(KT_CORRECT); k_ec_read<<<(unsigned)(slots / threads), threads, 0, rt.stream>>>(P); }
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
			{ KTime kt(KT_CORRECT_REDO); k_ec_read<<<(unsigned)(rs / 32), 32, 0, rt.stream>>>(Q); }
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
