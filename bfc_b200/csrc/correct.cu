// correct.cu -- the correction phase: what the reference runs inside
// kt_for(..., worker_ec, ...) (correct.c:587): bfc_ec1 per read (correct.c:388-472)
// with its two bfc_ec1dir heap searches (correct.c:249-386).
//
// The read stream of a batch window is turned into BIT PLANES (one bit per stream
// position) and a 16-bit flag word per base; the search then works on those:
//
//   K5  k_ec_lookup  bfc_ec_kcov's one-lookup-per-k-mer (correct.c:106) for the whole window at
//                    once: the window is staged as bit planes in shared memory, the k-mer ending
//                    at a position is cut out of them, hashed and looked up -> planes B0 B1 NB Q
//                    (base bits, non-ACGT, Q >= q) and SOL HS A H (what the search ever asks
//                    about a read k-mer's table value)
//   K5b k_ec_cov     lcov / hcov thresholds (correct.c:109-112) as sliding popcounts
//                    over SOL / HS -> the per-base flag word, and the two "jump" planes
//                    (positions a lone search state steps over without any decision)
//   K6a k_ec_setup   one read per thread: >5 % N, longest solid island (correct.c:119-130)
//                    by run scanning, the rare single-edit rescue (correct.c:63-94, 405-421)
//   K6a' k_ec_ext    the lookups past the end of every read (they depend on its last k-1 bases only)
//   K6b k_ec_search  the heap search, one (read, direction) JOB per thread at a time,
//                    persistent threads, jobs handed out dynamically.  The thread is a resumable state machine whose
//                    only table lookup sits at ONE place in the loop, so the lanes of
//                    a warp hash and probe together whatever step each of them is in.
//   K6c k_ec_merge   one read per warp: merge the two directions, rewrite seq / qual
//                    (correct.c:443-459), pack aux / aux2 (correct.c:552-553); coalesced.
//
// What keeps a search step cheap without changing its result:
//   * the lookup of the READ's own k-mer at a position (`os`, correct.c:299) is what K5
//     fetched whenever the path has matched the original read for the last k-1 bases;
//   * while the search holds a single state and the next positions are "fixed" with no
//     penalty, the state moves over all of them at once (count-trailing-ones on the
//     jump plane) and its k-mer is re-extracted from the base planes in O(1);
//   * the heap orders 4-byte keys (penalty, slot); states sit in a slot pool; a lone
//     state never leaves registers.  States enter the heap in exactly the reference's
//     push order and the sift rules are klib's (ksort.h:125-146), so ties pop identically;
//   * the reference's per-push stack (correct.c:162-167) is only ever read to recover
//     the edited bases, the absent count and the final penalty of a path: states carry
//     the latter two, and only pushes that CHANGE a base get a (parent-linked) entry.
// A job whose edit list overflows its fixed scratch is re-run by the same kernel with a
// larger one (never on the CPU).
//
// Pop/push order and every threshold follow the reference exactly: the `ec:Z:` tag
// exposes max_heap / n_absent, so the search internals are part of byte parity.
#include "common.cuh"
#include "enum.cuh"
#include <algorithm>
#include <climits>
#include <vector>

// plane indices (bit i of a plane = stream position i - PL_PAD of the window)
enum { PL_B0 = 0, PL_B1, PL_NB, PL_Q, PL_SOL, PL_HS, PL_A, PL_H, PL_J0, PL_J1,
       PL_E0V, PL_E0L, PL_E0H, PL_E1V, PL_E1L, PL_E1H, PL_N };
#define PL_PAD 64  // leading zero bits so that "64 bits ending at position p" never underflows

// per-base flag word (k_ec_cov)
#define FL_OB(v)   ((v) & 7)   // original base code 0..4
#define FL_Q       8           // Q >= q (0 for non-ACGT)
#define FL_LC      16          // lcov >= min_cov + 1
#define FL_HC      32          // hcov > 0.75 k
#define FL_SOL     64          // the k-mer ENDING here: in the table with cnt >= min_cov
#define FL_A       128         //   (os & 0xff) >= min_cov + 1, os = -1 counting as 255 (correct.c:299-300)
#define FL_H       256         //   in the table with high count >= min_cov

struct ReadDesc {              // k_ec_setup -> k_ec_search / k_ec_merge
	int start0, start1;        // search starts of the two directions; start0 < 0: no search
	int brute;                 // rescue edit: pos << 2 | base, or -1
	int code;                  // ec_code decided by the setup (2, 3) or 0
};

struct EcState {               // reference correct.c:153-160 (echeap1_t); see the header for `edit` / `n_absent`
	int tot_pen, i;
	int edit;                  // most recent edit entry on this path, -1 = none
	int n_absent;              // pushes with pen.absent at positions < n along this path
	int clean;                 // trailing path bases equal to the ORIGINAL read (memoised lookups)
	int ecpos_high[BFC_EC_HIST_HIGH];
	int ecpos[BFC_EC_HIST];
	uint64_t x[4];
};

struct EcParams {
	const uint64_t *off;
	uint8_t *seq, *qual;
	int64_t n_reads;
	uint64_t base0;              // stream offset of the window: plane / flag index = off - base0
	uint32_t *aux;
	const uint64_t *pl;          // PL_N planes of pl_words 64-bit words each
	uint64_t pl_words;
	const uint16_t *fl;          // per window position
	ReadDesc *desc;
	int4 *jobs;                  // per job: (window offset of the read, length, search start or -1, rescue edit or -1)
	int2 *res;                   // per job: (return value of bfc_ec1dir, max_heap)
	uint64_t *ext;               // per job: the lookups past the end of the read (k_ec_ext), 0 = not available
	TabView tab;
	int k, q, min_cov, win_multi_ec, max_end_ext;
	int w_ec, w_ec_high, w_absent, w_absent_high, max_path_diff, max_heap, mode;
	int refine;                  // -R: aux holds the earlier stats on entry (correct.c:438-442, 470)
	EcState *pool;               // heap_cap states per thread slot
	uint32_t *heapk;             // heap_cap keys per thread slot: tot_pen << 12 | pool slot
	uint2 *edits;                // edit_cap entries per thread slot: (parent, pos << 3 | base), forward coordinates
	int heap_cap, edit_cap;
	const uint32_t *redo;        // when set: thread job list (job = read * 2 + dir); n_jobs = its length
	int64_t n_jobs;
	uint32_t *overflow;          // jobs whose edit list overflowed
	unsigned long long *ctr;     // [0] n_overflow, [1] n_lookups, [2] next job to hand out, [3] lookups of k_ec_search
};

__device__ __forceinline__ int comp_b(int b) { return b < 4 ? 3 - b : 4; }

__device__ __forceinline__ const uint64_t *plane(const EcParams &P, int which) { return P.pl + (uint64_t)which * P.pl_words; }

// 64 plane bits starting at window position `pos` (pos >= -PL_PAD)
__device__ __forceinline__ uint64_t bits64(const uint64_t *pl, int64_t pos)
{
	const uint64_t p = (uint64_t)(pos + PL_PAD);
	const uint64_t *w = pl + (p >> 6);
	const int r = (int)(p & 63);
	const uint64_t lo = __ldg(w), hi = __ldg(w + 1);
	return r ? (lo >> r) | (hi << (64 - r)) : lo;
}

__device__ __forceinline__ int bit1(const uint64_t *pl, int64_t pos)
{
	const uint64_t p = (uint64_t)(pos + PL_PAD);
	return (int)(__ldg(pl + (p >> 6)) >> (p & 63)) & 1;
}

// ------------------------------------------------------------------ K5: batched k-mer lookups -> planes

struct LookupParams {
	TabView tab;
	const uint8_t *seq, *qual;
	uint64_t n_pos;
	uint64_t *pl;
	uint64_t pl_words;
	int k, q, min_cov;
	int refine;                  // -R: corrected bases are taken back from the quality string (correct.c:31)
	unsigned long long *ctr;
};

// One CTA per EL_SEG stream positions: the window is staged as bit planes in shared memory (enum.cuh), the k-mer
// ending at each position is cut out of them, hashed and looked up (bfc_ec_kcov's lookup, correct.c:106), and the
// four things the search ever asks about a read k-mer's table value leave as plane words next to the base planes.
__global__ void __launch_bounds__(EL_THREADS) k_ec_lookup(LookupParams p)
{
	__shared__ uint32_t s_pl[4][EL_WORDS];
	const int64_t seg0 = (int64_t)blockIdx.x * EL_SEG;
	el_stage_planes(s_pl, p.seq, p.qual, p.n_pos, seg0, p.q, p.refine != 0);
	uint32_t *pl32 = (uint32_t*)p.pl;
	const uint64_t stride32 = p.pl_words * 2, w0 = (uint64_t)(seg0 + PL_PAD) / 32;
	// the base planes of the segment: shared-memory word 2 + t = positions seg0 + 32 t ..
	{
		const int t = threadIdx.x;
		pl32[PL_B0 * stride32 + w0 + t] = s_pl[0][2 + t], pl32[PL_B1 * stride32 + w0 + t] = s_pl[1][2 + t];
		pl32[PL_NB * stride32 + w0 + t] = s_pl[2][2 + t], pl32[PL_Q * stride32 + w0 + t] = s_pl[3][2 + t];
	}
	const int k = p.k;
	const uint64_t kmask = (1ULL << k) - 1;
	unsigned long long n_lookups = 0;
#pragma unroll 2
	for (int j = 0; j < EL_ITERS; ++j) {
		const uint32_t pp = (uint32_t)(j * EL_THREADS + threadIdx.x);
		uint32_t f = 4; // bit 0 SOL, 1 HS, 2 A, 3 H; no k-mer or an absent one counts as "A" (os = -1 -> 255, correct.c:299-300)
		uint64_t y[2];
		if (el_kmer_at(s_pl, pp + EL_LEAD - (uint32_t)(k - 1), k, kmask, y)) {
			const int r = tab_get(p.tab, y[0], y[1]);
			if (r >= 0) {
				const int cnt = r & 0xff, high = r >> 8 & 0x3f;
				f = (cnt >= p.min_cov ? 1u : 0u) | (cnt >= p.min_cov && high >= p.min_cov + 1 ? 2u : 0u) |
				    (cnt >= p.min_cov + 1 ? 4u : 0u) | (high >= p.min_cov ? 8u : 0u);
			}
			++n_lookups;
		}
		const uint32_t sol = __ballot_sync(0xffffffffu, f & 1), hs = __ballot_sync(0xffffffffu, f & 2);
		const uint32_t fa = __ballot_sync(0xffffffffu, f & 4), fh = __ballot_sync(0xffffffffu, f & 8);
		if ((threadIdx.x & 31) == 0) {
			uint32_t *w = pl32 + w0 + (pp >> 5);
			w[PL_SOL * stride32] = sol, w[PL_HS * stride32] = hs, w[PL_A * stride32] = fa, w[PL_H * stride32] = fh;
		}
	}
	block_add(p.ctr + 1, n_lookups);
}

// ------------------------------------------------------------------ K5b: coverage flags + jump planes

// reference correct.c:109-112: lcov[j] / hcov[j] = number of solid (solid && high_end)
// k-mers covering base j = those ENDING in [j, j+k-1].  A window never picks up bits of
// the next read: no k-mer ends on a terminator or on the first k-1 bases of a read.
__global__ void __launch_bounds__(256) k_ec_cov(EcParams P, uint64_t n_pos, uint16_t *fl, uint64_t *pl_out)
{
	// one thread per plane word = 32 stream positions p0 .. p0 + 31 (n_pos is a multiple of 32): everything is
	// word arithmetic on the planes; the only per-position work is the two sliding popcounts
	const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (w * 32 >= n_pos) return;
	const int k = P.k;
	const uint32_t *pl32 = (const uint32_t*)P.pl;
	const uint64_t s32 = P.pl_words * 2, g = w + PL_PAD / 32;
#define PLW(which, d) __ldg(pl32 + (uint64_t)(which) * s32 + g + (d))
	const uint32_t b0 = PLW(PL_B0, 0), b1 = PLW(PL_B1, 0), nb = PLW(PL_NB, 0), q = PLW(PL_Q, 0);
	const uint32_t s0 = PLW(PL_SOL, 0), s1 = PLW(PL_SOL, 1), s2 = PLW(PL_SOL, 2);
	const uint32_t t0 = PLW(PL_HS, 0), t1 = PLW(PL_HS, 1), t2 = PLW(PL_HS, 2);
	const uint32_t a0 = PLW(PL_A, 0), a1 = PLW(PL_A, 1), a2 = PLW(PL_A, 2);
	const uint32_t h0 = PLW(PL_H, 0), h1 = PLW(PL_H, 1), h2 = PLW(PL_H, 2);
#undef PLW
	const uint64_t kmask = (1ULL << k) - 1;
	const uint64_t slo = s0 | (uint64_t)s1 << 32, tlo = t0 | (uint64_t)t1 << 32;
	// lcov[p] / hcov[p] = solid (solid && high_end) k-mers ENDING in [p, p + k - 1] (correct.c:109-112)
	int cs = __popcll(slo & kmask), ct = __popcll(tlo & kmask);
	uint32_t lc = 0, hc = 0;
#pragma unroll
	for (int i = 0; i < 32; ++i) {
		lc |= (uint32_t)(cs >= P.min_cov + 1) << i;
		hc |= (uint32_t)(4 * ct > 3 * k) << i; // hcov > k * .75
		const int j = i + k; // the bit entering the window; bit i leaves it
		cs += (int)(j < 64 ? slo >> j & 1 : (uint64_t)(s2 >> (j - 64) & 1)) - (int)(slo >> i & 1);
		ct += (int)(j < 64 ? tlo >> j & 1 : (uint64_t)(t2 >> (j - 64) & 1)) - (int)(tlo >> i & 1);
	}
	// a lone state whose last k-1 bases are the read's steps over a base without a decision when the base is
	// "fixed" (correct.c:299-301) and its own k-mer costs nothing (correct.c:334-336)
	const uint32_t fixed = ~nb & ((q & lc) | hc);
	const uint32_t j0 = ~nb & ((q & a0 & lc) | hc) & s0 & h0;
	// reverse direction: the k-mer ending (in search order) on a base ends k-1 positions further in the stream
	const int sh = k - 1;
	const uint64_t alo = a0 | (uint64_t)a1 << 32, hlo = h0 | (uint64_t)h1 << 32;
	const uint32_t sr = sh ? (uint32_t)((slo >> sh) | ((uint64_t)s2 << (64 - sh))) : s0;
	const uint32_t ar = sh ? (uint32_t)((alo >> sh) | ((uint64_t)a2 << (64 - sh))) : a0;
	const uint32_t hr = sh ? (uint32_t)((hlo >> sh) | ((uint64_t)h2 << (64 - sh))) : h0;
	const uint32_t j1 = ~nb & ((q & ar & lc) | hc) & sr & hr;
	(void)fixed;
	uint32_t *jw = (uint32_t*)pl_out + g;
	jw[(uint64_t)PL_J0 * s32] = j0, jw[(uint64_t)PL_J1 * s32] = j1;
	// the flag word of every base
	uint4 *out = (uint4*)(fl + w * 32);
#pragma unroll
	for (int v = 0; v < 4; ++v) {
		uint32_t o[4];
#pragma unroll
		for (int u = 0; u < 4; ++u) {
			uint32_t pair = 0;
#pragma unroll
			for (int e = 0; e < 2; ++e) {
				const int i = v * 8 + u * 2 + e;
				const uint32_t n = nb >> i & 1;
				const uint32_t f = (n ? 4u : (b0 >> i & 1) | (b1 >> i & 1) << 1) | (q >> i & 1 ? (uint32_t)FL_Q : 0u) | (lc >> i & 1 ? (uint32_t)FL_LC : 0u) |
				                   (hc >> i & 1 ? (uint32_t)FL_HC : 0u) | (s0 >> i & 1 ? (uint32_t)FL_SOL : 0u) | (a0 >> i & 1 ? (uint32_t)FL_A : 0u) | (h0 >> i & 1 ? (uint32_t)FL_H : 0u);
				pair |= f << (16 * e);
			}
			o[u] = pair;
		}
		out[v] = make_uint4(o[0], o[1], o[2], o[3]);
	}
}

// ------------------------------------------------------------------ k-mer extraction from the base planes

// The 4-plane k-mer state (kmer.h:10-17) after appending, to an empty state, the m <= 63
// ORIGINAL read bases that end at search position e of direction dir (read at window
// position o, length n).  `bp`/`bb`: optional base override (rescue edit) in forward coordinates.
__device__ __forceinline__ void extract_kmer(const EcParams &P, int64_t o, int n, int dir, int e, int m, int bp, int bb, uint64_t x[4])
{
	const int ws = dir ? n - 1 - e : e - m + 1; // forward position of the lowest window bit
	const uint64_t mm = m >= 64 ? ~0ULL : (1ULL << m) - 1;
	uint64_t w0 = bits64(plane(P, PL_B0), o + ws) & mm, w1 = bits64(plane(P, PL_B1), o + ws) & mm;
	if (bp >= ws && bp < ws + m) {
		const int t = bp - ws;
		w0 = (w0 & ~(1ULL << t)) | (uint64_t)(bb & 1) << t;
		w1 = (w1 & ~(1ULL << t)) | (uint64_t)(bb >> 1) << t;
	}
	const uint64_t r0 = m ? __brevll(w0) >> (64 - m) : 0, r1 = m ? __brevll(w1) >> (64 - m) : 0;
	const uint64_t c0 = ~w0 & mm, c1 = ~w1 & mm;
	const int up = P.k - m;
	if (!dir) x[0] = r0, x[1] = r1, x[2] = c0 << up, x[3] = c1 << up;
	else x[0] = c0, x[1] = c1, x[2] = r0 << up, x[3] = r1 << up;
}

// ------------------------------------------------------------------ K6a: per-read setup

// reference correct.c:82-94: position of the k-th base of the first run of k ACGT bases at or after `start` (n if none)
__device__ int ec_first_kmer_end(const EcParams &P, int64_t o, int n, int start)
{
	const int k = P.k;
	const uint64_t kmask = (1ULL << k) - 1;
	int p = start;
	while (p + k <= n) {
		const uint64_t w = bits64(plane(P, PL_NB), o + p) & kmask;
		if (w == 0) return p + k - 1;
		p += 64 - __clzll(w); // just past the last non-ACGT base of the window
	}
	return n;
}

__global__ void __launch_bounds__(128) k_ec_setup(EcParams P)
{
	const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long n_lookups = 0;
	if (r < P.n_reads) {
		const uint64_t ob = P.off[r];
		const int64_t o = (int64_t)(ob - P.base0);
		const int n = (int)(P.off[r + 1] - ob - 1), k = P.k;
		ReadDesc d;
		d.start0 = d.start1 = -1, d.brute = -1, d.code = 0;
		int n_n = 0;
		for (int p = 0; p < n; p += 64) {
			const int len = n - p < 64 ? n - p : 64;
			n_n += __popcll(bits64(plane(P, PL_NB), o + p) & (len == 64 ? ~0ULL : (1ULL << len) - 1));
		}
		if (n_n > n * .05) d.code = 2; // correct.c:397-402
		else {
			// bfc_ec_best_island (correct.c:119-130): longest run of solid k-mer ends in [k-1, n), first wins
			int l = 0, max = 0, max_i = -1;
			for (int p = k - 1; p < n; p += 64) {
				const int len = n - p < 64 ? n - p : 64;
				const uint64_t w = bits64(plane(P, PL_SOL), o + p) & (len == 64 ? ~0ULL : (1ULL << len) - 1);
				int c = 0;
				while (c < len) {
					const uint64_t v = w >> c;
					if (v & 1) {
						const uint64_t nv = ~v;
						int ones = nv ? __ffsll((long long)nv) - 1 : 64;
						if (ones > len - c) ones = len - c;
						l += ones, c += ones;
					} else {
						if (l > max) max = l, max_i = p + c;
						l = 0;
						int zeros = v ? __ffsll((long long)v) - 1 : 64;
						if (zeros > len - c) zeros = len - c;
						c += zeros;
					}
				}
			}
			if (l > max) max = l, max_i = n;
			if (max > 0) d.start0 = max_i - max - k + 1, d.start1 = n - max_i;
			else { // no solid k-mer: the single-edit rescue (correct.c:405-421) tries 3k edits per k-mer -- a warp's job (k_ec_rescue)
				P.overflow[atomicAdd(P.ctr + 4, 1ULL)] = (uint32_t)r; // (the search's overflow list is not in use yet)
				d.code = 3;                                           // until the rescue finds an edit
			}
		}
		P.desc[r] = d;
		P.jobs[2 * r] = make_int4((int)o, n, d.start0, d.brute);
		P.jobs[2 * r + 1] = make_int4((int)o, n, d.start0 < 0 ? -1 : d.start1, d.brute);
	}
	block_add(P.ctr + 1, n_lookups);
}

// ------------------------------------------------------------------ K6a+: the single-edit rescue, one read per warp

// bfc_ec_greedy_k (correct.c:63-80) tries every single-base edit of a k-mer -- 3k lookups -- and bfc_ec1 walks the read
// in steps of k/2 until an edit is accepted (correct.c:405-421).  The lanes of a warp share the 3k lookups of one k-mer;
// what the sequential loop keeps -- the FIRST edit (in its i, j order) with the largest count, and the second-largest
// count, equal counts included -- is order-free apart from that "first", so it reduces across lanes exactly.
__global__ void __launch_bounds__(256) k_ec_rescue(EcParams P)
{
	const unsigned lane = threadIdx.x & 31;
	const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	const uint64_t n_list = P.ctr[4];
	const int k = P.k;
	unsigned long long n_lookups = 0;
	for (uint64_t li = warp; li < n_list; li += n_warps) {
		const int64_t r = (int64_t)P.overflow[li];
		const uint64_t ob = P.off[r];
		const int64_t o = (int64_t)(ob - P.base0);
		const int n = (int)(P.off[r + 1] - ob - 1);
		int start = 0, end, ec = -1;
		while ((end = ec_first_kmer_end(P, o, n, start)) < n) {
			uint64_t x[4];
			extract_kmer(P, o, n, 0, end, k, -1, 0, x);
			// candidate t = i * 4 + j: base i of the k-mer (kmer.h bit order) changed to j, the sequential loop's order
			int best = 0, best_id = -1, second = 0;
			for (int t = (int)lane; t < 4 * k; t += 32) {
				const int i = t >> 2, j = t & 3, c = (int)((x[1] >> i & 1) << 1 | (x[0] >> i & 1));
				if (j == c) continue;
				uint64_t y[4] = { x[0], x[1], x[2], x[3] };
				bfc_kmer_change(k, y, i, j);
				const int ret = tab_kmer_occ(P.tab, y);
				++n_lookups;
				if (ret < 0) continue;
				const int cnt = ret & 0xff;
				if (cnt > best) second = best, best = cnt, best_id = t;
				else if (cnt > second) second = cnt;
			}
			for (int s = 16; s > 0; s >>= 1) { // merge the lanes' (largest, where first, second largest)
				const int ob_ = __shfl_xor_sync(0xffffffffu, best, s), oi = __shfl_xor_sync(0xffffffffu, best_id, s), os = __shfl_xor_sync(0xffffffffu, second, s);
				if (ob_ > best) second = best > os ? best : os, best = ob_, best_id = oi;
				else if (ob_ < best) second = second > ob_ ? second : ob_;
				else { // the same largest count on both sides: it is also the second largest; the earlier edit stays
					second = best;
					if (oi >= 0 && (best_id < 0 || oi < best_id)) best_id = oi;
				}
			}
			ec = best * 3 > P.mode && second < 3 ? best_id : -1; // correct.c:79 (best_id = i << 2 | j)
			if (ec >= 0) break;
			if (end + (k >> 1) >= n) break;
			start = end - (k >> 1);
		}
		if (lane == 0 && ec >= 0) {
			ReadDesc d;
			d.brute = (end - (ec >> 2)) << 2 | (ec & 3);
			++end;
			d.start0 = end - k, d.start1 = n - end, d.code = 0;
			P.desc[r] = d;
			P.jobs[2 * r] = make_int4((int)o, n, d.start0, d.brute);
			P.jobs[2 * r + 1] = make_int4((int)o, n, d.start1, d.brute);
		}
	}
	block_add(P.ctr + 1, n_lookups);
}

// ------------------------------------------------------------------ K6a': lookups past the end of the read

#define EC_EXT_STEPS 7 // positions n .. n + 6 fit the 8 bytes of a memo; used when max_end_ext < 7

// Past the end of the read (z.i >= n) bfc_ec1dir tries all four bases at every position (correct.c:318-333) and
// goes on only while exactly one of them is solid (correct.c:358-372), for at most max_end_ext + 1 positions
// (correct.c:289).  For a state whose last k-1 bases are the read's own, those lookups depend on nothing else in
// the search: they are made here for every job at once -- all lanes in step, four independent probes per position --
// and k_ec_search reads them back.  Byte s of the memo = position n + s: bits 0-3 "base b is solid"
// (cnt >= min_cov), bits 4-7 "its high count is below min_cov"; bit 63 = memo present.
__global__ void __launch_bounds__(256) k_ec_ext(EcParams P)
{
	const int64_t job = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long n_lookups = 0;
	if (job < P.n_jobs) {
		const int4 jr = P.jobs[job];
		const int k = P.k, n = jr.y, dir = (int)(job & 1);
		uint64_t memo = 0;
		if (jr.z >= 0 && n >= k) {
			uint64_t x[4];
			extract_kmer(P, (int64_t)(uint32_t)jr.x, n, dir, n - 1, k - 1, -1, 0, x);
			memo = 1ULL << 63;
			for (int s = 0; s <= P.max_end_ext; ++s) {
				uint32_t bits = 0;
				int n_solid = 0, last = 0;
#pragma unroll
				for (int b = 0; b < 4; ++b) {
					uint64_t y[4] = { x[0], x[1], x[2], x[3] };
					bfc_kmer_append(k, y, b);
					const int res = tab_kmer_occ(P.tab, y);
					if (res >= 0 && (res & 0xff) >= P.min_cov) {
						bits |= 1u << b;
						if ((res >> 8 & 0xff) < P.min_cov) bits |= 16u << b;
						++n_solid, last = b;
					}
				}
				n_lookups += 4;
				memo |= (uint64_t)bits << (8 * s);
				if (n_solid != 1) break;
				bfc_kmer_append(k, x, last);
			}
		}
		P.ext[job] = memo;
	}
	block_add(P.ctr + 1, n_lookups);
}

// ------------------------------------------------------------------ K6b: the search

enum { PC_NEWJOB = 0, PC_POP, PC_STEP, PC_OWN_DONE, PC_AFTER_OWN, PC_ALT_NEXT, PC_ALT_DONE, PC_FINISH, PC_EXIT };

#define EC_OVERFLOW (-100)

__device__ __forceinline__ uint32_t hk_pen(uint32_t key) { return key >> 12; }

// The key heap of a thread: keys 0..HK_SMEM-1 in shared memory (interleaved by thread, conflict-free),
// the rest in the thread's global scratch.  Heaps hold a handful of states almost always.
#define HK_SMEM 8
#define EC_THREADS 128
#ifndef EC_CTAS_PER_SM
#define EC_CTAS_PER_SM 6   // 80 registers per thread (measured best: 5 CTAs x 96 regs and 8 x 64 are slower)
#endif
struct KeyHeap {
	uint32_t *sm;   // &s_hk[0][threadIdx.x]
	uint32_t *gl;   // global scratch of heap_cap keys
	__device__ __forceinline__ uint32_t get(int j) const { return j < HK_SMEM ? sm[j * EC_THREADS] : gl[j]; }
	__device__ __forceinline__ void set(int j, uint32_t v) const { if (j < HK_SMEM) sm[j * EC_THREADS] = v; else gl[j] = v; }
};

// klib heap with "less" = larger tot_pen: root = smallest penalty (correct.c:179, ksort.h:125-146)
__device__ __forceinline__ void heapk_down(const KeyHeap &l, int n)
{
	int i = 0, c;
	const uint32_t tmp = l.get(0);
	while ((c = 2 * i + 1) < n) {
		uint32_t vc = l.get(c);
		if (c != n - 1) {
			const uint32_t vr = l.get(c + 1);
			if (hk_pen(vc) > hk_pen(vr)) ++c, vc = vr;
		}
		if (hk_pen(vc) > hk_pen(tmp)) break;
		l.set(i, vc); i = c;
	}
	l.set(i, tmp);
}

__device__ __forceinline__ void heapk_up(const KeyHeap &l, int n)
{
	int c = n - 1;
	const uint32_t tmp = l.get(c);
	while (c) {
		const int par = (c - 1) >> 1;
		const uint32_t vp = l.get(par);
		if (hk_pen(tmp) > hk_pen(vp)) break;
		l.set(c, vp); c = par;
	}
	l.set(c, tmp);
}

// packed candidate of one step: bit 0 valid, 1 ec, 2 ec_high, 3 absent, 4 absent_high
__device__ __forceinline__ int cand_weight(const EcParams &P, uint32_t c)
{
	return P.w_ec * (int)(c >> 1 & 1) + P.w_ec_high * (int)(c >> 2 & 1) + P.w_absent * (int)(c >> 3 & 1) + P.w_absent_high * (int)(c >> 4 & 1);
}

__global__ void __launch_bounds__(EC_THREADS, EC_CTAS_PER_SM) k_ec_search(EcParams P)
{
	const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	EcState *const pool = P.pool + slot * P.heap_cap;
	__shared__ uint32_t s_hk[HK_SMEM][EC_THREADS];
	KeyHeap heapk;
	heapk.sm = &s_hk[0][threadIdx.x], heapk.gl = P.heapk + slot * P.heap_cap;
	uint2 *const edits = P.edits + slot * P.edit_cap;
	const int k = P.k;

	// job context
	int64_t job = 0, o = 0;
	int jid = 0, n = 0, dir = 0, bp = -1, bb = 0;
	bool memo_dir = false;
	uint64_t ext = 0;
	// search context (reference bfc_ec1dir locals)
	EcState z;
	int heap_n = 0, n_init = 0, n_edits = 0, max_heap = 0, n_fail = 0, n_paths = 0, rvl = -1;
	int best_pen = INT_MAX, best_edit = -1, best_absent = 0;
	bool top_valid = false, have_best = false;
	int z_id = -1; // pool slot of a successor that is still only in registers (its key is in the heap)
	// step context
	int cb = -1, cob = -1, ff = 0, osf = 0, alt_mask = 0, other_ext = 0, cur_alt = 0;
	uint32_t cand = 0; // 4 x 8 bits, one per base
	bool has_c = false, fixed = false;
	// lookup hand-over
	int req_b = 0, res = -1;
	int pc = PC_NEWJOB;
	unsigned long long n_lookups = 0;
	z.tot_pen = z.i = z.n_absent = z.clean = 0, z.edit = -1;
	z.x[0] = z.x[1] = z.x[2] = z.x[3] = 0;
#pragma unroll
	for (int t = 0; t < BFC_EC_HIST; ++t) z.ecpos[t] = -1;
#pragma unroll
	for (int t = 0; t < BFC_EC_HIST_HIGH; ++t) z.ecpos_high[t] = -1;

	for (;;) {
		// ---------------- stage 0: take the lookup result in (correct.c:299 / :320)
		if (pc == PC_OWN_DONE) {
			osf = (res >= 0 && (res & 0xff) >= P.min_cov ? FL_SOL : 0) | ((res & 0xff) >= P.min_cov + 1 ? FL_A : 0) |
			      (res >= 0 && (res >> 8 & 0xff) >= P.min_cov ? FL_H : 0);
			pc = PC_AFTER_OWN;
		} else if (pc == PC_ALT_DONE) { // correct.c:320-333
			if (res >= 0 && (res & 0xff) >= P.min_cov) {
				const uint32_t ec = has_c && cb < 4 ? 1u : 0u;
				cand |= (1u | ec << 1 | (ec && (ff & FL_Q) ? 4u : 0u) | ((res >> 8 & 0xff) < P.min_cov ? 16u : 0u)) << (8 * cur_alt);
				++other_ext;
			}
			pc = PC_ALT_NEXT;
		}
		// ---------------- the step pipeline: ONE pass per round through the stages in the order a search step
		// runs through them, every lane executing the stages its state is due for, so the lanes of a warp stay
		// together whatever each of them is doing.  A lane that ends the pass without a lookup request (its
		// k-mer was memoised, a path ended, a new job started) just sits out this round's lookup.
		bool yield = false;
		{
			bool job_done = false;
			if (pc == PC_AFTER_OWN) {
				fixed = z.i > n; // correct.c:295 with end == n
				if (has_c && cb < 4) {
					if ((ff & FL_Q) && (osf & FL_A) && (ff & FL_LC)) fixed = true; // correct.c:299-301
					else if (ff & FL_HC) fixed = true;
					// the read base itself (correct.c:334-337)
					cand |= (1u | ((osf & FL_SOL) ? 0u : 8u) | ((osf & FL_H) ? 0u : 16u)) << (8 * cb);
				}
				alt_mask = 0;
				if (!(fixed && has_c)) {
					bool allowed = true;
					if (has_c) { // correct.c:316-317
						if ((ff & FL_Q) && z.ecpos_high[BFC_EC_HIST_HIGH - 1] >= 0 && z.i - z.ecpos_high[BFC_EC_HIST_HIGH - 1] < P.win_multi_ec) allowed = false;
						if (z.ecpos[BFC_EC_HIST - 1] >= 0 && z.i - z.ecpos[BFC_EC_HIST - 1] < P.win_multi_ec) allowed = false;
					}
					if (allowed) alt_mask = has_c && cb < 4 ? 0xF & ~(1 << cb) : 0xF;
				}
				pc = PC_ALT_NEXT;
			}
			if (pc == PC_ALT_NEXT) {
				if (alt_mask) {
					cur_alt = __ffs(alt_mask) - 1;
					alt_mask &= alt_mask - 1;
					req_b = cur_alt; pc = PC_ALT_DONE; yield = true;
				} else pc = PC_FINISH;
			}
			if (pc == PC_FINISH) {
				const int n_added = (int)((cand & 1) + (cand >> 8 & 1) + (cand >> 16 & 1) + (cand >> 24 & 1));
				pc = PC_POP;
				if (!fixed && other_ext == 0) ++n_fail;
				if (n_fail > n * 2) { rvl = -3; job_done = true; } // correct.c:342-347
				else if (has_c || n_added == 1) {
					uint32_t push = cand;
					if (n_added > 1 && heap_n > P.max_heap) { // keep only the cheapest extension, first on ties (correct.c:349-355)
						int min = INT_MAX, min_b = -1;
#pragma unroll
						for (int b = 0; b < 4; ++b)
							if (cand >> (8 * b) & 1) {
								const int t = cand_weight(P, cand >> (8 * b) & 0xff);
								if (min > t) min = t, min_b = b;
							}
						push = cand & (0xffu << (8 * min_b));
					}
					const uint32_t pm = push & 0x01010101u;
					if (pm != 0 && (pm & (pm - 1)) == 0) {
						// a single successor replaces z in registers (buf_update, correct.c:198-230); with other states alive
						// its KEY goes through the heap like any other, the state is stored only if another one pops first
						const int b = (__ffs(pm) - 1) >> 3;
						const uint32_t c = push >> (8 * b) & 0xff;
						const int zi = z.i;
						if (has_c && b != cb) { // the path changes this base: remember it (forward coordinates)
							if (n_edits >= P.edit_cap) { rvl = EC_OVERFLOW; job_done = true; }
							else {
								const int f = dir ? n - 1 - zi : zi;
								edits[n_edits] = make_uint2((uint32_t)z.edit, (uint32_t)f << 3 | (uint32_t)(dir ? 3 - b : b));
								z.edit = n_edits++;
							}
						}
						z.i = zi + 1;
						z.tot_pen += cand_weight(P, c);
						if (c & 4) z.ecpos_high[1] = z.ecpos_high[0], z.ecpos_high[0] = zi;
						if (c & 2) {
#pragma unroll
							for (int t = BFC_EC_HIST - 1; t > 0; --t) z.ecpos[t] = z.ecpos[t - 1];
							z.ecpos[0] = zi;
						}
						z.clean = b == cob ? z.clean + 1 : 0;
						if (has_c) z.n_absent += (int)(c >> 3 & 1);
						bfc_kmer_append(k, z.x, b);
						if (heap_n == 0) top_valid = true;
						else if (heap_n <= 3 && (uint32_t)z.tot_pen <= hk_pen(heapk.get(0))) {
							// Not above the cheapest other state: pushed, the successor rises to the root (ties keep rising
							// in ks_heapup) and the pop that follows hands it straight back.  The up-to-3 other keys end
							// where they were, except that two keys of equal penalty swap places (ksort.h:125-146 worked
							// through, as for the jumps below) -- so z stays in registers and the heap is not touched.
							if (heap_n == 2) {
								const uint32_t k0 = heapk.get(0), k1 = heapk.get(1);
								if (hk_pen(k0) == hk_pen(k1)) heapk.set(0, k1), heapk.set(1, k0);
							}
							top_valid = true;
						} else if (!job_done) {
							uint32_t id;
							if (heap_n < n_init) id = heapk.get(heap_n) & 0xfff;
							else id = (uint32_t)n_init++;
							heapk.set(heap_n++, (uint32_t)z.tot_pen << 12 | id);
							heapk_up(heapk, heap_n);
							z_id = (int)id;
						}
					} else {
						for (int b = 0; b < 4 && !job_done; ++b) { // buf_update (correct.c:198-230), in base order
							const uint32_t c = push >> (8 * b) & 0xff;
							if (!(c & 1)) continue;
							EcState s = z;
							s.i = z.i + 1;
							s.tot_pen = z.tot_pen + cand_weight(P, c);
							if (c & 4) s.ecpos_high[0] = z.i, s.ecpos_high[1] = z.ecpos_high[0];
							if (c & 2) {
								s.ecpos[0] = z.i;
#pragma unroll
								for (int t = 1; t < BFC_EC_HIST; ++t) s.ecpos[t] = z.ecpos[t - 1];
							}
							s.clean = b == cob ? z.clean + 1 : 0;
							if (has_c) {
								s.n_absent = z.n_absent + (int)(c >> 3 & 1);
								if (b != cb) {
									if (n_edits >= P.edit_cap) { rvl = EC_OVERFLOW; job_done = true; break; }
									const int f = dir ? n - 1 - z.i : z.i;
									edits[n_edits] = make_uint2((uint32_t)z.edit, (uint32_t)f << 3 | (uint32_t)(dir ? 3 - b : b));
									s.edit = n_edits++;
								}
							}
							bfc_kmer_append(k, s.x, b);
							uint32_t id;
							if (heap_n < n_init) id = heapk.get(heap_n) & 0xfff;
							else id = (uint32_t)n_init++;
							pool[id] = s;
							heapk.set(heap_n++, (uint32_t)s.tot_pen << 12 | id);
							heapk_up(heapk, heap_n);
						}
					}
				} else { // past the end of the read with 0 or >= 2 extensions: the path ends here (correct.c:360-372)
					const int fin = z.tot_pen + (n_added == 0 ? P.w_absent * (P.max_end_ext - (z.i - n)) : 0);
					if (fin < best_pen) best_pen = fin, best_edit = z.edit, best_absent = z.n_absent, have_best = true;
					if (++n_paths == BFC_MAX_PATHS) job_done = true;
				}
			}
			if (pc == PC_POP && !job_done) {
				const int hs = heap_n + (top_valid ? 1 : 0);
				max_heap = max_heap > 255 ? 255 : max_heap > hs ? max_heap : hs; // correct.c:276
				if (hs == 0) { rvl = -2; job_done = true; }
				else {
					if (top_valid) top_valid = false;
					else { // ks_heapdown-based pop (correct.c:281-283); the freed slot id parks behind the live keys
						const uint32_t key = heapk.get(0);
						--heap_n;
						heapk.set(0, heapk.get(heap_n));
						heapk.set(heap_n, key);
						if (heap_n > 1) heapk_down(heapk, heap_n);
						if ((int)(key & 0xfff) != z_id) {
							if (z_id >= 0) pool[z_id] = z; // the register state lost the pop: now it needs its slot
							z = pool[key & 0xfff];
						}
						z_id = -1;
					}
					if (have_best && z.tot_pen > best_pen + P.max_path_diff) job_done = true; // correct.c:288
					else if (z.i - n > P.max_end_ext) { // correct.c:289, 366-372; then the next pop
						if (z.tot_pen < best_pen) best_pen = z.tot_pen, best_edit = z.edit, best_absent = z.n_absent, have_best = true;
						if (++n_paths == BFC_MAX_PATHS) job_done = true;
					} else pc = PC_STEP;
				}
			}
			if (job_done) {
				if (rvl == EC_OVERFLOW) P.overflow[atomicAdd(P.ctr, 1ULL)] = (uint32_t)jid;
				else {
					int rv = rvl;
					if (n_paths > 0) { // buf_backtrack (correct.c:232-247): only the changed bases need recording
						uint32_t *ev = (uint32_t*)(P.pl + (uint64_t)(dir ? PL_E1V : PL_E0V) * P.pl_words);
						const uint64_t s32 = P.pl_words * 2;
						for (int e = best_edit; e >= 0;) {
							const uint2 ed = edits[e];
							const uint64_t pos = (uint64_t)(o + (int64_t)(ed.y >> 3) + PL_PAD);
							const uint32_t bit = 1u << (pos & 31), b = ed.y & 7;
							uint32_t *w = ev + (pos >> 5);
							atomicOr(w, bit);
							if (b & 1) atomicOr(w + s32, bit);
							if (b & 2) atomicOr(w + 2 * s32, bit);
							e = (int)ed.x;
						}
						rv = best_absent;
					}
					P.res[jid] = make_int2(rv, max_heap);
				}
				pc = PC_NEWJOB;
			}
			if (pc == PC_NEWJOB) {
				// jobs are handed out dynamically (their lengths vary by orders of magnitude): the lanes that arrive
				// here together take consecutive jobs with one atomic
				const unsigned am = __activemask(), ln = threadIdx.x & 31;
				const int leader = __ffs(am) - 1;
				unsigned long long first = 0;
				if ((int)ln == leader) first = atomicAdd(P.ctr + 2, (unsigned long long)__popc(am));
				first = __shfl_sync(am, first, leader);
				job = (int64_t)(first + __popc(am & ((1u << ln) - 1)));
				if (job >= P.n_jobs) { pc = PC_EXIT; yield = true; }
				else {
					jid = P.redo ? (int)P.redo[job] : (int)job;
					const int4 jr = P.jobs[jid];
					if (jr.z >= 0) { // else: nothing to search for this read
						dir = jid & 1;
						o = (int64_t)(uint32_t)jr.x;
						n = jr.y;
						bp = jr.w >= 0 ? jr.w >> 2 : -1, bb = jr.w & 3;
						// the reverse-complement k-mer hashes like the forward one only for odd k (kmer.h:81)
						memo_dir = dir == 0 || (k & 1) != 0;
						ext = P.ext ? __ldg(P.ext + jid) : 0;
						const int start = jr.z;
						heap_n = n_init = n_edits = 0, max_heap = 0, n_fail = 0, n_paths = 0, rvl = -1, z_id = -1;
						best_pen = INT_MAX, best_edit = -1, best_absent = 0, have_best = false;
						// seed: the k-1 bases before position z.i (correct.c:260-267); [start, start+k) is a k-mer of ACGT
						z.i = start + k - 1;
						z.tot_pen = 0, z.edit = -1, z.n_absent = 0;
#pragma unroll
						for (int t = 0; t < BFC_EC_HIST; ++t) z.ecpos[t] = -1;
#pragma unroll
						for (int t = 0; t < BFC_EC_HIST_HIGH; ++t) z.ecpos_high[t] = -1;
						if (start < 0 || z.i >= n) { P.res[jid] = make_int2(-1, 0); } // the reference asserts
						else {
							extract_kmer(P, o, n, dir, z.i - 1, k - 1, bp, bb, z.x);
							z.clean = k - 1;
							if (bp >= 0) { // bases after the rescue edit are the original ones
								const int ib = dir ? n - 1 - bp : bp; // search position of the edit
								if (ib >= start && ib <= z.i - 1) z.clean = z.i - 1 - ib;
							}
							top_valid = true;
							pc = PC_POP;
						}
					}
				}
			}
			while (pc == PC_STEP) { // repeats only after a jump
				has_c = z.i < n;
				cand = 0, other_ext = 0, osf = 0, cb = cob = -1, ff = 0;
				pc = PC_AFTER_OWN;
				if (has_c) {
					const int f = dir ? n - 1 - z.i : z.i;
					ff = __ldg(P.fl + o + f);
					const int ob = FL_OB(ff), cur = f == bp ? bb : ob;
					cb = dir ? comp_b(cur) : cur, cob = dir ? comp_b(ob) : ob;
					if (cb < 4) {
						if (memo_dir && z.clean >= k - 1 && cb == cob) { // the read's own k-mer: K5 fetched it
							// Step over every decision-free base at once.  Each such step pushes one successor with the
							// same penalty, which is popped right back (ties keep rising in ks_heapup), so only z moves.
							// With up to 3 OTHER states alive the push + pop leaves their heap order untouched, except
							// that two states of equal penalty swap places each time (ksort.h:125-146 worked through).
							int run = 0;
							if (heap_n <= 3) {
								if (!dir) { const uint64_t w = ~bits64(plane(P, PL_J0), o + f); run = w ? __ffsll((long long)w) - 1 : 64; }
								else { const uint64_t w = ~bits64(plane(P, PL_J1), o + f - 63); run = w ? __clzll((long long)w) : 64; }
								if (bp >= 0) { // the rescued base is not the original one: stop in front of it
									const int ib = dir ? n - 1 - bp : bp;
									if (ib >= z.i && ib < z.i + run) run = ib - z.i;
								}
							}
							if (run > 0) {
								const int hs = heap_n + 1;
								max_heap = max_heap > 255 ? 255 : max_heap > hs ? max_heap : hs;
								if (heap_n == 2 && (run & 1)) {
									const uint32_t k0 = heapk.get(0), k1 = heapk.get(1);
									if (hk_pen(k0) == hk_pen(k1)) heapk.set(0, k1), heapk.set(1, k0);
								}
								z.i += run, z.clean += run;
								extract_kmer(P, o, n, dir, z.i - 1, k - 1, -1, 0, z.x);
								pc = PC_STEP; // again, at the new position
							} else osf = __ldg(P.fl + o + (dir ? f + k - 1 : f));
						} else { req_b = cb; pc = PC_OWN_DONE; yield = true; }
					}
				} else if ((ext >> 63) && z.clean >= k - 1) {
					// past the end with the read's own last k-1 bases (and k_ec_ext's bases after them): the four
					// lookups of this position (correct.c:318-333) were made by k_ec_ext
					const uint32_t m8 = (uint32_t)(ext >> (8 * (z.i - n))) & 0xff;
#pragma unroll
					for (int b = 0; b < 4; ++b)
						if (m8 >> b & 1) cand |= (1u | ((m8 >> (4 + b) & 1) ? 16u : 0u)) << (8 * b), ++other_ext;
					const uint32_t sol = m8 & 15;
					cob = sol != 0 && (sol & (sol - 1)) == 0 ? __ffs(sol) - 1 : -1; // the base k_ec_ext went on with
					fixed = z.i > n; // correct.c:295
					pc = PC_FINISH;
				}
			}
		}
		if (pc == PC_EXIT) break;
		// ---------------- the one lookup site of the loop
		if (yield) {
			uint64_t x[4] = { z.x[0], z.x[1], z.x[2], z.x[3] };
			bfc_kmer_append(k, x, req_b);
			res = tab_kmer_occ(P.tab, x);
			++n_lookups;
		}
	}
	block_add(P.ctr + 1, n_lookups);
	block_add(P.ctr + 3, n_lookups);
}

// ------------------------------------------------------------------ K6b': the search, several lookups per round
//
// Same search, same results as k_ec_search; what changes is how often a thread has to come round to the lookup site.
// k_ec_search makes ONE table lookup per pass of its stage loop, so a position that tries the three other bases costs
// four passes, and the k-1 positions after an edit (whose k-mers contain the new base, so K5 never fetched them) cost
// one pass each -- and a pass is expensive, because the 32 lanes of a warp are at 32 different places of the search.
// Here a pass ends with up to EC_RQ lookups per thread:
//   * the alternatives of a position (correct.c:318-333) are looked up together, and together with the position's own
//     k-mer whenever the per-base flags alone already rule out "fixed" (correct.c:299-301);
//   * while the path follows the read after an edit, the own k-mers of the NEXT positions are looked up ahead (they
//     only depend on the read's bases) and kept in a 4-entry cache that stays valid for as long as the state keeps
//     taking the read's base; any other base, or another state popped from the heap, drops it.
// Every result is reduced at once to the three bits the search ever asks of a table value (occ_code).
// The push / pop / heap logic below the step is k_ec_search's, statement for statement.

#define EC_RQ 4
#ifndef EC2_CTAS_PER_SM
#define EC2_CTAS_PER_SM 5
#endif

// bit 0: in the table with cnt >= min_cov (solid); bit 1: (os & 0xff) >= min_cov + 1 with os = -1 counting as 255
// (correct.c:299-300); bit 2: in the table with high count >= min_cov
__device__ __forceinline__ uint32_t occ_code(int res, int min_cov)
{
	return (res >= 0 && (res & 0xff) >= min_cov ? 1u : 0u) | ((res & 0xff) >= min_cov + 1 ? 2u : 0u) | (res >= 0 && (res >> 8 & 0xff) >= min_cov ? 4u : 0u);
}
__device__ __forceinline__ int code_flags(uint32_t c) { return (c & 1 ? FL_SOL : 0) | (c & 2 ? FL_A : 0) | (c & 4 ? FL_H : 0); }

enum { P2_NEWJOB = 0, P2_POP, P2_STEP, P2_RES, P2_FINISH, P2_EXIT };

template <int CTAS>
__global__ void __launch_bounds__(EC_THREADS, CTAS) k_ec_search2(EcParams P)
{
	const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	EcState *const pool = P.pool + slot * P.heap_cap;
	__shared__ uint32_t s_hk[HK_SMEM][EC_THREADS];
	KeyHeap heapk;
	heapk.sm = &s_hk[0][threadIdx.x], heapk.gl = P.heapk + slot * P.heap_cap;
	uint2 *const edits = P.edits + slot * P.edit_cap;
	const int k = P.k;

	// job context
	int64_t job = 0, o = 0;
	int jid = 0, n = 0, dir = 0, bp = -1, bb = 0;
	bool memo_dir = false;
	uint64_t ext = 0;
	// search context (reference bfc_ec1dir locals)
	EcState z;
	int heap_n = 0, n_init = 0, n_edits = 0, max_heap = 0, n_fail = 0, n_paths = 0, rvl = -1;
	int best_pen = INT_MAX, best_edit = -1, best_absent = 0;
	bool top_valid = false, have_best = false;
	int z_id = -1;
	// step context
	int cb = -1, cob = -1, ff = 0, osf = 0, other_ext = 0;
	uint32_t cand = 0, alt_mask = 0;
	bool has_c = false, fixed = false, own_pend = false, allowed = true;
	// this round's requests and their results
	uint32_t rq_cur = 0;                 // bases to try at the current position (appended to z.x)
	uint32_t rq_done = 0;                // ... that have been looked up for this position so far
	uint32_t rbits = 0;                  // occ_code of base b at the current position: bits 3b .. 3b+2
	uint32_t ch_len = 0, ch_skip = 0, ch_bases = 0; // chain along the read from z.i: 2-bit bases; the first ch_skip are cached
	// look-ahead cache: own k-mers of positions la_pos .. la_pos + 3 of the path that keeps taking the read's bases
	int la_pos = INT_MIN;
	uint32_t la_valid = 0, la_code = 0;  // entry j: bit j of la_valid, bits 3j .. 3j+2 of la_code
	int pc = P2_NEWJOB;
	unsigned long long n_lookups = 0;
	z.tot_pen = z.i = z.n_absent = z.clean = 0, z.edit = -1;
	z.x[0] = z.x[1] = z.x[2] = z.x[3] = 0;
#pragma unroll
	for (int t = 0; t < BFC_EC_HIST; ++t) z.ecpos[t] = -1;
#pragma unroll
	for (int t = 0; t < BFC_EC_HIST_HIGH; ++t) z.ecpos_high[t] = -1;

	for (;;) {
		bool yield = false;
		{
			bool job_done = false;
			// ---------------- the lookups of the last round are in
			if (pc == P2_RES) {
				if (ch_len > ch_skip) la_valid = (1u << ch_len) - 1; // the chain's results went straight into the cache
				pc = P2_FINISH;
				if (own_pend) { // correct.c:299-301, 334-337 with the own k-mer's value
					own_pend = false;
					osf = code_flags(la_code & 7);
					if (((ff & FL_Q) && (osf & FL_A) && (ff & FL_LC)) || (ff & FL_HC)) fixed = true;
					cand |= (1u | ((osf & FL_SOL) ? 0u : 8u) | ((osf & FL_H) ? 0u : 16u)) << (8 * cb);
					alt_mask = !fixed && allowed ? 0xFu & ~(1u << cb) : 0u;
					const uint32_t need = alt_mask & ~rq_done;
					if (need) { rq_cur = need, ch_len = ch_skip = 0; pc = P2_RES; yield = true; } // the flags left "fixed" open and the k-mer closed it: second round
				}
				if (pc == P2_FINISH) { // correct.c:320-333
					const uint32_t ec = has_c && cb < 4 ? 1u : 0u;
					for (uint32_t m = alt_mask; m; m &= m - 1) {
						const int b = __ffs(m) - 1;
						const uint32_t c = rbits >> (3 * b) & 7;
						if (c & 1) {
							cand |= (1u | ec << 1 | (ec && (ff & FL_Q) ? 4u : 0u) | ((c & 4) ? 0u : 16u)) << (8 * b);
							++other_ext;
						}
					}
				}
			}
			if (pc == P2_FINISH) {
				const int n_added = (int)((cand & 1) + (cand >> 8 & 1) + (cand >> 16 & 1) + (cand >> 24 & 1));
				pc = P2_POP;
				if (!fixed && other_ext == 0) ++n_fail;
				if (n_fail > n * 2) { rvl = -3; job_done = true; } // correct.c:342-347
				else if (has_c || n_added == 1) {
					uint32_t push = cand;
					if (n_added > 1 && heap_n > P.max_heap) { // keep only the cheapest extension, first on ties (correct.c:349-355)
						int min = INT_MAX, min_b = -1;
#pragma unroll
						for (int b = 0; b < 4; ++b)
							if (cand >> (8 * b) & 1) {
								const int t = cand_weight(P, cand >> (8 * b) & 0xff);
								if (min > t) min = t, min_b = b;
							}
						push = cand & (0xffu << (8 * min_b));
					}
					const uint32_t pm = push & 0x01010101u;
					if (pm != 0 && (pm & (pm - 1)) == 0) {
						// a single successor replaces z in registers (see k_ec_search)
						const int b = (__ffs(pm) - 1) >> 3;
						const uint32_t c = push >> (8 * b) & 0xff;
						const int zi = z.i;
						if (has_c && b != cb) {
							if (n_edits >= P.edit_cap) { rvl = EC_OVERFLOW; job_done = true; }
							else {
								const int f = dir ? n - 1 - zi : zi;
								edits[n_edits] = make_uint2((uint32_t)z.edit, (uint32_t)f << 3 | (uint32_t)(dir ? 3 - b : b));
								z.edit = n_edits++;
							}
						}
						if (b != cb) la_valid = 0; // off the read: the k-mers looked up ahead are not this path's
						z.i = zi + 1;
						z.tot_pen += cand_weight(P, c);
						if (c & 4) z.ecpos_high[1] = z.ecpos_high[0], z.ecpos_high[0] = zi;
						if (c & 2) {
#pragma unroll
							for (int t = BFC_EC_HIST - 1; t > 0; --t) z.ecpos[t] = z.ecpos[t - 1];
							z.ecpos[0] = zi;
						}
						z.clean = b == cob ? z.clean + 1 : 0;
						if (has_c) z.n_absent += (int)(c >> 3 & 1);
						bfc_kmer_append(k, z.x, b);
						if (heap_n == 0) top_valid = true;
						else if (heap_n <= 3 && (uint32_t)z.tot_pen <= hk_pen(heapk.get(0))) {
							if (heap_n == 2) {
								const uint32_t k0 = heapk.get(0), k1 = heapk.get(1);
								if (hk_pen(k0) == hk_pen(k1)) heapk.set(0, k1), heapk.set(1, k0);
							}
							top_valid = true;
						} else if (!job_done) {
							uint32_t id;
							if (heap_n < n_init) id = heapk.get(heap_n) & 0xfff;
							else id = (uint32_t)n_init++;
							heapk.set(heap_n++, (uint32_t)z.tot_pen << 12 | id);
							heapk_up(heapk, heap_n);
							z_id = (int)id;
						}
					} else {
						for (int b = 0; b < 4 && !job_done; ++b) { // buf_update (correct.c:198-230), in base order
							const uint32_t c = push >> (8 * b) & 0xff;
							if (!(c & 1)) continue;
							EcState s = z;
							s.i = z.i + 1;
							s.tot_pen = z.tot_pen + cand_weight(P, c);
							if (c & 4) s.ecpos_high[0] = z.i, s.ecpos_high[1] = z.ecpos_high[0];
							if (c & 2) {
								s.ecpos[0] = z.i;
#pragma unroll
								for (int t = 1; t < BFC_EC_HIST; ++t) s.ecpos[t] = z.ecpos[t - 1];
							}
							s.clean = b == cob ? z.clean + 1 : 0;
							if (has_c) {
								s.n_absent = z.n_absent + (int)(c >> 3 & 1);
								if (b != cb) {
									if (n_edits >= P.edit_cap) { rvl = EC_OVERFLOW; job_done = true; break; }
									const int f = dir ? n - 1 - z.i : z.i;
									edits[n_edits] = make_uint2((uint32_t)z.edit, (uint32_t)f << 3 | (uint32_t)(dir ? 3 - b : b));
									s.edit = n_edits++;
								}
							}
							bfc_kmer_append(k, s.x, b);
							uint32_t id;
							if (heap_n < n_init) id = heapk.get(heap_n) & 0xfff;
							else id = (uint32_t)n_init++;
							pool[id] = s;
							heapk.set(heap_n++, (uint32_t)s.tot_pen << 12 | id);
							heapk_up(heapk, heap_n);
						}
					}
				} else { // past the end of the read with 0 or >= 2 extensions: the path ends here (correct.c:360-372)
					const int fin = z.tot_pen + (n_added == 0 ? P.w_absent * (P.max_end_ext - (z.i - n)) : 0);
					if (fin < best_pen) best_pen = fin, best_edit = z.edit, best_absent = z.n_absent, have_best = true;
					if (++n_paths == BFC_MAX_PATHS) job_done = true;
				}
			}
			if (pc == P2_POP && !job_done) {
				const int hs = heap_n + (top_valid ? 1 : 0);
				max_heap = max_heap > 255 ? 255 : max_heap > hs ? max_heap : hs; // correct.c:276
				if (hs == 0) { rvl = -2; job_done = true; }
				else {
					if (top_valid) top_valid = false;
					else {
						const uint32_t key = heapk.get(0);
						--heap_n;
						heapk.set(0, heapk.get(heap_n));
						heapk.set(heap_n, key);
						if (heap_n > 1) heapk_down(heapk, heap_n);
						if ((int)(key & 0xfff) != z_id) {
							if (z_id >= 0) pool[z_id] = z;
							z = pool[key & 0xfff];
							la_valid = 0; // another path
						}
						z_id = -1;
					}
					if (have_best && z.tot_pen > best_pen + P.max_path_diff) job_done = true; // correct.c:288
					else if (z.i - n > P.max_end_ext) { // correct.c:289, 366-372; then the next pop
						if (z.tot_pen < best_pen) best_pen = z.tot_pen, best_edit = z.edit, best_absent = z.n_absent, have_best = true;
						if (++n_paths == BFC_MAX_PATHS) job_done = true;
					} else pc = P2_STEP;
				}
			}
			if (job_done) {
				if (rvl == EC_OVERFLOW) P.overflow[atomicAdd(P.ctr, 1ULL)] = (uint32_t)jid;
				else {
					int rv = rvl;
					if (n_paths > 0) { // buf_backtrack (correct.c:232-247): only the changed bases need recording
						uint32_t *ev = (uint32_t*)(P.pl + (uint64_t)(dir ? PL_E1V : PL_E0V) * P.pl_words);
						const uint64_t s32 = P.pl_words * 2;
						for (int e = best_edit; e >= 0;) {
							const uint2 ed = edits[e];
							const uint64_t pos = (uint64_t)(o + (int64_t)(ed.y >> 3) + PL_PAD);
							const uint32_t bit = 1u << (pos & 31), b = ed.y & 7;
							uint32_t *w = ev + (pos >> 5);
							atomicOr(w, bit);
							if (b & 1) atomicOr(w + s32, bit);
							if (b & 2) atomicOr(w + 2 * s32, bit);
							e = (int)ed.x;
						}
						rv = best_absent;
					}
					P.res[jid] = make_int2(rv, max_heap);
				}
				pc = P2_NEWJOB;
			}
			if (pc == P2_NEWJOB) {
				const unsigned am = __activemask(), ln = threadIdx.x & 31;
				const int leader = __ffs(am) - 1;
				unsigned long long first = 0;
				if ((int)ln == leader) first = atomicAdd(P.ctr + 2, (unsigned long long)__popc(am));
				first = __shfl_sync(am, first, leader);
				job = (int64_t)(first + __popc(am & ((1u << ln) - 1)));
				if (job >= P.n_jobs) pc = P2_EXIT;
				else {
					jid = P.redo ? (int)P.redo[job] : (int)job;
					const int4 jr = P.jobs[jid];
					if (jr.z >= 0) { // else: nothing to search for this read
						dir = jid & 1;
						o = (int64_t)(uint32_t)jr.x;
						n = jr.y;
						bp = jr.w >= 0 ? jr.w >> 2 : -1, bb = jr.w & 3;
						memo_dir = dir == 0 || (k & 1) != 0; // the reverse-complement k-mer hashes like the forward one only for odd k (kmer.h:81)
						ext = P.ext ? __ldg(P.ext + jid) : 0;
						const int start = jr.z;
						heap_n = n_init = n_edits = 0, max_heap = 0, n_fail = 0, n_paths = 0, rvl = -1, z_id = -1;
						best_pen = INT_MAX, best_edit = -1, best_absent = 0, have_best = false;
						la_valid = 0;
						z.i = start + k - 1; // seed: the k-1 bases before position z.i (correct.c:260-267)
						z.tot_pen = 0, z.edit = -1, z.n_absent = 0;
#pragma unroll
						for (int t = 0; t < BFC_EC_HIST; ++t) z.ecpos[t] = -1;
#pragma unroll
						for (int t = 0; t < BFC_EC_HIST_HIGH; ++t) z.ecpos_high[t] = -1;
						if (start < 0 || z.i >= n) { P.res[jid] = make_int2(-1, 0); } // the reference asserts
						else {
							extract_kmer(P, o, n, dir, z.i - 1, k - 1, bp, bb, z.x);
							z.clean = k - 1;
							if (bp >= 0) { // bases after the rescue edit are the original ones
								const int ib = dir ? n - 1 - bp : bp;
								if (ib >= start && ib <= z.i - 1) z.clean = z.i - 1 - ib;
							}
							top_valid = true;
							pc = P2_POP;
						}
					}
				}
			}
			if (pc == P2_STEP) {
				bool memo = false, own_known = false;
				int f = 0;
				for (;;) { // repeats only after stepping over decision-free bases
					// Step over every decision-free base at once (see k_ec_search): possible while the path's last k-1
					// bases are the read's own, which is what the jump planes were computed for
					int run = 0;
					if (z.i < n && memo_dir && z.clean >= k - 1 && heap_n <= 3) {
						const int fz = dir ? n - 1 - z.i : z.i;
						if (!dir) { const uint64_t w = ~bits64(plane(P, PL_J0), o + fz); run = w ? __ffsll((long long)w) - 1 : 64; }
						else { const uint64_t w = ~bits64(plane(P, PL_J1), o + fz - 63); run = w ? __clzll((long long)w) : 64; }
						if (bp >= 0) { // the rescued base is not the original one: stop in front of it
							const int ib = dir ? n - 1 - bp : bp;
							if (ib >= z.i && ib < z.i + run) run = ib - z.i;
						}
						if (run > 0) {
							const int hs = heap_n + 1;
							max_heap = max_heap > 255 ? 255 : max_heap > hs ? max_heap : hs;
							if (heap_n == 2 && (run & 1)) {
								const uint32_t k0 = heapk.get(0), k1 = heapk.get(1);
								if (hk_pen(k0) == hk_pen(k1)) heapk.set(0, k1), heapk.set(1, k0);
							}
							z.i += run, z.clean += run;
						}
					}
					has_c = z.i < n;
					cb = cob = -1, ff = 0, osf = 0, allowed = true, memo = own_known = false;
					int ff2 = 0;
					if (has_c) { // the flags of the position travel while the k-mer is cut out of the planes
						f = dir ? n - 1 - z.i : z.i;
						ff = __ldg(P.fl + o + f);
						ff2 = dir && memo_dir && z.clean >= k - 1 ? (int)__ldg(P.fl + o + f + k - 1) : ff; // the k-mer ending here (in search order)
					}
					if (run > 0) extract_kmer(P, o, n, dir, z.i - 1, k - 1, -1, 0, z.x);
					if (!has_c) break;
					const int ob = FL_OB(ff), cur = f == bp ? bb : ob;
					cb = dir ? comp_b(cur) : cur, cob = dir ? comp_b(ob) : ob;
					if (cb > 3) break;
					memo = memo_dir && z.clean >= k - 1 && cb == cob; // the read's own k-mer: K5 fetched it
					if (memo) { osf = ff2 & (FL_SOL | FL_A | FL_H); own_known = true; break; }
					// looked up ahead?  Entry 0 of the cache becomes this position.
					const int d = z.i - la_pos;
					if (la_valid && d > 0 && d < 4) la_valid >>= d, la_code >>= 3 * d;
					else if (d != 0) la_valid = 0;
					if (!la_valid) la_code = 0;
					la_pos = z.i;
					if (!(la_valid & 1)) break;
					osf = code_flags(la_code & 7), own_known = true;
					// A lone successor at no cost -- the base is "fixed" (correct.c:299-301) and its k-mer solid with a solid
					// high count (correct.c:334-337) -- is pushed and popped right back, exactly as in a jump: only z moves
					if (heap_n > 3 || !(osf & FL_SOL) || !(osf & FL_H) || !(((ff & FL_Q) && (osf & FL_A) && (ff & FL_LC)) || (ff & FL_HC))) break;
					{
						const int hs = heap_n + 1;
						max_heap = max_heap > 255 ? 255 : max_heap > hs ? max_heap : hs;
						if (heap_n == 2) {
							const uint32_t k0 = heapk.get(0), k1 = heapk.get(1);
							if (hk_pen(k0) == hk_pen(k1)) heapk.set(0, k1), heapk.set(1, k0);
						}
						z.clean = cb == cob ? z.clean + 1 : 0;
						bfc_kmer_append(k, z.x, cb);
						++z.i;
					}
				}
				cand = 0, other_ext = 0, alt_mask = 0;
				rq_cur = rq_done = rbits = 0, ch_len = ch_skip = 0, own_pend = false;
				fixed = z.i > n; // correct.c:295 with end == n
				pc = P2_FINISH;
				if (has_c) {
					// correct.c:316-317
					if ((ff & FL_Q) && z.ecpos_high[BFC_EC_HIST_HIGH - 1] >= 0 && z.i - z.ecpos_high[BFC_EC_HIST_HIGH - 1] < P.win_multi_ec) allowed = false;
					if (z.ecpos[BFC_EC_HIST - 1] >= 0 && z.i - z.ecpos[BFC_EC_HIST - 1] < P.win_multi_ec) allowed = false;
					if (cb < 4) {
						if (own_known) {
							if (((ff & FL_Q) && (osf & FL_A) && (ff & FL_LC)) || (ff & FL_HC)) fixed = true; // correct.c:299-301
							cand |= (1u | ((osf & FL_SOL) ? 0u : 8u) | ((osf & FL_H) ? 0u : 16u)) << (8 * cb); // correct.c:334-337
							alt_mask = !fixed && allowed ? 0xFu & ~(1u << cb) : 0u;
							rq_cur = alt_mask;
						} else {
							own_pend = true;
							// the flags alone already rule out "fixed": the alternatives go along with the own k-mer
							if (allowed && !(ff & FL_HC) && !((ff & FL_Q) && (ff & FL_LC))) rq_cur = 0xFu & ~(1u << cb);
						}
						if (!memo) {
							// the chain: own k-mers of z.i, z.i + 1, ... for as long as the path would keep the read's bases,
							// the next position is not K5's again, and this round has lookups to spare
							const int c1 = cb == cob ? z.clean + 1 : 0;
							int len = n - z.i < 4 ? n - z.i : 4;
							if (memo_dir && k - c1 < len) len = k - c1 < 1 ? 1 : k - c1;
							ch_skip = (uint32_t)__ffs(~la_valid) - 1; // cached entries come first
							const int room = EC_RQ - __popc(rq_cur);
							if ((int)ch_skip + room < len) len = (int)ch_skip + room;
							ch_bases = (uint32_t)cb;
							int obs[3];
#pragma unroll
							for (int m = 1; m < 4; ++m) { // (three independent loads)
								const int fj = dir ? f - m : f + m;
								obs[m - 1] = m < len && fj != bp ? FL_OB(__ldg(P.fl + o + fj)) : 7;
							}
							int m = 1;
#pragma unroll
							for (int t = 0; t < 3; ++t)
								if (m == t + 1 && obs[t] <= 3) ch_bases |= (uint32_t)(dir ? 3 - obs[t] : obs[t]) << (2 * m), ++m;
							ch_len = (uint32_t)m;
							if (ch_len <= ch_skip) ch_len = ch_skip = 0; // nothing new to look up
						}
					} else { // a non-ACGT read base has no own candidate
						alt_mask = allowed ? 0xFu : 0u;
						rq_cur = alt_mask;
					}
				} else if ((ext >> 63) && z.clean >= k - 1) {
					// past the end with the read's own last k-1 bases: k_ec_ext made the four lookups of this position
					const uint32_t m8 = (uint32_t)(ext >> (8 * (z.i - n))) & 0xff;
#pragma unroll
					for (int b = 0; b < 4; ++b)
						if (m8 >> b & 1) cand |= (1u | ((m8 >> (4 + b) & 1) ? 16u : 0u)) << (8 * b), ++other_ext;
					const uint32_t sol = m8 & 15;
					cob = sol != 0 && (sol & (sol - 1)) == 0 ? __ffs(sol) - 1 : -1; // the base k_ec_ext went on with
				} else { // past the end: all four bases (correct.c:318-333)
					alt_mask = 0xFu;
					rq_cur = alt_mask;
				}
				if (rq_cur | (ch_len > ch_skip ? 1u : 0u)) { pc = P2_RES; yield = true; }
			}
		}
		if (pc == P2_EXIT) break;
		// ---------------- the lookup site: this round's requests of every lane, one after the other
		if (yield) {
			uint64_t xr[4] = { z.x[0], z.x[1], z.x[2], z.x[3] };
			uint32_t m = rq_cur, cj = 0;
			rq_done |= rq_cur;
			for (; cj < ch_skip; ++cj) bfc_kmer_append(k, xr, (int)(ch_bases >> (2 * cj) & 3));
			while (m | (cj < ch_len ? 1u : 0u)) {
				uint64_t x[4];
				uint32_t sh;
				if (m) {
					const int b = __ffs(m) - 1;
					m &= m - 1;
					x[0] = z.x[0], x[1] = z.x[1], x[2] = z.x[2], x[3] = z.x[3];
					bfc_kmer_append(k, x, b);
					sh = 3 * b;
				} else {
					bfc_kmer_append(k, xr, (int)(ch_bases >> (2 * cj) & 3));
					x[0] = xr[0], x[1] = xr[1], x[2] = xr[2], x[3] = xr[3];
					sh = 16 + 3 * cj;
					++cj;
				}
				const uint32_t c = occ_code(tab_kmer_occ(P.tab, x), P.min_cov);
				++n_lookups;
				if (sh < 16) rbits |= c << sh;
				else la_code |= c << (sh - 16);
			}
			rq_cur = 0;
		}
	}
	block_add(P.ctr + 1, n_lookups);
	block_add(P.ctr + 3, n_lookups);
}

// ------------------------------------------------------------------ K6c: merge + rewrite, one read per warp

__global__ void __launch_bounds__(256) k_ec_merge(EcParams P)
{
	const int lane = threadIdx.x & 31;
	const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
	for (int64_t r = warp; r < P.n_reads; r += n_warps) {
		const uint64_t ob0 = P.off[r];
		const int64_t o = (int64_t)(ob0 - P.base0);
		const int n = (int)(P.off[r + 1] - ob0 - 1), k = P.k;
		const ReadDesc d = P.desc[r];
		uint32_t ec_code = (uint32_t)d.code, brute = d.brute >= 0, n_absent = 0, mh = 0;
		int2 r0 = make_int2(0, 0), r1 = r0;
		if (ec_code == 0) {
			r0 = P.res[2 * r], r1 = P.res[2 * r + 1];
			const int bad = r0.x < 0 ? r0.x : r1.x < 0 ? r1.x : 0; // the reference stops at the first failing direction
			if (bad < 0) ec_code = bad == -2 ? 4 : bad == -3 ? 5 : 1;
		}
		uint32_t n_ec = 0, n_ec_high = 0, rf_code = P.refine ? 1u : 0u;
		if (P.refine && ec_code == 0) { // correct.c:438-442: more absent k-mers than the earlier round left => keep that one
			const uint32_t oa = P.aux[2 * r], oa2 = P.aux[2 * r + 1];
			if ((oa & 7) == 0 && (uint32_t)(r0.x + r1.x) > oa2 >> 10) {
				__syncwarp();
				if (lane == 0) P.aux[2 * r + 1] = (oa2 & ~(3u << 8)) | 2u << 8; // aux[2r] stays: the earlier stats, rf_code 2
				continue;
			}
			rf_code = 3; // correct.c:470
		}
		__syncwarp(); // (every lane has read the earlier stats before lane 0 overwrites them)
		if (ec_code == 0) {
			uint8_t *seq = P.seq + ob0;
			uint8_t *qual = P.qual && n > 0 && P.qual[ob0] != 0xFF ? P.qual + ob0 : 0;
			const int bp = d.brute >= 0 ? d.brute >> 2 : -1, bb = d.brute & 3;
			const int lim0 = d.start0 + k, lim1 = n - d.start1 - k; // masked: i < lim0 (forward), i >= lim1 (reverse), correct.c:378-379
			mh = (uint32_t)(r0.y > r1.y ? r0.y : r1.y), n_absent = (uint32_t)(r0.x + r1.x);
			for (int i = lane; i < n; i += 32) {
				const int ff = __ldg(P.fl + o + i);
				const int ob = FL_OB(ff), cur = i == bp ? bb : ob, q = (ff & FL_Q) != 0;
				int f = cur, g = cur;
				if (bit1(plane(P, PL_E0V), o + i)) f = bit1(plane(P, PL_E0L), o + i) | bit1(plane(P, PL_E0H), o + i) << 1;
				if (bit1(plane(P, PL_E1V), o + i)) g = bit1(plane(P, PL_E1L), o + i) | bit1(plane(P, PL_E1H), o + i) << 1;
				if (i < lim0) f = 4;
				if (i >= lim1) g = 4;
				int nb; // correct.c:443-450
				if (f == g) nb = f > 3 ? cur : f;
				else if (g > 3) nb = f;
				else if (f > 3) nb = g;
				else nb = ob;
				const bool diff = nb != ob;
				n_ec += diff, n_ec_high += diff && q;
				seq[i] = (uint8_t)((diff ? "acgtn" : "ACGTN")[nb]);
				if (qual) qual[i] = (uint8_t)(diff ? 34 + ob : (q ? '?' : '+'));
			}
			for (int s = 16; s > 0; s >>= 1) {
				n_ec += __shfl_down_sync(0xffffffffu, n_ec, s);
				n_ec_high += __shfl_down_sync(0xffffffffu, n_ec_high, s);
			}
		}
		if (lane == 0) { // correct.c:552-553
			P.aux[2 * r] = (n_ec & 0x3fff) << 18 | (n_ec_high & 0x3fff) << 4 | brute << 3 | ec_code;
			P.aux[2 * r + 1] = (n_absent & 0x3fffff) << 10 | rf_code << 8 | (mh & 0xff);
		}
	}
}

// ------------------------------------------------------------------ host side

// window starts of a device batch: out[0] = number of entries, then (read index, byte offset) pairs, the last one = the end
__global__ void k_ec_cuts(const uint64_t *off, int64_t n, uint64_t limit, uint64_t cap, unsigned long long *out)
{
	uint64_t m = 0;
	for (int64_t r0 = 0; r0 < n && m + 1 < cap;) {
		// largest r1 in (r0, n] with off[r1] - off[r0] <= limit, at least r0 + 1, at most r0 + 2^30
		int64_t lo = r0 + 1, hi = n < r0 + (1LL << 30) ? n : r0 + (1LL << 30);
		const uint64_t b0 = off[r0];
		while (lo < hi) {
			const int64_t mid = lo + (hi - lo + 1) / 2;
			if (off[mid] - b0 <= limit) lo = mid; else hi = mid - 1;
		}
		out[1 + 2 * m] = (unsigned long long)r0, out[2 + 2 * m] = b0, ++m;
		r0 = lo;
		if (r0 >= n) { out[1 + 2 * m] = (unsigned long long)n, out[2 + 2 * m] = off[n], ++m; }
	}
	out[0] = m;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static uint64_t batch_bytes_limit()
{
	const char *e = getenv("BFC_B200_EC_BATCH");
	return e && atoll(e) >= 4096 && atoll(e) <= (1LL << 30) ? (uint64_t)atoll(e) : 1ULL << 30; // window offsets are 31-bit
}

static int edit_cap0()
{
	const char *e = getenv("BFC_B200_EC_EDITS");
	return e && atoi(e) >= 4 ? atoi(e) : 192;
}

extern "C" int bfcg_correct_batch(const bfc_opt_t *opt, const bfc_ch_t *ch, int mode, bfcg_batch_t *batch,
                                  uint32_t *aux, bfcg_stats_t *stats)
{
	int r;
	if ((r = bfcg_rt_init()) != BFCG_OK) return r;
	BfcgRuntime &rt = bfcg_rt();
	if (!opt || !ch || !batch || !aux || !batch->off || bfc_ch_get_k(ch) != opt->k || opt->max_heap < 1 || opt->max_heap > 4000)
		return bfcg_fail(__func__, "invalid arguments", cudaSuccess), BFCG_ERR_ARG;
	if (batch->n_reads == 0) return BFCG_OK;
	const bool host = batch->where == BFCG_HOST;
	const int64_t n = batch->n_reads;

	// a host batch is cut into at least three windows (of at least 16 MB) so that its copies in and out overlap the
	// search even when it is small (the command line's batches are ~120 MB)
	uint64_t limit = batch_bytes_limit();
	if (host && getenv("BFC_B200_EC_BATCH") == 0) limit = std::min<uint64_t>(limit, std::max<uint64_t>((uint64_t)16 << 20, batch->n_bytes / 3 + 1));
	// windows of at most `limit` bytes / 2^30 reads, cut at read boundaries; for a device batch the cuts are
	// found on the device (one thread, a binary search per window) so the offsets never travel to the host
	std::vector<uint64_t> cut_r, cut_b; // read index / byte offset of every window start, plus the end
	if (host) {
		for (int64_t r0 = 0; r0 < n;) { // largest r1 in (r0, n] with off[r1] - off[r0] <= limit, at least r0 + 1
			int64_t lo = r0 + 1, hi = n < r0 + (1LL << 30) ? n : r0 + (1LL << 30);
			while (lo < hi) {
				const int64_t mid = lo + (hi - lo + 1) / 2;
				if (batch->off[mid] - batch->off[r0] <= limit) lo = mid; else hi = mid - 1;
			}
			cut_r.push_back(r0), cut_b.push_back(batch->off[r0]);
			r0 = lo;
		}
		cut_r.push_back(n), cut_b.push_back(batch->off[n]);
	} else {
		const uint64_t cap = batch->n_bytes / limit * 2 + (uint64_t)(n >> 30) + 8;
		unsigned long long *d_cuts = (unsigned long long*)bfcg_arena((2 * cap + 1) * 8);
		if (!d_cuts) return BFCG_ERR_NOMEM;
		k_ec_cuts<<<1, 1, 0, rt.stream>>>(batch->off, n, limit, cap, d_cuts);
		BFCG_LAUNCH_CHECK();
		std::vector<unsigned long long> h(2 * cap + 1);
		BFCG_CUDA(cudaMemcpyAsync(h.data(), d_cuts, (2 * cap + 1) * 8, cudaMemcpyDeviceToHost, rt.stream));
		BFCG_CUDA(cudaStreamSynchronize(rt.stream));
		if (h[0] == 0 || h[0] > cap) return bfcg_fail(__func__, "window cut search failed", cudaSuccess);
		for (uint64_t i = 0; i < h[0]; ++i) cut_r.push_back(h[1 + 2 * i]), cut_b.push_back(h[2 + 2 * i]);
	}
	const int threads = EC_THREADS;
	const bool search_v1 = getenv("BFC_B200_EC_V1") != 0; // A/B: the one-lookup-per-round search kernel
	int ctas = search_v1 ? EC_CTAS_PER_SM : EC2_CTAS_PER_SM;
	{ const char *e = getenv("BFC_B200_EC_CTAS"); if (e && atoi(e) >= 1 && atoi(e) <= ctas) ctas = atoi(e); }
	const int64_t max_slots = (int64_t)rt.sm_count * ctas * threads; // persistent threads: exactly what is resident
	const int heap_cap = opt->max_heap + 6; // the search never holds more than max_heap + 4 states
	const int edit_cap = edit_cap0();

	// one arena for every window: three sets of in/out buffers for host batches (window w+1 is copied in and
	// window w-1 copied out, on their own streams, while window w is searched) + one set of scratch
	uint64_t nb_max = 0;
	int64_t nr_max = 0;
	for (size_t w = 0; w + 1 < cut_r.size(); ++w) {
		nb_max = std::max<uint64_t>(nb_max, cut_b[w + 1] - cut_b[w]);
		nr_max = std::max<int64_t>(nr_max, (int64_t)(cut_r[w + 1] - cut_r[w]));
	}
	const uint64_t n_rec_max = el_padded(nb_max), pl_words_max = (n_rec_max + PL_PAD) / 64 + 4;
	const int64_t slots_max = std::min<int64_t>(max_slots, (2 * nr_max + threads - 1) / threads * threads);
	enum { NBUF = 3 };
	size_t tot = 0, o_seq[NBUF] = {0}, o_qual[NBUF] = {0}, o_off[NBUF] = {0}, o_aux[NBUF] = {0};
	size_t o_pl, o_fl, o_desc, o_jobs, o_res, o_ext, o_pool, o_heapk, o_edits, o_ovf, o_ctr;
	if (host)
		for (int b = 0; b < NBUF; ++b) {
			o_seq[b] = tot; tot = align_up(tot + nb_max, 256);
			o_qual[b] = tot; tot = align_up(tot + nb_max, 256);
			o_off[b] = tot; tot = align_up(tot + (nr_max + 1) * 8, 256);
			o_aux[b] = tot; tot = align_up(tot + nr_max * 8, 256);
		}
	o_pl = tot; tot = align_up(tot + (size_t)PL_N * pl_words_max * 8, 256);
	o_fl = tot; tot = align_up(tot + n_rec_max * 2, 256);
	o_desc = tot; tot = align_up(tot + nr_max * sizeof(ReadDesc), 256);
	o_jobs = tot; tot = align_up(tot + 2 * nr_max * sizeof(int4), 256);
	o_res = tot; tot = align_up(tot + 2 * nr_max * sizeof(int2), 256);
	o_ext = tot; tot = align_up(tot + 2 * nr_max * 8, 256);
	o_pool = tot; tot = align_up(tot + (size_t)slots_max * heap_cap * sizeof(EcState), 256);
	o_heapk = tot; tot = align_up(tot + (size_t)slots_max * heap_cap * 4, 256);
	o_edits = tot; tot = align_up(tot + (size_t)slots_max * edit_cap * sizeof(uint2), 256);
	o_ovf = tot; tot = align_up(tot + 2 * nr_max * 4, 256);
	o_ctr = tot; tot += 256;
	uint8_t *a = (uint8_t*)bfcg_arena(tot);
	if (!a) return BFCG_ERR_NOMEM;

	const size_t n_win = cut_r.size() - 1;
	// the per-read stats of a host batch come back through a pinned ring: a copy straight into pageable caller
	// memory would block the host until the window is done, and the next window's kernels with it
	uint32_t *pin_aux = 0;
	if (host && !(pin_aux = (uint32_t*)bfcg_pinned(0, (size_t)NBUF * nr_max * 8))) return BFCG_ERR_NOMEM;
	auto drain_aux = [&](size_t w) { // window w's stats: pinned ring -> caller
		const int64_t r0 = (int64_t)cut_r[w], nr = (int64_t)cut_r[w + 1] - r0;
		cudaEventSynchronize(rt.ev_out[w % NBUF]);
		memcpy(aux + 2 * r0, pin_aux + (size_t)(w % NBUF) * nr_max * 2, (size_t)nr * 8);
	};
	auto issue_copy_in = [&](size_t w) -> cudaError_t {
		const int64_t r0 = (int64_t)cut_r[w], nr = (int64_t)cut_r[w + 1] - r0;
		const uint64_t b0 = cut_b[w], nb = cut_b[w + 1] - b0;
		const int b = (int)(w % NBUF);
		cudaError_t ce;
		if (w >= NBUF && (ce = cudaStreamWaitEvent(rt.copy_in, rt.ev_out[b], 0)) != cudaSuccess) return ce; // buffer b was read out
		if ((ce = cudaMemcpyAsync(a + o_seq[b], batch->seq + b0, nb, cudaMemcpyHostToDevice, rt.copy_in)) != cudaSuccess) return ce;
		if (batch->qual && (ce = cudaMemcpyAsync(a + o_qual[b], batch->qual + b0, nb, cudaMemcpyHostToDevice, rt.copy_in)) != cudaSuccess) return ce;
		if ((ce = cudaMemcpyAsync(a + o_off[b], batch->off + r0, (nr + 1) * 8, cudaMemcpyHostToDevice, rt.copy_in)) != cudaSuccess) return ce;
		if (opt->refine_ec && (ce = cudaMemcpyAsync(a + o_aux[b], aux + 2 * r0, nr * 8, cudaMemcpyHostToDevice, rt.copy_in)) != cudaSuccess) return ce; // earlier stats in
		return cudaEventRecord(rt.ev_in[b], rt.copy_in);
	};

	BfcgTimer timer(stats);
	if (host) {
		BFCG_CUDA(cudaStreamSynchronize(rt.stream)); // earlier work on the engine's stream may still use the arena
		BFCG_CUDA(issue_copy_in(0));
	}
	int rc = BFCG_OK;
	for (size_t w = 0; w < n_win; ++w) {
		const int64_t r0 = (int64_t)cut_r[w], r1 = (int64_t)cut_r[w + 1];
		const int64_t nr = r1 - r0;
		const uint64_t b0 = cut_b[w], nb = cut_b[w + 1] - b0;
		const uint64_t n_rec = el_padded(nb);
		const uint64_t pl_words = (n_rec + PL_PAD) / 64 + 4;
		const int64_t slots = std::min<int64_t>(max_slots, (2 * nr + threads - 1) / threads * threads);
		const int ib = (int)(w % NBUF);

		EcParams P;
		memset(&P, 0, sizeof(P));
		if (host) { // offsets stay absolute (as in a device batch): the staged window starts at stream offset b0
			const cudaError_t ce = cudaStreamWaitEvent(rt.stream, rt.ev_in[ib], 0);
			if (ce != cudaSuccess) { rc = bfcg_fail(__func__, "staging copy", ce); break; }
			P.off = (const uint64_t*)(a + o_off[ib]), P.seq = a + o_seq[ib] - b0, P.qual = batch->qual ? a + o_qual[ib] - b0 : 0;
			P.aux = (uint32_t*)(a + o_aux[ib]);
		} else {
			P.off = batch->off + r0, P.seq = batch->seq, P.qual = batch->qual;
			P.aux = aux + 2 * r0;
		}
		P.base0 = b0;
		P.pl = (const uint64_t*)(a + o_pl), P.pl_words = pl_words;
		P.fl = (const uint16_t*)(a + o_fl);
		P.desc = (ReadDesc*)(a + o_desc), P.jobs = (int4*)(a + o_jobs), P.res = (int2*)(a + o_res);
		P.n_reads = nr, P.n_jobs = 2 * nr;
		P.ext = opt->max_end_ext >= 0 && opt->max_end_ext < EC_EXT_STEPS && !getenv("BFC_B200_EC_NOEXT") ? (uint64_t*)(a + o_ext) : 0;
		P.tab = tab_view(ch);
		P.k = opt->k, P.q = opt->q, P.min_cov = opt->min_cov, P.win_multi_ec = opt->win_multi_ec, P.max_end_ext = opt->max_end_ext;
		P.w_ec = opt->w_ec, P.w_ec_high = opt->w_ec_high, P.w_absent = opt->w_absent, P.w_absent_high = opt->w_absent_high;
		P.max_path_diff = opt->max_path_diff, P.max_heap = opt->max_heap, P.mode = mode, P.refine = opt->refine_ec != 0;
		P.pool = (EcState*)(a + o_pool), P.heapk = (uint32_t*)(a + o_heapk), P.edits = (uint2*)(a + o_edits);
		P.heap_cap = heap_cap, P.edit_cap = edit_cap;
		P.overflow = (uint32_t*)(a + o_ovf), P.ctr = (unsigned long long*)(a + o_ctr);
		BFCG_CUDA(cudaMemsetAsync(P.ctr, 0, 64, rt.stream));
		BFCG_CUDA(cudaMemsetAsync(a + o_pl, 0, (size_t)PL_N * pl_words * 8, rt.stream));

		const uint8_t *w_seq = host ? a + o_seq[ib] : batch->seq + b0;
		const uint8_t *w_qual = host ? (batch->qual ? a + o_qual[ib] : 0) : (batch->qual ? batch->qual + b0 : 0);
		LookupParams lp;
		lp.tab = P.tab, lp.seq = w_seq, lp.qual = w_qual, lp.n_pos = nb;
		lp.pl = (uint64_t*)(a + o_pl), lp.pl_words = pl_words, lp.k = opt->k, lp.q = opt->q, lp.min_cov = opt->min_cov, lp.refine = opt->refine_ec != 0, lp.ctr = P.ctr;
		{ KTime kt(KT_EC_LOOKUP); k_ec_lookup<<<(unsigned)(n_rec / EL_SEG), EL_THREADS, 0, rt.stream>>>(lp); }
		BFCG_LAUNCH_CHECK();
		{
			KTime kt(KT_EC_SETUP);
			k_ec_cov<<<(unsigned)((n_rec / 32 + 255) / 256), 256, 0, rt.stream>>>(P, n_rec, (uint16_t*)(a + o_fl), (uint64_t*)(a + o_pl));
			k_ec_setup<<<(unsigned)((nr + 127) / 128), 128, 0, rt.stream>>>(P);
			k_ec_rescue<<<rt.sm_count * 8, 256, 0, rt.stream>>>(P);
		}
		++rt.n_launches;
		BFCG_LAUNCH_CHECK();
		++rt.n_launches;
		if (P.ext) {
			{ KTime kt(KT_EC_EXT); k_ec_ext<<<(unsigned)((2 * nr + 255) / 256), 256, 0, rt.stream>>>(P); }
			BFCG_LAUNCH_CHECK();
		}
		{
			KTime kt(KT_CORRECT);
			if (search_v1) k_ec_search<<<(unsigned)(slots / threads), threads, 0, rt.stream>>>(P);
			else if (ctas >= 5) k_ec_search2<5><<<(unsigned)(slots / threads), threads, 0, rt.stream>>>(P);
			else k_ec_search2<4><<<(unsigned)(slots / threads), threads, 0, rt.stream>>>(P);
		}
		BFCG_LAUNCH_CHECK();
		if (host && w + 1 < n_win) { // the next window travels while this one is searched (issued after the launches:
			const cudaError_t ce = issue_copy_in(w + 1); // a copy from pageable memory blocks the host, not the GPU)
			if (ce != cudaSuccess) { rc = bfcg_fail(__func__, "staging copy", ce); break; }
		}
		if (host && w >= 1) drain_aux(w - 1);
		unsigned long long c[4];
		BFCG_CUDA(cudaMemcpyAsync(c, P.ctr, sizeof(c), cudaMemcpyDeviceToHost, rt.stream));
		BFCG_CUDA(cudaStreamSynchronize(rt.stream));
		// jobs whose edit list outgrew the scratch: same kernel, fewer threads, larger lists
		uint32_t *redo = 0;
		uint2 *big = 0;
		for (int cap = edit_cap * 16; c[0] > 0; cap *= 16) {
			const uint64_t n_redo = c[0];
			if (stats) stats->n_redo += n_redo;
			if (cap > (1 << 26)) { cudaFree(redo); cudaFree(big); return bfcg_fail(__func__, "search edit list overflow beyond 2^26 entries", cudaSuccess), BFCG_ERR_OVERFLOW; }
			const int64_t rs = std::min<int64_t>(std::min<int64_t>(slots, (int64_t)((n_redo + 31) / 32 * 32)), std::max<int64_t>(32, (int64_t)((4ULL << 30) / ((uint64_t)cap * sizeof(uint2))) / 32 * 32));
			cudaFree(redo); cudaFree(big);
			redo = 0, big = 0;
			BFCG_CUDA(cudaMalloc(&redo, n_redo * 4));
			BFCG_CUDA(cudaMalloc(&big, (size_t)rs * cap * sizeof(uint2)));
			BFCG_CUDA(cudaMemcpyAsync(redo, P.overflow, n_redo * 4, cudaMemcpyDeviceToDevice, rt.stream));
			BFCG_CUDA(cudaMemsetAsync(P.ctr, 0, 8, rt.stream));
			BFCG_CUDA(cudaMemsetAsync(P.ctr + 2, 0, 8, rt.stream));
			EcParams Q = P;
			Q.redo = redo, Q.n_jobs = (int64_t)n_redo, Q.edits = big, Q.edit_cap = cap;
			{
				KTime kt(KT_CORRECT_REDO);
				if (search_v1) k_ec_search<<<(unsigned)(rs / 32), 32, 0, rt.stream>>>(Q);
				else k_ec_search2<4><<<(unsigned)(rs / 32), 32, 0, rt.stream>>>(Q);
			}
			BFCG_LAUNCH_CHECK();
			BFCG_CUDA(cudaMemcpyAsync(c, P.ctr, sizeof(c), cudaMemcpyDeviceToHost, rt.stream));
			BFCG_CUDA(cudaStreamSynchronize(rt.stream));
		}
		cudaFree(redo); cudaFree(big);
		{ KTime kt(KT_EC_MERGE); k_ec_merge<<<(unsigned)std::min<int64_t>((nr + 7) / 8, (int64_t)rt.sm_count * 16), 256, 0, rt.stream>>>(P); }
		BFCG_LAUNCH_CHECK();
		if (stats) stats->n_lookups += c[1], stats->n_search_lookups += c[3];
		if (host) { // the corrected window leaves on the copy-out stream while the next one is searched
			cudaError_t ce = cudaEventRecord(rt.ev_done[ib], rt.stream);
			if (ce == cudaSuccess) ce = cudaStreamWaitEvent(rt.copy_out, rt.ev_done[ib], 0);
			if (ce == cudaSuccess) ce = cudaMemcpyAsync(batch->seq + b0, a + o_seq[ib], nb, cudaMemcpyDeviceToHost, rt.copy_out);
			if (ce == cudaSuccess && batch->qual) ce = cudaMemcpyAsync(batch->qual + b0, a + o_qual[ib], nb, cudaMemcpyDeviceToHost, rt.copy_out);
			if (ce == cudaSuccess) ce = cudaMemcpyAsync(pin_aux + (size_t)ib * nr_max * 2, a + o_aux[ib], nr * 8, cudaMemcpyDeviceToHost, rt.copy_out);
			if (ce == cudaSuccess) ce = cudaEventRecord(rt.ev_out[ib], rt.copy_out);
			if (ce != cudaSuccess) { rc = bfcg_fail(__func__, "copy out", ce); break; }
		}
	}
	if (host) { cudaStreamSynchronize(rt.copy_in); cudaStreamSynchronize(rt.copy_out); }
	if (rc != BFCG_OK) { cudaStreamSynchronize(rt.stream); return rc; }
	if (host) drain_aux(n_win - 1);
	timer.stop();
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	return BFCG_OK;
}
