/* TEST INFRASTRUCTURE ONLY -- oracle restatement of the reference correction phase
 * (correct.c) and of the `-1` trim lookup.  Sequential, plain C, growable scratch.
 * Pinned against oracle/_ref (the unmodified reference) by tests/test_oracle_vs_ref.py. */
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include "oracle_kmer.h"

/* ---------------------------------------------------------------- per-base record */

typedef struct { /* reference correct.c:14-19 (ecbase_t), fields unpacked */
	uint8_t b, q, ob, oq;
	uint8_t lcov, hcov;       /* 6-bit fields in the reference; never exceed k <= 63 */
	uint8_t solid_end, high_end;
	uint8_t ec, absent;
	int i;
} obase_t;

typedef struct { size_t n, m; obase_t *a; } oseq_t;

static void oseq_reserve(oseq_t *s, size_t n)
{
	if (s->m < n) { s->m = n + 16; s->a = (obase_t*)realloc(s->a, s->m * sizeof(obase_t)); }
}

/* reference correct.c:23-37 */
static int seq_convert(const char *s, const char *q, int qthres, oseq_t *seq, int b_from_q)
{
	int i, l = (int)strlen(s);
	oseq_reserve(seq, l);
	seq->n = l;
	for (i = 0; i < l; ++i) {
		obase_t *c = &seq->a[i];
		int b;
		memset(c, 0, sizeof(*c));
		if (b_from_q && q && q[i] - 33 <= 5) b = (q[i] - 34) & 7; /* 3-bit field */
		else b = orc_nt6[(uint8_t)s[i]] - 1;
		c->b = c->ob = (uint8_t)b;
		c->q = c->oq = !q ? 1 : q[i] - 33 >= qthres ? 1 : 0;
		if (c->b > 3) c->q = c->oq = 0;
		c->i = i;
	}
	return l;
}

/* reference correct.c:39-57 */
static inline void base_comp(obase_t *c)
{
	c->b = c->b < 4 ? 3 - c->b : 4;
	c->ob = c->ob < 4 ? 3 - c->ob : 4;
}

static void seq_revcomp(oseq_t *s)
{
	size_t i, n = s->n;
	for (i = 0; i < n >> 1; ++i) {
		obase_t t = s->a[i];
		s->a[i] = s->a[n - 1 - i];
		s->a[n - 1 - i] = t;
	}
	for (i = 0; i < n; ++i) base_comp(&s->a[i]);
}

/* ---------------------------------------------------------------- search state */

typedef struct { uint8_t ec, ec_high, absent, absent_high, b; } open_t; /* correct.c:149-151 */

typedef struct { /* correct.c:153-160 */
	int tot_pen, i, k;
	int32_t ecpos_high[2], ecpos[5];
	orc_kmer_t x;
} oheap_t;

typedef struct { /* correct.c:162-167 */
	int parent, i, tot_pen;
	uint8_t b;
	open_t pen;
	uint16_t cnt;
} ostack_t;

struct orc_ecbuf_s {
	const orc_opt_t *opt;
	const orc_ch_t *ch;
	int mode;
	size_t heap_n, heap_m, stack_n, stack_m;
	oheap_t *heap;
	ostack_t *stack;
	oseq_t seq, ec[2];
	uint64_t counters[3]; /* lookups, pops, max stack */
};

orc_ecbuf_t *orc_ecbuf_new(const orc_opt_t *opt, const orc_ch_t *ch, int mode)
{
	orc_ecbuf_t *e = (orc_ecbuf_t*)calloc(1, sizeof(orc_ecbuf_t));
	e->opt = opt, e->ch = ch, e->mode = mode;
	return e;
}

void orc_ecbuf_free(orc_ecbuf_t *e)
{
	if (!e) return;
	free(e->heap); free(e->stack); free(e->seq.a); free(e->ec[0].a); free(e->ec[1].a);
	free(e);
}

const uint64_t *orc_ecbuf_counters(const orc_ecbuf_t *e) { return e->counters; }

static inline int occ(orc_ecbuf_t *e, const orc_kmer_t *x)
{
	++e->counters[0];
	return orc_ch_kmer_occ(e->ch, x);
}

/* klib binary heap, "less" = larger tot_pen (correct.c:179), so the root is the
 * smallest penalty.  Tie behaviour follows ksort.h:125-146 exactly. */
static void heap_sift_down(oheap_t *l, size_t n)
{
	size_t i = 0, c;
	oheap_t tmp = l[0];
	while ((c = 2 * i + 1) < n) {
		if (c != n - 1 && l[c].tot_pen > l[c + 1].tot_pen) ++c; /* right child only if strictly smaller */
		if (l[c].tot_pen > tmp.tot_pen) break;                   /* ties keep sinking */
		l[i] = l[c]; i = c;
	}
	l[i] = tmp;
}

static void heap_sift_up(oheap_t *l, size_t n)
{
	size_t c = n - 1;
	oheap_t tmp = l[c];
	while (c) {
		size_t p = (c - 1) >> 1;
		if (tmp.tot_pen > l[p].tot_pen) break;                   /* ties keep rising */
		l[c] = l[p]; c = p;
	}
	l[c] = tmp;
}

static inline int pen_weight(const orc_opt_t *o, open_t p)
{
	return o->w_ec * p.ec + o->w_ec_high * p.ec_high + o->w_absent * p.absent + o->w_absent_high * p.absent_high;
}

/* reference correct.c:198-230 (buf_update) */
static void push_state(orc_ecbuf_t *e, const oheap_t *prev, open_t pen, int cnt)
{
	ostack_t *q;
	oheap_t *r;
	if (e->stack_n == e->stack_m) {
		e->stack_m = e->stack_m ? e->stack_m << 1 : 256;
		e->stack = (ostack_t*)realloc(e->stack, e->stack_m * sizeof(ostack_t));
	}
	q = &e->stack[e->stack_n++];
	q->parent = prev->k, q->i = prev->i, q->b = pen.b, q->pen = pen;
	q->cnt = cnt > 0 ? cnt & 0xff : 0;
	q->tot_pen = prev->tot_pen + pen_weight(e->opt, pen);
	if (e->stack_n > e->counters[2]) e->counters[2] = e->stack_n;
	if (e->heap_n == e->heap_m) {
		e->heap_m = e->heap_m ? e->heap_m << 1 : 128;
		e->heap = (oheap_t*)realloc(e->heap, e->heap_m * sizeof(oheap_t));
	}
	r = &e->heap[e->heap_n++];
	r->i = prev->i + 1;
	r->k = (int)e->stack_n - 1;
	r->x = prev->x;
	if (pen.ec_high) {
		r->ecpos_high[1] = prev->ecpos_high[0];
		r->ecpos_high[0] = prev->i;
	} else memcpy(r->ecpos_high, prev->ecpos_high, sizeof(r->ecpos_high));
	if (pen.ec) {
		memcpy(r->ecpos + 1, prev->ecpos, 4 * sizeof(int32_t));
		r->ecpos[0] = prev->i;
	} else memcpy(r->ecpos, prev->ecpos, sizeof(r->ecpos));
	r->tot_pen = q->tot_pen;
	okm_append(e->opt->k, r->x.x, pen.b);
	heap_sift_up(e->heap, e->heap_n);
}

/* reference correct.c:232-247 */
static int backtrack(const ostack_t *s, int end, const oseq_t *seq, oseq_t *path)
{
	int n_absent = 0;
	oseq_reserve(path, seq->n);
	path->n = seq->n;
	while (end >= 0) {
		int i = s[end].i;
		if ((size_t)i < seq->n) {
			path->a[i].b = s[end].b;
			path->a[i].ec = s[end].pen.ec;
			path->a[i].absent = s[end].pen.absent;
			n_absent += s[end].pen.absent;
		}
		end = s[end].parent;
	}
	return n_absent;
}

/* reference correct.c:249-386 (bfc_ec1dir) */
static int search_dir(orc_ecbuf_t *e, const oseq_t *seq, oseq_t *ec, int start, int end, int *max_heap)
{
	const orc_opt_t *o = e->opt;
	const int k = o->k, n = (int)seq->n;
	oheap_t z;
	int i, run, rv = -1, paths[4], n_paths = 0, best = -1, best_pen = INT_MAX, n_fail = 0;

	e->heap_n = e->stack_n = 0;
	*max_heap = 0;
	memset(&z, 0, sizeof(z));
	oseq_reserve(ec, n);
	ec->n = n;
	/* seed: the k-1 bases before position z.i (correct.c:260-267) */
	for (z.i = start, run = 0; z.i < end; ++z.i) {
		int c = seq->a[z.i].b;
		if (c < 4) {
			if (++run == k) break;
			okm_append(k, z.x.x, c);
		} else run = 0, memset(&z.x, 0, sizeof(z.x));
	}
	if (z.i >= end) abort(); /* the reference asserts a solid k-mer exists */
	z.k = -1;
	for (i = 0; i < 5; ++i) z.ecpos[i] = -1;
	for (i = 0; i < 2; ++i) z.ecpos_high[i] = -1;
	if (e->heap_m == 0) { e->heap_m = 128; e->heap = (oheap_t*)malloc(e->heap_m * sizeof(oheap_t)); }
	e->heap[e->heap_n++] = z;
	for (i = 0; i < n; ++i) ec->a[i] = seq->a[i], ec->a[i].ec = ec->a[i].absent = 0;

	for (;;) {
		int stop = 0;
		*max_heap = *max_heap > 255 ? 255 : *max_heap > (int)e->heap_n ? *max_heap : (int)e->heap_n;
		if (e->heap_n == 0) { rv = -2; break; }
		z = e->heap[0];
		e->heap[0] = e->heap[--e->heap_n];
		heap_sift_down(e->heap, e->heap_n);
		++e->counters[1];
		if (best >= 0 && z.tot_pen > best_pen + o->max_path_diff) break;
		if (z.i - end > o->max_end_ext) stop = 1;
		if (!stop) {
			const obase_t *c = z.i < n ? &seq->a[z.i] : 0;
			int b, os = -1, fixed = 0, other_ext = 0, n_added = 0, added_cnt[4];
			open_t added[4];
			if (z.i > end) fixed = 1;
			if (c && c->b < 4) {
				orc_kmer_t x = z.x;
				okm_append(k, x.x, c->b);
				os = occ(e, &x);
				if (c->q && (os & 0xff) >= o->min_cov + 1 && c->lcov >= o->min_cov + 1) fixed = 1;
				else if (c->hcov > k * .75) fixed = 1;
			}
			for (b = 0; b < 4; ++b) {
				open_t pen;
				if (fixed && c && b != c->b) continue;
				if (c == 0 || b != c->b) {
					int s;
					orc_kmer_t x = z.x;
					if (c) {
						if (c->q && z.ecpos_high[1] >= 0 && z.i - z.ecpos_high[1] < o->win_multi_ec) continue;
						if (z.ecpos[4] >= 0 && z.i - z.ecpos[4] < o->win_multi_ec) continue;
					}
					okm_append(k, x.x, b);
					s = occ(e, &x);
					if (s < 0 || (s & 0xff) < o->min_cov) continue;
					pen.ec = c && c->b < 4 ? 1 : 0;
					pen.ec_high = pen.ec ? c->oq : 0;
					pen.absent = 0;
					pen.absent_high = ((s >> 8 & 0xff) < o->min_cov);
					pen.b = (uint8_t)b;
					added_cnt[n_added] = s;
					added[n_added++] = pen;
					++other_ext;
				} else {
					pen.ec = pen.ec_high = 0;
					pen.absent = (os < 0 || (os & 0xff) < o->min_cov);
					pen.absent_high = (os < 0 || (os >> 8 & 0xff) < o->min_cov);
					pen.b = (uint8_t)b;
					added_cnt[n_added] = os;
					added[n_added++] = pen;
				}
			}
			if (fixed == 0 && other_ext == 0) ++n_fail;
			if (n_fail > n * 2) { rv = -3; break; }
			if (c || n_added == 1) {
				if (n_added > 1 && (int)e->heap_n > o->max_heap) {
					int min_b = -1, min = INT_MAX;
					for (b = 0; b < n_added; ++b) {
						int t = pen_weight(o, added[b]);
						if (min > t) min = t, min_b = b;
					}
					push_state(e, &z, added[min_b], added_cnt[min_b]);
				} else {
					for (b = 0; b < n_added; ++b) push_state(e, &z, added[b], added_cnt[b]);
				}
			} else {
				if (n_added == 0)
					e->stack[z.k].tot_pen += o->w_absent * (o->max_end_ext - (z.i - end));
				stop = 1;
			}
		}
		if (stop) {
			if (e->stack[z.k].tot_pen < best_pen) best_pen = e->stack[z.k].tot_pen, best = n_paths;
			paths[n_paths++] = z.k;
			if (n_paths == 4) break;
		}
	}
	if (n_paths == 0) return rv;
	rv = backtrack(e->stack, paths[best], seq, ec);
	for (i = 0; i < n; ++i)
		if (i < start + k || i >= end) ec->a[i].b = 4;
	return rv;
}

/* reference correct.c:63-80 */
static int greedy_k(orc_ecbuf_t *e, const orc_kmer_t *x)
{
	int i, j, k = e->opt->k, max = 0, max_ec = -1, max2 = 0;
	for (i = 0; i < k; ++i) {
		int c = (int)((x->x[1] >> i & 1) << 1 | (x->x[0] >> i & 1));
		for (j = 0; j < 4; ++j) {
			orc_kmer_t y = *x;
			int ret;
			if (j == c) continue;
			okm_change(k, y.x, i, j);
			ret = occ(e, &y);
			if (ret < 0) continue;
			if ((max & 0xff) < (ret & 0xff)) max2 = max, max = ret, max_ec = i << 2 | j;
			else if ((max2 & 0xff) < (ret & 0xff)) max2 = ret;
		}
	}
	return (max & 0xff) * 3 > e->mode && (max2 & 0xff) < 3 ? max_ec : -1;
}

/* reference correct.c:82-94 */
static int first_kmer(int k, const oseq_t *s, int start, orc_kmer_t *x)
{
	int i, l;
	memset(x, 0, sizeof(*x));
	for (i = start, l = 0; i < (int)s->n; ++i) {
		if (s->a[i].b < 4) {
			okm_append(k, x->x, s->a[i].b);
			if (++l == k) break;
		} else l = 0, memset(x, 0, sizeof(*x));
	}
	return i;
}

/* reference correct.c:96-117 */
static void kmer_cov(orc_ecbuf_t *e, oseq_t *s)
{
	int i, j, l, r, k = e->opt->k, min_occ = e->opt->min_cov;
	orc_kmer_t x;
	memset(&x, 0, sizeof(x));
	for (i = 0; i < (int)s->n; ++i)
		s->a[i].high_end = s->a[i].solid_end = s->a[i].lcov = s->a[i].hcov = 0;
	for (i = l = 0; i < (int)s->n; ++i) {
		obase_t *c = &s->a[i];
		if (c->b >= 4) { l = 0; memset(&x, 0, sizeof(x)); continue; }
		okm_append(k, x.x, c->b);
		if (++l < k) continue;
		if ((r = occ(e, &x)) < 0) continue;
		if ((r >> 8 & 0x3f) >= min_occ + 1) c->high_end = 1;
		if ((r & 0xff) >= min_occ) {
			c->solid_end = 1;
			for (j = i - k + 1; j <= i; ++j) ++s->a[j].lcov, s->a[j].hcov += c->high_end;
		}
	}
}

/* reference correct.c:119-130 */
static uint64_t best_island(int k, const oseq_t *s)
{
	int i, l = 0, max = 0, max_i = -1;
	for (i = k - 1; i < (int)s->n; ++i) {
		if (!s->a[i].solid_end) {
			if (l > max) max = l, max_i = i;
			l = 0;
		} else ++l;
	}
	if (l > max) max = l, max_i = i;
	return max > 0 ? (uint64_t)(max_i - max - k + 1) << 32 | (uint32_t)max_i : 0;
}

/* reference correct.c:388-472 (bfc_ec1) + the packing of correct.c:552-553.
 * Refine mode (-R): `ori` = the read's earlier stats (e->ori_st in the reference, correct.c:176, 543) packed like
 * the result: ori[0] = aux, ori[1] = aux2.  Bases that an earlier round corrected are taken back from the quality
 * string (correct.c:31); if this round leaves more absent k-mers than the earlier one the earlier stats come back
 * with rf_code 2 and the read stays as it is (correct.c:438-442), else rf_code is 3 (correct.c:470). */
uint64_t orc_ec1(orc_ecbuf_t *e, char *seq, char *qual, const uint32_t *ori)
{
	const orc_opt_t *o = e->opt;
	int i, start = 0, end = 0, n_n = 0, rv[2], max_heap[2], n;
	uint32_t ec_code = 1, brute = 0, n_ec = 0, n_ec_high = 0, n_absent = 0, mh = 0, rf_code = o->refine_ec ? 1 : 0;
	uint64_t r;

	seq_convert(seq, qual, o->q, &e->seq, o->refine_ec);
	n = (int)e->seq.n;
	for (i = 0; i < n; ++i) n_n += e->seq.a[i].ob > 3;
	if (n_n > n * .05) { ec_code = 2; goto done; }
	kmer_cov(e, &e->seq);
	r = best_island(o->k, &e->seq);
	if (r == 0) {
		orc_kmer_t x;
		int ec = -1;
		while ((end = first_kmer(o->k, &e->seq, start, &x)) < n) {
			ec = greedy_k(e, &x);
			if (ec >= 0) break;
			if (end + (o->k >> 1) >= n) break;
			start = end - (o->k >> 1);
		}
		if (ec >= 0) {
			e->seq.a[end - (ec >> 2)].b = ec & 3;
			++end; start = end - o->k;
			brute = 1;
		} else { ec_code = 3; goto done; }
	} else start = (int)(r >> 32), end = (int)(uint32_t)r;
	if ((rv[0] = search_dir(e, &e->seq, &e->ec[0], start, n, &max_heap[0])) < 0) {
		ec_code = rv[0] == -2 ? 4 : rv[0] == -3 ? 5 : 1;
		goto done;
	}
	seq_revcomp(&e->seq);
	if ((rv[1] = search_dir(e, &e->seq, &e->ec[1], n - end, n, &max_heap[1])) < 0) {
		ec_code = rv[1] == -2 ? 4 : rv[1] == -3 ? 5 : 1;
		goto done;
	}
	mh = max_heap[0] > max_heap[1] ? max_heap[0] : max_heap[1];
	ec_code = 0, n_absent = rv[0] + rv[1];
	seq_revcomp(&e->ec[1]);
	seq_revcomp(&e->seq);
	if (o->refine_ec && ori && (ori[0] & 7) == 0 && n_absent > ori[1] >> 10) /* correct.c:438-442 */
		return (uint64_t)((ori[1] & ~(3u << 8)) | 2u << 8) << 32 | ori[0];
	for (i = 0; i < n; ++i) {
		obase_t *c = &e->seq.a[i];
		int f = e->ec[0].a[i].b, g = e->ec[1].a[i].b;
		if (f == g) c->b = f > 3 ? c->b : f;
		else if (g > 3) c->b = f;
		else if (f > 3) c->b = g;
		else c->b = c->ob;
	}
	for (i = 0; i < n; ++i) {
		const obase_t *c = &e->seq.a[i];
		int diff = c->b != c->ob;
		if (diff) { ++n_ec; if (c->q) ++n_ec_high; }
		seq[i] = (diff ? "acgtn" : "ACGTN")[c->b];
		if (qual) qual[i] = diff ? 34 + c->ob : "+?"[c->q];
	}
	if (o->refine_ec) rf_code = 3; /* correct.c:470 */
done:
	{
		uint32_t aux = (n_ec & 0x3fff) << 18 | (n_ec_high & 0x3fff) << 4 | brute << 3 | ec_code;
		uint32_t aux2 = (n_absent & 0x3fffff) << 10 | rf_code << 8 | (mh & 0xff);
		return (uint64_t)aux2 << 32 | aux;
	}
}

void orc_correct_batch(const orc_opt_t *opt, const orc_ch_t *ch, int mode, int64_t n_reads,
                       const uint64_t *off, uint8_t *seq, uint8_t *qual, uint32_t *aux, uint64_t counters[3])
{
	int64_t i;
	orc_ecbuf_t *e = orc_ecbuf_new(opt, ch, mode);
	for (i = 0; i < n_reads; ++i) {
		int len = (int)(off[i + 1] - off[i] - 1);
		char *q = qual && (len == 0 || qual[off[i]] != 0) ? (char*)qual + off[i] : 0;
		uint64_t r = orc_ec1(e, (char*)seq + off[i], q, opt->refine_ec ? aux + 2 * i : 0); /* refine: aux is in/out */
		aux[2 * i] = (uint32_t)r, aux[2 * i + 1] = (uint32_t)(r >> 32);
	}
	if (counters) memcpy(counters, e->counters, sizeof(e->counters));
	orc_ecbuf_free(e);
}

/* ---------------------------------------------------------------- trim */

/* reference correct.c:478-497 */
uint64_t orc_max_streak(int k, const orc_bf_t *bf, const char *seq, int l_seq)
{
	int i, l = 0;
	uint64_t max = 0, t = 0, x[4] = {0, 0, 0, 0};
	for (i = 0; i < l_seq; ++i) {
		int c = orc_nt6[(uint8_t)seq[i]] - 1;
		if (c < 4) {
			okm_append(k, x, c);
			if (++l >= k) {
				uint64_t y[2], hash = okm_hash(k, x, y);
				if (orc_bf_get(bf, hash) == bf->n_hashes) t += 1ULL << 32;
				else t = i + 1;
			} else t = i + 1;
		} else l = 0, memset(x, 0, sizeof(x)), t = i + 1;
		max = max > t ? max : t;
	}
	return max;
}

/* reference correct.c:554-569 */
void orc_trim_batch(const orc_opt_t *opt, const orc_bf_t *bf, int64_t n_reads, const uint64_t *off,
                    const uint8_t *seq, uint8_t *keep, int32_t *tstart, int32_t *tend)
{
	int64_t i;
	for (i = 0; i < n_reads; ++i) {
		int len = (int)(off[i + 1] - off[i] - 1);
		uint64_t max = orc_max_streak(opt->k, bf, (const char*)seq + off[i], len);
		keep[i] = 0, tstart[i] = tend[i] = 0;
		if (max >> 32 && (double)((max >> 32) + opt->k) / len > opt->min_frac) {
			int start = (int)(uint32_t)max, end = start + (int)(max >> 32);
			start -= opt->k - 1;
			keep[i] = 1, tstart[i] = start, tend[i] = end;
		}
	}
}
