/* flat.c -- gathers one pipeline batch of bseq1_t records into the flat layout the
 * engine takes (bfc_b200.h) and scatters corrected bases back.  The copies are spread
 * over the -t helper threads with kt_for. */
#include <stdlib.h>
#include <string.h>
#include "bfc.h"
#include "flat.h"

static void copy_worker(void *data, long i, int tid)
{
	flat_batch_t *f = (flat_batch_t*)data;
	const bseq1_t *s = &f->seqs[i];
	const uint64_t o = f->off[i];
	(void)tid;
	if (!f->to_reads) {
		memcpy(f->b.seq + o, s->seq, (size_t)s->l_seq);
		f->b.seq[o + s->l_seq] = 0;
		if (f->b.qual) {
			if (s->qual) memcpy(f->b.qual + o, s->qual, (size_t)s->l_seq);
			else memset(f->b.qual + o, 0xFF, (size_t)s->l_seq); /* "no quality" marker */
			f->b.qual[o + s->l_seq] = 0;
		}
	} else {
		memcpy(s->seq, f->b.seq + o, (size_t)s->l_seq);
		if (s->qual && f->b.qual) memcpy(s->qual, f->b.qual + o, (size_t)s->l_seq);
	}
}

int flat_from_reads(flat_batch_t *f, const bseq1_t *seqs, int n, int n_threads)
{
	int i, any_qual = 0;
	uint64_t tot = 0;
	memset(f, 0, sizeof(*f));
	f->off = (uint64_t*)malloc((size_t)(n + 1) * sizeof(uint64_t));
	for (i = 0; i < n; ++i) {
		f->off[i] = tot;
		tot += (uint64_t)seqs[i].l_seq + 1;
		any_qual |= seqs[i].qual != 0;
	}
	f->off[n] = tot;
	f->b.n_reads = n, f->b.n_bytes = tot, f->b.where = BFCG_HOST, f->b.off = f->off;
	f->b.seq = (uint8_t*)bfcg_host_alloc_pinned(tot + 1);
	f->pinned = f->b.seq != 0;
	if (!f->pinned) f->b.seq = (uint8_t*)malloc(tot + 1);
	if (any_qual) f->b.qual = f->pinned ? (uint8_t*)bfcg_host_alloc_pinned(tot + 1) : (uint8_t*)malloc(tot + 1);
	if (f->b.seq == 0 || (any_qual && f->b.qual == 0)) return -1;
	f->seqs = seqs, f->to_reads = 0;
	kt_for(n_threads, copy_worker, f, n);
	return 0;
}

void flat_to_reads(const flat_batch_t *f, bseq1_t *seqs, int n_threads)
{
	flat_batch_t g = *f;
	g.seqs = seqs, g.to_reads = 1;
	kt_for(n_threads, copy_worker, &g, (long)f->b.n_reads);
}

void flat_free(flat_batch_t *f)
{
	if (f->pinned) { bfcg_host_free_pinned(f->b.seq); bfcg_host_free_pinned(f->b.qual); }
	else { free(f->b.seq); free(f->b.qual); }
	free(f->off);
	memset(f, 0, sizeof(*f));
}
