/* count_host.c -- bfc_count(): the count-phase driver (interface: reference bfc.h:39;
 * shape: reference count.c:91-157).  The reference's two-step kt_pipeline (read a batch | count it) becomes
 * three steps -- read and split a block of text | pack it into a pinned flat batch | count it on the GPU -- so that
 * the host's packing of batch i+1 runs beside the GPU's work on batch i; batches are blocks of input text split by
 * all -t threads (fqblock.c), and what the reference does inside kt_for(worker_count) is one
 * bfcg_count_batch() call on the GPU.  Progress lines on
 * stderr keep the reference's wording (count.c:98, 110-115). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "bfc.h"
#include "bfc_b200.h"
#include "fqblock.h"

#define N_FLAT 3 /* = the batches the pipeline keeps in flight */

typedef struct {
	const bfc_opt_t *opt;
	fq_reader_t *ks;
	bfc_bf_t *bf, *bf_high;
	bfc_ch_t *ch;
	fq_flat_t flat[N_FLAT]; /* pinned, reused round robin: a batch holds one from step 1 (pack) to the end of step 2 (GPU) */
	long n_batches;
	int keep_comment;       /* what bfc_correct will ask the reader for: the blocks are kept for it */
	bfcg_stats_t stats;
} cnt_shared_t;

typedef struct {
	fq_block_t blk;
	fq_flat_t *flat;
} cnt_step_t;

/* text per batch: the reference batches by bases (-L, count.c:103); here a batch is a block of input text -- 250 MB
 * at the default -L: the GPU's per-batch costs (one sweep of the filter and of the table, ~10 ms) stay far below what
 * the host needs to read and pack it, and the pinned staging buffers and the recycled text buffers stay small enough
 * that filling the pool with first-touch pages is over after a fraction of a second */
static size_t batch_text_bytes(const bfc_opt_t *opt) { return (size_t)opt->chunk_size * 5 / 2; }

#define STAMP(what) do { if (bfc_verbose >= 4) fprintf(stderr, "[D::%s @%.3f] %s\n", __func__, realtime() - bfc_real_time, what); } while (0)

static void *count_cb(void *shared, int step, void *_data)
{
	cnt_shared_t *cs = (cnt_shared_t*)shared;
	if (step == 0) {
		cnt_step_t *d = (cnt_step_t*)calloc(1, sizeof(cnt_step_t));
		int ok;
		STAMP("read begins");
		ok = fq_next(cs->ks, batch_text_bytes(cs->opt), cs->keep_comment, &d->blk);
		STAMP("block read and split");
		if (ok < 0) { fprintf(stderr, "[E::%s] out of host memory while reading\n", "bfc_count"); exit(1); }
		fprintf(stderr, "[M::%s] read %ld sequences\n", "bfc_count_cb", (long)d->blk.n);
		if (ok) return d;
		free(d);
	} else if (step == 1) {
		cnt_step_t *d = (cnt_step_t*)_data;
		d->flat = &cs->flat[cs->n_batches++ % N_FLAT];
		STAMP("pack begins");
		if (fq_flat_fill(d->flat, &d->blk, 0, cs->opt->n_threads) < 0) {
			fprintf(stderr, "[E::%s] out of host memory while packing\n", "bfc_count");
			exit(1);
		}
		STAMP("batch packed");
		return d;
	} else if (step == 2) {
		cnt_step_t *d = (cnt_step_t*)_data;
		double rt, eff;
		STAMP("count begins");
		if (bfcg_count_batch(cs->opt, cs->bf, cs->bf_high, cs->ch, &d->flat->b, &cs->stats) != BFCG_OK) {
			fprintf(stderr, "[E::%s] GPU count failed: %s\n", "bfc_count", bfcg_last_error());
			exit(1);
		}
		STAMP("batch counted");
		rt = realtime() - bfc_real_time;
		eff = 100. * cputime() / (rt + 1e-6);
		if (cs->ch)
			fprintf(stderr, "[M::%s @%.1f*%.1f%%] processed %ld sequences; # distinct k-mers: %ld\n",
					"bfc_count_cb", rt, eff, (long)d->blk.n, (long)bfc_ch_count(cs->ch));
		else
			fprintf(stderr, "[M::%s @%.1f*%.1f%%] processed %ld sequences\n", "bfc_count_cb", rt, eff, (long)d->blk.n);
		if (!fq_keep_add(&d->blk)) fq_block_free(&d->blk);
		free(d);
	}
	return 0;
}

void *bfc_count(const char *fn, const bfc_opt_t *opt)
{
	cnt_shared_t cs;
	void *ret;
	memset(&cs, 0, sizeof(cs));
	cs.opt = opt;
	STAMP("start");
	cs.bf = bfc_bf_init(opt->bf_shift, opt->n_hashes);
	STAMP("filter allocated");
	if (cs.bf == 0) {
		fprintf(stderr, "[E::%s] cannot create the Bloom filter (-b %d): %s\n", __func__, opt->bf_shift, bfcg_last_error());
		exit(1);
	}
	if (!opt->filter_mode) ret = cs.ch = bfc_ch_init(opt->k, opt->l_pre);
	else ret = cs.bf_high = bfc_bf_init(opt->bf_shift, opt->n_hashes);
	if (ret == 0) {
		fprintf(stderr, "[E::%s] cannot create the k-mer table / second filter: %s\n", __func__, bfcg_last_error());
		exit(1);
	}
	STAMP("table allocated");
	cs.ks = fq_open(fn, opt->n_threads);
	if (cs.ks == 0) {
		fprintf(stderr, "[E::%s] cannot open '%s'\n", __func__, fn);
		exit(1);
	}
	cs.keep_comment = opt->filter_mode || opt->refine_ec;
	fq_keep_begin(fn, cs.keep_comment);
	kt_pipeline(opt->no_mt_io ? 1 : N_FLAT, count_cb, &cs, 3);
	fq_keep_end(1);
	fq_close(cs.ks);
	{ int i; for (i = 0; i < N_FLAT; ++i) fq_flat_free(&cs.flat[i]); }
	if (bfc_verbose >= 3)
		fprintf(stderr, "[M::%s] k-mer occurrences: %llu; passed the first filter: %llu; replayed in order: %llu\n", __func__,
				(unsigned long long)cs.stats.n_kmers, (unsigned long long)cs.stats.n_pass, (unsigned long long)cs.stats.n_conflict);
	if (bfc_verbose >= 3) { /* filter occupancy: how full -b left the first filter, and what that means for the table */
		double load = 0, blocks = 0, fp = 0;
		if (bfcg_bf_load(cs.bf, 1, &load, &blocks, &fp) == BFCG_OK) {
			fprintf(stderr, "[M::%s] first Bloom filter (-b %d): %.2f%% of its bits set, %.1f%% of its blocks touched; a new k-mer passes it with p = %.2e\n",
					__func__, opt->bf_shift, 100. * load, 100. * blocks, fp);
			if (fp > 0.05) {
				const uint64_t n_est = (uint64_t)(-(double)((uint64_t)1 << opt->bf_shift) * (504. / 512.) / opt->n_hashes * log(1. - (load < 0.999999 ? load : 0.999999)));
				fprintf(stderr, "[W::%s] the first filter is saturated (about %llu distinct k-mers went in): singletons leak into the %s; -b %d would keep p below 1%%\n",
						__func__, (unsigned long long)n_est, cs.ch ? "k-mer table" : "second filter", bfcg_bf_suggest_shift(n_est, opt->n_hashes, 0.01));
			}
		}
	}
	bfc_bf_destroy(cs.bf); /* the first filter never leaves this function (reference count.c:155) */
	return ret;
}
