/* correct_host.c -- bfc_correct(): the correction / trimming driver (interface:
 * reference bfc.h:40; shape: reference correct.c:573-655).  The reference's three-step
 * kt_pipeline with the packing as a step of its own: read | pack into a pinned flat batch | correct on the GPU
 * (bfcg_correct_batch / bfcg_trim_batch replace kt_for(worker_ec)) | print.  The printer reproduces correct.c:591-611 byte
 * for byte, including the ec:Z: tag. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bfc.h"
#include "bfc_b200.h"
#include "fqblock.h"

#define N_FLAT 4 /* = the batches in flight: a batch's flat buffers live from step 1 (pack) to the end of step 3 (write) */

typedef struct {
	const bfc_opt_t *opt;
	fq_reader_t *ks;
	const bfc_bf_t *bf;
	const bfc_ch_t *ch;
	int mode;
	long n_batches;
	long kept, next_kept;   /* blocks the count pass kept for this file (-1: none, read it again) */
	uint32_t ori[2];        /* -R: the stats of the latest tagged read (e->ori_st of the reference, correct.c:176, 543) */
	fq_flat_t flat[N_FLAT];
	bfcg_stats_t stats;
} ec_shared_t;

typedef struct {
	fq_block_t blk;
	fq_flat_t *flat;
	uint32_t *aux;          /* normal mode: 2 per read (correct.c:552-553) */
	uint8_t *keep;          /* filter mode */
	int32_t *ts, *te;
} ec_step_t;

static size_t batch_text_bytes(const bfc_opt_t *opt) { return (size_t)opt->chunk_size * 5 / 2; }

/* parse_stats (reference correct.c:517-531) over the text after "ec:Z:", packed as worker_ec packs the result
 * (correct.c:552-553): numbers separated by one character each; what is missing counts as 0 */
static void parse_ec_tag(const char *s, size_t l, uint32_t ori[2])
{
	long v[6] = {0, 0, 0, 0, 0, 0};
	size_t i = 0;
	int f;
	for (f = 0; f < 6 && i < l; ++f) {
		int neg = 0;
		if (f) ++i; /* the separator */
		if (i < l && s[i] == '-') neg = 1, ++i;
		while (i < l && s[i] >= '0' && s[i] <= '9') v[f] = v[f] * 10 + (s[i++] - '0');
		if (neg) v[f] = -v[f];
		if (f == 0 && v[0] != 0) break; /* the other fields are only read for ec_code 0 */
	}
	if (v[0] != 0) ori[0] = (uint32_t)v[0] & 7, ori[1] = 0;
	else {
		ori[0] = ((uint32_t)v[4] & 0x3fff) << 18 | ((uint32_t)v[5] & 0x3fff) << 4 | ((uint32_t)v[3] & 1) << 3;
		ori[1] = ((uint32_t)v[1] & 0x3fffff) << 10 | 1u << 8 | ((uint32_t)v[2] & 0xff);
	}
}

#define STAMP(what) do { if (bfc_verbose >= 4) fprintf(stderr, "[D::%s @%.3f] %s\n", __func__, realtime() - bfc_real_time, what); } while (0)

static void ec_pack(ec_shared_t *es, ec_step_t *data)
{
	const bfc_opt_t *opt = es->opt;
	const size_t n = (size_t)data->blk.n;
	uint8_t *skip = 0;
	data->flat = &es->flat[es->n_batches++ % N_FLAT];
	STAMP("pack begins");
	if (!opt->filter_mode) data->aux = (uint32_t*)malloc((n ? n : 1) * 2 * sizeof(uint32_t));
	if (opt->refine_ec && !opt->filter_mode) { /* worker_ec's refine branch (correct.c:542-550), in read order */
		const fq_block_t *b = &data->blk;
		size_t r, m = 0;
		skip = (uint8_t*)calloc(n ? n : 1, 1);
		for (r = 0; r < n; ++r) {
			if (b->com_off[r] != FQ_NONE && b->com_len[r] >= 5 && memcmp(b->buf + b->com_off[r], "ec:Z:", 5) == 0) {
				parse_ec_tag(b->buf + b->com_off[r] + 5, b->com_len[r] - 5, es->ori);
				skip[r] = (es->ori[0] & 7) == 0 && (es->ori[1] & 0xff) < 50; /* fine as it is */
			}
			if (!skip[r]) data->aux[2 * m] = es->ori[0], data->aux[2 * m + 1] = es->ori[1], ++m; /* in: the earlier stats */
		}
	}
	if (fq_flat_fill(data->flat, &data->blk, skip, opt->n_threads) < 0) {
		fprintf(stderr, "[E::%s] out of host memory\n", "bfc_correct");
		exit(1);
	}
	free(skip);
	STAMP("batch packed");
}

static void ec_gpu(ec_shared_t *es, ec_step_t *data)
{
	const bfc_opt_t *opt = es->opt;
	const size_t n = (size_t)data->blk.n;
	int rc = BFCG_OK;
	if (!opt->filter_mode) {
		if (data->flat->b.n_reads) rc = bfcg_correct_batch(opt, es->ch, es->mode, &data->flat->b, data->aux, &es->stats);
	} else {
		data->keep = (uint8_t*)malloc(n ? n : 1);
		data->ts = (int32_t*)malloc((n ? n : 1) * 4), data->te = (int32_t*)malloc((n ? n : 1) * 4);
		rc = bfcg_trim_batch(opt, es->bf, &data->flat->b, data->keep, data->ts, data->te, &es->stats);
	}
	if (rc != BFCG_OK) {
		fprintf(stderr, "[E::%s] GPU correction failed: %s\n", "bfc_correct", bfcg_last_error());
		exit(1);
	}
}

static void *ec_cb(void *shared, int step, void *_data)
{
	ec_shared_t *es = (ec_shared_t*)shared;
	if (step == 0) {
		ec_step_t *ret = (ec_step_t*)calloc(1, sizeof(ec_step_t));
		const int keep_comment = (es->opt->filter_mode || es->opt->refine_ec);
		int ok;
		STAMP("read begins");
		if (es->kept >= 0) ok = fq_keep_take(es->next_kept++, &ret->blk);
		else ok = fq_next(es->ks, batch_text_bytes(es->opt), keep_comment, &ret->blk);
		STAMP("block read and split");
		if (ok < 0) { fprintf(stderr, "[E::%s] out of host memory while reading\n", "bfc_correct"); exit(1); }
		fprintf(stderr, "[M::%s] read %ld sequences\n", "bfc_ec_cb", (long)ret->blk.n);
		if (ok) return ret;
		free(ret);
	} else if (step == 1) {
		ec_pack(es, (ec_step_t*)_data);
		return _data;
	} else if (step == 2) {
		ec_step_t *data = (ec_step_t*)_data;
		STAMP("batch taken");
		ec_gpu(es, data);
		STAMP("batch corrected");
		fprintf(stderr, "[M::%s @%.1f*%.1f%%] processed %ld sequences\n", "bfc_ec_cb", realtime() - bfc_real_time,
				100. * cputime() / (realtime() - bfc_real_time + 1e-6), (long)data->blk.n);
		return data;
	} else if (step == 3) { /* correct.c:591-616 */
		ec_step_t *data = (ec_step_t*)_data;
		fq_out_t o;
		memset(&o, 0, sizeof(o));
		STAMP("write begins");
		o.filter_mode = es->opt->filter_mode, o.discard = es->opt->discard, o.no_qual = es->opt->no_qual;
		o.refine = es->opt->refine_ec && !es->opt->filter_mode;
		o.aux = data->aux, o.keep = data->keep, o.tstart = data->ts, o.tend = data->te;
		if (fq_write(stdout, &data->blk, data->flat, &o, es->opt->n_threads) < 0) {
			fprintf(stderr, "[E::%s] writing the output failed\n", "bfc_correct");
			exit(1);
		}
		STAMP("batch written");
		fq_block_free(&data->blk);
		free(data->aux); free(data->keep); free(data->ts); free(data->te);
		free(data);
	}
	return 0;
}

void bfc_correct(const char *fn, const bfc_opt_t *opt, const void *ptr)
{
	ec_shared_t es;
	memset(&es, 0, sizeof(es));
	es.opt = opt;
	if (bfc_verbose >= 3)
		fprintf(stderr, "[M::%s @%.1f*%.1f%%] Starting...\n", __func__, realtime() - bfc_real_time,
				100. * cputime() / (realtime() - bfc_real_time + 1e-6));
	if (!opt->filter_mode) {
		uint64_t hist[256], hist_high[64];
		int i;
		es.ch = (const bfc_ch_t*)ptr;
		es.mode = bfc_ch_hist(es.ch, hist, hist_high);
		if (bfc_verbose >= 4)
			for (i = 0; i < 256; ++i) {
				if (i < 64) fprintf(stderr, "[M::%s] %3d : %llu : %llu\n", __func__, i, (unsigned long long)hist[i], (unsigned long long)hist_high[i]);
				else fprintf(stderr, "[M::%s] %3d : %llu\n", __func__, i, (unsigned long long)hist[i]);
			}
	} else es.bf = (const bfc_bf_t*)ptr;
	es.kept = fq_keep_match(fn, opt->filter_mode || opt->refine_ec);
	if (es.kept < 0) {
		fq_keep_drop();
		es.ks = fq_open(fn, opt->n_threads);
		if (es.ks == 0) {
			fprintf(stderr, "[E::%s] cannot open '%s'\n", __func__, fn);
			exit(1);
		}
	} else if (bfc_verbose >= 4) fprintf(stderr, "[M::%s] %ld blocks of '%s' are still in memory from the count pass\n", __func__, es.kept, fn);
	kt_pipeline(opt->no_mt_io ? 1 : N_FLAT, ec_cb, &es, 4);
	if (es.ks) fq_close(es.ks);
	fq_keep_drop();
	{ int i; for (i = 0; i < N_FLAT; ++i) fq_flat_free(&es.flat[i]); }
	if (bfc_verbose >= 3 && !opt->filter_mode)
		fprintf(stderr, "[M::%s] table lookups: %llu; reads re-run with a larger search stack: %llu\n", __func__,
				(unsigned long long)es.stats.n_lookups, (unsigned long long)es.stats.n_redo);
}
