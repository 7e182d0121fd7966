/* correct_host.c -- bfc_correct(): the correction / trimming driver (interface:
 * reference bfc.h:40; shape: reference correct.c:573-655).  Three-step kt_pipeline as
 * in the reference: read | correct on the GPU (bfcg_correct_batch / bfcg_trim_batch
 * replace kt_for(worker_ec)) | print.  The printer reproduces correct.c:591-611 byte
 * for byte, including the ec:Z: tag. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bfc.h"
#include "bfc_b200.h"
#include "fqblock.h"

#define N_FLAT 3 /* a batch's flat buffers live from step 1 (GPU) to the end of step 2 (write) */

typedef struct {
	const bfc_opt_t *opt;
	fq_reader_t *ks;
	const bfc_bf_t *bf;
	const bfc_ch_t *ch;
	int mode;
	long n_batches;
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

static size_t batch_text_bytes(const bfc_opt_t *opt) { return (size_t)opt->chunk_size * 10; }

static void ec_step1(ec_shared_t *es, ec_step_t *data)
{
	const bfc_opt_t *opt = es->opt;
	const size_t n = (size_t)data->blk.n;
	int rc;
	data->flat = &es->flat[es->n_batches++ % N_FLAT];
	if (fq_flat_fill(data->flat, &data->blk, opt->n_threads) < 0) {
		fprintf(stderr, "[E::%s] out of host memory\n", "bfc_correct");
		exit(1);
	}
	if (!opt->filter_mode) {
		data->aux = (uint32_t*)malloc((n ? n : 1) * 2 * sizeof(uint32_t));
		rc = bfcg_correct_batch(opt, es->ch, es->mode, &data->flat->b, data->aux, &es->stats);
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
		const int ok = fq_next(es->ks, batch_text_bytes(es->opt), keep_comment, &ret->blk);
		fprintf(stderr, "[M::%s] read %ld sequences\n", "bfc_ec_cb", (long)ret->blk.n);
		if (ok) return ret;
		free(ret);
	} else if (step == 1) {
		ec_step_t *data = (ec_step_t*)_data;
		ec_step1(es, data);
		fprintf(stderr, "[M::%s @%.1f*%.1f%%] processed %ld sequences\n", "bfc_ec_cb", realtime() - bfc_real_time,
				100. * cputime() / (realtime() - bfc_real_time + 1e-6), (long)data->blk.n);
		return data;
	} else if (step == 2) { /* correct.c:591-616 */
		ec_step_t *data = (ec_step_t*)_data;
		fq_out_t o;
		memset(&o, 0, sizeof(o));
		o.filter_mode = es->opt->filter_mode, o.discard = es->opt->discard, o.no_qual = es->opt->no_qual;
		o.aux = data->aux, o.keep = data->keep, o.tstart = data->ts, o.tend = data->te;
		if (fq_write(stdout, &data->blk, data->flat, &o, es->opt->n_threads) < 0) {
			fprintf(stderr, "[E::%s] writing the output failed\n", "bfc_correct");
			exit(1);
		}
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
	if (opt->refine_ec) {
		fprintf(stderr, "[E::%s] refine mode (-R) is not part of the GPU path yet\n", __func__);
		exit(1);
	}
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
	es.ks = fq_open(fn, opt->n_threads);
	if (es.ks == 0) {
		fprintf(stderr, "[E::%s] cannot open '%s'\n", __func__, fn);
		exit(1);
	}
	kt_pipeline(opt->no_mt_io ? 1 : 3, ec_cb, &es, 3);
	fq_close(es.ks);
	{ int i; for (i = 0; i < N_FLAT; ++i) fq_flat_free(&es.flat[i]); }
	if (bfc_verbose >= 3 && !opt->filter_mode)
		fprintf(stderr, "[M::%s] table lookups: %llu; reads re-run with a larger search stack: %llu\n", __func__,
				(unsigned long long)es.stats.n_lookups, (unsigned long long)es.stats.n_redo);
}
