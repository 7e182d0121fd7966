/* correct_host.c -- bfc_correct(): the correction / trimming driver (interface:
 * reference bfc.h:40; shape: reference correct.c:573-655).  Three-step kt_pipeline as
 * in the reference: read | correct on the GPU (bfcg_correct_batch / bfcg_trim_batch
 * replace kt_for(worker_ec)) | print.  The printer reproduces correct.c:591-611 byte
 * for byte, including the ec:Z: tag. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <assert.h>
#include "bfc.h"
#include "bfc_b200.h"
#include "flat.h"

typedef struct {
	const bfc_opt_t *opt;
	bseq_file_t *ks;
	const bfc_bf_t *bf;
	const bfc_ch_t *ch;
	int mode;
	bfcg_stats_t stats;
} ec_shared_t;

typedef struct {
	int n_seqs;
	bseq1_t *seqs;
} ec_step_t;

static void ec_step1(ec_shared_t *es, ec_step_t *data)
{
	const bfc_opt_t *opt = es->opt;
	flat_batch_t f;
	int i, n = data->n_seqs, rc;
	if (flat_from_reads(&f, data->seqs, n, opt->n_threads) < 0) {
		fprintf(stderr, "[E::%s] out of host memory\n", "bfc_correct");
		exit(1);
	}
	if (!opt->filter_mode) {
		uint32_t *aux = (uint32_t*)malloc((size_t)n * 2 * sizeof(uint32_t));
		rc = bfcg_correct_batch(opt, es->ch, es->mode, &f.b, aux, &es->stats);
		if (rc == BFCG_OK) {
			flat_to_reads(&f, data->seqs, opt->n_threads);
			for (i = 0; i < n; ++i) {
				bseq1_t *s = &data->seqs[i];
				if (s->comment) { free(s->comment); s->comment = 0; } /* correct.c:547-550 */
				s->aux = aux[2 * i], s->aux2 = aux[2 * i + 1];
			}
		}
		free(aux);
	} else {
		uint8_t *keep = (uint8_t*)malloc((size_t)n);
		int32_t *ts = (int32_t*)malloc((size_t)n * 4), *te = (int32_t*)malloc((size_t)n * 4);
		rc = bfcg_trim_batch(opt, es->bf, &f.b, keep, ts, te, &es->stats);
		if (rc == BFCG_OK) {
			for (i = 0; i < n; ++i) { /* correct.c:557-569 */
				bseq1_t *s = &data->seqs[i];
				if (keep[i]) {
					const int start = ts[i], end = te[i];
					assert(start >= 0 && end <= s->l_seq);
					memmove(s->seq, s->seq + start, (size_t)(end - start));
					s->l_seq = end - start;
					s->seq[s->l_seq] = 0;
					if (s->qual) {
						memmove(s->qual, s->qual + start, (size_t)s->l_seq);
						s->qual[s->l_seq] = 0;
					}
					s->aux = 0;
				} else s->aux = 1;
			}
		}
		free(keep); free(ts); free(te);
	}
	flat_free(&f);
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
		ret->seqs = bseq_read(es->ks, es->opt->chunk_size, keep_comment, &ret->n_seqs);
		fprintf(stderr, "[M::%s] read %d sequences\n", "bfc_ec_cb", ret->n_seqs);
		if (ret->seqs) return ret;
		free(ret);
	} else if (step == 1) {
		ec_step_t *data = (ec_step_t*)_data;
		ec_step1(es, data);
		fprintf(stderr, "[M::%s @%.1f*%.1f%%] processed %d sequences\n", "bfc_ec_cb", realtime() - bfc_real_time,
				100. * cputime() / (realtime() - bfc_real_time + 1e-6), data->n_seqs);
		return data;
	} else if (step == 2) {
		ec_step_t *data = (ec_step_t*)_data;
		const bfc_opt_t *opt = es->opt;
		int i;
		for (i = 0; i < data->n_seqs; ++i) {
			bseq1_t *s = &data->seqs[i];
			const int is_fq = (s->qual && !opt->no_qual);
			int skip = 0;
			if (!opt->filter_mode) {
				if (opt->discard && (s->aux & 7)) skip = 1;
				else {
					printf("%c%s", is_fq ? '@' : '>', s->name);
					if (!s->comment) {
						printf("\tec:Z:%d", s->aux & 7);
						if ((s->aux & 7) == 0)
							printf("_%d:%d_%d_%d:%d_%d", s->aux2 >> 10, s->aux2 & 0xff, s->aux >> 3 & 1,
								   s->aux >> 18 & 0x3fff, s->aux >> 4 & 0x3fff, s->aux2 >> 8 & 3);
					} else printf("\t%s", s->comment);
				}
			} else {
				if (s->aux) skip = 1;
				else {
					printf("%c%s", is_fq ? '@' : '>', s->name);
					if (s->comment) printf("\t%s", s->comment);
				}
			}
			if (!skip) {
				putchar('\n'); puts(s->seq);
				if (is_fq) { puts("+"); puts(s->qual); }
			}
			free(s->seq); free(s->qual); free(s->comment); free(s->name);
		}
		free(data->seqs); free(data);
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
	es.ks = bseq_open(fn);
	if (es.ks == 0) {
		fprintf(stderr, "[E::%s] cannot open '%s'\n", __func__, fn);
		exit(1);
	}
	kt_pipeline(opt->no_mt_io ? 1 : 2, ec_cb, &es, 3);
	bseq_close(es.ks);
	if (bfc_verbose >= 3 && !opt->filter_mode)
		fprintf(stderr, "[M::%s] table lookups: %llu; reads re-run with a larger search stack: %llu\n", __func__,
				(unsigned long long)es.stats.n_lookups, (unsigned long long)es.stats.n_redo);
}
