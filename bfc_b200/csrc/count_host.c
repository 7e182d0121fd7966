/* count_host.c -- bfc_count(): the count-phase driver (interface: reference bfc.h:39;
 * shape: reference count.c:91-157).  The host keeps the reference's two-step
 * kt_pipeline (read a batch | count it); what the reference does inside
 * kt_for(worker_count) is one bfcg_count_batch() call on the GPU.  Progress lines on
 * stderr keep the reference's wording (count.c:98, 110-115). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bfc.h"
#include "bfc_b200.h"
#include "flat.h"

typedef struct {
	const bfc_opt_t *opt;
	bseq_file_t *ks;
	bfc_bf_t *bf, *bf_high;
	bfc_ch_t *ch;
	bfcg_stats_t stats;
} cnt_shared_t;

typedef struct {
	int n_seqs;
	bseq1_t *seqs;
} cnt_step_t;

static void free_reads(bseq1_t *seqs, int n)
{
	int i;
	for (i = 0; i < n; ++i) { free(seqs[i].seq); free(seqs[i].qual); free(seqs[i].comment); free(seqs[i].name); }
	free(seqs);
}

static void *count_cb(void *shared, int step, void *_data)
{
	cnt_shared_t *cs = (cnt_shared_t*)shared;
	if (step == 0) {
		cnt_step_t *ret = (cnt_step_t*)calloc(1, sizeof(cnt_step_t));
		ret->seqs = bseq_read(cs->ks, cs->opt->chunk_size, 0, &ret->n_seqs);
		fprintf(stderr, "[M::%s] read %d sequences\n", "bfc_count_cb", ret->n_seqs);
		if (ret->seqs) return ret;
		free(ret);
	} else if (step == 1) {
		cnt_step_t *data = (cnt_step_t*)_data;
		flat_batch_t f;
		double rt, eff;
		if (flat_from_reads(&f, data->seqs, data->n_seqs, cs->opt->n_threads) < 0 ||
			bfcg_count_batch(cs->opt, cs->bf, cs->bf_high, cs->ch, &f.b, &cs->stats) != BFCG_OK) {
			fprintf(stderr, "[E::%s] GPU count failed: %s\n", "bfc_count", bfcg_last_error());
			exit(1);
		}
		flat_free(&f);
		rt = realtime() - bfc_real_time;
		eff = 100. * cputime() / (rt + 1e-6);
		if (cs->ch)
			fprintf(stderr, "[M::%s @%.1f*%.1f%%] processed %d sequences; # distinct k-mers: %ld\n",
					"bfc_count_cb", rt, eff, data->n_seqs, (long)bfc_ch_count(cs->ch));
		else
			fprintf(stderr, "[M::%s @%.1f*%.1f%%] processed %d sequences\n", "bfc_count_cb", rt, eff, data->n_seqs);
		free_reads(data->seqs, data->n_seqs);
		free(data);
	}
	return 0;
}

void *bfc_count(const char *fn, const bfc_opt_t *opt)
{
	cnt_shared_t cs;
	void *ret;
	memset(&cs, 0, sizeof(cs));
	cs.opt = opt;
	cs.bf = bfc_bf_init(opt->bf_shift, opt->n_hashes);
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
	cs.ks = bseq_open(fn);
	if (cs.ks == 0) {
		fprintf(stderr, "[E::%s] cannot open '%s'\n", __func__, fn);
		exit(1);
	}
	kt_pipeline(opt->no_mt_io ? 1 : 2, count_cb, &cs, 2);
	bseq_close(cs.ks);
	if (bfc_verbose >= 3)
		fprintf(stderr, "[M::%s] k-mer occurrences: %llu; passed the first filter: %llu; replayed in order: %llu\n", __func__,
				(unsigned long long)cs.stats.n_kmers, (unsigned long long)cs.stats.n_pass, (unsigned long long)cs.stats.n_conflict);
	bfc_bf_destroy(cs.bf); /* the first filter never leaves this function (reference count.c:155) */
	return ret;
}
