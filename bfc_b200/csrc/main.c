/* main.c -- the `bfc` command line (reference bfc.c:55-158): same option letters,
 * same defaults and -s rule, same phase order, same stderr summary lines. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "bfc.h"

#define BFC_VERSION "r181-b200"

static void usage(FILE *fp, const bfc_opt_t *o)
{
	fprintf(fp, "Usage: bfc [options] <to-count.fq> [to-correct.fq]\n");
	fprintf(fp, "Options:\n");
	fprintf(fp, "  -s FLOAT     approx genome size (k/m/g allowed; change -k and -b) [unset]\n");
	fprintf(fp, "  -k INT       k-mer length [%d]\n", o->k);
	fprintf(fp, "  -t INT       number of threads [%d]\n", o->n_threads);
	fprintf(fp, "  -b INT       set Bloom filter size to pow(2,INT) bits [%d]\n", o->bf_shift);
	fprintf(fp, "  -H INT       use INT hash functions for Bloom filter [%d]\n", o->n_hashes);
	fprintf(fp, "  -d FILE      dump hash table to FILE [null]\n");
	fprintf(fp, "  -E           skip error correction\n");
	fprintf(fp, "  -R           refine bfc-corrected reads\n");
	fprintf(fp, "  -r FILE      restore hash table from FILE [null]\n");
	fprintf(fp, "  -w INT       no more than %d ec or 2 highQ ec in INT-bp window [%d]\n", BFC_EC_HIST, o->win_multi_ec);
	fprintf(fp, "  -c INT       min k-mer coverage [%d]\n", o->min_cov);
	fprintf(fp, "  -Q           force FASTA output\n");
	fprintf(fp, "  -1           drop reads containing unique k-mers\n");
	fprintf(fp, "  -v           show version number\n");
	fprintf(fp, "  -h           show command line help\n");
}

static double parse_size(const char *arg)
{
	char *p;
	double x = strtod(arg, &p);
	if (*p == 'G' || *p == 'g') x *= 1e9;
	else if (*p == 'M' || *p == 'm') x *= 1e6;
	else if (*p == 'K' || *p == 'k') x *= 1e3;
	return x;
}

int main(int argc, char *argv[])
{
	bfc_opt_t opt;
	bfc_bf_t *bf = 0;
	bfc_ch_t *ch = 0;
	int i, c, no_ec = 0;
	const char *in_hash = 0, *out_hash = 0, *next_fn;

	bfc_real_time = realtime();
	bfc_opt_init(&opt);
	while ((c = getopt(argc, argv, "hvV:Ed:k:s:b:L:t:C:H:q:Jr:c:w:D1QR")) >= 0) {
		switch (c) {
		case 'd': out_hash = optarg; break;
		case 'r': in_hash = optarg; break;
		case 'q': opt.q = atoi(optarg); break;
		case 'b': opt.bf_shift = atoi(optarg); break;
		case 't': opt.n_threads = atoi(optarg); break;
		case 'H': opt.n_hashes = atoi(optarg); break;
		case 'c': opt.min_cov = atoi(optarg); break;
		case 'w': opt.win_multi_ec = atoi(optarg); break;
		case 'R': opt.refine_ec = 1; break;
		case 'D': opt.discard = 1; break;
		case '1': opt.filter_mode = 1; break;
		case 'Q': opt.no_qual = 1; break;
		case 'J': opt.no_mt_io = 1; break;
		case 'E': no_ec = 1; break;
		case 'V': bfc_verbose = atoi(optarg); break;
		case 'k':
			opt.k = atoi(optarg);
			fprintf(stderr, "[M::%s] set k to %d\n", __func__, opt.k);
			break;
		case 'h': usage(stdout, &opt); return 0;
		case 'v': printf("%s\n", BFC_VERSION); return 0;
		case 's':
			bfc_opt_by_size(&opt, (long)parse_size(optarg) + 1);
			fprintf(stderr, "[M::%s] applied `-k %d -b %d'\n", __func__, opt.k, opt.bf_shift);
			break;
		case 'L': opt.chunk_size = (int)((long)parse_size(optarg) + 1); break;
		default: break; /* -C is accepted and ignored, as in the reference */
		}
	}
	if (optind == argc) {
		usage(stderr, &opt);
		return 1;
	}
	if (opt.k < 1 || opt.k > BFC_MAX_KMER) {
		fprintf(stderr, "[E::%s] -k must be in 1..%d\n", __func__, BFC_MAX_KMER);
		return 1;
	}

	if (opt.filter_mode) bf = (bfc_bf_t*)bfc_count(argv[optind], &opt);
	else if (!in_hash) ch = (bfc_ch_t*)bfc_count(argv[optind], &opt);
	else {
		ch = bfc_ch_restore(in_hash);
		if (ch == 0) {
			fprintf(stderr, "[E::%s] failed to restore the hash table from '%s'\n", __func__, in_hash);
			return 1;
		}
		if (opt.k != bfc_ch_get_k(ch)) {
			opt.k = bfc_ch_get_k(ch);
			if (bfc_verbose >= 2)
				fprintf(stderr, "[W::%s] hash table was constructed with a different k; set k to %d\n", __func__, opt.k);
		}
	}

	next_fn = optind + 1 < argc ? argv[optind + 1] : argv[optind];
	if (ch) {
		if (out_hash) bfc_ch_dump(ch, out_hash);
		if (!no_ec) bfc_correct(next_fn, &opt, ch);
	} else if (bf) bfc_correct(next_fn, &opt, bf);

	fprintf(stderr, "[M::%s] Version: %s\n", __func__, BFC_VERSION);
	fprintf(stderr, "[M::%s] CMD:", __func__);
	for (i = 0; i < argc; ++i) fprintf(stderr, " %s", argv[i]);
	fprintf(stderr, "\n[M::%s] Real time: %.3f sec; CPU: %.3f sec\n", __func__, realtime() - bfc_real_time, cputime());
	/* The process ends here without tearing anything down: releasing the table, the pinned buffers and the CUDA context
	 * one by one took 0.3-1.5 s on the B200 box, and the kernel and the driver reclaim all of it at exit anyway. */
	if (fflush(stdout) != 0 || ferror(stdout)) {
		fprintf(stderr, "[E::%s] writing the output failed\n", __func__);
		_exit(1);
	}
	_exit(0);
}
