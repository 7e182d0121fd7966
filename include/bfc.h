/* bfc.h -- phase drivers and options of the count + correct path.
 *
 * Drop-in for the reference's bfc.h (bfc.h:8-45): same limits, same option struct
 * (field for field), same globals and the same entry points.  `bfc_count` returns a
 * bfc_ch_t* (or, with opt->filter_mode, the bfc_bf_t* of high-occurrence k-mers) that
 * the caller owns; `bfc_correct` borrows it and writes corrected / trimmed reads to
 * stdout byte for byte as the reference does with `-t1`.
 */
#ifndef BFC_B200_BFC_H
#define BFC_B200_BFC_H

#include "bbf.h"
#include "htab.h"
#include "bseq.h"

#define BFC_MAX_KMER     63
#define BFC_MAX_BF_SHIFT 37

#define BFC_MAX_PATHS 4
#define BFC_EC_HIST 5
#define BFC_EC_HIST_HIGH 2

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	int chunk_size;            /* bases per pipeline batch (-L) */
	int n_threads, no_mt_io;   /* -t (host-side helper threads), -J */
	int q, k;                  /* -q quality threshold, -k k-mer length */

	int filter_mode, refine_ec, no_qual; /* -1, -R, -Q */
	float min_frac;            /* trim mode: keep a read if (streak+k)/len exceeds this */

	int l_pre, bf_shift, n_hashes; /* table prefix bits, -b, -H */

	int discard;               /* -D */
	int max_end_ext;
	int win_multi_ec;          /* -w */
	int min_cov;               /* -c */

	int w_ec, w_ec_high, w_absent, w_absent_high; /* penalty weights */
	int max_path_diff, max_heap;
} bfc_opt_t;

extern int bfc_verbose;
extern double bfc_real_time;
extern bfc_kmer_t bfc_kmer_null;

void bfc_opt_init(bfc_opt_t *opt);                 /* reference bfc.c:17-40 */
void bfc_opt_by_size(bfc_opt_t *opt, long size);   /* reference bfc.c:42-53 */

void *bfc_count(const char *fn, const bfc_opt_t *opt);
void bfc_correct(const char *fn, const bfc_opt_t *opt, const void *ptr);

void kt_for(int n_threads, void (*func)(void*,long,int), void *data, long n);
void kt_pipeline(int n_threads, void *(*func)(void*, int, void*), void *shared_data, int n_steps);
double cputime(void);
double realtime(void);

#ifdef __cplusplus
}
#endif

#endif
