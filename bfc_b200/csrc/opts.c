/* opts.c -- option defaults and the -s rule (reference bfc.c:17-53). */
#include <math.h>
#include <string.h>
#include "bfc.h"

void bfc_opt_init(bfc_opt_t *o)
{
	memset(o, 0, sizeof(*o));
	o->chunk_size = 100000000;
	o->n_threads = 1;
	o->q = 20, o->k = 33;
	o->l_pre = 20, o->bf_shift = 33, o->n_hashes = 4;
	o->min_frac = .9;
	o->min_cov = 3, o->win_multi_ec = 10, o->max_end_ext = 5;
	o->w_ec = 1, o->w_ec_high = 7, o->w_absent = 3, o->w_absent_high = 1;
	o->max_path_diff = 15, o->max_heap = 100;
}

/* k = odd(floor(log2(size) + 1)) <= 63, b = floor(log2(size) + 8) <= 37 */
void bfc_opt_by_size(bfc_opt_t *o, long size)
{
	const double bits = log(size) / log(2);
	o->k = (int)(bits + 1.);
	if (!(o->k & 1)) ++o->k;
	if (o->k > BFC_MAX_KMER) o->k = BFC_MAX_KMER;
	o->bf_shift = (int)(bits + 8.);
	if (o->bf_shift > BFC_MAX_BF_SHIFT) o->bf_shift = BFC_MAX_BF_SHIFT;
}
