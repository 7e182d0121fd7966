/* TEST INFRASTRUCTURE ONLY -- oracle restatement of the reference count phase. */
#include <string.h>
#include "oracle_kmer.h"

/* reference bseq.c:9-26: A/a=1 C/c=2 G/g=3 T/t=4, everything else 5 */
const unsigned char orc_nt6[256] = {
	5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5, 5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,
	5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5, 5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,
	5,1,5,2,5,5,5,3,5,5,5,5,5,5,5,5, 5,5,5,5,4,5,5,5,5,5,5,5,5,5,5,5,
	5,1,5,2,5,5,5,3,5,5,5,5,5,5,5,5, 5,5,5,5,4,5,5,5,5,5,5,5,5,5,5,5,
	5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5, 5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,
	5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5, 5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,
	5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5, 5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,
	5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5, 5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5
};

/* One read, reference count.c:72-89 (worker_count) with the cascade of
 * count.c:54-70 (bfc_kmer_insert) inlined; the retry buffer of count.c:42-52 is a
 * lock-contention device with no effect on a sequential run. */
static void count_read(const orc_opt_t *o, orc_bf_t *bf, orc_bf_t *bf_high, orc_ch_t *ch,
                       const uint8_t *seq, const uint8_t *qual, int len, uint64_t stats[2])
{
	int i, run = 0, k = o->k;
	uint64_t x[4] = {0, 0, 0, 0}, qmer = 0, m = ORC_MASK(k);
	for (i = 0; i < len; ++i) {
		int c = orc_nt6[seq[i]] - 1;
		if (c >= 4) { run = 0, qmer = 0; memset(x, 0, sizeof(x)); continue; }
		okm_append(k, x, c);
		qmer = (qmer << 1 | (uint64_t)(qual == 0 || (int)qual[i] - 33 >= o->q)) & m;
		if (++run >= k) {
			uint64_t y[2], hash = okm_hash(k, x, y);
			++stats[0];
			if (orc_bf_insert(bf, hash) == o->n_hashes) {
				++stats[1];
				if (ch) orc_ch_insert(ch, y, qmer == m);
				else if (bf_high) orc_bf_insert(bf_high, hash);
			}
		}
	}
}

void orc_count_batch(const orc_opt_t *opt, orc_bf_t *bf, orc_bf_t *bf_high, orc_ch_t *ch,
                     const orc_batch_t *b, uint64_t stats[2])
{
	int64_t i;
	for (i = 0; i < b->n_reads; ++i) { /* read order == `-t1` order (kthread.c:39) */
		int len = (int)(b->off[i + 1] - b->off[i] - 1);
		const uint8_t *q = b->qual && (len == 0 || b->qual[b->off[i]] != 0) ? b->qual + b->off[i] : 0;
		count_read(opt, bf, bf_high, ch, b->seq + b->off[i], q, len, stats);
	}
}
