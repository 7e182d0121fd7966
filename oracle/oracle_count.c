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

/* ---- the same cascade split in two, for the sharded (multi-GPU) protocol tests ----
 * A record is what travels between ranks: y[0] | is_high << 63 and y[1] (the two words of
 * bfc_kmer_hash, kmer.h:79-88); the 64-bit Bloom hash is rebuilt from them. */

/* k-mer records of a batch in stream order (count.c:72-89); returns their number; y0/y1 may be NULL to count */
uint64_t orc_enum_records(const orc_opt_t *o, const orc_batch_t *b, uint64_t *y0, uint64_t *y1)
{
	int64_t r;
	uint64_t n = 0, m = ORC_MASK(o->k);
	for (r = 0; r < b->n_reads; ++r) {
		int i, run = 0, k = o->k, len = (int)(b->off[r + 1] - b->off[r] - 1);
		const uint8_t *seq = b->seq + b->off[r];
		const uint8_t *qual = b->qual && (len == 0 || b->qual[b->off[r]] != 0) ? b->qual + b->off[r] : 0;
		uint64_t x[4] = {0, 0, 0, 0}, qmer = 0;
		for (i = 0; i < len; ++i) {
			int c = orc_nt6[seq[i]] - 1;
			if (c >= 4) { run = 0, qmer = 0; memset(x, 0, sizeof(x)); continue; }
			okm_append(k, x, c);
			qmer = (qmer << 1 | (uint64_t)(qual == 0 || (int)qual[i] - 33 >= o->q)) & m;
			if (++run >= k) {
				uint64_t y[2];
				okm_hash(k, x, y);
				if (y0) y0[n] = y[0] | (uint64_t)(qmer == m) << 63, y1[n] = y[1];
				++n;
			}
		}
	}
	return n;
}

/* inverse of the last two lines of bfc_kmer_hash (kmer.h:85-86) */
uint64_t orc_hash_from_y(int k, uint64_t y0, uint64_t y1)
{
	uint64_t m = ORC_MASK(k), h0 = (y0 - y1) & m;
	return ((h0 ^ y1) << k) | y0;
}

/* the Bloom -> table cascade (count.c:54-70) over records, in the order given */
void orc_count_records(const orc_opt_t *o, orc_bf_t *bf, orc_bf_t *bf_high, orc_ch_t *ch, uint64_t n,
                       const uint64_t *y0, const uint64_t *y1, uint64_t stats[2])
{
	uint64_t i;
	for (i = 0; i < n; ++i) {
		uint64_t y[2], hash;
		y[0] = y0[i] & ~(1ULL << 63), y[1] = y1[i];
		hash = orc_hash_from_y(o->k, y[0], y[1]);
		++stats[0];
		if (orc_bf_insert(bf, hash) == o->n_hashes) {
			++stats[1];
			if (ch) orc_ch_insert(ch, y, (int)(y0[i] >> 63));
			else if (bf_high) orc_bf_insert(bf_high, hash);
		}
	}
}
