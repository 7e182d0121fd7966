/* oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference's count + correct (+ `-1` trim) path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load liboracle.so; the product library never does.
 *
 * Parity of this restatement is PINNED against the unmodified reference compiled
 * into oracle/_ref/ (see oracle/Makefile, tests/test_oracle_golden.py) and against
 * the golden fixtures under tests/golden/ that were produced by that reference
 * binary (tools/make_golden.py).  The reference ships no tests of its own.
 *
 * Each function cites the reference file:line it follows.
 */
#ifndef ORACLE_H
#define ORACLE_H

#include <stdint.h>
#include <stddef.h>

/* same field order as bfc_opt_t, reference bfc.h:15-33 */
typedef struct {
	int chunk_size;
	int n_threads, no_mt_io;
	int q, k;
	int filter_mode, refine_ec, no_qual;
	float min_frac;
	int l_pre, bf_shift, n_hashes;
	int discard;
	int max_end_ext;
	int win_multi_ec;
	int min_cov;
	int w_ec, w_ec_high, w_absent, w_absent_high;
	int max_path_diff, max_heap;
} orc_opt_t;

typedef struct { uint64_t x[4]; } orc_kmer_t;

/* ---- k-mer math (oracle_kmer.h) ---- */
void     orc_kmer_append(int k, uint64_t x[4], int c);
void     orc_kmer_change(int k, uint64_t x[4], int d, int c);
uint64_t orc_hash_64(uint64_t key, uint64_t mask);
uint64_t orc_kmer_hash(int k, const uint64_t x[4], uint64_t h[2]);

/* ---- blocked Bloom filter (oracle_tab.c; reference bbf.c) ---- */
typedef struct { int n_shift, n_hashes; uint8_t *b; } orc_bf_t;
orc_bf_t *orc_bf_new(int n_shift, int n_hashes);
void      orc_bf_free(orc_bf_t *b);
int       orc_bf_insert(orc_bf_t *b, uint64_t hash);
int       orc_bf_get(const orc_bf_t *b, uint64_t hash);

/* ---- counting table (oracle_tab.c; reference htab.c) ---- */
typedef struct orc_ch_s orc_ch_t;
orc_ch_t *orc_ch_new(int k, int l_pre);
void      orc_ch_free(orc_ch_t *ch);
int       orc_ch_k(const orc_ch_t *ch);
int       orc_ch_lpre(const orc_ch_t *ch);
void      orc_ch_subkey(const orc_ch_t *ch, const uint64_t y[2], uint32_t *sub, uint64_t *key);
int       orc_ch_insert(orc_ch_t *ch, const uint64_t y[2], int is_high);
int       orc_ch_get(const orc_ch_t *ch, const uint64_t y[2]);
int       orc_ch_kmer_occ(const orc_ch_t *ch, const orc_kmer_t *z);
uint64_t  orc_ch_count(const orc_ch_t *ch);
int       orc_ch_hist(const orc_ch_t *ch, uint64_t cnt[256], uint64_t high[64]);
/* entries as (sub, key50<<14|val14), sorted by (sub, key); returns n; arrays may be NULL to query n */
uint64_t  orc_ch_export(const orc_ch_t *ch, uint32_t *sub, uint64_t *key);
void      orc_ch_put_raw(orc_ch_t *ch, uint32_t sub, uint64_t key); /* restore path */

/* ---- host batch: reads are NUL-terminated strings back to back ----
 * read i = seq[off[i] .. off[i+1]-2]; qual == NULL or qual[off[i]] == 0 => no quality */
typedef struct {
	int64_t n_reads;
	const uint64_t *off;
	const uint8_t *seq, *qual;
} orc_batch_t;

/* ---- count (oracle_count.c; reference count.c:54-89) ----
 * stats[0] += k-mer occurrences, stats[1] += occurrences that passed the first Bloom */
void orc_count_batch(const orc_opt_t *opt, orc_bf_t *bf, orc_bf_t *bf_high, orc_ch_t *ch,
                     const orc_batch_t *batch, uint64_t stats[2]);

/* the same cascade in two halves, for the sharded protocol tests: records = (y[0] | is_high << 63, y[1]) */
uint64_t orc_enum_records(const orc_opt_t *opt, const orc_batch_t *batch, uint64_t *y0, uint64_t *y1);
uint64_t orc_hash_from_y(int k, uint64_t y0, uint64_t y1);
void orc_count_records(const orc_opt_t *opt, orc_bf_t *bf, orc_bf_t *bf_high, orc_ch_t *ch, uint64_t n,
                       const uint64_t *y0, const uint64_t *y1, uint64_t stats[2]);

/* ---- correct (oracle_correct.c; reference correct.c) ---- */
typedef struct orc_ecbuf_s orc_ecbuf_t;
orc_ecbuf_t *orc_ecbuf_new(const orc_opt_t *opt, const orc_ch_t *ch, int mode);
void         orc_ecbuf_free(orc_ecbuf_t *e);
/* edits seq/qual in place exactly as bfc_ec1 does; returns aux | (uint64_t)aux2 << 32
 * packed like worker_ec (correct.c:552-553) */
uint64_t orc_ec1(orc_ecbuf_t *e, char *seq, char *qual, const uint32_t *ori /* refine mode: earlier aux, aux2 */);
/* counters accumulated over orc_ec1 calls: [0] table lookups, [1] heap pops, [2] max stack */
const uint64_t *orc_ecbuf_counters(const orc_ecbuf_t *e);
/* whole batch, sequentially; seq/qual are edited in place; aux[2*i], aux[2*i+1] per read */
void orc_correct_batch(const orc_opt_t *opt, const orc_ch_t *ch, int mode, int64_t n_reads,
                       const uint64_t *off, uint8_t *seq, uint8_t *qual, uint32_t *aux, uint64_t counters[3]);

/* ---- trim (reference correct.c:478-497, 554-569) ----
 * returns streak<<32 | start as max_streak does */
uint64_t orc_max_streak(int k, const orc_bf_t *bf, const char *seq, int l_seq);
/* keep[i] = 1 and [tstart[i], tend[i]) = kept range, or keep[i] = 0 */
void orc_trim_batch(const orc_opt_t *opt, const orc_bf_t *bf, int64_t n_reads, const uint64_t *off,
                    const uint8_t *seq, uint8_t *keep, int32_t *tstart, int32_t *tend);

extern const unsigned char orc_nt6[256];

#endif
