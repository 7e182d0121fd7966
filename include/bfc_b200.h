/* bfc_b200.h -- batch-level C ABI of the B200 count / correct / trim engine.
 *
 * This is the boundary a host program (the `bfc` CLI in csrc/, the Python mirror
 * in bfc_b200/, or a maintainer's patch to the reference) binds:
 * plain pointers and sizes, no C++ or torch types.  It replaces what runs INSIDE the
 * reference's two parallel loops:
 *
 *   bfcg_count_batch    <- kt_for(..., worker_count, ...)   reference count.c:106 (+ :54-89, :116-117)
 *   bfcg_correct_batch  <- kt_for(..., worker_ec, ...)      reference correct.c:587 (+ :388-472, :533-553)
 *   bfcg_trim_batch     <- worker_ec, filter_mode branch    reference correct.c:554-569 (+ :478-497)
 *
 * Results are those of the reference run with `-t1` (strict read order), bit for bit.
 * There is no CPU fallback: every entry point returns BFCG_ERR_CUDA (and prints a
 * `[E::...]` line) when no CUDA device / kernel image is usable.
 *
 * Host batch layout ("reads back to back"):
 *   read i occupies seq[off[i] .. off[i+1]-2]; seq[off[i+1]-1] is a 0 terminator.
 *   qual uses the same offsets; qual == NULL means no read has qualities; a read
 *   without quality inside a batch that has a qual array carries 0xFF in all of
 *   its qual bytes.  off[] has n_reads+1 entries; off[n_reads] = total bytes.
 */
#ifndef BFC_B200_ABI_H
#define BFC_B200_ABI_H

#include <stdint.h>
#include "bfc.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { BFCG_OK = 0, BFCG_ERR_CUDA = -1, BFCG_ERR_ARG = -2, BFCG_ERR_NOMEM = -3, BFCG_ERR_OVERFLOW = -4 };

enum { BFCG_HOST = 0, BFCG_DEVICE = 1 };

typedef struct {
	int64_t n_reads;
	uint64_t n_bytes;        /* = off[n_reads]: total bytes of seq (and of qual) */
	int where;               /* BFCG_HOST: pointers are host memory; BFCG_DEVICE: device memory */
	const uint64_t *off;     /* n_reads + 1 (not needed by bfcg_count_batch) */
	uint8_t *seq;            /* n_bytes bytes; edited in place by bfcg_correct_batch */
	uint8_t *qual;           /* same, or NULL */
} bfcg_batch_t;

/* counters filled by the batch calls (all cumulative adds) */
typedef struct {
	uint64_t n_kmers;        /* k-mer occurrences enumerated                               */
	uint64_t n_pass;         /* occurrences whose Bloom bits were all set (count.c:60)      */
	uint64_t n_pending;      /* occurrences that set at least one new bit                   */
	uint64_t n_conflict;     /* of those, resolved by the ordered replay                    */
	uint64_t n_lookups;      /* bfc_ch_kmer_occ / bfc_bf_get calls made by correct / trim   */
	uint64_t n_redo;         /* reads re-run with a larger search scratch                   */
	uint64_t n_launches;     /* kernels launched                                            */
	double   kernel_ms;      /* device time of those kernels (CUDA events), when timing on  */
	uint64_t n_search_lookups; /* of n_lookups: those made inside the search kernel (k_ec_search) */
} bfcg_stats_t;

/* device selection / info ------------------------------------------------------- */
int  bfcg_device_count(void);
int  bfcg_set_device(int dev);                 /* default 0 (or LOCAL_RANK when set) */
int  bfcg_sync(void);
void bfcg_set_timing(int on);                  /* time kernels with CUDA events into stats.kernel_ms */
const char *bfcg_last_error(void);
/* per-kernel device time accumulated since the last call (needs bfcg_set_timing(1)); indices:
 * 0 count_probe 1 count_resolve 2 conflict sort 3 count_replay 4 correct (search) 5 correct redo 6 trim
 * 7 table rehash 8 table hist 9 table deferred-apply 10 k-mer enumeration 11 correct: batched lookups
 * 12 correct: coverage flags + per-read setup 13 correct: merge + rewrite 14 sharded count: owner bucketing
 * 15 count: partition kernel (Bloom slices in shared memory) 16 count: partition bounds 17 count: stream-order enumeration
 * 18 correct: lookups past the end of the reads.
 * Returns the number of valid entries. */
int  bfcg_kernel_times(double *ms, uint64_t *launches, int n);
/* CUDA events on the engine's stream: record into slot 0..7, elapsed ms between two slots */
int  bfcg_event_record(int slot);
double bfcg_event_elapsed_ms(int a, int b);

/* count: insert every k-mer of the batch, in read order, as count.c:54-89 does.
 * bf: first Bloom filter; exactly one of ch / bf_high is non-NULL
 * (normal mode / opt->filter_mode). */
int bfcg_count_batch(const bfc_opt_t *opt, bfc_bf_t *bf, bfc_bf_t *bf_high, bfc_ch_t *ch,
                     const bfcg_batch_t *batch, bfcg_stats_t *stats);

/* correct: bfc_ec1 on every read (correct.c:388-472).  seq/qual are rewritten in
 * place exactly as there (untouched on failure); aux[2*i] = s->aux, aux[2*i+1] =
 * s->aux2 as packed by worker_ec (correct.c:552-553).  `mode` is the return value
 * of bfc_ch_hist.  aux lives where the batch lives (host or device).
 * Refine mode (opt->refine_ec, `bfc -R`): aux is IN/OUT -- on entry aux[2*i], aux[2*i+1] hold the read's earlier
 * stats in the same packing (what parse_stats reads from the ec:Z: tag into e->ori_st, correct.c:517-531, 543);
 * bases an earlier round corrected are taken back from the quality string (correct.c:31), and a read whose new
 * search leaves more absent k-mers than the earlier one keeps its bytes and its earlier stats with rf_code 2
 * (correct.c:438-442).  Which reads are left out altogether (correct.c:544-545) is the caller's business. */
int bfcg_correct_batch(const bfc_opt_t *opt, const bfc_ch_t *ch, int mode,
                       bfcg_batch_t *batch, uint32_t *aux, bfcg_stats_t *stats);

/* trim: max_streak + the keep rule (correct.c:478-497, 555-569).  keep[i] = 1 and
 * the kept bases are [tstart[i], tend[i]), or keep[i] = 0.  The caller moves the
 * bytes (as worker_ec's memmove does). */
int bfcg_trim_batch(const bfc_opt_t *opt, const bfc_bf_t *bf_high, const bfcg_batch_t *batch,
                    uint8_t *keep, int32_t *tstart, int32_t *tend, bfcg_stats_t *stats);

/* sharded counting: one rank of N (N = 1, 2, 4, 8), one process per GPU -----------
 * The owner of a k-mer is given by the top log2(N) bits of its Bloom block index (bbf.c:27-28), so rank r
 * holds blocks [r, r+1) * 2^(b-9) / N of the first filter (and of bf_high) and the table entries of exactly
 * those k-mers.  Records are 16 bytes: y0 | is_high << 63 and y1 (the two words of bfc_kmer_hash).
 *
 * bfcg_enum_records: enumerate the k-mers of a batch (count.c:72-89) and bucket them by owner; inside a bucket the
 * records of one Bloom block keep the stream order (see bfcg_count_record_runs).  d_y0 / d_y1: device arrays with room for batch->n_bytes records; counts[n_owners]
 * (host) receives the bucket sizes; bucket o starts at counts[0] + ... + counts[o-1].
 * bfcg_count_records: the Bloom -> table cascade (count.c:54-70) over records in stream order (what the
 * all-to-all delivers: the pieces of the ranks concatenated in rank order), against this rank's shard. */
int bfcg_enum_records(const bfc_opt_t *opt, const bfcg_batch_t *batch, int n_owners,
                      uint64_t *d_y0, uint64_t *d_y1, uint64_t *counts);
int bfcg_count_records(const bfc_opt_t *opt, bfc_bf_t *bf, bfc_bf_t *bf_high, bfc_ch_t *ch, uint64_t n_rec,
                       const uint64_t *d_y0, const uint64_t *d_y1, int n_owners, bfcg_stats_t *stats);
/* bfcg_count_record_runs: the same cascade over exactly what an all-to-all of bfcg_enum_records buckets delivers: n_runs
 * pieces back to back in source-rank order (= global read order), piece i holding run_counts[i] records.  When the Bloom
 * block index is a bit field of y0 (n_shift - 9 <= k) bfcg_enum_records orders every bucket by count partition (stable,
 * so the order inside a Bloom block is still the read order) and this call replays the pieces without sorting again;
 * otherwise both keep stream order.  d_y1 is overwritten. */
int bfcg_count_record_runs(const bfc_opt_t *opt, bfc_bf_t *bf, bfc_bf_t *bf_high, bfc_ch_t *ch, int n_runs, const uint64_t *run_counts,
                           uint64_t *d_y0, uint64_t *d_y1, int n_owners, bfcg_stats_t *stats);
bfc_bf_t *bfcg_bf_init_shard(int n_shift, int n_hashes, int n_owners);   /* 2^(n_shift-3) / n_owners bytes */
/* declare an EMPTY table to be shard `owner` of `n_owners`: it then allocates (and sizes itself for) only the 1/n_owners
 * of the sub-table regions its k-mers can fall into (the owner bits are sub-table index bits, htab.c:45-58) */
int bfcg_ch_set_shard(bfc_ch_t *ch, int n_owners, int owner);
/* table exchange: entries as (sub-table index, key50<<14 | val14) in DEVICE arrays, unsorted; returns n
 * (pass NULLs to get n); import adds entries that are not present yet (shards are disjoint) */
uint64_t bfcg_ch_export_device(const bfc_ch_t *ch, uint32_t *d_sub, uint64_t *d_key);
int      bfcg_ch_import_device(bfc_ch_t *ch, uint64_t n, const uint32_t *d_sub, const uint64_t *d_key);

/* sharded counting with the exchange INSIDE the library (csrc/dist.cu): one rank per GPU, NCCL over NVLink ---------
 * The caller only brings the ranks together: one rank makes an id (bfcg_dist_unique_id), hands its 128 bytes to the
 * others by whatever means it has (MPI, a file, torch.distributed's store), and every rank calls bfcg_dist_init after
 * selecting its device.  Then, per global chunk of reads, EVERY rank calls bfcg_dist_count_piece with its piece (a
 * collective; an empty piece when it has run out of reads): the piece is enumerated and sorted by (owner, count
 * partition) into packed records, the buckets travel by grouped ncclSend / ncclRecv on a stream of their own while the
 * engine's stream runs the cascade over the previous chunk's records and enumerates the next piece.
 * bfcg_dist_count_finish drains the pipeline; bf / bf_high are shards (bfcg_bf_init_shard), ch a shard table
 * (bfcg_ch_set_shard).  Results: the reference's `-t1` run over all reads in global order, bit for bit.
 * bfcg_dist_gather_table / _filter replicate the result on every rank (what the correction / trim phase needs). */
#define BFCG_DIST_ID_BYTES 128
int  bfcg_dist_unique_id(void *id);
int  bfcg_dist_init(int rank, int world, const void *id);
void bfcg_dist_finalize(void);
int  bfcg_dist_rank(void);
int  bfcg_dist_world(void);
int  bfcg_dist_count_piece(const bfc_opt_t *opt, bfc_bf_t *bf, bfc_bf_t *bf_high, bfc_ch_t *ch,
                           const bfcg_batch_t *piece, bfcg_stats_t *stats);
int  bfcg_dist_count_finish(const bfc_opt_t *opt, bfc_bf_t *bf, bfc_bf_t *bf_high, bfc_ch_t *ch, bfcg_stats_t *stats);
int  bfcg_dist_gather_table(const bfc_ch_t *shard, bfc_ch_t *full);
int  bfcg_dist_gather_filter(const bfc_bf_t *shard, bfc_bf_t *full);
int  bfcg_dist_allreduce_sum_u64(uint64_t *v, int n);      /* host arrays, n <= 64 */
int  bfcg_dist_allreduce_max_f64(double *v, int n);
int  bfcg_dist_barrier(void);
int  bfcg_dist_stats(uint64_t *sent_records, uint64_t *received_records, double *exchange_ms);

/* device memory helpers for callers that keep batches resident in HBM -------------- */
void *bfcg_dev_alloc(uint64_t bytes);
void  bfcg_dev_free(void *p);
int   bfcg_h2d(void *dst, const void *src, uint64_t bytes);
int   bfcg_d2h(void *dst, const void *src, uint64_t bytes);
void *bfcg_host_alloc_pinned(uint64_t bytes);
void  bfcg_host_free_pinned(void *p);

/* inspection (tests, dumps) --------------------------------------------------------- */
int      bfcg_bf_download(const bfc_bf_t *bf, uint8_t *dst);          /* 2^(n_shift-3) bytes */
int      bfcg_bf_upload(bfc_bf_t *bf, const uint8_t *src);
int      bfcg_bf_clear(bfc_bf_t *bf);
/* occupancy telemetry (a working version of what the reference's unused bfc_bf_load, bbf.c:65-79, was for): fraction of
 * data bits set, fraction of blocks touched, false-positive rate of the filter at that load; n_owners > 1: `bf` is a shard */
int      bfcg_bf_load(const bfc_bf_t *bf, int n_owners, double *bit_load, double *block_load, double *fp_rate);
/* the -b (log2 bits) that keeps the false-positive rate <= target_fp after n_distinct different k-mers (at most 37) */
int      bfcg_bf_suggest_shift(uint64_t n_distinct, int n_hashes, double target_fp);
/* all entries as (sub-table index, key50<<14 | val14), sorted by (sub, key); pass NULLs to get n */
uint64_t bfcg_ch_export(const bfc_ch_t *ch, uint32_t *sub, uint64_t *key);
int      bfcg_ch_l_pre(const bfc_ch_t *ch);
int      bfcg_ch_capacity_log2(const bfc_ch_t *ch);
int      bfcg_ch_clear(bfc_ch_t *ch);
int      bfcg_ch_reserve(bfc_ch_t *ch, uint64_t n_keys);              /* pre-size for n_keys distinct keys */
/* the count phase's stable radix partition on its own (csrc/partition.cuh): n < 2^30 records = 64-bit keys + vb-byte
 * values (vb = 1, 2, 4, 8) in device arrays, ordered by bits [begin, end) of the key, ties in input order */
int      bfcg_partition_records(const uint64_t *d_key_in, const void *d_val_in, uint64_t *d_key_out, void *d_val_out,
                                uint64_t n, int vb, int begin, int end);
/* batched lookups: y = n pairs of k-bit words (bfc_kmer_hash output); out[i] = bfc_ch_get */
int      bfcg_ch_get_batch(const bfc_ch_t *ch, int where, uint64_t n, const uint64_t *y, int32_t *out);

/* synthetic input for HBM-resident benchmarks (csrc/synth.cu; numpy twin in bfc_b200/synth.py) */
int bfcg_synth_genome(uint8_t *d_genome, uint64_t G, uint64_t seed);
int bfcg_synth_reads(const uint8_t *d_genome, uint64_t G, uint64_t seed, uint64_t first_read, int64_t n_reads, int L,
                     double err, double n_rate, uint8_t *d_seq, uint8_t *d_qual, uint64_t *d_off);

#ifdef __cplusplus
}
#endif

#endif
