/* fqblock.h -- block-wise FASTA/FASTQ ingest and egress for the phase drivers.
 *
 * Replaces, on the way in, the record-at-a-time bseq_read() (reference bseq.c:52-76 over
 * kseq.h:185-224: one getc state machine and four mallocs per read) and, on the way out,
 * the printf-per-read writer (reference correct.c:591-616).  Input is read in blocks of
 * text; a block that is plain four-line FASTQ -- every record "@name[ comment]", one
 * sequence line, "+...", one quality line of the same length, "\n" line ends -- is split
 * by all -t threads at once (the newlines of every 2 MB piece are counted while it is read; the counts give the
 * line number of every piece's first byte; each piece then parses the records that start in it -- one pass over the
 * text, no index of line starts).  Anything
 * else (FASTA, multi-line records, "\r\n", blank lines, a truncated tail) switches the
 * reader, from the start of that block on, to the tolerant sequential parser in bseq.c,
 * so the records delivered are always exactly those bseq_read() would deliver, including
 * kseq's two quirks (sticky comments; stop at the first record whose quality length
 * differs).  Records never get their own allocations: a block owns one text buffer and
 * arrays of offsets into it.
 */
#ifndef BFC_B200_FQBLOCK_H
#define BFC_B200_FQBLOCK_H

#include <stdint.h>
#include <stdio.h>
#include "bfc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define FQ_NONE UINT64_MAX

typedef struct {
	char *buf;                 /* the text the offsets point into (owned) */
	size_t buf_len;
	int64_t n;                 /* records */
	uint64_t *name_off, *com_off, *seq_off, *qual_off; /* com_off / qual_off = FQ_NONE: no comment / no quality */
	uint32_t *name_len, *com_len, *seq_len;
	uint64_t n_bases;
	int any_qual;
} fq_block_t;

struct fq_reader_s;
typedef struct fq_reader_s fq_reader_t;

fq_reader_t *fq_open(const char *fn, int n_threads);   /* plain or gzip'd, "-" = stdin; NULL on failure */
void fq_close(fq_reader_t *r);
/* next block of about target_bytes of input text: 1, 0 at end of input (blk zeroed), -1 out of memory */
int  fq_next(fq_reader_t *r, size_t target_bytes, int keep_comment, fq_block_t *blk);
void fq_block_free(fq_block_t *blk);
int  fq_reader_is_fast(const fq_reader_t *r);           /* still on the parallel path (tests) */

/* blocks kept from the count pass for the correct pass over the same file (both passes must ask for the same
 * keep_comment).  By default only for gzip'd input, and everything is dropped again when the blocks do not fit into
 * a quarter of the host's memory; BFC_B200_KEEP_MAX=<bytes> sets the budget for every regular file (0 = never keep) */
void fq_keep_begin(const char *fn, int keep_comment);   /* forgets what was kept; fn must be a regular file for anything to be kept */
int  fq_keep_add(fq_block_t *blk);                      /* 1 = kept: the block now belongs to the cache and *blk is zeroed */
void fq_keep_end(int complete);                         /* complete: every block of the input was offered */
long fq_keep_match(const char *fn, int keep_comment);   /* number of blocks kept for exactly this file (size, mtime), or -1 */
int  fq_keep_take(long i, fq_block_t *blk);             /* moves block i out */
void fq_keep_drop(void);

/* the flat batch of bfc_b200.h over a block: pinned buffers that are reused from batch to batch */
typedef struct {
	bfcg_batch_t b;            /* b.qual is NULL for a batch without any quality string */
	uint64_t *off;
	int64_t *flat_idx;         /* per record of the block: its index in the batch, -1 = left out (fq_flat_fill's skip) */
	uint8_t *seq_buf, *qual_buf;
	size_t cap_bytes, cap_reads;
	int pinned;
} fq_flat_t;

/* skip: NULL, or one byte per record, non-zero = the record does not go into the batch.  0 ok, -1 out of memory */
int  fq_flat_fill(fq_flat_t *f, const fq_block_t *blk, const uint8_t *skip, int n_threads);
void fq_flat_free(fq_flat_t *f);

/* the writer of correct.c:591-611 over a block + the (corrected / trimmed) flat batch: formats on n_threads
 * threads, then writes the pieces in order.  aux/aux2 as packed by worker_ec (correct.c:552-553); in filter mode
 * keep/tstart/tend give the kept stretch (aux = !keep). */
typedef struct {
	int filter_mode, discard, no_qual;
	int refine;                 /* -R: records left out of the batch are written as they came (comment included, correct.c:544-545);
	                               the others lose their comment and get a fresh tag (correct.c:547-550) */
	const uint32_t *aux;        /* 2 per read OF THE BATCH (normal mode) */
	const uint8_t *keep;        /* filter mode */
	const int32_t *tstart, *tend;
} fq_out_t;

int fq_write(FILE *fp, const fq_block_t *blk, const fq_flat_t *flat, const fq_out_t *o, int n_threads);

#ifdef __cplusplus
}
#endif

#endif
