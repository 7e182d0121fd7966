/* bseq.h -- FASTA/FASTQ batch reader feeding the count/correct pipeline.
 *
 * Drop-in for the reference's bseq.h (bseq.h:9-19): same record type, same three
 * functions and the same nt6 table.  bseq_read() returns a malloc'd array of at
 * least `chunk_size` bases worth of records (all of them malloc'd strings the caller
 * frees), NULL/0 at end of input.  Plain and gzip'd input, "-" = stdin.
 */
#ifndef BFC_B200_BSEQ_H
#define BFC_B200_BSEQ_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

struct bseq_file_s;
typedef struct bseq_file_s bseq_file_t;

typedef struct {
	int l_seq;
	uint32_t aux, aux2;
	char *name, *comment, *seq, *qual;
} bseq1_t;

extern unsigned char seq_nt6_table[256];

bseq_file_t *bseq_open(const char *fn);
void bseq_close(bseq_file_t *fp);
bseq1_t *bseq_read(bseq_file_t *fp, int chunk_size, int keep_comment, int *n_);

#ifdef __cplusplus
}
#endif

#endif
