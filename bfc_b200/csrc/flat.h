/* flat.h -- bseq1_t[] <-> the flat host batch of bfc_b200.h ("reads back to back"). */
#ifndef BFC_B200_FLAT_H
#define BFC_B200_FLAT_H

#include "bfc_b200.h"

typedef struct {
	bfcg_batch_t b;      /* b.off / b.seq / b.qual are owned (pinned when possible) */
	uint64_t *off;
	int pinned;
	const bseq1_t *seqs; /* scratch for the copy workers */
	int to_reads;
} flat_batch_t;

int  flat_from_reads(flat_batch_t *f, const bseq1_t *seqs, int n, int n_threads);
void flat_to_reads(const flat_batch_t *f, bseq1_t *seqs, int n_threads); /* copy seq/qual back (same lengths) */
void flat_free(flat_batch_t *f);

#endif
