/* bbf.h -- blocked Bloom filter of the count path, resident in GPU memory.
 *
 * Drop-in for the reference's bbf.h (same struct layout, same four entry points,
 * same return values):
 *
 *   bfc_bf_t        reference bbf.h:9-12   {n_shift, n_hashes, b}
 *   bfc_bf_init     reference bbf.c:5-17   NULL when n_shift + 9 > 64 or n_shift < 9
 *   bfc_bf_destroy  reference bbf.c:19-23  accepts NULL
 *   bfc_bf_insert   reference bbf.c:25-45  returns how many of the n_hashes bits were already set
 *   bfc_bf_get      reference bbf.c:47-63  returns how many of the n_hashes bits are set
 *
 * The filter has 2^n_shift bits in 64-byte blocks; bits 0..7 of a block (the
 * reference's spin-lock byte) never hold data and are always 0 between calls.
 * DIFFERENCE: `b` is a DEVICE pointer (cudaMalloc).  Host code must not
 * dereference it; use bfc_bf_insert/get (one-element kernels, for API
 * compatibility) or the batch entry points in bfc_b200.h, or copy it out with
 * bfcg_bf_download().  There is no CPU fallback: every call needs a CUDA device.
 */
#ifndef BFC_B200_BBF_H
#define BFC_B200_BBF_H

#include <stdint.h>

#define BFC_BLK_SHIFT  9                          /* 512-bit = 64-byte blocks */
#define BFC_BLK_MASK   ((1 << (BFC_BLK_SHIFT)) - 1)

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	int n_shift, n_hashes;
	uint8_t *b;                                   /* device memory, 2^(n_shift-3) bytes */
} bfc_bf_t;

bfc_bf_t *bfc_bf_init(int n_shift, int n_hashes);
void bfc_bf_destroy(bfc_bf_t *b);
int bfc_bf_insert(bfc_bf_t *b, uint64_t hash);
int bfc_bf_get(const bfc_bf_t *b, uint64_t hash);

#ifdef __cplusplus
}
#endif

#endif
