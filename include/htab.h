/* htab.h -- counting hash table of the count/correct path, resident in GPU memory.
 *
 * Drop-in for the reference's htab.h: same opaque type and the same ten entry
 * points with the same argument meaning and return values.
 *
 *   bfc_ch_init      reference htab.c:19-34   (l_pre is adjusted exactly as there)
 *   bfc_ch_destroy   reference htab.c:36-43   accepts NULL
 *   bfc_ch_insert    reference htab.c:60-82   0 = done; never "busy" here, `forced` is ignored
 *   bfc_ch_get       reference htab.c:84-92   -1 = absent, else high6<<8 | cnt8
 *   bfc_ch_kmer_occ  reference htab.c:94-99
 *   bfc_ch_count     reference htab.c:101-108 number of distinct keys
 *   bfc_ch_hist      reference htab.c:110-127 count / high-count histograms, returns the mode (>=3) or -1
 *   bfc_ch_dump      reference htab.c:129-149 same file format (restorable by the reference's -r)
 *   bfc_ch_restore   reference htab.c:151-176
 *   bfc_ch_get_k     reference htab.c:178-181
 *
 * A key is identified by (sub-table index, key >> 14) exactly as in the reference
 * (htab.c:45-58), including the lossy XOR fold for k > 32; the value is the low 14
 * bits (6-bit high-quality count, 8-bit count, both saturating).  Instead of 2^l_pre
 * khash sets the keys live in ONE open-addressing array in HBM whose 2^l_pre equal
 * regions make the sub-table index implicit in the slot position, so a slot stays
 * 8 bytes and four slots share one 32-byte DRAM sector.
 */
#ifndef BFC_B200_HTAB_H
#define BFC_B200_HTAB_H

#include <stdint.h>
#include "kmer.h"

#define BFC_CH_KEYBITS 50
#define BFC_CH_MAXPRE  24

#ifdef __cplusplus
extern "C" {
#endif

struct bfc_ch_s;
typedef struct bfc_ch_s bfc_ch_t;

bfc_ch_t *bfc_ch_init(int k, int l_pre);
void bfc_ch_destroy(bfc_ch_t *ch);
int bfc_ch_insert(bfc_ch_t *ch, const uint64_t x[2], int is_high, int forced);
int bfc_ch_get(const bfc_ch_t *ch, const uint64_t x[2]);
uint64_t bfc_ch_count(const bfc_ch_t *ch);
int bfc_ch_hist(const bfc_ch_t *ch, uint64_t cnt[256], uint64_t high[64]);
int bfc_ch_dump(const bfc_ch_t *ch, const char *fn);
bfc_ch_t *bfc_ch_restore(const char *fn);
int bfc_ch_get_k(const bfc_ch_t *ch);

int bfc_ch_kmer_occ(const bfc_ch_t *ch, const bfc_kmer_t *z);

#ifdef __cplusplus
}
#endif

#endif
