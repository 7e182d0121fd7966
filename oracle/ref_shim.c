/* TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * kmer.h of the reference is header-only `static inline`, so its functions have no
 * symbols.  This shim (compiled together with the unmodified reference sources into
 * oracle/_ref/libbfcref.so, see oracle/Makefile) gives them C symbols so that the
 * tests can pin this repo's restatement against the reference's own code.
 */
#include <stdint.h>
#include "kmer.h"   /* the reference's, via -I$(REF) */

void ref_kmer_append(int k, uint64_t x[4], int c) { bfc_kmer_append(k, x, c); }
void ref_kmer_change(int k, uint64_t x[4], int d, int c) { bfc_kmer_change(k, x, d, c); }
uint64_t ref_hash_64(uint64_t key, uint64_t mask) { return bfc_hash_64(key, mask); }
uint64_t ref_hash_64_inv(uint64_t key, uint64_t mask) { return bfc_hash_64_inv(key, mask); }
uint64_t ref_kmer_hash(int k, const uint64_t x[4], uint64_t h[2]) { return bfc_kmer_hash(k, x, h); }
void ref_kmer_hash_inv(int k, const uint64_t h[2], uint64_t y[2]) { bfc_kmer_hash_inv(k, h, y); }
