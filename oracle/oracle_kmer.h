/* TEST INFRASTRUCTURE ONLY -- oracle restatement of reference kmer.h. */
#ifndef ORACLE_KMER_H
#define ORACLE_KMER_H
#include "oracle.h"

#define ORC_MASK(k) ((1ULL << (k)) - 1ULL)

/* reference kmer.h:10-17 */
static inline void okm_append(int k, uint64_t x[4], int c)
{
	uint64_t m = ORC_MASK(k);
	x[0] = (x[0] << 1 | (uint64_t)(c & 1)) & m;
	x[1] = (x[1] << 1 | (uint64_t)(c >> 1 & 1)) & m;
	x[2] = x[2] >> 1 | (uint64_t)(1 - (c & 1)) << (k - 1);
	x[3] = x[3] >> 1 | (uint64_t)(1 - (c >> 1 & 1)) << (k - 1);
}

/* reference kmer.h:19-27 */
static inline void okm_change(int k, uint64_t x[4], int d, int c)
{
	int e = k - 1 - d;
	x[0] = (x[0] & ~(1ULL << d)) | (uint64_t)(c & 1) << d;
	x[1] = (x[1] & ~(1ULL << d)) | (uint64_t)(c >> 1 & 1) << d;
	x[2] = (x[2] & ~(1ULL << e)) | (uint64_t)(1 - (c & 1)) << e;
	x[3] = (x[3] & ~(1ULL << e)) | (uint64_t)(1 - (c >> 1 & 1)) << e;
}

/* reference kmer.h:30-40 */
static inline uint64_t okm_mix(uint64_t v, uint64_t m)
{
	v = ((v << 21) - v - 1) & m;
	v ^= v >> 24;
	v = (v * 265) & m;
	v ^= v >> 14;
	v = (v * 21) & m;
	v ^= v >> 28;
	v = (v + (v << 31)) & m;
	return v;
}

/* reference kmer.h:79-88 */
static inline uint64_t okm_hash(int k, const uint64_t x[4], uint64_t h[2])
{
	uint64_t m = ORC_MASK(k), a, b, h0, h1;
	int fwd_mid_hi = (int)(x[1] >> (k >> 1) & 1), rev_mid_hi = (int)(x[3] >> (k >> 1) & 1);
	if (fwd_mid_hi > rev_mid_hi) a = x[2], b = x[3];
	else a = x[0], b = x[1];
	h0 = okm_mix((a + b) & m, m);
	h1 = okm_mix(h0 ^ b, m);
	h[0] = (h0 + h1) & m;
	h[1] = h1;
	return (h0 ^ h1) << k | h[0];
}

#endif
