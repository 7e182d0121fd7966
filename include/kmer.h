/* kmer.h -- bit-sliced k-mer arithmetic of the count/correct path (host + device).
 *
 * Drop-in for the reference's kmer.h: same type and inline function names, same
 * results bit for bit, so code written against the reference header compiles
 * unchanged.  Every function is usable from CUDA device code as well (BFC_HD).
 *
 *   bfc_kmer_t        reference kmer.h:6-8    four bit planes of <=63 bits
 *   bfc_kmer_append   reference kmer.h:10-17  roll one base in (both strands)
 *   bfc_kmer_change   reference kmer.h:19-27  point edit d bases from the 3'-end
 *   bfc_hash_64       reference kmer.h:30-40  invertible 64-bit mix under a k-bit mask
 *   bfc_hash_64_inv   reference kmer.h:42-77
 *   bfc_kmer_hash     reference kmer.h:79-88  strand pick + two mixes -> (hash, y[2])
 *   bfc_kmer_hash_inv reference kmer.h:90-95
 *   bfc_kmer_2str     reference kmer.h:97-104
 *
 * Plane layout: x[0]/x[1] hold the low/high bit of each base of the forward strand
 * with the newest base at bit 0; x[2]/x[3] hold the reverse complement with the
 * newest base at bit k-1.
 */
#ifndef BFC_B200_KMER_H
#define BFC_B200_KMER_H

#include <stdint.h>

#if defined(__CUDACC__)
#define BFC_HD __host__ __device__ __forceinline__
#else
#define BFC_HD static inline
#endif

typedef struct {
	uint64_t x[4];
} bfc_kmer_t;

/* c must be 0..3 */
BFC_HD void bfc_kmer_append(int k, uint64_t x[4], int c)
{
	const uint64_t m = (1ULL << k) - 1;
	const uint64_t lo = (uint64_t)(c & 1), hi = (uint64_t)(c >> 1);
	x[0] = ((x[0] << 1) | lo) & m;
	x[1] = ((x[1] << 1) | hi) & m;
	x[2] = (x[2] >> 1) | ((lo ^ 1ULL) << (k - 1));
	x[3] = (x[3] >> 1) | ((hi ^ 1ULL) << (k - 1));
}

/* replace the base d positions from the 3'-end (0 <= d < k) by c (0..3) */
BFC_HD void bfc_kmer_change(int k, uint64_t x[4], int d, int c)
{
	const uint64_t lo = (uint64_t)(c & 1), hi = (uint64_t)(c >> 1);
	const int r = k - 1 - d;
	x[0] = (x[0] & ~(1ULL << d)) | (lo << d);
	x[1] = (x[1] & ~(1ULL << d)) | (hi << d);
	x[2] = (x[2] & ~(1ULL << r)) | ((lo ^ 1ULL) << r);
	x[3] = (x[3] & ~(1ULL << r)) | ((hi ^ 1ULL) << r);
}

/* Thomas Wang's 64-bit mix restricted to the low bits selected by `mask` */
BFC_HD uint64_t bfc_hash_64(uint64_t key, uint64_t mask)
{
	key = (~key + (key << 21)) & mask;
	key ^= key >> 24;
	key = (key + (key << 3) + (key << 8)) & mask;   /* * 265 */
	key ^= key >> 14;
	key = (key + (key << 2) + (key << 4)) & mask;   /* * 21 */
	key ^= key >> 28;
	key = (key + (key << 31)) & mask;
	return key;
}

BFC_HD uint64_t bfc_hash_64_inv(uint64_t key, uint64_t mask)
{
	uint64_t t;
	/* undo key += key << 31 */
	t = key - (key << 31);
	key = (key - (t << 31)) & mask;
	/* undo key ^= key >> 28 */
	t = key ^ (key >> 28);
	key ^= t >> 28;
	/* undo * 21 (multiplicative inverse mod 2^64) */
	key = (key * 14933078535860113213ULL) & mask;
	/* undo key ^= key >> 14 */
	t = key ^ (key >> 14);
	t = key ^ (t >> 14);
	t = key ^ (t >> 14);
	key ^= t >> 14;
	/* undo * 265 */
	key = (key * 15244667743933553977ULL) & mask;
	/* undo key ^= key >> 24 */
	t = key ^ (key >> 24);
	key ^= t >> 24;
	/* undo key = ~key + (key << 21) */
	t = ~key;
	t = ~(key - (t << 21));
	t = ~(key - (t << 21));
	key = ~(key - (t << 21)) & mask;
	return key;
}

/* Returns the 64-bit Bloom hash; h[0], h[1] receive the two k-bit table words. */
BFC_HD uint64_t bfc_kmer_hash(int k, const uint64_t x[4], uint64_t h[2])
{
	const int mid = k >> 1;
	const int u = ((x[1] >> mid) & 1) > ((x[3] >> mid) & 1); /* pick strand by the middle base */
	const uint64_t m = (1ULL << k) - 1;
	const uint64_t a = x[u << 1], b = x[(u << 1) | 1];
	const uint64_t h0 = bfc_hash_64((a + b) & m, m);
	const uint64_t h1 = bfc_hash_64(h0 ^ b, m);
	const uint64_t s = (h0 + h1) & m;
	h[0] = s;
	h[1] = h1;
	return ((h0 ^ h1) << k) | s;
}

BFC_HD void bfc_kmer_hash_inv(int k, const uint64_t h[2], uint64_t y[2])
{
	const uint64_t m = (1ULL << k) - 1, t = (h[0] - h[1]) & m;
	y[1] = bfc_hash_64_inv(h[1], m) ^ t;
	y[0] = (bfc_hash_64_inv(t, m) - y[1]) & m;
}

BFC_HD char *bfc_kmer_2str(int k, const uint64_t y[2], char *buf)
{
	int l;
	for (l = 0; l < k; ++l) {
		const int c = (int)(((y[1] >> l) & 1) << 1 | ((y[0] >> l) & 1));
		buf[k - 1 - l] = (char)(c == 0 ? 'A' : c == 1 ? 'C' : c == 2 ? 'G' : 'T');
	}
	buf[k] = 0;
	return buf;
}

#endif
