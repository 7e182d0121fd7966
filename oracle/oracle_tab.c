/* TEST INFRASTRUCTURE ONLY -- oracle restatement of reference bbf.c and htab.c. */
#include <stdlib.h>
#include <string.h>
#include <assert.h>
#include "oracle_kmer.h"

void orc_kmer_append(int k, uint64_t x[4], int c) { okm_append(k, x, c); }
void orc_kmer_change(int k, uint64_t x[4], int d, int c) { okm_change(k, x, d, c); }
uint64_t orc_hash_64(uint64_t key, uint64_t mask) { return okm_mix(key, mask); }
uint64_t orc_kmer_hash(int k, const uint64_t x[4], uint64_t h[2]) { return okm_hash(k, x, h); }

/* ------------------------------------------------------------------ Bloom */

/* reference bbf.c:5-17 */
orc_bf_t *orc_bf_new(int n_shift, int n_hashes)
{
	orc_bf_t *b;
	if (n_shift + 9 > 64 || n_shift < 9) return 0;
	b = (orc_bf_t*)calloc(1, sizeof(orc_bf_t));
	b->n_shift = n_shift, b->n_hashes = n_hashes;
	b->b = (uint8_t*)calloc((size_t)1 << (n_shift - 3), 1);
	return b;
}

void orc_bf_free(orc_bf_t *b) { if (b) { free(b->b); free(b); } }

/* probe sequence shared by insert and get: reference bbf.c:27-33, 35-38 */
static inline uint8_t *bf_locate(const orc_bf_t *b, uint64_t hash, int *h1, int *h2)
{
	int x = b->n_shift - 9;
	uint64_t blk = hash & ((1ULL << x) - 1);
	*h1 = (int)(hash >> x & 511);
	*h2 = (int)(hash >> b->n_shift & 511);
	if ((*h2 & 31) == 0) *h2 = (*h2 + 1) & 511;
	return b->b + (blk << 6);
}

/* reference bbf.c:25-45 (single-threaded: the byte-0 lock is a no-op and ends 0) */
int orc_bf_insert(orc_bf_t *b, uint64_t hash)
{
	int h1, h2, z, done = 0, n_set = 0;
	uint8_t *p = bf_locate(b, hash, &h1, &h2);
	for (z = h1; done < b->n_hashes; z = (z + h2) & 511) {
		if (z < 8) continue; /* byte 0 is the lock; the probe does not count */
		n_set += p[z >> 3] >> (z & 7) & 1;
		p[z >> 3] |= (uint8_t)(1 << (z & 7));
		++done;
	}
	return n_set;
}

/* reference bbf.c:47-63 */
int orc_bf_get(const orc_bf_t *b, uint64_t hash)
{
	int h1, h2, z, done = 0, n_set = 0;
	const uint8_t *p = bf_locate(b, hash, &h1, &h2);
	for (z = h1; done < b->n_hashes; z = (z + h2) & 511) {
		if (z < 8) continue;
		n_set += p[z >> 3] >> (z & 7) & 1;
		++done;
	}
	return n_set;
}

/* ------------------------------------------------------------------ table
 * One flat open-addressing map keyed by (sub-table index, key>>14); value = low 14
 * bits.  Only the *results* of get/insert follow the reference (htab.c); the slot
 * layout of khash is not reproduced. */

typedef struct { uint64_t key; uint32_t sub; uint32_t used; } och_ent_t;

struct orc_ch_s {
	int k, l_pre;
	uint64_t cap, n; /* cap is a power of two */
	och_ent_t *a;
};

/* reference htab.c:19-34 */
orc_ch_t *orc_ch_new(int k, int l_pre)
{
	orc_ch_t *ch;
	assert(k <= 63);
	if (k * 2 - l_pre > 50) l_pre = k * 2 - 50;
	if (l_pre > 24) l_pre = 24;
	assert(k - l_pre < 50);
	ch = (orc_ch_t*)calloc(1, sizeof(orc_ch_t));
	ch->k = k, ch->l_pre = l_pre;
	ch->cap = 1 << 16;
	ch->a = (och_ent_t*)calloc(ch->cap, sizeof(och_ent_t));
	return ch;
}

void orc_ch_free(orc_ch_t *ch) { if (ch) { free(ch->a); free(ch); } }
int orc_ch_k(const orc_ch_t *ch) { return ch->k; }
int orc_ch_lpre(const orc_ch_t *ch) { return ch->l_pre; }

/* reference htab.c:45-58 */
void orc_ch_subkey(const orc_ch_t *ch, const uint64_t y[2], uint32_t *sub, uint64_t *key)
{
	int k = ch->k;
	if (k <= 32) {
		int t = 2 * k - ch->l_pre;
		uint64_t z = y[0] << k | y[1];
		*key = (z & ((1ULL << t) - 1)) << 14 | 1;
		*sub = (uint32_t)(z >> t);
	} else {
		int t = k - ch->l_pre;
		int shift = t + k < 50 ? k : 50 - t;
		*key = ((y[0] & ((1ULL << t) - 1)) << shift ^ y[1]) << 14 | 1;
		*sub = (uint32_t)(y[0] >> t);
	}
}

static inline uint64_t och_slot(uint32_t sub, uint64_t key, uint64_t cap)
{
	uint64_t h = (key >> 14) * 0x9E3779B97F4A7C15ULL ^ (uint64_t)sub * 0xC2B2AE3D27D4EB4FULL;
	h ^= h >> 29;
	return h & (cap - 1);
}

static och_ent_t *och_find(const orc_ch_t *ch, uint32_t sub, uint64_t key)
{
	uint64_t i = och_slot(sub, key, ch->cap);
	for (;;) {
		och_ent_t *e = &ch->a[i];
		if (!e->used) return e;
		if (e->sub == sub && e->key >> 14 == key >> 14) return e;
		i = (i + 1) & (ch->cap - 1);
	}
}

static void och_grow(orc_ch_t *ch)
{
	uint64_t old_cap = ch->cap, i;
	och_ent_t *old = ch->a;
	ch->cap <<= 1;
	ch->a = (och_ent_t*)calloc(ch->cap, sizeof(och_ent_t));
	for (i = 0; i < old_cap; ++i)
		if (old[i].used) *och_find(ch, old[i].sub, old[i].key) = old[i];
	free(old);
}

/* reference htab.c:60-82 (locks omitted: sequential) */
int orc_ch_insert(orc_ch_t *ch, const uint64_t y[2], int is_high)
{
	uint32_t sub; uint64_t key;
	och_ent_t *e;
	orc_ch_subkey(ch, y, &sub, &key);
	if ((ch->n + 1) * 2 > ch->cap) och_grow(ch);
	e = och_find(ch, sub, key);
	if (!e->used) {
		e->used = 1, e->sub = sub, e->key = key; /* count field starts at 1 */
		if (is_high) e->key |= 1 << 8;
		++ch->n;
	} else {
		if ((e->key & 0xff) != 0xff) ++e->key;
		if (is_high && (e->key >> 8 & 0x3f) != 0x3f) e->key += 1 << 8;
	}
	return 0;
}

void orc_ch_put_raw(orc_ch_t *ch, uint32_t sub, uint64_t key)
{
	och_ent_t *e;
	if ((ch->n + 1) * 2 > ch->cap) och_grow(ch);
	e = och_find(ch, sub, key);
	if (!e->used) ++ch->n;
	e->used = 1, e->sub = sub, e->key = key;
}

/* reference htab.c:84-92 */
int orc_ch_get(const orc_ch_t *ch, const uint64_t y[2])
{
	uint32_t sub; uint64_t key;
	const och_ent_t *e;
	orc_ch_subkey(ch, y, &sub, &key);
	e = och_find(ch, sub, key);
	return e->used ? (int)(e->key & 0x3fff) : -1;
}

/* reference htab.c:94-99 */
int orc_ch_kmer_occ(const orc_ch_t *ch, const orc_kmer_t *z)
{
	uint64_t y[2];
	okm_hash(ch->k, z->x, y);
	return orc_ch_get(ch, y);
}

/* reference htab.c:101-108 */
uint64_t orc_ch_count(const orc_ch_t *ch) { return ch->n; }

/* reference htab.c:110-127 */
int orc_ch_hist(const orc_ch_t *ch, uint64_t cnt[256], uint64_t high[64])
{
	uint64_t i, best = 0;
	int c, mode = -1;
	memset(cnt, 0, 256 * sizeof(uint64_t));
	memset(high, 0, 64 * sizeof(uint64_t));
	for (i = 0; i < ch->cap; ++i)
		if (ch->a[i].used) ++cnt[ch->a[i].key & 0xff], ++high[ch->a[i].key >> 8 & 0x3f];
	for (c = 3; c < 256; ++c)
		if (cnt[c] > best) best = cnt[c], mode = c;
	return mode;
}

static int och_cmp(const void *pa, const void *pb)
{
	const och_ent_t *a = (const och_ent_t*)pa, *b = (const och_ent_t*)pb;
	if (a->sub != b->sub) return a->sub < b->sub ? -1 : 1;
	if (a->key != b->key) return a->key < b->key ? -1 : 1;
	return 0;
}

uint64_t orc_ch_export(const orc_ch_t *ch, uint32_t *sub, uint64_t *key)
{
	uint64_t i, n = 0;
	och_ent_t *tmp;
	if (sub == 0 || key == 0) return ch->n;
	tmp = (och_ent_t*)malloc((ch->n + 1) * sizeof(och_ent_t));
	for (i = 0; i < ch->cap; ++i)
		if (ch->a[i].used) tmp[n++] = ch->a[i];
	qsort(tmp, n, sizeof(och_ent_t), och_cmp);
	for (i = 0; i < n; ++i) sub[i] = tmp[i].sub, key[i] = tmp[i].key;
	free(tmp);
	return n;
}
