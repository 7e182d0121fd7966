/* TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * Link-time hooks for the UNMODIFIED reference objects (built by oracle/Makefile
 * from the sources where they lie under /root/reference into oracle/_ref/).
 *
 * The reference frees its first Bloom filter inside bfc_count() (count.c:155), so
 * the only way to observe the Bloom bit-vector without editing reference sources
 * is `ld --wrap=bfc_bf_destroy`: every call the reference makes to
 * bfc_bf_destroy() lands here first.  When BFC_REF_BLOOM_DUMP=<prefix> is set,
 * the filter bytes are written to <prefix>.<n> (n = 0 for the first filter
 * destroyed, 1 for the second, ...) before the real destructor runs.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

typedef struct { int n_shift, n_hashes; uint8_t *b; } ref_bf_t; /* layout: bbf.h:9-12 */

void __real_bfc_bf_destroy(ref_bf_t *b);

void __wrap_bfc_bf_destroy(ref_bf_t *b)
{
	static int n_dumped = 0;
	const char *prefix = getenv("BFC_REF_BLOOM_DUMP");
	if (b && prefix) {
		char fn[4096];
		FILE *fp;
		snprintf(fn, sizeof(fn), "%s.%d", prefix, n_dumped++);
		if ((fp = fopen(fn, "wb")) != 0) {
			int32_t hdr[2] = { b->n_shift, b->n_hashes };
			fwrite(hdr, 4, 2, fp);
			fwrite(b->b, 1, (size_t)1 << (b->n_shift - 3), fp);
			fclose(fp);
		}
	}
	__real_bfc_bf_destroy(b);
}
