/* fqblock.c -- block-wise FASTA/FASTQ ingest and egress (see fqblock.h). */
#include <zlib.h>
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <fcntl.h>
#include <sys/stat.h>
#include "bfc.h"
#include "fqblock.h"

bseq_file_t *bseq_open_from(void *gz, unsigned char *pre, size_t pre_len, const char *comment); /* bseq.c */
int bseq_at_eof(const bseq_file_t *f);

/* ---------------------------------------------------------------- recycled large buffers
 * A batch is a quarter of a gigabyte of text plus tens of megabytes of offsets and as much formatted output again.
 * Fresh from malloc, every one of those pages is mapped, zeroed and faulted in on first touch and unmapped on free --
 * per batch -- which cost more than reading and splitting the text.  Large buffers therefore go back to a small pool
 * and are handed out again (first fit within 2x), up to a fixed total. */
#include <pthread.h>
#include <time.h>
#define BIG_MIN   ((size_t)1 << 20)
#define BIG_SLOTS 96
#define BIG_POOL_MAX ((size_t)12 << 30)
typedef struct { size_t cap, pooled; char pad[48]; } big_hdr_t; /* 64 bytes in front of every buffer */
static big_hdr_t *big_pool[BIG_SLOTS];
static size_t big_pool_bytes;
static pthread_mutex_t big_mu = PTHREAD_MUTEX_INITIALIZER;

static void *big_alloc(size_t n)
{
	big_hdr_t *h = 0;
	if (n >= BIG_MIN) {
		int i, best = -1;
		pthread_mutex_lock(&big_mu);
		for (i = 0; i < BIG_SLOTS; ++i)
			if (big_pool[i] && big_pool[i]->cap >= n && big_pool[i]->cap <= 2 * n + BIG_MIN && (best < 0 || big_pool[i]->cap < big_pool[best]->cap)) best = i;
		if (best >= 0) h = big_pool[best], big_pool[best] = 0, big_pool_bytes -= h->cap;
		pthread_mutex_unlock(&big_mu);
	}
	if (h == 0) {
		const size_t cap = n >= BIG_MIN ? n + n / 8 : n;
		h = (big_hdr_t*)malloc(sizeof(big_hdr_t) + cap + 1);
		if (h == 0) return 0;
		h->cap = cap, h->pooled = n >= BIG_MIN;
	}
	return h + 1;
}

static void big_free(void *p)
{
	big_hdr_t *h;
	if (p == 0) return;
	h = (big_hdr_t*)p - 1;
	if (h->pooled) {
		int i, done = 0;
		pthread_mutex_lock(&big_mu);
		if (big_pool_bytes + h->cap <= BIG_POOL_MAX)
			for (i = 0; i < BIG_SLOTS && !done; ++i)
				if (big_pool[i] == 0) big_pool[i] = h, big_pool_bytes += h->cap, done = 1;
		pthread_mutex_unlock(&big_mu);
		if (done) return;
	}
	free(h);
}

#define PREAD_PIECE_DEFAULT ((size_t)2 << 20) /* BFC_B200_READ_PIECE=<bytes> (tests): the grain of the parallel read and split */

struct fq_reader_s {
	gzFile fp;
	int fd;                     /* >= 0: an uncompressed regular file, read with parallel pread()s; the gzFile then only
	                               serves the hand-over to the tolerant parser (repositioned with gzseek) */
	int64_t pos, size;
	int n_threads, eof, fast;
	size_t piece;               /* bytes per work item of the parallel read and split */
	char *carry;                /* text read but not yet delivered (the incomplete tail of the previous block) */
	size_t carry_len;
	char *last_comment;         /* kseq's sticky comment (bseq.c header) */
	bseq_file_t *slow;          /* the tolerant parser, once the input stopped being plain four-line FASTQ */
};

fq_reader_t *fq_open(const char *fn, int n_threads)
{
	gzFile g = fn && strcmp(fn, "-") ? gzopen(fn, "r") : gzdopen(fileno(stdin), "r");
	fq_reader_t *r;
	if (g == 0) return 0;
	gzbuffer(g, 1 << 20);
	r = (fq_reader_t*)calloc(1, sizeof(fq_reader_t));
	r->fp = g, r->n_threads = n_threads < 1 ? 1 : n_threads, r->fast = 1, r->fd = -1;
	r->piece = getenv("BFC_B200_READ_PIECE") && atoll(getenv("BFC_B200_READ_PIECE")) >= 16 ? (size_t)atoll(getenv("BFC_B200_READ_PIECE")) : PREAD_PIECE_DEFAULT;
	if (fn && strcmp(fn, "-")) { /* plain regular file? (zlib would copy it through its own buffer on one thread) */
		struct stat st;
		unsigned char magic[2] = {0, 0};
		const int fd = open(fn, O_RDONLY);
		if (fd >= 0 && fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && !(pread(fd, magic, 2, 0) == 2 && magic[0] == 0x1f && magic[1] == 0x8b))
			r->fd = fd, r->size = (int64_t)st.st_size;
		else if (fd >= 0) close(fd);
	}
	return r;
}

typedef struct { int fd; char *dst; int64_t pos, len, piece; int err; size_t *nl_cnt; } pread_t;

static size_t count_newlines(const char *p, const char *e)
{
	size_t c = 0;
	for (; p < e && (p = (const char*)memchr(p, '\n', (size_t)(e - p))) != 0; ++p) ++c;
	return c;
}

static void pread_worker(void *data, long i, int tid)
{
	pread_t *p = (pread_t*)data;
	const int64_t o0 = i * p->piece;
	int64_t o = o0, e = o + p->piece < p->len ? o + p->piece : p->len;
	(void)tid;
	while (o < e) {
		const ssize_t got = pread(p->fd, p->dst + o, (size_t)(e - o), (off_t)(p->pos + o));
		if (got <= 0) { p->err = 1; return; }
		o += got;
	}
	if (p->nl_cnt) p->nl_cnt[i] = count_newlines(p->dst + o0, p->dst + e); /* the piece is still in this core's cache */
}

/* up to `want` more bytes of input at dst; 0 at end of input.  nl_cnt (optional, room for a count per rd->piece bytes of
 * `want`): filled with the newlines of every piece when the bytes came from parallel pread()s, *n_pieces says how many */
static size_t rd_more(fq_reader_t *rd, char *dst, size_t want, size_t *nl_cnt, long *n_pieces)
{
	if (n_pieces) *n_pieces = 0;
	if (rd->fd >= 0) {
		pread_t p;
		const long np = (long)(((rd->size - rd->pos < (int64_t)want ? rd->size - rd->pos : (int64_t)want) + (int64_t)rd->piece - 1) / (int64_t)rd->piece);
		p.piece = (int64_t)rd->piece;
		p.fd = rd->fd, p.dst = dst, p.pos = rd->pos, p.err = 0, p.nl_cnt = nl_cnt;
		p.len = rd->size - rd->pos < (int64_t)want ? rd->size - rd->pos : (int64_t)want;
		if (p.len <= 0) return 0;
		kt_for(rd->n_threads, pread_worker, &p, np);
		if (p.err) return 0;
		rd->pos += p.len;
		if (n_pieces && nl_cnt) *n_pieces = np;
		return (size_t)p.len;
	} else {
		const int got = gzread(rd->fp, dst, (unsigned)(want > (1u << 30) ? (1u << 30) : want));
		return got > 0 ? (size_t)got : 0;
	}
}

void fq_close(fq_reader_t *r)
{
	if (r == 0) return;
	if (r->slow) bseq_close(r->slow); /* owns the stream by now */
	else gzclose(r->fp);
	if (r->fd >= 0) close(r->fd);
	free(r->carry); free(r->last_comment);
	free(r);
}

int fq_reader_is_fast(const fq_reader_t *r) { return r->fast; }

void fq_block_free(fq_block_t *b)
{
	big_free(b->buf); big_free(b->name_off); big_free(b->com_off); big_free(b->seq_off); big_free(b->qual_off);
	big_free(b->name_len); big_free(b->com_len); big_free(b->seq_len);
	memset(b, 0, sizeof(*b));
}

static int blk_alloc(fq_block_t *b, int64_t n)
{
	const size_t m = n > 0 ? (size_t)n : 1;
	b->n = n;
	b->name_off = (uint64_t*)big_alloc(m * 8), b->com_off = (uint64_t*)big_alloc(m * 8);
	b->seq_off = (uint64_t*)big_alloc(m * 8), b->qual_off = (uint64_t*)big_alloc(m * 8);
	b->name_len = (uint32_t*)big_alloc(m * 4), b->com_len = (uint32_t*)big_alloc(m * 4), b->seq_len = (uint32_t*)big_alloc(m * 4);
	return b->name_off && b->com_off && b->seq_off && b->qual_off && b->name_len && b->com_len && b->seq_len ? 0 : -1;
}

/* ---------------------------------------------------------------- the parallel path */

typedef struct {
	const char *s;
	size_t len;
	int virt_nl;        /* the input ended without a newline: one is imagined at s[len] */
	long n_parts;
	size_t *lo;         /* part i = [lo[i], lo[i + 1]) */
	size_t *cnt;        /* newlines per part, then their exclusive prefix = the index of the line that holds byte lo[i] */
	fq_block_t *b;
	int keep_comment, bad;
	size_t used;        /* end of the last complete record */
	uint64_t *part_bases;
	int64_t *part_last_com; /* the last record of the part that has a comment (-1: none), and where that comment is */
	uint64_t *part_com_off;
	uint32_t *part_com_len;
} split_t;

static void count_nl_worker(void *data, long i, int tid)
{
	split_t *sp = (split_t*)data;
	(void)tid;
	sp->cnt[i] = count_newlines(sp->s + sp->lo[i], sp->s + sp->lo[i + 1]);
}

/* Part i delivers the records whose '@' line starts inside it (reading on into the next part for the rest of its last
 * record).  Which line a byte belongs to follows from the newline counts alone, so no index of line starts is built:
 * one pass over the text.  Anything kseq would read differently flags the block. */
static void parse_worker(void *data, long i, int tid)
{
	split_t *sp = (split_t*)data;
	fq_block_t *b = sp->b;
	const char *s = sp->s, *q;
	const size_t len = sp->len, hi = sp->lo[i + 1];
	size_t p = sp->lo[i];
	uint64_t line = sp->cnt[i], bases = 0, last_off = 0;
	uint32_t last_len = 0;
	int64_t last_com = -1;
	(void)tid;
	sp->part_bases[i] = 0, sp->part_last_com[i] = -1;
	if (p > 0 && s[p - 1] != '\n') { /* the line that holds lo[i] began in an earlier part */
		if ((q = (const char*)memchr(s + p, '\n', len - p)) == 0) return;
		p = (size_t)(q - s) + 1, ++line;
	}
	for (; line & 3; ++line) { /* on to the next '@' line */
		if (p >= hi || (q = (const char*)memchr(s + p, '\n', len - p)) == 0) return;
		p = (size_t)(q - s) + 1;
	}
	for (; p < hi && (int64_t)(line >> 2) < b->n; line += 4) {
		const int64_t r = (int64_t)(line >> 2);
		const uint64_t l0 = p;
		uint64_t e[4], c;
		int j;
		for (j = 0; j < 4; ++j) {
			q = p < len ? (const char*)memchr(s + p, '\n', len - p) : 0;
			if (q == 0 && !(sp->virt_nl && j == 3)) { sp->bad = 1; return; } /* (cannot happen: r < n counts whole records) */
			e[j] = q ? (uint64_t)(q - s) : len;
			p = (size_t)e[j] + 1;
		}
		{
			const uint64_t l1 = e[0] + 1, l2 = e[1] + 1, l3 = e[2] + 1;
			if (e[0] == l0 || s[l0] != '@' || e[1] == l1 || s[l2] != '+' || e[2] == l2 || e[3] - l3 != e[1] - l1 ||
				s[l1] == '>' || s[l1] == '+' || s[l1] == '@' || s[e[0] - 1] == '\r' || s[e[1] - 1] == '\r' || s[e[2] - 1] == '\r' || s[e[3] - 1] == '\r' ||
				e[1] - l1 > 0x7fffffffULL) { sp->bad = 1; return; }
			for (c = l0 + 1; c < e[0] && !isspace((unsigned char)s[c]); ++c) {}
			b->name_off[r] = l0 + 1, b->name_len[r] = (uint32_t)(c - (l0 + 1));
			b->com_off[r] = FQ_NONE, b->com_len[r] = 0;
			if (c < e[0]) { /* the rest of the line after ONE delimiter */
				last_com = r, last_off = c + 1, last_len = (uint32_t)(e[0] - (c + 1));
				if (sp->keep_comment) b->com_off[r] = last_off, b->com_len[r] = last_len;
			}
			b->seq_off[r] = l1, b->seq_len[r] = (uint32_t)(e[1] - l1), b->qual_off[r] = l3;
			bases += e[1] - l1;
			if (r == b->n - 1) sp->used = e[3] + 1 > len ? len : (size_t)e[3] + 1;
		}
	}
	sp->part_bases[i] = bases, sp->part_last_com[i] = last_com, sp->part_com_off[i] = last_off, sp->part_com_len[i] = last_len;
}

/* kseq never clears its comment buffer: a record without one inherits the latest (bseq.c header).  `last` = the last
 * record of the block that carries a comment of its own (-1: none) */
static void sticky_comments(fq_reader_t *rd, fq_block_t *b, int keep_comment, int64_t last, uint64_t last_off, uint32_t last_len)
{
	int64_t r;
	if (keep_comment) {
		int64_t seen = -1;
		for (r = 0; r < b->n; ++r) {
			if (b->com_off[r] != FQ_NONE) seen = r;
			else if (seen >= 0) b->com_off[r] = b->com_off[seen], b->com_len[r] = b->com_len[seen];
			else if (rd->last_comment) b->com_off[r] = FQ_NONE - 1; /* patched below: lives outside this block */
		}
	}
	if (keep_comment && rd->last_comment) { /* comments inherited from an earlier block: append the text to the block */
		const size_t l = strlen(rd->last_comment);
		int any = 0;
		for (r = 0; r < b->n && b->com_off[r] == FQ_NONE - 1; ++r) any = 1;
		if (any) {
			char *nb = (char*)big_alloc(b->buf_len + l + 1); /* (rare: one block in a file whose first records lack comments) */
			memcpy(nb, b->buf, b->buf_len);
			big_free(b->buf);
			b->buf = nb;
			memcpy(b->buf + b->buf_len, rd->last_comment, l + 1);
			for (r = 0; r < b->n && b->com_off[r] == FQ_NONE - 1; ++r) b->com_off[r] = b->buf_len, b->com_len[r] = (uint32_t)l;
			b->buf_len += l + 1;
		}
	}
	if (last >= 0) {
		free(rd->last_comment);
		rd->last_comment = (char*)malloc((size_t)last_len + 1);
		memcpy(rd->last_comment, b->buf + last_off, last_len);
		rd->last_comment[last_len] = 0;
	}
}

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
extern int bfc_verbose;

/* 1 = block delivered, 0 = not plain four-line FASTQ (nothing consumed; *data is handed back), -1 = out of memory.
 * pre_cnt/pre_n: newline counts the reader already took of the pieces [pre_lo + j * rd->piece, ...) */
static int fast_block(fq_reader_t *rd, char *data, size_t len, int keep_comment, fq_block_t *b, size_t *used,
                      const size_t *pre_cnt, long pre_n, size_t pre_lo)
{
	split_t sp;
	size_t run = 0, n_lines;
	long i;
	int64_t last = -1;
	uint64_t last_off = 0;
	uint32_t last_len = 0;
	double t[4];
	memset(&sp, 0, sizeof(sp));
	if (len == 0 || data[0] != '@') return 0;
	t[0] = now_s();
	sp.s = data, sp.len = len, sp.keep_comment = keep_comment;
	sp.virt_nl = rd->eof && data[len - 1] != '\n'; /* last line without a newline */
	const size_t piece = rd->piece;
	if (pre_n > 0 && pre_lo + (size_t)pre_n * piece >= len && pre_lo + (size_t)(pre_n - 1) * piece < len) {
		/* parts = the carried-over head, then the pieces as they were read (and counted) */
		sp.n_parts = pre_n + (pre_lo > 0);
		sp.lo = (size_t*)malloc(((size_t)sp.n_parts + 1) * sizeof(size_t)), sp.cnt = (size_t*)calloc((size_t)sp.n_parts + 1, sizeof(size_t));
		if (pre_lo > 0) sp.lo[0] = 0, sp.cnt[0] = count_newlines(data, data + pre_lo);
		for (i = 0; i < pre_n; ++i) sp.lo[i + (pre_lo > 0)] = pre_lo + (size_t)i * piece, sp.cnt[i + (pre_lo > 0)] = pre_cnt[i];
		sp.lo[sp.n_parts] = len;
	} else {
		sp.n_parts = (long)((len + piece - 1) / piece);
		sp.lo = (size_t*)malloc(((size_t)sp.n_parts + 1) * sizeof(size_t)), sp.cnt = (size_t*)calloc((size_t)sp.n_parts + 1, sizeof(size_t));
		for (i = 0; i < sp.n_parts; ++i) sp.lo[i] = (size_t)i * piece;
		sp.lo[sp.n_parts] = len;
		kt_for(rd->n_threads, count_nl_worker, &sp, sp.n_parts);
	}
	for (i = 0; i < sp.n_parts; ++i) { const size_t c = sp.cnt[i]; sp.cnt[i] = run; run += c; }
	n_lines = run + (size_t)sp.virt_nl;
	t[1] = now_s();
	if (n_lines < 4 || (rd->eof && n_lines % 4 != 0)) { free(sp.cnt); free(sp.lo); return 0; }
	memset(b, 0, sizeof(*b));
	sp.part_bases = (uint64_t*)calloc((size_t)sp.n_parts, 8), sp.part_last_com = (int64_t*)calloc((size_t)sp.n_parts, 8);
	sp.part_com_off = (uint64_t*)calloc((size_t)sp.n_parts, 8), sp.part_com_len = (uint32_t*)calloc((size_t)sp.n_parts, 4);
	if (blk_alloc(b, (int64_t)(n_lines / 4)) < 0) {
		free(sp.cnt); free(sp.lo); free(sp.part_bases); free(sp.part_last_com); free(sp.part_com_off); free(sp.part_com_len);
		fq_block_free(b);
		return -1;
	}
	sp.b = b;
	kt_for(rd->n_threads, parse_worker, &sp, sp.n_parts);
	t[2] = now_s();
	for (i = 0; i < sp.n_parts; ++i) {
		b->n_bases += sp.part_bases[i];
		if (sp.part_last_com[i] > last) last = sp.part_last_com[i], last_off = sp.part_com_off[i], last_len = sp.part_com_len[i];
	}
	*used = sp.used;
	free(sp.cnt); free(sp.lo); free(sp.part_bases); free(sp.part_last_com); free(sp.part_com_off); free(sp.part_com_len);
	if (sp.bad || sp.used == 0) { fq_block_free(b); return 0; }
	b->buf = data, b->buf_len = len, b->any_qual = 1;
	sticky_comments(rd, b, keep_comment, last, last_off, last_len);
	t[3] = now_s();
	if (bfc_verbose >= 5)
		fprintf(stderr, "[D::%s] newline counts %.1f ms, records %.1f, rest %.1f\n", __func__,
				1e3 * (t[1] - t[0]), 1e3 * (t[2] - t[1]), 1e3 * (t[3] - t[2]));
	return 1;
}

/* ---------------------------------------------------------------- the tolerant path */

static int slow_block(fq_reader_t *rd, size_t target, int keep_comment, fq_block_t *b)
{
	int n = 0, i;
	const long chunk = target > 0x7fffffffUL / 2 ? 0x7fffffff / 2 : (long)(target / 2 + 1); /* bases ~ half of the text */
	bseq1_t *seqs = 0;
	size_t tot = 0, at = 0;
	memset(b, 0, sizeof(*b));
	/* bseq_read() ends a batch at a record whose quality length is wrong (kseq's -2) and goes on behind it on the next
	 * call; a batch that is empty for that reason is not the end of the input (the reference, whose batches are 100 M
	 * bases, would only stop there if the bad record happened to be the first of a batch) */
	for (;;) {
		seqs = bseq_read(rd->slow, (int)chunk, keep_comment, &n);
		if (seqs && n > 0) break;
		free(seqs);
		if (bseq_at_eof(rd->slow)) return 0;
	}
	for (i = 0; i < n; ++i)
		tot += strlen(seqs[i].name) + 1 + (seqs[i].comment ? strlen(seqs[i].comment) + 1 : 0) + (size_t)seqs[i].l_seq * (seqs[i].qual ? 2 : 1) + 2;
	if (blk_alloc(b, n) < 0 || (b->buf = (char*)big_alloc(tot + 1)) == 0) { fq_block_free(b); return -1; }
	for (i = 0; i < n; ++i) {
		bseq1_t *s = &seqs[i];
		size_t l = strlen(s->name);
		memcpy(b->buf + at, s->name, l + 1); b->name_off[i] = at, b->name_len[i] = (uint32_t)l; at += l + 1;
		if (s->comment) { l = strlen(s->comment); memcpy(b->buf + at, s->comment, l + 1); b->com_off[i] = at, b->com_len[i] = (uint32_t)l; at += l + 1; }
		else b->com_off[i] = FQ_NONE, b->com_len[i] = 0;
		memcpy(b->buf + at, s->seq, (size_t)s->l_seq + 1); b->seq_off[i] = at, b->seq_len[i] = (uint32_t)s->l_seq; at += (size_t)s->l_seq + 1;
		if (s->qual) { memcpy(b->buf + at, s->qual, (size_t)s->l_seq + 1); b->qual_off[i] = at; at += (size_t)s->l_seq + 1; b->any_qual = 1; }
		else b->qual_off[i] = FQ_NONE;
		b->n_bases += (uint64_t)s->l_seq;
		free(s->name); free(s->comment); free(s->seq); free(s->qual);
	}
	b->buf_len = at;
	free(seqs);
	return 1;
}

int fq_next(fq_reader_t *rd, size_t target, int keep_comment, fq_block_t *b)
{
	memset(b, 0, sizeof(*b));
	if (target < 4096) target = 4096;
	while (rd->fast) {
		const size_t carry0 = rd->carry_len;
		const double t0 = now_s();
		char *data = (char*)big_alloc(carry0 + target + 1);
		size_t len = carry0, used = 0;
		int rc;
		if (data == 0) return -1;
		if (carry0) memcpy(data, rd->carry, carry0);
		free(rd->carry); rd->carry = 0, rd->carry_len = 0;
		size_t *pre_cnt = (size_t*)malloc((target / rd->piece + 2) * sizeof(size_t));
		long pre_n = 0, calls = 0;
		while (!rd->eof && len < carry0 + target) {
			long np = 0;
			const size_t got = rd_more(rd, data + len, carry0 + target - len, calls == 0 ? pre_cnt : 0, &np);
			if (got == 0) { rd->eof = 1; break; }
			pre_n = calls++ == 0 ? np : 0; /* counts are only good when one call brought everything */
			len += got;
		}
		if (len == 0) { big_free(data); free(pre_cnt); return 0; }
		if (bfc_verbose >= 5) fprintf(stderr, "[D::%s] %.1f MB of text in %.1f ms\n", __func__, len / 1e6, 1e3 * (now_s() - t0));
		rc = fast_block(rd, data, len, keep_comment, b, &used, pre_cnt, pre_n, carry0);
		free(pre_cnt);
		if (rc == 1) {
			if (used < len) { /* the incomplete tail waits for the next block */
				rd->carry_len = len - used;
				rd->carry = (char*)malloc(rd->carry_len);
				memcpy(rd->carry, b->buf + used, rd->carry_len); /* (b->buf: sticky_comments may have moved the text) */
			}
			return 1;
		}
		if (rc < 0) { big_free(data); return -1; }
		/* not plain four-line FASTQ: the tolerant parser takes the stream over, starting with this block's text */
		rd->fast = 0;
		if (rd->fd >= 0) gzseek(rd->fp, (z_off_t)rd->pos, SEEK_SET); /* the stream goes on where the pread()s stopped */
		{
			unsigned char *pre = (unsigned char*)malloc(len + 1); /* (bseq.c frees it with free()) */
			if (pre == 0) { big_free(data); return -1; }
			memcpy(pre, data, len);
			big_free(data);
			rd->slow = bseq_open_from(rd->fp, pre, len, rd->last_comment);
		}
	}
	return slow_block(rd, target, keep_comment, b); /* 1, 0 at end of input, -1 out of memory */
}

/* ---------------------------------------------------------------- blocks kept from one phase to the next */

/* bfc reads its input twice (count, then correct: reference bfc.c:131-148).  For gzip'd input, where the second pass
 * would inflate everything again on one thread, the parsed blocks of the first pass are kept when they fit into a
 * quarter of the host's memory, and the second pass takes them from here.  For plain files this is off by default:
 * on the B200 box (16 cores) keeping the blocks saves the correct pass 4 ms per 250 MB block and costs the count pass
 * 8 ms, because kept buffers cannot be recycled and every new one is paid for in page faults.
 * BFC_B200_KEEP_MAX=<bytes> switches it on for every regular file (0: off for all). */
static struct {
	char *fn;
	off_t size;
	struct timespec mtime;
	int keep_comment, complete, over;
	fq_block_t *blk;
	long n, m;
	size_t bytes, budget;
} kept;

static size_t block_bytes(const fq_block_t *b) { return b->buf_len + (size_t)(b->n > 0 ? b->n : 1) * 44; }

void fq_keep_drop(void)
{
	long i;
	for (i = 0; i < kept.n; ++i) fq_block_free(&kept.blk[i]);
	free(kept.blk); free(kept.fn);
	memset(&kept, 0, sizeof(kept));
}

void fq_keep_begin(const char *fn, int keep_comment)
{
	struct stat st;
	const char *e = getenv("BFC_B200_KEEP_MAX"); /* bytes; 0 switches the cache off */
	fq_keep_drop();
	if (fn == 0 || strcmp(fn, "-") == 0 || stat(fn, &st) != 0 || !S_ISREG(st.st_mode)) return;
	if (e) kept.budget = (size_t)strtoull(e, 0, 10);
	else {
		unsigned char magic[2] = {0, 0};
		FILE *fp = fopen(fn, "rb");
		const int gz = fp && fread(magic, 1, 2, fp) == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
		if (fp) fclose(fp);
		kept.budget = gz ? (size_t)sysconf(_SC_PHYS_PAGES) / 4 * (size_t)sysconf(_SC_PAGESIZE) : 0;
	}
	if (kept.budget == 0) return;
	kept.fn = strdup(fn), kept.size = st.st_size, kept.mtime = st.st_mtim, kept.keep_comment = keep_comment;
}

int fq_keep_add(fq_block_t *b)
{
	if (kept.fn == 0 || kept.over) return 0;
	if (kept.bytes + block_bytes(b) > kept.budget) { /* too big for this host: the second pass reads the file again */
		long i;
		for (i = 0; i < kept.n; ++i) fq_block_free(&kept.blk[i]);
		kept.n = 0, kept.bytes = 0, kept.over = 1;
		return 0;
	}
	if (kept.n == kept.m) {
		fq_block_t *nb = (fq_block_t*)realloc(kept.blk, (size_t)(kept.m ? kept.m * 2 : 64) * sizeof(fq_block_t));
		if (nb == 0) { kept.over = 1; return 0; }
		kept.blk = nb, kept.m = kept.m ? kept.m * 2 : 64;
	}
	kept.bytes += block_bytes(b);
	kept.blk[kept.n++] = *b;
	memset(b, 0, sizeof(*b));
	return 1;
}

void fq_keep_end(int complete) { kept.complete = complete && kept.fn && !kept.over; }

long fq_keep_match(const char *fn, int keep_comment)
{
	struct stat st;
	if (!kept.complete || fn == 0 || strcmp(fn, kept.fn) != 0 || keep_comment != kept.keep_comment) return -1;
	if (stat(fn, &st) != 0 || st.st_size != kept.size || st.st_mtim.tv_sec != kept.mtime.tv_sec || st.st_mtim.tv_nsec != kept.mtime.tv_nsec) return -1;
	return kept.n;
}

int fq_keep_take(long i, fq_block_t *b)
{
	if (i < 0 || i >= kept.n) return 0;
	*b = kept.blk[i];
	memset(&kept.blk[i], 0, sizeof(fq_block_t));
	return 1;
}

/* ---------------------------------------------------------------- flat batch */

#define REC_PER_ITEM 4096

typedef struct { fq_flat_t *f; const fq_block_t *b; } fill_t;

static void fill_worker(void *data, long i, int tid)
{
	fill_t *ft = (fill_t*)data;
	const fq_block_t *b = ft->b;
	fq_flat_t *f = ft->f;
	int64_t r, r1 = (i + 1) * (int64_t)REC_PER_ITEM < b->n ? (i + 1) * (int64_t)REC_PER_ITEM : b->n;
	(void)tid;
	for (r = i * (int64_t)REC_PER_ITEM; r < r1; ++r) {
		const uint32_t l = b->seq_len[r];
		uint64_t o;
		if (f->flat_idx[r] < 0) continue;
		o = f->off[f->flat_idx[r]];
		memcpy(f->b.seq + o, b->buf + b->seq_off[r], l);
		f->b.seq[o + l] = 0;
		if (f->b.qual) {
			if (b->qual_off[r] != FQ_NONE) memcpy(f->b.qual + o, b->buf + b->qual_off[r], l);
			else memset(f->b.qual + o, 0xFF, l); /* "no quality" marker (bfc_b200.h) */
			f->b.qual[o + l] = 0;
		}
	}
}

/* pinning host memory costs ~0.4 s per GB (cudaMallocHost, B200 box): buffers released by one phase are parked here for
 * the next one.  (Pinning them ahead on a helper thread while the filter is being created was measured and dropped:
 * the driver serialises it with the context's own allocations, the first batch came out 0.3 s later, not earlier.) */
#define FLAT_CACHE 4
static struct { uint8_t *seq, *qual; size_t cap; } flat_cache[FLAT_CACHE];
static pthread_mutex_t flat_mu = PTHREAD_MUTEX_INITIALIZER;

static int flat_cache_take(fq_flat_t *f, size_t need)
{
	int i, got = 0;
	pthread_mutex_lock(&flat_mu);
	for (i = 0; i < FLAT_CACHE && !got; ++i)
		if (flat_cache[i].seq && flat_cache[i].cap >= need) {
			f->seq_buf = flat_cache[i].seq, f->qual_buf = flat_cache[i].qual, f->cap_bytes = flat_cache[i].cap, f->pinned = 1;
			flat_cache[i].seq = 0;
			got = 1;
		}
	pthread_mutex_unlock(&flat_mu);
	return got;
}

static int flat_cache_put(uint8_t *seq, uint8_t *qual, size_t cap)
{
	int i, put = 0;
	pthread_mutex_lock(&flat_mu);
	for (i = 0; i < FLAT_CACHE && !put; ++i)
		if (flat_cache[i].seq == 0) flat_cache[i].seq = seq, flat_cache[i].qual = qual, flat_cache[i].cap = cap, put = 1;
	pthread_mutex_unlock(&flat_mu);
	return put;
}

int fq_flat_fill(fq_flat_t *f, const fq_block_t *b, const uint8_t *skip, int n_threads)
{
	const size_t need = (size_t)(b->n_bases + (uint64_t)b->n) + 1;
	int64_t r, m = 0;
	int any_qual = 0;
	uint64_t tot = 0;
	fill_t ft;
	if (need > f->cap_bytes) { /* grow-only, reused from batch to batch: pinning memory is expensive */
		fq_flat_t old = *f;
		if (!flat_cache_take(f, need)) {
			f->cap_bytes = need + need / 8;
			f->seq_buf = (uint8_t*)bfcg_host_alloc_pinned(f->cap_bytes);
			f->qual_buf = f->seq_buf ? (uint8_t*)bfcg_host_alloc_pinned(f->cap_bytes) : 0;
			f->pinned = f->seq_buf && f->qual_buf;
			if (!f->pinned) { /* pageable memory works too: the copies just do not overlap */
				if (f->seq_buf) bfcg_host_free_pinned(f->seq_buf);
				f->seq_buf = (uint8_t*)malloc(f->cap_bytes), f->qual_buf = (uint8_t*)malloc(f->cap_bytes);
			}
		}
		if (old.pinned) { bfcg_host_free_pinned(old.seq_buf); bfcg_host_free_pinned(old.qual_buf); }
		else { free(old.seq_buf); free(old.qual_buf); }
		if (f->seq_buf == 0 || f->qual_buf == 0) return -1;
	}
	if ((size_t)b->n + 1 > f->cap_reads) {
		free(f->off); free(f->flat_idx);
		f->cap_reads = (size_t)b->n + 1 + (size_t)b->n / 8;
		f->off = (uint64_t*)malloc(f->cap_reads * 8);
		f->flat_idx = (int64_t*)malloc(f->cap_reads * 8);
		if (f->off == 0 || f->flat_idx == 0) return -1;
	}
	for (r = 0; r < b->n; ++r) {
		if (skip && skip[r]) { f->flat_idx[r] = -1; continue; }
		f->flat_idx[r] = m, f->off[m++] = tot;
		tot += (uint64_t)b->seq_len[r] + 1;
		any_qual |= b->qual_off[r] != FQ_NONE;
	}
	f->off[m] = tot;
	f->b.n_reads = m, f->b.n_bytes = tot, f->b.where = BFCG_HOST, f->b.off = f->off;
	f->b.seq = f->seq_buf, f->b.qual = any_qual ? f->qual_buf : 0;
	ft.f = f, ft.b = b;
	kt_for(n_threads, fill_worker, &ft, (long)((b->n + REC_PER_ITEM - 1) / REC_PER_ITEM));
	return 0;
}

void fq_flat_free(fq_flat_t *f)
{
	if (f->pinned && f->seq_buf && flat_cache_put(f->seq_buf, f->qual_buf, f->cap_bytes)) {}
	else if (f->pinned) { bfcg_host_free_pinned(f->seq_buf); bfcg_host_free_pinned(f->qual_buf); }
	else { free(f->seq_buf); free(f->qual_buf); }
	free(f->off); free(f->flat_idx);
	memset(f, 0, sizeof(*f));
}

/* ---------------------------------------------------------------- writer */

typedef struct {
	const fq_block_t *b;
	const fq_flat_t *flat;
	const fq_out_t *o;
	int n_items;
	char **piece;
	size_t *piece_len;
	int oom;
} wr_t;

static inline char *put_uint(char *p, unsigned v)
{
	char t[12];
	int n = 0;
	do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v);
	while (n) *p++ = t[--n];
	return p;
}

static void write_worker(void *data, long i, int tid)
{
	wr_t *w = (wr_t*)data;
	const fq_block_t *b = w->b;
	const fq_out_t *o = w->o;
	const int64_t per = (b->n + w->n_items - 1) / w->n_items;
	const int64_t r0 = i * per, r1 = r0 + per < b->n ? r0 + per : b->n;
	int64_t r;
	size_t cap = 16;
	char *out, *p;
	(void)tid;
	for (r = r0; r < r1; ++r) cap += (size_t)b->name_len[r] + b->com_len[r] + 2 * (size_t)b->seq_len[r] + 96;
	out = p = (char*)big_alloc(cap);
	if (out == 0) { w->oom = 1; w->piece[i] = 0, w->piece_len[i] = 0; return; }
	for (r = r0; r < r1; ++r) {
		const int64_t j = w->flat->flat_idx[r]; /* index in the batch; < 0: the record was left out of it */
		const uint64_t fo = j >= 0 ? w->flat->b.off[j] : 0;
		const int has_qual = b->qual_off[r] != FQ_NONE && (j < 0 || w->flat->b.qual != 0);
		const int is_fq = has_qual && !o->no_qual;
		const uint8_t *seq = j >= 0 ? w->flat->b.seq + fo : (const uint8_t*)b->buf + b->seq_off[r];
		const uint8_t *qual = !has_qual ? 0 : j >= 0 ? w->flat->b.qual + fo : (const uint8_t*)b->buf + b->qual_off[r];
		uint32_t l = b->seq_len[r];
		if (j < 0) { /* -R: untouched, comment kept (correct.c:544-545, 602) */
			*p++ = is_fq ? '@' : '>';
			memcpy(p, b->buf + b->name_off[r], b->name_len[r]); p += b->name_len[r];
			if (b->com_off[r] != FQ_NONE) { *p++ = '\t'; memcpy(p, b->buf + b->com_off[r], b->com_len[r]); p += b->com_len[r]; }
		} else if (!o->filter_mode) { /* correct.c:596-603 */
			const uint32_t aux = o->aux[2 * j], aux2 = o->aux[2 * j + 1];
			if (o->discard && (aux & 7)) continue;
			*p++ = is_fq ? '@' : '>';
			memcpy(p, b->buf + b->name_off[r], b->name_len[r]); p += b->name_len[r];
			if (o->refine || b->com_off[r] == FQ_NONE) {
				memcpy(p, "\tec:Z:", 6); p += 6;
				p = put_uint(p, aux & 7);
				if ((aux & 7) == 0) {
					*p++ = '_'; p = put_uint(p, aux2 >> 10); *p++ = ':'; p = put_uint(p, aux2 & 0xff);
					*p++ = '_'; p = put_uint(p, aux >> 3 & 1);
					*p++ = '_'; p = put_uint(p, aux >> 18 & 0x3fff); *p++ = ':'; p = put_uint(p, aux >> 4 & 0x3fff);
					*p++ = '_'; p = put_uint(p, aux2 >> 8 & 3);
				}
			} else { *p++ = '\t'; memcpy(p, b->buf + b->com_off[r], b->com_len[r]); p += b->com_len[r]; }
		} else { /* correct.c:604-608; the kept stretch as worker_ec's memmove leaves it (correct.c:557-567) */
			if (!o->keep[r]) continue;
			*p++ = is_fq ? '@' : '>';
			memcpy(p, b->buf + b->name_off[r], b->name_len[r]); p += b->name_len[r];
			if (b->com_off[r] != FQ_NONE) { *p++ = '\t'; memcpy(p, b->buf + b->com_off[r], b->com_len[r]); p += b->com_len[r]; }
			seq += o->tstart[r];
			if (qual) qual += o->tstart[r];
			l = (uint32_t)(o->tend[r] - o->tstart[r]);
		}
		*p++ = '\n';
		memcpy(p, seq, l); p += l; *p++ = '\n';
		if (is_fq) { *p++ = '+'; *p++ = '\n'; memcpy(p, qual, l); p += l; *p++ = '\n'; }
	}
	w->piece[i] = out, w->piece_len[i] = (size_t)(p - out);
}

int fq_write(FILE *fp, const fq_block_t *b, const fq_flat_t *flat, const fq_out_t *o, int n_threads)
{
	wr_t w;
	int i, rc = 0;
	if (b->n == 0) return 0;
	memset(&w, 0, sizeof(w));
	w.b = b, w.flat = flat, w.o = o;
	w.n_items = n_threads < 1 ? 1 : n_threads * 4;
	if ((int64_t)w.n_items > b->n) w.n_items = (int)b->n;
	w.piece = (char**)calloc((size_t)w.n_items, sizeof(char*));
	w.piece_len = (size_t*)calloc((size_t)w.n_items, sizeof(size_t));
	kt_for(n_threads, write_worker, &w, w.n_items);
	for (i = 0; i < w.n_items; ++i) {
		if (!w.oom && w.piece_len[i] && fwrite(w.piece[i], 1, w.piece_len[i], fp) != w.piece_len[i]) rc = -1;
		big_free(w.piece[i]);
	}
	free(w.piece); free(w.piece_len);
	return w.oom ? -1 : rc;
}
