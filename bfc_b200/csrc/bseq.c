/* bseq.c -- FASTA/FASTQ batch reader (interface: reference bseq.h:9-19).
 *
 * Accepts what the reference's reader (bseq.c:52-76 over kseq.h:185-224) accepts and
 * yields the same records: multi-line FASTA/FASTQ, "\r\n" line ends, blank lines,
 * mixed FASTA and FASTQ records, gzip or plain input, "-" for stdin.  Two quirks of
 * that reader are kept because they are visible in the output of `-1` mode:
 *   - a record without a comment inherits the most recent comment seen in the file
 *     (kseq never clears comment.s; bseq.c:66 copies whatever is there);
 *   - reading stops at the first record whose quality length differs from its
 *     sequence length (kseq_read returns -2, bseq.c:58).
 */
#include <zlib.h>
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bseq.h"

/* A/a 1, C/c 2, G/g 3, T/t 4, everything else 5 (reference bseq.c:9-26) */
unsigned char seq_nt6_table[256] = {
	[0 ... 255] = 5,
	['A'] = 1, ['a'] = 1, ['C'] = 2, ['c'] = 2, ['G'] = 3, ['g'] = 3, ['T'] = 4, ['t'] = 4
};

typedef struct { size_t l, m; char *s; } str_t;

#define RD_BUF (1 << 20)

struct bseq_file_s {
	gzFile fp;
	unsigned char *buf;
	const unsigned char *pre; /* text to consume before the file (fqblock.c hands its unread block over), owned */
	size_t pre_len, pre_pos;
	int begin, end, eof;
	int pending;            /* header character already consumed ('>' or '@'), or 0 */
	int comment_seen;       /* comment.s is valid (sticky, see above) */
	str_t name, comment, seq, qual;
};

static inline void str_reserve(str_t *s, size_t need)
{
	if (s->m < need) {
		s->m = need + (need >> 1) + 64;
		s->s = (char*)realloc(s->s, s->m);
	}
}

static int rd_fill(bseq_file_t *f)
{
	if (f->pre_pos < f->pre_len) {
		const size_t n = f->pre_len - f->pre_pos < RD_BUF ? f->pre_len - f->pre_pos : RD_BUF;
		memcpy(f->buf, f->pre + f->pre_pos, n);
		f->pre_pos += n, f->begin = 0, f->end = (int)n;
		return 1;
	}
	if (f->eof) return 0;
	f->begin = 0;
	f->end = gzread(f->fp, f->buf, RD_BUF);
	if (f->end < RD_BUF) f->eof = 1;
	if (f->end <= 0) { f->end = 0; return 0; }
	return 1;
}

static inline int rd_getc(bseq_file_t *f)
{
	if (f->begin >= f->end && !rd_fill(f)) return -1;
	return f->buf[f->begin++];
}

/* append up to (not including) the next delimiter; `line`: delimiter is '\n', otherwise any
 * white space.  Returns -1 if the input was exhausted on entry, else the string length.
 * *dret = the delimiter that ended the token, 0 at end of input. */
static long rd_until(bseq_file_t *f, int line, str_t *str, int *dret, int append)
{
	int gotany = 0;
	if (dret) *dret = 0;
	if (!append) str->l = 0;
	if (f->begin >= f->end && f->eof && f->pre_pos >= f->pre_len) return -1;
	for (;;) {
		int i;
		if (f->begin >= f->end && !rd_fill(f)) break;
		gotany = 1;
		if (line) {
			unsigned char *nl = (unsigned char*)memchr(f->buf + f->begin, '\n', f->end - f->begin);
			i = nl ? (int)(nl - f->buf) : f->end;
		} else for (i = f->begin; i < f->end && !isspace(f->buf[i]); ++i) {}
		str_reserve(str, str->l + (i - f->begin) + 2);
		memcpy(str->s + str->l, f->buf + f->begin, i - f->begin);
		str->l += i - f->begin;
		f->begin = i + 1;
		if (i < f->end) {
			if (dret) *dret = f->buf[i];
			break;
		}
	}
	if (!gotany) return -1; /* the input ended before this token began (kseq.h: !gotany && ks_eof) */
	str_reserve(str, str->l + 2);
	if (line && str->l > 1 && str->s[str->l - 1] == '\r') --str->l;
	str->s[str->l] = 0;
	return (long)str->l;
}

/* >= 0 sequence length; -1 end of input; -2 malformed quality */
static long rd_record(bseq_file_t *f)
{
	int c;
	if (f->pending == 0) {
		while ((c = rd_getc(f)) != -1 && c != '>' && c != '@') {}
		if (c == -1) return -1;
		f->pending = c;
	}
	f->seq.l = f->qual.l = 0;
	if (rd_until(f, 0, &f->name, &c, 0) < 0) return -1;
	if (c != '\n') { rd_until(f, 1, &f->comment, 0, 0); f->comment_seen = 1; }
	str_reserve(&f->seq, 256);
	while ((c = rd_getc(f)) != -1 && c != '>' && c != '+' && c != '@') {
		if (c == '\n') continue;
		str_reserve(&f->seq, f->seq.l + 2);
		f->seq.s[f->seq.l++] = (char)c;
		rd_until(f, 1, &f->seq, 0, 1);
	}
	if (c == '>' || c == '@') f->pending = c;
	f->seq.s[f->seq.l] = 0;
	if (c != '+') return (long)f->seq.l;
	while ((c = rd_getc(f)) != -1 && c != '\n') {}
	if (c == -1) return -2;
	while (rd_until(f, 1, &f->qual, 0, 1) >= 0 && f->qual.l < f->seq.l) {}
	f->pending = 0;
	if (f->seq.l != f->qual.l) return -2;
	return (long)f->seq.l;
}

bseq_file_t *bseq_open(const char *fn)
{
	bseq_file_t *f;
	gzFile g = fn && strcmp(fn, "-") ? gzopen(fn, "r") : gzdopen(fileno(stdin), "r");
	if (g == 0) return 0;
	gzbuffer(g, 1 << 18);
	f = (bseq_file_t*)calloc(1, sizeof(bseq_file_t));
	f->fp = g;
	f->buf = (unsigned char*)malloc(RD_BUF);
	return f;
}

/* fqblock.c: go on with an open stream in the tolerant parser; `pre` (malloc'd, taken over) is read before the file,
 * `comment` is kseq's sticky comment so far (NULL = none seen) */
bseq_file_t *bseq_open_from(void *gz, unsigned char *pre, size_t pre_len, const char *comment)
{
	bseq_file_t *f = (bseq_file_t*)calloc(1, sizeof(bseq_file_t));
	f->fp = (gzFile)gz;
	f->buf = (unsigned char*)malloc(RD_BUF);
	f->pre = pre, f->pre_len = pre_len;
	if (comment) {
		const size_t l = strlen(comment);
		str_reserve(&f->comment, l + 2);
		memcpy(f->comment.s, comment, l + 1);
		f->comment.l = l, f->comment_seen = 1;
	}
	return f;
}

/* nothing left to read (an empty batch from bseq_read() otherwise means "stopped at a malformed record") */
int bseq_at_eof(const bseq_file_t *f) { return f->begin >= f->end && f->eof && f->pre_pos >= f->pre_len; }

void bseq_close(bseq_file_t *f)
{
	if (f == 0) return;
	gzclose(f->fp);
	free((void*)f->pre);
	free(f->name.s); free(f->comment.s); free(f->seq.s); free(f->qual.s); free(f->buf);
	free(f);
}

static char *dup_n(const char *s, size_t l)
{
	char *r = (char*)malloc(l + 1);
	memcpy(r, s, l);
	r[l] = 0;
	return r;
}

bseq1_t *bseq_read(bseq_file_t *f, int chunk_size, int keep_comment, int *n_)
{
	int size = 0, m = 0, n = 0;
	bseq1_t *seqs = 0;
	while (rd_record(f) >= 0) {
		bseq1_t *s;
		if (n >= m) {
			m = m ? m << 1 : 256;
			seqs = (bseq1_t*)realloc(seqs, (size_t)m * sizeof(bseq1_t));
		}
		s = &seqs[n++];
		s->name = dup_n(f->name.s, f->name.l);
		s->comment = f->comment.s && f->comment_seen && keep_comment ? dup_n(f->comment.s, strlen(f->comment.s)) : 0; /* (bseq.c:66 tests the pointer) */
		s->seq = dup_n(f->seq.s, f->seq.l);
		s->qual = f->qual.l ? dup_n(f->qual.s, f->qual.l) : 0;
		s->l_seq = (int)f->seq.l;
		s->aux = s->aux2 = 0;
		size += s->l_seq;
		if (size >= chunk_size) break;
	}
	*n_ = n;
	return seqs;
}
