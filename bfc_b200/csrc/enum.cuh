// enum.cuh -- K0 k_enum: canonical k-mer enumeration over a read stream.
//
// What worker_count (reference count.c:72-89) and bfc_ec_kcov (correct.c:96-117) do per
// base -- map the character (bseq.c:9-26), roll the 4-plane k-mer (kmer.h:10-17), reset
// on a non-ACGT character, track "all k bases have Q >= q" -- and, once k bases are in,
// bfc_kmer_hash (kmer.h:79-88).  The stream is "reads back to back" (bfc_b200.h): the
// 0 terminator between reads is a non-ACGT character, so read boundaries need no
// special handling and the kernel never looks at the offsets.
//
// Work split: a CTA stages 256 x 36 stream positions (+ k-1 of warm-up) as 1-byte
// codes in shared memory with coalesced loads; each thread then rolls over its own 36
// positions (36 bytes apart = 9 words: conflict-free).  Output is one 16-byte record
// per position, y1 = ~0 where no k-mer ends:
//     rec_y0[i] = y[0] | is_high << 63,   rec_y1[i] = y[1]      (y as in bfc_kmer_hash)
// written in a blocked order (record i of a segment = iteration j * 256 + thread t) so
// that the stores of a warp are contiguous; enum_pos_of_record() maps back.
#pragma once
#include "common.cuh"

#define ENUM_THREADS 256
#define ENUM_CHUNK   36
#define ENUM_SEG     (ENUM_THREADS * ENUM_CHUNK)
#define ENUM_HALO_MAX 64

struct EnumParams {
	const uint8_t *seq, *qual;   // stream window (device); qual may be 0 (= every base has high quality)
	uint64_t len;                // bytes in the window
	uint64_t emit_from;          // window positions < emit_from only warm the rolling k-mer up
	int k, q;
	unsigned long long *rec_y0, *rec_y1;
};

static inline uint64_t enum_padded(uint64_t n_positions) { return (n_positions + ENUM_SEG - 1) / ENUM_SEG * ENUM_SEG; }

// stream position (relative to emit_from) of record r
__host__ __device__ __forceinline__ uint32_t enum_pos_of_record(uint32_t r)
{
	const uint32_t seg = r / ENUM_SEG, rem = r % ENUM_SEG;
	return seg * ENUM_SEG + (rem % ENUM_THREADS) * ENUM_CHUNK + rem / ENUM_THREADS;
}

// record index of stream position pos (relative to emit_from)
__host__ __device__ __forceinline__ uint64_t enum_record_of_pos(uint64_t pos)
{
	const uint64_t seg = pos / ENUM_SEG, rem = pos % ENUM_SEG;
	return seg * ENUM_SEG + (rem % ENUM_CHUNK) * ENUM_THREADS + rem / ENUM_CHUNK;
}

// ---- stream-order enumeration from bit planes (k_enum_lin in count_part.cu, k_ec_lookup in correct.cu)
#define EL_THREADS 256
#define EL_ITERS   32
#define EL_SEG     (EL_THREADS * EL_ITERS)       // stream positions per CTA
#define EL_LEAD    64                            // plane bits in front of the segment (>= k - 1)
#define EL_WORDS   ((EL_SEG + EL_LEAD) / 32 + 2)

static inline uint64_t el_padded(uint64_t n_positions) { return (n_positions + EL_SEG - 1) / EL_SEG * EL_SEG; }

#ifdef __CUDACC__
// 64 plane bits starting at bit index `bit`
__device__ __forceinline__ uint64_t win64(const uint32_t *pl, uint32_t bit)
{
	const uint32_t w = bit >> 5, r = bit & 31;
	const uint32_t a = pl[w], b = pl[w + 1], c = pl[w + 2];
	return (uint64_t)__funnelshift_r(b, c, r) << 32 | __funnelshift_r(a, b, r);
}

// Stage EL_LEAD + EL_SEG stream positions starting at seg0 - EL_LEAD as four bit planes in shared memory: B0, B1
// (base code bits, bseq.c:9-26), NB (not ACGT, or outside [0, len)), Q (an ACGT base with Q >= q; every ACGT base
// when qual == 0).  Bit i of a plane = position seg0 - EL_LEAD + i.  Ends with a __syncthreads().
// b_from_q (refine mode, correct.c:31): a base whose quality is <= 5 was corrected by an earlier round, which left the
// ORIGINAL base in the quality string as 34 + base -- that is the base to start from again.
__device__ __forceinline__ void el_stage_planes(uint32_t (*s_pl)[EL_WORDS], const uint8_t *seq, const uint8_t *qual, uint64_t len, int64_t seg0, int q_min,
                                                bool b_from_q = false)
{
	const unsigned lane = threadIdx.x & 31;
	if (threadIdx.x < 8) s_pl[threadIdx.x & 3][EL_WORDS - 1 - (threadIdx.x >> 2)] = 0;
	for (int i = threadIdx.x; i < EL_SEG + EL_LEAD; i += EL_THREADS) { // whole warps in or out
		const int64_t pos = seg0 - EL_LEAD + i;
		uint32_t c = 4, q = 0;
		if (pos >= 0 && (uint64_t)pos < len) {
			const uint8_t sc = __ldg(seq + pos);
			const int qv = qual ? (int)__ldg(qual + pos) : 0xFF; // 0xFF = this read has no quality string (bfc_b200.h)
			c = base_code(sc);
			if (b_from_q && qual && qv != 0xFF && sc != 0 && qv - 33 <= 5) { c = (uint32_t)(qv - 34) & 7; if (c > 3) c = 4; } // (3-bit field)
			q = c < 4 && (qual == 0 || qv - 33 >= q_min);
		}
		const uint32_t b0 = __ballot_sync(0xffffffffu, c & 1), b1 = __ballot_sync(0xffffffffu, c & 2);
		const uint32_t nb = __ballot_sync(0xffffffffu, c > 3), bq = __ballot_sync(0xffffffffu, q);
		if (lane == 0) s_pl[0][i >> 5] = b0, s_pl[1][i >> 5] = b1, s_pl[2][i >> 5] = nb, s_pl[3][i >> 5] = bq;
	}
	__syncthreads();
}

// The same for the NB plane alone (what deciding WHERE k-mers end needs)
__device__ __forceinline__ void el_stage_nb(uint32_t *s_nb, const uint8_t *seq, uint64_t len, int64_t seg0)
{
	const unsigned lane = threadIdx.x & 31;
	if (threadIdx.x < 2) s_nb[EL_WORDS - 1 - threadIdx.x] = 0;
	for (int i = threadIdx.x; i < EL_SEG + EL_LEAD; i += EL_THREADS) {
		const int64_t pos = seg0 - EL_LEAD + i;
		const bool not_acgt = !(pos >= 0 && (uint64_t)pos < len && base_code(__ldg(seq + pos)) < 4);
		const uint32_t nb = __ballot_sync(0xffffffffu, not_acgt);
		if (lane == 0) s_nb[i >> 5] = nb;
	}
	__syncthreads();
}

// Word w (0 .. EL_SEG/32 - 1) of the segment's "a k-mer ends here" plane: position p qualifies when none of the k
// bases p-k+1 .. p is non-ACGT, i.e. when the NB plane dilated by k-1 positions is 0 at p.  The dilation doubles:
// OR of shifts 0 .. k-1 of the 96 NB bits ending with this word.
__device__ __forceinline__ uint32_t el_valid_word(const uint32_t *s_nb, int w, int k)
{
	uint64_t lo = s_nb[w] | (uint64_t)s_nb[w + 1] << 32, hi = s_nb[w + 2]; // plane words w .. w+2 = positions 32w-64 .. 32w+31
	for (int covered = 1; covered < k;) {
		const int step = covered < k - covered ? covered : k - covered;
		hi |= (hi << step) | (lo >> (64 - step));
		lo |= lo << step;
		covered += step;
	}
	return ~(uint32_t)hi;
}

// exclusive prefix sum over the EL_THREADS values of a CTA (one per thread); *total = their sum.  Two barriers.
__device__ __forceinline__ uint32_t el_block_scan(uint32_t v, uint32_t *s_warp /* EL_THREADS / 32 + 1 */, uint32_t *total)
{
	const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t inc = v;
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane >= (unsigned)d) inc += u;
	}
	if (lane == 31) s_warp[warp] = inc;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t run = 0;
		for (int i = 0; i < EL_THREADS / 32; ++i) { const uint32_t t = s_warp[i]; s_warp[i] = run; run += t; }
		s_warp[EL_THREADS / 32] = run;
	}
	__syncthreads();
	*total = s_warp[EL_THREADS / 32];
	return s_warp[warp] + inc - v;
}

// The canonical k-mer hash (kmer.h:79-88) of the k bases whose oldest one is plane bit `bit`; false when one of them
// is not ACGT.  The 4-plane k-mer (kmer.h:10-17) is cut out of the base planes: forward planes = the window
// bit-reversed (newest base at bit 0), reverse-complement planes = the window complemented (newest base at bit k-1).
__device__ __forceinline__ void el_kmer_hash_at(const uint32_t (*s_pl)[EL_WORDS], uint32_t bit, int k, uint64_t kmask, uint64_t y[2])
{
	const uint64_t w0 = win64(s_pl[0], bit) & kmask, w1 = win64(s_pl[1], bit) & kmask;
	const uint64_t x[4] = { __brevll(w0) >> (64 - k), __brevll(w1) >> (64 - k), ~w0 & kmask, ~w1 & kmask };
	bfc_kmer_hash(k, x, y);
}

__device__ __forceinline__ bool el_kmer_at(const uint32_t (*s_pl)[EL_WORDS], uint32_t bit, int k, uint64_t kmask, uint64_t y[2])
{
	if ((win64(s_pl[2], bit) & kmask) != 0) return false;
	el_kmer_hash_at(s_pl, bit, k, kmask, y);
	return true;
}

static __global__ void __launch_bounds__(ENUM_THREADS) k_enum(EnumParams p)
{
	__shared__ uint8_t s_code[ENUM_SEG + ENUM_HALO_MAX];
	const int halo = p.k - 1;
	const int64_t seg0 = (int64_t)p.emit_from + (int64_t)blockIdx.x * ENUM_SEG;

	// stage the CTA's window as codes: bits 0-2 base (4 = not ACGT / outside), bit 3 Q >= q
	for (int i = threadIdx.x; i < ENUM_SEG + halo; i += ENUM_THREADS) {
		const int64_t pos = seg0 - halo + i;
		uint32_t c = 4;
		if (pos >= 0 && (uint64_t)pos < p.len) {
			c = base_code(__ldg(p.seq + pos));
			if (c < 4 && (p.qual == 0 || (int)__ldg(p.qual + pos) - 33 >= p.q)) c |= 8;
		}
		s_code[i] = (uint8_t)c;
	}
	__syncthreads();

	const int base = threadIdx.x * ENUM_CHUNK;
	const int k = p.k;
	const uint64_t mask = (1ULL << k) - 1;
	uint64_t x[4] = {0, 0, 0, 0}, qmer = 0;
	int l = 0;
	unsigned long long *o0 = p.rec_y0 + (uint64_t)blockIdx.x * ENUM_SEG + threadIdx.x;
	unsigned long long *o1 = p.rec_y1 + (uint64_t)blockIdx.x * ENUM_SEG + threadIdx.x;

	for (int j = 0; j < halo; ++j) { // warm-up: roll only
		const uint32_t c = s_code[base + j];
		if ((c & 7) < 4) {
			bfc_kmer_append(k, x, c & 3);
			qmer = (qmer << 1 | (c >> 3)) & mask;
			++l;
		} else l = 0, qmer = 0, x[0] = x[1] = x[2] = x[3] = 0;
	}
	for (int j = 0; j < ENUM_CHUNK; ++j) {
		const uint32_t c = s_code[base + halo + j];
		uint64_t y[2] = {0, ~0ULL};
		if ((c & 7) < 4) {
			bfc_kmer_append(k, x, c & 3);
			qmer = (qmer << 1 | (c >> 3)) & mask;
			if (++l >= k) {
				bfc_kmer_hash(k, x, y);
				y[0] |= (unsigned long long)(qmer == mask) << 63;
			}
		} else l = 0, qmer = 0, x[0] = x[1] = x[2] = x[3] = 0;
		o0[j * ENUM_THREADS] = y[0];
		o1[j * ENUM_THREADS] = y[1];
	}
}
#endif
