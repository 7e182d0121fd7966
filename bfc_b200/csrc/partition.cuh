// partition.cuh -- stable radix partition of the count phase's k-mer records, written for their shape: a 64-bit key
// (whose low bits are y0, the Bloom block index among them) and a 1 / 2 / 4 / 8-byte value.  It replaces the
// library radix sort round 1 used (3 passes of 8 bits over 20 partition bits): here the digits are up to 10 bits wide,
// so the 20 bits of the usual geometry (k = 33, 2^37-bit filter, 16 KB slices) take TWO passes.
//
// A pass is one sweep (single read, single write) in the manner of a decoupled look-back scan:
//   * k_rp_hist has counted every digit of every pass beforehand (one read of the keys), k_rp_scan turned the counts
//     into the first output position of every digit;
//   * tiles of RP_TILE records are taken in order (a ticket), each by one CTA: keys and values are read coalesced,
//     warp w owning records [w * 256, (w + 1) * 256) of the tile, 32 at a time, in order (one CTA of 1024 threads per SM);
//   * the rank of a record among the records of its digit in the tile comes from ballots over the digit's bits (peers of the same
//     digit inside a warp step, lower lanes first) + a per-warp running count per digit in shared memory + a scan over
//     the warps: stream order is preserved inside every digit = the partition is STABLE, which is what keeps the
//     records of one Bloom block in read order (the reference's -t1 semantics, count.c:54-70);
//   * the tile publishes its per-digit counts, looks back over the earlier tiles' counts / running totals to get its
//     own offset inside every digit, and publishes its running totals;
//   * the records are put in digit order in shared memory first, so that what leaves for a digit's output range is a
//     run of consecutive addresses.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#define RP_THREADS 1024                         // one thread per digit in the look-back; one CTA per SM
#define RP_ITEMS   8
#define RP_TILE    (RP_THREADS * RP_ITEMS)      // records per tile
#define RP_WARPS   (RP_THREADS / 32)
#define RP_MAX_D   10                           // widest digit
#define RP_MAX_NB  (1 << RP_MAX_D)
#define RP_MAX_PASSES 4
#define RP_AGG     0x40000000u                  // status word: low 30 bits = a count, bit 30 = "tile's own count",
#define RP_PREFIX  0x80000000u                  //              bit 31 = "running total up to and including this tile"
#define RP_VALUE   0x3fffffffu

struct RpPlan { // digits of a partition by bits [begin, end) of the key, least significant first
	int n_passes, shift[RP_MAX_PASSES], bits[RP_MAX_PASSES];
};

static inline RpPlan rp_plan(int begin, int end)
{
	RpPlan p;
	const int nb = end - begin;
	p.n_passes = nb <= 0 ? 0 : (nb + RP_MAX_D - 1) / RP_MAX_D;
	int at = begin;
	for (int i = 0; i < p.n_passes; ++i) { // as even as possible
		const int left = end - at, passes_left = p.n_passes - i;
		p.shift[i] = at, p.bits[i] = (left + passes_left - 1) / passes_left;
		at += p.bits[i];
	}
	return p;
}

static inline uint64_t rp_tiles(uint64_t n) { return (n + RP_TILE - 1) / RP_TILE; }

// scratch: histograms / digit bases of every pass, the ticket, and one status word per (tile, digit) of a pass
static inline size_t rp_scratch_bytes(uint64_t n, int begin, int end)
{
	const RpPlan p = rp_plan(begin, end);
	if (p.n_passes == 0) return 256;
	return (size_t)RP_MAX_PASSES * RP_MAX_NB * 4 + 256 + (size_t)rp_tiles(n) * RP_MAX_NB * 4 + 256;
}

struct RpHistParams { const unsigned long long *key; uint64_t n; int n_passes, shift[RP_MAX_PASSES], bits[RP_MAX_PASSES]; uint32_t *hist; };

// digit counts of all passes in one read of the keys (shared-memory histograms, one global add per bin and CTA)
__global__ void __launch_bounds__(512) k_rp_hist(RpHistParams p)
{
	__shared__ uint32_t s_h[RP_MAX_PASSES][RP_MAX_NB];
	for (int i = threadIdx.x; i < RP_MAX_PASSES * RP_MAX_NB; i += blockDim.x) (&s_h[0][0])[i] = 0;
	__syncthreads();
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (uint64_t)gridDim.x * blockDim.x) {
		const unsigned long long k = __ldg(p.key + i);
#pragma unroll
		for (int j = 0; j < RP_MAX_PASSES; ++j)
			if (j < p.n_passes) atomicAdd(&s_h[j][(uint32_t)(k >> p.shift[j]) & ((1u << p.bits[j]) - 1)], 1u);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < RP_MAX_PASSES * RP_MAX_NB; i += blockDim.x) {
		const uint32_t v = (&s_h[0][0])[i];
		if (v) atomicAdd(p.hist + i, v);
	}
}

// exclusive scan of every pass's histogram, in place: first output position of every digit (one CTA of RP_MAX_NB threads)
__global__ void __launch_bounds__(RP_MAX_NB) k_rp_scan(uint32_t *hist, int n_passes)
{
	__shared__ uint32_t s_w[RP_MAX_NB / 32];
	const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int j = 0; j < n_passes; ++j) {
		const uint32_t v = hist[j * RP_MAX_NB + threadIdx.x];
		uint32_t inc = v;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (unsigned)d) inc += u; }
		if (lane == 31) s_w[warp] = inc;
		__syncthreads();
		if (warp == 0) {
			uint32_t t = s_w[lane], ti = t;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, ti, d); if (lane >= (unsigned)d) ti += u; }
			s_w[lane] = ti - t;
		}
		__syncthreads();
		hist[j * RP_MAX_NB + threadIdx.x] = s_w[warp] + inc - v;
		__syncthreads();
	}
}

template <typename VT>
struct RpPassParams {
	const unsigned long long *key_in;
	const VT *val_in;
	unsigned long long *key_out;
	VT *val_out;
	uint64_t n;
	int shift, bits;
	const uint32_t *base;        // first output position of every digit (k_rp_scan)
	uint32_t *status;            // [tile][1 << bits], zero on entry
	unsigned int *ticket;        // zero on entry
};

template <typename VT>
__global__ void __launch_bounds__(RP_THREADS, 1) k_rp_pass(RpPassParams<VT> p)
{
	extern __shared__ __align__(128) unsigned char s_raw[];
	// layout: per-warp digit counts (u16) | digit starts inside the sorted tile, then output offsets (u32) | keys | values
	uint16_t *const s_wh = (uint16_t*)s_raw; // [warp][digit] (lanes of a warp spread over the banks by their digits)
	uint32_t *const s_start = (uint32_t*)(s_raw + sizeof(uint16_t) * RP_WARPS * RP_MAX_NB);
	unsigned long long *const s_key = (unsigned long long*)(s_start + RP_MAX_NB);
	VT *const s_val = (VT*)(s_key + RP_TILE);
	__shared__ uint32_t s_scan[RP_WARPS + 1];
	__shared__ unsigned int s_tile;
	const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t nb = 1u << p.bits, dmask = nb - 1;

	if (tid == 0) s_tile = atomicAdd(p.ticket, 1u); // tiles start in order: whatever a tile waits for is already running
	for (uint32_t i = tid; i < RP_WARPS * RP_MAX_NB / 8; i += RP_THREADS) ((uint4*)s_raw)[i] = make_uint4(0, 0, 0, 0);
	__syncthreads();
	const uint64_t tile = s_tile, t0 = tile * RP_TILE;
	const uint32_t n_tile = (uint32_t)(p.n - t0 < RP_TILE ? p.n - t0 : RP_TILE);

	// ---- read; rank of every record among the records of its digit that this WARP holds, in stream order
	unsigned long long key[RP_ITEMS];
	VT val[RP_ITEMS];
	uint32_t dr[RP_ITEMS]; // digit | rank inside the warp << 16
#pragma unroll
	for (int j = 0; j < RP_ITEMS; ++j) {
		const uint32_t it = warp * (32 * RP_ITEMS) + j * 32 + lane;
		const bool valid = it < n_tile;
		key[j] = valid ? __ldg(p.key_in + t0 + it) : 0ULL;
		val[j] = valid ? __ldg(p.val_in + t0 + it) : (VT)0;
	}
#pragma unroll
	for (int j = 0; j < RP_ITEMS; ++j) {
		const uint32_t it = warp * (32 * RP_ITEMS) + j * 32 + lane;
		const bool valid = it < n_tile;
		const uint32_t d = valid ? (uint32_t)(key[j] >> p.shift) & dmask : 0, idx = warp * RP_MAX_NB + d;
		// The lanes of this step that hold the same digit.  32 records over 2^bits digits: usually none do, and that is
		// found out through the count itself -- every lane reads it, tags it with its lane number, and sees after a
		// warp barrier whether its tag survived.  Only when some lane lost does the warp work the groups out, with one
		// ballot per digit bit (match.any is far slower than that here).  (The tag stores of lanes with the same digit
		// go to the same 16-bit word in the same instruction -- compute-sanitizer's racecheck reports them as write-write
		// hazards, which they are by design: exactly one of the stores lands, whole, and the losers find out.)
		const uint32_t before = valid ? s_wh[idx] : 0; // < 512: a warp holds 256 records of a tile
		__syncwarp();
		if (valid) s_wh[idx] = (uint16_t)(lane << 9 | before);
		__syncwarp();
		const bool lost = valid && (uint32_t)(s_wh[idx] >> 9) != lane;
		unsigned peers = 1u << lane;
		if (__any_sync(0xffffffffu, lost)) {
			peers = __ballot_sync(0xffffffffu, valid);
			for (int b = 0; b < p.bits; ++b) {
				const unsigned m = __ballot_sync(0xffffffffu, d >> b & 1);
				peers &= (d >> b & 1) ? m : ~m;
			}
			if (!valid) peers = 1u << lane; // (a record past the end matches nobody)
		}
		__syncwarp();
		if (valid && (int)lane == __ffs(peers) - 1) s_wh[idx] = (uint16_t)(before + __popc(peers));
		dr[j] = valid ? d | (before + __popc(peers & ((1u << lane) - 1))) << 16 : 0xffffffffu;
		__syncwarp();
	}
	__syncthreads();

	// ---- digit `tid`: counts of the warps -> offsets of the warps; count of the tile
	uint32_t cnt = 0;
	if (tid < nb) {
#pragma unroll 8
		for (int w = 0; w < RP_WARPS; ++w) { const uint32_t c = s_wh[w * RP_MAX_NB + tid]; s_wh[w * RP_MAX_NB + tid] = (uint16_t)cnt; cnt += c; }
	}
	// ---- publish the tile's count, look back for its offset inside the digit, publish the running total.  The earlier
	// tiles are read four at a time (the reads of a walk are independent; only the decisions are in order).
	uint32_t excl = 0;
	if (tid < nb) {
		volatile uint32_t *st = p.status + tid;
		st[tile * nb] = cnt | (tile == 0 ? RP_PREFIX : RP_AGG);
		int64_t t = (int64_t)tile - 1;
		bool done = t < 0;
		while (!done) {
			uint32_t v[4];
#pragma unroll
			for (int q = 0; q < 4; ++q) v[q] = t - q >= 0 ? st[(uint64_t)(t - q) * nb] : RP_PREFIX;
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				if (done) break;
				if ((v[q] & (RP_AGG | RP_PREFIX)) == 0) break; // not published yet: read again from here
				excl += v[q] & RP_VALUE;
				--t;
				if ((v[q] & RP_PREFIX) || t < 0) done = true;
			}
		}
		if (tile) st[tile * nb] = ((excl + cnt) & RP_VALUE) | RP_PREFIX;
	}
	// ---- where every digit starts inside the tile once it is in digit order (exclusive scan of the counts)
	{
		uint32_t inc = cnt;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (unsigned)d) inc += u; }
		if (lane == 31) s_scan[warp] = inc;
		__syncthreads();
		if (warp == 0) {
			const uint32_t tsum = s_scan[lane];
			uint32_t ti = tsum;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, ti, d); if (lane >= (unsigned)d) ti += u; }
			s_scan[lane] = ti - tsum;
		}
		__syncthreads();
		if (tid < nb) s_start[tid] = s_scan[warp] + inc - cnt;
	}
	__syncthreads();
	// ---- into digit order in shared memory (stable: digit start + offset of the warp + rank inside the warp)
#pragma unroll
	for (int j = 0; j < RP_ITEMS; ++j)
		if (dr[j] != 0xffffffffu) {
			const uint32_t d = dr[j] & 0xffff, at = s_start[d] + s_wh[warp * RP_MAX_NB + d] + (dr[j] >> 16);
			s_key[at] = key[j], s_val[at] = val[j];
		}
	__syncthreads();
	// s_start[d] becomes: output position of the digit's first record of this tile - its position in the tile
	if (tid < nb) s_start[tid] = __ldg(p.base + tid) + excl - s_start[tid];
	__syncthreads();
	for (uint32_t i = tid; i < n_tile; i += RP_THREADS) {
		const unsigned long long k = s_key[i];
		const uint32_t at = s_start[(uint32_t)(k >> p.shift) & dmask] + i;
		p.key_out[at] = k, p.val_out[at] = s_val[i];
	}
}

template <typename VT> static inline size_t rp_pass_smem() { return sizeof(uint16_t) * RP_WARPS * RP_MAX_NB + 4 * RP_MAX_NB + (size_t)RP_TILE * (8 + sizeof(VT)); }

// Stable partition of n records by bits [begin, end) of the key: (key_in, val_in) -> (key_out, val_out); tmp_key / tmp_val
// (n records) hold the intermediate when there is more than one pass; scratch: rp_scratch_bytes().  n < 2^30.
template <typename VT>
static cudaError_t rp_partition(cudaStream_t stream, int sm_count, const unsigned long long *key_in, const VT *val_in, unsigned long long *key_out, VT *val_out,
                                unsigned long long *tmp_key, VT *tmp_val, uint8_t *scratch, uint64_t n, int begin, int end, uint64_t *n_launches)
{
	const RpPlan pl = rp_plan(begin, end);
	cudaError_t e;
	if (pl.n_passes == 0 || n == 0) return cudaSuccess;
	if (pl.n_passes > RP_MAX_PASSES || n >= (1ULL << 30)) return cudaErrorInvalidValue;
	static bool attr_done = false; // (per record type; the attribute is per device, set again is harmless)
	(void)attr_done;
	if ((e = cudaFuncSetAttribute(k_rp_pass<VT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rp_pass_smem<VT>())) != cudaSuccess) return e;
	uint32_t *hist = (uint32_t*)scratch;
	unsigned int *ticket = (unsigned int*)(scratch + (size_t)RP_MAX_PASSES * RP_MAX_NB * 4);
	uint32_t *status = (uint32_t*)(scratch + (size_t)RP_MAX_PASSES * RP_MAX_NB * 4 + 256);
	const uint64_t n_tiles = rp_tiles(n);
	if ((e = cudaMemsetAsync(hist, 0, (size_t)RP_MAX_PASSES * RP_MAX_NB * 4, stream)) != cudaSuccess) return e;
	RpHistParams hp;
	hp.key = key_in, hp.n = n, hp.n_passes = pl.n_passes, hp.hist = hist;
	for (int i = 0; i < RP_MAX_PASSES; ++i) hp.shift[i] = i < pl.n_passes ? pl.shift[i] : 0, hp.bits[i] = i < pl.n_passes ? pl.bits[i] : 0;
	k_rp_hist<<<sm_count * 4, 512, 0, stream>>>(hp);
	k_rp_scan<<<1, RP_MAX_NB, 0, stream>>>(hist, pl.n_passes);
	if (n_launches) *n_launches += 2;
	const unsigned long long *ki = key_in;
	const VT *vi = val_in;
	for (int i = 0; i < pl.n_passes; ++i) {
		const bool to_out = ((pl.n_passes - 1 - i) & 1) == 0; // the last pass lands in the output
		RpPassParams<VT> pp;
		pp.key_in = ki, pp.val_in = vi, pp.key_out = to_out ? key_out : tmp_key, pp.val_out = to_out ? val_out : tmp_val;
		pp.n = n, pp.shift = pl.shift[i], pp.bits = pl.bits[i], pp.base = hist + i * RP_MAX_NB, pp.status = status, pp.ticket = ticket;
		if ((e = cudaMemsetAsync(ticket, 0, 4, stream)) != cudaSuccess) return e;
		if ((e = cudaMemsetAsync(status, 0, (size_t)n_tiles * ((size_t)1 << pl.bits[i]) * 4, stream)) != cudaSuccess) return e;
		k_rp_pass<VT><<<(unsigned)n_tiles, RP_THREADS, rp_pass_smem<VT>(), stream>>>(pp);
		if (n_launches) ++*n_launches;
		ki = pp.key_out, vi = pp.val_out;
	}
	return cudaGetLastError();
}
