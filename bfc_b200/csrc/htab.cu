// htab.cu -- the reference's bfc_ch_* entry points (htab.h:13-23) over one
// open-addressing array in HBM (layout: include/htab.h, slot functions: common.cuh).
#include "common.cuh"
#include <algorithm>
#include <vector>

static const int TAB_MIN_RBITS = 6;         // >= 64 slots per region
static const uint64_t TAB_DEF_CAP = 1 << 20; // parked inserts per kernel before growth is forced

// ------------------------------------------------------------------ kernels

// reg_hi = the region index bits a shard leaves out (own_val << (l_pre - own_bits)), same for the old and the new table
__global__ void k_tab_rehash(const unsigned long long *old_slots, int old_rbits, int old_rot, uint32_t reg_hi, uint64_t old_n, TabView nt, unsigned long long *fail)
{
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < old_n; i += (uint64_t)gridDim.x * blockDim.x) {
		const unsigned long long s = old_slots[i];
		if (s && !tab_put_raw(nt, tab_region_inv(nt.l_pre, old_rot, reg_hi | (uint32_t)(i >> old_rbits)), s)) atomicAdd(fail, 1ULL);
	}
}

__global__ void k_tab_apply(TabView t, const unsigned long long *rec, uint64_t n)
{
	unsigned long long added = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		const unsigned long long y0 = rec[2 * i], y1 = rec[2 * i + 1];
		added += tab_upsert(t, y0 & ~(1ULL << 63), y1, (int)(y0 >> 63)) == 1;
	}
	block_add(t.counters, added);
}

__global__ void k_tab_put_raw(TabView t, const uint32_t *sub, const unsigned long long *key, uint64_t n, unsigned long long *fail)
{
	unsigned long long added = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		if (tab_put_raw(t, sub[i], key[i])) ++added;
		else atomicAdd(fail, 1ULL);
	}
	block_add(t.counters, added);
}

__global__ void k_tab_insert1(TabView t, uint64_t y0, uint64_t y1, int is_high)
{
	if (tab_upsert(t, y0, y1, is_high) == 1) atomicAdd(t.counters, 1ULL);
}

__global__ void k_tab_get(TabView t, const uint64_t *y, uint64_t n, int32_t *out)
{
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		out[i] = tab_get(t, y[2 * i], y[2 * i + 1]);
}

__global__ void k_tab_kmer_occ1(TabView t, bfc_kmer_t z, int32_t *out)
{
	*out = tab_kmer_occ(t, z.x);
}

// reference htab.c:110-122: histogram of the 8-bit counts and the 6-bit high counts
__global__ void k_tab_hist(const unsigned long long *slots, uint64_t n, unsigned long long *hist /* 256 + 64 */)
{
	__shared__ unsigned int s_h[320];
	for (int i = threadIdx.x; i < 320; i += blockDim.x) s_h[i] = 0;
	__syncthreads();
	// each CTA covers < 2^32 slots, so 32-bit shared counters cannot overflow
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		const unsigned long long s = __ldg(slots + i);
		if (s) {
			atomicAdd(&s_h[s & 0xff], 1u);
			atomicAdd(&s_h[256 + ((s >> 8) & 0x3f)], 1u);
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < 320; i += blockDim.x)
		if (s_h[i]) atomicAdd(hist + i, (unsigned long long)s_h[i]);
}

__global__ void k_tab_export(const unsigned long long *slots, int rbits, int l_pre, int rot, uint32_t reg_hi, uint64_t n, uint32_t *sub, unsigned long long *key,
                             unsigned long long *cursor)
{
	for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x; i0 < n; i0 += (uint64_t)gridDim.x * blockDim.x) { // (n is a multiple of 64: whole warps)
		const uint64_t i = i0 + threadIdx.x;
		const unsigned long long s = i < n ? __ldg(slots + i) : 0ULL;
		const unsigned m = __ballot_sync(0xffffffffu, s != 0), lane = threadIdx.x & 31;
		if (m == 0) continue;
		unsigned long long at = 0;
		if (lane == (unsigned)(__ffs(m) - 1)) at = atomicAdd(cursor, (unsigned long long)__popc(m)); // one atomic per warp
		at = __shfl_sync(0xffffffffu, at, __ffs(m) - 1) + __popc(m & ((1u << lane) - 1));
		if (s) {
			sub[at] = tab_region_inv(l_pre, rot, reg_hi | (uint32_t)(i >> rbits));
			key[at] = s;
		}
	}
}

// ------------------------------------------------------------------ growth

static inline uint64_t tab_capacity(const bfc_ch_s *ch) { return bfcg_tab_capacity(ch); }
static inline uint32_t tab_reg_hi(const bfc_ch_s *ch) { return ch->own_bits ? ch->own_val << (ch->l_pre - ch->own_bits) : 0; }

static int tab_read_counters(const bfc_ch_s *ch, unsigned long long c[2])
{
	BfcgRuntime &rt = bfcg_rt();
	BFCG_CUDA(cudaMemcpyAsync(c, ch->counters, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, rt.stream));
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	return BFCG_OK;
}

static int tab_resize(bfc_ch_s *ch, int new_rbits, int new_rot = -1)
{
	if (new_rot < 0) new_rot = ch->rot;
	BfcgRuntime &rt = bfcg_rt();
	unsigned long long *ns = 0, *fail = ch->counters + 2, h_fail = 0; // own scratch word: callers may hold the arena
	const uint64_t new_cap = 1ULL << (ch->l_pre - ch->own_bits + new_rbits);
	if (bfc_verbose >= 4)
		fprintf(stderr, "[M::%s] growing the k-mer table: 2^%d -> 2^%d slots\n", __func__, ch->l_pre - ch->own_bits + ch->rbits, ch->l_pre - ch->own_bits + new_rbits);
	if (cudaMalloc(&ns, new_cap * 8) != cudaSuccess)
		return bfcg_fail(__func__, "cudaMalloc(larger k-mer table)", cudaErrorMemoryAllocation);
	BFCG_CUDA(cudaMemsetAsync(ns, 0, new_cap * 8, rt.stream));
	BFCG_CUDA(cudaMemsetAsync(fail, 0, 8, rt.stream));
	TabView nt = tab_view(ch);
	nt.slots = ns, nt.rbits = new_rbits, nt.rot = new_rot;
	{ KTime kt(KT_TAB_REHASH); k_tab_rehash<<<rt.sm_count * 8, 256, 0, rt.stream>>>(ch->slots, ch->rbits, ch->rot, tab_reg_hi(ch), tab_capacity(ch), nt, fail); }
	BFCG_LAUNCH_CHECK();
	BFCG_CUDA(cudaMemcpyAsync(&h_fail, fail, 8, cudaMemcpyDeviceToHost, rt.stream));
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	if (h_fail) return bfcg_fail(__func__, "rehash lost keys", cudaSuccess);
	cudaFree(ch->slots);
	ch->slots = ns, ch->rbits = new_rbits, ch->rot = new_rot;
	return BFCG_OK;
}

// The low x bits of y0 are the Bloom block index (when x <= k); sub-table index bit j is y0 bit j + k - l_pre
// (htab.c:45-58; for k <= 32 and l_pre > k it is y0 bit j - (l_pre - k)).  Rotating the sub-table index right by the
// number of its bits that lie below y0 bit x makes them the top bits of the region index.
// A shard request (bfcg_ch_set_shard) is settled here, while the table is still empty: the owner bits are the top
// owner_bits bits of the block index = y0 bits [x - owner_bits, x) = the top bits of the region index after the
// rotation, provided they are sub-table bits at all (rot >= owner_bits).
int bfcg_tab_align_to_filter(bfc_ch_s *ch, int x)
{
	int rot = x - (ch->k - ch->l_pre);
	if (x > ch->k || rot <= 0 || rot >= ch->l_pre) rot = 0;
	int own_bits = 0, skew = 1;
	if (ch->req_owners > 1) {
		int ob = 0;
		while ((1 << ob) < ch->req_owners) ++ob;
		if (rot >= ob) own_bits = ob;
		else skew = ch->req_owners;
	}
	if (rot == ch->rot && own_bits == ch->own_bits) { ch->skew = skew; return BFCG_OK; }
	if (bfc_ch_count(ch) == 0) {
		if (own_bits != ch->own_bits) { // another number of regions: a fresh, empty array
			BfcgRuntime &rt = bfcg_rt();
			unsigned long long *ns = 0;
			const uint64_t cap = 1ULL << (ch->l_pre - own_bits + ch->rbits);
			if (cudaMalloc(&ns, cap * 8) != cudaSuccess) return bfcg_fail(__func__, "cudaMalloc(k-mer table)", cudaErrorMemoryAllocation);
			BFCG_CUDA(cudaMemsetAsync(ns, 0, cap * 8, rt.stream));
			BFCG_CUDA(cudaStreamSynchronize(rt.stream));
			cudaFree(ch->slots);
			ch->slots = ns;
		}
		ch->rot = rot, ch->own_bits = own_bits, ch->own_val = own_bits ? (uint32_t)ch->req_owner : 0, ch->skew = skew;
		return BFCG_OK;
	}
	if (own_bits != ch->own_bits) return bfcg_fail(__func__, "the shard geometry of a non-empty k-mer table cannot change", cudaSuccess);
	ch->skew = skew;
	return tab_resize(ch, ch->rbits, rot);
}

int bfcg_tab_reserve(bfc_ch_s *ch, uint64_t extra)
{
	unsigned long long c[2];
	int r;
	if ((r = tab_read_counters(ch, c)) != BFCG_OK) return r;
	int rbits = ch->rbits;
	const uint64_t skew = ch->skew > 1 ? (uint64_t)ch->skew : 1; // a shard that keeps every region fills 1/skew of them
	while (2 * (c[0] + extra) * skew > (1ULL << (ch->l_pre - ch->own_bits + rbits))) ++rbits;
	return rbits != ch->rbits ? tab_resize(ch, rbits) : BFCG_OK;
}

int bfcg_tab_grow(bfc_ch_s *ch) { return tab_resize(ch, ch->rbits + 1); }

int bfcg_tab_set_rbits(bfc_ch_s *ch, int rbits) { return rbits > ch->rbits ? tab_resize(ch, rbits) : BFCG_OK; }

// Make `full` (an ordinary table, its contents are dropped) the table whose slot array is the concatenation, in owner
// order, of the slot arrays of the 2^own_bits shards shaped like `shard`: same region size, same region placement.
// The shards' owner bits are the top bits of the region index (bfcg_tab_align_to_filter), so region r of shard o IS
// region o << (l_pre - own_bits) | r of the whole table, probe sequences included.
int bfcg_tab_shape_like_shards(bfc_ch_s *full, const bfc_ch_s *shard)
{
	BfcgRuntime &rt = bfcg_rt();
	if (full->k != shard->k || full->l_pre != shard->l_pre) return bfcg_fail(__func__, "tables of different k", cudaSuccess);
	if (full->own_bits != 0 || full->rbits != shard->rbits || full->rot != shard->rot) {
		unsigned long long *ns = 0;
		const uint64_t cap = 1ULL << (full->l_pre + shard->rbits);
		cudaFree(full->slots);
		full->slots = 0;
		if (cudaMalloc(&ns, cap * 8) != cudaSuccess) return bfcg_fail(__func__, "cudaMalloc(k-mer table)", cudaErrorMemoryAllocation);
		full->slots = ns, full->rbits = shard->rbits, full->rot = shard->rot, full->own_bits = 0, full->own_val = 0, full->skew = 1;
		full->req_owners = 0;
	}
	BFCG_CUDA(cudaMemsetAsync(full->counters, 0, 16, rt.stream));
	full->prev_new = 0, full->have_prev = 0;
	return BFCG_OK;
}

int bfcg_tab_set_count(bfc_ch_s *ch, uint64_t n)
{
	BfcgRuntime &rt = bfcg_rt();
	unsigned long long v = n;
	BFCG_CUDA(cudaMemcpyAsync(ch->counters, &v, 8, cudaMemcpyHostToDevice, rt.stream));
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	return BFCG_OK;
}

int bfcg_tab_drain_deferred(bfc_ch_s *ch)
{
	BfcgRuntime &rt = bfcg_rt();
	for (int round = 0; round < 40; ++round) {
		unsigned long long c[2];
		int r;
		if ((r = tab_read_counters(ch, c)) != BFCG_OK) return r;
		if (c[1] == 0) return BFCG_OK;
		if (c[1] > ch->def_cap) {
			snprintf(rt.err, sizeof(rt.err), "k-mer table: %llu inserts hit full regions (limit %llu)", c[1], (unsigned long long)ch->def_cap);
			fprintf(stderr, "[E::%s] %s\n", __func__, rt.err);
			return BFCG_ERR_OVERFLOW;
		}
		// copy the parked inserts aside, grow, re-apply
		unsigned long long *tmp = 0;
		BFCG_CUDA(cudaMalloc(&tmp, c[1] * 16));
		BFCG_CUDA(cudaMemcpyAsync(tmp, ch->deferred, c[1] * 16, cudaMemcpyDeviceToDevice, rt.stream));
		BFCG_CUDA(cudaMemsetAsync(ch->counters + 1, 0, 8, rt.stream));
		if ((r = tab_resize(ch, ch->rbits + 1)) != BFCG_OK) { cudaFree(tmp); return r; }
		k_tab_apply<<<(unsigned)std::min<uint64_t>((c[1] + 255) / 256, 65535), 256, 0, rt.stream>>>(tab_view(ch), tmp, c[1]);
		BFCG_LAUNCH_CHECK();
		BFCG_CUDA(cudaStreamSynchronize(rt.stream));
		cudaFree(tmp);
	}
	return bfcg_fail(__func__, "deferred inserts did not drain", cudaSuccess);
}

// ------------------------------------------------------------------ C entry points

extern "C" {

// reference htab.c:19-34
bfc_ch_t *bfc_ch_init(int k, int l_pre)
{
	if (k > BFC_MAX_KMER || k < 1) {
		fprintf(stderr, "[E::%s] k=%d is outside 1..%d\n", __func__, k, BFC_MAX_KMER);
		return 0;
	}
	if (k * 2 - l_pre > BFC_CH_KEYBITS) l_pre = k * 2 - BFC_CH_KEYBITS;
	if (l_pre > BFC_CH_MAXPRE) l_pre = BFC_CH_MAXPRE;
	if (k - l_pre >= BFC_CH_KEYBITS || (k <= 32 && 2 * k <= l_pre) || (k > 32 && k <= l_pre) || l_pre < 0) {
		fprintf(stderr, "[E::%s] unsupported (k=%d, l_pre=%d)\n", __func__, k, l_pre);
		return 0;
	}
	if (bfcg_rt_init() != BFCG_OK) return 0;
	BfcgRuntime &rt = bfcg_rt();
	bfc_ch_s *ch = (bfc_ch_s*)calloc(1, sizeof(bfc_ch_s));
	ch->k = k, ch->l_pre = l_pre, ch->rbits = TAB_MIN_RBITS, ch->def_cap = TAB_DEF_CAP, ch->skew = 1;
	{ const char *dc = getenv("BFC_B200_TAB_DEFCAP"); if (dc && atoll(dc) >= 16) ch->def_cap = (uint64_t)atoll(dc); } // tests: a small parking list
	const char *e = getenv("BFC_B200_TAB_LOG2"); // optional pre-sizing: log2(slots)
	if (e && atoi(e) - l_pre > ch->rbits && atoi(e) <= 36) ch->rbits = atoi(e) - l_pre;
	if (cudaMalloc(&ch->slots, tab_capacity(ch) * 8) != cudaSuccess ||
		cudaMalloc(&ch->counters, 64) != cudaSuccess ||
		cudaMalloc(&ch->deferred, ch->def_cap * 16) != cudaSuccess) {
		bfcg_fail(__func__, "cudaMalloc(k-mer table)", cudaErrorMemoryAllocation);
		cudaFree(ch->slots); cudaFree(ch->counters); cudaFree(ch->deferred); free(ch);
		return 0;
	}
	cudaMemsetAsync(ch->slots, 0, tab_capacity(ch) * 8, rt.stream);
	cudaMemsetAsync(ch->counters, 0, 64, rt.stream);
	cudaStreamSynchronize(rt.stream);
	return ch;
}

// reference htab.c:36-43
void bfc_ch_destroy(bfc_ch_t *ch)
{
	if (ch == 0) return;
	cudaFree(ch->slots); cudaFree(ch->counters); cudaFree(ch->deferred);
	free(ch);
}

// reference htab.c:60-82; never busy, so `forced` has no effect and the result is always 0
int bfc_ch_insert(bfc_ch_t *ch, const uint64_t x[2], int is_high, int forced)
{
	(void)forced;
	if (bfcg_rt_init() != BFCG_OK) return -1;
	BfcgRuntime &rt = bfcg_rt();
	if (bfcg_tab_reserve(ch, 1) != BFCG_OK) return -1;
	k_tab_insert1<<<1, 1, 0, rt.stream>>>(tab_view(ch), x[0], x[1], is_high != 0);
	++rt.n_launches;
	if (cudaStreamSynchronize(rt.stream) != cudaSuccess) { bfcg_fail(__func__, "kernel", cudaGetLastError()); return -1; }
	return bfcg_tab_drain_deferred(ch) == BFCG_OK ? 0 : -1;
}

int bfcg_ch_get_batch(const bfc_ch_t *ch, int where, uint64_t n, const uint64_t *y, int32_t *out)
{
	if (bfcg_rt_init() != BFCG_OK) return BFCG_ERR_CUDA;
	BfcgRuntime &rt = bfcg_rt();
	if (n == 0) return BFCG_OK;
	const uint64_t *d_y = y;
	int32_t *d_out = out;
	if (where == BFCG_HOST) {
		uint8_t *a = (uint8_t*)bfcg_arena(n * 16 + n * 4 + 256);
		if (!a) return BFCG_ERR_NOMEM;
		d_y = (const uint64_t*)a, d_out = (int32_t*)(a + n * 16);
		BFCG_CUDA(cudaMemcpyAsync((void*)d_y, y, n * 16, cudaMemcpyHostToDevice, rt.stream));
	}
	k_tab_get<<<(unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)rt.sm_count * 32), 256, 0, rt.stream>>>(tab_view(ch), d_y, n, d_out);
	BFCG_LAUNCH_CHECK();
	if (where == BFCG_HOST) BFCG_CUDA(cudaMemcpyAsync(out, d_out, n * 4, cudaMemcpyDeviceToHost, rt.stream));
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	return BFCG_OK;
}

// reference htab.c:84-92
int bfc_ch_get(const bfc_ch_t *ch, const uint64_t x[2])
{
	int32_t out = -1;
	if (bfcg_ch_get_batch(ch, BFCG_HOST, 1, x, &out) != BFCG_OK) return -1;
	return out;
}

// reference htab.c:94-99
int bfc_ch_kmer_occ(const bfc_ch_t *ch, const bfc_kmer_t *z)
{
	if (bfcg_rt_init() != BFCG_OK) return -1;
	BfcgRuntime &rt = bfcg_rt();
	int32_t *d = (int32_t*)bfcg_arena(256), out = -1;
	if (!d) return -1;
	k_tab_kmer_occ1<<<1, 1, 0, rt.stream>>>(tab_view(ch), *z, d);
	++rt.n_launches;
	cudaMemcpyAsync(&out, d, 4, cudaMemcpyDeviceToHost, rt.stream);
	if (cudaStreamSynchronize(rt.stream) != cudaSuccess) { bfcg_fail(__func__, "kernel", cudaGetLastError()); return -1; }
	return out;
}

// reference htab.c:101-108
uint64_t bfc_ch_count(const bfc_ch_t *ch)
{
	unsigned long long c[2] = {0, 0};
	if (bfcg_rt_init() != BFCG_OK) return 0;
	tab_read_counters(ch, c);
	return c[0];
}

// reference htab.c:110-127
int bfc_ch_hist(const bfc_ch_t *ch, uint64_t cnt[256], uint64_t high[64])
{
	memset(cnt, 0, 256 * 8);
	memset(high, 0, 64 * 8);
	if (bfcg_rt_init() != BFCG_OK) return -1;
	BfcgRuntime &rt = bfcg_rt();
	unsigned long long *d = (unsigned long long*)bfcg_arena(320 * 8), h[320];
	if (!d) return -1;
	cudaMemsetAsync(d, 0, 320 * 8, rt.stream);
	{ KTime kt(KT_TAB_HIST); k_tab_hist<<<rt.sm_count * 8, 256, 0, rt.stream>>>(ch->slots, tab_capacity(ch), d); }
	++rt.n_launches;
	cudaMemcpyAsync(h, d, 320 * 8, cudaMemcpyDeviceToHost, rt.stream);
	if (cudaStreamSynchronize(rt.stream) != cudaSuccess) { bfcg_fail(__func__, "kernel", cudaGetLastError()); return -1; }
	for (int i = 0; i < 256; ++i) cnt[i] = h[i];
	for (int i = 0; i < 64; ++i) high[i] = h[256 + i];
	int max_i = -1;
	uint64_t max = 0;
	for (int i = 3; i < 256; ++i)
		if (cnt[i] > max) max = cnt[i], max_i = i;
	return max_i;
}

int bfc_ch_get_k(const bfc_ch_t *ch) { return ch->k; }
int bfcg_ch_l_pre(const bfc_ch_t *ch) { return ch->l_pre; }
int bfcg_ch_capacity_log2(const bfc_ch_t *ch) { return ch->l_pre - ch->own_bits + ch->rbits; }

// this table is shard `owner` of `n_owners` (1, 2, 4, 8) of a sharded count; must be called while it is empty
int bfcg_ch_set_shard(bfc_ch_t *ch, int n_owners, int owner)
{
	if (!ch || n_owners < 1 || n_owners > 8 || (n_owners & (n_owners - 1)) || owner < 0 || owner >= n_owners)
		return bfcg_fail(__func__, "invalid arguments", cudaSuccess), BFCG_ERR_ARG;
	if (bfc_ch_count(ch) != 0) return bfcg_fail(__func__, "the table is not empty", cudaSuccess), BFCG_ERR_ARG;
	ch->req_owners = n_owners, ch->req_owner = owner;
	return BFCG_OK;
}

int bfcg_ch_clear(bfc_ch_t *ch)
{
	if (bfcg_rt_init() != BFCG_OK) return BFCG_ERR_CUDA;
	BfcgRuntime &rt = bfcg_rt();
	BFCG_CUDA(cudaMemsetAsync(ch->slots, 0, tab_capacity(ch) * 8, rt.stream));
	BFCG_CUDA(cudaMemsetAsync(ch->counters, 0, 16, rt.stream));
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	ch->prev_new = 0, ch->have_prev = 0;
	return BFCG_OK;
}

int bfcg_ch_reserve(bfc_ch_t *ch, uint64_t n_keys)
{
	if (bfcg_rt_init() != BFCG_OK) return BFCG_ERR_CUDA;
	return bfcg_tab_reserve(ch, n_keys);
}

uint64_t bfcg_ch_export(const bfc_ch_t *ch, uint32_t *sub, uint64_t *key)
{
	if (bfcg_rt_init() != BFCG_OK) return 0;
	BfcgRuntime &rt = bfcg_rt();
	unsigned long long c[2] = {0, 0};
	if (tab_read_counters(ch, c) != BFCG_OK) return 0;
	const uint64_t n = c[0];
	if (sub == 0 || key == 0 || n == 0) return n;
	uint8_t *a = (uint8_t*)bfcg_arena(n * 12 + 512);
	if (!a) return 0;
	unsigned long long *d_key = (unsigned long long*)a, *cursor = (unsigned long long*)(a + n * 8);
	uint32_t *d_sub = (uint32_t*)(a + n * 8 + 256);
	cudaMemsetAsync(cursor, 0, 8, rt.stream);
	k_tab_export<<<rt.sm_count * 8, 256, 0, rt.stream>>>(ch->slots, ch->rbits, ch->l_pre, ch->rot, tab_reg_hi(ch), tab_capacity(ch), d_sub, d_key, cursor);
	++rt.n_launches;
	std::vector<uint32_t> hs(n);
	std::vector<uint64_t> hk(n);
	cudaMemcpyAsync(hs.data(), d_sub, n * 4, cudaMemcpyDeviceToHost, rt.stream);
	cudaMemcpyAsync(hk.data(), d_key, n * 8, cudaMemcpyDeviceToHost, rt.stream);
	if (cudaStreamSynchronize(rt.stream) != cudaSuccess) { bfcg_fail(__func__, "kernel", cudaGetLastError()); return 0; }
	std::vector<uint64_t> order(n);
	for (uint64_t i = 0; i < n; ++i) order[i] = i;
	std::sort(order.begin(), order.end(), [&](uint64_t a_, uint64_t b_) {
		return hs[a_] != hs[b_] ? hs[a_] < hs[b_] : hk[a_] < hk[b_];
	});
	for (uint64_t i = 0; i < n; ++i) sub[i] = hs[order[i]], key[i] = hk[order[i]];
	return n;
}

// all entries as (sub-table index, slot) pairs in DEVICE arrays, unsorted (multi-GPU table exchange)
uint64_t bfcg_ch_export_device(const bfc_ch_t *ch, uint32_t *d_sub, uint64_t *d_key)
{
	if (bfcg_rt_init() != BFCG_OK) return 0;
	BfcgRuntime &rt = bfcg_rt();
	unsigned long long c[2] = {0, 0};
	if (tab_read_counters(ch, c) != BFCG_OK) return 0;
	if (d_sub == 0 || d_key == 0 || c[0] == 0) return c[0];
	unsigned long long *cursor = ch->counters + 3;
	cudaMemsetAsync(cursor, 0, 8, rt.stream);
	k_tab_export<<<rt.sm_count * 8, 256, 0, rt.stream>>>(ch->slots, ch->rbits, ch->l_pre, ch->rot, tab_reg_hi(ch), tab_capacity(ch), d_sub, (unsigned long long*)d_key, cursor);
	++rt.n_launches;
	if (cudaStreamSynchronize(rt.stream) != cudaSuccess) { bfcg_fail(__func__, "kernel", cudaGetLastError()); return 0; }
	return c[0];
}

// add n entries (none of them present yet) from DEVICE arrays
int bfcg_ch_import_device(bfc_ch_t *ch, uint64_t n, const uint32_t *d_sub, const uint64_t *d_key)
{
	int r;
	if ((r = bfcg_rt_init()) != BFCG_OK) return r;
	BfcgRuntime &rt = bfcg_rt();
	if (n == 0) return BFCG_OK;
	if ((r = bfcg_tab_reserve(ch, n)) != BFCG_OK) return r;
	unsigned long long *fail = ch->counters + 2, h_fail = 0;
	BFCG_CUDA(cudaMemsetAsync(fail, 0, 8, rt.stream));
	k_tab_put_raw<<<(unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)rt.sm_count * 32), 256, 0, rt.stream>>>(tab_view(ch), d_sub, (const unsigned long long*)d_key, n, fail);
	BFCG_LAUNCH_CHECK();
	BFCG_CUDA(cudaMemcpyAsync(&h_fail, fail, 8, cudaMemcpyDeviceToHost, rt.stream));
	BFCG_CUDA(cudaStreamSynchronize(rt.stream));
	if (h_fail) return bfcg_fail(__func__, "import found full regions", cudaSuccess);
	return BFCG_OK;
}

// reference htab.c:129-149.  Same container: u32 k, u32 l_pre, then per sub-table
// u32 n_buckets, u32 size, size raw u64 keys.  Key order inside a sub-table is
// ascending here (khash slot order there); n_buckets is the smallest khash size that
// holds `size` keys below its 0.75 load bound, so the reference's -r / hash2cnt accept it.
int bfc_ch_dump(const bfc_ch_t *ch, const char *fn)
{
	FILE *fp = strcmp(fn, "-") ? fopen(fn, "wb") : stdout;
	if (fp == 0) return -1;
	const uint64_t n = bfc_ch_count(ch);
	std::vector<uint32_t> sub(n ? n : 1);
	std::vector<uint64_t> key(n ? n : 1);
	if (n && bfcg_ch_export(ch, sub.data(), key.data()) != n) { if (fp != stdout) fclose(fp); return -1; }
	uint32_t t[2] = { (uint32_t)ch->k, (uint32_t)ch->l_pre };
	fwrite(t, 4, 2, fp);
	uint64_t i = 0;
	for (uint32_t s = 0; s < 1u << ch->l_pre; ++s) {
		uint64_t j = i;
		while (j < n && sub[j] == s) ++j;
		const uint32_t size = (uint32_t)(j - i);
		uint32_t nb = 0;
		if (size) for (nb = 4; size >= (nb >> 2) + (nb >> 1); nb <<= 1) {}
		t[0] = nb, t[1] = size;
		fwrite(t, 4, 2, fp);
		if (size) fwrite(&key[i], 8, size, fp);
		i = j;
	}
	fprintf(stderr, "[M::%s] dumpped the hash table to file '%s'.\n", __func__, fn);
	if (fp != stdout) fclose(fp);
	return 0;
}

// reference htab.c:151-176
bfc_ch_t *bfc_ch_restore(const char *fn)
{
	FILE *fp = fopen(fn, "rb");
	uint32_t t[2];
	if (fp == 0) return 0;
	if (fread(t, 4, 2, fp) != 2) { fclose(fp); return 0; }
	bfc_ch_t *ch = bfc_ch_init(t[0], t[1]);
	if (ch == 0 || (int)t[1] != ch->l_pre) { fclose(fp); bfc_ch_destroy(ch); return 0; }
	std::vector<uint32_t> sub;
	std::vector<uint64_t> key;
	bool whole = true; // a short read anywhere = a truncated or corrupt dump: no table (the reference asserts, htab.c:161-170)
	for (uint32_t s = 0; s < 1u << ch->l_pre; ++s) {
		if (fread(t, 4, 2, fp) != 2) { whole = false; break; }
		const size_t at = key.size();
		if (t[1] > t[0] || at + t[1] > (1ULL << 40)) { whole = false; break; } // more keys than buckets
		key.resize(at + t[1]);
		sub.resize(at + t[1], s);
		if (t[1] && fread(&key[at], 8, t[1], fp) != t[1]) { whole = false; break; }
	}
	fclose(fp);
	if (!whole) {
		fprintf(stderr, "[E::%s] '%s' is truncated or not a k-mer table dump\n", __func__, fn);
		bfc_ch_destroy(ch);
		return 0;
	}
	const uint64_t n = key.size();
	BfcgRuntime &rt = bfcg_rt();
	if (n) {
		if (bfcg_tab_reserve(ch, n) != BFCG_OK) { bfc_ch_destroy(ch); return 0; }
		uint8_t *a = (uint8_t*)bfcg_arena(n * 12 + 512);
		if (!a) { bfc_ch_destroy(ch); return 0; }
		unsigned long long *d_key = (unsigned long long*)a, *fail = (unsigned long long*)(a + n * 8), h_fail = 0;
		uint32_t *d_sub = (uint32_t*)(a + n * 8 + 256);
		cudaMemcpyAsync(d_key, key.data(), n * 8, cudaMemcpyHostToDevice, rt.stream);
		cudaMemcpyAsync(d_sub, sub.data(), n * 4, cudaMemcpyHostToDevice, rt.stream);
		cudaMemsetAsync(fail, 0, 8, rt.stream);
		k_tab_put_raw<<<(unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)rt.sm_count * 32), 256, 0, rt.stream>>>(tab_view(ch), d_sub, d_key, n, fail);
		++rt.n_launches;
		cudaMemcpyAsync(&h_fail, fail, 8, cudaMemcpyDeviceToHost, rt.stream);
		if (cudaStreamSynchronize(rt.stream) != cudaSuccess || h_fail) {
			bfcg_fail(__func__, "restore kernel", cudaGetLastError());
			bfc_ch_destroy(ch);
			return 0;
		}
	}
	fprintf(stderr, "[M::%s] restored the hash table from file '%s'.\n", __func__, fn);
	return ch;
}

} // extern "C"
