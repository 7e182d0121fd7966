// runtime.cu -- device selection, the one stream all phase kernels run on, grow-only
// scratch arena / pinned staging, error reporting.  Fails loudly: there is no CPU path.
#include "common.cuh"
#include <mutex>

static BfcgRuntime g_rt = {};
static std::mutex g_rt_mutex;

BfcgRuntime &bfcg_rt() { return g_rt; }

int bfcg_fail(const char *func, const char *what, cudaError_t e)
{
	snprintf(g_rt.err, sizeof(g_rt.err), "%s: %s: %s", func, what, e == cudaSuccess ? "failed" : cudaGetErrorString(e));
	fprintf(stderr, "[E::%s] %s: %s\n", func, what, e == cudaSuccess ? "failed" : cudaGetErrorString(e));
	return e == cudaErrorMemoryAllocation ? BFCG_ERR_NOMEM : BFCG_ERR_CUDA;
}

// Every public entry point comes through here.  The GPU steps of bfc_count / bfc_correct run on kt_pipeline worker
// threads, and a host thread's current CUDA device defaults to 0: make the engine's device current on whichever thread
// calls (once per thread).
static thread_local int tl_dev = -1;

int bfcg_rt_init()
{
	if (g_rt.ready && tl_dev == g_rt.dev) return BFCG_OK;
	std::lock_guard<std::mutex> lock(g_rt_mutex);
	if (g_rt.ready) {
		BFCG_CUDA(cudaSetDevice(g_rt.dev));
		tl_dev = g_rt.dev;
		return BFCG_OK;
	}
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0)
		return bfcg_fail(__func__, "no CUDA device (this library has no CPU fallback)", e);
	if (g_rt.dev < 0 || g_rt.dev >= n) g_rt.dev = 0;
	else if (g_rt.dev == 0) {
		const char *lr = getenv("LOCAL_RANK"); // one process per GPU under torchrun
		if (lr && atoi(lr) >= 0 && atoi(lr) < n) g_rt.dev = atoi(lr);
	}
	BFCG_CUDA(cudaSetDevice(g_rt.dev));
	cudaDeviceProp prop;
	BFCG_CUDA(cudaGetDeviceProperties(&prop, g_rt.dev));
	g_rt.sm_count = prop.multiProcessorCount;
	if (prop.major < 10)
		return bfcg_fail(__func__, "device is not sm_100 (kernels are built for sm_100a only)", cudaSuccess);
	// random 32-byte sector gathers dominate: do not let L2 widen the DRAM fetch
	cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
	BFCG_CUDA(cudaStreamCreateWithFlags(&g_rt.stream, cudaStreamNonBlocking));
	BFCG_CUDA(cudaStreamCreateWithFlags(&g_rt.copy_in, cudaStreamNonBlocking));
	BFCG_CUDA(cudaStreamCreateWithFlags(&g_rt.copy_out, cudaStreamNonBlocking));
	BFCG_CUDA(cudaEventCreate(&g_rt.ev0));
	BFCG_CUDA(cudaEventCreate(&g_rt.ev1));
	for (int i = 0; i < 3; ++i) {
		BFCG_CUDA(cudaEventCreateWithFlags(&g_rt.ev_in[i], cudaEventDisableTiming));
		BFCG_CUDA(cudaEventCreateWithFlags(&g_rt.ev_free[i], cudaEventDisableTiming));
		BFCG_CUDA(cudaEventCreateWithFlags(&g_rt.ev_done[i], cudaEventDisableTiming));
		BFCG_CUDA(cudaEventCreateWithFlags(&g_rt.ev_out[i], cudaEventDisableTiming));
	}
	g_rt.ready = true;
	tl_dev = g_rt.dev;
	return BFCG_OK;
}

static cudaEvent_t ev_get()
{
	cudaEvent_t e;
	if (!g_rt.ev_pool.empty()) { e = g_rt.ev_pool.back(); g_rt.ev_pool.pop_back(); return e; }
	cudaEventCreate(&e);
	return e;
}

int bfcg_kt_begin(int id)
{
	if (!g_rt.timing) return -1;
	BfcgRuntime::Span s;
	s.id = id, s.a = ev_get(), s.b = ev_get();
	cudaEventRecord(s.a, g_rt.stream);
	g_rt.spans.push_back(s);
	return (int)g_rt.spans.size() - 1;
}

void bfcg_kt_end(int idx)
{
	if (idx >= 0) cudaEventRecord(g_rt.spans[idx].b, g_rt.stream);
}

static void kt_collect()
{
	if (g_rt.spans.empty()) return;
	cudaStreamSynchronize(g_rt.stream);
	for (size_t i = 0; i < g_rt.spans.size(); ++i) {
		float ms = 0;
		const BfcgRuntime::Span &s = g_rt.spans[i];
		if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) g_rt.kt_ms[s.id] += ms, ++g_rt.kt_n[s.id];
		g_rt.ev_pool.push_back(s.a), g_rt.ev_pool.push_back(s.b);
	}
	g_rt.spans.clear();
}

void *bfcg_arena(size_t bytes)
{
	if (bytes <= g_rt.arena_bytes) return g_rt.arena;
	if (g_rt.arena) { cudaDeviceSynchronize(); cudaFree(g_rt.arena); g_rt.arena = 0; g_rt.arena_bytes = 0; }
	size_t want = bytes < ((size_t)1 << 30) ? bytes + (bytes >> 2) : bytes; // head-room only for small arenas
	if (cudaMalloc(&g_rt.arena, want) != cudaSuccess) {
		want = bytes;
		cudaGetLastError();
		if (cudaMalloc(&g_rt.arena, want) != cudaSuccess) {
			bfcg_fail(__func__, "cudaMalloc(scratch arena)", cudaErrorMemoryAllocation);
			g_rt.arena = 0;
			return 0;
		}
	}
	g_rt.arena_bytes = want;
	return g_rt.arena;
}

uint8_t *bfcg_pinned(int which, size_t bytes)
{
	if (bytes <= g_rt.pin_bytes[which]) return g_rt.pin[which];
	if (g_rt.pin[which]) { cudaStreamSynchronize(g_rt.stream); cudaFreeHost(g_rt.pin[which]); g_rt.pin[which] = 0; g_rt.pin_bytes[which] = 0; }
	if (cudaMallocHost(&g_rt.pin[which], bytes) != cudaSuccess) {
		bfcg_fail(__func__, "cudaMallocHost(staging)", cudaErrorMemoryAllocation);
		g_rt.pin[which] = 0;
		return 0;
	}
	g_rt.pin_bytes[which] = bytes;
	return g_rt.pin[which];
}

extern "C" {

int bfcg_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

int bfcg_set_device(int dev)
{
	if (g_rt.ready && dev != g_rt.dev)
		return bfcg_fail(__func__, "device already selected for this process", cudaSuccess);
	g_rt.dev = dev;
	return bfcg_rt_init();
}

int bfcg_sync(void)
{
	if (bfcg_rt_init() != BFCG_OK) return BFCG_ERR_CUDA;
	BFCG_CUDA(cudaStreamSynchronize(g_rt.stream));
	return BFCG_OK;
}

void bfcg_set_timing(int on) { g_rt.timing = on != 0; }

// per-kernel device time since the last call (indices: KT_* in common.cuh); resets the accumulators
int bfcg_kernel_times(double *ms, uint64_t *launches, int n)
{
	kt_collect();
	for (int i = 0; i < n && i < KT_N; ++i) ms[i] = g_rt.kt_ms[i], launches[i] = g_rt.kt_n[i];
	memset(g_rt.kt_ms, 0, sizeof(g_rt.kt_ms));
	memset(g_rt.kt_n, 0, sizeof(g_rt.kt_n));
	return KT_N;
}

// step timing on the engine's own stream (torch.cuda.Event only sees torch's stream)
int bfcg_event_record(int slot)
{
	if (bfcg_rt_init() != BFCG_OK || slot < 0 || slot >= 8) return BFCG_ERR_ARG;
	if (!g_rt.user_ev[slot]) BFCG_CUDA(cudaEventCreate(&g_rt.user_ev[slot]));
	BFCG_CUDA(cudaEventRecord(g_rt.user_ev[slot], g_rt.stream));
	return BFCG_OK;
}

double bfcg_event_elapsed_ms(int a, int b)
{
	float ms = -1;
	if (a < 0 || b < 0 || a >= 8 || b >= 8 || !g_rt.user_ev[a] || !g_rt.user_ev[b]) return -1;
	if (cudaEventSynchronize(g_rt.user_ev[b]) != cudaSuccess) return -1;
	if (cudaEventElapsedTime(&ms, g_rt.user_ev[a], g_rt.user_ev[b]) != cudaSuccess) return -1;
	return ms;
}

const char *bfcg_last_error(void) { return g_rt.err; }

void *bfcg_dev_alloc(uint64_t bytes)
{
	void *p = 0;
	if (bfcg_rt_init() != BFCG_OK) return 0;
	if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { bfcg_fail(__func__, "cudaMalloc", cudaErrorMemoryAllocation); return 0; }
	return p;
}

void bfcg_dev_free(void *p) { if (p) cudaFree(p); }

int bfcg_h2d(void *dst, const void *src, uint64_t bytes)
{
	if (bfcg_rt_init() != BFCG_OK) return BFCG_ERR_CUDA;
	BFCG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_rt.stream));
	BFCG_CUDA(cudaStreamSynchronize(g_rt.stream));
	return BFCG_OK;
}

int bfcg_d2h(void *dst, const void *src, uint64_t bytes)
{
	if (bfcg_rt_init() != BFCG_OK) return BFCG_ERR_CUDA;
	BFCG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_rt.stream));
	BFCG_CUDA(cudaStreamSynchronize(g_rt.stream));
	return BFCG_OK;
}

void *bfcg_host_alloc_pinned(uint64_t bytes)
{
	void *p = 0;
	if (bfcg_rt_init() != BFCG_OK) return 0;
	if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { bfcg_fail(__func__, "cudaMallocHost", cudaErrorMemoryAllocation); return 0; }
	return p;
}

void bfcg_host_free_pinned(void *p) { if (p) cudaFreeHost(p); }

} // extern "C"

// globals of the reference's bfc.c:13-15, owned by the library so that every host
// program linking it (the CLI, a patched reference, the Python mirror) shares them
extern "C" {
int bfc_verbose = 3;
double bfc_real_time = 0;
bfc_kmer_t bfc_kmer_null = {{0, 0, 0, 0}};
}
