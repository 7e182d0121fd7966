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

int bfcg_rt_init()
{
	std::lock_guard<std::mutex> lock(g_rt_mutex);
	if (g_rt.ready) return BFCG_OK;
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
	BFCG_CUDA(cudaEventCreate(&g_rt.ev0));
	BFCG_CUDA(cudaEventCreate(&g_rt.ev1));
	g_rt.ready = true;
	return BFCG_OK;
}

void *bfcg_arena(size_t bytes)
{
	if (bytes <= g_rt.arena_bytes) return g_rt.arena;
	if (g_rt.arena) { cudaStreamSynchronize(g_rt.stream); cudaFree(g_rt.arena); g_rt.arena = 0; g_rt.arena_bytes = 0; }
	size_t want = bytes + (bytes >> 2);
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
