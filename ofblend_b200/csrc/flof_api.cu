// flof_api.cu -- context, device memory pool and copies of the C ABI (include/flof_b200.h).
//
// ref: FluidSolver owns a stack-like pool of grid buffers (fluidsolver.cpp:24-53, 94-126).
// B200-native equivalent: one CUDA stream per context and the device's stream-ordered memory
// pool with an unlimited release threshold, so Grid4d temporaries are recycled without
// cudaMalloc/cudaFree round trips; 180 GB of HBM3e hold every grid of a 128^4 solve resident.
#include <stdarg.h>
#include <stdlib.h>

#include "flof_common.cuh"

static char g_create_err[512] = "";
void flof_seq_release(flof_ctx *ctx);  // flof_solve.cu

int flof_fail(flof_ctx *ctx, int code, const char *fmt, ...)
{
	char *buf = ctx ? ctx->err : g_create_err;
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, 512, fmt, ap);
	va_end(ap);
	return code;
}

extern "C" {

int flof_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
}

int flof_ctx_create(flof_ctx **out, int device)
{
	flof_ctx *ctx = NULL;
	if (!out) return flof_fail(NULL, FLOF_ERR_ARG, "flof_ctx_create: out is NULL");
	*out = NULL;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0)
		return flof_fail(NULL, FLOF_ERR_CUDA,
		                 "flof_ctx_create: no CUDA device (%s); libflof_b200 has no CPU fallback",
		                 e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
	if (device < 0 || device >= n)
		return flof_fail(NULL, FLOF_ERR_ARG, "flof_ctx_create: device %d out of range (0..%d)",
		                 device, n - 1);
	flof_ctx *c = (flof_ctx *)calloc(1, sizeof(flof_ctx));
	if (!c) return flof_fail(NULL, FLOF_ERR_NOMEM, "flof_ctx_create: out of host memory");
	c->device = device;
#define CCK(call)                                                                             \
	do {                                                                                      \
		cudaError_t e__ = (call);                                                             \
		if (e__ != cudaSuccess) {                                                             \
			flof_fail(NULL, FLOF_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));  \
			free(c);                                                                          \
			return FLOF_ERR_CUDA;                                                             \
		}                                                                                     \
	} while (0)
	CCK(cudaSetDevice(device));
	cudaDeviceProp prop;
	CCK(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10) {
		flof_fail(NULL, FLOF_ERR_CUDA,
		          "flof_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only",
		          device, prop.major, prop.minor);
		free(c);
		return FLOF_ERR_CUDA;
	}
	c->sm_count = prop.multiProcessorCount;
	c->nranks = 1;
	c->rank = 0;
	// break-even of a CG iteration (replicated: 41 us per Mcell; sharded over P: that / P + ~0.1 ms of halo and
	// reduction latency) lies near 3-4 Mcells: 32^4 stays replicated, 64^4 and up are cut along t
	c->shard_min_cells = (int64_t)1 << 22;
	c->opt.expol_mode = getenv("FLOF_EXPOL_MODE") ? atoi(getenv("FLOF_EXPOL_MODE")) : -1;  // -1: by grid size (flof_blur.cu)
	c->opt.expol_variant = getenv("FLOF_EXPOL_VARIANT") ? atoi(getenv("FLOF_EXPOL_VARIANT")) : 0;
	c->opt.apply_variant = getenv("FLOF_APPLY_VARIANT") ? atoi(getenv("FLOF_APPLY_VARIANT")) : 11;
	c->opt.apply_zchunk = getenv("FLOF_APPLY_ZCHUNK") ? atoi(getenv("FLOF_APPLY_ZCHUNK")) : -1;
	c->opt.dot_mode = getenv("FLOF_DOT_MODE") ? atoi(getenv("FLOF_DOT_MODE")) : 1;
	c->opt.no_p2p = getenv("FLOF_NO_P2P") ? 1 : 0;
	c->opt.sweep_overlap = getenv("FLOF_SWEEP_OVERLAP") ? atoi(getenv("FLOF_SWEEP_OVERLAP")) : 1;
	c->opt.host_result_rank = -1;
	c->opt.blur_mode = getenv("FLOF_BLUR_MODE") ? atoi(getenv("FLOF_BLUR_MODE")) : 0;
	CCK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	CCK(cudaDeviceGetDefaultMemPool(&c->pool, device));
	uint64_t thr = UINT64_MAX;
	CCK(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &thr));
	CCK(cudaMalloc((void **)&c->red, sizeof(flof_reduce_scratch)));
	CCK(cudaMemset(c->red, 0, sizeof(flof_reduce_scratch)));
	CCK(cudaMalloc((void **)&c->cg, sizeof(flof_cg_state)));
	CCK(cudaMemset(c->cg, 0, sizeof(flof_cg_state)));
	CCK(cudaMallocHost(&c->pinned, 4096));
	for (int i = 0; i < 4; ++i) CCK(cudaEventCreate(&c->ev[i]));
#undef CCK
	(void)ctx;
	*out = c;
	return FLOF_OK;
}

int flof_ctx_destroy(flof_ctx *ctx)
{
	if (!ctx) return FLOF_OK;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	flof_ctx_comm_destroy(ctx);
	flof_seq_release(ctx);
	for (int i = 0; i < 4; ++i) cudaEventDestroy(ctx->ev[i]);
	if (ctx->stream_hi) {
		for (int i = 0; i < 3; ++i) cudaEventDestroy(ctx->ev_ov[i]);
		cudaStreamDestroy(ctx->stream_hi);
	}
	cudaFreeHost(ctx->pinned);
	cudaFree(ctx->cg);
	cudaFree(ctx->red);
	cudaStreamDestroy(ctx->stream);
	free(ctx);
	return FLOF_OK;
}

const char *flof_last_error(flof_ctx *ctx) { return ctx ? ctx->err : g_create_err; }
void *flof_ctx_stream(flof_ctx *ctx) { return (void *)ctx->stream; }
long long flof_ctx_launch_count(flof_ctx *ctx) { return ctx->launches; }
int flof_ctx_sm_count(flof_ctx *ctx) { return ctx->sm_count; }
int flof_ctx_set_option(flof_ctx *ctx, const char *name, int value)
{
	FLOF_ARG(name != NULL, "flof_ctx_set_option: name is NULL");
	if (!strcmp(name, "expol_mode")) ctx->opt.expol_mode = value;
	else if (!strcmp(name, "expol_variant")) ctx->opt.expol_variant = value;
	else if (!strcmp(name, "apply_variant")) ctx->opt.apply_variant = value;
	else if (!strcmp(name, "dot_mode")) ctx->opt.dot_mode = value;
	else if (!strcmp(name, "no_p2p")) ctx->opt.no_p2p = value;
	else if (!strcmp(name, "apply_zchunk")) ctx->opt.apply_zchunk = value;
	else if (!strcmp(name, "sweep_overlap")) ctx->opt.sweep_overlap = value;
	else if (!strcmp(name, "host_result_rank")) ctx->opt.host_result_rank = value;
	else if (!strcmp(name, "blur_mode")) ctx->opt.blur_mode = value;
	else return flof_fail(ctx, FLOF_ERR_ARG, "flof_ctx_set_option: unknown option '%s'", name);
	return FLOF_OK;
}
int flof_ctx_get_option(flof_ctx *ctx, const char *name, int *value)
{
	FLOF_ARG(name != NULL && value != NULL, "flof_ctx_get_option: NULL argument");
	if (!strcmp(name, "expol_mode")) *value = ctx->opt.expol_mode;
	else if (!strcmp(name, "expol_variant")) *value = ctx->opt.expol_variant;
	else if (!strcmp(name, "apply_variant")) *value = ctx->opt.apply_variant;
	else if (!strcmp(name, "dot_mode")) *value = ctx->opt.dot_mode;
	else if (!strcmp(name, "no_p2p")) *value = ctx->opt.no_p2p;
	else if (!strcmp(name, "apply_zchunk")) *value = ctx->opt.apply_zchunk;
	else if (!strcmp(name, "sweep_overlap")) *value = ctx->opt.sweep_overlap;
	else if (!strcmp(name, "host_result_rank")) *value = ctx->opt.host_result_rank;
	else if (!strcmp(name, "blur_mode")) *value = ctx->opt.blur_mode;
	else return flof_fail(ctx, FLOF_ERR_ARG, "flof_ctx_get_option: unknown option '%s'", name);
	return FLOF_OK;
}
int flof_ctx_set_shard_min_cells(flof_ctx *ctx, int64_t cells)
{
	ctx->shard_min_cells = cells;
	return FLOF_OK;
}

int flof_malloc(flof_ctx *ctx, void **dptr, size_t bytes)
{
	FLOF_ARG(dptr != NULL, "flof_malloc: dptr is NULL");
	return flof_tmp_alloc(ctx, dptr, bytes, true);
}
int flof_free(flof_ctx *ctx, void *dptr) { return flof_tmp_free(ctx, dptr); }

int flof_memcpy_h2d(flof_ctx *ctx, void *dst, const void *src, size_t bytes)
{
	FLOF_CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
	return FLOF_OK;
}
int flof_memcpy_d2h(flof_ctx *ctx, void *dst, const void *src, size_t bytes)
{
	FLOF_CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	FLOF_CK(cudaStreamSynchronize(ctx->stream));
	return FLOF_OK;
}
int flof_memcpy_d2d(flof_ctx *ctx, void *dst, const void *src, size_t bytes)
{
	FLOF_CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
	return FLOF_OK;
}
int flof_memset0(flof_ctx *ctx, void *dst, size_t bytes)
{
	FLOF_CK(cudaMemsetAsync(dst, 0, bytes, ctx->stream));
	return FLOF_OK;
}
int flof_sync(flof_ctx *ctx)
{
	FLOF_CK(cudaStreamSynchronize(ctx->stream));
	return FLOF_OK;
}
int flof_host_alloc(flof_ctx *ctx, void **hptr, size_t bytes)
{
	FLOF_ARG(hptr != NULL, "flof_host_alloc: hptr is NULL");
	FLOF_CK(cudaMallocHost(hptr, bytes ? bytes : 16));
	return FLOF_OK;
}
int flof_host_free(flof_ctx *ctx, void *hptr)
{
	if (hptr) FLOF_CK(cudaFreeHost(hptr));
	return FLOF_OK;
}

int flof_profile_begin(flof_ctx *ctx)
{
	if (!ctx->prof_e0) {
		ctx->prof_cap = 1 << 16;
		ctx->prof_e0 = (cudaEvent_t *)calloc(ctx->prof_cap, sizeof(cudaEvent_t));
		ctx->prof_e1 = (cudaEvent_t *)calloc(ctx->prof_cap, sizeof(cudaEvent_t));
		ctx->prof_name = (const char **)calloc(ctx->prof_cap, sizeof(char *));
		ctx->prof_cells_of = (int64_t *)calloc(ctx->prof_cap, sizeof(int64_t));
		if (!ctx->prof_e0 || !ctx->prof_e1 || !ctx->prof_name || !ctx->prof_cells_of)
			return flof_fail(ctx, FLOF_ERR_NOMEM, "flof_profile_begin: out of host memory");
		for (int i = 0; i < ctx->prof_cap; ++i) {
			FLOF_CK(cudaEventCreate(&ctx->prof_e0[i]));
			FLOF_CK(cudaEventCreate(&ctx->prof_e1[i]));
		}
	}
	ctx->prof_n = 0;
	ctx->prof_on = 1;
	return FLOF_OK;
}
int flof_profile_end(flof_ctx *ctx, flof_kernel_stat *out, int max_out, int *n_out)
{
	ctx->prof_on = 0;
	FLOF_CK(cudaStreamSynchronize(ctx->stream));
	int n = 0;
	for (int i = 0; i < ctx->prof_n; ++i) {
		float ms = 0.f;
		FLOF_CK(cudaEventElapsedTime(&ms, ctx->prof_e0[i], ctx->prof_e1[i]));
		int k = -1;
		for (int q = 0; q < n; ++q)
			if (out[q].cells == ctx->prof_cells_of[i] && !strncmp(out[q].name, ctx->prof_name[i], sizeof(out[q].name) - 1)) {
				k = q;
				break;
			}
		if (k < 0) {
			if (n >= max_out) continue;
			k = n++;
			memset(&out[k], 0, sizeof(out[k]));
			strncpy(out[k].name, ctx->prof_name[i], sizeof(out[k].name) - 1);
			out[k].cells = ctx->prof_cells_of[i];
		}
		out[k].launches++;
		out[k].total_ms += ms;
	}
	for (int a = 0; a < n; ++a)  // sort by total time, descending
		for (int b = a + 1; b < n; ++b)
			if (out[b].total_ms > out[a].total_ms) {
				flof_kernel_stat t = out[a];
				out[a] = out[b];
				out[b] = t;
			}
	if (n_out) *n_out = n;
	return FLOF_OK;
}

} /* extern "C" */

int flof_tmp_alloc(flof_ctx *ctx, void **p, size_t bytes, bool zero)
{
	if (bytes == 0) bytes = 16;
	cudaError_t e = cudaMallocAsync(p, bytes, ctx->stream);
	if (e != cudaSuccess) {
		*p = NULL;
		return flof_fail(ctx, e == cudaErrorMemoryAllocation ? FLOF_ERR_NOMEM : FLOF_ERR_CUDA,
		                 "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
	}
	if (zero) FLOF_CK(cudaMemsetAsync(*p, 0, bytes, ctx->stream));
	return FLOF_OK;
}
int flof_tmp_free(flof_ctx *ctx, void *p)
{
	if (!p) return FLOF_OK;
	FLOF_CK(cudaFreeAsync(p, ctx->stream));
	return FLOF_OK;
}
