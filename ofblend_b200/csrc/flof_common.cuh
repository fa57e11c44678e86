// flof_common.cuh -- shared declarations of libflof_b200.so (sm_100a only).
//
// Context object, launch helpers and the device-side interpolation primitives.
// Compile flags (see Makefile): -gencode arch=compute_100a,code=sm_100a -fmad=false.
// -fmad=false matters: the reference is built for baseline x86-64 (mul then add, each
// rounded), and the CG stopping iteration / projectCell branches depend on it.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/flof_b200.h"

#define FLOF_VECTOR_EPSILON (1e-6f)  // ref: util/vectorbase.h:53
#define FLOF_MAX_PARTIALS 2048       // upper bound of blocks in any reducing kernel

// device scratch for deterministic two-stage reductions ("last block finishes")
struct flof_reduce_scratch {
	double dsum[3][FLOF_MAX_PARTIALS];  // partial fp64 sums, three independent slots
	double asum[FLOF_MAX_PARTIALS];     // partial sums of |product| (sequential-order dot products, flof_seqsum.cuh)
	float fmax[FLOF_MAX_PARTIALS];
	float fmin[FLOF_MAX_PARTIALS];
	unsigned int counter[4];            // arrival counters (self-resetting)
	double out_d[4];                    // finished results
	float out_f[4];
	int out_i[4];
};

// device-resident CG state (ref: locals of GridCGOptflow4d::solve optflow4d.cpp:274-329)
struct flof_cg_state {
	double sigma[2];   // ring: sigma of iteration it lives in sigma[it & 1]
	double acc;        // accuracy * residual0
	double resIni;
	double alpha1;     // dot(srch, A srch) of the current iteration
	double sigmaNew;
	float residual;    // signed max of res
	float relResidual; // ret_residual
	int iter;          // completed iterations
	int done;          // 1 = stop (converged, early-out or failed)
	int status;        // 0 running, 1 converged, 2 early-out residual0 < eps, 3 sigma == 0 / NaN
	int seq_inexact;   // a sequential-order dot product exceeded its capacities on a range too large for the plain loop
};

// ---- NVLink peer mailboxes (flof_comm.cu): every rank owns one cudaMalloc'ed, IPC-exported buffer that all
// other ranks map; halos and the CG scalars travel as plain peer stores + system-scope flags.
#define FLOF_P2P_MAX 16
struct flof_p2p_dev {           // by-value kernel argument
	char *peer[FLOF_P2P_MAX];   // mailbox base of every rank in this process's address space (peer[rank] = own)
	int rank, nranks;
	unsigned int *err;          // device word, set when a spin-wait timed out
	unsigned int *ar_seq;       // device word: sequence number of the last all-reduce
	unsigned int *chain_seq;    // device word: sequence number of the last sequential-order dot product (flof_seqsum)
};
struct flof_mbox_hdr {
	unsigned int halo_flag[2][2];  // [0: from rank-1, 1: from rank+1][parity] = sequence number of the data
	unsigned int pad0[12];
	struct {
		double v[4];
		unsigned int seq;
		unsigned int pad[7];
	} ar[2][FLOF_P2P_MAX];         // [parity][source rank]: that rank's contribution to all-reduce #seq
	// sequential-order dot products (flof_seqsum_kernels.cuh): the exact running sum travels rank 0 -> 1 -> ... (chain),
	// the last rank stores the total into every rank's `total` slot
	struct {
		double v;
		unsigned int seq;
		unsigned int pad[5];
	} chain[2], total[2];          // [parity]
};
#define FLOF_MBOX_HDR_BYTES 4096

struct flof_seq;
struct flof_ctx {
	int device;
	int sm_count;
	cudaStream_t stream;
	cudaMemPool_t pool;
	flof_reduce_scratch *red;  // device
	flof_cg_state *cg;         // device
	struct flof_seq *seq;      // sequential-order dot products: descriptors, leaf records, piece pool (flof_seqsum.cuh)
	void *pinned;              // 4 KB pinned host scratch for scalar read-back
	cudaEvent_t ev[4];
	long long launches;
	char err[512];
	// optional per-launch CUDA-event timing (flof_profile_begin/end): bench.py uses it to measure
	// each kernel's average duration live, outside any profiler
	int prof_on, prof_n, prof_cap;
	cudaEvent_t *prof_e0, *prof_e1;
	const char **prof_name;
	int64_t *prof_cells_of;  // per launch: cells of the grid level being processed
	int64_t prof_cells;      // set by the multi-scale driver at each pyramid level
	// multi-GPU: one context per rank, NCCL communicator over NVLink (flof_comm.cu)
	void *comm;              // ncclComm_t
	int rank, nranks;
	struct {
		int enabled;          // 1 = mailboxes mapped on every rank
		char *mbox;           // own mailbox: header + 4 halo buffers [from][parity] of `cap` bytes
		size_t cap;
		flof_p2p_dev dev;
		unsigned int halo_seq;           // advances identically on every rank (SPMD call sequence)
		unsigned int *counter;           // device: arrival counter of the push kernel + error word
	} p2p;
	// kernel selection knobs (flof_ctx_set_option; defaults from FLOF_EXPOL_MODE / FLOF_EXPOL_VARIANT /
	// FLOF_APPLY_VARIANT at context creation).  Every choice is bit-identical; they exist for A/B timing and tests.
	struct {
		int expol_mode;     // 1 Vec4 work list (default), 0 component planes, 2 dense kernel
		int expol_variant;  // register budget / unrolling variant of the chosen extrapolation kernel
		int apply_variant;  // CG apply: 11 (default) = streaming hints + 6 CTAs/SM (40 registers, no spills); 7 = 8 CTAs/SM; 0, 1, 3, 5, 9, 10 other occupancy / unrolling points
		int no_p2p;         // 1: keep NCCL for halos and CG scalars (no NVLink peer mailboxes; tree-order dot products)
		int apply_zchunk;   // CG apply: z-planes per chunk of the leaf order (-1 = by grid size, 0 = plain index order)
		int sweep_overlap;  // sharded extrapolation sweeps: 1 (default) = boundary items + halo exchange on a high-priority
		                    // side stream, overlapped with the interior items; 0 = exchange, then one launch
		int host_result_rank;  // flof_optical_flow_multiscale4d_host on N ranks: -1 (default) every rank downloads the result,
		                       // r >= 0: only rank r does
		int blur_mode;      // Gaussian blur: 0 (default) = the reference's (2S+1)^4-tap sums, bit for bit; 1 = separable fp32
		                    // passes (opt-in, ~6e-6 rel-L2 per blur: measures what exactness costs)
		int dot_mode;       // CG dot products: 1 (default) = the reference's sequential summation order, bit for bit
		                    // (flof_seqsum); 0 = tree reductions (faster, last bits of the fp64 sums differ)
	} opt;
	// sharded extrapolation sweeps: a high-priority side stream for the boundary items + halo exchange, so that they
	// overlap the interior items on `stream` (flof_blur.cu); created on first use
	cudaStream_t stream_hi;
	cudaEvent_t ev_ov[3];
	int64_t shard_min_cells; // smaller pyramid levels are computed replicated on every rank
	// t-sharding of the pyramid level currently being processed (set by the multi-scale driver):
	// every rank keeps full-size grids in a global index space but computes and owns only the
	// t-slices [ta, tb); ghost slices are refreshed by NCCL send/recv where a stencil needs them.
	struct {
		int active;
		int nt;       // T of the sharded level: only grids with this T are sliced
		int64_t n3;   // nx*ny*nz of the level
		int ta, tb;
	} sh;
};

// high-priority side stream + three untimed events for the overlapped exchanges (sweeps: flof_blur.cu, CG: flof_solve.cu)
static inline int flof_side_stream_ensure(flof_ctx *ctx)
{
	if (ctx->stream_hi) return 0;
	int lo = 0, hi = 0;
	if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) return 1;
	if (cudaStreamCreateWithPriority(&ctx->stream_hi, cudaStreamNonBlocking, hi) != cudaSuccess) return 1;
	for (int i = 0; i < 3; ++i)
		if (cudaEventCreateWithFlags(&ctx->ev_ov[i], cudaEventDisableTiming) != cudaSuccess) return 1;
	return 0;
}

// slab of this rank for a grid with `nt` slices: [0, nt) unless the grid belongs to the sharded level
static inline void flof_slab(const flof_ctx *ctx, int nt, int *ta, int *tb)
{
	if (ctx->sh.active && nt == ctx->sh.nt) {
		*ta = ctx->sh.ta;
		*tb = ctx->sh.tb;
	} else {
		*ta = 0;
		*tb = nt;
	}
}
// same for flat kernels addressed by cell count
static inline void flof_flat_range(const flof_ctx *ctx, int64_t cells, int64_t *c0, int64_t *c1)
{
	if (ctx->sh.active && cells == ctx->sh.n3 * ctx->sh.nt) {
		*c0 = ctx->sh.n3 * ctx->sh.ta;
		*c1 = ctx->sh.n3 * ctx->sh.tb;
	} else {
		*c0 = 0;
		*c1 = cells;
	}
}
static inline bool flof_sharded(const flof_ctx *ctx, int nt) { return ctx->sh.active && nt == ctx->sh.nt; }

// communication helpers (flof_comm.cu); all are no-ops when the grid is not sharded
int flof_halo_exchange(flof_ctx *ctx, void *grid, int nt, size_t slice_bytes, int h);
int flof_p2p_ensure(flof_ctx *ctx, size_t need);  // collective: maps the peer mailboxes (halo buffers >= need bytes)
int flof_allgather_slabs(flof_ctx *ctx, void *grid, int nt, size_t slice_bytes);
int flof_allgather_bytes(flof_ctx *ctx, void *buf, size_t off, size_t n);
int flof_allreduce_f64_sum(flof_ctx *ctx, double *dev, int n);
int flof_allreduce_f32_max(flof_ctx *ctx, float *dev, int n);
int flof_allreduce_f32_min(flof_ctx *ctx, float *dev, int n);

static inline int flof_prof_pre(flof_ctx *ctx, const char *name)
{
	if (!ctx->prof_on || ctx->prof_n >= ctx->prof_cap) return -1;
	const int i = ctx->prof_n++;
	ctx->prof_name[i] = name;
	ctx->prof_cells_of[i] = ctx->prof_cells;
	cudaEventRecord(ctx->prof_e0[i], ctx->stream);
	return i;
}
static inline void flof_prof_post(flof_ctx *ctx, int i)
{
	if (i >= 0) cudaEventRecord(ctx->prof_e1[i], ctx->stream);
}

int flof_fail(flof_ctx *ctx, int code, const char *fmt, ...);

#define FLOF_CK(call)                                                                         \
	do {                                                                                      \
		cudaError_t e__ = (call);                                                             \
		if (e__ != cudaSuccess)                                                               \
			return flof_fail(ctx, FLOF_ERR_CUDA, "%s failed: %s (%s:%d)", #call,              \
			                 cudaGetErrorString(e__), __FILE__, __LINE__);                    \
	} while (0)

#define FLOF_RET(call)                                                                        \
	do {                                                                                      \
		int r__ = (call);                                                                     \
		if (r__ != FLOF_OK) return r__;                                                       \
	} while (0)

#define FLOF_ARG(cond, ...)                                                                   \
	do {                                                                                      \
		if (!(cond)) return flof_fail(ctx, FLOF_ERR_ARG, __VA_ARGS__);                        \
	} while (0)

// kernel<<<grid, block, smem, ctx->stream>>>(args) + launch accounting + error check
#define FLOF_LAUNCH(kernel, grid, block, smem, ...)                                           \
	do {                                                                                      \
		const int pi__ = flof_prof_pre(ctx, #kernel);                                         \
		kernel<<<(grid), (block), (smem), ctx->stream>>>(__VA_ARGS__);                        \
		flof_prof_post(ctx, pi__);                                                            \
		ctx->launches++;                                                                      \
		FLOF_CK(cudaGetLastError());                                                          \
	} while (0)

static inline int64_t flof_cells(flof_dim4 d) { return (int64_t)d.nx * d.ny * d.nz * d.nt; }
static inline int64_t flof_cells3(flof_dim3 d) { return (int64_t)d.nx * d.ny * d.nz; }

// temporaries come from the stream-ordered pool (the FluidSolver grid stack of the reference)
int flof_tmp_alloc(flof_ctx *ctx, void **p, size_t bytes, bool zero);
int flof_tmp_free(flof_ctx *ctx, void *p);

// ---- launch geometry -------------------------------------------------------------------
// "plane-linear" mapping for stencil/gather kernels: threads run over the x-y plane
// (coalesced along x, one 32-bit division per thread), blockIdx.y = k, blockIdx.z = t.
#define FLOF_BLOCK 256
static inline dim3 flof_grid4(flof_dim4 d)
{
	return dim3((unsigned)(((int64_t)d.nx * d.ny + FLOF_BLOCK - 1) / FLOF_BLOCK), (unsigned)d.nz,
	            (unsigned)d.nt);
}
// kernel-side dims of a (possibly sliced) launch: global sizes + the first t-slice this launch covers
struct flof_kd : flof_dim4 {
	int t0;
};
static inline flof_kd flof_kdim(const flof_ctx *ctx, flof_dim4 d, dim3 *grid)
{
	int ta, tb;
	flof_slab(ctx, d.nt, &ta, &tb);
	flof_kd k;
	k.nx = d.nx; k.ny = d.ny; k.nz = d.nz; k.nt = d.nt;
	k.t0 = ta;
	*grid = dim3((unsigned)(((int64_t)d.nx * d.ny + FLOF_BLOCK - 1) / FLOF_BLOCK), (unsigned)d.nz, (unsigned)(tb - ta));
	return k;
}
static inline dim3 flof_grid3(flof_dim3 d)
{
	return dim3((unsigned)(((int64_t)d.nx * d.ny + FLOF_BLOCK - 1) / FLOF_BLOCK), (unsigned)d.nz, 1);
}
// flat grid-stride kernels: a few CTAs per SM, a multiple of the SM count
static inline int flof_flat_blocks(flof_ctx *ctx, int64_t work_items, int per_sm)
{
	int64_t need = (work_items + FLOF_BLOCK - 1) / FLOF_BLOCK;
	int64_t cap = (int64_t)ctx->sm_count * per_sm;
	if (cap > FLOF_MAX_PARTIALS) cap = FLOF_MAX_PARTIALS;
	if (need < 1) need = 1;
	return (int)(need < cap ? need : cap);
}

#ifdef __CUDACC__

__device__ __forceinline__ bool flof_cell_ijkt(flof_dim4 d, int &i, int &j, int &k, int &t)
{
	const unsigned p = blockIdx.x * FLOF_BLOCK + threadIdx.x;
	if (p >= (unsigned)(d.nx * d.ny)) return false;
	j = (int)(p / (unsigned)d.nx);
	i = (int)(p - (unsigned)j * (unsigned)d.nx);
	k = (int)blockIdx.y;
	t = (int)blockIdx.z;
	return true;
}
__device__ __forceinline__ bool flof_cell_ijkt(const flof_kd &d, int &i, int &j, int &k, int &t)
{
	const unsigned p = blockIdx.x * FLOF_BLOCK + threadIdx.x;
	if (p >= (unsigned)(d.nx * d.ny)) return false;
	j = (int)(p / (unsigned)d.nx);
	i = (int)(p - (unsigned)j * (unsigned)d.nx);
	k = (int)blockIdx.y;
	t = (int)blockIdx.z + d.t0;
	return true;
}
__device__ __forceinline__ int64_t flof_idx(flof_dim4 d, int i, int j, int k, int t)
{
	return (int64_t)i + (int64_t)d.nx * (j + (int64_t)d.ny * (k + (int64_t)d.nz * t));
}
__device__ __forceinline__ bool flof_in_bounds(flof_dim4 d, int i, int j, int k, int t, int b)
{  // ref: Grid4dBase::isInBounds grid4d.h:319-326
	return i >= b && j >= b && i < d.nx - b && j < d.ny - b && k >= b && k < d.nz - b && t >= b &&
	       t < d.nt - b;
}

__device__ __forceinline__ float4 operator*(float4 a, float s)
{
	return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}
__device__ __forceinline__ float4 operator+(float4 a, float4 b)
{
	return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// quadrilinear weights + base index   ref: BUILD_INDEX_4D util/vector4d.h:422-475
struct flof_ipol4 {
	int64_t idx;
	float s0, s1, t0, t1, f0, f1, g0, g1;
};
__device__ __forceinline__ flof_ipol4 flof_build_index4(flof_dim4 d, float x, float y, float z,
                                                        float w)
{
	flof_ipol4 q;
	const float px = x - 0.5f, py = y - 0.5f, pz = z - 0.5f, pt = w - 0.5f;
	int xi = (int)px, yi = (int)py, zi = (int)pz, ti = (int)pt;
	q.s1 = px - (float)xi; q.s0 = 1.0f - q.s1;
	q.t1 = py - (float)yi; q.t0 = 1.0f - q.t1;
	q.f1 = pz - (float)zi; q.f0 = 1.0f - q.f1;
	q.g1 = pt - (float)ti; q.g0 = 1.0f - q.g1;
	if (px < 0.f) { xi = 0; q.s0 = 1.f; q.s1 = 0.f; }
	if (py < 0.f) { yi = 0; q.t0 = 1.f; q.t1 = 0.f; }
	if (pz < 0.f) { zi = 0; q.f0 = 1.f; q.f1 = 0.f; }
	if (pt < 0.f) { ti = 0; q.g0 = 1.f; q.g1 = 0.f; }
	if (xi >= d.nx - 1) { xi = d.nx - 2; q.s0 = 0.f; q.s1 = 1.f; }
	if (yi >= d.ny - 1) { yi = d.ny - 2; q.t0 = 0.f; q.t1 = 1.f; }
	if (zi >= d.nz - 1) { zi = d.nz - 2; q.f0 = 0.f; q.f1 = 1.f; }
	if (ti >= d.nt - 1) { ti = d.nt - 2; q.g0 = 0.f; q.g1 = 1.f; }
	q.idx = flof_idx(d, xi, yi, zi, ti);
	return q;
}

// ref: interpol4d util/vector4d.h:487-513 -- evaluation order ((y)x)z)t, fp32 mul/add
template <class T>
__device__ __forceinline__ T flof_interpol4d(const T *__restrict__ data, flof_dim4 d, float x,
                                             float y, float z, float w)
{
	const flof_ipol4 q = flof_build_index4(d, x, y, z, w);
	const int64_t sX = 1, sY = d.nx, sZ = (int64_t)d.nx * d.ny, sT = sZ * d.nz;
	const T *p = data + q.idx;
	const T a0 = (__ldg(p) * q.t0 + __ldg(p + sY) * q.t1) * q.s0 +
	             (__ldg(p + sX) * q.t0 + __ldg(p + sX + sY) * q.t1) * q.s1;
	const T a1 = (__ldg(p + sZ) * q.t0 + __ldg(p + sY + sZ) * q.t1) * q.s0 +
	             (__ldg(p + sX + sZ) * q.t0 + __ldg(p + sX + sY + sZ) * q.t1) * q.s1;
	const T *r = p + sT;
	const T b0 = (__ldg(r) * q.t0 + __ldg(r + sY) * q.t1) * q.s0 +
	             (__ldg(r + sX) * q.t0 + __ldg(r + sX + sY) * q.t1) * q.s1;
	const T b1 = (__ldg(r + sZ) * q.t0 + __ldg(r + sY + sZ) * q.t1) * q.s0 +
	             (__ldg(r + sX + sZ) * q.t0 + __ldg(r + sX + sY + sZ) * q.t1) * q.s1;
	return (a0 * q.f0 + a1 * q.f1) * q.g0 + (b0 * q.f0 + b1 * q.f1) * q.g1;
}

// same without the read-only path, for grids that are written by the running kernel's launch
template <class T>
__device__ __forceinline__ T flof_interpol4d_rw(const T *data, flof_dim4 d, float x, float y,
                                                float z, float w)
{
	const flof_ipol4 q = flof_build_index4(d, x, y, z, w);
	const int64_t sX = 1, sY = d.nx, sZ = (int64_t)d.nx * d.ny, sT = sZ * d.nz;
	const T *p = data + q.idx;
	const T a0 = (p[0] * q.t0 + p[sY] * q.t1) * q.s0 + (p[sX] * q.t0 + p[sX + sY] * q.t1) * q.s1;
	const T a1 = (p[sZ] * q.t0 + p[sY + sZ] * q.t1) * q.s0 +
	             (p[sX + sZ] * q.t0 + p[sX + sY + sZ] * q.t1) * q.s1;
	const T *r = p + sT;
	const T b0 = (r[0] * q.t0 + r[sY] * q.t1) * q.s0 + (r[sX] * q.t0 + r[sX + sY] * q.t1) * q.s1;
	const T b1 = (r[sZ] * q.t0 + r[sY + sZ] * q.t1) * q.s0 +
	             (r[sX + sZ] * q.t0 + r[sX + sY + sZ] * q.t1) * q.s1;
	return (a0 * q.f0 + a1 * q.f1) * q.g0 + (b0 * q.f0 + b1 * q.f1) * q.g1;
}

// ref: interpol util/interpol.h:57-116 (trilinear; z clamp only if nz > 1)
template <class T>
__device__ __forceinline__ T flof_interpol3d(const T *__restrict__ data, flof_dim3 d, float x,
                                             float y, float z)
{
	const float px = x - 0.5f, py = y - 0.5f, pz = z - 0.5f;
	int xi = (int)px, yi = (int)py, zi = (int)pz;
	float s1 = px - (float)xi, s0 = 1.0f - s1;
	float t1 = py - (float)yi, t0 = 1.0f - t1;
	float f1 = pz - (float)zi, f0 = 1.0f - f1;
	if (px < 0.f) { xi = 0; s0 = 1.f; s1 = 0.f; }
	if (py < 0.f) { yi = 0; t0 = 1.f; t1 = 0.f; }
	if (pz < 0.f) { zi = 0; f0 = 1.f; f1 = 0.f; }
	if (xi >= d.nx - 1) { xi = d.nx - 2; s0 = 0.f; s1 = 1.f; }
	if (yi >= d.ny - 1) { yi = d.ny - 2; t0 = 0.f; t1 = 1.f; }
	if (d.nz > 1) {
		if (zi >= d.nz - 1) { zi = d.nz - 2; f0 = 0.f; f1 = 1.f; }
	}
	const int64_t X = 1, Y = d.nx, Z = d.nz > 1 ? (int64_t)d.nx * d.ny : 0;  // mStrideZ of a 2D grid is 0 (grid.cpp:64)
	const T *p = data + ((int64_t)xi + Y * yi + Z * zi);
	return ((__ldg(p) * t0 + __ldg(p + Y) * t1) * s0 + (__ldg(p + X) * t0 + __ldg(p + X + Y) * t1) * s1) * f0 +
	       ((__ldg(p + Z) * t0 + __ldg(p + Y + Z) * t1) * s0 +
	        (__ldg(p + X + Z) * t0 + __ldg(p + X + Y + Z) * t1) * s1) *
	           f1;
}

// ---- block reductions (warp shuffle + shared), deterministic for a fixed block size -------
__device__ __forceinline__ double flof_warp_sum(double v)
{
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ float flof_warp_max(float v)
{
	for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_down_sync(0xffffffffu, v, o));
	return v;
}
__device__ __forceinline__ float flof_warp_min(float v)
{
	for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_down_sync(0xffffffffu, v, o));
	return v;
}
// result valid in thread 0
__device__ __forceinline__ double flof_block_sum(double v, double *sh /* >= 32 */)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
	v = flof_warp_sum(v);
	__syncthreads();
	if (lane == 0) sh[wid] = v;
	__syncthreads();
	v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
	if (wid == 0) v = flof_warp_sum(v);
	return v;
}
__device__ __forceinline__ float flof_block_max(float v, float *sh)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
	v = flof_warp_max(v);
	__syncthreads();
	if (lane == 0) sh[wid] = v;
	__syncthreads();
	v = (threadIdx.x < nw) ? sh[threadIdx.x] : -3.402823466e+38f;
	if (wid == 0) v = flof_warp_max(v);
	return v;
}
__device__ __forceinline__ float flof_block_min(float v, float *sh)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
	v = flof_warp_min(v);
	__syncthreads();
	if (lane == 0) sh[wid] = v;
	__syncthreads();
	v = (threadIdx.x < nw) ? sh[threadIdx.x] : 3.402823466e+38f;
	if (wid == 0) v = flof_warp_min(v);
	return v;
}
// "last block finishes": every block publishes its partial(s), the block that arrives last
// returns true and then reduces all partials in index order (deterministic).
__device__ __forceinline__ bool flof_last_block(unsigned int *counter)
{
	__shared__ bool s_last;
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0) {
		const unsigned int ticket = atomicAdd(counter, 1u);
		s_last = (ticket == gridDim.x * gridDim.y * gridDim.z - 1);
		if (s_last) *counter = 0;  // self-reset for the next launch
	}
	__syncthreads();
	if (s_last) __threadfence();
	return s_last;
}
#endif  // __CUDACC__
