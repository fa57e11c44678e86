// flof_advect.cu -- 4D semi-Lagrangian quadrilinear advection.
// ref: semiLagrange4d / advectSemiLagrange4d / advect4d / advectCflHelper4d
//      optflow4d.cpp:1275-1321 and advectCflTemplate<Grid4d...> :2170-2180.
//
// Gather kernel: one thread per destination cell, coalesced 128-bit read of vel and
// coalesced write of dst; the 16 corner reads of src go through L1/L2 (neighbouring threads
// back-trace to neighbouring positions, so corner reads of a warp fall into a few lines).
// Algorithmic traffic: vel 16 B + src elem*4 B + dst elem*4 B per cell.
#include <math.h>

#include "flof_common.cuh"

int flof_min_max_device(flof_ctx *ctx, const float *a, int64_t cells, int elem);

template <class T> __device__ __forceinline__ T zero_of();
template <> __device__ __forceinline__ float zero_of<float>() { return 0.f; }
template <> __device__ __forceinline__ float4 zero_of<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }

template <class T>
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_semi_lagrange4d(const float4 *__restrict__ vel, const T *__restrict__ src, T *__restrict__ dst,
                      flof_kd d, float dt)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t)) return;
	const int64_t c = flof_idx(d, i, j, k, t);
	if (!flof_in_bounds(d, i, j, k, t, 1)) {
		dst[c] = zero_of<T>();  // fresh zero grid + bnd=1 kernel + swap (ref: :1287-1289)
		return;
	}
	const float4 v = __ldg(vel + c);
	// ref :1278: Vec4(i+0.5f, ...) - vel*dt
	const float x = ((float)i + 0.5f) - v.x * dt, y = ((float)j + 0.5f) - v.y * dt;
	const float z = ((float)k + 0.5f) - v.z * dt, w = ((float)t + 0.5f) - v.w * dt;
	dst[c] = flof_interpol4d<T>(src, d, x, y, z, w);
}

extern "C" int flof_semi_lagrange4d(flof_ctx *ctx, const float *vel, const float *src, float *dst,
                                    flof_dim4 d, int elem, float dt)
{
	FLOF_ARG(elem == 1 || elem == 4, "advect4d: Grid Type is not supported (only Real, Vec4)");
	FLOF_ARG(src != dst, "flof_semi_lagrange4d: dst must not alias src");
	FLOF_ARG(d.nx >= 2 && d.ny >= 2 && d.nz >= 2 && d.nt >= 2, "advect4d: grid too small");
	// sharded level: dst and vel are touched on this rank's slices only; src must be complete (the
	// back-traced positions leave the slab), see flof_advect_cfl4d_ex
	dim3 g;
	const flof_kd kd = flof_kdim(ctx, d, &g);
	if (elem == 4)
		FLOF_LAUNCH(k_semi_lagrange4d<float4>, g, FLOF_BLOCK, 0, (const float4 *)vel, (const float4 *)src, (float4 *)dst, kd, dt);
	else
		FLOF_LAUNCH(k_semi_lagrange4d<float>, g, FLOF_BLOCK, 0, (const float4 *)vel, src, dst, kd, dt);
	return FLOF_OK;
}

extern "C" int flof_advect4d(flof_ctx *ctx, const float *vel, float *grid, flof_dim4 d, int elem,
                             float dt)
{
	const size_t bytes = sizeof(float) * (size_t)elem * (size_t)flof_cells(d);
	void *fwd = NULL;
	FLOF_RET(flof_tmp_alloc(ctx, &fwd, bytes, false));
	int rc = flof_semi_lagrange4d(ctx, vel, grid, (float *)fwd, d, elem, dt);
	if (rc == FLOF_OK) rc = flof_memcpy_d2d(ctx, grid, fwd, bytes);
	flof_tmp_free(ctx, fwd);
	return rc;
}

__global__ void k_scale_vec4(const float4 *__restrict__ src, float4 *__restrict__ dst, int64_t cells, float f)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += stride) {
		float4 v = __ldg(src + i);
		v.x *= f; v.y *= f; v.z *= f; v.w *= f;
		dst[i] = v;
	}
}

int flof_advect_cfl4d_ex(flof_ctx *ctx, float cfl, const float *vel, float *grid, flof_dim4 d, int elem,
                         float velFactor, int grid_complete);

extern "C" int flof_advect_cfl4d(flof_ctx *ctx, float cfl, const float *vel, float *grid,
                                 flof_dim4 d, int elem, float velFactor)
{
	return flof_advect_cfl4d_ex(ctx, cfl, vel, grid, d, elem, velFactor, 0);
}

// grid_complete: on a sharded level, `grid` is known to hold valid data on ALL slices (e.g. a fresh
// copy of an input), so no all-gather is needed before the first gather pass.  The result is valid
// on this rank's slices only.
int flof_advect_cfl4d_ex(flof_ctx *ctx, float cfl, const float *vel, float *grid, flof_dim4 d, int elem,
                         float velFactor, int grid_complete)
{
	FLOF_ARG(elem == 1 || elem == 4, "advect4d: Grid Type is not supported (only Real, Vec4)");
	const int64_t cells = flof_cells(d);
	const size_t gbytes = sizeof(float) * (size_t)elem * (size_t)cells;
	// ref :2175-2177 velTmp = vel * velFactor  (skipped bit-exactly when the factor is 1)
	const float *v = vel;
	void *velTmp = NULL;
	if (velFactor != 1.0f) {
		FLOF_RET(flof_tmp_alloc(ctx, &velTmp, sizeof(float) * 4 * (size_t)cells, false));
		int64_t c0, c1;
		flof_flat_range(ctx, cells, &c0, &c1);
		FLOF_LAUNCH(k_scale_vec4, flof_flat_blocks(ctx, c1 - c0, 8), FLOF_BLOCK, 0, (const float4 *)vel + c0,
		            (float4 *)velTmp + c0, c1 - c0, velFactor);
		v = (const float *)velTmp;
	}
	// ref :1311-1316: maxVel = getMaxValue()*dt; steps = int(maxVel/cfl)+1; dt = 1/steps
	int steps = 1;
	if (cfl < 1e30f) {
		FLOF_RET(flof_min_max_device(ctx, v, cells, 4));
		float *h = (float *)ctx->pinned;
		FLOF_CK(cudaMemcpyAsync(h, ctx->red->out_f, 2 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
		FLOF_CK(cudaStreamSynchronize(ctx->stream));
		const float orgDt = 1.0f;
		const float maxVel = sqrtf(h[1]) * orgDt;
		steps = (int)(maxVel / cfl) + 1;
	}
	const float dt = 1.0f / (float)steps;
	void *fwd = NULL;
	FLOF_RET(flof_tmp_alloc(ctx, &fwd, gbytes, false));
	float *cur = grid, *nxt = (float *)fwd;
	int rc = FLOF_OK;
	const size_t slice_bytes = sizeof(float) * (size_t)elem * (size_t)d.nx * d.ny * d.nz;
	for (int s = 0; s < steps && rc == FLOF_OK; ++s) {
		if (!(s == 0 && grid_complete)) rc = flof_allgather_slabs(ctx, cur, d.nt, slice_bytes);
		if (rc == FLOF_OK) rc = flof_semi_lagrange4d(ctx, v, cur, nxt, d, elem, dt * 1.f);
		float *sw = cur; cur = nxt; nxt = sw;
	}
	if (rc == FLOF_OK && cur != grid) rc = flof_memcpy_d2d(ctx, grid, cur, gbytes);
	flof_tmp_free(ctx, fwd);
	flof_tmp_free(ctx, velTmp);
	return rc;
}
