// flof_grid.cu -- Grid4d<T> storage-level operators on device-resident grids:
// element-wise ops, reductions, boundary fills, slices and resampling.
// ref: source/grid4d.{h,cpp}, source/grid.cpp:462-473, source/test.cpp:199-249.
//
// All of these are pure HBM streams: flat grid-stride kernels, 128-bit accesses where the
// element type allows, grid = a multiple of the SM count.
#include <float.h>
#include <math.h>

#include "flof_common.cuh"

// ------------------------------------------------------------------ element-wise ----------
// cells*elem floats are processed as float4 when possible (elem == 4 always; elem == 1 when
// the length is a multiple of 4), the scalar tail handles the rest.
template <int OP>
__global__ void k_binary(float *__restrict__ a, const float *__restrict__ b, int64_t n)
{
	const int64_t n4 = n >> 2;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	float4 *a4 = reinterpret_cast<float4 *>(a);
	const float4 *b4 = reinterpret_cast<const float4 *>(b);
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
		float4 x = a4[i];
		const float4 y = __ldg(b4 + i);
		if (OP == FLOF_OP_ADD) { x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; }
		if (OP == FLOF_OP_SUB) { x.x -= y.x; x.y -= y.y; x.z -= y.z; x.w -= y.w; }
		if (OP == FLOF_OP_MULT) { x.x *= y.x; x.y *= y.y; x.z *= y.z; x.w *= y.w; }
		if (OP == FLOF_OP_MIN) {  // ref: KnJoin a = min(a,b) levelset.cpp:114-117
			x.x = y.x < x.x ? y.x : x.x; x.y = y.y < x.y ? y.y : x.y;
			x.z = y.z < x.z ? y.z : x.z; x.w = y.w < x.w ? y.w : x.w;
		}
		a4[i] = x;
	}
	for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		float x = a[i];
		const float y = b[i];
		if (OP == FLOF_OP_ADD) x += y;
		if (OP == FLOF_OP_SUB) x -= y;
		if (OP == FLOF_OP_MULT) x *= y;
		if (OP == FLOF_OP_MIN) x = y < x ? y : x;
		a[i] = x;
	}
}

// UN: 0 a += f*b, 1 a *= f, 2 a += f, 3 a = f, 4 clamp(a, f.x, f.y)
// PERCOMP: f applies per Vec4 component (elem 4); otherwise f.x for every float
template <int UN, bool PERCOMP>
__global__ void k_unary(float *__restrict__ a, const float *__restrict__ b, int64_t n, float4 f)
{
	const int64_t n4 = n >> 2;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	float4 *a4 = reinterpret_cast<float4 *>(a);
	const float4 *b4 = reinterpret_cast<const float4 *>(b);
	const float4 g = PERCOMP ? f : make_float4(f.x, f.x, f.x, f.x);
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
		float4 x = (UN == 3) ? g : a4[i];
		if (UN == 0) {  // ref: Grid4dScaledAdd me += factor * other  grid4d.h:368-372
			const float4 y = __ldg(b4 + i);
			x.x += g.x * y.x; x.y += g.y * y.y; x.z += g.z * y.z; x.w += g.w * y.w;
		}
		if (UN == 1) { x.x *= g.x; x.y *= g.y; x.z *= g.z; x.w *= g.w; }
		if (UN == 2) { x.x += g.x; x.y += g.y; x.z += g.z; x.w += g.w; }
		if (UN == 4) {  // ref: clamp general.h:175 (val < min ? min : val > max ? max : val)
			x.x = x.x < f.x ? f.x : (x.x > f.y ? f.y : x.x);
			x.y = x.y < f.x ? f.x : (x.y > f.y ? f.y : x.y);
			x.z = x.z < f.x ? f.x : (x.z > f.y ? f.y : x.z);
			x.w = x.w < f.x ? f.x : (x.w > f.y ? f.y : x.w);
		}
		a4[i] = x;
	}
	for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		float x = (UN == 3) ? f.x : a[i];
		if (UN == 0) x += f.x * b[i];
		if (UN == 1) x *= f.x;
		if (UN == 2) x += f.x;
		if (UN == 4) x = x < f.x ? f.x : (x > f.y ? f.y : x);
		a[i] = x;
	}
}

__global__ void k_set_const_int(int *__restrict__ a, int64_t n, int v)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a[i] = v;
}

static int check_elem(flof_ctx *ctx, int elem, const void *a)
{
	FLOF_ARG(elem == 1 || elem == 4 || elem == 3, "elem must be 1, 3 or 4 (got %d)", elem);
	FLOF_ARG(((uintptr_t)a & 15) == 0, "grid pointer %p is not 16-byte aligned", a);
	return FLOF_OK;
}

extern "C" int flof_grid_binary(flof_ctx *ctx, float *a, const float *b, int64_t cells, int elem,
                                int op)
{
	FLOF_RET(check_elem(ctx, elem, a));
	FLOF_RET(check_elem(ctx, elem, b));
	int64_t c0, c1;
	flof_flat_range(ctx, cells, &c0, &c1);  // t-slab of this rank if the grid belongs to the sharded level
	a += c0 * elem;
	b += c0 * elem;
	const int64_t n = (c1 - c0) * elem;
	const int blocks = flof_flat_blocks(ctx, (n + 3) / 4, 8);
	switch (op) {
	case FLOF_OP_ADD: FLOF_LAUNCH(k_binary<FLOF_OP_ADD>, blocks, FLOF_BLOCK, 0, a, b, n); break;
	case FLOF_OP_SUB: FLOF_LAUNCH(k_binary<FLOF_OP_SUB>, blocks, FLOF_BLOCK, 0, a, b, n); break;
	case FLOF_OP_MULT: FLOF_LAUNCH(k_binary<FLOF_OP_MULT>, blocks, FLOF_BLOCK, 0, a, b, n); break;
	case FLOF_OP_MIN: FLOF_LAUNCH(k_binary<FLOF_OP_MIN>, blocks, FLOF_BLOCK, 0, a, b, n); break;
	default: return flof_fail(ctx, FLOF_ERR_ARG, "flof_grid_binary: unknown op %d", op);
	}
	return FLOF_OK;
}

template <int UN>
static int unary(flof_ctx *ctx, float *a, const float *b, int64_t cells, int elem, float4 f)
{
	FLOF_RET(check_elem(ctx, elem, a));
	int64_t c0, c1;
	flof_flat_range(ctx, cells, &c0, &c1);
	a += c0 * elem;
	if (b) b += c0 * elem;
	const int64_t n = (c1 - c0) * elem;
	const int blocks = flof_flat_blocks(ctx, (n + 3) / 4, 8);
	if (elem == 4)
		FLOF_LAUNCH((k_unary<UN, true>), blocks, FLOF_BLOCK, 0, a, b, n, f);
	else
		FLOF_LAUNCH((k_unary<UN, false>), blocks, FLOF_BLOCK, 0, a, b, n, f);
	return FLOF_OK;
}

extern "C" int flof_grid_add_scaled(flof_ctx *ctx, float *a, const float *b, int64_t cells,
                                    int elem, const float f[4])
{
	FLOF_RET(check_elem(ctx, elem, b));
	return unary<0>(ctx, a, b, cells, elem, make_float4(f[0], f[1], f[2], f[3]));
}
extern "C" int flof_grid_mult_const(flof_ctx *ctx, float *a, int64_t cells, int elem,
                                    const float f[4])
{
	return unary<1>(ctx, a, NULL, cells, elem, make_float4(f[0], f[1], f[2], f[3]));
}
extern "C" int flof_grid_add_const(flof_ctx *ctx, float *a, int64_t cells, int elem,
                                   const float f[4])
{
	return unary<2>(ctx, a, NULL, cells, elem, make_float4(f[0], f[1], f[2], f[3]));
}
extern "C" int flof_grid_set_const(flof_ctx *ctx, float *a, int64_t cells, int elem,
                                   const float f[4])
{
	return unary<3>(ctx, a, NULL, cells, elem, make_float4(f[0], f[1], f[2], f[3]));
}
extern "C" int flof_grid_clamp(flof_ctx *ctx, float *a, int64_t cells, int elem, float lo, float hi)
{
	return unary<4>(ctx, a, NULL, cells, elem, make_float4(lo, hi, 0.f, 0.f));
}
extern "C" int flof_grid_set_const_int(flof_ctx *ctx, int *a, int64_t cells, int v)
{
	FLOF_LAUNCH(k_set_const_int, flof_flat_blocks(ctx, cells, 8), FLOF_BLOCK, 0, a, cells, v);
	return FLOF_OK;
}

// ------------------------------------------------------------------ reductions ------------
// ref: kn4dMinReal/kn4dMaxReal/kn4dMinVec/kn4dMaxVec grid4d.cpp:143-191
template <int ELEM>
__global__ void k_min_max(const float *__restrict__ a, int64_t cells, flof_reduce_scratch *red)
{
	__shared__ float sh[32];
	float mn = FLT_MAX, mx = -FLT_MAX;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += stride) {
		float s;
		if (ELEM == 4) {
			const float4 v = __ldg(reinterpret_cast<const float4 *>(a) + i);
			s = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;  // normSquare vector4d.h:311
		} else {
			s = __ldg(a + i);
		}
		if (s < mn) mn = s;
		if (s > mx) mx = s;
	}
	mn = flof_block_min(mn, sh);
	mx = flof_block_max(mx, sh);
	if (threadIdx.x == 0) {
		red->fmin[blockIdx.x] = mn;
		red->fmax[blockIdx.x] = mx;
	}
	if (flof_last_block(&red->counter[0])) {
		mn = FLT_MAX;
		mx = -FLT_MAX;
		for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
			mn = fminf(mn, red->fmin[b]);
			mx = fmaxf(mx, red->fmax[b]);
		}
		mn = flof_block_min(mn, sh);
		mx = flof_block_max(mx, sh);
		if (threadIdx.x == 0) {
			red->out_f[0] = mn;
			red->out_f[1] = mx;
		}
	}
}

// device-side min/max left in ctx->red->out_f[0..1]; no sync (used by advectCfl / corrVels)
int flof_min_max_device(flof_ctx *ctx, const float *a, int64_t cells, int elem)
{
	int64_t c0, c1;
	flof_flat_range(ctx, cells, &c0, &c1);
	const int blocks = flof_flat_blocks(ctx, c1 - c0, 8);
	if (elem == 4)
		FLOF_LAUNCH(k_min_max<4>, blocks, FLOF_BLOCK, 0, a + c0 * 4, c1 - c0, ctx->red);
	else
		FLOF_LAUNCH(k_min_max<1>, blocks, FLOF_BLOCK, 0, a + c0, c1 - c0, ctx->red);
	if (c1 - c0 != cells) {  // sharded level: combine the slabs
		FLOF_RET(flof_allreduce_f32_min(ctx, &ctx->red->out_f[0], 1));
		FLOF_RET(flof_allreduce_f32_max(ctx, &ctx->red->out_f[1], 1));
	}
	return FLOF_OK;
}

extern "C" int flof_grid_min_max(flof_ctx *ctx, const float *a, int64_t cells, int elem,
                                 float out[3])
{
	FLOF_ARG(elem == 1 || elem == 4, "flof_grid_min_max: elem must be 1 or 4");
	FLOF_RET(flof_min_max_device(ctx, a, cells, elem));
	float *h = (float *)ctx->pinned;
	FLOF_CK(cudaMemcpyAsync(h, ctx->red->out_f, 2 * sizeof(float), cudaMemcpyDeviceToHost,
	                        ctx->stream));
	FLOF_CK(cudaStreamSynchronize(ctx->stream));
	if (elem == 4) {  // ref: grid4d.cpp:274-285
		out[0] = sqrtf(h[0]);
		out[1] = sqrtf(h[1]);
		out[2] = sqrtf(h[1]);
	} else {  // ref: grid4d.cpp:266-273
		out[0] = h[0];
		out[1] = h[1];
		out[2] = fmaxf(fabsf(h[0]), fabsf(h[1]));
	}
	return FLOF_OK;
}

__global__ void k_min_max_int(const int *__restrict__ a, int64_t cells, int *out)
{
	// tiny helper (Grid4d<int> is only the extrapolation marker): one block
	__shared__ int smn[FLOF_BLOCK], smx[FLOF_BLOCK];
	int mn = INT_MAX, mx = INT_MIN;
	for (int64_t i = threadIdx.x; i < cells; i += blockDim.x) {
		const int v = a[i];
		mn = v < mn ? v : mn;
		mx = v > mx ? v : mx;
	}
	smn[threadIdx.x] = mn;
	smx[threadIdx.x] = mx;
	__syncthreads();
	for (int o = blockDim.x / 2; o > 0; o >>= 1) {
		if (threadIdx.x < o) {
			smn[threadIdx.x] = min(smn[threadIdx.x], smn[threadIdx.x + o]);
			smx[threadIdx.x] = max(smx[threadIdx.x], smx[threadIdx.x + o]);
		}
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		out[0] = smn[0];
		out[1] = smx[0];
	}
}
extern "C" int flof_grid_min_max_int(flof_ctx *ctx, const int *a, int64_t cells, int out[2])
{
	FLOF_LAUNCH(k_min_max_int, 1, FLOF_BLOCK, 0, a, cells, ctx->red->out_i);
	int *h = (int *)ctx->pinned;
	FLOF_CK(cudaMemcpyAsync(h, ctx->red->out_i, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	FLOF_CK(cudaStreamSynchronize(ctx->stream));
	out[0] = h[0];
	out[1] = h[1];
	return FLOF_OK;
}

// ref: grid4dMaxDiff / Vec4 variant grid4d.cpp:419-464 (sum of |component diffs| in double)
__global__ void k_max_diff(const float *__restrict__ a, const float *__restrict__ b, int64_t cells,
                           int elem, flof_reduce_scratch *red)
{
	__shared__ double sh[32];
	double mx = 0.;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += stride) {
		double dsum = 0.;
		if (elem == 1) {
			dsum = (double)fabsf(a[i] - b[i]);
		} else if (elem == -1) {  // Grid4d<int>, ref grid4dMaxDiffInt grid4d.cpp:429-437
			dsum = fabs((double)((const int *)a)[i] - (double)((const int *)b)[i]);
		} else {
			for (int c = 0; c < elem; ++c) dsum += fabs((double)a[i * elem + c] - (double)b[i * elem + c]);
		}
		mx = dsum > mx ? dsum : mx;
	}
	// max via the sum helper's layout: use shuffles directly
	for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
	if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmax(mx, sh[w]);
		red->dsum[0][blockIdx.x] = mx;
	}
	if (flof_last_block(&red->counter[0])) {
		if (threadIdx.x == 0) {
			double m = 0.;
			for (int bb = 0; bb < (int)gridDim.x; ++bb) m = fmax(m, red->dsum[0][bb]);
			red->out_d[0] = m;
		}
	}
}
extern "C" int flof_grid_max_diff(flof_ctx *ctx, const float *a, const float *b, int64_t cells,
                                  int elem, double *out)
{
	FLOF_LAUNCH(k_max_diff, flof_flat_blocks(ctx, cells, 4), FLOF_BLOCK, 0, a, b, cells, elem, ctx->red);
	double *h = (double *)ctx->pinned;
	FLOF_CK(cudaMemcpyAsync(h, ctx->red->out_d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	FLOF_CK(cudaStreamSynchronize(ctx->stream));
	*out = h[0];
	return FLOF_OK;
}

// ref: debugGridAvg4d test.cpp:199-208 (double accumulate over bnd region)
__global__ void k_grid_avg(const float *__restrict__ phi, flof_dim4 d, int brd, flof_reduce_scratch *red)
{
	__shared__ double sh[32];
	int i, j, k, t;
	double v = 0.;
	if (flof_cell_ijkt(d, i, j, k, t) && flof_in_bounds(d, i, j, k, t, brd)) v = phi[flof_idx(d, i, j, k, t)];
	v = flof_block_sum(v, sh);
	const int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
	// many blocks: accumulate with atomics per slot (order-insensitive use: debug print only)
	if (threadIdx.x == 0) atomicAdd(&red->out_d[1], v);
	(void)bid;
}
extern "C" int flof_debug_grid_avg4d(flof_ctx *ctx, const float *phi, flof_dim4 d, int brd, float *out)
{
	FLOF_CK(cudaMemsetAsync(&ctx->red->out_d[1], 0, sizeof(double), ctx->stream));
	FLOF_LAUNCH(k_grid_avg, flof_grid4(d), FLOF_BLOCK, 0, phi, d, brd, ctx->red);
	double *h = (double *)ctx->pinned;
	FLOF_CK(cudaMemcpyAsync(h, &ctx->red->out_d[1], sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	FLOF_CK(cudaStreamSynchronize(ctx->stream));
	const double cnt = (double)(d.nx - 2 * brd) * (d.ny - 2 * brd) * (d.nz - 2 * brd) * (d.nt - 2 * brd);
	*out = (float)(h[0] * 1000000. / cnt);
	return FLOF_OK;
}

// ref: debugVelAvg4d test.cpp:210-219: mean of norm(v) over the bnd region, accumulated in double; norm() snaps to
// exactly 1 when |l - 1| < eps^2 (util/vector4d.h:304-308).  Debug print only: block partials are combined with atomics.
__global__ void k_vel_avg(const float4 *__restrict__ v, flof_dim4 d, int brd, flof_reduce_scratch *red)
{
	__shared__ double sh[32];
	int i, j, k, t;
	double a = 0.;
	if (flof_cell_ijkt(d, i, j, k, t) && flof_in_bounds(d, i, j, k, t, brd)) {
		const float4 q = v[flof_idx(d, i, j, k, t)];
		const float l = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
		a = (fabs((double)l - 1.) < (double)(FLOF_VECTOR_EPSILON * FLOF_VECTOR_EPSILON)) ? 1. : (double)sqrtf(l);
	}
	a = flof_block_sum(a, sh);
	if (threadIdx.x == 0) atomicAdd(&red->out_d[1], a);
}
extern "C" int flof_debug_vel_avg4d(flof_ctx *ctx, const float *v, flof_dim4 d, int brd, float *out)
{
	FLOF_CK(cudaMemsetAsync(&ctx->red->out_d[1], 0, sizeof(double), ctx->stream));
	FLOF_LAUNCH(k_vel_avg, flof_grid4(d), FLOF_BLOCK, 0, (const float4 *)v, d, brd, ctx->red);
	double *h = (double *)ctx->pinned;
	FLOF_CK(cudaMemcpyAsync(h, &ctx->red->out_d[1], sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	FLOF_CK(cudaStreamSynchronize(ctx->stream));
	const double cnt = (double)(d.nx - 2 * brd) * (d.ny - 2 * brd) * (d.nz - 2 * brd) * (d.nt - 2 * brd);
	*out = (float)(h[0] * 1. / cnt);
	return FLOF_OK;
}

// ref: calcObfDiff optflow4d.cpp:1762-1777 (3D): phiDiff = |phi1 - phi2|, velDiff = norm((vel1 - vel2, velt1 - velt2))
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_calc_obf_diff(const float *__restrict__ phi1, const float *__restrict__ phi2, float *__restrict__ phiDiff,
                    const float *__restrict__ vel1, const float *__restrict__ vel2, const float *__restrict__ velt1,
                    const float *__restrict__ velt2, float *__restrict__ velDiff, flof_dim3 d, int bnd)
{
	const unsigned p = blockIdx.x * FLOF_BLOCK + threadIdx.x;
	if (p >= (unsigned)(d.nx * d.ny)) return;
	const int j = (int)(p / (unsigned)d.nx), i = (int)(p - (unsigned)j * (unsigned)d.nx), k = (int)blockIdx.y;
	if (i < bnd || j < bnd || i >= d.nx - bnd || j >= d.ny - bnd) return;
	if (d.nz > 1 && (k < bnd || k >= d.nz - bnd)) return;  // FOR_IJK_BND: z border only for 3D grids
	const int64_t c = (int64_t)i + (int64_t)d.nx * (j + (int64_t)d.ny * k);
	phiDiff[c] = fabsf(phi1[c] - phi2[c]);
	const float d0 = vel1[3 * c] - vel2[3 * c], d1 = vel1[3 * c + 1] - vel2[3 * c + 1], d2 = vel1[3 * c + 2] - vel2[3 * c + 2];
	const float d3 = velt1[c] - velt2[c];
	const float l = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
	velDiff[c] = (fabs((double)l - 1.) < (double)(FLOF_VECTOR_EPSILON * FLOF_VECTOR_EPSILON)) ? 1.f : sqrtf(l);
}
extern "C" int flof_calc_obf_diff(flof_ctx *ctx, const float *phi1, const float *phi2, float *phiDiff, const float *vel1,
                                  const float *vel2, const float *velt1, const float *velt2, float *velDiff, flof_dim3 d, int bnd)
{
	FLOF_LAUNCH(k_calc_obf_diff, flof_grid3(d), FLOF_BLOCK, 0, phi1, phi2, phiDiff, vel1, vel2, velt1, velt2, velDiff, d, bnd);
	return FLOF_OK;
}

// ------------------------------------------------------------------ boundaries ------------
// ref: knSetBnd4d grid4d.cpp:355-363 (`<= w`: w+1 shells)
template <class T> __global__ void k_set_bound4d(T *__restrict__ a, flof_kd d, T v, int w)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t)) return;
	const bool bnd = (i <= w || i >= d.nx - 1 - w || j <= w || j >= d.ny - 1 - w || k <= w ||
	                  k >= d.nz - 1 - w || t <= w || t >= d.nt - 1 - w);
	if (bnd) a[flof_idx(d, i, j, k, t)] = v;
}
extern "C" int flof_grid4d_set_bound(flof_ctx *ctx, float *a, flof_dim4 d, int elem,
                                     const float v[4], int w)
{
	FLOF_ARG(elem == 1 || elem == 4, "flof_grid4d_set_bound: elem must be 1 or 4");
	dim3 g;
	const flof_kd kd = flof_kdim(ctx, d, &g);
	if (elem == 4)
		FLOF_LAUNCH(k_set_bound4d<float4>, g, FLOF_BLOCK, 0, (float4 *)a, kd, make_float4(v[0], v[1], v[2], v[3]), w);
	else
		FLOF_LAUNCH(k_set_bound4d<float>, g, FLOF_BLOCK, 0, a, kd, v[0], w);
	return FLOF_OK;
}
extern "C" int flof_grid4d_set_bound_int(flof_ctx *ctx, int *a, flof_dim4 d, int v, int w)
{
	dim3 g;
	const flof_kd kd = flof_kdim(ctx, d, &g);
	FLOF_LAUNCH(k_set_bound4d<int>, g, FLOF_BLOCK, 0, a, kd, v, w);
	return FLOF_OK;
}

// ref: knSetBnd4dNeumann grid4d.cpp:370-407.  Source cells are never boundary cells
// themselves (for sizes >= 2w+3), so the in-place update is race-free.
template <class T> __global__ void k_set_bound_neumann4d(T *a, flof_kd d, int w)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t)) return;
	bool set = false;
	int si = i, sj = j, sk = k, st = t;
	if (i <= w) { si = w + 1; set = true; }
	if (i >= d.nx - 1 - w) { si = d.nx - 1 - w - 1; set = true; }
	if (j <= w) { sj = w + 1; set = true; }
	if (j >= d.ny - 1 - w) { sj = d.ny - 1 - w - 1; set = true; }
	if (k <= w) { sk = w + 1; set = true; }
	if (k >= d.nz - 1 - w) { sk = d.nz - 1 - w - 1; set = true; }
	if (t <= w) { st = w + 1; set = true; }
	if (t >= d.nt - 1 - w) { st = d.nt - 1 - w - 1; set = true; }
	if (set) a[flof_idx(d, i, j, k, t)] = a[flof_idx(d, si, sj, sk, st)];
}
extern "C" int flof_grid4d_set_bound_neumann(flof_ctx *ctx, float *a, flof_dim4 d, int elem, int w)
{
	FLOF_ARG(elem == 1 || elem == 4, "flof_grid4d_set_bound_neumann: elem must be 1 or 4");
	// n >= 2w + 3 per axis: the source index (w+1 or n-2-w) is then no boundary index itself
	FLOF_ARG(d.nx >= 2 * w + 3 && d.ny >= 2 * w + 3 && d.nz >= 2 * w + 3 && d.nt >= 2 * w + 3,
	         "flof_grid4d_set_bound_neumann: grid too small for width %d", w);
	// sharded: the source slice of a t-border cell (w+1 / nt-2-w) must lie in the same slab
	if (flof_sharded(ctx, d.nt)) FLOF_ARG(ctx->sh.tb - ctx->sh.ta >= w + 2, "setBoundNeumann: slab thinner than the boundary width");
	dim3 g;
	const flof_kd kd = flof_kdim(ctx, d, &g);
	if (elem == 4)
		FLOF_LAUNCH(k_set_bound_neumann4d<float4>, g, FLOF_BLOCK, 0, (float4 *)a, kd, w);
	else
		FLOF_LAUNCH(k_set_bound_neumann4d<float>, g, FLOF_BLOCK, 0, a, kd, w);
	return FLOF_OK;
}

// ref: knSetBoundary grid.cpp:462-468 (3D Grid<Real>)
__global__ void k_set_bound3(float *__restrict__ a, flof_dim3 d, float v, int w)
{
	const unsigned p = blockIdx.x * FLOF_BLOCK + threadIdx.x;
	if (p >= (unsigned)(d.nx * d.ny)) return;
	const int j = p / d.nx, i = p - j * d.nx, k = blockIdx.y;
	const bool bnd = (i <= w || i >= d.nx - 1 - w || j <= w || j >= d.ny - 1 - w ||
	                  (d.nz > 1 && (k <= w || k >= d.nz - 1 - w)));
	if (bnd) a[(int64_t)i + (int64_t)d.nx * (j + (int64_t)d.ny * k)] = v;
}
extern "C" int flof_grid3_set_bound(flof_ctx *ctx, float *a, flof_dim3 d, float v, int w)
{
	FLOF_LAUNCH(k_set_bound3, flof_grid3(d), FLOF_BLOCK, 0, a, d, v, w);
	return FLOF_OK;
}

// ------------------------------------------------------------------ slices / components ---
__global__ void k_get_comp(const float *__restrict__ src, float *__restrict__ dst, int64_t cells, int c)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += stride)
		dst[i] = src[i * 4 + c];
}
__global__ void k_set_comp(const float *__restrict__ src, float *__restrict__ dst, int64_t cells, int c)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += stride)
		dst[i * 4 + c] = src[i];
}
extern "C" int flof_get_comp4d(flof_ctx *ctx, const float *src, float *dst, int64_t cells, int c)
{  // ref: getComp4d grid4d.cpp:346
	FLOF_ARG(c >= 0 && c < 4, "getComp4d: component %d out of range", c);
	FLOF_LAUNCH(k_get_comp, flof_flat_blocks(ctx, cells, 8), FLOF_BLOCK, 0, src, dst, cells, c);
	return FLOF_OK;
}
extern "C" int flof_set_comp4d(flof_ctx *ctx, const float *src, float *dst, int64_t cells, int c)
{  // ref: setComp4d grid4d.cpp:350
	FLOF_ARG(c >= 0 && c < 4, "setComp4d: component %d out of range", c);
	FLOF_LAUNCH(k_set_comp, flof_flat_blocks(ctx, cells, 8), FLOF_BLOCK, 0, src, dst, cells, c);
	return FLOF_OK;
}

// ref: knSetRegion4d grid4d.cpp:467-474 (float compare of integer coordinates)
template <class T>
__global__ void k_set_region(T *__restrict__ a, flof_dim4 d, float4 s, float4 e, T v)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t)) return;
	const float p[4] = { (float)i, (float)j, (float)k, (float)t };
	const float ss[4] = { s.x, s.y, s.z, s.w }, ee[4] = { e.x, e.y, e.z, e.w };
	for (int c = 0; c < 4; ++c)
		if (p[c] < ss[c] || p[c] > ee[c]) return;
	a[flof_idx(d, i, j, k, t)] = v;
}
extern "C" int flof_set_region4d(flof_ctx *ctx, float *dst, flof_dim4 d, int elem,
                                 const float start[4], const float end[4], const float value[4])
{
	const float4 s = make_float4(start[0], start[1], start[2], start[3]);
	const float4 e = make_float4(end[0], end[1], end[2], end[3]);
	if (elem == 4)
		FLOF_LAUNCH(k_set_region<float4>, flof_grid4(d), FLOF_BLOCK, 0, (float4 *)dst, d, s, e,
		            make_float4(value[0], value[1], value[2], value[3]));
	else
		FLOF_LAUNCH(k_set_region<float>, flof_grid4(d), FLOF_BLOCK, 0, dst, d, s, e, value[0]);
	return FLOF_OK;
}

extern "C" int flof_get_slice_from4d(flof_ctx *ctx, const float *src, flof_dim4 d, int srct, float *dst3)
{  // ref: getSliceFrom4d grid4d.cpp:488-499 -- a t-slice is contiguous: one D2D copy
	if (srct < 0 || srct >= d.nt) return FLOF_OK;  // reference returns silently
	const int64_t n3 = (int64_t)d.nx * d.ny * d.nz;
	return flof_memcpy_d2d(ctx, dst3, src + n3 * srct, sizeof(float) * n3);
}
__global__ void k_slice_vec(const float4 *__restrict__ src, float *__restrict__ xyz, float *__restrict__ tt, int64_t n3)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += stride) {
		const float4 v = src[i];
		xyz[i * 3 + 0] = v.x;
		xyz[i * 3 + 1] = v.y;
		xyz[i * 3 + 2] = v.z;
		if (tt) tt[i] = v.w;
	}
}
extern "C" int flof_get_slice_from4d_vec(flof_ctx *ctx, const float *src, flof_dim4 d, int srct,
                                         float *dst_xyz, float *dst_t)
{  // ref: getSliceFrom4dVec grid4d.cpp:501-515
	if (srct < 0 || srct >= d.nt) return FLOF_OK;
	const int64_t n3 = (int64_t)d.nx * d.ny * d.nz;
	FLOF_LAUNCH(k_slice_vec, flof_flat_blocks(ctx, n3, 8), FLOF_BLOCK, 0,
	            (const float4 *)src + n3 * srct, dst_xyz, dst_t, n3);
	return FLOF_OK;
}
extern "C" int flof_place_grid3d(flof_ctx *ctx, const float *src3, float *dst, flof_dim4 d, int dstt)
{  // ref: placeGrid3d grid4d.cpp:517-524
	if (dstt < 0 || dstt >= d.nt) return FLOF_OK;
	const int64_t n3 = (int64_t)d.nx * d.ny * d.nz;
	return flof_memcpy_d2d(ctx, dst + n3 * dstt, src3, sizeof(float) * n3);
}
__global__ void k_vec_from_scalar(const float *__restrict__ src, float4 *__restrict__ dst, int64_t cells)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += stride) {
		const float v = src[i];
		dst[i] = make_float4(v, v, v, v);
	}
}
extern "C" int flof_init_vec_from_scalar(flof_ctx *ctx, const float *src, float *dst, int64_t cells)
{  // ref: initVecFromScalar test.cpp:221
	FLOF_LAUNCH(k_vec_from_scalar, flof_flat_blocks(ctx, cells, 8), FLOF_BLOCK, 0, src, (float4 *)dst, cells);
	return FLOF_OK;
}
// ref: initTestCheckerboard test.cpp:231-249
__global__ void k_checker(float *__restrict__ val, float4 *__restrict__ vec, flof_dim4 d, int brd)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t) || !flof_in_bounds(d, i, j, k, t, brd)) return;
	const float num = 4.f;
	const int ci = (int)(i / (d.nx / num)), cj = (int)(j / (d.ny / num));
	const int ck = (int)(k / (d.nz / num)), ct = (int)(t / (d.nt / num));
	float v = -1.f;
	if ((ci + cj + ck + ct) % 2 == 1) v = 1.f;
	const int64_t c = flof_idx(d, i, j, k, t);
	val[c] = v;
	if (vec) vec[c] = make_float4(v, v, v, v);
}
extern "C" int flof_init_test_checkerboard(flof_ctx *ctx, float *val, float *vec, flof_dim4 d, int brd)
{
	FLOF_LAUNCH(k_checker, flof_grid4(d), FLOF_BLOCK, 0, val, (float4 *)vec, d, brd);
	return FLOF_OK;
}

// ------------------------------------------------------------------ resampling ------------
// ref: gridFactor4d grid4d.cpp:559-569 (host arithmetic, fp32 with the reference's roundings)
extern "C" void flof_grid_factor4d(const float s1[4], const float s2in[4], const float optSize[4],
                                   const float scale[4], float srcFac[4], float off[4])
{
	for (int c = 0; c < 4; ++c) {
		float s2 = s2in[c];
		if (optSize[c] > 0.) s2 = optSize[c];
		const volatile float q = s1[c] / s2;
		srcFac[c] = q / scale[c];
		const volatile float a = -off[c] * srcFac[c];
		const volatile float b = (float)(srcFac[c] * 0.5);
		off[c] = a + b;
	}
}

// ref: knInterpol4d grid4d.cpp:531-537 / KnInterpolateGrid4dTempl grid4d.h:463-471.
// One thread per target cell; gathers hit L1/L2 (down-sampling reads each source cell once,
// up-sampling re-reads a 16x smaller source).
template <class T>
__global__ void k_interpol4d(T *__restrict__ dst, flof_kd td, const T *__restrict__ src,
                             flof_dim4 sd, float4 fac, float4 off)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(td, i, j, k, t)) return;
	const float x = (float)i * fac.x + off.x, y = (float)j * fac.y + off.y;
	const float z = (float)k * fac.z + off.z, w = (float)t * fac.w + off.w;
	dst[flof_idx(td, i, j, k, t)] = flof_interpol4d<T>(src, sd, x, y, z, w);
}
extern "C" int flof_kn_interpol4d(flof_ctx *ctx, float *dst, flof_dim4 td, const float *src,
                                  flof_dim4 sd, int elem, const float srcFac[4], const float off[4])
{
	FLOF_ARG(elem == 1 || elem == 4, "flof_kn_interpol4d: elem must be 1 or 4");
	FLOF_ARG(sd.nx >= 2 && sd.ny >= 2 && sd.nz >= 2 && sd.nt >= 2, "interpolation source too small");
	const float4 f = make_float4(srcFac[0], srcFac[1], srcFac[2], srcFac[3]);
	const float4 o = make_float4(off[0], off[1], off[2], off[3]);
	dim3 g;
	const flof_kd kd = flof_kdim(ctx, td, &g);  // sharded target: only this rank's slices (source must be complete)
	if (elem == 4)
		FLOF_LAUNCH(k_interpol4d<float4>, g, FLOF_BLOCK, 0, (float4 *)dst, kd, (const float4 *)src, sd, f, o);
	else
		FLOF_LAUNCH(k_interpol4d<float>, g, FLOF_BLOCK, 0, dst, kd, src, sd, f, o);
	return FLOF_OK;
}
extern "C" int flof_interpolate_grid4d(flof_ctx *ctx, float *dst, flof_dim4 td, const float *src,
                                       flof_dim4 sd, int elem, const float offset[4],
                                       const float scale[4], const float size[4])
{
	const float s1[4] = { (float)sd.nx, (float)sd.ny, (float)sd.nz, (float)sd.nt };
	const float s2[4] = { (float)td.nx, (float)td.ny, (float)td.nz, (float)td.nt };
	float fac[4], off[4] = { offset[0], offset[1], offset[2], offset[3] };
	flof_grid_factor4d(s1, s2, size, scale, fac, off);
	return flof_kn_interpol4d(ctx, dst, td, src, sd, elem, fac, off);
}
extern "C" int flof_interpol_grid_templ(flof_ctx *ctx, float *dst, flof_dim4 td, const float *src,
                                        flof_dim4 sd, int elem)
{  // ref: optflow4d.cpp:40-57 with calcGridSizeFactor4d(Vec4i,Vec4i) grid4d.h:275-279
	float fac[4] = { (float)sd.nx / td.nx, (float)sd.ny / td.ny, (float)sd.nz / td.nz,
		             (float)sd.nt / td.nt };
	float off[4];
	for (int c = 0; c < 4; ++c) off[c] = (float)(fac[c] * 0.5);
	return flof_kn_interpol4d(ctx, dst, td, src, sd, elem, fac, off);
}
