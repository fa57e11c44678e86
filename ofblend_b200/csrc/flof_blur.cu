// flof_blur.cu -- truncated 4D Gaussian blur, 81-tap extrapolation blur, border reset.
// ref: gaussianWeight / knGaussianBlur / gaussianBlurGeneric optflow4d.cpp:128-174,
//      knCvExpolBlur4d :613-626 (driver loop :770-780), border reset :544-551.
//
// Parity mode: the Gaussian is evaluated exactly like the reference -- the full (2s+1)^4 box,
// taps visited in the order vt, zk, yj, xi, each `val += w*a` an fp32 multiply then add, the
// weight sum accumulated in the same order over the in-bounds taps only -- so the result is
// bit-identical to the CPU.  That makes this kernel FP32-issue bound rather than HBM bound
// (625 taps x 9 instructions per cell for s = 2); the data it touches per pass is still just
// 16 B in + 16 B out per cell and stays L1/L2 resident across the taps.
#include <math.h>
#include <stdlib.h>

#include "flof_common.cuh"

#define FLOF_BLUR_MAXS 8  // half-widths 1..4 are compile-time instantiations, 5..8 run the same kernel with a run-time width
// weight by integer squared distance, ref gaussianWeight :128-131: exp(-dSqr/(2.*sigma*sigma))
__constant__ float c_gauss_w[4 * FLOF_BLUR_MAXS * FLOF_BLUR_MAXS + 1];

#include "flof_blur_tiled.cuh"

template <class T> __device__ __forceinline__ T blur_zero();
template <> __device__ __forceinline__ float blur_zero<float>() { return 0.f; }
template <> __device__ __forceinline__ float4 blur_zero<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void blur_acc(float &v, float w, float a) { v += w * a; }
__device__ __forceinline__ void blur_acc(float4 &v, float w, const float4 &a)
{
	v.x += w * a.x; v.y += w * a.y; v.z += w * a.z; v.w += w * a.w;
}
__device__ __forceinline__ float blur_div(float v, float w) { return v / w; }
__device__ __forceinline__ float4 blur_div(const float4 &v, float w)
{
	return make_float4(v.x / w, v.y / w, v.z / w, v.w / w);
}

// one thread per interior cell (bnd 1); dst cells on the border shell are left untouched,
// which reproduces the zero border of the fresh tmp grid after pass 1 and the original border
// after pass 2 (ref :160-174) because the caller ping-pongs the same two buffers.
template <class T, int ST>
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_gauss_blur4d(const T *__restrict__ a, T *__restrict__ tmp, flof_kd d, int s_rt)
{
	const int S = ST ? ST : s_rt;  // ST == 0: run-time half-width (postVelBlur > 9)
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t)) return;
	// KERNEL(fourd, bnd = 1): a one-slice grid is not 4D in the reference (kernel.h:62-68), t is unbounded there --
	// the 3D instantiation gaussianBlurGeneric<Grid<Vec3>> runs through this kernel with d.nt == 1
	// (and a one-plane grid is 2D: z unbounded as well)
	if (d.nt == 1 ? !(i >= 1 && j >= 1 && i < d.nx - 1 && j < d.ny - 1 && (d.nz == 1 || (k >= 1 && k < d.nz - 1)))
	              : !flof_in_bounds(d, i, j, k, t, 1))
		return;
	T val = blur_zero<T>();
	float weight = 0.f;
	const int x0 = max(i - S, 0), x1 = min(i + S, d.nx - 1);
	for (int vt = t - S; vt <= t + S; ++vt) {
		if (vt < 0 || vt >= d.nt) continue;
		const int dt2 = (vt - t) * (vt - t);
		for (int zk = k - S; zk <= k + S; ++zk) {
			if (zk < 0 || zk >= d.nz) continue;
			const int dz2 = dt2 + (zk - k) * (zk - k);
			for (int yj = j - S; yj <= j + S; ++yj) {
				if (yj < 0 || yj >= d.ny) continue;
				const int dy2 = dz2 + (yj - j) * (yj - j);
				const T *row = a + flof_idx(d, 0, yj, zk, vt);
				for (int xi = x0; xi <= x1; ++xi) {
					const float wcurr = c_gauss_w[dy2 + (xi - i) * (xi - i)];
					weight += wcurr;
					blur_acc(val, wcurr, __ldg(row + xi));
				}
			}
		}
	}
	const int64_t c = flof_idx(d, i, j, k, t);
	if (weight > FLOF_VECTOR_EPSILON)
		tmp[c] = blur_div(val, weight);
	else
		tmp[c] = __ldg(a + c);
}

template <class T>
static int launch_blur(flof_ctx *ctx, const T *a, T *tmp, flof_dim4 d, int s)
{
	dim3 g;
	const flof_kd kd = flof_kdim(ctx, d, &g);
	switch (s) {
	case 1: FLOF_LAUNCH((k_gauss_blur4d<T, 1>), g, FLOF_BLOCK, 0, a, tmp, kd, s); break;
	case 2: FLOF_LAUNCH((k_gauss_blur4d<T, 2>), g, FLOF_BLOCK, 0, a, tmp, kd, s); break;
	case 3: FLOF_LAUNCH((k_gauss_blur4d<T, 3>), g, FLOF_BLOCK, 0, a, tmp, kd, s); break;
	case 4: FLOF_LAUNCH((k_gauss_blur4d<T, 4>), g, FLOF_BLOCK, 0, a, tmp, kd, s); break;
	default:
		if (s > FLOF_BLUR_MAXS) return flof_fail(ctx, FLOF_ERR_ARG, "gaussianBlur: kernel half-width %d > %d unsupported", s, FLOF_BLUR_MAXS);
		FLOF_LAUNCH((k_gauss_blur4d<T, 0>), g, FLOF_BLOCK, 0, a, tmp, kd, s);
		break;
	}
	return FLOF_OK;
}

// ---- opt-in SEPARABLE evaluation (option blur_mode = 1; NOT the default, NOT bit-exact) ------------------------------
// The Gaussian weight factorises, exp(-(dx^2+dy^2+dz^2+dt^2) / 2 sigma^2) = w(dx) w(dy) w(dz) w(dt), and so do the window
// clipping and the normalising weight sum; in real arithmetic the reference's (2S+1)^4-tap pass equals four 1D passes and
// one division.  In fp32 the result differs in the last bits (rel-L2 ~6e-6), which the projection
// amplifies, so the bit-exact kernels stay the default; this path exists to measure what exactness costs: 4 x 32 B/cell of
// HBM traffic per pass instead of 625 x 4 multiply-adds per cell.
// One axis pass: dst(c) = sum over the in-bounds taps k in [-S, S] of w(k) * src(c + k * stride); the last pass (t) divides
// by Wx(i) Wy(j) Wz(k) Wt(t) and writes interior cells only (KERNEL(fourd, bnd = 1): the caller ping-pongs like the exact path).
template <class T>
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_gauss_axis(const T *__restrict__ src, T *__restrict__ dst, flof_kd d, int S, int axis)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t)) return;
	const int pos = axis == 0 ? i : (axis == 1 ? j : (axis == 2 ? k : t));
	const int n = axis == 0 ? d.nx : (axis == 1 ? d.ny : (axis == 2 ? d.nz : d.nt));
	const int64_t stride = axis == 0 ? 1 : (axis == 1 ? d.nx : (axis == 2 ? (int64_t)d.nx * d.ny : (int64_t)d.nx * d.ny * d.nz));
	const int64_t c = flof_idx(d, i, j, k, t);
	if (axis == 3 && !flof_in_bounds(d, i, j, k, t, 1)) return;
	T val = blur_zero<T>();
	for (int q = max(-S, -pos); q <= min(S, n - 1 - pos); ++q) blur_acc(val, c_gauss_w[q * q], __ldg(src + c + q * stride));
	if (axis == 3) {
		float W = 1.f;
		const int ps[4] = { i, j, k, t }, ns[4] = { d.nx, d.ny, d.nz, d.nt };
#pragma unroll
		for (int ax = 0; ax < 4; ++ax) {
			float w1 = 0.f;
			for (int q = max(-S, -ps[ax]); q <= min(S, ns[ax] - 1 - ps[ax]); ++q) w1 += c_gauss_w[q * q];
			W *= w1;
		}
		dst[c] = W > FLOF_VECTOR_EPSILON ? blur_div(val, W) : __ldg(src + c);
	} else {
		dst[c] = val;
	}
}
// one blur pass cur -> oth (interior cells of oth); s1, s2: grid-sized scratch
template <class T>
static int gauss_pass_separable(flof_ctx *ctx, const T *cur, T *oth, T *s1, T *s2, flof_dim4 d, int s, size_t slice_bytes)
{
	dim3 g;
	const flof_kd kd = flof_kdim(ctx, d, &g);  // sharded level: the x, y, z passes run on the own slices ...
	FLOF_LAUNCH(k_gauss_axis<T>, g, FLOF_BLOCK, 0, cur, s1, kd, s, 0);
	FLOF_LAUNCH(k_gauss_axis<T>, g, FLOF_BLOCK, 0, (const T *)s1, s2, kd, s, 1);
	FLOF_LAUNCH(k_gauss_axis<T>, g, FLOF_BLOCK, 0, (const T *)s2, s1, kd, s, 2);
	FLOF_RET(flof_halo_exchange(ctx, s1, d.nt, slice_bytes, s));  // ... and the t pass needs +-s slices of THEIR result
	FLOF_LAUNCH(k_gauss_axis<T>, g, FLOF_BLOCK, 0, (const T *)s1, oth, kd, s, 3);
	return FLOF_OK;
}

int flof_gaussian_blur4d_impl(flof_ctx *ctx, float *a, flof_dim4 d, int elem, float sigma, int iter)
{
	FLOF_ARG(elem == 1 || elem == 4, "gaussianBlur: elem must be 1 or 4");
	int s = (int)(1. * sigma + 0.5);  // ref :165
	if (s == 0) s = 1;
	FLOF_ARG(s <= FLOF_BLUR_MAXS, "gaussianBlur: sigma %g too large (half-width %d > %d)", sigma, s, FLOF_BLUR_MAXS);
	float w[4 * FLOF_BLUR_MAXS * FLOF_BLUR_MAXS + 1];
	for (int q = 0; q <= 4 * s * s; ++q) {
		const float dSqr = (float)q;
		w[q] = (float)exp(-dSqr / (2. * sigma * sigma));
	}
	FLOF_CK(cudaMemcpyToSymbolAsync(c_gauss_w, w, sizeof(float) * (4 * s * s + 1), 0, cudaMemcpyHostToDevice,
	                                ctx->stream));
	const size_t bytes = sizeof(float) * (size_t)elem * (size_t)flof_cells(d);
	void *tmp = NULL;
	FLOF_RET(flof_tmp_alloc(ctx, &tmp, bytes, true));  // GRID tmp(parent): zero-initialised
	float *cur = a, *oth = (float *)tmp;
	int rc = FLOF_OK;
	const size_t slice_bytes = sizeof(float) * (size_t)elem * (size_t)d.nx * d.ny * d.nz;
	const bool separable = ctx->opt.blur_mode == 1 && d.nt > 1;
	void *s1 = NULL, *s2 = NULL;
	if (separable) {
		rc = flof_tmp_alloc(ctx, &s1, bytes, false);
		if (rc == FLOF_OK) rc = flof_tmp_alloc(ctx, &s2, bytes, false);
	}
	for (int numIt = 0; numIt < 2 * iter && rc == FLOF_OK; ++numIt) {
		if (separable) {
			rc = elem == 4 ? gauss_pass_separable<float4>(ctx, (const float4 *)cur, (float4 *)oth, (float4 *)s1, (float4 *)s2, d, s, slice_bytes)
			               : gauss_pass_separable<float>(ctx, cur, oth, (float *)s1, (float *)s2, d, s, slice_bytes);
			float *sw = cur; cur = oth; oth = sw;
			continue;
		}
		rc = flof_halo_exchange(ctx, cur, d.nt, slice_bytes, s);  // sharded level: +-s ghost slices of the source
		if (rc != FLOF_OK) break;
		int tiled = 0;
		if (elem == 4) {
			tiled = flof_launch_gauss_tiled(ctx, cur, oth, d, s, w, numIt == 0);
			if (tiled < 0) rc = FLOF_ERR_CUDA;
		}
		if (tiled != 0) {
			// done (or failed) in the register-tiled kernel
		} else if (elem == 4)
			rc = launch_blur<float4>(ctx, (const float4 *)cur, (float4 *)oth, d, s);
		else
			rc = launch_blur<float>(ctx, cur, oth, d, s);
		float *sw = cur; cur = oth; oth = sw;  // a.swap(tmp)
	}
	// 2*iter swaps: `cur` is the caller's buffer again
	flof_tmp_free(ctx, s2);
	flof_tmp_free(ctx, s1);
	flof_tmp_free(ctx, tmp);
	return rc;
}

extern "C" int flof_gaussian_blur4d(flof_ctx *ctx, float *a, flof_dim4 d, int elem, float sigma, int iter)
{
	return flof_gaussian_blur4d_impl(ctx, a, d, elem, sigma, iter);
}

// ------------------------------------------------------------------ 81-tap extrapolation ---
// ref knCvExpolBlur4d :613-626: where marker == 0 (interior, bnd 1) tmp = (sum of the 3^4
// neighbourhood in vt,zk,yj,xi order) * (1./81.); elsewhere tmp keeps the copy of a.
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_cv_expol_blur4d(const float4 *__restrict__ a, float4 *__restrict__ tmp, const float *__restrict__ mark,
                      flof_dim4 d)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t)) return;
	const int64_t c = flof_idx(d, i, j, k, t);
	if (!flof_in_bounds(d, i, j, k, t, 1) || __ldg(mark + c) != 0.f) {
		tmp[c] = __ldg(a + c);  // tmp.copyFrom(dst) fused in
		return;
	}
	float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
	for (int vt = t - 1; vt <= t + 1; ++vt)
		for (int zk = k - 1; zk <= k + 1; ++zk)
			for (int yj = j - 1; yj <= j + 1; ++yj) {
				const float4 *row = a + flof_idx(d, i - 1, yj, zk, vt);
#pragma unroll
				for (int xi = 0; xi < 3; ++xi) {
					const float4 q = __ldg(row + xi);
					val.x += q.x; val.y += q.y; val.z += q.z; val.w += q.w;
				}
			}
	const double f = 1. / 81.0;  // Vec4 * double, rounded per component
	tmp[c] = make_float4((float)(val.x * f), (float)(val.y * f), (float)(val.z * f), (float)(val.w * f));
}

// Sharded level: the +-1 ghost slices a sweep needs are the neighbours' outputs of the previous sweep.  Instead of
// "exchange, then sweep" (the exchange -- 33.5 MB per direction over NVLink at 128^4 -- sat between two sweeps: 0.14 ms
// against 0.42 ms of compute at 8 GPUs), the work list is split once per pass into the items of the slab's first and last
// slice and all the others.  Per sweep the BOUNDARY items run first on a high-priority side stream, followed there by the
// push / pull of the two slices they just produced; the INTERIOR items (which read no ghost slice) run on the main stream
// at the same time.  Dependencies (events): boundary(s) needs interior(s-1) -- its slice ta+1 / tb-2 -- and, by stream
// order, exchange(s-1); interior(s) needs boundary(s-1).  Same kernels, same arithmetic: results are bit-identical.
static int expol_sweeps_overlapped(flof_ctx *ctx, float *a, float *tmp, const float *marker, flof_dim4 d, int tz, int shfl,
                                   uint2 *items, unsigned int *count, int sweeps, size_t bytes, size_t slice_bytes)
{
	if (flof_side_stream_ensure(ctx)) return flof_fail(ctx, FLOF_ERR_CUDA, "side stream for the overlapped sweeps could not be created");
	cudaStream_t smain = ctx->stream, shi = ctx->stream_hi;
	cudaEvent_t evStart = ctx->ev_ov[0], evB = ctx->ev_ov[1], evI = ctx->ev_ov[2];
	int nB = 0, nI = 0;
	FLOF_RET(flof_expol_zn_build(ctx, marker, d, tz, items, count, &nB, 1));
	FLOF_RET(flof_expol_zn_build(ctx, marker, d, tz, items + nB, count, &nI, 2));
	float *cur = a, *oth = tmp;
	FLOF_RET(flof_memcpy_d2d(ctx, oth, cur, bytes));
	FLOF_RET(flof_halo_exchange(ctx, cur, d.nt, slice_bytes, 1));  // ghost slices of the start field
	FLOF_CK(cudaEventRecord(evStart, smain));
	FLOF_CK(cudaStreamWaitEvent(shi, evStart, 0));
	int rc = FLOF_OK;
	for (int sIt = 0; sIt < sweeps && rc == FLOF_OK; ++sIt) {
		// interior items on the main stream (after the boundary items of the previous sweep)
		if (sIt > 0) FLOF_CK(cudaStreamWaitEvent(smain, evB, 0));
		rc = flof_launch_expol_zn(ctx, cur, oth, items + nB, nI, d, tz, shfl);
		// boundary items + exchange on the side stream (after the interior items of the previous sweep)
		if (sIt > 0) FLOF_CK(cudaStreamWaitEvent(shi, evI, 0));
		FLOF_CK(cudaEventRecord(evI, smain));
		{
			struct OnSide {  // the launch helpers use ctx->stream: point it at the side stream for this scope only
				flof_ctx *c;
				cudaStream_t back;
				OnSide(flof_ctx *cc, cudaStream_t s) : c(cc), back(cc->stream) { c->stream = s; }
				~OnSide() { c->stream = back; }
			} side(ctx, shi);
			if (rc == FLOF_OK) rc = flof_launch_expol_zn(ctx, cur, oth, items, nB, d, tz, shfl);
			cudaEventRecord(evB, shi);
			if (rc == FLOF_OK && sIt + 1 < sweeps) rc = flof_halo_exchange(ctx, oth, d.nt, slice_bytes, 1);
		}
		float *sw = cur; cur = oth; oth = sw;
	}
	// join: everything on the side stream, then the last interior launch is already on the main stream
	cudaEventRecord(evB, shi);
	cudaStreamWaitEvent(smain, evB, 0);
	if (rc == FLOF_OK && cur != a) rc = flof_memcpy_d2d(ctx, a, cur, bytes);
	return rc;
}

extern "C" int flof_cv_expol_blur4d(flof_ctx *ctx, float *a, const float *marker, flof_dim4 d, int sweeps)
{
	if (sweeps <= 0) return FLOF_OK;
	const size_t bytes = sizeof(float) * 4 * (size_t)flof_cells(d);
	const size_t slice_bytes = sizeof(float) * 4 * (size_t)d.nx * d.ny * d.nz;
	// Work-list kernels, all bit-identical (tests/test_gpu_fullsize.py).  1: items of 4 y-cells; 3 / 4: items of 4y x 2z /
	// 4y x 4z cells (27 / 20.25 row loads per output instead of 40.5; measured 3.08 / 2.94 ms per sweep at 128^4 against
	// 4.36 ms for mode 1, 0.154 / 0.159 / 0.203 ms at 64^4); 0: component planes (4.1 ms at 128^4, plus two layout passes);
	// 2: dense kernel (no work list).  -1 (default): 4 on large grids, 3 below.
	int mode = ctx->opt.expol_mode;
	if (mode < 0) mode = flof_cells(d) >= ((int64_t)1 << 25) ? 4 : 3;
	const int64_t cap4 = mode == 0 ? flof_expol_planes_capacity(ctx, d) : 0;
	const int tz = (mode == 4 || mode == 6) ? 4 : 2;  // 5 / 6: like 3 / 4, the x-neighbour columns come from lane shuffles
	const int shfl = mode >= 5 ? 1 : 0;
	const int64_t capz = mode >= 3 ? flof_expol_zn_capacity(ctx, d, tz) : 0;  // 3 / 4: Vec4 work list with 4y x 2z / 4y x 4z items
	const int64_t cap1 = ((mode <= 1 || mode >= 3) && cap4 == 0 && capz == 0) ? flof_expol_item_capacity(ctx, d) : 0;
	void *tmp = NULL, *tmp2 = NULL, *items = NULL, *count = NULL;
	int rc = flof_tmp_alloc(ctx, &tmp, bytes, false);
	int n = 0;
	if (rc == FLOF_OK && (cap4 > 0 || cap1 > 0 || capz > 0)) {
		rc = flof_tmp_alloc(ctx, &items, (cap4 > 0 ? sizeof(uint2) * (size_t)cap4 : (capz > 0 ? sizeof(uint2) * (size_t)capz : sizeof(uint32_t) * (size_t)cap1)), false);
		if (rc == FLOF_OK) rc = flof_tmp_alloc(ctx, &count, sizeof(unsigned int), false);
	}
	if (rc == FLOF_OK && cap4 > 0) {
		// component-plane path: a -> planes (cur), planes copied once (oth), sweeps ping-pong, planes -> a
		rc = flof_tmp_alloc(ctx, &tmp2, bytes, false);
		float *cur = (float *)tmp, *oth = (float *)tmp2;
		if (rc == FLOF_OK) rc = flof_expol_planes_build(ctx, marker, d, (uint2 *)items, (unsigned int *)count, &n);
		if (rc == FLOF_OK) rc = flof_expol_to_planes(ctx, a, cur, d);
		if (rc == FLOF_OK) rc = flof_memcpy_d2d(ctx, oth, cur, bytes);
		for (int sIt = 0; sIt < sweeps && rc == FLOF_OK; ++sIt) {
			rc = flof_halo_exchange(ctx, cur, d.nt, slice_bytes, 1);  // sharded level: +-1 ghost slice per sweep
			if (rc == FLOF_OK) rc = flof_launch_expol_planes(ctx, cur, oth, (const uint2 *)items, n, d);
			float *sw = cur; cur = oth; oth = sw;
		}
		if (rc == FLOF_OK) rc = flof_expol_from_planes(ctx, a, cur, d);
	} else if (rc == FLOF_OK && capz > 0 && ctx->opt.sweep_overlap && ctx->nranks > 1 && flof_sharded(ctx, d.nt) &&
	           ctx->sh.tb - ctx->sh.ta >= 4) {
		rc = expol_sweeps_overlapped(ctx, a, (float *)tmp, marker, d, tz, shfl, (uint2 *)items, (unsigned int *)count, sweeps,
		                             bytes, slice_bytes);
	} else if (rc == FLOF_OK) {
		float *cur = a, *oth = (float *)tmp;
		if (cap1 > 0 || capz > 0) {
			// Vec4 work list: the cells that never change are copied into the second buffer once
			if (capz > 0)
				rc = flof_expol_zn_build(ctx, marker, d, tz, (uint2 *)items, (unsigned int *)count, &n);
			else
				rc = flof_expol_build(ctx, marker, d, (uint32_t *)items, (unsigned int *)count, &n);
			if (rc == FLOF_OK) rc = flof_memcpy_d2d(ctx, oth, cur, bytes);
		}
		for (int sIt = 0; sIt < sweeps && rc == FLOF_OK; ++sIt) {
			rc = flof_halo_exchange(ctx, cur, d.nt, slice_bytes, 1);
			if (rc != FLOF_OK) break;
			if (capz > 0)
				rc = flof_launch_expol_zn(ctx, cur, oth, (const uint2 *)items, n, d, tz, shfl);
			else if (cap1 > 0)
				rc = flof_launch_expol_items(ctx, cur, oth, (const uint32_t *)items, n, d);
			else
				rc = flof_launch_expol_tiled(ctx, cur, oth, marker, d);
			float *sw = cur; cur = oth; oth = sw;
		}
		if (rc == FLOF_OK && cur != a) rc = flof_memcpy_d2d(ctx, a, cur, bytes);
	}
	if (count) flof_tmp_free(ctx, count);
	if (items) flof_tmp_free(ctx, items);
	if (tmp2) flof_tmp_free(ctx, tmp2);
	if (tmp) flof_tmp_free(ctx, tmp);
	return rc;
}

// ------------------------------------------------------------------ border reset -----------
// ref :544-551: everything outside isInBounds(resetBnd) -> 0
__global__ void k_reset_border_vec4(float4 *__restrict__ vel, flof_kd d, int resetBnd)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t)) return;
	if (flof_in_bounds(d, i, j, k, t, resetBnd)) return;
	vel[flof_idx(d, i, j, k, t)] = make_float4(0.f, 0.f, 0.f, 0.f);
}
int flof_reset_border_vec4(flof_ctx *ctx, float *vel, flof_dim4 d, int resetBnd)
{
	dim3 g;
	const flof_kd kd = flof_kdim(ctx, d, &g);
	FLOF_LAUNCH(k_reset_border_vec4, g, FLOF_BLOCK, 0, (float4 *)vel, kd, resetBnd);
	return FLOF_OK;
}
