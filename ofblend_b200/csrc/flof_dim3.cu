// flof_dim3.cu -- the 3D instantiations of the optical-flow templates.
// ref (optflow4d.cpp): opticalFlowMultiscale3d :1175-1188 -> opticalFlowMultiscaleTemplate<FlagGrid, Grid<Real>, Grid<Vec3>>
//      :936-1173; corrVelsOf3d :803-812 -> corrVelsOfTempl :737-802 with knProjectCells :713-734, knCvExpolBlur3d :627-638;
//      advectCent3d / advectSemiLagrangeCfl :820-872; calcLsDiff3d :928-933; Grid<T> pieces: interpol util/interpol.h:57-116,
//      knInterpolateGridTempl grid.h:830-838, knSetBoundary / knSetBoundaryNeumann grid.cpp:462-506.
//
// How the 3D path is built here:
//  * Velocities cross the C ABI as Vec3 AoS (3 floats per cell, the reference's Grid<Vec3>) and are padded to float4 with
//    w = 0 on the device, so that one 128-bit access moves a cell and the Vec4 element-wise / min-max kernels serve both.
//  * The Jacobi-PCG solve re-uses the 4D kernels unchanged (including the sequential-order dot products): the 3D inputs are
//    laid into the middle slice of a THREE-slice 4D problem whose outer slices repeat it.  Then grad.t = (i1 - i1) * dxf is
//    exactly 0, the t-neighbours of the middle slice are identity rows that stay 0, every 4th component stays +-0, and all
//    extra terms of the matrix-vector products and of the fp64 dot products are +-0 added to a non-zero running sum: the
//    iterates of the middle slice are the reference's DIM = 3 iterates bit for bit.  Only the diagonal constant differs
//    ((2*DIM) * wSmooth + wEnergy, ref :478-480) and is passed explicitly.
//  * Everything else is a one-thread-per-cell 3D kernel below (3D grids are small: performance is not on the north-star
//    path, parity is).  The multi-scale driver is the same function as in 4D (flof_multiscale.cu): a level with nt == 1
//    dispatches every operator to its 3D form.
//  * 2D grids (nz == 1, the reference's DIM = 2 instantiation used by scenes/ofblend2dTest.py) take the same route: every
//    kernel below follows the reference's `is3D()` switches, the solve embeds the plane in z and in t.
#include <math.h>

#include "flof_common.cuh"

int flof_min_max_device(flof_ctx *ctx, const float *a, int64_t cells, int elem);
int flof_gaussian_blur4d_impl(flof_ctx *ctx, float *a, flof_dim4 d, int elem, float sigma, int iter);
int flof_of_cg_dim(flof_ctx *ctx, float *x, const float *grad, const float *rhs, flof_dim4 d, float wSmooth, float wEnergy,
                   int dim, float accuracy, int maxIter, int *iters, float *relResidual);
int flof_multiscale_run3d(flof_ctx *ctx, float *vel4, const float *i0, const float *i1, flof_dim4 d,
                          const flof_multiscale_params *p, flof_multiscale_trace *tr, float *err_out);

namespace {

__device__ __forceinline__ bool in3(const flof_dim4 &d, int i, int j, int k, int b)
{  // ref GridBase::isInBounds(Vec3i, bnd) grid.h:630-640: z is bounded on a 3D grid, must be 0 on a 2D one (nz == 1)
	if (!(i >= b && j >= b && i < d.nx - b && j < d.ny - b)) return false;
	return d.nz > 1 ? (k >= b && k < d.nz - b) : (k == 0);
}
__device__ __forceinline__ flof_dim3 d3of(const flof_dim4 &d)
{
	flof_dim3 r = { d.nx, d.ny, d.nz };
	return r;
}

// ---- Vec3 AoS <-> padded float4 -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FLOF_BLOCK) k3_pad(float4 *__restrict__ dst, const float *__restrict__ src, int64_t cells)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += stride)
		dst[c] = make_float4(__ldg(src + 3 * c), __ldg(src + 3 * c + 1), __ldg(src + 3 * c + 2), 0.f);
}
__global__ void __launch_bounds__(FLOF_BLOCK) k3_unpad(float *__restrict__ dst, const float4 *__restrict__ src, int64_t cells)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += stride) {
		const float4 v = __ldg(src + c);
		dst[3 * c] = v.x; dst[3 * c + 1] = v.y; dst[3 * c + 2] = v.z;
	}
}

// ---- resampling: knInterpolateGridTempl grid.h:830-838, pos = Vec3(i,j,k) * sourceFactor + offset ---------------------
template <class T>
__global__ void __launch_bounds__(FLOF_BLOCK)
    k3_interpol_grid(T *__restrict__ dst, flof_kd td, const T *__restrict__ src, flof_dim3 sd, float3 fac, float3 off)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(td, i, j, k, t)) return;
	const float px = (float)i * fac.x + off.x, py = (float)j * fac.y + off.y;
	const float pz = sd.nz > 1 ? (float)k * fac.z + off.z : 0.f;  // if (!source.is3D()) pos[2] = 0, grid.h:836
	dst[flof_idx(td, i, j, k, 0)] = flof_interpol3d<T>(src, sd, px, py, pz);
}

// ---- advection: semiLagrangeCent3d :820-825, KERNEL(bnd = 1) into a fresh (zero) grid, then swap ----------------------
template <class T> __device__ __forceinline__ T zero_of();
template <> __device__ __forceinline__ float zero_of<float>() { return 0.f; }
template <> __device__ __forceinline__ float4 zero_of<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }
template <class T>
__global__ void __launch_bounds__(FLOF_BLOCK)
    k3_semi_lagrange(const float4 *__restrict__ vel, const T *__restrict__ src, T *__restrict__ dst, flof_kd d, float dt)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t)) return;
	const int64_t c = flof_idx(d, i, j, k, 0);
	if (!in3(d, i, j, k, 1)) {
		dst[c] = zero_of<T>();
		return;
	}
	const float4 v = __ldg(vel + c);
	const float px = ((float)i + 0.5f) - v.x * dt, py = ((float)j + 0.5f) - v.y * dt, pz = ((float)k + 0.5f) - v.z * dt;
	dst[c] = flof_interpol3d<T>(src, d3of(d), px, py, pz);
}

// ---- boundaries: knSetBoundaryNeumann grid.cpp:475-506, knSetBoundary :462-468 ---------------------------------------
template <class T>
__global__ void __launch_bounds__(FLOF_BLOCK) k3_set_bound_neumann(T *grid, flof_kd d, int w)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t)) return;
	bool set = false;
	int si = i, sj = j, sk = k;
	if (i <= w) { si = w + 1; set = true; }
	if (i >= d.nx - 1 - w) { si = d.nx - 1 - w - 1; set = true; }
	if (j <= w) { sj = w + 1; set = true; }
	if (j >= d.ny - 1 - w) { sj = d.ny - 1 - w - 1; set = true; }
	if (d.nz > 1) {  // grid.is3D()
		if (k <= w) { sk = w + 1; set = true; }
		if (k >= d.nz - 1 - w) { sk = d.nz - 1 - w - 1; set = true; }
	}
	// the source cell has no clamped coordinate left, so it is never written by this launch
	if (set) grid[flof_idx(d, i, j, k, 0)] = grid[flof_idx(d, si, sj, sk, 0)];
}
__global__ void __launch_bounds__(FLOF_BLOCK) k3_set_bound_zero(float4 *grid, flof_kd d, int w)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t)) return;
	const bool bnd = i <= w || i >= d.nx - 1 - w || j <= w || j >= d.ny - 1 - w || (d.nz > 1 && (k <= w || k >= d.nz - 1 - w));
	if (bnd) grid[flof_idx(d, i, j, k, 0)] = make_float4(0.f, 0.f, 0.f, 0.f);
}
// ref :544-551: everything outside isInBounds(resetBnd) -> 0
__global__ void __launch_bounds__(FLOF_BLOCK) k3_reset_border(float4 *vel, flof_kd d, int resetBnd)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t)) return;
	if (in3(d, i, j, k, resetBnd)) return;
	vel[flof_idx(d, i, j, k, 0)] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---- projection: getNormalInterpTempl :651-673, projectCell :675-710, knProjectCells :713-734 (Vec3 instantiation) ----
__global__ void __launch_bounds__(FLOF_BLOCK)
    k3_project_cells(float4 *__restrict__ dst, const float4 *__restrict__ vel, const float *__restrict__ phiOrg,
                     const float *__restrict__ phiTarget, float *__restrict__ marker, flof_kd d, float threshPhi,
                     int maxIter, float dt)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t)) return;
	if (!in3(d, i, j, k, 2)) return;  // KERNEL(fourd, bnd = 2) on a 3D grid: no bound on t
	const flof_dim3 s = d3of(d);
	const int64_t c = flof_idx(d, i, j, k, 0);
	const float4 v = __ldg(vel + c);
	const float px = ((float)i + 0.5f) - v.x * dt, py = ((float)j + 0.5f) - v.y * dt, pz = ((float)k + 0.5f) - v.z * dt;
	const float h = 0.5f;
	float n0 = flof_interpol3d<float>(phiOrg, s, px + h, py, pz) - flof_interpol3d<float>(phiOrg, s, px - h, py, pz);
	float n1 = flof_interpol3d<float>(phiOrg, s, px, py + h, pz) - flof_interpol3d<float>(phiOrg, s, px, py - h, pz);
	float n2 = 0.f;  // `VEC n;` is zero-initialised (vectorbase.h:92); n[2] is only set if (phi.is3D()), ref :662-665
	if (d.nz > 1) n2 = flof_interpol3d<float>(phiOrg, s, px, py, pz + h) - flof_interpol3d<float>(phiOrg, s, px, py, pz - h);
	{  // normalize(Vector3D), ref util/vectorbase.h:401-415
		const float l = n0 * n0 + n1 * n1 + n2 * n2;
		const double eps2 = (double)(FLOF_VECTOR_EPSILON * FLOF_VECTOR_EPSILON);
		if (fabs((double)l - 1.) < eps2) {
			// normalized "enough"
		} else if ((double)l > eps2) {
			const float nrm = sqrtf(l);
			const float fac = (float)(1. / (double)nrm);
			n0 *= fac; n1 *= fac; n2 *= fac;
		} else {
			n0 = n1 = n2 = 0.f;
		}
	}
	float step = 0.25f;
	if (maxIter < 50) step = 0.5f;
	int lastDir = 0;
	const float targetVal = __ldg(phiTarget + c);
	float vn = flof_interpol3d<float>(phiOrg, s, px, py, pz);
	float dd = vn - targetVal;
	// flat regions (n == 0): the probe position never moves, see k_project_cells in flof_project.cu
	const bool flat = n0 == 0.f && n1 == 0.f && n2 == 0.f && fabsf(dd) <= 3.0e38f;
	for (int it = 0; it < maxIter; ++it) {
		if (!flat || it == 0) {
			const float qx = px + (n0 * dd) * dt, qy = py + (n1 * dd) * dt, qz = pz + (n2 * dd) * dt;
			if (!in3(d, (int)qx, (int)qy, (int)qz, 0)) {  // isInBounds(Vec3) truncates toward zero first, grid.h:72-75
				dd = 9999.f;
				break;
			}
			vn = flof_interpol3d<float>(phiOrg, s, qx, qy, qz);
		}
		if (vn < targetVal) {
			if (lastDir < 0) step *= 0.5f;
			dd += step;
			lastDir = 1;
		} else {
			if (lastDir > 0) step *= 0.5f;
			dd -= step;
			lastDir = -1;
		}
		if ((double)step < 1e-01) break;
	}
	if (fabsf(dd) > threshPhi) {
		dst[c] = make_float4(0.f, 0.f, 0.f, 0.f);
		return;
	}
	dst[c] = make_float4(n0 * dd, n1 * dd, n2 * dd, 0.f);
	marker[c] = 1.f;
}

// ---- knCvExpolBlur3d :627-638: 27 taps in (zk, yj, xi) order, * (1./27.0); knCvExpolBlur2d :639-649: 9 taps, * (1./9.0);
// other cells keep the copy -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FLOF_BLOCK)
    k3_cv_expol_blur(const float4 *__restrict__ a, float4 *__restrict__ tmp, const float *__restrict__ mark, flof_kd d)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t)) return;
	const int64_t c = flof_idx(d, i, j, k, 0);
	if (!in3(d, i, j, k, 1) || __ldg(mark + c) != 0.f) {
		tmp[c] = __ldg(a + c);  // tmp.copyFrom(dst) fused in
		return;
	}
	float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
	const int z0 = d.nz > 1 ? k - 1 : 0, z1 = d.nz > 1 ? k + 1 : 0;
	for (int zk = z0; zk <= z1; ++zk)
		for (int yj = j - 1; yj <= j + 1; ++yj) {
			const float4 *row = a + flof_idx(d, i - 1, yj, zk, 0);
#pragma unroll
			for (int xi = 0; xi < 3; ++xi) {
				const float4 q = __ldg(row + xi);
				val.x += q.x; val.y += q.y; val.z += q.z;
			}
		}
	const double f = d.nz > 1 ? 1. / 27.0 : 1. / 9.0;
	tmp[c] = make_float4((float)(val.x * f), (float)(val.y * f), (float)(val.z * f), 0.f);
}

// ---- optical flow: the 3D data embedded as the middle slice of a 3-slice 4D problem (see the header) -----------------
// grad.t = 0 and rhs.t = -0 * tderiv are produced by k_of_assemble itself from the repeated slices.
__global__ void __launch_bounds__(FLOF_BLOCK)
    k3_copy_back(float4 *__restrict__ vel, const float4 *__restrict__ x, const float4 *__restrict__ rhs,
                 float *__restrict__ rhsT, int64_t cells, float mDx)
{  // ref :520-529 on the middle slice (x, rhs point at it)
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += stride) {
		const float4 v = __ldg(x + c);
		vel[c] = make_float4(v.x / mDx, v.y / mDx, v.z / mDx, 0.f);
		if (rhsT) rhsT[c] = __ldg(rhs + c).x;
	}
}

struct Tmp {  // RAII for pool temporaries
	flof_ctx *ctx;
	void *p;
	explicit Tmp(flof_ctx *c) : ctx(c), p(NULL) {}
	~Tmp() { flof_tmp_free(ctx, p); }
	int alloc(size_t bytes, bool zero) { return flof_tmp_alloc(ctx, &p, bytes, zero); }
	float *f() { return (float *)p; }
};

flof_kd kd3(flof_dim4 d, dim3 *g)
{
	flof_kd k;
	k.nx = d.nx; k.ny = d.ny; k.nz = d.nz; k.nt = 1;
	k.t0 = 0;
	*g = dim3((unsigned)(((int64_t)d.nx * d.ny + FLOF_BLOCK - 1) / FLOF_BLOCK), (unsigned)d.nz, 1);
	return k;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// operators on padded (float4) device grids; d.nt == 1.  Used by the multi-scale driver's 3D table and by the C ABI.
int flof3_interpol_grid(flof_ctx *ctx, float *dst, flof_dim4 td, const float *src, flof_dim4 sd, int elem)
{
	// calcGridSizeFactor grid.h:490-493 (float / int), offset = sourceFactor * 0.5
	const float3 fac = make_float3((float)sd.nx / (float)td.nx, (float)sd.ny / (float)td.ny, (float)sd.nz / (float)td.nz);
	const float3 off = make_float3(fac.x * 0.5f, fac.y * 0.5f, fac.z * 0.5f);
	dim3 g;
	const flof_kd kd = kd3(td, &g);
	const flof_dim3 s3 = { sd.nx, sd.ny, sd.nz };
	if (elem == 4)
		FLOF_LAUNCH(k3_interpol_grid<float4>, g, FLOF_BLOCK, 0, (float4 *)dst, kd, (const float4 *)src, s3, fac, off);
	else
		FLOF_LAUNCH(k3_interpol_grid<float>, g, FLOF_BLOCK, 0, dst, kd, src, s3, fac, off);
	return FLOF_OK;
}

// advectCflHelperCentered :841-853 (velFactor applied to a copy of vel first, :866-870)
int flof3_advect_cfl(flof_ctx *ctx, float cfl, const float *vel, float *grid, flof_dim4 d, int elem, float velFactor)
{
	const int64_t cells = flof_cells(d);
	const size_t gb = sizeof(float) * (size_t)elem * (size_t)cells, vb = sizeof(float) * 4 * (size_t)cells;
	Tmp velTmp(ctx), fwd(ctx);
	const float *v = vel;
	if (velFactor != 1.f) {
		FLOF_RET(velTmp.alloc(vb, false));
		FLOF_RET(flof_memcpy_d2d(ctx, velTmp.p, vel, vb));
		const float f4[4] = { velFactor, velFactor, velFactor, velFactor };
		FLOF_RET(flof_grid_mult_const(ctx, velTmp.f(), cells, 4, f4));
		v = velTmp.f();
	}
	FLOF_RET(flof_min_max_device(ctx, v, cells, 4));
	float *h = (float *)ctx->pinned;
	FLOF_CK(cudaMemcpyAsync(h, ctx->red->out_f, 2 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
	FLOF_CK(cudaStreamSynchronize(ctx->stream));
	const float orgDt = 1.0f;
	const float maxVel = sqrtf(h[1]) * orgDt;  // Grid<Vec3>::getMax = sqrt(max normSquare), grid.cpp:261
	const int steps = (int)(maxVel / cfl) + 1;
	const float dt = orgDt / (float)steps;
	FLOF_RET(fwd.alloc(gb, false));
	float *cur = grid, *oth = fwd.f();
	dim3 g;
	const flof_kd kd = kd3(d, &g);
	for (int s = 0; s < steps; ++s) {
		if (elem == 4)
			FLOF_LAUNCH(k3_semi_lagrange<float4>, g, FLOF_BLOCK, 0, (const float4 *)v, (const float4 *)cur, (float4 *)oth, kd, dt);
		else
			FLOF_LAUNCH(k3_semi_lagrange<float>, g, FLOF_BLOCK, 0, (const float4 *)v, cur, oth, kd, dt);
		float *sw = cur; cur = oth; oth = sw;  // orig.swap(fwd)
	}
	if (cur != grid) FLOF_RET(flof_memcpy_d2d(ctx, grid, cur, gb));
	return FLOF_OK;
}

int flof3_set_bound_neumann(flof_ctx *ctx, float *grid, flof_dim4 d, int elem, int w)
{
	FLOF_ARG(d.nx >= 2 * w + 3 && d.ny >= 2 * w + 3 && (d.nz == 1 || d.nz >= 2 * w + 3), "setBoundNeumann: grid too small for width %d", w);
	dim3 g;
	const flof_kd kd = kd3(d, &g);
	if (elem == 4)
		FLOF_LAUNCH(k3_set_bound_neumann<float4>, g, FLOF_BLOCK, 0, (float4 *)grid, kd, w);
	else
		FLOF_LAUNCH(k3_set_bound_neumann<float>, g, FLOF_BLOCK, 0, grid, kd, w);
	return FLOF_OK;
}

// opticalFlowDim<Grid<Real>, Grid<Vec3>, 3> :361-553; a 2D grid (nz == 1, DIM = 2) is embedded the same way in z AND t:
// the centre of a 3 x 3 block of copies, diagonal constant of DIM = 2
int flof3_optical_flow(flof_ctx *ctx, float *vel, const float *i0, const float *i1, float *rhsT, flof_dim4 d, float wSmooth,
                       float wEnergy, float postVelBlur, float cgAccuracy, float resetBndWidth, int vel_is_zero, int *cgIters,
                       float *cgRes)
{
	FLOF_ARG(d.nx >= 3 && d.ny >= 3 && (d.nz >= 3 || d.nz == 1) && d.nt == 1, "opticalFlow (3D): grid too small");
	const int64_t cells = flof_cells(d);
	const bool two = d.nz == 1;
	const flof_dim4 e = { d.nx, d.ny, two ? 3 : d.nz, 3 };  // the embedding: the outer slices repeat the middle one
	const int copies = two ? 9 : 3, mid = two ? 4 : 1;
	const size_t rb = sizeof(float) * (size_t)cells, vb = rb * 4;
	Tmp i0e(ctx), i1e(ctx), ve(ctx), grad(ctx), rhs(ctx), x(ctx);
	FLOF_RET(i0e.alloc(copies * rb, false));
	FLOF_RET(i1e.alloc(copies * rb, false));
	for (int s = 0; s < copies; ++s) {
		FLOF_RET(flof_memcpy_d2d(ctx, (char *)i0e.p + s * rb, i0, rb));
		FLOF_RET(flof_memcpy_d2d(ctx, (char *)i1e.p + s * rb, i1, rb));
	}
	if (!vel_is_zero) {
		FLOF_RET(ve.alloc(copies * vb, false));
		for (int s = 0; s < copies; ++s) FLOF_RET(flof_memcpy_d2d(ctx, (char *)ve.p + s * vb, vel, vb));
	}
	FLOF_RET(grad.alloc(copies * vb, false));
	FLOF_RET(rhs.alloc(copies * vb, false));
	FLOF_RET(x.alloc(copies * vb, false));
	FLOF_RET(flof_of_assemble(ctx, grad.f(), rhs.f(), i0e.f(), i1e.f(), vel_is_zero ? NULL : ve.f(), e, wSmooth, wEnergy));
	int it = 0;
	float rr = 1e10f;
	FLOF_RET(flof_of_cg_dim(ctx, x.f(), grad.f(), rhs.f(), e, wSmooth, wEnergy, two ? 2 : 3, cgAccuracy, 1000, &it, &rr));
	if (rr != rr) FLOF_RET(flof_memset0(ctx, x.p, copies * vb));  // ref :509-514
	const float mDx = (float)(1. / d.nx);
	FLOF_LAUNCH(k3_copy_back, flof_flat_blocks(ctx, cells, 8), FLOF_BLOCK, 0, (float4 *)vel, (const float4 *)x.p + mid * cells,
	            (const float4 *)rhs.p + mid * cells, rhsT, cells, mDx);
	if (postVelBlur > 0.f) FLOF_RET(flof_gaussian_blur4d_impl(ctx, vel, d, 4, (float)(0.5 * postVelBlur), 1));
	if (resetBndWidth > 0.f) {
		const int resetBnd = (int)(resetBndWidth * d.nx) + 1;
		dim3 g;
		const flof_kd kd = kd3(d, &g);
		FLOF_LAUNCH(k3_reset_border, g, FLOF_BLOCK, 0, (float4 *)vel, kd, resetBnd);
	}
	if (cgIters) *cgIters = it;
	if (cgRes) *cgRes = rr;
	return FLOF_OK;
}

// corrVelsOfTempl<Grid<Real>, Grid<Vec3>> :737-802
int flof3_corr_vels(flof_ctx *ctx, float *dst, float *vel, const float *phiOrg, const float *phiTarget, flof_dim4 d,
                    float threshPhi, float postVelBlur, float resetBndWidth, int maxIter)
{
	FLOF_ARG(d.nx >= 5 && d.ny >= 5 && (d.nz >= 5 || d.nz == 1) && d.nt == 1, "corrVelsOf3d: grid too small");
	const int64_t cells = flof_cells(d);
	const size_t vb = sizeof(float) * 4 * (size_t)cells;
	const float blurThreshold = 0.98f;
	FLOF_RET(flof_min_max_device(ctx, vel, cells, 4));
	float *h = (float *)ctx->pinned;
	FLOF_CK(cudaMemcpyAsync(h, ctx->red->out_f, 2 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
	FLOF_CK(cudaStreamSynchronize(ctx->stream));
	const int doCvExtrapol = (int)(sqrtf(h[1]) + 4);  // ref :748
	Tmp marker(ctx), tmp(ctx);
	FLOF_RET(marker.alloc(sizeof(float) * (size_t)cells, false));
	FLOF_RET(tmp.alloc(vb, false));
	dim3 g;
	const flof_kd kd = kd3(d, &g);
	int singleStep = postVelBlur < blurThreshold ? 1 : 0;
	while ((postVelBlur >= blurThreshold) || singleStep) {
		FLOF_RET(flof_memset0(ctx, dst, vb));
		FLOF_RET(flof_memset0(ctx, marker.p, sizeof(float) * (size_t)cells));
		FLOF_LAUNCH(k3_project_cells, g, FLOF_BLOCK, 0, (float4 *)dst, (const float4 *)vel, phiOrg, phiTarget, marker.f(), kd,
		            threshPhi, maxIter, 1.0f);
		float *cur = dst, *oth = tmp.f();
		for (int s = 0; s < doCvExtrapol; ++s) {
			FLOF_LAUNCH(k3_cv_expol_blur, g, FLOF_BLOCK, 0, (const float4 *)cur, (float4 *)oth, (const float *)marker.p, kd);
			float *sw = cur; cur = oth; oth = sw;
		}
		if (cur != dst) FLOF_RET(flof_memcpy_d2d(ctx, dst, cur, vb));
		if (postVelBlur > FLOF_VECTOR_EPSILON) FLOF_RET(flof_gaussian_blur4d_impl(ctx, dst, d, 4, (float)(0.5 * postVelBlur), 1));
		const int resetBnd = (int)(resetBndWidth * d.nx) + 1;
		if (resetBndWidth > 0.f) FLOF_LAUNCH(k3_set_bound_zero, g, FLOF_BLOCK, 0, (float4 *)dst, kd, resetBnd);
		const float m1[4] = { -1.f, -1.f, -1.f, -1.f };
		FLOF_RET(flof_grid_add_scaled(ctx, vel, dst, cells, 4, m1));  // vel.addScaled(dst, -1)
		postVelBlur *= 0.5f;
		if (singleStep > 0) singleStep--;
	}
	return FLOF_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// C ABI: Vec3 AoS velocities in device memory, padded to float4 temporaries here
static int pad3(flof_ctx *ctx, float *dst4, const float *src3, int64_t cells)
{
	FLOF_LAUNCH(k3_pad, flof_flat_blocks(ctx, cells, 8), FLOF_BLOCK, 0, (float4 *)dst4, src3, cells);
	return FLOF_OK;
}
static int unpad3(flof_ctx *ctx, float *dst3, const float *src4, int64_t cells)
{
	FLOF_LAUNCH(k3_unpad, flof_flat_blocks(ctx, cells, 8), FLOF_BLOCK, 0, dst3, (const float4 *)src4, cells);
	return FLOF_OK;
}
#define D4(d) flof_dim4 { (d).nx, (d).ny, (d).nz, 1 }

extern "C" int flof_optical_flow_multiscale3d(flof_ctx *ctx, float *vel3, const float *i0, const float *i1, flof_dim3 d,
                                              const flof_multiscale_params *p, flof_multiscale_trace *tr, float *err_out)
{
	FLOF_ARG(p != NULL, "opticalFlowMultiscale3d: params is NULL");
	FLOF_ARG(d.nx >= 3 && d.ny >= 3 && (d.nz >= 3 || d.nz == 1), "opticalFlowMultiscale3d: grid too small");
	const int64_t cells = flof_cells3(d);
	Tmp v4(ctx);
	FLOF_RET(v4.alloc(sizeof(float) * 4 * (size_t)cells, false));
	FLOF_RET(pad3(ctx, v4.f(), vel3, cells));
	FLOF_RET(flof_multiscale_run3d(ctx, v4.f(), i0, i1, D4(d), p, tr, err_out));
	return unpad3(ctx, vel3, v4.f(), cells);
}

extern "C" int flof_corr_vels_of3d(flof_ctx *ctx, float *dst3, float *vel3, const float *phiOrg, const float *phiTarget,
                                   flof_dim3 d, float threshPhi, float postVelBlur, float resetBndWidth, int maxIter)
{
	const int64_t cells = flof_cells3(d);
	Tmp d4(ctx), v4(ctx);
	FLOF_RET(d4.alloc(sizeof(float) * 4 * (size_t)cells, false));
	FLOF_RET(v4.alloc(sizeof(float) * 4 * (size_t)cells, false));
	FLOF_RET(pad3(ctx, v4.f(), vel3, cells));
	FLOF_RET(flof3_corr_vels(ctx, d4.f(), v4.f(), phiOrg, phiTarget, D4(d), threshPhi, postVelBlur, resetBndWidth, maxIter));
	FLOF_RET(unpad3(ctx, dst3, d4.f(), cells));
	return unpad3(ctx, vel3, v4.f(), cells);
}

// elem 1: Grid<Real>, elem 3: Grid<Vec3> payload (advectCent3d :836-846 = one step: pass a cfl above max|vel|)
extern "C" int flof_advect_semi_lagrange_cfl3d(flof_ctx *ctx, float cfl, const float *vel3, float *grid, int elem, flof_dim3 d,
                                               float velFactor)
{
	FLOF_ARG(elem == 1 || elem == 3, "advectSemiLagrangeCfl: Grid Type is not supported (only Real, Vec3)");
	FLOF_ARG(d.nx >= 3 && d.ny >= 3 && (d.nz >= 3 || d.nz == 1), "advectSemiLagrangeCfl: grid too small");
	FLOF_ARG(cfl > 0.f, "advectSemiLagrangeCfl: cfl must be positive");
	const int64_t cells = flof_cells3(d);
	Tmp v4(ctx), g4(ctx);
	FLOF_RET(v4.alloc(sizeof(float) * 4 * (size_t)cells, false));
	FLOF_RET(pad3(ctx, v4.f(), vel3, cells));
	if (elem == 1) return flof3_advect_cfl(ctx, cfl, v4.f(), grid, D4(d), 1, velFactor);
	FLOF_RET(g4.alloc(sizeof(float) * 4 * (size_t)cells, false));
	FLOF_RET(pad3(ctx, g4.f(), grid, cells));
	FLOF_RET(flof3_advect_cfl(ctx, cfl, v4.f(), g4.f(), D4(d), 4, velFactor));
	return unpad3(ctx, grid, g4.f(), cells);
}

extern "C" int flof_calc_ls_diff3d(flof_ctx *ctx, const float *i0, const float *i1, float *out, flof_dim3 d, float correction,
                                   int bnd, float *result)
{
	FLOF_ARG(result != NULL, "calcLsDiff3d: result is NULL");
	return flof_calc_ls_diff4d(ctx, i0, i1, out, D4(d), correction, bnd, result);  // nt == 1: the 3D rules (kernel.h:62-68)
}
