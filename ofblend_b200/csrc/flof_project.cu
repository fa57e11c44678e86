// flof_project.cu -- SDF projection of the deformation (final correction step of mode 1).
// ref: getNormalInterpTempl :651-673, projectCell :675-710, knProjectCells :713-734,
//      corrVelsOfTempl :737-802, corrVelsOf4d :2121-2130 (optflow4d.cpp).
//
// knProjectCells is a data-dependent search: per cell 9+ quadrilinear look-ups into phiOrg
// (8 for the normal, then up to 40 for the bisection-like walk along the normal).  It is
// latency/gather bound, not streaming; one thread per cell, phiOrg stays in L2 (4 B/cell).
#include <math.h>

#include "flof_common.cuh"

int flof_min_max_device(flof_ctx *ctx, const float *a, int64_t cells, int elem);
int flof_gaussian_blur4d_impl(flof_ctx *ctx, float *a, flof_dim4 d, int elem, float sigma, int iter);

// one axis of BUILD_INDEX_4D (util/vector4d.h:422-475): cell-centred position -> base index + the two weights, with the
// reference's clamp rules; identical arithmetic to flof_build_index4, split per axis
struct proj_axis {
	int i;
	float w0, w1;
};
__device__ __forceinline__ proj_axis proj_axis_of(float x, int n)
{
	proj_axis a;
	const float p = x - 0.5f;
	a.i = (int)p;
	a.w1 = p - (float)a.i;
	a.w0 = 1.0f - a.w1;
	if (p < 0.f) { a.i = 0; a.w0 = 1.f; a.w1 = 0.f; }
	if (a.i >= n - 1) { a.i = n - 2; a.w0 = 0.f; a.w1 = 1.f; }
	return a;
}
// interpol4d (util/vector4d.h:487-513) from four prepared axes: evaluation order ((y)x)z)t, fp32 mul/add; 32-bit cell
// offsets (the caller checks cells < 2^31)
__device__ __forceinline__ float proj_interpol(const float *__restrict__ data, const flof_dim4 &d, const proj_axis &X,
                                               const proj_axis &Y, const proj_axis &Z, const proj_axis &T)
{
	const int sY = d.nx, sZ = d.nx * d.ny, sT = sZ * d.nz;
	const float *p = data + (X.i + d.nx * (Y.i + d.ny * (Z.i + d.nz * T.i)));
	const float a0 = (__ldg(p) * Y.w0 + __ldg(p + sY) * Y.w1) * X.w0 + (__ldg(p + 1) * Y.w0 + __ldg(p + 1 + sY) * Y.w1) * X.w1;
	const float a1 = (__ldg(p + sZ) * Y.w0 + __ldg(p + sY + sZ) * Y.w1) * X.w0 +
	                 (__ldg(p + 1 + sZ) * Y.w0 + __ldg(p + 1 + sY + sZ) * Y.w1) * X.w1;
	const float *r = p + sT;
	const float b0 = (__ldg(r) * Y.w0 + __ldg(r + sY) * Y.w1) * X.w0 + (__ldg(r + 1) * Y.w0 + __ldg(r + 1 + sY) * Y.w1) * X.w1;
	const float b1 = (__ldg(r + sZ) * Y.w0 + __ldg(r + sY + sZ) * Y.w1) * X.w0 +
	                 (__ldg(r + 1 + sZ) * Y.w0 + __ldg(r + 1 + sY + sZ) * Y.w1) * X.w1;
	return (a0 * Z.w0 + a1 * Z.w1) * T.w0 + (b0 * Z.w0 + b1 * Z.w1) * T.w1;
}

__global__ void __launch_bounds__(FLOF_BLOCK)
    k_project_cells(float4 *__restrict__ dst, const float4 *__restrict__ vel, const float *__restrict__ phiOrg,
                    const float *__restrict__ phiTarget, float *__restrict__ marker, flof_kd d,
                    float threshPhi, int maxIter, float dt)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t)) return;
	if (!flof_in_bounds(d, i, j, k, t, 2)) return;  // KERNEL(fourd, bnd = 2)
	const int64_t c = flof_idx(d, i, j, k, t);
	const float4 v = __ldg(vel + c);
	const float px = ((float)i + 0.5f) - v.x * dt, py = ((float)j + 0.5f) - v.y * dt;
	const float pz = ((float)k + 0.5f) - v.z * dt, pt = ((float)t + 0.5f) - v.w * dt;
	// normal from central differences of interpolated phiOrg at +-0.5 (ref :651-673).  The nine probes (eight here, the
	// centre one below) differ in one coordinate at a time: index, clamp and weights are evaluated once per axis and
	// offset (12 instead of 36 evaluations), each probe then runs the reference's interpolation arithmetic unchanged.
	const float h = 0.5f;
	const proj_axis ax0 = proj_axis_of(px, d.nx), axm = proj_axis_of(px - h, d.nx), axp = proj_axis_of(px + h, d.nx);
	const proj_axis ay0 = proj_axis_of(py, d.ny), aym = proj_axis_of(py - h, d.ny), ayp = proj_axis_of(py + h, d.ny);
	const proj_axis az0 = proj_axis_of(pz, d.nz), azm = proj_axis_of(pz - h, d.nz), azp = proj_axis_of(pz + h, d.nz);
	const proj_axis at0 = proj_axis_of(pt, d.nt), atm = proj_axis_of(pt - h, d.nt), atp = proj_axis_of(pt + h, d.nt);
	float n0 = proj_interpol(phiOrg, d, axp, ay0, az0, at0) - proj_interpol(phiOrg, d, axm, ay0, az0, at0);
	float n1 = proj_interpol(phiOrg, d, ax0, ayp, az0, at0) - proj_interpol(phiOrg, d, ax0, aym, az0, at0);
	float n2 = proj_interpol(phiOrg, d, ax0, ay0, azp, at0) - proj_interpol(phiOrg, d, ax0, ay0, azm, at0);
	float n3 = proj_interpol(phiOrg, d, ax0, ay0, az0, atp) - proj_interpol(phiOrg, d, ax0, ay0, az0, atm);
	{  // normalize(), ref util/vector4d.h:331-345
		const float l = n0 * n0 + n1 * n1 + n2 * n2 + n3 * n3;
		const double eps2 = (double)(FLOF_VECTOR_EPSILON * FLOF_VECTOR_EPSILON);
		if (fabs((double)l - 1.) < eps2) {
			// normalized "enough"
		} else if ((double)l > eps2) {
			const float nrm = sqrtf(l);
			const float fac = (float)(1. / (double)nrm);
			n0 *= fac; n1 *= fac; n2 *= fac; n3 *= fac;
		} else {
			n0 = n1 = n2 = n3 = 0.f;
		}
	}
	// projectCell, ref :675-710
	float step = 0.25f;
	if (maxIter < 50) step = 0.5f;
	int lastDir = 0;
	const float targetVal = __ldg(phiTarget + c);
	float vn = proj_interpol(phiOrg, d, ax0, ay0, az0, at0);
	float dd = vn - targetVal;
	// Flat regions of phiOrg (the clamped far field: all eight normal probes equal, so n == 0 exactly) never leave
	// their position: p + (0 * dd) * dt == p for every finite dd, the probe returns the same value in every
	// iteration, the walk never changes direction and runs all maxIter steps.  There the look-up is done once and
	// the remaining iterations only repeat the reference's fp32 adds of `step` (same values, same order).  On the
	// synthetic pair that is 23 % (64^4) to 45 % (128^4) of the cells -- the ones that end up unmarked.
	const bool flat = n0 == 0.f && n1 == 0.f && n2 == 0.f && n3 == 0.f && fabsf(dd) <= 3.0e38f;
	for (int s = 0; s < maxIter; ++s) {
		if (!flat || s == 0) {
			const float qx = px + (n0 * dd) * dt, qy = py + (n1 * dd) * dt;
			const float qz = pz + (n2 * dd) * dt, qt = pt + (n3 * dd) * dt;
			// isInBounds(Vec4) truncates toward zero first (toVec4i), ref grid4d.h:72-75
			if (!flof_in_bounds(d, (int)qx, (int)qy, (int)qz, (int)qt, 0)) {
				dd = 9999.f;
				break;
			}
			vn = proj_interpol(phiOrg, d, proj_axis_of(qx, d.nx), proj_axis_of(qy, d.ny), proj_axis_of(qz, d.nz),
			                   proj_axis_of(qt, d.nt));
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
		if ((double)step < 1e-01) break;  // converged
	}
	if (fabsf(dd) > threshPhi) {
		dst[c] = make_float4(0.f, 0.f, 0.f, 0.f);
		return;
	}
	dst[c] = make_float4(n0 * dd, n1 * dd, n2 * dd, n3 * dd);
	marker[c] = 1.f;
}

extern "C" int flof_project_cells(flof_ctx *ctx, float *dst, const float *vel, const float *phiOrg,
                                  const float *phiTarget, float *marker, flof_dim4 d, float threshPhi,
                                  int maxIter)
{
	FLOF_ARG(d.nx >= 2 && d.ny >= 2 && d.nz >= 2 && d.nt >= 2, "corrVelsOf4d: grid too small");
	FLOF_ARG(flof_cells(d) < ((int64_t)1 << 31), "corrVelsOf4d: more than 2^31 cells");
	dim3 g;
	const flof_kd kd = flof_kdim(ctx, d, &g);  // sharded: phiOrg must be complete, everything else is per cell
	FLOF_LAUNCH(k_project_cells, g, FLOF_BLOCK, 0, (float4 *)dst, (const float4 *)vel, phiOrg, phiTarget, marker, kd,
	            threshPhi, maxIter, 1.0f);
	return FLOF_OK;
}

extern "C" int flof_corr_vels_of4d(flof_ctx *ctx, float *dst, float *vel, const float *phiOrg,
                                   const float *phiTarget, flof_dim4 d, float threshPhi, float postVelBlur,
                                   float resetBndWidth, int maxIter)
{
	const int64_t cells = flof_cells(d);
	const float blurThreshold = 0.98f;
	// ref :748 doCvExtrapol = vel.getMaxAbsValue() + 4 (float add, truncated)
	FLOF_RET(flof_min_max_device(ctx, vel, cells, 4));
	float *h = (float *)ctx->pinned;
	FLOF_CK(cudaMemcpyAsync(h, ctx->red->out_f, 2 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
	FLOF_CK(cudaStreamSynchronize(ctx->stream));
	const int doCvExtrapol = (int)(sqrtf(h[1]) + 4);

	void *marker = NULL;
	FLOF_RET(flof_tmp_alloc(ctx, &marker, sizeof(float) * (size_t)cells, false));
	int singleStep = 0;
	if (postVelBlur < blurThreshold) singleStep = 1;
	int rc = FLOF_OK;
	while (rc == FLOF_OK && ((postVelBlur >= blurThreshold) || singleStep)) {
		rc = flof_memset0(ctx, dst, sizeof(float) * 4 * (size_t)cells);
		if (rc == FLOF_OK) rc = flof_memset0(ctx, marker, sizeof(float) * (size_t)cells);
		if (rc == FLOF_OK) rc = flof_project_cells(ctx, dst, vel, phiOrg, phiTarget, (float *)marker, d, threshPhi, maxIter);
		if (rc == FLOF_OK && doCvExtrapol) rc = flof_cv_expol_blur4d(ctx, dst, (const float *)marker, d, doCvExtrapol);
		if (rc == FLOF_OK && postVelBlur > FLOF_VECTOR_EPSILON)
			rc = flof_gaussian_blur4d_impl(ctx, dst, d, 4, (float)(0.5 * postVelBlur), 1);
		const int resetBnd = (int)(resetBndWidth * d.nx) + 1;
		const float zero[4] = { 0.f, 0.f, 0.f, 0.f };
		if (rc == FLOF_OK && resetBndWidth > 0.f) rc = flof_grid4d_set_bound(ctx, dst, d, 4, zero, resetBnd);
		const float m1[4] = { -1.f, -1.f, -1.f, -1.f };
		if (rc == FLOF_OK) rc = flof_grid_add_scaled(ctx, vel, dst, cells, 4, m1);  // vel.addScaled(dst, -1)
		postVelBlur *= 0.5f;
		if (singleStep > 0) singleStep--;
	}
	flof_tmp_free(ctx, marker);
	return rc;
}
