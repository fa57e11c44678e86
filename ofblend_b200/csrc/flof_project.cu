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
	// normal from central differences of interpolated phiOrg at +-0.5 (ref :651-673)
	const float h = 0.5f;
	float n0 = flof_interpol4d<float>(phiOrg, d, px + h, py, pz, pt) - flof_interpol4d<float>(phiOrg, d, px - h, py, pz, pt);
	float n1 = flof_interpol4d<float>(phiOrg, d, px, py + h, pz, pt) - flof_interpol4d<float>(phiOrg, d, px, py - h, pz, pt);
	float n2 = flof_interpol4d<float>(phiOrg, d, px, py, pz + h, pt) - flof_interpol4d<float>(phiOrg, d, px, py, pz - h, pt);
	float n3 = flof_interpol4d<float>(phiOrg, d, px, py, pz, pt + h) - flof_interpol4d<float>(phiOrg, d, px, py, pz, pt - h);
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
	float vn = flof_interpol4d<float>(phiOrg, d, px, py, pz, pt);
	float dd = vn - targetVal;
	for (int s = 0; s < maxIter; ++s) {
		const float qx = px + (n0 * dd) * dt, qy = py + (n1 * dd) * dt;
		const float qz = pz + (n2 * dd) * dt, qt = pt + (n3 * dd) * dt;
		// isInBounds(Vec4) truncates toward zero first (toVec4i), ref grid4d.h:72-75
		if (!flof_in_bounds(d, (int)qx, (int)qy, (int)qz, (int)qt, 0)) {
			dd = 9999.f;
			break;
		}
		vn = flof_interpol4d<float>(phiOrg, d, qx, qy, qz, qt);
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
