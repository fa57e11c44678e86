/* oracle/flof_oracle3.c -- TEST INFRASTRUCTURE.  Plain-C restatement of the 3D / 2D instantiations of the reference's
 * optical-flow templates: opticalFlowMultiscale3d, corrVelsOf3d, advectSemiLagrangeCfl, calcLsDiff3d on
 * Grid<Real> / Grid<Vec3> (3 floats per cell, x fastest).  nz == 1 is a 2D grid: the reference's is3D() switches and its
 * DIM = 2 template instantiation apply.  Pinned bit for bit to the compiled reference by tests/test_oracle_vs_ref.py.
 *
 * Citations are relative to /root/reference/source/.  Float / double mixing follows the C++ usual arithmetic conversions
 * of the cited lines (see flof_oracle.c); everything here is serial -- 3D grids are small.
 */
#include "flof_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define VECTOR_EPSILON (1e-6f) /* util/vectorbase.h:53 */

typedef long long i64;
typedef struct { int nx, ny, nz; } dim3g;

static i64 n3(dim3g d) { return (i64)d.nx * d.ny * d.nz; }
static i64 c3(dim3g d, int i, int j, int k) { return (i64)i + (i64)d.nx * j + (i64)d.nx * d.ny * k; }
static int is3(dim3g d) { return d.nz > 1; }
/* GridBase::isInBounds(Vec3i, bnd) grid.h:630-640 */
static int inb(dim3g d, int i, int j, int k, int b)
{
	int ret = (i >= b && j >= b && i < d.nx - b && j < d.ny - b);
	if (is3(d))
		ret &= (k >= b && k < d.nz - b);
	else
		ret &= (k == 0);
	return ret;
}
/* loop bounds of KERNEL(bnd) / FOR_IJK_BND on a Grid: z is bounded only on a 3D grid (kernel.cpp:20-27, kernel.h:62-68) */
static int zlo(dim3g d, int b) { return is3(d) ? b : 0; }
static int zhi(dim3g d, int b) { return is3(d) ? d.nz - b : 1; }
static float *fal(i64 n) { return (float *)calloc((size_t)(n > 0 ? n : 1), sizeof(float)); }

/* util/interpol.h:57-116 with mStrideZ = 0 on a 2D grid (grid.cpp:64) */
static void interpol(const float *data, dim3g d, int elem, const float pos[3], float *out)
{
	float px = pos[0] - 0.5f, py = pos[1] - 0.5f, pz = pos[2] - 0.5f;
	int xi = (int)px, yi = (int)py, zi = (int)pz;
	float s1 = px - (float)xi, s0 = 1. - s1;
	float t1 = py - (float)yi, t0 = 1. - t1;
	float f1 = pz - (float)zi, f0 = 1. - f1;
	if (px < 0.) { xi = 0; s0 = 1.0; s1 = 0.0; }
	if (py < 0.) { yi = 0; t0 = 1.0; t1 = 0.0; }
	if (pz < 0.) { zi = 0; f0 = 1.0; f1 = 0.0; }
	if (xi >= d.nx - 1) { xi = d.nx - 2; s0 = 0.0; s1 = 1.0; }
	if (yi >= d.ny - 1) { yi = d.ny - 2; t0 = 0.0; t1 = 1.0; }
	if (d.nz > 1) {
		if (zi >= d.nz - 1) { zi = d.nz - 2; f0 = 0.0; f1 = 1.0; }
	}
	const i64 X = 1, Y = d.nx, Z = is3(d) ? (i64)d.nx * d.ny : 0;
	const i64 idx = (i64)xi + Y * yi + Z * zi;
	for (int c = 0; c < elem; ++c) {
#define D(o) data[(idx + (o)) * elem + c]
		out[c] = ((D(0) * t0 + D(Y) * t1) * s0 + (D(X) * t0 + D(X + Y) * t1) * s1) * f0 +
		         ((D(Z) * t0 + D(Y + Z) * t1) * s0 + (D(X + Z) * t0 + D(X + Y + Z) * t1) * s1) * f1;
#undef D
	}
}

/* interpolGridTempl<Grid<T>> grid.h:845-860, knInterpolateGridTempl :830-838, calcGridSizeFactor :490-493 */
static void interpol_grid(float *dst, dim3g td, const float *src, dim3g sd, int elem)
{
	const float fac[3] = { (float)sd.nx / td.nx, (float)sd.ny / td.ny, (float)sd.nz / td.nz };
	const float off[3] = { fac[0] * 0.5f, fac[1] * 0.5f, fac[2] * 0.5f };
	for (int k = 0; k < td.nz; ++k)
		for (int j = 0; j < td.ny; ++j)
			for (int i = 0; i < td.nx; ++i) {
				float pos[3] = { (float)i * fac[0] + off[0], (float)j * fac[1] + off[1], (float)k * fac[2] + off[2] };
				if (!is3(sd)) pos[2] = 0;
				interpol(src, sd, elem, pos, dst + c3(td, i, j, k) * elem);
			}
}

/* Grid<Vec3>::getMax = sqrt(max normSquare) grid.cpp:174-181, 261 */
static float vec_max(const float *v, dim3g d)
{
	float m = -3.402823466e+38f;
	for (i64 c = 0; c < n3(d); ++c) {
		const float s = v[c * 3] * v[c * 3] + v[c * 3 + 1] * v[c * 3 + 1] + v[c * 3 + 2] * v[c * 3 + 2];
		if (s > m) m = s;
	}
	return sqrtf(m);
}

/* advectSemiLagrangeCfl :863-872 -> advectCflHelperCentered :841-853 -> semiLagrangeCent3d :820-825 (KERNEL bnd = 1) */
void orc3_advect_cfl(float cfl, const float *vel3, float *grid, int elem, int nx, int ny, int nz, float velFactor)
{
	const dim3g d = { nx, ny, nz };
	const i64 n = n3(d);
	float *v = fal(n * 3);
	for (i64 q = 0; q < n * 3; ++q) v[q] = vel3[q] * velFactor; /* velTmp.multConst(Vec3(velFactor)) */
	const float orgDt = 1.0f;
	const float maxVel = vec_max(v, d) * orgDt;
	const int steps = (int)(maxVel / cfl) + 1;
	const float dt = orgDt / (float)steps;
	for (int s = 0; s < steps; ++s) {
		float *fwd = fal(n * elem); /* GRID fwd(parent): zero */
		for (int k = zlo(d, 1); k < zhi(d, 1); ++k)
			for (int j = 1; j < ny - 1; ++j)
				for (int i = 1; i < nx - 1; ++i) {
					const i64 c = c3(d, i, j, k);
					const float pos[3] = { (i + 0.5f) - v[c * 3] * dt, (j + 0.5f) - v[c * 3 + 1] * dt, (k + 0.5f) - v[c * 3 + 2] * dt };
					interpol(grid, d, elem, pos, fwd + c * elem);
				}
		memcpy(grid, fwd, sizeof(float) * n * elem); /* orig.swap(fwd) */
		free(fwd);
	}
	free(v);
}

/* knSetBoundaryNeumann grid.cpp:475-506 (serial in-place, index order) */
static void set_bound_neumann(float *a, dim3g d, int w)
{
	for (int k = 0; k < d.nz; ++k)
		for (int j = 0; j < d.ny; ++j)
			for (int i = 0; i < d.nx; ++i) {
				int set = 0, si = i, sj = j, sk = k;
				if (i <= w) { si = w + 1; set = 1; }
				if (i >= d.nx - 1 - w) { si = d.nx - 1 - w - 1; set = 1; }
				if (j <= w) { sj = w + 1; set = 1; }
				if (j >= d.ny - 1 - w) { sj = d.ny - 1 - w - 1; set = 1; }
				if (is3(d)) {
					if (k <= w) { sk = w + 1; set = 1; }
					if (k >= d.nz - 1 - w) { sk = d.nz - 1 - w - 1; set = 1; }
				}
				if (set) a[c3(d, i, j, k)] = a[c3(d, si, sj, sk)];
			}
}

/* calcLsDiff3d :928-933 -> calcLsDiffTempl :895-927 */
float orc3_calc_ls_diff(const float *i0, const float *i1, float *out, int nx, int ny, int nz, float correction, int bnd)
{
	const dim3g d = { nx, ny, nz };
	double accu = 0.;
	for (int k = zlo(d, bnd); k < zhi(d, bnd); ++k)
		for (int j = bnd; j < ny - bnd; ++j)
			for (int i = bnd; i < nx - bnd; ++i) {
				const i64 c = c3(d, i, j, k);
				if ((i0[c] < 0. && i1[c] < 0.) || (i0[c] >= 0. && i1[c] >= 0.)) {
					if (out) out[c] = 0.;
				} else {
					float dv = fabsf(i0[c] - i1[c]) * correction;
					if (dv > 1.) dv = 1.;
					accu += dv;
					if (out) out[c] = dv;
				}
			}
	const int sx = nx - 2 * bnd, sy = ny - 2 * bnd, sz = is3(d) ? nz - 2 * bnd : nz, st = 1;
	accu *= 1000.;
	accu *= 1. / (double)(sx * sy * sz * st);
	return (float)accu;
}

/* gaussianBlurGeneric<Grid<Vec3>> :160-174 with knGaussianBlur :133-158 on a 3D / 2D grid (size[3] = 0, size[2] = 0 in 2D) */
static void gaussian_blur(float *a, dim3g d, float sigma)
{
	const i64 n = n3(d);
	int s = (int)(1. * sigma + 0.5);
	if (s == 0) s = 1;
	const int sz = is3(d) ? s : 0;
	float *bufA = a, *bufT = fal(n * 3);
	for (int numIt = 0; numIt < 2; ++numIt) {
		for (int k = zlo(d, 1); k < zhi(d, 1); ++k)
			for (int j = 1; j < d.ny - 1; ++j)
				for (int i = 1; i < d.nx - 1; ++i) {
					float val[3] = { 0.f, 0.f, 0.f };
					float weight = 0.f;
					for (int zk = k - sz; zk <= k + sz; ++zk)
						for (int yj = j - s; yj <= j + s; ++yj)
							for (int xi = i - s; xi <= i + s; ++xi) {
								if (!inb(d, xi, yj, zk, 0)) continue;
								float dd = (xi - i) * (xi - i) + (yj - j) * (yj - j);
								if (is3(d)) dd += (zk - k) * (zk - k);
								const float wcurr = exp(-dd / (2. * sigma * sigma));
								weight += wcurr;
								const float *src = bufA + c3(d, xi, yj, zk) * 3;
								for (int c = 0; c < 3; ++c) val[c] += wcurr * src[c];
							}
					float *o = bufT + c3(d, i, j, k) * 3;
					if (weight > VECTOR_EPSILON) {
						for (int c = 0; c < 3; ++c) o[c] = val[c] / weight;
					} else {
						for (int c = 0; c < 3; ++c) o[c] = bufA[c3(d, i, j, k) * 3 + c];
					}
				}
		float *sw = bufA;
		bufA = bufT;
		bufT = sw;
	}
	free(bufT); /* two swaps: bufA == a again */
}

/* opticalFlowDim<Grid<Real>, Grid<Vec3>, DIM> :361-553 with FixedMatrixOF / applyMat / GridCGOptflow4d :179-355 */
static int optical_flow(float *vel, const float *i0, const float *i1, dim3g d, float wSmooth, float wEnergy, float postVelBlur,
                        float cgAccuracy, float resetBndWidth)
{
	const int DIM = is3(d) ? 3 : 2;
	const float mDt = 1.0f;
	const float mDx = 1. / d.nx;
	const float mDx2Inv = 1.;
	const int dim[4] = { d.nx, d.ny, d.nz, 1 };
	const i64 cells = n3(d), N = cells * DIM;
	float *offd = fal(N), *blockd = fal(N * DIM), *rhs = fal(N), *result = fal(N);
	static const int nbx[] = { 0, 0, 0, -1, +1, 0, 0, 0 };
	static const int nby[] = { 0, 0, -1, 0, 0, +1, 0, 0 };
	static const int nbz[] = { 0, -1, 0, 0, 0, 0, +1, 0 };
	static const int nbt[] = { -1, 0, 0, 0, 0, 0, 0, +1 };
	i64 indices[8] = { 0 };
	const int p[3] = { 1, 1, is3(d) ? 1 : 0 };
	for (int m = 4 - DIM; m < (4 - DIM) + 2 * DIM; ++m) {
		const int ti = p[0] + nbx[m], tj = p[1] + nby[m], tk = p[2] + nbz[m], tt = nbt[m];
		if (tt == 0 && inb(d, ti, tj, tk, 0)) indices[m] = (c3(d, ti, tj, tk) - c3(d, p[0], p[1], p[2])) * DIM;
	}
	for (int k = 0; k < d.nz; ++k)
		for (int j = 0; j < d.ny; ++j)
			for (int i = 0; i < d.nx; ++i) {
				const i64 c = c3(d, i, j, k);
				if (!inb(d, i, j, k, 1)) {
					for (int dd = 0; dd < DIM; ++dd) blockd[(c * DIM + dd) * DIM + dd] = 1.;
					continue;
				}
				const float tderiv = (i1[c] - i0[c]) / mDt;
				const float dxf = 1. / (2. * mDx);
				float grad[3] = { 0.f, 0.f, 0.f };
				grad[0] = (i1[c3(d, i + 1, j, k)] - i1[c3(d, i - 1, j, k)]) * dxf;
				grad[1] = (i1[c3(d, i, j + 1, k)] - i1[c3(d, i, j - 1, k)]) * dxf;
				if (DIM > 2) grad[2] = (i1[c3(d, i, j, k + 1)] - i1[c3(d, i, j, k - 1)]) * dxf;
				for (int dd = 0; dd < DIM; ++dd) {
					const i64 v = c * DIM + dd;
					rhs[v] = -grad[dd] * tderiv;
					for (int m = 4 - DIM; m < (4 - DIM) + 2 * DIM; ++m) {
						int ti = i + nbx[m], tj = j + nby[m], tk = k + nbz[m];
						ti = ti < 0 ? 0 : (ti > dim[0] - 1 ? dim[0] - 1 : ti);
						tj = tj < 0 ? 0 : (tj > dim[1] - 1 ? dim[1] - 1 : tj);
						tk = tk < 0 ? 0 : (tk > dim[2] - 1 ? dim[2] - 1 : tk);
						rhs[v] -= wSmooth * (vel[c * 3 + dd] - vel[c3(d, ti, tj, tk) * 3 + dd]) * mDx * mDx2Inv;
					}
					offd[v] = -wSmooth * mDx2Inv;
					float diag = 0.;
					diag += (float)(2 * DIM) * wSmooth * mDx2Inv;
					diag += wEnergy;
					for (int m = 0; m < DIM; ++m) {
						if (dd == m)
							blockd[v * DIM + m] = grad[dd] * grad[m] + diag;
						else
							blockd[v * DIM + m] = grad[dd] * grad[m];
					}
					rhs[v] -= wEnergy * vel[c * 3 + dd] * mDx;
				}
			}
	float cgRes = 1e10f;
	int cgIter = -1;
	{
		float *srch = fal(N), *res = fal(N), *tmp = fal(N), *precond = fal(N);
		memcpy(res, rhs, sizeof(float) * N);
		double residual = res[0];
		for (i64 q = 1; q < N; ++q)
			if (res[q] > residual) residual = res[q];
		if (residual < VECTOR_EPSILON) {
			cgRes = 0;
			cgIter = 0;
		} else {
			const double acc = cgAccuracy * residual, resIni = residual;
			for (i64 q = 0; q < N; ++q) precond[q] = 1. / blockd[q * DIM + (q - (q / DIM) * DIM)];
			for (i64 q = 0; q < N; ++q) tmp[q] = res[q] * precond[q];
			double sigma = 0.;
			for (i64 q = 0; q < N; ++q) sigma += tmp[q] * res[q];
			if (sigma == 0 || sigma != sigma) {
				cgIter = 0;
			} else {
				memcpy(srch, tmp, sizeof(float) * N);
				for (int iter = 0; iter < 1000; ++iter) {
					for (i64 q = 0; q < N; ++q) { /* applyMat :211-232 */
						float v = 0;
						const float of = offd[q];
						if (of != 0.)
							for (int m = 4 - DIM; m < (4 - DIM) + 2 * DIM; ++m) v += of * srch[q + indices[m]];
						const i64 blockIdx = (q / DIM) * DIM;
						for (int m = 0; m < DIM; ++m) v += blockd[q * DIM + m] * srch[blockIdx + m];
						tmp[q] = v;
					}
					double alpha1 = 0.;
					for (i64 q = 0; q < N; ++q) alpha1 += srch[q] * tmp[q];
					const double alpha = sigma / alpha1;
					for (i64 q = 0; q < N; ++q) result[q] += alpha * srch[q];
					for (i64 q = 0; q < N; ++q) res[q] += -alpha * tmp[q];
					float mx = res[0];
					for (i64 q = 1; q < N; ++q)
						if (res[q] > mx) mx = res[q];
					residual = mx;
					cgRes = residual / resIni;
					cgIter = iter + 1;
					if (residual <= acc) break;
					for (i64 q = 0; q < N; ++q) tmp[q] = res[q] * precond[q];
					double sigmaNew = 0.;
					for (i64 q = 0; q < N; ++q) sigmaNew += tmp[q] * res[q];
					const double beta = sigmaNew / sigma;
					for (i64 q = 0; q < N; ++q) tmp[q] += beta * srch[q];
					memcpy(srch, tmp, sizeof(float) * N);
					sigma = sigmaNew;
				}
			}
		}
		free(srch);
		free(res);
		free(tmp);
		free(precond);
	}
	if (cgRes != cgRes) memset(result, 0, sizeof(float) * N);
	for (i64 c = 0; c < cells; ++c)
		for (int dd = 0; dd < DIM; ++dd) vel[c * 3 + dd] = result[c * DIM + dd] / mDx;
	free(offd);
	free(blockd);
	free(rhs);
	free(result);
	if (postVelBlur > 0.) gaussian_blur(vel, d, (float)(0.5 * postVelBlur));
	if (resetBndWidth > 0.) {
		const int resetBnd = (int)(resetBndWidth * d.nx) + 1;
		for (int k = 0; k < d.nz; ++k)
			for (int j = 0; j < d.ny; ++j)
				for (int i = 0; i < d.nx; ++i) {
					if (inb(d, i, j, k, resetBnd)) continue;
					float *v = vel + c3(d, i, j, k) * 3;
					v[0] = v[1] = v[2] = 0.f;
				}
	}
	return cgIter;
}

/* normalize(Vector3D) util/vectorbase.h:401-415 */
static void normalize3(float v[3])
{
	const float l = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
	if (fabs(l - 1.) < VECTOR_EPSILON * VECTOR_EPSILON) {
		/* normalized "enough" */
	} else if (l > VECTOR_EPSILON * VECTOR_EPSILON) {
		const float norm = sqrt(l);
		const float fac = 1. / norm;
		for (int c = 0; c < 3; ++c) v[c] *= fac;
	} else {
		v[0] = v[1] = v[2] = 0.f;
	}
}

/* corrVelsOf3d :803-812 -> corrVelsOfTempl :737-802; knProjectCells :713-734 (KERNEL fourd, bnd = 2), getNormalInterpTempl
 * :651-673, projectCell :675-710, knCvExpolBlur3d / 2d :627-649 */
void orc3_corr_vels(float *dst, float *vel, const float *phiOrg, const float *phiTarget, int nx, int ny, int nz, float threshPhi,
                    float postVelBlur, float resetBndWidth, int maxIter)
{
	const dim3g d = { nx, ny, nz };
	const i64 n = n3(d);
	const float blurThreshold = 0.98;
	const float dt = 1.0f;
	const int doCvExtrapol = vec_max(vel, d) + 4;
	float *marker = fal(n), *tmp = fal(n * 3);
	int singleStep = 0;
	if (postVelBlur < blurThreshold) singleStep = 1;
	while ((postVelBlur >= blurThreshold) || singleStep) {
		memset(dst, 0, sizeof(float) * n * 3);
		memset(marker, 0, sizeof(float) * n);
		for (int k = zlo(d, 2); k < zhi(d, 2); ++k)
			for (int j = 2; j < ny - 2; ++j)
				for (int i = 2; i < nx - 2; ++i) {
					const i64 c = c3(d, i, j, k);
					const float *v = vel + c * 3;
					const float pos[3] = { (i + 0.5f) - v[0] * dt, (j + 0.5f) - v[1] * dt, (k + 0.5f) - v[2] * dt };
					const float h = 0.5;
					float nrm[3] = { 0.f, 0.f, 0.f }; /* `VEC n;` is zero-initialised (vectorbase.h:92) */
					for (int a = 0; a < (is3(d) ? 3 : 2); ++a) {
						float pp[3] = { pos[0], pos[1], pos[2] }, pm[3] = { pos[0], pos[1], pos[2] };
						pp[a] = pos[a] + h;
						pm[a] = pos[a] - h;
						float vp, vm;
						interpol(phiOrg, d, 1, pp, &vp);
						interpol(phiOrg, d, 1, pm, &vm);
						nrm[a] = vp - vm;
					}
					normalize3(nrm);
					float step = 0.25;
					if (maxIter < 50) step = 0.5;
					int lastDir = 0;
					const float targetVal = phiTarget[c];
					float vn;
					interpol(phiOrg, d, 1, pos, &vn);
					float dd = vn - targetVal;
					for (int s = 0; s < maxIter; ++s) {
						float pn[3];
						for (int a = 0; a < 3; ++a) pn[a] = pos[a] + (nrm[a] * dd) * dt;
						if (!inb(d, (int)pn[0], (int)pn[1], (int)pn[2], 0)) { /* isInBounds(Vec3): toVec3i truncation, grid.h:72-75 */
							s = maxIter;
							dd = 9999.;
							continue;
						}
						interpol(phiOrg, d, 1, pn, &vn);
						if (vn < targetVal) {
							if (lastDir < 0) step *= 0.5;
							dd += step;
							lastDir = 1;
						} else {
							if (lastDir > 0) step *= 0.5;
							dd -= step;
							lastDir = -1;
						}
						if (step < 1e-01) s = maxIter + 1;
					}
					if (fabsf(dd) > threshPhi) {
						dst[c * 3] = dst[c * 3 + 1] = dst[c * 3 + 2] = 0.f;
						continue;
					}
					for (int a = 0; a < 3; ++a) dst[c * 3 + a] = nrm[a] * dd;
					marker[c] = 1.;
				}
		for (int sIt = 0; sIt < doCvExtrapol; ++sIt) {
			memcpy(tmp, dst, sizeof(float) * n * 3); /* tmp.copyFrom(dst) */
			for (int k = zlo(d, 1); k < zhi(d, 1); ++k)
				for (int j = 1; j < ny - 1; ++j)
					for (int i = 1; i < nx - 1; ++i) {
						const i64 c = c3(d, i, j, k);
						if (marker[c] != 0.) continue;
						float val[3] = { 0.f, 0.f, 0.f };
						for (int zk = (is3(d) ? k - 1 : 0); zk <= (is3(d) ? k + 1 : 0); ++zk)
							for (int yj = j - 1; yj <= j + 1; ++yj)
								for (int xi = i - 1; xi <= i + 1; ++xi) {
									const float *s = dst + c3(d, xi, yj, zk) * 3;
									for (int q = 0; q < 3; ++q) val[q] += s[q];
								}
						for (int q = 0; q < 3; ++q) tmp[c * 3 + q] = val[q] * (is3(d) ? (1. / 27.0) : (1. / 9.0));
					}
			memcpy(dst, tmp, sizeof(float) * n * 3); /* dst.swap(tmp) */
		}
		if (postVelBlur > VECTOR_EPSILON) gaussian_blur(dst, d, (float)(0.5 * postVelBlur));
		const int resetBnd = (int)(resetBndWidth * nx) + 1;
		if (resetBndWidth > 0.) { /* Grid::setBound(VEC(0.), resetBnd): knSetBoundary grid.cpp:462-468 */
			const int w = resetBnd;
			for (int k = 0; k < nz; ++k)
				for (int j = 0; j < ny; ++j)
					for (int i = 0; i < nx; ++i) {
						const int bnd = (i <= w || i >= nx - 1 - w || j <= w || j >= ny - 1 - w || (is3(d) && (k <= w || k >= nz - 1 - w)));
						if (bnd) {
							float *v = dst + c3(d, i, j, k) * 3;
							v[0] = v[1] = v[2] = 0.f;
						}
					}
		}
		for (i64 q = 0; q < n * 3; ++q) vel[q] += dst[q] * -1.f; /* vel.addScaled(dst, VEC(-1.)) grid.h:229-232 */
		postVelBlur *= 0.5;
		if (singleStep > 0) singleStep--;
	}
	free(marker);
	free(tmp);
}

typedef struct {
	int *cgIters, nIters;
	float *errs;
	int nErrs;
} trace3;

/* opticalFlowMultiscaleTemplate<FlagGrid, Grid<Real>, Grid<Vec3>> :936-1173 */
static float multiscale(float *vel, const float *i0, const float *i1, dim3g d, float wSmooth, float wEnergy, int level,
                        float postVelBlur, float cgAccuracy, float cfl, float resetBndWidth, int multiStep, int projSizeThresh,
                        int minGridSize, int doFinalProject, trace3 *tr)
{
	const i64 n = n3(d);
	const int resetBnd = resetBndWidth > 0 ? (int)(resetBndWidth * d.nx) + 1 : 0;
	const float projMaxDist = 4., projMaxIter = 40.;
	const float lsDiffFac = 0.1 / 20.;
	float *i0warped = fal(n);
	memcpy(i0warped, i0, sizeof(float) * n);
	if (d.nx > minGridSize) {
		const dim3g s = { d.nx / 2, d.ny / 2, is3(d) ? d.nz / 2 : 1 };
		const i64 ns = n3(s);
		float *velSm = fal(ns * 3), *i0Sm = fal(ns), *i1Sm = fal(ns);
		interpol_grid(i0Sm, s, i0, d, 1);
		interpol_grid(i1Sm, s, i1, d, 1);
		interpol_grid(velSm, s, vel, d, 3);
		for (i64 q = 0; q < ns * 3; ++q) velSm[q] *= 0.5f;
		multiscale(velSm, i0Sm, i1Sm, s, wSmooth, wEnergy, level + 1, postVelBlur, cgAccuracy, cfl, resetBndWidth, multiStep,
		           projSizeThresh, minGridSize, doFinalProject, tr);
		interpol_grid(vel, d, velSm, s, 3);
		for (i64 q = 0; q < n * 3; ++q) vel[q] *= 2.f;
		free(velSm);
		free(i0Sm);
		free(i1Sm);
	}
	orc3_advect_cfl(cfl, vel, i0warped, 1, d.nx, d.ny, d.nz, 1.f);
	set_bound_neumann(i0warped, d, 0);
	float errCurr = orc3_calc_ls_diff(i0warped, i1, NULL, d.nx, d.ny, d.nz, lsDiffFac, resetBnd);
	int doProject = 0;
	if (d.nx > projSizeThresh) {
		doProject = 1;
		multiStep = 1;
		if (doFinalProject) doFinalProject = 0;
	}
	if (multiStep > 1) {
		float *vs[10], *vs2[10];
		for (int of = 0; of < 10; ++of) vs[of] = vs2[of] = NULL;
		for (int of = 0; of < multiStep; ++of) vs2[of] = fal(n * 3);
		float *tmpVel = fal(n * 3), *i0warp2 = fal(n);
		float velBlur = postVelBlur;
		float errLast = errCurr;
		int ofStepsCurr = multiStep;
		for (int of = 0; of < ofStepsCurr; ++of) {
			vs[of] = fal(n * 3);
			const int it = optical_flow(vs[of], i0warped, i1, d, wSmooth, wEnergy, velBlur, cgAccuracy, resetBndWidth);
			if (tr && tr->cgIters) tr->cgIters[tr->nIters++] = it;
			velBlur *= (3. / 4.);
			if (velBlur < 2.) velBlur = 2.;
			for (int k = of; k >= 0; --k) memcpy(vs2[k], vs[k], sizeof(float) * n * 3);
			for (int k = of - 1; k >= 0; --k)
				for (int l = 0; l < k; ++l) orc3_advect_cfl(cfl, vs2[k], vs2[l], 3, d.nx, d.ny, d.nz, 1.f);
			memcpy(tmpVel, vel, sizeof(float) * n * 3);
			for (int k = of; k >= 0; --k)
				for (i64 q = 0; q < n * 3; ++q) tmpVel[q] += vs2[k][q];
			memcpy(i0warp2, i0, sizeof(float) * n);
			orc3_advect_cfl(cfl, tmpVel, i0warp2, 1, d.nx, d.ny, d.nz, 1.f);
			set_bound_neumann(i0warp2, d, 0);
			const float errC = orc3_calc_ls_diff(i0warp2, i1, NULL, d.nx, d.ny, d.nz, lsDiffFac, resetBnd);
			if (tr && tr->errs) tr->errs[tr->nErrs++] = errC;
			memcpy(i0warped, i0warp2, sizeof(float) * n);
			if (of > 0 && (errC / errLast) > 0.95) {
				memset(vs[of], 0, sizeof(float) * n * 3);
				ofStepsCurr = of + 1;
			}
			errLast = errC;
		}
		for (int of = ofStepsCurr - 1; of >= 0; --of)
			for (int l = 0; l < of; ++l) orc3_advect_cfl(cfl, vs[of], vs[l], 3, d.nx, d.ny, d.nz, 1.f);
		for (int of = 0; of < ofStepsCurr; ++of)
			for (i64 q = 0; q < n * 3; ++q) vel[q] += vs[of][q];
		for (int of = 0; of < 10; ++of) {
			free(vs[of]);
			free(vs2[of]);
		}
		free(tmpVel);
		free(i0warp2);
	} else {
		float *velCurr = fal(n * 3);
		if (!doProject) {
			const int it = optical_flow(velCurr, i0warped, i1, d, wSmooth, wEnergy, postVelBlur, cgAccuracy, resetBndWidth);
			if (tr && tr->cgIters) tr->cgIters[tr->nIters++] = it;
			for (i64 q = 0; q < n * 3; ++q) vel[q] += velCurr[q];
		} else {
			float *velTmp2 = fal(n * 3);
			orc3_corr_vels(velCurr, velTmp2, i0warped, i1, d.nx, d.ny, d.nz, projMaxDist, postVelBlur, resetBndWidth, (int)projMaxIter);
			for (i64 q = 0; q < n * 3; ++q) vel[q] += velTmp2[q];
			for (i64 q = 0; q < n * 3; ++q) velCurr[q] *= -1.f;
			for (i64 q = 0; q < n * 3; ++q) vel[q] += velCurr[q];
			free(velTmp2);
		}
		free(velCurr);
	}
	if ((level == 0) && doFinalProject) {
		float *velCurr = fal(n * 3);
		const float finalProjBlur = 4.;
		orc3_corr_vels(velCurr, vel, i0, i1, d.nx, d.ny, d.nz, projMaxDist, finalProjBlur, resetBndWidth, (int)projMaxIter);
		free(velCurr);
	}
	if (level == 0) {
		memcpy(i0warped, i0, sizeof(float) * n);
		orc3_advect_cfl(cfl, vel, i0warped, 1, d.nx, d.ny, d.nz, 1.f);
		set_bound_neumann(i0warped, d, 0);
		errCurr = orc3_calc_ls_diff(i0warped, i1, NULL, d.nx, d.ny, d.nz, lsDiffFac, resetBnd);
		if (tr && tr->errs) tr->errs[tr->nErrs++] = errCurr;
	}
	free(i0warped);
	return errCurr;
}

/* opticalFlowMultiscale3d :1175-1188.  cgIters / errs (optional, >= 64 entries each) receive the trace. */
float orc3_optical_flow_multiscale(float *vel3, const float *i0, const float *i1, int nx, int ny, int nz, float wSmooth,
                                   float wEnergy, float postVelBlur, float cgAccuracy, float cfl, float resetBndWidth,
                                   int multiStep, int projSizeThresh, int minGridSize, int doFinalProject, int *cgIters,
                                   int *nIters, float *errs, int *nErrs)
{
	const dim3g d = { nx, ny, nz };
	trace3 tr = { cgIters, 0, errs, 0 };
	const float e = multiscale(vel3, i0, i1, d, wSmooth, wEnergy, 0, postVelBlur, cgAccuracy, cfl, resetBndWidth, multiStep,
	                           projSizeThresh, minGridSize, doFinalProject, &tr);
	if (nIters) *nIters = tr.nIters;
	if (nErrs) *nErrs = tr.nErrs;
	return e;
}
