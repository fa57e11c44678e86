/* oracle/flof_oracle.c -- TEST INFRASTRUCTURE.  Plain-C restatement of the reference's
 * FlOF 4D hot path.  See flof_oracle.h for the scope / parity-pin statement.
 *
 * Reference citations are relative to /root/reference/source/.  The float/double mixing of
 * every expression follows the C++ usual arithmetic conversions of the cited line: literals
 * like `1.` and `0.5` are double, `Real` is float, a double result assigned to a Real is
 * rounded once.  Loops that the reference runs as KERNELs are threaded over t with OpenMP
 * exactly like the reference's OPENMP build (preprocessor/codegen_kernel.cpp:159-178); the
 * CG and the matrix assembly are serial, as in the reference.
 */
#include "flof_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define VECTOR_EPSILON (1e-6f) /* util/vectorbase.h:53 */

typedef long long i64;

int orc_set_threads(int n)
{
#ifdef _OPENMP
	if (n > 0) omp_set_num_threads(n);
	return omp_get_max_threads();
#else
	(void)n;
	return 1;
#endif
}

static i64 ncells(orc_dim4 d) { return (i64)d.nx * d.ny * d.nz * d.nt; }
static i64 cidx(orc_dim4 d, int i, int j, int k, int t)
{
	return (i64)i + (i64)d.nx * j + (i64)d.nx * d.ny * k + (i64)d.nx * d.ny * d.nz * t;
}
static int in_bounds(orc_dim4 d, int i, int j, int k, int t, int b)
{ /* grid4d.h:319-326 */
	return i >= b && j >= b && i < d.nx - b && j < d.ny - b && k >= b && k < d.nz - b &&
	       t >= b && t < d.nt - b;
}
static float *falloc(i64 n) { return (float *)calloc((size_t)(n > 0 ? n : 1), sizeof(float)); }

/* ------------------------------------------------------------------ interpolation ---- */

/* util/vector4d.h:422-513 */
void orc_interpol4d(const float *data, orc_dim4 d, int elem, const float pos[4], float *out)
{
	float px = pos[0] - 0.5f, py = pos[1] - 0.5f, pz = pos[2] - 0.5f, pt = pos[3] - 0.5f;
	int xi = (int)px, yi = (int)py, zi = (int)pz, ti = (int)pt;
	float s1 = px - (float)xi, s0 = 1. - s1;
	float t1 = py - (float)yi, t0 = 1. - t1;
	float f1 = pz - (float)zi, f0 = 1. - f1;
	float g1 = pt - (float)ti, g0 = 1. - g1;
	if (px < 0.) { xi = 0; s0 = 1.0; s1 = 0.0; }
	if (py < 0.) { yi = 0; t0 = 1.0; t1 = 0.0; }
	if (pz < 0.) { zi = 0; f0 = 1.0; f1 = 0.0; }
	if (pt < 0.) { ti = 0; g0 = 1.0; g1 = 0.0; }
	if (xi >= d.nx - 1) { xi = d.nx - 2; s0 = 0.0; s1 = 1.0; }
	if (yi >= d.ny - 1) { yi = d.ny - 2; t0 = 0.0; t1 = 1.0; }
	if (zi >= d.nz - 1) { zi = d.nz - 2; f0 = 0.0; f1 = 1.0; }
	if (ti >= d.nt - 1) { ti = d.nt - 2; g0 = 0.0; g1 = 1.0; }
	const i64 sX = 1, sY = d.nx, sZ = (i64)d.nx * d.ny, sT = sZ * d.nz;
	const i64 idx = (i64)xi + sY * yi + sZ * zi + sT * ti;
	for (int c = 0; c < elem; ++c) {
#define D(o) data[(idx + (o)) * elem + c]
		out[c] = (((D(0) * t0 + D(sY) * t1) * s0 + (D(sX) * t0 + D(sX + sY) * t1) * s1) * f0 +
		          ((D(sZ) * t0 + D(sY + sZ) * t1) * s0 +
		           (D(sX + sZ) * t0 + D(sX + sY + sZ) * t1) * s1) *
		              f1) *
		             g0 +
		         (((D(sT) * t0 + D(sT + sY) * t1) * s0 +
		           (D(sT + sX) * t0 + D(sT + sX + sY) * t1) * s1) *
		              f0 +
		          ((D(sT + sZ) * t0 + D(sT + sY + sZ) * t1) * s0 +
		           (D(sT + sX + sZ) * t0 + D(sT + sX + sY + sZ) * t1) * s1) *
		              f1) *
		             g1;
#undef D
	}
}

/* util/interpol.h:57-116 */
void orc_interpol3d(const float *data, int nx, int ny, int nz, int elem, const float pos[3],
                    float *out)
{
	float px = pos[0] - 0.5f, py = pos[1] - 0.5f, pz = pos[2] - 0.5f;
	int xi = (int)px, yi = (int)py, zi = (int)pz;
	float s1 = px - (float)xi, s0 = 1. - s1;
	float t1 = py - (float)yi, t0 = 1. - t1;
	float f1 = pz - (float)zi, f0 = 1. - f1;
	if (px < 0.) { xi = 0; s0 = 1.0; s1 = 0.0; }
	if (py < 0.) { yi = 0; t0 = 1.0; t1 = 0.0; }
	if (pz < 0.) { zi = 0; f0 = 1.0; f1 = 0.0; }
	if (xi >= nx - 1) { xi = nx - 2; s0 = 0.0; s1 = 1.0; }
	if (yi >= ny - 1) { yi = ny - 2; t0 = 0.0; t1 = 1.0; }
	if (nz > 1) {
		if (zi >= nz - 1) { zi = nz - 2; f0 = 0.0; f1 = 1.0; }
	}
	const i64 X = 1, Y = nx, Z = (i64)nx * ny;
	const i64 idx = (i64)xi + Y * yi + Z * zi;
	for (int c = 0; c < elem; ++c) {
#define D(o) data[(idx + (o)) * elem + c]
		out[c] = ((D(0) * t0 + D(Y) * t1) * s0 + (D(X) * t0 + D(X + Y) * t1) * s1) * f0 +
		         ((D(Z) * t0 + D(Y + Z) * t1) * s0 + (D(X + Z) * t0 + D(X + Y + Z) * t1) * s1) *
		             f1;
#undef D
	}
}

/* grid4d.cpp:559-569 (+ grid4d.h:280-283) */
void orc_grid_factor4d(const float s1[4], const float s2in[4], const float optSize[4],
                       const float scale[4], float srcFac[4], float off[4])
{
	for (int c = 0; c < 4; ++c) {
		float s2 = s2in[c];
		if (optSize[c] > 0.) s2 = optSize[c];
		srcFac[c] = (s1[c] / s2) / scale[c];
		float a = -off[c] * srcFac[c];
		float b = srcFac[c] * 0.5; /* Vec4 * double, rounded per component */
		off[c] = a + b;
	}
}

/* grid4d.cpp:531-537 knInterpol4d == grid4d.h:463-471 KnInterpolateGrid4dTempl */
void orc_kn_interpol4d(float *dst, orc_dim4 td, const float *src, orc_dim4 sd, int elem,
                       const float srcFac[4], const float off[4])
{
#pragma omp parallel for schedule(static)
	for (int t = 0; t < td.nt; ++t)
		for (int k = 0; k < td.nz; ++k)
			for (int j = 0; j < td.ny; ++j)
				for (int i = 0; i < td.nx; ++i) {
					float pos[4] = { (float)i * srcFac[0] + off[0], (float)j * srcFac[1] + off[1],
						             (float)k * srcFac[2] + off[2], (float)t * srcFac[3] + off[3] };
					orc_interpol4d(src, sd, elem, pos, dst + cidx(td, i, j, k, t) * elem);
				}
}

void orc_interpolate_grid4d(float *dst, orc_dim4 td, const float *src, orc_dim4 sd, int elem,
                            const float offset[4], const float scale[4], const float size[4])
{
	float s1[4] = { (float)sd.nx, (float)sd.ny, (float)sd.nz, (float)sd.nt };
	float s2[4] = { (float)td.nx, (float)td.ny, (float)td.nz, (float)td.nt };
	float fac[4], off[4] = { offset[0], offset[1], offset[2], offset[3] };
	orc_grid_factor4d(s1, s2, size, scale, fac, off);
	orc_kn_interpol4d(dst, td, src, sd, elem, fac, off);
}

/* optflow4d.cpp:40-57 with grid4d.h:275-279 (integer sizes) */
void orc_interpol_grid_templ(float *dst, orc_dim4 td, const float *src, orc_dim4 sd, int elem)
{
	float fac[4] = { (float)sd.nx / td.nx, (float)sd.ny / td.ny, (float)sd.nz / td.nz,
		             (float)sd.nt / td.nt };
	float off[4];
	for (int c = 0; c < 4; ++c) off[c] = fac[c] * 0.5;
	orc_kn_interpol4d(dst, td, src, sd, elem, fac, off);
}

/* ------------------------------------------------------------------ grid helpers ----- */

/* grid4d.cpp:355-368: note `<= w`, i.e. w+1 layers */
void orc_set_bound4d(float *a, orc_dim4 d, int elem, const float value[4], int w)
{
#pragma omp parallel for schedule(static)
	for (int t = 0; t < d.nt; ++t)
		for (int k = 0; k < d.nz; ++k)
			for (int j = 0; j < d.ny; ++j)
				for (int i = 0; i < d.nx; ++i) {
					int bnd = (i <= w || i >= d.nx - 1 - w || j <= w || j >= d.ny - 1 - w ||
					           k <= w || k >= d.nz - 1 - w || t <= w || t >= d.nt - 1 - w);
					if (bnd)
						for (int c = 0; c < elem; ++c) a[cidx(d, i, j, k, t) * elem + c] = value[c];
				}
}
void orc_set_bound4d_int(int *a, orc_dim4 d, int value, int w)
{
	for (int t = 0; t < d.nt; ++t)
		for (int k = 0; k < d.nz; ++k)
			for (int j = 0; j < d.ny; ++j)
				for (int i = 0; i < d.nx; ++i) {
					int bnd = (i <= w || i >= d.nx - 1 - w || j <= w || j >= d.ny - 1 - w ||
					           k <= w || k >= d.nz - 1 - w || t <= w || t >= d.nt - 1 - w);
					if (bnd) a[cidx(d, i, j, k, t)] = value;
				}
}

/* grid4d.cpp:370-412 */
void orc_set_bound_neumann4d(float *a, orc_dim4 d, int elem, int w)
{
#pragma omp parallel for schedule(static)
	for (int t = 0; t < d.nt; ++t)
		for (int k = 0; k < d.nz; ++k)
			for (int j = 0; j < d.ny; ++j)
				for (int i = 0; i < d.nx; ++i) {
					int set = 0, si = i, sj = j, sk = k, st = t;
					if (i <= w) { si = w + 1; set = 1; }
					if (i >= d.nx - 1 - w) { si = d.nx - 1 - w - 1; set = 1; }
					if (j <= w) { sj = w + 1; set = 1; }
					if (j >= d.ny - 1 - w) { sj = d.ny - 1 - w - 1; set = 1; }
					if (k <= w) { sk = w + 1; set = 1; }
					if (k >= d.nz - 1 - w) { sk = d.nz - 1 - w - 1; set = 1; }
					if (t <= w) { st = w + 1; set = 1; }
					if (t >= d.nt - 1 - w) { st = d.nt - 1 - w - 1; set = 1; }
					if (set)
						for (int c = 0; c < elem; ++c)
							a[cidx(d, i, j, k, t) * elem + c] = a[cidx(d, si, sj, sk, st) * elem + c];
				}
}

/* grid4d.cpp:143-191, 266-285 */
void orc_min_max4d(const float *a, orc_dim4 d, int elem, float out[3])
{
	const i64 n = ncells(d);
	float mn = FLT_MAX, mx = -FLT_MAX;
	if (elem == 1) {
		for (i64 i = 0; i < n; ++i) {
			if (a[i] < mn) mn = a[i];
			if (a[i] > mx) mx = a[i];
		}
		out[0] = mn;
		out[1] = mx;
		float fa = fabsf(mn), fb = fabsf(mx);
		out[2] = fa > fb ? fa : fb;
	} else {
		for (i64 i = 0; i < n; ++i) {
			const float *v = a + i * 4;
			const float s = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
			if (s < mn) mn = s;
			if (s > mx) mx = s;
		}
		out[0] = sqrtf(mn);
		out[1] = sqrtf(mx);
		out[2] = sqrtf(mx);
	}
}

/* grid4d.h:338-372, grid4d.cpp:213-264 */
void orc_grid_op4d(float *a, const float *b, orc_dim4 d, int elem, int op, const float f[4])
{
	const i64 n = ncells(d);
#pragma omp parallel for schedule(static)
	for (i64 i = 0; i < n; ++i)
		for (int c = 0; c < elem; ++c) {
			const i64 x = i * elem + c;
			switch (op) {
			case 0: a[x] += b[x]; break;
			case 1: a[x] -= b[x]; break;
			case 2: a[x] *= b[x]; break;
			case 3: a[x] += f[c] * b[x]; break;
			case 4: a[x] *= f[c]; break;
			case 5: a[x] += f[c]; break;
			case 6: a[x] = a[x] < f[0] ? f[0] : (a[x] > f[1] ? f[1] : a[x]); break;
			}
		}
}

/* ------------------------------------------------------------------ advection -------- */

/* optflow4d.cpp:1275-1290: bnd 1, result in a fresh zeroed grid that replaces `grid` */
void orc_advect4d(const float *vel, float *grid, orc_dim4 d, int elem, float dt)
{
	const i64 n = ncells(d);
	float *fwd = falloc(n * elem);
#pragma omp parallel for schedule(static)
	for (int t = 1; t < d.nt - 1; ++t)
		for (int k = 1; k < d.nz - 1; ++k)
			for (int j = 1; j < d.ny - 1; ++j)
				for (int i = 1; i < d.nx - 1; ++i) {
					const i64 c = cidx(d, i, j, k, t);
					const float *v = vel + c * 4;
					float pos[4] = { (i + 0.5f) - v[0] * dt, (j + 0.5f) - v[1] * dt,
						             (k + 0.5f) - v[2] * dt, (t + 0.5f) - v[3] * dt };
					orc_interpol4d(grid, d, elem, pos, fwd + c * elem);
				}
	memcpy(grid, fwd, sizeof(float) * n * elem);
	free(fwd);
}

/* optflow4d.cpp:2170-2180 + 1309-1321 (solver dt is 1) */
void orc_advect_cfl4d(float cfl, const float *vel, float *grid, orc_dim4 d, int elem,
                      float velFactor)
{
	const i64 n = ncells(d);
	float *vt = (float *)malloc(sizeof(float) * n * 4);
	for (i64 i = 0; i < n * 4; ++i) vt[i] = vel[i] * velFactor;
	float mm[3];
	orc_min_max4d(vt, d, 4, mm);
	const float orgDt = 1.0f;
	const float maxVel = mm[1] * orgDt;
	const int steps = (int)(maxVel / cfl) + 1;
	const float dt = orgDt / (float)steps;
	for (int s = 0; s < steps; ++s) orc_advect4d(vt, grid, d, elem, dt * 1.f);
	free(vt);
}

/* ------------------------------------------------------------------ Gaussian blur ---- */

/* optflow4d.cpp:128-174 */
void orc_gaussian_blur4d(float *a, orc_dim4 d, int elem, float sigma, int iter)
{
	const i64 n = ncells(d);
	int s = (int)(1. * sigma + 0.5);
	if (s == 0) s = 1;
	float *bufA = a, *bufT = falloc(n * elem);
	/* weights depend only on the integer squared distance (optflow4d.cpp:144-148) */
	const int maxd = 4 * s * s;
	float *wtab = (float *)malloc(sizeof(float) * (maxd + 1));
	for (int q = 0; q <= maxd; ++q) {
		const float dSqr = (float)q;
		wtab[q] = (float)exp(-dSqr / (2. * sigma * sigma));
	}
	for (int numIt = 0; numIt < 2 * iter; ++numIt) {
#pragma omp parallel for schedule(static)
		for (int t = 1; t < d.nt - 1; ++t)
			for (int k = 1; k < d.nz - 1; ++k)
				for (int j = 1; j < d.ny - 1; ++j)
					for (int i = 1; i < d.nx - 1; ++i) {
						float val[4] = { 0.f, 0.f, 0.f, 0.f };
						float weight = 0.f;
						for (int vt = t - s; vt <= t + s; ++vt)
							for (int zk = k - s; zk <= k + s; ++zk)
								for (int yj = j - s; yj <= j + s; ++yj)
									for (int xi = i - s; xi <= i + s; ++xi) {
										if (!in_bounds(d, xi, yj, zk, vt, 0)) continue;
										const int q = (xi - i) * (xi - i) + (yj - j) * (yj - j) +
										              (zk - k) * (zk - k) + (vt - t) * (vt - t);
										const float wcurr = wtab[q];
										weight += wcurr;
										const float *src = bufA + cidx(d, xi, yj, zk, vt) * elem;
										for (int c = 0; c < elem; ++c) val[c] += wcurr * src[c];
									}
						float *out = bufT + cidx(d, i, j, k, t) * elem;
						if (weight > VECTOR_EPSILON) {
							for (int c = 0; c < elem; ++c) out[c] = val[c] / weight;
						} else {
							for (int c = 0; c < elem; ++c) out[c] = bufA[cidx(d, i, j, k, t) * elem + c];
						}
					}
		float *sw = bufA;
		bufA = bufT;
		bufT = sw; /* a.swap(tmp) */
	}
	/* 2*iter swaps: bufA == a again, bufT is the scratch grid */
	free(bufT);
	free(wtab);
}

/* ------------------------------------------------------------------ OF solve ---------- */

/* optflow4d.cpp:211-232 with the matrix recomputed on the fly (identical fp32 values) */
typedef struct {
	orc_dim4 d;
	const float *offd;   /* N   */
	const float *blockd; /* 4N  */
	i64 off[8];
} orc_mat;

static void apply_mat(const orc_mat *m, const float *x, float *y, i64 N)
{
	for (i64 i = 0; i < N; ++i) {
		float v = 0;
		const float offd = m->offd[i];
		if (offd != 0.) {
			for (int q = 0; q < 8; ++q) v += offd * x[i + m->off[q]];
		}
		const i64 blockIdx = (i / 4) * 4;
		for (int q = 0; q < 4; ++q) v += m->blockd[i * 4 + q] * x[blockIdx + q];
		y[i] = v;
	}
}
/* Sensitivity probe (tests only): 0 = the reference's sequential fp64 sum, 1 = the same products
 * summed in 4096-element blocks first.  Used to measure how much the reference's own result moves
 * when only the order of the fp64 dot-product summation changes (DESIGN.md "conditioning"). */
static int g_dot_mode = 0;
void orc_set_dot_mode(int m) { g_dot_mode = m; }

/* test probes for the CUDA path's sequential-order dot products (flof_dot_seq): the reference's loops, nothing else.
 * kind 0: dotProd(a, b) (:234-241).  kind 1: precondInit + precondApply + dotProd(tmp, res) (:331-354, :296, :319) with
 * a = res and the Jacobi diagonal of the matrix-free form: grad_d^2 + diag, 1 on identity rows (grad.x = NaN marker). */
double orc_dot_seq(const float *a, const float *b, long long cells, int kind, float diag)
{
	double d = 0.;
	for (i64 c = 0; c < cells; ++c)
		for (int q = 0; q < 4; ++q) {
			const i64 i = c * 4 + q;
			if (kind == 0) {
				d += a[i] * b[i];
			} else {
				const float g = b[i];
				float pc = 1.f;
				if (b[c * 4] == b[c * 4]) pc = 1. / (g * g + diag);
				const float t = a[i] * pc;
				d += t * a[i];
			}
		}
	return d;
}
static double dot_prod(const float *a, const float *b, i64 N)
{
	if (g_dot_mode == 1) {
		double tot = 0.;
		for (i64 s = 0; s < N; s += 4096) {
			double blk = 0.;
			const i64 e = s + 4096 < N ? s + 4096 : N;
			for (i64 i = s; i < e; ++i) blk += a[i] * b[i];
			tot += blk;
		}
		return tot;
	}
	double d = 0.;
	for (i64 i = 0; i < N; ++i) d += a[i] * b[i]; /* fp32 product, fp64 sum (:234-241) */
	return d;
}
static float max_norm(const float *a, i64 N)
{
	float m = a[0];
	for (i64 i = 1; i < N; ++i)
		if (a[i] > m) m = a[i];
	return m; /* signed (:243-250) */
}
static void add_scaled(float *a, double b, const float *c, i64 N)
{
	for (i64 i = 0; i < N; ++i) a[i] += b * c[i]; /* double, rounded once (:253-258) */
}

int orc_optical_flow4d(float *vel, const float *i0, const float *i1, float *rhsT, orc_dim4 d,
                       float wSmooth, float wEnergy, float postVelBlur, float cgAccuracy,
                       float resetBndWidth, float *cgResOut)
{
	const int DIM = 4;
	const float mDt = 1.0f;
	const float mDx = 1. / d.nx;
	const float mDx2Inv = 1.;
	const int dim[4] = { d.nx, d.ny, d.nz, d.nt };
	const i64 cells = ncells(d), N = cells * DIM;
	float *offd = falloc(N), *blockd = falloc(N * 4), *rhs = falloc(N), *result = falloc(N);
	static const int nbx[] = { 0, 0, 0, -1, +1, 0, 0, 0 };
	static const int nby[] = { 0, 0, -1, 0, 0, +1, 0, 0 };
	static const int nbz[] = { 0, -1, 0, 0, 0, 0, +1, 0 };
	static const int nbt[] = { -1, 0, 0, 0, 0, 0, 0, +1 };
	orc_mat A;
	A.d = d;
	A.offd = offd;
	A.blockd = blockd;
	for (int m = 0; m < 8; ++m)
		A.off[m] = (cidx(d, 1 + nbx[m], 1 + nby[m], 1 + nbz[m], 1 + nbt[m]) - cidx(d, 1, 1, 1, 1)) * DIM;

	/* :422-493 */
	for (int t = 0; t < d.nt; ++t)
		for (int k = 0; k < d.nz; ++k)
			for (int j = 0; j < d.ny; ++j)
				for (int i = 0; i < d.nx; ++i) {
					const i64 c = cidx(d, i, j, k, t);
					if (!in_bounds(d, i, j, k, t, 1)) {
						for (int dd = 0; dd < DIM; ++dd) blockd[(c * DIM + dd) * 4 + dd] = 1.;
						continue;
					}
					const float tderiv = (i1[c] - i0[c]) / mDt;
					const float dxf = 1. / (2. * mDx);
					float grad[4];
					grad[0] = (i1[cidx(d, i + 1, j, k, t)] - i1[cidx(d, i - 1, j, k, t)]) * dxf;
					grad[1] = (i1[cidx(d, i, j + 1, k, t)] - i1[cidx(d, i, j - 1, k, t)]) * dxf;
					grad[2] = (i1[cidx(d, i, j, k + 1, t)] - i1[cidx(d, i, j, k - 1, t)]) * dxf;
					grad[3] = (i1[cidx(d, i, j, k, t + 1)] - i1[cidx(d, i, j, k, t - 1)]) * dxf;
					for (int dd = 0; dd < DIM; ++dd) {
						const i64 v = c * DIM + dd;
						rhs[v] = -grad[dd] * tderiv;
						for (int m = 0; m < 8; ++m) {
							int ti = i + nbx[m], tj = j + nby[m], tk = k + nbz[m], tt = t + nbt[m];
							ti = ti < 0 ? 0 : (ti > dim[0] - 1 ? dim[0] - 1 : ti);
							tj = tj < 0 ? 0 : (tj > dim[1] - 1 ? dim[1] - 1 : tj);
							tk = tk < 0 ? 0 : (tk > dim[2] - 1 ? dim[2] - 1 : tk);
							tt = tt < 0 ? 0 : (tt > dim[3] - 1 ? dim[3] - 1 : tt);
							rhs[v] -= wSmooth * (vel[c * 4 + dd] - vel[cidx(d, ti, tj, tk, tt) * 4 + dd]) *
							          mDx * mDx2Inv;
						}
						offd[v] = -wSmooth * mDx2Inv;
						float diag = 0.;
						diag += (float)(2 * DIM) * wSmooth * mDx2Inv;
						diag += wEnergy;
						for (int m = 0; m < DIM; ++m) {
							if (dd == m)
								blockd[v * 4 + m] = grad[dd] * grad[m] + diag;
							else
								blockd[v * 4 + m] = grad[dd] * grad[m];
						}
						rhs[v] -= wEnergy * vel[c * 4 + dd] * mDx;
					}
				}

	/* :274-329 GridCGOptflow4d::solve */
	float cgRes = 1e10f;
	int cgIter = -1;
	{
		float *srch = falloc(N), *res = falloc(N), *tmp = falloc(N), *precond = falloc(N);
		memset(result, 0, sizeof(float) * N);
		memcpy(res, rhs, sizeof(float) * N);
		double residual = max_norm(res, N);
		if (residual < VECTOR_EPSILON) {
			cgRes = 0;
			cgIter = 0;
		} else {
			const double acc = cgAccuracy * residual;
			const double resIni = residual;
			for (i64 i = 0; i < N; ++i) {
				const float diag = blockd[i * 4 + (i - (i / DIM) * DIM)];
				precond[i] = 1. / diag;
			}
			for (i64 i = 0; i < N; ++i) tmp[i] = res[i] * precond[i];
			double sigma = dot_prod(tmp, res, N);
			if (sigma == 0 || sigma != sigma) {
				cgIter = 0;
			} else {
				memcpy(srch, tmp, sizeof(float) * N);
				int done = 0;
				for (int iter = 0; iter < 1000 && !done; ++iter) {
					apply_mat(&A, srch, tmp, N);
					const double alpha1 = dot_prod(srch, tmp, N);
					const double alpha = sigma / alpha1;
					add_scaled(result, alpha, srch, N);
					add_scaled(res, -alpha, tmp, N);
					residual = max_norm(res, N);
					cgRes = residual / resIni;
					cgIter = iter + 1;
					if (residual <= acc) {
						done = 1;
						break;
					}
					for (i64 i = 0; i < N; ++i) tmp[i] = res[i] * precond[i];
					const double sigmaNew = dot_prod(tmp, res, N);
					const double beta = sigmaNew / sigma;
					add_scaled(tmp, beta, srch, N);
					memcpy(srch, tmp, sizeof(float) * N);
					sigma = sigmaNew;
				}
				if (!done) cgRes = residual / resIni;
			}
		}
		free(srch);
		free(res);
		free(tmp);
		free(precond);
	}
	if (cgRes != cgRes) memset(result, 0, sizeof(float) * N); /* :509-514 */

	/* :520-529 */
	for (i64 c = 0; c < cells; ++c) {
		for (int dd = 0; dd < DIM; ++dd) vel[c * 4 + dd] = result[c * DIM + dd] / mDx;
		if (rhsT) rhsT[c] = rhs[c * DIM];
	}
	free(offd);
	free(blockd);
	free(rhs);
	free(result);

	/* :532-541 */
	if (postVelBlur > 0.) orc_gaussian_blur4d(vel, d, 4, (float)(0.5 * postVelBlur), 1);
	/* :544-551 */
	if (resetBndWidth > 0.) {
		const int resetBnd = (int)(resetBndWidth * d.nx) + 1;
		for (int t = 0; t < d.nt; ++t)
			for (int k = 0; k < d.nz; ++k)
				for (int j = 0; j < d.ny; ++j)
					for (int i = 0; i < d.nx; ++i) {
						if (in_bounds(d, i, j, k, t, resetBnd)) continue;
						float *v = vel + cidx(d, i, j, k, t) * 4;
						v[0] = v[1] = v[2] = v[3] = 0.f;
					}
	}
	if (cgResOut) *cgResOut = cgRes;
	return cgIter;
}

/* ------------------------------------------------------------------ projection -------- */

/* util/vector4d.h:331-345 */
static void normalize4(float v[4])
{
	const float l = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
	if (fabs(l - 1.) < VECTOR_EPSILON * VECTOR_EPSILON) {
		/* unchanged */
	} else if (l > VECTOR_EPSILON * VECTOR_EPSILON) {
		const float norm = sqrtf(l);
		const float fac = 1. / norm;
		for (int c = 0; c < 4; ++c) v[c] *= fac;
	} else {
		v[0] = v[1] = v[2] = v[3] = 0.f;
	}
}

/* optflow4d.cpp:651-734 */
void orc_project_cells(float *dst, const float *vel, const float *phiOrg,
                       const float *phiTarget, float *marker, orc_dim4 d, float threshPhi,
                       int maxIter)
{
	const float dt = 1.0f;
#pragma omp parallel for schedule(static)
	for (int t = 2; t < d.nt - 2; ++t)
		for (int k = 2; k < d.nz - 2; ++k)
			for (int j = 2; j < d.ny - 2; ++j)
				for (int i = 2; i < d.nx - 2; ++i) {
					const i64 c = cidx(d, i, j, k, t);
					const float *v = vel + c * 4;
					const float pos[4] = { (i + 0.5f) - v[0] * dt, (j + 0.5f) - v[1] * dt,
						                   (k + 0.5f) - v[2] * dt, (t + 0.5f) - v[3] * dt };
					/* getNormalInterpTempl :651-673 */
					const float h = 0.5;
					float n[4];
					for (int a = 0; a < 4; ++a) {
						float pp[4] = { pos[0], pos[1], pos[2], pos[3] }, pm[4] = { pos[0], pos[1], pos[2], pos[3] };
						pp[a] = pos[a] + h;
						pm[a] = pos[a] - h;
						float vp, vm;
						orc_interpol4d(phiOrg, d, 1, pp, &vp);
						orc_interpol4d(phiOrg, d, 1, pm, &vm);
						n[a] = vp - vm;
					}
					normalize4(n);
					/* projectCell :675-710 */
					float step = 0.25;
					if (maxIter < 50) step = 0.5;
					int lastDir = 0;
					const float targetVal = phiTarget[c];
					float vn;
					orc_interpol4d(phiOrg, d, 1, pos, &vn);
					float dd = vn - targetVal;
					for (int s = 0; s < maxIter; ++s) {
						float pn[4];
						for (int a = 0; a < 4; ++a) pn[a] = pos[a] + (n[a] * dd) * dt;
						if (!in_bounds(d, (int)pn[0], (int)pn[1], (int)pn[2], (int)pn[3], 0)) {
							s = maxIter;
							dd = 9999.;
							continue;
						}
						orc_interpol4d(phiOrg, d, 1, pn, &vn);
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
						dst[c * 4 + 0] = dst[c * 4 + 1] = dst[c * 4 + 2] = dst[c * 4 + 3] = 0.f;
						continue;
					}
					for (int a = 0; a < 4; ++a) dst[c * 4 + a] = n[a] * dd;
					marker[c] = 1.;
				}
}

/* optflow4d.cpp:613-626 wrapped as in :770-780 */
void orc_cv_expol_blur4d(float *a, const float *marker, orc_dim4 d, int sweeps)
{
	const i64 n = ncells(d);
	float *cur = a, *tmp = (float *)malloc(sizeof(float) * n * 4);
	for (int sIt = 0; sIt < sweeps; ++sIt) {
		memcpy(tmp, cur, sizeof(float) * n * 4);
#pragma omp parallel for schedule(static)
		for (int t = 1; t < d.nt - 1; ++t)
			for (int k = 1; k < d.nz - 1; ++k)
				for (int j = 1; j < d.ny - 1; ++j)
					for (int i = 1; i < d.nx - 1; ++i) {
						const i64 c = cidx(d, i, j, k, t);
						if (marker[c] != 0.) continue;
						float val[4] = { 0.f, 0.f, 0.f, 0.f };
						for (int vt = t - 1; vt <= t + 1; ++vt)
							for (int zk = k - 1; zk <= k + 1; ++zk)
								for (int yj = j - 1; yj <= j + 1; ++yj)
									for (int xi = i - 1; xi <= i + 1; ++xi) {
										const float *s = cur + cidx(d, xi, yj, zk, vt) * 4;
										for (int q = 0; q < 4; ++q) val[q] += s[q];
									}
						for (int q = 0; q < 4; ++q) tmp[c * 4 + q] = val[q] * (1. / 81.0);
					}
		float *sw = cur;
		cur = tmp;
		tmp = sw;
	}
	if (cur != a) {
		memcpy(a, cur, sizeof(float) * n * 4);
		free(cur);
	} else {
		free(tmp);
	}
}

/* optflow4d.cpp:737-802 */
void orc_corr_vels_of4d(float *dst, float *vel, const float *phiOrg, const float *phiTarget,
                        orc_dim4 d, float threshPhi, float postVelBlur, float resetBndWidth,
                        int maxIter)
{
	const i64 n = ncells(d);
	const float blurThreshold = 0.98;
	float mm[3];
	orc_min_max4d(vel, d, 4, mm);
	const int doCvExtrapol = mm[2] + 4;
	float *marker = falloc(n);
	int singleStep = 0;
	if (postVelBlur < blurThreshold) singleStep = 1;
	while ((postVelBlur >= blurThreshold) || singleStep) {
		memset(dst, 0, sizeof(float) * n * 4);
		memset(marker, 0, sizeof(float) * n);
		orc_project_cells(dst, vel, phiOrg, phiTarget, marker, d, threshPhi, maxIter);
		if (doCvExtrapol) orc_cv_expol_blur4d(dst, marker, d, doCvExtrapol);
		if (postVelBlur > VECTOR_EPSILON)
			orc_gaussian_blur4d(dst, d, 4, (float)(0.5 * postVelBlur), 1);
		const int resetBnd = (int)(resetBndWidth * d.nx) + 1;
		const float zero[4] = { 0.f, 0.f, 0.f, 0.f };
		if (resetBndWidth > 0.) orc_set_bound4d(dst, d, 4, zero, resetBnd);
		const float m1[4] = { -1.f, -1.f, -1.f, -1.f };
		orc_grid_op4d(vel, dst, d, 4, 3, m1);
		postVelBlur *= 0.5;
		if (singleStep > 0) singleStep--;
	}
	free(marker);
}

/* ------------------------------------------------------------------ error metric ------ */

/* optflow4d.cpp:895-927 */
float orc_calc_ls_diff4d(const float *i0, const float *i1, float *out, orc_dim4 d,
                         float correction, int bnd)
{
	double accu = 0.;
	for (int t = bnd; t < d.nt - bnd; ++t)
		for (int k = bnd; k < d.nz - bnd; ++k)
			for (int j = bnd; j < d.ny - bnd; ++j)
				for (int i = bnd; i < d.nx - bnd; ++i) {
					const i64 c = cidx(d, i, j, k, t);
					if ((i0[c] < 0. && i1[c] < 0.) || (i0[c] >= 0. && i1[c] >= 0.)) {
						if (out) out[c] = 0.;
					} else {
						float dv = fabsf(i0[c] - i1[c]) * correction;
						if (dv > 1.) dv = 1.;
						accu += dv;
						if (out) out[c] = dv;
					}
				}
	const int sx = d.nx - 2 * bnd, sy = d.ny - 2 * bnd, sz = d.nz - 2 * bnd, st = d.nt - 2 * bnd;
	accu *= 1000.;
	if (d.nt > 1) accu *= 1000.;
	accu *= 1. / (double)(sx * sy * sz * st);
	return (float)accu;
}

/* ------------------------------------------------------------------ multi-scale ------- */

typedef struct {
	int *cgIters, nIters;
	float *errs;
	int nErrs;
} orc_trace;

static void tr_iter(orc_trace *tr, int it)
{
	if (tr && tr->cgIters && tr->nIters < 64) tr->cgIters[tr->nIters] = it;
	if (tr) tr->nIters++;
}
static void tr_err(orc_trace *tr, float e)
{
	if (tr && tr->errs && tr->nErrs < 64) tr->errs[tr->nErrs] = e;
	if (tr) tr->nErrs++;
}

/* optflow4d.cpp:936-1173 */
static float multiscale(float *vel, const float *i0, const float *i1, orc_dim4 d, float wSmooth,
                        float wEnergy, int level, float postVelBlur, float cgAccuracy, float cfl,
                        float resetBndWidth, int multiStep, int projSizeThresh, int minGridSize,
                        int doFinalProject, orc_trace *tr)
{
	const i64 n = ncells(d);
	const int resetBnd = resetBndWidth > 0 ? (int)(resetBndWidth * d.nx) + 1 : 0;
	const float projMaxDist = 4.;
	const float projMaxIter = 40.;
	const float lsDiffFac = 0.1 / 20.;
	const float one4[4] = { 1.f, 1.f, 1.f, 1.f };
	(void)one4;

	float *i0warped = (float *)malloc(sizeof(float) * n);
	memcpy(i0warped, i0, sizeof(float) * n);
	const float errPreOf = orc_calc_ls_diff4d(i0, i1, NULL, d, lsDiffFac, resetBnd);

	if (d.nx > minGridSize) {
		orc_dim4 s = { d.nx / 2, d.ny / 2, d.nz / 2, d.nt / 2 };
		const i64 ns = ncells(s);
		float *velSm = falloc(ns * 4), *i0Sm = falloc(ns), *i1Sm = falloc(ns);
		orc_interpol_grid_templ(i0Sm, s, i0, d, 1);
		orc_interpol_grid_templ(i1Sm, s, i1, d, 1);
		orc_interpol_grid_templ(velSm, s, vel, d, 4);
		const float half[4] = { 0.5f, 0.5f, 0.5f, 0.5f };
		orc_grid_op4d(velSm, NULL, s, 4, 4, half);
		multiscale(velSm, i0Sm, i1Sm, s, wSmooth, wEnergy, level + 1, postVelBlur, cgAccuracy, cfl,
		           resetBndWidth, multiStep, projSizeThresh, minGridSize, doFinalProject, tr);
		orc_interpol_grid_templ(vel, d, velSm, s, 4);
		const float two[4] = { 2.f, 2.f, 2.f, 2.f };
		orc_grid_op4d(vel, NULL, d, 4, 4, two);
		free(velSm);
		free(i0Sm);
		free(i1Sm);
	}

	orc_advect_cfl4d(cfl, vel, i0warped, d, 1, 1.f);
	orc_set_bound_neumann4d(i0warped, d, 1, 0);
	float errCurr = orc_calc_ls_diff4d(i0warped, i1, NULL, d, lsDiffFac, resetBnd);

	int doProject = 0;
	if (d.nx > projSizeThresh) {
		doProject = 1;
		multiStep = 1;
		if (doFinalProject) doFinalProject = 0;
	}

	if (multiStep > 1) {
		float *vs[10], *vs2[10];
		for (int of = 0; of < multiStep; ++of) {
			vs[of] = NULL;
			vs2[of] = falloc(n * 4);
		}
		float *tmpVel = (float *)malloc(sizeof(float) * n * 4);
		memcpy(tmpVel, vel, sizeof(float) * n * 4);
		float velBlur = postVelBlur;
		float errLast = errCurr;
		int ofStepsCurr = multiStep;
		for (int of = 0; of < ofStepsCurr; ++of) {
			vs[of] = falloc(n * 4);
			tr_iter(tr, orc_optical_flow4d(vs[of], i0warped, i1, NULL, d, wSmooth, wEnergy, velBlur,
			                               cgAccuracy, resetBndWidth, NULL));
			velBlur *= (3. / 4.);
			if (velBlur < 2.) velBlur = 2.;
			for (int k = of; k >= 0; --k) memcpy(vs2[k], vs[k], sizeof(float) * n * 4);
			for (int k = of - 1; k >= 0; --k)
				for (int l = 0; l < k; ++l) orc_advect_cfl4d(cfl, vs2[k], vs2[l], d, 4, 1.f);
			memcpy(tmpVel, vel, sizeof(float) * n * 4);
			for (int k = of; k >= 0; --k) orc_grid_op4d(tmpVel, vs2[k], d, 4, 0, NULL);
			float *i0warp2 = (float *)malloc(sizeof(float) * n);
			memcpy(i0warp2, i0, sizeof(float) * n);
			orc_advect_cfl4d(cfl, tmpVel, i0warp2, d, 1, 1.f);
			orc_set_bound_neumann4d(i0warp2, d, 1, 0);
			const float errC = orc_calc_ls_diff4d(i0warp2, i1, NULL, d, lsDiffFac, resetBnd);
			tr_err(tr, errC);
			memcpy(i0warped, i0warp2, sizeof(float) * n);
			free(i0warp2);
			if (of > 0 && (errC / errLast) > 0.95) {
				memset(vs[of], 0, sizeof(float) * n * 4);
				ofStepsCurr = of + 1;
			}
			errLast = errC;
		}
		for (int of = ofStepsCurr - 1; of >= 0; --of)
			for (int l = 0; l < of; ++l) orc_advect_cfl4d(cfl, vs[of], vs[l], d, 4, 1.f);
		for (int of = 0; of < ofStepsCurr; ++of) {
			orc_grid_op4d(vel, vs[of], d, 4, 0, NULL);
			free(vs[of]);
		}
		for (int of = 0; of < multiStep; ++of) free(vs2[of]);
		free(tmpVel);
	} else {
		float *velCurr = falloc(n * 4);
		if (!doProject) {
			tr_iter(tr, orc_optical_flow4d(velCurr, i0warped, i1, NULL, d, wSmooth, wEnergy,
			                               postVelBlur, cgAccuracy, resetBndWidth, NULL));
			orc_grid_op4d(vel, velCurr, d, 4, 0, NULL);
		} else {
			float *velTmp2 = falloc(n * 4);
			orc_corr_vels_of4d(velCurr, velTmp2, i0warped, i1, d, projMaxDist, postVelBlur,
			                   resetBndWidth, (int)projMaxIter);
			orc_grid_op4d(vel, velTmp2, d, 4, 0, NULL);
			const float m1[4] = { -1.f, -1.f, -1.f, -1.f };
			orc_grid_op4d(velCurr, NULL, d, 4, 4, m1);
			orc_grid_op4d(vel, velCurr, d, 4, 0, NULL);
			free(velTmp2);
		}
		free(velCurr);
	}

	if ((level == 0) && doFinalProject) {
		float *velCurr = falloc(n * 4);
		const float finalProjBlur = 4.;
		orc_corr_vels_of4d(velCurr, vel, i0, i1, d, projMaxDist, finalProjBlur, resetBndWidth,
		                   (int)projMaxIter);
		free(velCurr);
	}

	if (level == 0) {
		memcpy(i0warped, i0, sizeof(float) * n);
		orc_advect_cfl4d(cfl, vel, i0warped, d, 1, 1.f);
		orc_set_bound_neumann4d(i0warped, d, 1, 0);
		const float errFinal = orc_calc_ls_diff4d(i0warped, i1, NULL, d, lsDiffFac, resetBnd);
		errCurr = errFinal;
		tr_err(tr, errFinal);
		(void)errPreOf;
	}
	free(i0warped);
	return errCurr;
}

float orc_optical_flow_multiscale4d(float *vel, const float *i0, const float *i1, orc_dim4 d,
                                    float wSmooth, float wEnergy, float postVelBlur,
                                    float cgAccuracy, float cfl, float resetBndWidth,
                                    int multiStep, int projSizeThresh, int minGridSize,
                                    int doFinalProject, int *cgIters, int *nIters,
                                    float *errs, int *nErrs)
{
	orc_trace tr = { cgIters, 0, errs, 0 };
	const float e = multiscale(vel, i0, i1, d, wSmooth, wEnergy, 0, postVelBlur, cgAccuracy, cfl,
	                           resetBndWidth, multiStep, projSizeThresh, minGridSize,
	                           doFinalProject, &tr);
	if (nIters) *nIters = tr.nIters;
	if (nErrs) *nErrs = tr.nErrs;
	return e;
}

/* ------------------------------------------------------------------ pre-processing ---- */

/* optflow4d.cpp:1254-1266 (sequential, in place, like the reference) */
void orc_repeat_frame4d(float *phi, orc_dim4 d, float srct, float range, int bnd)
{
	for (int t = (int)(srct - range + 0.5); t <= (int)(srct); t++)
		for (int k = bnd; k < d.nz - bnd; ++k)
			for (int j = bnd; j < d.ny - bnd; j++)
				for (int i = bnd; i < d.nx - bnd; i++) {
					const float pos[4] = { (float)i + 0.5f, (float)j + 0.5f, (float)k + 0.5f,
						                   srct + 0.5f };
					float v;
					orc_interpol4d(phi, d, 1, pos, &v);
					phi[cidx(d, i, j, k, t)] = v;
				}
}

static const int NBS4D[8][4] = { { -1, 0, 0, 0 }, { 1, 0, 0, 0 }, { 0, -1, 0, 0 }, { 0, 1, 0, 0 },
	                             { 0, 0, -1, 0 }, { 0, 0, 1, 0 }, { 0, 0, 0, -1 }, { 0, 0, 0, 1 } };

/* optflow4d.cpp:1330-1352 */
static void kn_extrap(float *val, int elem, int *tmp, orc_dim4 d, int dd, const float *direction)
{
	for (int t = 1; t < d.nt - 1; ++t)
		for (int k = 1; k < d.nz - 1; ++k)
			for (int j = 1; j < d.ny - 1; ++j)
				for (int i = 1; i < d.nx - 1; ++i) {
					const i64 c = cidx(d, i, j, k, t);
					if (tmp[c] != 0) continue;
					int nbs = 0;
					float avg[4] = { 0.f, 0.f, 0.f, 0.f };
					for (int n = 0; n < 8; ++n) {
						const i64 q = cidx(d, i + NBS4D[n][0], j + NBS4D[n][1], k + NBS4D[n][2],
						                   t + NBS4D[n][3]);
						if (tmp[q] == dd) {
							for (int e = 0; e < elem; ++e) avg[e] += val[q * elem + e];
							nbs++;
						}
					}
					if (nbs > 0) {
						tmp[c] = dd + 1;
						for (int e = 0; e < elem; ++e) val[c * elem + e] = avg[e] / nbs + direction[e];
					}
				}
}

static void mark_first_ring(int *tmp, orc_dim4 d)
{ /* optflow4d.cpp:1387-1397 */
	for (int t = 1; t < d.nt - 1; ++t)
		for (int k = 1; k < d.nz - 1; ++k)
			for (int j = 1; j < d.ny - 1; ++j)
				for (int i = 1; i < d.nx - 1; ++i) {
					const i64 c = cidx(d, i, j, k, t);
					if (tmp[c]) continue;
					for (int n = 0; n < 8; ++n) {
						const i64 q = cidx(d, i + NBS4D[n][0], j + NBS4D[n][1], k + NBS4D[n][2],
						                   t + NBS4D[n][3]);
						if (tmp[q] == 1) {
							tmp[c] = 2;
							n = 8;
						}
					}
				}
}

/* optflow4d.cpp:1361-1406 */
void orc_extrap4d_ls_simple(float *phi, orc_dim4 d, int distance, int inside, int *marker)
{
	const i64 n = ncells(d);
	int *tmp = (int *)calloc((size_t)n, sizeof(int));
	float direction = 1.;
	if (inside) direction = -1.;
	for (int t = 1; t < d.nt - 1; ++t)
		for (int k = 1; k < d.nz - 1; ++k)
			for (int j = 1; j < d.ny - 1; ++j)
				for (int i = 1; i < d.nx - 1; ++i) {
					const i64 c = cidx(d, i, j, k, t);
					if (!inside && phi[c] < 0.) tmp[c] = 1;
					if (inside && phi[c] > 0.) tmp[c] = 1;
				}
	mark_first_ring(tmp, d);
	for (int dd = 2; dd < 1 + distance; ++dd) kn_extrap(phi, 1, tmp, d, dd, &direction);
	const float set = (float)(direction * (distance + 2));
	for (int t = 1; t < d.nt - 1; ++t)
		for (int k = 1; k < d.nz - 1; ++k)
			for (int j = 1; j < d.ny - 1; ++j)
				for (int i = 1; i < d.nx - 1; ++i) {
					const i64 c = cidx(d, i, j, k, t);
					if (tmp[c] != 0) continue;
					phi[c] = set;
				}
	if (marker) memcpy(marker, tmp, sizeof(int) * n);
	free(tmp);
}

/* optflow4d.cpp:1408-1439 */
void orc_extrapolate_vec4_simple(float *vel, const float *phi, orc_dim4 d, int distance)
{
	const i64 n = ncells(d);
	int *tmp = (int *)calloc((size_t)n, sizeof(int));
	const float zero[4] = { 0.f, 0.f, 0.f, 0.f };
	for (int t = 1; t < d.nt - 1; ++t)
		for (int k = 1; k < d.nz - 1; ++k)
			for (int j = 1; j < d.ny - 1; ++j)
				for (int i = 1; i < d.nx - 1; ++i) {
					const i64 c = cidx(d, i, j, k, t);
					if (phi[c] < 0.) tmp[c] = 1;
				}
	mark_first_ring(tmp, d);
	for (int dd = 2; dd < distance + 1; ++dd) kn_extrap(vel, 4, tmp, d, dd, zero);
	for (int t = 1; t < d.nt - 1; ++t)
		for (int k = 1; k < d.nz - 1; ++k)
			for (int j = 1; j < d.ny - 1; ++j)
				for (int i = 1; i < d.nx - 1; ++i) {
					const i64 c = cidx(d, i, j, k, t);
					if (tmp[c] != 0) continue;
					vel[c * 4 + 0] = vel[c * 4 + 1] = vel[c * 4 + 2] = vel[c * 4 + 3] = 0.f;
				}
	free(tmp);
}

/* ------------------------------------------------------------------ 3D output ops ----- */

/* test.cpp:93-131 */
void orc_simple_blur_special(float *a, int nx, int ny, int nz, int iter, float thresh, int bord)
{
	const i64 n = (i64)nx * ny * nz;
	float *cur = a, *tmp = falloc(n);
	for (int it = 0; it < iter; ++it) {
#pragma omp parallel for schedule(static)
		for (int k = 1; k < nz - 1; ++k)
			for (int j = 1; j < ny - 1; ++j)
				for (int i = 1; i < nx - 1; ++i) {
					const i64 c = (i64)i + (i64)nx * j + (i64)nx * ny * k;
					const int inb = i >= bord && j >= bord && k >= bord && i < nx - bord &&
					                j < ny - bord && k < nz - bord;
					if (!inb || cur[c] < thresh) {
						tmp[c] = cur[c];
						continue;
					}
					float v = cur[c + 1] + cur[c - 1] + cur[c + nx] + cur[c - nx];
					v += cur[c + (i64)nx * ny] + cur[c - (i64)nx * ny];
					v *= 1. / 6.;
					tmp[c] = v;
				}
		float *sw = cur;
		cur = tmp;
		tmp = sw;
	}
	if (cur != a) {
		memcpy(a, cur, sizeof(float) * n);
		/* the grid that was `a` now plays tmp: after an odd number of swaps its content is
		 * irrelevant to the caller */
		free(cur);
	} else {
		free(tmp);
	}
}

/* grid.cpp:462-473 */
void orc_grid3_set_bound(float *a, int nx, int ny, int nz, float value, int w)
{
	for (int k = 0; k < nz; ++k)
		for (int j = 0; j < ny; ++j)
			for (int i = 0; i < nx; ++i) {
				int bnd = (i <= w || i >= nx - 1 - w || j <= w || j >= ny - 1 - w ||
				           (nz > 1 && (k <= w || k >= nz - 1 - w)));
				if (bnd) a[(i64)i + (i64)nx * j + (i64)nx * ny * k] = value;
			}
}

/* levelset.cpp:114-118 */
void orc_levelset_join(float *a, const float *b, long n)
{
	for (long i = 0; i < n; ++i) a[i] = a[i] < b[i] ? a[i] : b[i];
}

/* ------------------------------------------------------------------ mode-3 loaders ---- */

/* optflow4d.cpp:1464-1595 */
void orc_load_place_grid4d(const float *slices, int sx, int sy, int sz, float *phi, orc_dim4 d,
                           const float offset[4], const float scale[4], int fileIdxStart,
                           int fileIdxEnd, int debugSkipLoad, float spread,
                           const float overrideSize[4], float overrideTimeOff,
                           int overrideGoodRegion, float loadTimeScale, int rescaleSdfValues,
                           float sdfIsoOff, float repeatStartFrame)
{
	const i64 n3 = (i64)d.nx * d.ny * d.nz, ns = (i64)sx * sy * sz;
	float v1[4] = { (float)sx, (float)sy, (float)sz, 1.f };
	float defoSize[4] = { (float)d.nx, (float)d.ny, (float)d.nz, (float)d.nt };
	if (overrideSize[0] > 0.) memcpy(defoSize, overrideSize, sizeof(defoSize));
	const float m1[4] = { -1.f, -1.f, -1.f, -1.f };
	float sourceFactor[4], off2[4] = { offset[0], offset[1], offset[2], offset[3] };
	orc_grid_factor4d(v1, defoSize, m1, scale, sourceFactor, off2);
	float valueScale = 1.;
	if (rescaleSdfValues) valueScale = 1. / sourceFactor[0];
	float *w = falloc(d.nt);
	float *tmp = (float *)malloc(sizeof(float) * ns);

	if (overrideGoodRegion <= 0)
		memset(phi, 0, sizeof(float) * n3 * d.nt);
	else
		memset(phi + n3 * overrideGoodRegion, 0, sizeof(float) * n3 * (d.nt - overrideGoodRegion));
	if (fileIdxStart < 0) fileIdxStart = 0;
	if (fileIdxEnd < 0) fileIdxEnd = defoSize[3];
	int repeatOff = 0;
	if (repeatStartFrame > 0) repeatOff = (int)((fileIdxEnd - fileIdxStart) * repeatStartFrame + 1.0);

	const int fend = fileIdxEnd < debugSkipLoad ? fileIdxEnd : debugSkipLoad;
	for (int fileid = fileIdxStart - repeatOff; fileid < fend; ++fileid) {
		const int fileidClamp = fileid > fileIdxStart ? fileid : fileIdxStart;
		int didLoad = 0;
		const float duration =
		    1. * (defoSize[3] * scale[3] / loadTimeScale) / (float)(fileIdxEnd - fileIdxStart);
		const float dstt = offset[3] + overrideTimeOff + duration * (fileid - fileIdxStart);
		for (int slice = (int)(dstt - spread * duration) + 0; slice < (int)(dstt + spread * duration) + 2;
		     ++slice) {
			if (slice < overrideGoodRegion || slice >= d.nt) continue;
			const float wc = 1. - fabsf((float)slice - dstt) / (duration * spread); /* std::fabs(float) */
			if (wc < VECTOR_EPSILON) continue;
			w[slice] += wc;
			if (!didLoad) {
				memcpy(tmp, slices + ns * (fileidClamp - fileIdxStart), sizeof(float) * ns);
				if (rescaleSdfValues && sdfIsoOff != 0.) {
					for (int k = 1; k < sz - 1; ++k)
						for (int j = 1; j < sy - 1; ++j)
							for (int i = 1; i < sx - 1; ++i)
								tmp[(i64)i + (i64)sx * j + (i64)sx * sy * k] += sdfIsoOff;
					orc_grid3_set_bound(tmp, sx, sy, sz, 1.f, 1);
				}
				didLoad = 1;
			}
			/* knLoadPlaceGridIpol :1445-1452, bnd 1 over the 3D extent of phi */
#pragma omp parallel for schedule(static)
			for (int k = 1; k < d.nz - 1; ++k)
				for (int j = 1; j < d.ny - 1; ++j)
					for (int i = 1; i < d.nx - 1; ++i) {
						const float p[3] = { (float)i * sourceFactor[0] + off2[0],
							                 (float)j * sourceFactor[1] + off2[1],
							                 (float)k * sourceFactor[2] + off2[2] };
						float v;
						orc_interpol3d(tmp, sx, sy, sz, 1, p, &v);
						phi[cidx(d, i, j, k, slice)] += wc * v;
					}
		}
	}
	for (int slice = 0; slice < d.nt; ++slice) {
		if (w[slice] < 1e-03) {
			w[slice] = 0.;
		} else {
			w[slice] = 1. / w[slice];
			w[slice] *= valueScale;
		}
	}
	/* knLoadPlaceGridRescale :1454-1459, fourd bnd 1 */
	for (int t = 1; t < d.nt - 1; ++t) {
		if (w[t] <= 0.) continue;
		for (int k = 1; k < d.nz - 1; ++k)
			for (int j = 1; j < d.ny - 1; ++j)
				for (int i = 1; i < d.nx - 1; ++i) phi[cidx(d, i, j, k, t)] *= w[t];
	}
	free(w);
	free(tmp);
}

/* optflow4d.cpp:1598-1616 */
void orc_shift_forw_grid4d(float *phi, orc_dim4 d, int overrideGoodRegion)
{
	const i64 n3 = (i64)d.nx * d.ny * d.nz;
	const int shift = d.nt - overrideGoodRegion;
	for (int slice = 0; slice < overrideGoodRegion; slice += shift) {
		int len = shift;
		if (slice + shift >= overrideGoodRegion) len = overrideGoodRegion - slice;
		memmove(phi + n3 * slice, phi + n3 * (slice + shift), sizeof(float) * n3 * len);
	}
}

/* optflow4d.cpp:1951-2105 (useDefoVols == false) + :1648-1665 */
void orc_load_advect_time_slice(const float *defo, orc_dim4 dd, float *dst, int nx, int ny,
                                int nz, const float *phi, orc_dim4 d, float time,
                                float blendAlpha, float loadTimeScale,
                                const float defoOffset[4], const float defoScale[4],
                                const float defoFactor[4], const float overrideSize[4],
                                float overrideTimeOff, int bordSkip, float defoAniFac)
{
	const int dimT = dd.nt;
	float dim4[4] = { (float)dd.nx, (float)dd.ny, (float)dd.nz, (float)dimT };
	float defoSize[4] = { (float)d.nx, (float)d.ny, (float)d.nz, (float)d.nt };
	if (overrideSize[0] > 0.)
		for (int c = 0; c < 4; ++c) defoSize[c] = (float)(int)overrideSize[c];
	const float m1[4] = { -1.f, -1.f, -1.f, -1.f };
	float sourceFactor[4], off2[4] = { defoOffset[0], defoOffset[1], defoOffset[2], defoOffset[3] };
	orc_grid_factor4d(dim4, defoSize, m1, defoScale, sourceFactor, off2);

	const float srcTime = (time)*sourceFactor[3] * loadTimeScale + off2[3] - 0.5;
	int t = (int)srcTime;
	int tp1 = t + 1;
	const float tw = srcTime - (float)t;
	t = t < dimT - 1 ? t : dimT - 1;
	tp1 = tp1 < dimT - 1 ? tp1 : dimT - 1;

	const i64 nv = (i64)dd.nx * dd.ny * dd.nz;
	const float *v1 = defo + nv * 4 * t, *v2 = defo + nv * 4 * tp1;
	float *vt = falloc(nv * 4);
	const float f1 = 1. - tw, f2 = tw;
	for (i64 i = 0; i < nv * 4; ++i) {
		vt[i] += f1 * v1[i];
		vt[i] += f2 * v2[i];
	}
	float fac[4];
	for (int c = 0; c < 4; ++c) fac[c] = defoFactor[c] * defoAniFac;
	const float tm = time + overrideTimeOff;
	const float dt = blendAlpha;
	const int b = bordSkip > 10 ? bordSkip : 10; /* KERNEL(bnd = 10) and the bordSkip test */
#pragma omp parallel for schedule(static)
	for (int k = b; k < nz - b; ++k)
		for (int j = b; j < ny - b; ++j)
			for (int i = b; i < nx - b; ++i) {
				const float p3[3] = { (float)i * sourceFactor[0] + off2[0],
					                  (float)j * sourceFactor[1] + off2[1],
					                  (float)k * sourceFactor[2] + off2[2] };
				float v[4];
				orc_interpol3d(vt, dd.nx, dd.ny, dd.nz, 4, p3, v);
				for (int c = 0; c < 4; ++c) v[c] *= fac[c];
				const float p4[4] = { (i + 0.5f) - v[0] * dt, (j + 0.5f) - v[1] * dt,
					                  (k + 0.5f) - v[2] * dt, (tm + 0.5f) - v[3] * dt };
				orc_interpol4d(phi, d, 1, p4, dst + ((i64)i + (i64)nx * j + (i64)nx * ny * k));
			}
	free(vt);
}

/* ---- deformation volumes: optflow4d.cpp:1822-1863 (updateDefoVol), :1951-2105 with useDefoVols ---------------------
 * vols[dv]: the complete deformation files (dd.nt slices each); the windows and lastT live for the n frames of this call,
 * exactly like the per-ID LoadAdvectData between _OptInit and _Finish. */
/* vt: the one-slice scratch grid (lats.tmp) every slice passes through; *filepos: next slice of the re-used gz handle
 * (fileio.cpp:903-919: once the handle exists, readGrid4dUni ignores the requested slice index and reads the next one;
 * past the end of the file it reads nothing and vt keeps its content), -1 = not open yet */
static void orc_defovol_update(float *win, i64 nv4, int Tw, const float *vol, int dimT, int t, int lastT, float *vt,
                               int *filepos)
{
	const int currt = Tw / 2;
	if (lastT == t) return;
	if (lastT + 1 == t) {
		int tl = 0;
		for (; tl < Tw - 1; ++tl) memcpy(win + nv4 * tl, win + nv4 * (tl + 1), sizeof(float) * nv4);
		tl = Tw - 1;
		int it = t - currt + tl;
		it = it < 0 ? 0 : it;
		it = it < dimT - 1 ? it : dimT - 1;
		if (*filepos >= 0) it = *filepos;
		if (it < dimT) {
			memcpy(vt, vol + nv4 * it, sizeof(float) * nv4);
			*filepos = it + 1;
		} else
			*filepos = it;
		memcpy(win + nv4 * tl, vt, sizeof(float) * nv4);
		return;
	}
	for (int tl = 0; tl < Tw; ++tl) {
		int it = t - currt + tl;
		it = it < 0 ? 0 : it;
		it = it < dimT - 1 ? it : dimT - 1;
		memcpy(vt, vol + nv4 * it, sizeof(float) * nv4);
		memcpy(win + nv4 * tl, vt, sizeof(float) * nv4);
	}
}

void orc_load_advect_defovols(const float *const *vols, int numDv, orc_dim4 dd, int doAligned, float partialLoadFac,
                              float *dst, int nx, int ny, int nz, const float *phi, orc_dim4 d, int n,
                              const float *times, float blendAlphaIn, float thirdAlpha, float fourthAlpha,
                              float loadTimeScale, const float defoOffset[4], const float defoScale[4],
                              const float defoFactor[4], const float overrideSize[4], float overrideTimeOff,
                              int bordSkip, float defoAniFac)
{
	const int dimT = dd.nt;
	const float defoVolWidth = 0.2f > partialLoadFac ? 0.2f : partialLoadFac; /* :1893 */
	const int Tw = (int)(dimT * defoVolWidth);
	const orc_dim4 wd = { dd.nx, dd.ny, dd.nz, Tw };
	const i64 nv = (i64)dd.nx * dd.ny * dd.nz, nv4 = nv * 4;
	float *win[3] = { NULL, NULL, NULL };
	for (int dv = 0; dv < numDv; ++dv) win[dv] = falloc(nv4 * Tw);
	float *dvt = doAligned ? falloc(nv4 * Tw) : NULL;
	float *vt = falloc(nv4);
	int lastT = -1;
	int filepos[3] = { -1, -1, -1 };
	for (int f = 0; f < n; ++f) {
		const float time = times[f];
		float blendAlpha = blendAlphaIn;
		float dim4[4] = { (float)dd.nx, (float)dd.ny, (float)dd.nz, (float)dimT };
		float defoSize[4] = { (float)d.nx, (float)d.ny, (float)d.nz, (float)d.nt };
		if (overrideSize[0] > 0.)
			for (int c = 0; c < 4; ++c) defoSize[c] = (float)(int)overrideSize[c];
		const float m1[4] = { -1.f, -1.f, -1.f, -1.f };
		float sourceFactor[4], off2[4] = { defoOffset[0], defoOffset[1], defoOffset[2], defoOffset[3] };
		orc_grid_factor4d(dim4, defoSize, m1, defoScale, sourceFactor, off2);
		const float srcTime = (time)*sourceFactor[3] * loadTimeScale + off2[3] - 0.5;
		int t = (int)srcTime;
		t = t < dimT - 1 ? t : dimT - 1;
		/* updateDefoVol(t) */
		const int defovolOff = t - Tw / 2;
		for (int dv = 0; dv < numDv; ++dv) orc_defovol_update(win[dv], nv4, Tw, vols[dv], dimT, t, lastT, vt, &filepos[dv]);
		const float tc = srcTime - defovolOff;
		if (numDv == 2 && !doAligned) {
#pragma omp parallel for schedule(static)
			for (int k = 0; k < dd.nz; ++k)
				for (int j = 0; j < dd.ny; ++j)
					for (int i = 0; i < dd.nx; ++i) {
						const float q[4] = { i + 0.5f, j + 0.5f, k + 0.5f, tc + 0.5f };
						float v2[4], a[4], r[4];
						orc_interpol4d(win[1], wd, 4, q, v2);
						for (int c = 0; c < 4; ++c) r[c] = q[c] - v2[c];
						orc_interpol4d(win[0], wd, 4, r, a);
						float *o = vt + 4 * ((i64)i + (i64)dd.nx * (j + (i64)dd.ny * k));
						for (int c = 0; c < 4; ++c) {
							const float x = blendAlpha * a[c], y = thirdAlpha * v2[c];
							o[c] = x + y;
						}
					}
		} else if (numDv == 2) {
#pragma omp parallel for schedule(static)
			for (int tl = 0; tl < Tw; ++tl)
				for (int k = 0; k < dd.nz; ++k)
					for (int j = 0; j < dd.ny; ++j)
						for (int i = 0; i < dd.nx; ++i) {
							const float q[4] = { i + 0.5f, j + 0.5f, k + 0.5f, tl + 0.5f };
							float v2[4], a[4], r[4];
							orc_interpol4d(win[1], wd, 4, q, v2);
							for (int c = 0; c < 4; ++c) r[c] = q[c] - v2[c];
							orc_interpol4d(win[0], wd, 4, r, a);
							float *o = dvt + 4 * ((i64)i + (i64)dd.nx * (j + (i64)dd.ny * (k + (i64)dd.nz * tl)));
							for (int c = 0; c < 4; ++c) {
								const float x = blendAlpha * a[c], y = thirdAlpha * v2[c];
								o[c] = x + y;
							}
						}
#pragma omp parallel for schedule(static)
			for (int k = 0; k < dd.nz; ++k)
				for (int j = 0; j < dd.ny; ++j)
					for (int i = 0; i < dd.nx; ++i) {
						const float q[4] = { i + 0.5f, j + 0.5f, k + 0.5f, tc + 0.5f };
						float v1[4], r[4];
						orc_interpol4d(win[0], wd, 4, q, v1);
						for (int c = 0; c < 4; ++c) {
							const float neg = v1[c] * -1;                       /* invert! */
							const float s = (float)((1. - blendAlpha) * neg);    /* double scalar * float component */
							r[c] = q[c] - s;
						}
						orc_interpol4d(dvt, wd, 4, r, vt + 4 * ((i64)i + (i64)dd.nx * (j + (i64)dd.ny * k)));
					}
		} else { /* numDv == 3 */
#pragma omp parallel for schedule(static)
			for (int k = 0; k < dd.nz; ++k)
				for (int j = 0; j < dd.ny; ++j)
					for (int i = 0; i < dd.nx; ++i) {
						const float q[4] = { i + 0.5f, j + 0.5f, k + 0.5f, tc + 0.5f };
						float v3[4], v2[4], v1[4], r[4], r2[4];
						orc_interpol4d(win[2], wd, 4, q, v3);
						for (int c = 0; c < 4; ++c) r[c] = q[c] - v3[c];
						orc_interpol4d(win[1], wd, 4, r, v2);
						for (int c = 0; c < 4; ++c) r2[c] = r[c] - v2[c];
						orc_interpol4d(win[0], wd, 4, r2, v1);
						float *o = vt + 4 * ((i64)i + (i64)dd.nx * (j + (i64)dd.ny * k));
						for (int c = 0; c < 4; ++c) {
							const float x = blendAlpha * v1[c], y = thirdAlpha * v2[c], z = fourthAlpha * v3[c];
							const float xy = x + y;
							o[c] = xy + z;
						}
					}
		}
		blendAlpha = 1.; /* :2088 already accumulated in vt */
		lastT = t;
		float fac[4];
		for (int c = 0; c < 4; ++c) fac[c] = defoFactor[c] * defoAniFac;
		const float tm = time + overrideTimeOff;
		const float dt = blendAlpha;
		const int b = bordSkip > 10 ? bordSkip : 10;
		float *out = dst + (i64)f * nx * ny * nz;
#pragma omp parallel for schedule(static)
		for (int k = b; k < nz - b; ++k)
			for (int j = b; j < ny - b; ++j)
				for (int i = b; i < nx - b; ++i) {
					const float p3[3] = { (float)i * sourceFactor[0] + off2[0], (float)j * sourceFactor[1] + off2[1],
						                  (float)k * sourceFactor[2] + off2[2] };
					float v[4];
					orc_interpol3d(vt, dd.nx, dd.ny, dd.nz, 4, p3, v);
					for (int c = 0; c < 4; ++c) v[c] *= fac[c];
					const float p4[4] = { (i + 0.5f) - v[0] * dt, (j + 0.5f) - v[1] * dt, (k + 0.5f) - v[2] * dt,
						                  (tm + 0.5f) - v[3] * dt };
					orc_interpol4d(phi, d, 1, p4, out + ((i64)i + (i64)nx * j + (i64)nx * ny * k));
				}
	}
	for (int dv = 0; dv < numDv; ++dv) free(win[dv]);
	free(dvt);
	free(vt);
}

/* optflow4d.cpp:1671-1760: the unoptimised loadAdvectTimeSlice -- the time-blended deformation slice, re-interpolated to
 * the output size (interior cells, bnd 1; vdst is cleared first) and scaled, returned as dbgVel (3 floats per cell) /
 * dbgVelT (either may be NULL), then knSemiLagrangeLookupSlice4d :1622-1629 over the interior cells of dst (the
 * generated KERNEL(fourd, bnd = 1) code runs a one-slice grid as a 3D kernel with t = 0). */
void orc_load_advect_time_slice_unopt(const float *defo, orc_dim4 dd, float *dst, float *dbgVel, float *dbgVelT, int nx,
                                      int ny, int nz, const float *phi, orc_dim4 d, float time, float blendAlpha,
                                      float loadTimeScale, const float defoOffset[4], const float defoScale[4],
                                      const float defoFactor[4], const float overrideSize[4], float overrideTimeOff,
                                      float defoAniFac, int zeroVel)
{
	const int dimT = dd.nt;
	float dim4[4] = { (float)dd.nx, (float)dd.ny, (float)dd.nz, (float)dimT };
	float defoSize[4] = { (float)d.nx, (float)d.ny, (float)d.nz, (float)d.nt };
	if (overrideSize[0] > 0.)
		for (int c = 0; c < 4; ++c) defoSize[c] = (float)(int)overrideSize[c];
	const float m1[4] = { -1.f, -1.f, -1.f, -1.f };
	float sourceFactor[4], off2[4] = { defoOffset[0], defoOffset[1], defoOffset[2], defoOffset[3] };
	orc_grid_factor4d(dim4, defoSize, m1, defoScale, sourceFactor, off2);
	const float srcTime = (time)*sourceFactor[3] * loadTimeScale + off2[3] - 0.5;
	int t = (int)srcTime;
	int tp1 = t + 1;
	const float tw = srcTime - (float)t;
	t = t < dimT - 1 ? t : dimT - 1;
	tp1 = tp1 < dimT - 1 ? tp1 : dimT - 1;
	const i64 nv = (i64)dd.nx * dd.ny * dd.nz;
	const float *s1 = defo + nv * 4 * t, *s2 = defo + nv * 4 * tp1;
	float *v1 = falloc(nv * 4);
	const float f1 = 1. - tw, f2 = tw;
	for (i64 i = 0; i < nv * 4; ++i) { /* v1.multConst(1-tw); v2.multConst(tw); v1.add(v2) */
		const float a = s1[i] * f1, b = s2[i] * f2;
		v1[i] = zeroVel ? 0.f : a + b;
	}
	for (i64 c = 0; c < (i64)nx * ny * nz; ++c) {
		if (dbgVel) dbgVel[3 * c] = dbgVel[3 * c + 1] = dbgVel[3 * c + 2] = 0.f;
		if (dbgVelT) dbgVelT[c] = 0.f;
	}
	for (int k = 1; k < nz - 1; ++k)
		for (int j = 1; j < ny - 1; ++j)
			for (int i = 1; i < nx - 1; ++i) {
				const float p3[3] = { (float)i * sourceFactor[0] + off2[0], (float)j * sourceFactor[1] + off2[1],
					                  (float)k * sourceFactor[2] + off2[2] };
				float v[4];
				orc_interpol3d(v1, dd.nx, dd.ny, dd.nz, 4, p3, v);
				for (int c = 0; c < 4; ++c) v[c] *= defoFactor[c] * defoAniFac;
				const i64 o = (i64)i + (i64)nx * (j + (i64)ny * k);
				if (dbgVel) {
					dbgVel[3 * o] = v[0];
					dbgVel[3 * o + 1] = v[1];
					dbgVel[3 * o + 2] = v[2];
				}
				if (dbgVelT) dbgVelT[o] = v[3];
				if (dst) {
					const float tm = time + overrideTimeOff, dt = blendAlpha;
					const float p4[4] = { (i + 0.5f) - v[0] * dt, (j + 0.5f) - v[1] * dt, (k + 0.5f) - v[2] * dt,
						                  (tm + 0.5f) - v[3] * dt };
					orc_interpol4d(phi, d, 1, p4, dst + o);
				}
			}
	free(v1);
}
