/* oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, never part of the product path.
 *
 * Thin C-ABI glue around the UNMODIFIED reference implementation (thunil/ofblend,
 * mantaflow, compiled from /root/reference in NOPYTHON mode by oracle/Makefile).
 * Every function here only copies host arrays into reference Grid4d<T>/Grid<T>
 * objects, calls the reference plugin function named in its comment, and copies the
 * result back.  No algorithmic code lives in this file.
 *
 * The reference plugin translation unit is #included (from the generated copy under
 * oracle/_ref/build) so that the file-local templates (gaussianBlurGeneric,
 * knProjectCells, knCvExpolBlur4d, ...) can be probed one by one.
 */
#include "plugin/optflow4d.cpp" /* reference source/plugin/optflow4d.cpp after `prep` */

#include <cstdio>
#include <cstring>
#include <string>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace Manta
{
/* reference plugins defined in other translation units (NOPYTHON => plain C++) */
void interpolateGrid4d(Grid4d<Real> &target, Grid4d<Real> &source, Vec4 offset,
                       Vec4 scale, Vec4 size);
void interpolateGrid4dVec(Grid4d<Vec4> &target, Grid4d<Vec4> &source,
                          Vec4 offset, Vec4 scale, Vec4 size);
void simpleBlurSpecial(Grid<Real> &a, int iter, Real thresh, int bord);
void initTestCheckerboard(Grid4d<Real> &val, Grid4d<Vec4> *vec, int brd);
Real debugGridAvg4d(Grid4d<Real> &phi, int brd);
void getSliceFrom4d(Grid4d<Real> &src, int srct, Grid<Real> &dst);
void getSliceFrom4dVec(Grid4d<Vec4> &src, int srct, Grid<Vec3> &dst,
                       Grid<Real> *dstt);
void setDebugLevel(int level);
} // namespace Manta

using namespace Manta;

static char g_err[2048] = { 0 };

#define REF_TRY try {
#define REF_CATCH                                                              \
	}                                                                          \
	catch (std::exception & e)                                                 \
	{                                                                          \
		snprintf(g_err, sizeof(g_err), "%s", e.what());                        \
		return -1;                                                             \
	}                                                                          \
	return 0;

template <class T> static size_t cells4(const Grid4d<T> &g)
{
	return (size_t)g.getSizeX() * g.getSizeY() * g.getSizeZ() * g.getSizeT();
}
template <class T> static void put(Grid4d<T> &g, const void *src)
{
	memcpy(&g[0], src, sizeof(T) * cells4(g));
}
template <class T> static void get(Grid4d<T> &g, void *dst)
{
	memcpy(dst, &g[0], sizeof(T) * cells4(g));
}
template <class T> static size_t cells3(const Grid<T> &g)
{
	return (size_t)g.getSizeX() * g.getSizeY() * g.getSizeZ();
}
template <class T> static void put3(Grid<T> &g, const void *src)
{
	memcpy(&g[0], src, sizeof(T) * cells3(g));
}
template <class T> static void get3(Grid<T> &g, void *dst)
{
	memcpy(dst, &g[0], sizeof(T) * cells3(g));
}
static Vec4 v4(const float *p) { return Vec4(p[0], p[1], p[2], p[3]); }

#define SOLVER(d) FluidSolver solver(Vec3i((d)[0], (d)[1], (d)[2]), 3, (d)[3])

extern "C" {

const char *ref_last_error() { return g_err; }

int ref_set_threads(int n)
{
#ifdef _OPENMP
	if (n > 0) omp_set_num_threads(n);
	return omp_get_max_threads();
#else
	return 1;
#endif
}

void ref_set_debug_level(int l) { gDebugLevel = l; }

/* grid4d.cpp:539 interpolateGrid4d / :549 interpolateGrid4dVec (elem 1 or 4) */
int ref_interpolate_grid4d(const int *sd, const float *src, const int *td,
                           float *dst, int elem, const float *offset,
                           const float *scale, const float *size)
{
	REF_TRY
	FluidSolver ss(Vec3i(sd[0], sd[1], sd[2]), 3, sd[3]);
	FluidSolver ts(Vec3i(td[0], td[1], td[2]), 3, td[3]);
	if (elem == 1) {
		Grid4d<Real> s(&ss), t(&ts);
		put(s, src);
		interpolateGrid4d(t, s, v4(offset), v4(scale), v4(size));
		get(t, dst);
	} else {
		Grid4d<Vec4> s(&ss), t(&ts);
		put(s, src);
		interpolateGrid4dVec(t, s, v4(offset), v4(scale), v4(size));
		get(t, dst);
	}
	REF_CATCH
}

/* optflow4d.cpp:40-57 interpolGridTempl (pyramid down/up-sampling) */
int ref_interpol_grid_templ(const int *sd, const float *src, const int *td,
                            float *dst, int elem)
{
	REF_TRY
	FluidSolver ss(Vec3i(sd[0], sd[1], sd[2]), 3, sd[3]);
	FluidSolver ts(Vec3i(td[0], td[1], td[2]), 3, td[3]);
	if (elem == 1) {
		Grid4d<Real> s(&ss), t(&ts);
		put(s, src);
		interpolGridTempl<Grid4d<Real> >(t, s);
		get(t, dst);
	} else {
		Grid4d<Vec4> s(&ss), t(&ts);
		put(s, src);
		interpolGridTempl<Grid4d<Vec4> >(t, s);
		get(t, dst);
	}
	REF_CATCH
}

/* optflow4d.cpp:1292 advect4d (elem 1 = Real payload, 4 = Vec4 payload) */
int ref_advect4d(const int *d, const float *vel, float *grid, int elem,
                 float dtFac, float dt)
{
	REF_TRY
	SOLVER(d);
	solver.mDt = dt;
	Grid4d<Vec4> v(&solver);
	put(v, vel);
	if (elem == 1) {
		Grid4d<Real> g(&solver);
		put(g, grid);
		advect4d(v, &g, dtFac);
		get(g, grid);
	} else {
		Grid4d<Vec4> g(&solver);
		put(g, grid);
		advect4d(v, &g, dtFac);
		get(g, grid);
	}
	REF_CATCH
}

/* optflow4d.cpp:2170 advectCflTemplate<Grid4d<int>,Grid4d<Vec4>,Grid4dBase> */
int ref_advect_cfl4d(const int *d, float cfl, const float *vel, float *grid,
                     int elem, float velFactor)
{
	REF_TRY
	SOLVER(d);
	Grid4d<Vec4> v(&solver);
	Grid4d<int> flags(&solver);
	put(v, vel);
	if (elem == 1) {
		Grid4d<Real> g(&solver);
		put(g, grid);
		advectCflTemplate<Grid4d<int>, Grid4d<Vec4>, Grid4dBase>(
		    cfl, &flags, v, &g, 1, velFactor, 1);
		get(g, grid);
	} else {
		Grid4d<Vec4> g(&solver);
		put(g, grid);
		advectCflTemplate<Grid4d<int>, Grid4d<Vec4>, Grid4dBase>(
		    cfl, &flags, v, &g, 1, velFactor, 1);
		get(g, grid);
	}
	REF_CATCH
}

/* optflow4d.cpp:2110 opticalFlow4d; rhsT may be NULL */
int ref_optical_flow4d(const int *d, float *vel, const float *i0,
                       const float *i1, float *rhsT, float wSmooth,
                       float wEnergy, float postVelBlur, float cgAccuracy,
                       int blurType, float resetBndWidth)
{
	REF_TRY
	SOLVER(d);
	Grid4d<Vec4> v(&solver);
	Grid4d<Real> a(&solver), b(&solver), r(&solver);
	put(v, vel);
	put(a, i0);
	put(b, i1);
	opticalFlow4d(v, a, b, rhsT ? &r : NULL, wSmooth, wEnergy, postVelBlur,
	              cgAccuracy, blurType, resetBndWidth);
	get(v, vel);
	if (rhsT) get(r, rhsT);
	REF_CATCH
}

/* optflow4d.cpp:160 gaussianBlurGeneric<Grid4d<Vec4>> (elem 4) / <Grid4d<Real>> */
int ref_gaussian_blur4d(const int *d, float *a, int elem, float sigma, int iter)
{
	REF_TRY
	SOLVER(d);
	if (elem == 4) {
		Grid4d<Vec4> g(&solver);
		put(g, a);
		gaussianBlurGeneric<Grid4d<Vec4> >(g, sigma, iter);
		get(g, a);
	} else {
		Grid4d<Real> g(&solver);
		put(g, a);
		gaussianBlurGeneric<Grid4d<Real> >(g, sigma, iter);
		get(g, a);
	}
	REF_CATCH
}

/* optflow4d.cpp:713 knProjectCells (one projection pass; dst/marker pre-zeroed) */
int ref_project_cells(const int *d, float *dst, const float *vel,
                      const float *phiOrg, const float *phiTarget,
                      float *marker, float threshPhi, int maxIter)
{
	REF_TRY
	SOLVER(d);
	Grid4d<Vec4> gd(&solver), gv(&solver);
	Grid4d<Real> po(&solver), pt(&solver), mk(&solver);
	put(gv, vel);
	put(po, phiOrg);
	put(pt, phiTarget);
	gd.setConst(Vec4(0.));
	mk.setConst(0);
	knProjectCells<Grid4d<Vec4>, Grid4d<Real>, Vec4>(gd, gv, po, pt, mk,
	                                                 threshPhi, maxIter);
	get(gd, dst);
	get(mk, marker);
	REF_CATCH
}

/* optflow4d.cpp:770-780: `sweeps` x { tmp.copyFrom(dst); knCvExpolBlur4d; swap } */
int ref_cv_expol_blur4d(const int *d, float *a, const float *marker, int sweeps)
{
	REF_TRY
	SOLVER(d);
	Grid4d<Vec4> g(&solver), tmp(&solver);
	Grid4d<Real> mk(&solver);
	put(g, a);
	put(mk, marker);
	for (int i = 0; i < sweeps; ++i) {
		tmp.copyFrom(g);
		knCvExpolBlur4d<Grid4d<Vec4>, Grid4d<Real>, Vec4>(g, tmp, mk);
		g.swap(tmp);
	}
	get(g, a);
	REF_CATCH
}

/* optflow4d.cpp:2121 corrVelsOf4d */
int ref_corr_vels_of4d(const int *d, float *dst, float *vel,
                       const float *phiOrg, const float *phiTarget,
                       float threshPhi, float postVelBlur, float resetBndWidth,
                       int maxIter)
{
	REF_TRY
	SOLVER(d);
	Grid4d<Vec4> gd(&solver), gv(&solver);
	Grid4d<Real> po(&solver), pt(&solver);
	put(gd, dst);
	put(gv, vel);
	put(po, phiOrg);
	put(pt, phiTarget);
	corrVelsOf4d(gd, gv, po, po, pt, threshPhi, 1e10, postVelBlur,
	             resetBndWidth, maxIter);
	get(gd, dst);
	get(gv, vel);
	REF_CATCH
}

/* optflow4d.cpp:2132 calcLsDiff4d */
int ref_calc_ls_diff4d(const int *d, const float *i0, const float *i1,
                       float *out, float correction, int bnd, float *result)
{
	REF_TRY
	SOLVER(d);
	Grid4d<Real> a(&solver), b(&solver), o(&solver);
	put(a, i0);
	put(b, i1);
	*result = calcLsDiff4d(a, b, out ? &o : NULL, correction, bnd);
	if (out) get(o, out);
	REF_CATCH
}

/* optflow4d.cpp:2182 opticalFlowMultiscale4d */
int ref_optical_flow_multiscale4d(const int *d, float *vel, const float *i0,
                                  const float *i1, float wSmooth, float wEnergy,
                                  float postVelBlur, float cgAccuracy,
                                  float cfl, float resetBndWidth, int multiStep,
                                  int projSizeThresh, int minGridSize,
                                  int doFinalProject)
{
	REF_TRY
	SOLVER(d);
	Grid4d<Vec4> v(&solver);
	Grid4d<Real> a(&solver), b(&solver);
	put(v, vel);
	put(a, i0);
	put(b, i1);
	opticalFlowMultiscale4d(v, a, b, NULL, wSmooth, wEnergy, 0, postVelBlur,
	                        cgAccuracy, 1, cfl, 1, 1, resetBndWidth, multiStep,
	                        projSizeThresh, minGridSize, doFinalProject != 0);
	get(v, vel);
	REF_CATCH
}

/* ---- the 3D instantiations: Grid<Real> / Grid<Vec3> on a plain 3D solver ---- */
#define SOLVER3(d) FluidSolver solver(Vec3i((d)[0], (d)[1], (d)[2]), (d)[2] > 1 ? 3 : 2, -1) /* nz == 1: a 2D solver */

/* optflow4d.cpp:1175 opticalFlowMultiscale3d */
int ref_optical_flow_multiscale3d(const int *d3, float *vel3, const float *i0, const float *i1, float wSmooth,
                                  float wEnergy, float postVelBlur, float cgAccuracy, float cfl, float resetBndWidth,
                                  int multiStep, int projSizeThresh, int minGridSize, int doFinalProject)
{
	REF_TRY
	SOLVER3(d3);
	Grid<Vec3> v(&solver);
	Grid<Real> a(&solver), b(&solver);
	put3(v, vel3);
	put3(a, i0);
	put3(b, i1);
	opticalFlowMultiscale3d(v, a, b, NULL, wSmooth, wEnergy, 0, postVelBlur, cgAccuracy, 1, cfl, 1, 1, resetBndWidth,
	                        multiStep, projSizeThresh, minGridSize, doFinalProject != 0);
	get3(v, vel3);
	REF_CATCH
}

/* optflow4d.cpp:803 corrVelsOf3d */
int ref_corr_vels_of3d(const int *d3, float *dst3, float *vel3, const float *phiOrg, const float *phiTarget,
                       float threshPhi, float postVelBlur, float resetBndWidth, int maxIter)
{
	REF_TRY
	SOLVER3(d3);
	Grid<Vec3> gd(&solver), gv(&solver);
	Grid<Real> po(&solver), pt(&solver);
	put3(gd, dst3);
	put3(gv, vel3);
	put3(po, phiOrg);
	put3(pt, phiTarget);
	corrVelsOf3d(gd, gv, po, po, pt, threshPhi, 1e10, postVelBlur, resetBndWidth, maxIter);
	get3(gd, dst3);
	get3(gv, vel3);
	REF_CATCH
}

/* optflow4d.cpp:863 advectSemiLagrangeCfl (elem 1: Grid<Real>, 3: Grid<Vec3> payload) */
int ref_advect_semi_lagrange_cfl3d(const int *d3, float cfl, const float *vel3, float *grid, int elem, float velFactor)
{
	REF_TRY
	SOLVER3(d3);
	FlagGrid flags(&solver);
	Grid<Vec3> v(&solver);
	put3(v, vel3);
	if (elem == 3) {
		Grid<Vec3> g(&solver);
		put3(g, grid);
		advectSemiLagrangeCfl(cfl, flags, v, &g, 1, velFactor, 1);
		get3(g, grid);
	} else {
		Grid<Real> g(&solver);
		put3(g, grid);
		advectSemiLagrangeCfl(cfl, flags, v, &g, 1, velFactor, 1);
		get3(g, grid);
	}
	REF_CATCH
}

/* optflow4d.cpp:928 calcLsDiff3d */
int ref_calc_ls_diff3d(const int *d3, const float *i0, const float *i1, float *out, float correction, int bnd,
                       float *result)
{
	REF_TRY
	SOLVER3(d3);
	Grid<Real> a(&solver), b(&solver), o(&solver);
	put3(a, i0);
	put3(b, i1);
	*result = calcLsDiff3d(a, b, out ? &o : NULL, correction, bnd);
	if (out) get3(o, out);
	REF_CATCH
}

/* optflow4d.cpp:1361 extrap4dLsSimple */
int ref_extrap4d_ls_simple(const int *d, float *phi, int distance, int inside)
{
	REF_TRY
	SOLVER(d);
	Grid4d<Real> p(&solver);
	put(p, phi);
	extrap4dLsSimple(p, distance, inside != 0);
	get(p, phi);
	REF_CATCH
}

/* Same call sequence as extrap4dLsSimple (optflow4d.cpp:1361-1406) issued with the
 * reference's own kernels, but returning the integer marker grid `tmp` too. */
int ref_extrap4d_ls_simple_marker(const int *d, float *phi, int distance,
                                  int inside, int *marker)
{
	REF_TRY
	SOLVER(d);
	Grid4d<Real> p(&solver);
	put(p, phi);
	Grid4d<int> tmp(&solver);
	tmp.clear();
	Real direction = inside ? -1. : 1.;
	FOR_IJKT_BND(p, 1)
	{
		if (!inside && p(i, j, k, t) < 0.) tmp(i, j, k, t) = 1;
		if (inside && p(i, j, k, t) > 0.) tmp(i, j, k, t) = 1;
	}
	FOR_IJKT_BND(p, 1)
	{
		Vec4i q(i, j, k, t);
		if (tmp(q)) continue;
		for (int n = 0; n < 8; ++n)
			if (tmp(q + nbs4d[n]) == 1) {
				tmp(i, j, k, t) = 2;
				n = 8;
			}
	}
	for (int dd = 2; dd < 1 + distance; ++dd)
		knExtrap4dLsSimple<Real>(p, distance, tmp, dd, direction);
	knSetRemaining4d<Real>(p, tmp, Real(direction * (distance + 2)));
	get(p, phi);
	get(tmp, marker);
	REF_CATCH
}

/* optflow4d.cpp:1408 extrapolateVec4Simple */
int ref_extrapolate_vec4_simple(const int *d, float *vel, const float *phi,
                                int distance)
{
	REF_TRY
	SOLVER(d);
	Grid4d<Vec4> v(&solver);
	Grid4d<Real> p(&solver);
	put(v, vel);
	put(p, phi);
	extrapolateVec4Simple(v, p, distance);
	get(v, vel);
	REF_CATCH
}

/* optflow4d.cpp:1254 repeatFrame4d */
int ref_repeat_frame4d(const int *d, float *phi, float srct, float range,
                       int bnd)
{
	REF_TRY
	SOLVER(d);
	Grid4d<Real> p(&solver);
	put(p, phi);
	repeatFrame4d(p, srct, range, bnd);
	get(p, phi);
	REF_CATCH
}

/* grid4d.cpp:365 Grid4d<T>::setBound (elem 1, 4; elem -1 = int grid) */
int ref_set_bound4d(const int *d, void *data, int elem, const float *value,
                    int w)
{
	REF_TRY
	SOLVER(d);
	if (elem == 1) {
		Grid4d<Real> g(&solver);
		put(g, data);
		g.setBound(value[0], w);
		get(g, data);
	} else if (elem == 4) {
		Grid4d<Vec4> g(&solver);
		put(g, data);
		g.setBound(v4(value), w);
		get(g, data);
	} else {
		Grid4d<int> g(&solver);
		put(g, data);
		g.setBound((int)value[0], w);
		get(g, data);
	}
	REF_CATCH
}

/* grid4d.cpp:409 Grid4d<T>::setBoundNeumann */
int ref_set_bound_neumann4d(const int *d, void *data, int elem, int w)
{
	REF_TRY
	SOLVER(d);
	if (elem == 1) {
		Grid4d<Real> g(&solver);
		put(g, data);
		g.setBoundNeumann(w);
		get(g, data);
	} else {
		Grid4d<Vec4> g(&solver);
		put(g, data);
		g.setBoundNeumann(w);
		get(g, data);
	}
	REF_CATCH
}

/* grid4d.cpp:266-285 getMin/getMax/getMaxAbs: out[0..2] */
int ref_min_max4d(const int *d, const void *data, int elem, float *out)
{
	REF_TRY
	SOLVER(d);
	if (elem == 1) {
		Grid4d<Real> g(&solver);
		put(g, data);
		out[0] = g.getMin();
		out[1] = g.getMax();
		out[2] = g.getMaxAbs();
	} else {
		Grid4d<Vec4> g(&solver);
		put(g, data);
		out[0] = g.getMin();
		out[1] = g.getMax();
		out[2] = g.getMaxAbs();
	}
	REF_CATCH
}

/* grid4d.h:338-372 element-wise ops on Grid4d<Real>/<Vec4>.
 * op: 0 add, 1 sub, 2 mult, 3 addScaled(factor), 4 multConst(factor), 5 addConst,
 *     6 clamp(factor[0],factor[1]) */
int ref_grid_op4d(const int *d, void *a, const void *b, int elem, int op,
                  const float *factor)
{
	REF_TRY
	SOLVER(d);
	if (elem == 1) {
		Grid4d<Real> ga(&solver), gb(&solver);
		put(ga, a);
		if (b) put(gb, b);
		switch (op) {
		case 0: ga.add(gb); break;
		case 1: ga.sub(gb); break;
		case 2: ga.mult(gb); break;
		case 3: ga.addScaled(gb, factor[0]); break;
		case 4: ga.multConst(factor[0]); break;
		case 5: ga.addConst(factor[0]); break;
		case 6: ga.clamp(factor[0], factor[1]); break;
		}
		get(ga, a);
	} else {
		Grid4d<Vec4> ga(&solver), gb(&solver);
		put(ga, a);
		if (b) put(gb, b);
		switch (op) {
		case 0: ga.add(gb); break;
		case 1: ga.sub(gb); break;
		case 2: ga.mult(gb); break;
		case 3: ga.addScaled(gb, v4(factor)); break;
		case 4: ga.multConst(v4(factor)); break;
		case 5: ga.addConst(v4(factor)); break;
		case 6: ga.clamp(factor[0], factor[1]); break;
		}
		get(ga, a);
	}
	REF_CATCH
}

/* test.cpp:127 simpleBlurSpecial on a 3D Grid<Real> */
int ref_simple_blur_special(const int *d3, float *a, int iter, float thresh,
                            int bord)
{
	REF_TRY
	FluidSolver solver(Vec3i(d3[0], d3[1], d3[2]), 3, 1);
	Grid<Real> g(&solver);
	put3(g, a);
	simpleBlurSpecial(g, iter, thresh, bord);
	get3(g, a);
	REF_CATCH
}

/* grid.cpp:470 Grid<Real>::setBound and levelset.cpp:118 LevelsetGrid::join */
int ref_grid3_set_bound(const int *d3, float *a, float value, int w)
{
	REF_TRY
	FluidSolver solver(Vec3i(d3[0], d3[1], d3[2]), 3, 1);
	Grid<Real> g(&solver);
	put3(g, a);
	g.setBound(value, w);
	get3(g, a);
	REF_CATCH
}
int ref_levelset_join(const int *d3, float *a, const float *b)
{
	REF_TRY
	FluidSolver solver(Vec3i(d3[0], d3[1], d3[2]), 3, 1);
	LevelsetGrid ga(&solver), gb(&solver);
	put3(ga, a);
	put3(gb, b);
	ga.join(gb);
	get3(ga, a);
	REF_CATCH
}

/* test.cpp:231 initTestCheckerboard */
int ref_init_test_checkerboard(const int *d, float *val, float *vec, int brd)
{
	REF_TRY
	SOLVER(d);
	Grid4d<Real> g(&solver);
	Grid4d<Vec4> v(&solver);
	initTestCheckerboard(g, vec ? &v : NULL, brd);
	get(g, val);
	if (vec) get(v, vec);
	REF_CATCH
}

/* fileio.cpp:834 writeGrid4dUni / :885 readGrid4dUni through Grid4d<T>::save/load */
int ref_grid4d_save(const int *d, const void *data, int elem, const char *name)
{
	REF_TRY
	SOLVER(d);
	if (elem == 1) {
		Grid4d<Real> g(&solver);
		put(g, data);
		g.save(name);
	} else {
		Grid4d<Vec4> g(&solver);
		put(g, data);
		g.save(name);
	}
	REF_CATCH
}
int ref_grid4d_load(const int *d, void *data, int elem, const char *name)
{
	REF_TRY
	SOLVER(d);
	if (elem == 1) {
		Grid4d<Real> g(&solver);
		g.load(name);
		get(g, data);
	} else {
		Grid4d<Vec4> g(&solver);
		g.load(name);
		get(g, data);
	}
	REF_CATCH
}
/* fileio.cpp writeGridUni / readGridUni through Grid<Real>::save/load (3D, MNT2) */
int ref_grid3_save(const int *d3, const float *data, const char *name)
{
	REF_TRY
	FluidSolver solver(Vec3i(d3[0], d3[1], d3[2]), 3, 1);
	Grid<Real> g(&solver);
	put3(g, data);
	g.save(name);
	REF_CATCH
}
int ref_grid3_load(const int *d3, float *data, const char *name)
{
	REF_TRY
	FluidSolver solver(Vec3i(d3[0], d3[1], d3[2]), 3, 1);
	Grid<Real> g(&solver);
	g.load(name);
	get3(g, data);
	REF_CATCH
}

/* optflow4d.cpp:1464 loadPlaceGrid4d (reads 3D .uni slices named by printf pattern) */
int ref_load_place_grid4d(const char *fname, const int *d, float *phi,
                          const float *offset, const float *scale,
                          int fileIdxStart, int fileIdxEnd, int debugSkipLoad,
                          float spread, const float *overrideSize,
                          float overrideTimeOff, int overrideGoodRegion,
                          float loadTimeScale, int rescaleSdfValues,
                          float sdfIsoOff, float repeatStartFrame)
{
	REF_TRY
	SOLVER(d);
	Grid4d<Real> p(&solver);
	put(p, phi);
	loadPlaceGrid4d(fname, p, v4(offset), v4(scale), fileIdxStart, fileIdxEnd,
	                debugSkipLoad, spread, v4(overrideSize), overrideTimeOff,
	                overrideGoodRegion, loadTimeScale, rescaleSdfValues != 0,
	                sdfIsoOff, repeatStartFrame);
	get(p, phi);
	REF_CATCH
}

/* optflow4d.cpp:1598 shiftForwGrid4d */
int ref_shift_forw_grid4d(const int *d, float *phi, int overrideGoodRegion)
{
	REF_TRY
	SOLVER(d);
	Grid4d<Real> p(&solver);
	put(p, phi);
	shiftForwGrid4d(p, overrideGoodRegion);
	get(p, phi);
	REF_CATCH
}

/* optflow4d.cpp:1871/1951/1930: _OptInit (useDefoVols=false) + n x _OptRun + _Finish.
 * times[n]; dst receives n consecutive 3D frames (d3 cells each). */
int ref_load_advect_time_slices_opt(const char *fname, const int *d3,
                                    float *dst, const int *d, const float *phi,
                                    int n, const float *times, float blendAlpha,
                                    float loadTimeScale,
                                    const float *defoOffset,
                                    const float *defoScale,
                                    const float *defoFactor,
                                    const float *overrideSize,
                                    float overrideTimeOff, int bordSkip,
                                    float defoAniFac)
{
	REF_TRY
	SOLVER(d);
	Grid4d<Real> p(&solver);
	put(p, phi);
	FluidSolver s3(Vec3i(d3[0], d3[1], d3[2]), 3, 1);
	Grid<Real> out(&s3);
	const int ID = 4711;
	loadAdvectTimeSlice_OptInit(ID, fname, false, false, 0.2);
	for (int f = 0; f < n; ++f) {
		out.clear();
		loadAdvectTimeSlice_OptRun(ID, fname, out, p, times[f], blendAlpha,
		                           loadTimeScale, v4(defoOffset),
		                           v4(defoScale), v4(defoFactor),
		                           v4(overrideSize), overrideTimeOff, NULL,
		                           NULL, false, 0., bordSkip, 0., defoAniFac);
		get3(out, dst + (size_t)f * cells3(out));
	}
	loadAdvectTimeSlice_Finish(ID);
	REF_CATCH
}

/* optflow4d.cpp:1871 _OptInit(useDefoVols = true) + :1914 _OptAdd per further file + n x :1951 _OptRun + :1930 _Finish:
 * 2 or 3 deformation volumes composed per frame (flof.py `thirdload`). */
int ref_load_advect_defovols(const char *const *fnames, int numDv, int doAligned,
                             float partialLoadFac, const int *d3, float *dst,
                             const int *d, const float *phi, int n,
                             const float *times, float blendAlpha,
                             float thirdAlpha, float fourthAlpha,
                             float loadTimeScale, const float *defoOffset,
                             const float *defoScale, const float *defoFactor,
                             const float *overrideSize, float overrideTimeOff,
                             int bordSkip, float defoAniFac)
{
	REF_TRY
	SOLVER(d);
	Grid4d<Real> p(&solver);
	put(p, phi);
	FluidSolver s3(Vec3i(d3[0], d3[1], d3[2]), 3, 1);
	Grid<Real> out(&s3);
	const int ID = 4712;
	loadAdvectTimeSlice_OptInit(ID, fnames[0], true, doAligned != 0, partialLoadFac);
	for (int dv = 1; dv < numDv; ++dv) loadAdvectTimeSlice_OptAdd(ID, fnames[dv]);
	for (int f = 0; f < n; ++f) {
		out.clear();
		loadAdvectTimeSlice_OptRun(ID, fnames[0], out, p, times[f], blendAlpha,
		                           loadTimeScale, v4(defoOffset),
		                           v4(defoScale), v4(defoFactor),
		                           v4(overrideSize), overrideTimeOff, NULL,
		                           NULL, false, thirdAlpha, bordSkip, fourthAlpha, defoAniFac);
		get3(out, dst + (size_t)f * cells3(out));
	}
	loadAdvectTimeSlice_Finish(ID);
	REF_CATCH
}

/* optflow4d.cpp:1671 loadAdvectTimeSlice (unoptimised twin, one frame) */
int ref_load_advect_time_slice_debug(const char *fname, const int *d3, float *dst, float *dbgVel, float *dbgVelT,
                                     const int *d, const float *phi, float time, float blendAlpha, float loadTimeScale,
                                     const float *defoOffset, const float *defoScale, const float *defoFactor,
                                     const float *overrideSize, float overrideTimeOff, float defoAniFac, int zeroVel)
{
	REF_TRY
	SOLVER(d);
	Grid4d<Real> p(&solver);
	put(p, phi);
	FluidSolver s3(Vec3i(d3[0], d3[1], d3[2]), 3, 1);
	Grid<Real> out(&s3), velT(&s3);
	Grid<Vec3> vel(&s3);
	loadAdvectTimeSlice(0, fname, out, p, time, blendAlpha, loadTimeScale, v4(defoOffset), v4(defoScale), v4(defoFactor),
	                    v4(overrideSize), overrideTimeOff, &vel, &velT, zeroVel != 0, 0., 1, 0., defoAniFac);
	get3(out, dst);
	get3(velT, dbgVelT);
	for (size_t c = 0; c < cells3(out); ++c) {
		dbgVel[3 * c] = vel[c].x;
		dbgVel[3 * c + 1] = vel[c].y;
		dbgVel[3 * c + 2] = vel[c].z;
	}
	REF_CATCH
}

int ref_load_advect_time_slice(const char *fname, const int *d3, float *dst,
                               const int *d, const float *phi, float time,
                               float blendAlpha, float loadTimeScale,
                               const float *defoOffset, const float *defoScale,
                               const float *defoFactor,
                               const float *overrideSize, float overrideTimeOff)
{
	REF_TRY
	SOLVER(d);
	Grid4d<Real> p(&solver);
	put(p, phi);
	FluidSolver s3(Vec3i(d3[0], d3[1], d3[2]), 3, 1);
	Grid<Real> out(&s3);
	loadAdvectTimeSlice(0, fname, out, p, time, blendAlpha, loadTimeScale,
	                    v4(defoOffset), v4(defoScale), v4(defoFactor),
	                    v4(overrideSize), overrideTimeOff);
	get3(out, dst);
	REF_CATCH
}

/* grid4d.cpp:488 getSliceFrom4d */
int ref_get_slice_from4d(const int *d, const float *src, int srct, float *dst3)
{
	REF_TRY
	SOLVER(d);
	Grid4d<Real> p(&solver);
	put(p, src);
	FluidSolver s3(Vec3i(d[0], d[1], d[2]), 3, 1);
	Grid<Real> out(&s3);
	getSliceFrom4d(p, srct, out);
	get3(out, dst3);
	REF_CATCH
}

} /* extern "C" */
