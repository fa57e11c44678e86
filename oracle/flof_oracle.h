/* oracle/flof_oracle.h -- TEST INFRASTRUCTURE (CPU restatement of the reference).
 *
 * Plain-C restatement of the FlOF 4D optical-flow hot path of thunil/ofblend
 * (source/plugin/optflow4d.cpp, source/grid4d.{h,cpp}, source/util/vector4d.h,
 * source/util/interpol.h, source/test.cpp).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may use it, and only as the checker.
 *
 * PARITY PIN: every function here is checked bit-for-bit (or to the stated tolerance)
 * against the reference itself (oracle/_ref/libofref.so, built from /root/reference by
 * oracle/Makefile) in tests/test_oracle_vs_ref.py, and against the committed fixtures under
 * tests/golden/ that were generated from the reference by tests/golden/make_golden.py.
 *
 * Storage: x fastest, idx = i + nx*(j + ny*(k + nz*t)) (grid4d.h:92-97); Vec4 grids are
 * arrays of 4 floats per cell (vector4d.h:164-181).  All arithmetic is fp32 with the
 * reference's operation order and its float/double promotions; compile with
 * -ffp-contract=off (the reference is built for baseline x86-64, i.e. without FMA).
 */
#ifndef FLOF_ORACLE_H
#define FLOF_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int nx, ny, nz, nt; } orc_dim4;

int orc_set_threads(int n);
/* CG dot-product summation order: 0 sequential (the reference), 1 blocked (conditioning probe: how far the
 * reference's own result moves when only the order of the fp64 sums changes) */
void orc_set_dot_mode(int m);
/* the reference's sequential dot-product loops on their own (checker of flof_dot_seq) */
double orc_dot_seq(const float *a, const float *b, long long cells, int kind, float diag);

/* vector4d.h:488 interpol4d / interpol.h:101 interpol (elem = 1 or 4 floats per cell) */
void orc_interpol4d(const float *data, orc_dim4 d, int elem, const float pos[4], float *out);
void orc_interpol3d(const float *data, int nx, int ny, int nz, int elem, const float pos[3],
                    float *out);

/* grid4d.cpp:559 gridFactor4d: srcFac out, off in/out */
void orc_grid_factor4d(const float s1[4], const float s2[4], const float optSize[4],
                       const float scale[4], float srcFac[4], float off[4]);
/* grid4d.cpp:531 knInterpol4d */
void orc_kn_interpol4d(float *dst, orc_dim4 td, const float *src, orc_dim4 sd, int elem,
                       const float srcFac[4], const float off[4]);
/* grid4d.cpp:539/549 interpolateGrid4d[Vec] */
void orc_interpolate_grid4d(float *dst, orc_dim4 td, const float *src, orc_dim4 sd, int elem,
                            const float offset[4], const float scale[4], const float size[4]);
/* optflow4d.cpp:40-57 interpolGridTempl */
void orc_interpol_grid_templ(float *dst, orc_dim4 td, const float *src, orc_dim4 sd, int elem);

/* grid4d.cpp:355-412, 213-305 */
void orc_set_bound4d(float *a, orc_dim4 d, int elem, const float value[4], int w);
void orc_set_bound4d_int(int *a, orc_dim4 d, int value, int w);
void orc_set_bound_neumann4d(float *a, orc_dim4 d, int elem, int w);
void orc_min_max4d(const float *a, orc_dim4 d, int elem, float out[3]); /* min, max, maxAbs */
/* op: 0 add 1 sub 2 mult 3 addScaled 4 multConst 5 addConst 6 clamp(f[0],f[1]) */
void orc_grid_op4d(float *a, const float *b, orc_dim4 d, int elem, int op, const float f[4]);

/* optflow4d.cpp:1275-1321, 2170-2180 */
void orc_advect4d(const float *vel, float *grid, orc_dim4 d, int elem, float dt);
void orc_advect_cfl4d(float cfl, const float *vel, float *grid, orc_dim4 d, int elem,
                      float velFactor);

/* optflow4d.cpp:128-174 */
void orc_gaussian_blur4d(float *a, orc_dim4 d, int elem, float sigma, int iter);

/* optflow4d.cpp:361-553 (DIM = 4).  Returns CG iterations; rhsT may be NULL. */
int orc_optical_flow4d(float *vel, const float *i0, const float *i1, float *rhsT, orc_dim4 d,
                       float wSmooth, float wEnergy, float postVelBlur, float cgAccuracy,
                       float resetBndWidth, float *cgResOut);

/* optflow4d.cpp:651-802 */
void orc_project_cells(float *dst, const float *vel, const float *phiOrg,
                       const float *phiTarget, float *marker, orc_dim4 d, float threshPhi,
                       int maxIter);
void orc_cv_expol_blur4d(float *a, const float *marker, orc_dim4 d, int sweeps);
void orc_corr_vels_of4d(float *dst, float *vel, const float *phiOrg, const float *phiTarget,
                        orc_dim4 d, float threshPhi, float postVelBlur, float resetBndWidth,
                        int maxIter);

/* optflow4d.cpp:895-927 */
float orc_calc_ls_diff4d(const float *i0, const float *i1, float *out, orc_dim4 d,
                         float correction, int bnd);

/* optflow4d.cpp:936-1173.  trace (optional, >= 64 ints/floats): CG iterations per solve and
 * error values in call order; ntrace returns the counts. */
float orc_optical_flow_multiscale4d(float *vel, const float *i0, const float *i1, orc_dim4 d,
                                    float wSmooth, float wEnergy, float postVelBlur,
                                    float cgAccuracy, float cfl, float resetBndWidth,
                                    int multiStep, int projSizeThresh, int minGridSize,
                                    int doFinalProject, int *cgIters, int *nIters,
                                    float *errs, int *nErrs);

/* optflow4d.cpp:1254-1266, 1323-1439 */
void orc_repeat_frame4d(float *phi, orc_dim4 d, float srct, float range, int bnd);
void orc_extrap4d_ls_simple(float *phi, orc_dim4 d, int distance, int inside, int *marker);
void orc_extrapolate_vec4_simple(float *vel, const float *phi, orc_dim4 d, int distance);

/* test.cpp:93-131, grid.cpp:462-473, levelset.cpp:114-118 (3D output ops) */
void orc_simple_blur_special(float *a, int nx, int ny, int nz, int iter, float thresh,
                             int bord);
void orc_grid3_set_bound(float *a, int nx, int ny, int nz, float value, int w);
void orc_levelset_join(float *a, const float *b, long n);

/* optflow4d.cpp:1464-1595 loadPlaceGrid4d with the 3D slices already in memory:
 * slices[(fileidClamp - fileIdxStart)] is an sx*sy*sz float volume. */
void orc_load_place_grid4d(const float *slices, int sx, int sy, int sz, float *phi, orc_dim4 d,
                           const float offset[4], const float scale[4], int fileIdxStart,
                           int fileIdxEnd, int debugSkipLoad, float spread,
                           const float overrideSize[4], float overrideTimeOff,
                           int overrideGoodRegion, float loadTimeScale, int rescaleSdfValues,
                           float sdfIsoOff, float repeatStartFrame);
void orc_shift_forw_grid4d(float *phi, orc_dim4 d, int overrideGoodRegion);

/* optflow4d.cpp:1951-2105 loadAdvectTimeSlice_OptRun (useDefoVols = false) with the
 * deformation volume `defo` (Vec4, dims dd) already in memory; one output frame. */
void orc_load_advect_time_slice(const float *defo, orc_dim4 dd, float *dst, int nx, int ny,
                                int nz, const float *phi, orc_dim4 d, float time,
                                float blendAlpha, float loadTimeScale,
                                const float defoOffset[4], const float defoScale[4],
                                const float defoFactor[4], const float overrideSize[4],
                                float overrideTimeOff, int bordSkip, float defoAniFac);

/* optflow4d.cpp:1822-1863, 1951-2105 with useDefoVols: n frames through 2 / 3 deformation volumes */
void orc_load_advect_defovols(const float *const *vols, int numDv, orc_dim4 dd, int doAligned, float partialLoadFac,
                              float *dst, int nx, int ny, int nz, const float *phi, orc_dim4 d, int n,
                              const float *times, float blendAlpha, float thirdAlpha, float fourthAlpha,
                              float loadTimeScale, const float defoOffset[4], const float defoScale[4],
                              const float defoFactor[4], const float overrideSize[4], float overrideTimeOff,
                              int bordSkip, float defoAniFac);
/* optflow4d.cpp:1671-1760: the unoptimised loadAdvectTimeSlice incl. its debugVel / debugVelT outputs */
void orc_load_advect_time_slice_unopt(const float *defo, orc_dim4 dd, float *dst, float *dbgVel, float *dbgVelT, int nx,
                                      int ny, int nz, const float *phi, orc_dim4 d, float time, float blendAlpha,
                                      float loadTimeScale, const float defoOffset[4], const float defoScale[4],
                                      const float defoFactor[4], const float overrideSize[4], float overrideTimeOff,
                                      float defoAniFac, int zeroVel);
/* ---- oracle/flof_oracle3.c: the 3D / 2D instantiations; grids nx*ny*nz, velocities 3 floats per cell ---- */
/* optflow4d.cpp:863-872 advectSemiLagrangeCfl (elem 1: Grid<Real>, 3: Grid<Vec3> payload) */
void orc3_advect_cfl(float cfl, const float *vel3, float *grid, int elem, int nx, int ny, int nz, float velFactor);
/* optflow4d.cpp:928-933 calcLsDiff3d */
float orc3_calc_ls_diff(const float *i0, const float *i1, float *out, int nx, int ny, int nz, float correction, int bnd);
/* optflow4d.cpp:803-812 corrVelsOf3d */
void orc3_corr_vels(float *dst3, float *vel3, const float *phiOrg, const float *phiTarget, int nx, int ny, int nz,
                    float threshPhi, float postVelBlur, float resetBndWidth, int maxIter);
/* optflow4d.cpp:1175-1188 opticalFlowMultiscale3d; cgIters / errs (optional, >= 64 entries) receive the trace */
float orc3_optical_flow_multiscale(float *vel3, const float *i0, const float *i1, int nx, int ny, int nz, float wSmooth,
                                   float wEnergy, float postVelBlur, float cgAccuracy, float cfl, float resetBndWidth,
                                   int multiStep, int projSizeThresh, int minGridSize, int doFinalProject, int *cgIters,
                                   int *nIters, float *errs, int *nErrs);
#ifdef __cplusplus
}
#endif
#endif
