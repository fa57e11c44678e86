/* flof_b200.h -- C ABI of the B200-native FlOF 4D optical-flow hot path.
 *
 * This is the drop-in boundary below mantaflow's host framework (Grid4d / FluidSolver /
 * pwrapper): every entry point replaces one PYTHON()-registered plugin function or Grid4d
 * method of the reference (thunil/ofblend) and cites it as `ref: file:line` relative to
 * /root/reference/source.  Plain C: raw DEVICE pointers, PODs and scalars; no C++ types, no
 * exceptions, no torch types.  Implemented by ofblend_b200/libflof_b200.so (hand-written
 * sm_100a CUDA, ofblend_b200/csrc/).  There is NO CPU fallback: without a CUDA device
 * flof_ctx_create() fails and nothing else can be called.
 *
 * Conventions
 *  - Grids are dense, x fastest: idx = i + nx*(j + ny*(k + nz*t)) (ref: grid4d.h:92-97).
 *    `elem` is the number of floats per cell: 1 = Grid4d<Real>, 4 = Grid4d<Vec4> (AoS float4,
 *    ref: util/vector4d.h:164-181; must be 16-byte aligned).  Grid4d<int> has its own calls.
 *  - All work is enqueued on the context's CUDA stream and is asynchronous unless the call
 *    returns a scalar to the host (documented per call) or is a *_host convenience call.
 *  - Every call returns FLOF_OK (0) or an error code; flof_last_error(ctx) gives the text.
 *    The host layer turns a non-zero code into the reference's Manta::Error -> RuntimeError
 *    (ref: general.h:87-99, pwrapper/pclass.cpp:50-54).
 *  - Arithmetic is fp32 with the reference's operation order and no FMA contraction;
 *    reductions that the reference accumulates in double are fp64 here too.
 */
#ifndef FLOF_B200_H
#define FLOF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLOF_OK 0
#define FLOF_ERR_CUDA 1  /* a CUDA runtime call or kernel failed */
#define FLOF_ERR_ARG 2   /* invalid argument (the reference's assertMsg/errMsg cases) */
#define FLOF_ERR_NOMEM 3 /* device allocation failed */

typedef struct flof_ctx flof_ctx; /* opaque: device, stream, grid memory pool, scratch */

typedef struct { int nx, ny, nz, nt; } flof_dim4; /* ref: Grid4dBase::mSize */
typedef struct { int nx, ny, nz; } flof_dim3;     /* ref: GridBase::mSize */

/* ---- context, device memory (ref: FluidSolver grid pool, fluidsolver.cpp:24-53, 94-126) ---- */
int flof_ctx_create(flof_ctx **out, int device);
int flof_ctx_destroy(flof_ctx *ctx);
const char *flof_last_error(flof_ctx *ctx); /* ctx may be NULL: last creation error */
void *flof_ctx_stream(flof_ctx *ctx);       /* the cudaStream_t all work is enqueued on */
int flof_device_count(void);
int flof_ctx_sm_count(flof_ctx *ctx);        /* SMs of the context's device (grid sizing, tests) */
/* number of kernels launched through this context since creation (bench.py gpu_launches) */
long long flof_ctx_launch_count(flof_ctx *ctx);

int flof_malloc(flof_ctx *ctx, void **dptr, size_t bytes); /* stream-ordered pool alloc, zeroed
                                                              (Grid4d ctor clears, grid4d.cpp:82) */
int flof_free(flof_ctx *ctx, void *dptr);
int flof_memcpy_h2d(flof_ctx *ctx, void *dst, const void *src, size_t bytes);
int flof_memcpy_d2h(flof_ctx *ctx, void *dst, const void *src, size_t bytes); /* synchronises */
int flof_memcpy_d2d(flof_ctx *ctx, void *dst, const void *src, size_t bytes);
int flof_memset0(flof_ctx *ctx, void *dst, size_t bytes); /* ref: Grid4d::clear grid4d.cpp:108 */
int flof_sync(flof_ctx *ctx);
/* pinned (page-locked) host buffers for the host<->device copies of the *_host entry points */
int flof_host_alloc(flof_ctx *ctx, void **hptr, size_t bytes);
int flof_host_free(flof_ctx *ctx, void *hptr);

/* Per-kernel device timing with CUDA events on the context stream (no profiler involved):
 * begin() starts bracketing every kernel launch, end() synchronises and returns one record per
 * kernel name, sorted by total time.  Used by bench.py for the live roofline numbers. */
typedef struct {
	char name[96];
	int64_t cells;   /* cells of the pyramid level the launches worked on (0 = outside the driver) */
	int launches;
	float total_ms;
} flof_kernel_stat; /* one record per (kernel, level) */
int flof_profile_begin(flof_ctx *ctx);
int flof_profile_end(flof_ctx *ctx, flof_kernel_stat *out, int max_out, int *n_out);

/* ---- multi-GPU (one process + one context per GPU, NCCL over NVLink; the reference is single process) ----
 * t-sharding: inside flof_optical_flow_multiscale4d every pyramid level whose T is divisible
 * by the rank count (and large enough) is cut along t; ranks own T/P contiguous slices, exchange ghost
 * slices per stencil sweep, all-reduce the CG scalars and all-gather gather sources.  Inputs must be
 * identical (replicated) on every rank at call time; the returned deformation is complete on every rank. */
#define FLOF_COMM_ID_BYTES 128
int flof_comm_unique_id(char out[FLOF_COMM_ID_BYTES]);  /* rank 0: ncclGetUniqueId, distribute out of band */
int flof_ctx_comm_init(flof_ctx *ctx, int nranks, int rank, const char id[FLOF_COMM_ID_BYTES]);
int flof_ctx_comm_destroy(flof_ctx *ctx);
int flof_ctx_rank(flof_ctx *ctx);
int flof_ctx_nranks(flof_ctx *ctx);
/* kernel selection knobs, all bit-identical (A/B timing, tests): "expol_mode" 1 Vec4 work list of 4y items / 3, 4 the same with
 * 4y x 2z, 4y x 4z items / 0 component planes / 2 dense; "expol_variant", "apply_variant" register-budget variants (apply: 11 = default).  Defaults come from the environment
 * variables FLOF_EXPOL_MODE, FLOF_EXPOL_VARIANT, FLOF_APPLY_VARIANT when the context is created. */
int flof_ctx_set_option(flof_ctx *ctx, const char *name, int value);
int flof_ctx_get_option(flof_ctx *ctx, const char *name, int *value);
/* "dot_mode": 1 (default, FLOF_DOT_MODE) = the CG's dot products are evaluated in the reference's SEQUENTIAL summation
 * order, bit for bit (ref: dotProd optflow4d.cpp:234-241 -- `for (i) d += a[i]*b[i]`, fp32 product, fp64 running sum),
 * which makes the whole mode-1 result bit-identical to the reference; 0 = tree reductions (last bits of the sums differ).
 * flof_dot_seq is that dot product on its own (tests, tools): kind 0 = sum a[i]*b[i], kind 1 = sum (a[i]*precond(b)[i])*a[i]
 * with the Jacobi reciprocal diagonal of grad = b (ref: precondInit/precondApply :331-354); `cells` Vec4 cells.
 * stats[15] = dot products, dirty leaves, raw products, pieces, fallbacks, failed consistency checks, segments that took
 * the careful (leaf-by-leaf) path, fallbacks that had to return the tree sum, OR of the reason flags of the fallbacks,
 * leaves kept as plain products (summed one by one), then five profiling sums of the resolver (cycles of its gather /
 * compose / walk / finish phases, walk steps) -- all since the context was created. */
int flof_dot_seq(flof_ctx *ctx, const float *a, const float *b, int64_t cells, int kind, float diag, double *result,
                 unsigned long long *stats);
int flof_seq_stats(flof_ctx *ctx, unsigned long long *stats);
/* levels with fewer cells are computed redundantly on every rank instead of being sharded (default 2^22) */
int flof_ctx_set_shard_min_cells(flof_ctx *ctx, int64_t cells);
void flof_slab_range(int nt, int nranks, int rank, int *ta, int *tb); /* slices owned by `rank` */
int flof_comm_barrier(flof_ctx *ctx);                                 /* stream sync + all ranks */
/* halos and CG scalars go through NVLink peer mailboxes (IPC-mapped, flof_comm.cu) when every rank could map them;
 * *enabled tells which path is active, a non-zero *timed_out (also returned as an error) that a peer never arrived */
int flof_comm_p2p_status(flof_ctx *ctx, int *enabled, int *timed_out);
int flof_comm_allreduce_max_host(flof_ctx *ctx, double *v);           /* max over ranks of a host scalar */

/* ---- element-wise Grid4d<T> ops (ref: grid4d.h:338-382, grid4d.cpp:213-264) -------------- */
#define FLOF_OP_ADD 0  /* a += b              ref: Grid4d::add        grid4d.cpp:230 */
#define FLOF_OP_SUB 1  /* a -= b              ref: Grid4d::sub        grid4d.cpp:234 */
#define FLOF_OP_MULT 2 /* a *= b              ref: Grid4d::mult       grid4d.cpp:256 */
#define FLOF_OP_MIN 3  /* a = min(a,b)        ref: LevelsetGrid::join levelset.cpp:114-118 */
int flof_grid_binary(flof_ctx *ctx, float *a, const float *b, int64_t cells, int elem, int op);
/* a += f*b (f per component)               ref: Grid4d::addScaled   grid4d.cpp:239 */
int flof_grid_add_scaled(flof_ctx *ctx, float *a, const float *b, int64_t cells, int elem,
                         const float f[4]);
int flof_grid_mult_const(flof_ctx *ctx, float *a, int64_t cells, int elem, const float f[4]);
int flof_grid_add_const(flof_ctx *ctx, float *a, int64_t cells, int elem, const float f[4]);
int flof_grid_set_const(flof_ctx *ctx, float *a, int64_t cells, int elem, const float f[4]);
int flof_grid_clamp(flof_ctx *ctx, float *a, int64_t cells, int elem, float lo, float hi);
int flof_grid_set_const_int(flof_ctx *ctx, int *a, int64_t cells, int v);
/* out = {min, max, maxAbs}; Vec4: sqrt of min/max squared norm (ref: grid4d.cpp:143-191,
 * 266-285).  Synchronises. */
int flof_grid_min_max(flof_ctx *ctx, const float *a, int64_t cells, int elem, float out[3]);
int flof_grid_min_max_int(flof_ctx *ctx, const int *a, int64_t cells, int out[2]);
/* max |a-b| summed over components (ref: grid4dMaxDiff* grid4d.cpp:419-464).  Synchronises. */
int flof_grid_max_diff(flof_ctx *ctx, const float *a, const float *b, int64_t cells, int elem,
                       double *out);

/* ---- boundaries (ref: grid4d.cpp:355-412; grid.cpp:462-473) ------------------------------ */
/* writes w+1 shells (`i <= w`), exactly like the reference */
int flof_grid4d_set_bound(flof_ctx *ctx, float *a, flof_dim4 d, int elem, const float v[4], int w);
int flof_grid4d_set_bound_int(flof_ctx *ctx, int *a, flof_dim4 d, int v, int w);
int flof_grid4d_set_bound_neumann(flof_ctx *ctx, float *a, flof_dim4 d, int elem, int w);
int flof_grid3_set_bound(flof_ctx *ctx, float *a, flof_dim3 d, float v, int w);

/* ---- slices / components / regions (ref: grid4d.cpp:338-353, 466-524; test.cpp:199-249) --- */
int flof_get_comp4d(flof_ctx *ctx, const float *src_vec4, float *dst, int64_t cells, int c);
int flof_set_comp4d(flof_ctx *ctx, const float *src, float *dst_vec4, int64_t cells, int c);
int flof_set_region4d(flof_ctx *ctx, float *dst, flof_dim4 d, int elem, const float start[4],
                      const float end[4], const float value[4]);
int flof_get_slice_from4d(flof_ctx *ctx, const float *src, flof_dim4 d, int srct, float *dst3);
/* dst_xyz: Vec3 grid (3 floats/cell), dst_t optional Real grid */
int flof_get_slice_from4d_vec(flof_ctx *ctx, const float *src_vec4, flof_dim4 d, int srct,
                              float *dst_xyz, float *dst_t);
int flof_place_grid3d(flof_ctx *ctx, const float *src3, float *dst, flof_dim4 d, int dstt);
int flof_init_vec_from_scalar(flof_ctx *ctx, const float *src, float *dst_vec4, int64_t cells);
int flof_init_test_checkerboard(flof_ctx *ctx, float *val, float *vec4_or_null, flof_dim4 d,
                                int brd);
/* ref: debugGridAvg4d test.cpp:199.  Synchronises. */
int flof_debug_grid_avg4d(flof_ctx *ctx, const float *phi, flof_dim4 d, int brd, float *out);
/* ref: debugVelAvg4d test.cpp:210-219 (mean |v| over the bnd region) */
int flof_debug_vel_avg4d(flof_ctx *ctx, const float *v, flof_dim4 d, int brd, float *out);
/* ref: calcObfDiff optflow4d.cpp:1762-1777 (3D difference images of two blended results; vel grids are Vec3 AoS) */
int flof_calc_obf_diff(flof_ctx *ctx, const float *phi1, const float *phi2, float *phiDiff, const float *vel1,
                       const float *vel2, const float *velt1, const float *velt2, float *velDiff, flof_dim3 d, int bnd);

/* ---- resampling (ref: grid4d.cpp:531-569, grid4d.h:275-283, 463-471; optflow4d.cpp:40-57) - */
/* host-side helper, pure arithmetic: srcFac out, off in/out (ref: gridFactor4d grid4d.cpp:559) */
void flof_grid_factor4d(const float s1[4], const float s2[4], const float optSize[4],
                        const float scale[4], float srcFac[4], float off[4]);
/* target(x) = source.interp(x*srcFac + off)      ref: knInterpol4d grid4d.cpp:531 */
int flof_kn_interpol4d(flof_ctx *ctx, float *dst, flof_dim4 td, const float *src, flof_dim4 sd,
                       int elem, const float srcFac[4], const float off[4]);
/* ref: interpolateGrid4d / interpolateGrid4dVec grid4d.cpp:539-557 */
int flof_interpolate_grid4d(flof_ctx *ctx, float *dst, flof_dim4 td, const float *src,
                            flof_dim4 sd, int elem, const float offset[4], const float scale[4],
                            const float size[4]);
/* pyramid down/up-sampling                       ref: interpolGridTempl optflow4d.cpp:40-57 */
int flof_interpol_grid_templ(flof_ctx *ctx, float *dst, flof_dim4 td, const float *src,
                             flof_dim4 sd, int elem);

/* ---- 4D semi-Lagrangian advection (ref: optflow4d.cpp:1275-1321, 2170-2180) --------------- */
/* dst(x) = src.interp(x + 0.5 - vel(x)*dt) on the interior, 0 on the 1-cell shell
 * (ref: semiLagrange4d :1275 into a fresh grid :1287).  dst must not alias src. */
int flof_semi_lagrange4d(flof_ctx *ctx, const float *vel, const float *src, float *dst,
                         flof_dim4 d, int elem, float dt);
/* in place (ref: advect4d :1292; dt = solver dt * dtFac) */
int flof_advect4d(flof_ctx *ctx, const float *vel, float *grid, flof_dim4 d, int elem, float dt);
/* ref: advectCflTemplate<Grid4d...> :2170 with solver dt 1.  Synchronises (max |vel|). */
int flof_advect_cfl4d(flof_ctx *ctx, float cfl, const float *vel, float *grid, flof_dim4 d,
                      int elem, float velFactor);

/* ---- optical-flow solve (ref: optflow4d.cpp:179-553) -------------------------------------- */
/* matrix-free system: grad (Vec4) and rhs (Vec4) per cell; border cells (bnd 1) get grad = 0,
 * rhs = 0 and are treated as identity rows (ref: assembly :398-493) */
int flof_of_assemble(flof_ctx *ctx, float *grad, float *rhs, const float *i0, const float *i1,
                     const float *vel, flof_dim4 d, float wSmooth, float wEnergy);
/* Jacobi-PCG with the reference's stopping rule (ref: GridCGOptflow4d::solve :274-329,
 * applyMat :211-232, dotProd :234, getMaxNorm :243, addScaled :253, precond :331-354).
 * x receives the solution (zero initial guess).  Synchronises. */
int flof_of_cg(flof_ctx *ctx, float *x, const float *grad, const float *rhs, flof_dim4 d,
               float wSmooth, float wEnergy, float accuracy, int maxIter, int *iters,
               float *relResidual);
/* ref: opticalFlow4d :2110 -> opticalFlowDim<.,.,4> :361-553.  rhsT may be NULL.  Synchronises. */
int flof_optical_flow4d(flof_ctx *ctx, float *vel, const float *i0, const float *i1, float *rhsT,
                        flof_dim4 d, float wSmooth, float wEnergy, float postVelBlur,
                        float cgAccuracy, float resetBndWidth, int *cgIters, float *cgRes);

/* ---- blurs (ref: optflow4d.cpp:128-174, 613-626) ------------------------------------------ */
int flof_gaussian_blur4d(flof_ctx *ctx, float *a, flof_dim4 d, int elem, float sigma, int iter);
/* `sweeps` x { tmp = a; 81-tap mean where marker == 0; swap }   (ref: :770-780) */
int flof_cv_expol_blur4d(flof_ctx *ctx, float *a, const float *marker, flof_dim4 d, int sweeps);

/* ---- SDF projection (ref: optflow4d.cpp:651-802, 2121-2130) ------------------------------- */
int flof_project_cells(flof_ctx *ctx, float *dst, const float *vel, const float *phiOrg,
                       const float *phiTarget, float *marker, flof_dim4 d, float threshPhi,
                       int maxIter);
/* Synchronises (max |vel| decides the number of extrapolation sweeps, :748). */
int flof_corr_vels_of4d(flof_ctx *ctx, float *dst, float *vel, const float *phiOrg,
                        const float *phiTarget, flof_dim4 d, float threshPhi, float postVelBlur,
                        float resetBndWidth, int maxIter);

/* ---- error metric (ref: calcLsDiff4d :2132, calcLsDiffTempl :895-927).  Synchronises. ------ */
int flof_calc_ls_diff4d(flof_ctx *ctx, const float *i0, const float *i1, float *out_or_null,
                        flof_dim4 d, float correction, int bnd, float *result);
/* ref: calcSmokeDiff4d :2163 */
int flof_calc_smoke_diff4d(flof_ctx *ctx, const float *i0, const float *i1, flof_dim4 d,
                           float correction, int bnd, float *result);

/* ---- multi-scale driver (ref: opticalFlowMultiscale4d :2182, template :936-1173) ---------- */
typedef struct {
	float wSmooth, wEnergy, postVelBlur, cgAccuracy, cfl, resetBndWidth;
	int multiStep, projSizeThresh, minGridSize, doFinalProject;
} flof_multiscale_params; /* defaults of the reference: flof_multiscale_defaults() */
typedef struct {
	int n_solves;
	int cg_iters[64];    /* per opticalFlow4d call, in call order */
	float cg_ms[64];     /* device time of each CG solve (ms) */
	int64_t cg_cells[64];/* cells of the level the solve ran on */
	int n_errs;
	float errs[64];      /* "Current error" values in call order, then the final error */
	float total_ms;      /* device time of the whole call */
} flof_multiscale_trace;
void flof_multiscale_defaults(flof_multiscale_params *p);
int flof_optical_flow_multiscale4d(flof_ctx *ctx, float *vel, const float *i0, const float *i1,
                                   flof_dim4 d, const flof_multiscale_params *p,
                                   flof_multiscale_trace *trace_or_null, float *err_out);
/* Same with HOST buffers: H2D of i0,i1,vel, solve, D2H of vel (the end-to-end plugin call). */
int flof_optical_flow_multiscale4d_host(flof_ctx *ctx, float *vel_h, const float *i0_h,
                                        const float *i1_h, flof_dim4 d,
                                        const flof_multiscale_params *p,
                                        flof_multiscale_trace *trace_or_null, float *err_out);

/* ---- pre-processing (ref: optflow4d.cpp:1254-1266, 1323-1439) ----------------------------- */
int flof_repeat_frame4d(flof_ctx *ctx, float *phi, flof_dim4 d, float srct, float range, int bnd);
/* marker_or_null receives the Grid4d<int> layer grid `tmp` (bit-exact with the reference) */
int flof_extrap4d_ls_simple(flof_ctx *ctx, float *phi, flof_dim4 d, int distance, int inside,
                            int *marker_or_null);
int flof_extrapolate_vec4_simple(flof_ctx *ctx, float *vel, const float *phi, flof_dim4 d,
                                 int distance);

/* ---- mode 3: sliced hi-res load + per-frame lookup (ref: optflow4d.cpp:1445-1665, 1951-2105) */
/* phi(.,.,.,slice) += wc * trilinear(src3)       ref: knLoadPlaceGridIpol :1445 */
int flof_load_place_slice(flof_ctx *ctx, float *phi, flof_dim4 d, int slice, const float *src3,
                          flof_dim3 sd, const float srcFac3[3], const float off3[3], float wc);
/* phi(.,.,.,t) *= w[t] where w[t] > 0, bnd 1     ref: knLoadPlaceGridRescale :1454 (w on host) */
int flof_load_place_rescale(flof_ctx *ctx, float *phi, flof_dim4 d, const float *w_host);
/* src3 interior += off, 2 outer shells = 1       ref: :1569-1578 */
int flof_slice_fix_levelset(flof_ctx *ctx, float *src3, flof_dim3 sd, float sdfIsoOff);
/* whole loadPlaceGrid4d with the 3D slices resident on the device (slices[i] = file
 * fileIdxStart+i)                                ref: loadPlaceGrid4d :1464-1595 */
int flof_load_place_grid4d(flof_ctx *ctx, const float *slices, int nslices, flof_dim3 sd,
                           float *phi, flof_dim4 d, const float offset[4], const float scale[4],
                           int fileIdxStart, int fileIdxEnd, int debugSkipLoad, float spread,
                           const float overrideSize[4], float overrideTimeOff,
                           int overrideGoodRegion, float loadTimeScale, int rescaleSdfValues,
                           float sdfIsoOff, float repeatStartFrame);
int flof_shift_forw_grid4d(flof_ctx *ctx, float *phi, flof_dim4 d, int overrideGoodRegion);
/* dst3(x) = phi.interp((x+.5, time+.5) - fac*trilinear(vslice)(x)*dt) for cells inside
 * max(10, bordSkip)                              ref: knSemiLagrangeLookupSlice4d_WithVel :1648 */
int flof_lookup_slice4d_with_vel(flof_ctx *ctx, float *dst3, flof_dim3 d3, const float *phi,
                                 flof_dim4 d, float time, float dt, const float *vslice_vec4,
                                 flof_dim3 vd, const float srcFac3[3], const float off3[3],
                                 const float fac[4], int bordSkip);
/* loadAdvectTimeSlice_OptRun without defo volumes, deformation resident on the device
 * (defo = whole Grid4d<Vec4> of the mode-1 result)   ref: :1951-2105 */
int flof_load_advect_time_slice(flof_ctx *ctx, const float *defo_vec4, flof_dim4 dd, float *dst3,
                                flof_dim3 d3, const float *phi, flof_dim4 d, float time,
                                float blendAlpha, float loadTimeScale, const float defoOffset[4],
                                const float defoScale[4], const float defoFactor[4],
                                const float overrideSize[4], float overrideTimeOff, int bordSkip,
                                float defoAniFac);
/* ---- deformation volumes (flof.py `thirdload`: two / three deformations composed per output frame) -----------------
 * The reference keeps, per deformation file, a window of Tw = int(dimT * max(0.2, partialLoadFac)) time slices around the
 * current source time (LoadAdvectData::updateDefoVol :1822-1863).  Here the whole volume is resident on the device and the
 * window is refreshed from it by the same rule (lastT == t: nothing; lastT + 1 == t: shift by one slice and fetch the
 * last; else refetch all; slice index clamp(t - Tw/2 + tl, 0, dimT-1)).  win: wd.nt slices of wd.nx*ny*nz Vec4.
 * vt: the one-slice scratch grid the slices pass through (the same grid flof_defovol_compose writes, like lats.tmp);
 * *filepos: position of the reference's re-used gz handle of this file (-1 before its first use, fileio.cpp:903-919). */
int flof_defovol_window_update(flof_ctx *ctx, float *win_vec4, flof_dim4 wd, const float *vol_vec4, int dimT, int t,
                               int lastT, float *vt_vec4, int *filepos);
/* vt (ONE slice, Vec4) from the windows: 2 volumes (dvol3 NULL; doAligned 0 :2019-2033, 1 :2035-2060 -- needs the
 * window-sized scratch dvt) or 3 volumes (:2063-2081).  tcoord = srcTime - (t - Tw/2).  The caller then runs
 * flof_lookup_slice4d_with_vel on vt with dt = 1 (ref :2086-2104). */
int flof_defovol_compose(flof_ctx *ctx, float *vt_vec4, const float *dvol1, const float *dvol2, const float *dvol3,
                         float *dvt_scratch, flof_dim4 wd, float tcoord, int doAligned, float blendAlpha, float thirdAlpha,
                         float fourthAlpha);
/* the source-time arithmetic of loadAdvectTimeSlice[_OptRun] (ref :1962-1985): host only, no device work */
int flof_lats_source_time(flof_dim4 dd, flof_dim4 d, float time, float loadTimeScale, const float defoOffset[4],
                          const float defoScale[4], const float overrideSize[4], float *srcTime, int *t, int *tp1,
                          float *tw, float srcFac3[3], float off3[3]);
/* the unoptimised loadAdvectTimeSlice (ref :1671-1760): look-up of every interior cell (bnd 1) of dst3 with dt =
 * blendAlpha, plus its optional debug outputs debugVel (Vec3, 3 floats per cell) / debugVelT (either may be NULL) */
int flof_load_advect_time_slice_unopt(flof_ctx *ctx, const float *defo_vec4, flof_dim4 dd, float *dst3, flof_dim3 d3,
                                      const float *phi, flof_dim4 d, float time, float blendAlpha, float loadTimeScale,
                                      const float defoOffset[4], const float defoScale[4], const float defoFactor[4],
                                      const float overrideSize[4], float overrideTimeOff, float defoAniFac, int zeroVel,
                                      float *dbgVel3, float *dbgVelT);
/* ---- 3D instantiations of the optical-flow templates (scenes/opticalFlowSimple3d.py) ----------------
 * Grids are nx*ny*nz arrays, x fastest; velocities are Vec3 AoS (3 floats per cell) like the reference's Grid<Vec3>.
 * nz == 1 is a 2D grid (the DIM = 2 instantiation of scenes/ofblend2dTest.py): the reference's is3D() switches apply. */
/* ref: opticalFlowMultiscale3d optflow4d.cpp:1175-1188 (template :936-1173 with Grid<Real> / Grid<Vec3>).  Synchronises. */
int flof_optical_flow_multiscale3d(flof_ctx *ctx, float *vel_vec3, const float *i0, const float *i1, flof_dim3 d,
                                   const flof_multiscale_params *p, flof_multiscale_trace *tr, float *err_out);
/* ref: corrVelsOf3d :803-812 (corrVelsOfTempl :737-802; phiCurr and threshNorm are unused there).  Synchronises. */
int flof_corr_vels_of3d(flof_ctx *ctx, float *dst_vec3, float *vel_vec3, const float *phiOrg, const float *phiTarget,
                        flof_dim3 d, float threshPhi, float postVelBlur, float resetBndWidth, int maxIter);
/* ref: advectSemiLagrangeCfl :863-872 (centred velocities, CFL sub-steps :841-853; flags / order / orderSpace are
 * unused there); elem 1 = Grid<Real> payload, 3 = Grid<Vec3>.  advectCent3d :836 = the same with one step (cfl above
 * max|vel|).  Synchronises. */
int flof_advect_semi_lagrange_cfl3d(flof_ctx *ctx, float cfl, const float *vel_vec3, float *grid, int elem, flof_dim3 d,
                                    float velFactor);
/* ref: calcLsDiff3d :928-933 (calcLsDiffTempl :895-927).  Synchronises. */
int flof_calc_ls_diff3d(flof_ctx *ctx, const float *i0, const float *i1, float *out_or_null, flof_dim3 d,
                        float correction, int bnd, float *result);

/* ref: simpleBlurSpecial test.cpp:127 */
int flof_simple_blur_special(flof_ctx *ctx, float *a, flof_dim3 d, int iter, float thresh,
                             int bord);

#ifdef __cplusplus
}
#endif
#endif /* FLOF_B200_H */
