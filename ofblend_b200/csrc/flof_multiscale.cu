// flof_multiscale.cu -- error metric and the coarse-to-fine multi-step driver of mode 1.
// ref: calcLsDiffTempl :895-927, calcSmokeDiffTempl :2140-2162,
//      opticalFlowMultiscaleTemplate :936-1173, opticalFlowMultiscale4d :2182-2195.
//
// The driver is host orchestration (like the reference's), but every grid it touches is
// device resident and every step is a kernel launch on the context stream; temporaries come
// from the stream-ordered pool, mirroring the throw-away FluidSolver of each pyramid level
// (ref :980-988).
#include <math.h>

#include "flof_common.cuh"

extern float g_flof_last_cg_ms;
int flof_advect_cfl4d_ex(flof_ctx *ctx, float cfl, const float *vel, float *grid, flof_dim4 d, int elem,
                         float velFactor, int grid_complete);
int flof_optical_flow4d_ex(flof_ctx *ctx, float *vel, const float *i0, const float *i1, float *rhsT, flof_dim4 d,
                            float wSmooth, float wEnergy, float postVelBlur, float cgAccuracy,
                            float resetBndWidth, int vel_is_zero, int *cgIters, float *cgRes);

// 3D operators (flof_dim3.cu): padded float4 velocities, d.nt == 1
int flof3_interpol_grid(flof_ctx *ctx, float *dst, flof_dim4 td, const float *src, flof_dim4 sd, int elem);
int flof3_advect_cfl(flof_ctx *ctx, float cfl, const float *vel, float *grid, flof_dim4 d, int elem, float velFactor);
int flof3_set_bound_neumann(flof_ctx *ctx, float *grid, flof_dim4 d, int elem, int w);
int flof3_optical_flow(flof_ctx *ctx, float *vel, const float *i0, const float *i1, float *rhsT, flof_dim4 d, float wSmooth,
                       float wEnergy, float postVelBlur, float cgAccuracy, float resetBndWidth, int vel_is_zero, int *cgIters,
                       float *cgRes);
int flof3_corr_vels(flof_ctx *ctx, float *dst, float *vel, const float *phiOrg, const float *phiTarget, flof_dim4 d,
                    float threshPhi, float postVelBlur, float resetBndWidth, int maxIter);

// ------------------------------------------------------------------ error metric ----------
// SMOKE = false: sign-mismatch masked, clamped |diff| (ref :902-912); SMOKE = true: plain |diff|.
// the metric kernels are launched over a 1-D grid of row-blocks; each block walks its rows in
// a fixed order so the fp64 sum is deterministic
template <bool SMOKE>
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_ls_diff_rows(const float *__restrict__ i0, const float *__restrict__ i1, float *__restrict__ out,
                   flof_dim4 d, float correction, int bnd, int row0, int row1, flof_reduce_scratch *red)
{
	__shared__ double sh[32];
	double acc = 0.;
	for (int row = row0 + blockIdx.x; row < row1; row += gridDim.x) {  // one row = nx cells
		const int j = row % d.ny, k = (row / d.ny) % d.nz, t = row / (d.ny * d.nz);
		for (int i = threadIdx.x; i < d.nx; i += blockDim.x) {
			// FOR_IJKT_BND kernel.h:62-68: t is bounded only on a 4D grid (sizeT > 1), z only on a 3D one
			if (i < bnd || j < bnd || i >= d.nx - bnd || j >= d.ny - bnd) continue;
			if (d.nz > 1 && (k < bnd || k >= d.nz - bnd)) continue;
			if (d.nt > 1 && (t < bnd || t >= d.nt - bnd)) continue;
			const int64_t c = flof_idx(d, i, j, k, t);
			const float a = __ldg(i0 + c), b = __ldg(i1 + c);
			if (SMOKE) {
				acc += (double)(fabsf(a - b) * correction);
			} else if ((a < 0.f && b < 0.f) || (a >= 0.f && b >= 0.f)) {
				if (out) out[c] = 0.f;
			} else {
				float dv = fabsf(a - b) * correction;
				if (dv > 1.f) dv = 1.f;
				acc += (double)dv;
				if (out) out[c] = dv;
			}
		}
	}
	acc = flof_block_sum(acc, sh);
	if (threadIdx.x == 0) red->dsum[2][blockIdx.x] = acc;
	if (flof_last_block(&red->counter[3])) {
		double s = 0.;
		for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) s += red->dsum[2][b];
		s = flof_block_sum(s, sh);
		if (threadIdx.x == 0) red->out_d[2] = s;
	}
}

template <bool SMOKE>
static int ls_diff(flof_ctx *ctx, const float *i0, const float *i1, float *out, flof_dim4 d, float correction,
                   int bnd, float *result)
{
	int ta, tb;
	flof_slab(ctx, d.nt, &ta, &tb);  // sharded level: sum this rank's slices, then all-reduce
	const int row0 = ta * d.ny * d.nz, row1 = tb * d.ny * d.nz;
	const int64_t rows = row1 - row0;
	int blocks = ctx->sm_count * 8;
	if (blocks > FLOF_MAX_PARTIALS) blocks = FLOF_MAX_PARTIALS;
	if (blocks > rows) blocks = (int)rows;
	const int threads = d.nx >= 128 ? 128 : (d.nx > 32 ? 64 : 32);
	FLOF_LAUNCH(k_ls_diff_rows<SMOKE>, blocks, threads, 0, i0, i1, out, d, correction, bnd, row0, row1, ctx->red);
	if (tb - ta != d.nt) FLOF_RET(flof_allreduce_f64_sum(ctx, &ctx->red->out_d[2], 1));
	double *h = (double *)ctx->pinned;
	FLOF_CK(cudaMemcpyAsync(h, &ctx->red->out_d[2], sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	FLOF_CK(cudaStreamSynchronize(ctx->stream));
	double accu = h[0];
	const int sx = d.nx - 2 * bnd, sy = d.ny - 2 * bnd, sz = d.nz > 1 ? d.nz - 2 * bnd : d.nz, st = d.nt > 1 ? d.nt - 2 * bnd : d.nt;  // ref :915-922
	accu *= 1000.;
	if (d.nt > 1) accu *= 1000.;
	accu *= 1. / (double)(sx * sy * sz * st);
	*result = (float)accu;
	return FLOF_OK;
}

extern "C" int flof_calc_ls_diff4d(flof_ctx *ctx, const float *i0, const float *i1, float *out, flof_dim4 d,
                                   float correction, int bnd, float *result)
{
	return ls_diff<false>(ctx, i0, i1, out, d, correction, bnd, result);
}
extern "C" int flof_calc_smoke_diff4d(flof_ctx *ctx, const float *i0, const float *i1, flof_dim4 d,
                                      float correction, int bnd, float *result)
{
	return ls_diff<true>(ctx, i0, i1, NULL, d, correction, bnd, result);
}

// ------------------------------------------------------------------ multi-scale driver ----
extern "C" void flof_multiscale_defaults(flof_multiscale_params *p)
{  // defaults of opticalFlowMultiscale4d, ref :2182-2189
	p->wSmooth = 0.f;
	p->wEnergy = 0.f;
	p->postVelBlur = 0.f;
	p->cgAccuracy = 1e-04f;
	p->cfl = 999.f;
	p->resetBndWidth = -1.f;
	p->multiStep = 1;
	p->projSizeThresh = 9999;
	p->minGridSize = 10;
	p->doFinalProject = 0;
}

namespace {

struct DevBuf {  // RAII for pool temporaries inside the recursive driver
	flof_ctx *ctx;
	void *p;
	DevBuf(flof_ctx *c) : ctx(c), p(NULL) {}
	~DevBuf() { flof_tmp_free(ctx, p); }
	int alloc(size_t bytes, bool zero) { return flof_tmp_alloc(ctx, &p, bytes, zero); }
	float *f() { return (float *)p; }
};

void tr_solve(flof_multiscale_trace *tr, int iters, int64_t cells)
{
	if (!tr) return;
	if (tr->n_solves < 64) {
		tr->cg_iters[tr->n_solves] = iters;
		tr->cg_ms[tr->n_solves] = g_flof_last_cg_ms;
		tr->cg_cells[tr->n_solves] = cells;
	}
	tr->n_solves++;
}
void tr_err(flof_multiscale_trace *tr, float e)
{
	if (!tr) return;
	if (tr->n_errs < 64) tr->errs[tr->n_errs] = e;
	tr->n_errs++;
}

#define MS_RET(call)                       \
	do {                                   \
		int r__ = (call);                  \
		if (r__ != FLOF_OK) return r__;    \
	} while (0)

// The driver serves the 4D template instantiation and the 3D one (Grid<Real> / Grid<Vec3>, ref :1175-1188): a level with
// d.nt == 1 is a 3D grid (velocities padded to float4) and every operator below dispatches to its 3D form (flof_dim3.cu).
inline bool is3(flof_dim4 d) { return d.nt == 1; }
// complete: `grid` is valid on all slices (fresh copy of an input) -> no all-gather on a sharded level
int advect_cfl(flof_ctx *ctx, float cfl, const float *vel, float *grid, flof_dim4 d, int elem, int complete)
{
	if (is3(d)) return flof3_advect_cfl(ctx, cfl, vel, grid, d, elem, 1.f);
	return flof_advect_cfl4d_ex(ctx, cfl, vel, grid, d, elem, 1.f, complete);
}
int resample(flof_ctx *ctx, float *dst, flof_dim4 td, const float *src, flof_dim4 sd, int elem)
{
	if (is3(td)) return flof3_interpol_grid(ctx, dst, td, src, sd, elem);
	return flof_interpol_grid_templ(ctx, dst, td, src, sd, elem);
}
int neumann(flof_ctx *ctx, float *grid, flof_dim4 d, int elem, int w)
{
	if (is3(d)) return flof3_set_bound_neumann(ctx, grid, d, elem, w);
	return flof_grid4d_set_bound_neumann(ctx, grid, d, elem, w);
}
int of_solve(flof_ctx *ctx, float *vel, const float *i0, const float *i1, flof_dim4 d, float wSmooth, float wEnergy,
             float postVelBlur, float cgAccuracy, float resetBndWidth, int *cgIters, float *cgRes)
{
	if (is3(d))
		return flof3_optical_flow(ctx, vel, i0, i1, NULL, d, wSmooth, wEnergy, postVelBlur, cgAccuracy, resetBndWidth, 1, cgIters,
		                          cgRes);
	return flof_optical_flow4d_ex(ctx, vel, i0, i1, NULL, d, wSmooth, wEnergy, postVelBlur, cgAccuracy, resetBndWidth, 1, cgIters,
	                              cgRes);
}
int corr_vels(flof_ctx *ctx, float *dst, float *vel, const float *phiOrg, const float *phiTarget, flof_dim4 d, float threshPhi,
              float postVelBlur, float resetBndWidth, int maxIter)
{
	if (is3(d)) return flof3_corr_vels(ctx, dst, vel, phiOrg, phiTarget, d, threshPhi, postVelBlur, resetBndWidth, maxIter);
	return flof_corr_vels_of4d(ctx, dst, vel, phiOrg, phiTarget, d, threshPhi, postVelBlur, resetBndWidth, maxIter);
}

// RAII: shard the level being processed along t over the ranks, restore the caller's state on exit
struct ShardScope {
	flof_ctx *ctx;
	decltype(flof_ctx::sh) saved;
	ShardScope(flof_ctx *c, flof_dim4 d) : ctx(c), saved(c->sh)
	{
		const int64_t n = flof_cells(d);
		const int P = ctx->nranks;
		const bool ok = ctx->comm && P > 1 && d.nt % P == 0 && d.nt / P >= 4 && n >= ctx->shard_min_cells;
		ctx->sh.active = ok ? 1 : 0;
		if (ok) {
			ctx->sh.nt = d.nt;
			ctx->sh.n3 = (int64_t)d.nx * d.ny * d.nz;
			flof_slab_range(d.nt, P, ctx->rank, &ctx->sh.ta, &ctx->sh.tb);
		}
	}
	void reapply() {}
	~ShardScope() { ctx->sh = saved; }
};

// ref opticalFlowMultiscaleTemplate :936-1173
int multiscale(flof_ctx *ctx, float *vel, const float *i0, const float *i1, flof_dim4 d,
               const flof_multiscale_params &P, int level, int multiStep, bool doFinalProject,
               flof_multiscale_trace *tr, float *errOut)
{
	const int64_t n = flof_cells(d);
	const size_t rb = sizeof(float) * (size_t)n, vb = rb * 4;
	const int resetBnd = P.resetBndWidth > 0 ? (int)(P.resetBndWidth * d.nx) + 1 : 0;
	const float projMaxDist = 4.f;
	const float projMaxIter = 40.f;
	const float lsDiffFac = (float)(0.1 / 20.);
	ctx->prof_cells = n;  // tag launches with the level they work on (flof_profile_*)
	// Multi-GPU: on entry vel, i0, i1 are complete on every rank.  If this level is sharded, every
	// operator below works on this rank's t-slab (ghost slices / all-gathers where a stencil or a
	// gather needs them); vel is all-gathered before returning so the contract holds for the caller.
	// Levels that are too small (or whose T is not divisible) run replicated on every rank.
	const size_t vslice = sizeof(float) * 4 * (size_t)d.nx * d.ny * d.nz;

	DevBuf i0warped(ctx);
	MS_RET(i0warped.alloc(rb, false));
	MS_RET(flof_memcpy_d2d(ctx, i0warped.p, i0, rb));
	float errPreOf = 0.f;
	MS_RET(flof_calc_ls_diff4d(ctx, i0, i1, NULL, d, lsDiffFac, resetBnd, &errPreOf));
	(void)errPreOf;

	if (d.nx > P.minGridSize) {
		// the coarser level decides about its own sharding; down-sampling runs on complete inputs
		ShardScope none(ctx, flof_dim4{ 0, 0, 0, 0 });
		flof_dim4 s = { d.nx / 2, d.ny / 2, d.nz > 1 ? d.nz / 2 : 1, is3(d) ? 1 : d.nt / 2 };  // ref :980-983
		if (s.nx < 3 || s.ny < 3 || (d.nz > 1 && s.nz < 3) || (!is3(d) && s.nt < 3))
			return flof_fail(ctx, FLOF_ERR_ARG, "opticalFlowMultiscale4d: coarse level %dx%dx%dx%d too small",
			                 s.nx, s.ny, s.nz, s.nt);
		const int64_t ns = flof_cells(s);
		DevBuf velSm(ctx), i0Sm(ctx), i1Sm(ctx);
		MS_RET(velSm.alloc(sizeof(float) * 4 * (size_t)ns, false));
		MS_RET(i0Sm.alloc(sizeof(float) * (size_t)ns, false));
		MS_RET(i1Sm.alloc(sizeof(float) * (size_t)ns, false));
		MS_RET(resample(ctx, i0Sm.f(), s, i0, d, 1));
		MS_RET(resample(ctx, i1Sm.f(), s, i1, d, 1));
		MS_RET(resample(ctx, velSm.f(), s, vel, d, 4));
		const float half[4] = { 0.5f, 0.5f, 0.5f, 0.5f };
		MS_RET(flof_grid_mult_const(ctx, velSm.f(), ns, 4, half));
		float eSm = 0.f;
		MS_RET(multiscale(ctx, velSm.f(), i0Sm.f(), i1Sm.f(), s, P, level + 1, multiStep, doFinalProject, tr, &eSm));
		ctx->prof_cells = n;
		MS_RET(resample(ctx, vel, d, velSm.f(), s, 4));  // complete (cheap), sliced below
		const float two[4] = { 2.f, 2.f, 2.f, 2.f };
		MS_RET(flof_grid_mult_const(ctx, vel, n, 4, two));
	}
	ShardScope shard(ctx, d);
	// peer mailboxes sized once per level for its largest halo (two Vec4 slices: Gaussian blur with s = 2)
	if (ctx->sh.active) MS_RET(flof_p2p_ensure(ctx, 2 * sizeof(float) * 4 * (size_t)ctx->sh.n3));

	// pre-warp (ref :1011-1018)
	MS_RET(advect_cfl(ctx, P.cfl, vel, i0warped.f(), d, 1, 1));
	MS_RET(neumann(ctx, i0warped.f(), d, 1, 0));
	float errCurr = 0.f;
	MS_RET(flof_calc_ls_diff4d(ctx, i0warped.f(), i1, NULL, d, lsDiffFac, resetBnd, &errCurr));

	bool doProject = false;
	if (d.nx > P.projSizeThresh) {
		doProject = true;
		multiStep = 1;
		if (doFinalProject) doFinalProject = false;
	}

	const int MAX_STEPS = 10;
	if (multiStep >= MAX_STEPS) return flof_fail(ctx, FLOF_ERR_ARG, "Too many of substeps!");
	if (multiStep > 1) {
		// ref :1031-1115
		DevBuf *vs[MAX_STEPS], *vs2[MAX_STEPS];
		for (int of = 0; of < MAX_STEPS; ++of) vs[of] = vs2[of] = NULL;
		struct Cleanup {
			DevBuf **a, **b;
			~Cleanup()
			{
				for (int q = 0; q < 10; ++q) {
					delete a[q];
					delete b[q];
				}
			}
		} cleanup = { vs, vs2 };
		for (int of = 0; of < multiStep; ++of) {
			vs2[of] = new DevBuf(ctx);
			MS_RET(vs2[of]->alloc(vb, true));
		}
		DevBuf tmpVel(ctx), i0warp2(ctx);
		MS_RET(tmpVel.alloc(vb, false));
		MS_RET(i0warp2.alloc(rb, false));
		float velBlur = P.postVelBlur;
		float errLast = errCurr;
		int ofStepsCurr = multiStep;
		for (int of = 0; of < ofStepsCurr; ++of) {
			vs[of] = new DevBuf(ctx);
			MS_RET(vs[of]->alloc(vb, true));  // zero init
			int iters = 0;
			float cgRes = 0.f;
			// vs[of] is all zero: the smoothness/Tikhonov rhs terms vanish, so the assembly may
			// skip reading it (bit-identical: every term is (+-)0 and rhs -= 0 leaves rhs unchanged)
			MS_RET(of_solve(ctx, vs[of]->f(), i0warped.f(), i1, d, P.wSmooth, P.wEnergy, velBlur, P.cgAccuracy, P.resetBndWidth,
			                &iters, &cgRes));
			tr_solve(tr, iters, n);
			velBlur *= (float)(3. / 4.);
			if (velBlur < 2.f) velBlur = 2.f;
			for (int k = of; k >= 0; --k) MS_RET(flof_memcpy_d2d(ctx, vs2[k]->p, vs[k]->p, vb));
			for (int k = of - 1; k >= 0; --k)
				for (int l = 0; l < k; ++l) MS_RET(advect_cfl(ctx, P.cfl, vs2[k]->f(), vs2[l]->f(), d, 4, 0));
			MS_RET(flof_memcpy_d2d(ctx, tmpVel.p, vel, vb));
			for (int k = of; k >= 0; --k) MS_RET(flof_grid_binary(ctx, tmpVel.f(), vs2[k]->f(), n, 4, FLOF_OP_ADD));
			MS_RET(flof_memcpy_d2d(ctx, i0warp2.p, i0, rb));
			MS_RET(advect_cfl(ctx, P.cfl, tmpVel.f(), i0warp2.f(), d, 1, 1));
			MS_RET(neumann(ctx, i0warp2.f(), d, 1, 0));
			float errC = 0.f;
			MS_RET(flof_calc_ls_diff4d(ctx, i0warp2.f(), i1, NULL, d, lsDiffFac, resetBnd, &errC));
			tr_err(tr, errC);
			MS_RET(flof_memcpy_d2d(ctx, i0warped.p, i0warp2.p, rb));
			if (of > 0 && (errC / errLast) > 0.95) {
				MS_RET(flof_memset0(ctx, vs[of]->p, vb));
				ofStepsCurr = of + 1;
			}
			errLast = errC;
		}
		for (int of = ofStepsCurr - 1; of >= 0; --of)
			for (int l = 0; l < of; ++l) MS_RET(advect_cfl(ctx, P.cfl, vs[of]->f(), vs[l]->f(), d, 4, 0));
		for (int of = 0; of < ofStepsCurr; ++of) MS_RET(flof_grid_binary(ctx, vel, vs[of]->f(), n, 4, FLOF_OP_ADD));
	} else {
		DevBuf velCurr(ctx);
		MS_RET(velCurr.alloc(vb, true));
		if (!doProject) {
			int iters = 0;
			float cgRes = 0.f;
			MS_RET(of_solve(ctx, velCurr.f(), i0warped.f(), i1, d, P.wSmooth, P.wEnergy, P.postVelBlur, P.cgAccuracy,
			                P.resetBndWidth, &iters, &cgRes));
			tr_solve(tr, iters, n);
			MS_RET(flof_grid_binary(ctx, vel, velCurr.f(), n, 4, FLOF_OP_ADD));
		} else {
			DevBuf velTmp2(ctx);
			MS_RET(velTmp2.alloc(vb, true));
			// the projection gathers phiOrg at back-traced positions that leave this rank's slab: complete i0warped first
			// (after the pre-warp it is valid on the own slices only)
			MS_RET(flof_allgather_slabs(ctx, i0warped.p, d.nt, sizeof(float) * (size_t)d.nx * d.ny * d.nz));
			MS_RET(corr_vels(ctx, velCurr.f(), velTmp2.f(), i0warped.f(), i1, d, projMaxDist, P.postVelBlur, P.resetBndWidth,
			                 (int)projMaxIter));
			MS_RET(flof_grid_binary(ctx, vel, velTmp2.f(), n, 4, FLOF_OP_ADD));
			const float m1[4] = { -1.f, -1.f, -1.f, -1.f };
			MS_RET(flof_grid_mult_const(ctx, velCurr.f(), n, 4, m1));
			MS_RET(flof_grid_binary(ctx, vel, velCurr.f(), n, 4, FLOF_OP_ADD));
		}
	}

	// final projection step (ref :1143-1153)
	if ((level == 0) && doFinalProject) {
		DevBuf velCurr(ctx);
		MS_RET(velCurr.alloc(vb, true));
		const float finalProjBlur = 4.f;
		MS_RET(corr_vels(ctx, velCurr.f(), vel, i0, i1, d, projMaxDist, finalProjBlur, P.resetBndWidth, (int)projMaxIter));
	}

	// re-advect and evaluate on the finest level (ref :1155-1170)
	if (level == 0) {
		MS_RET(flof_memcpy_d2d(ctx, i0warped.p, i0, rb));
		MS_RET(advect_cfl(ctx, P.cfl, vel, i0warped.f(), d, 1, 1));
		MS_RET(neumann(ctx, i0warped.f(), d, 1, 0));
		float errFinal = 0.f;
		MS_RET(flof_calc_ls_diff4d(ctx, i0warped.f(), i1, NULL, d, lsDiffFac, resetBnd, &errFinal));
		errCurr = errFinal;
		tr_err(tr, errFinal);
	}
	// leave the level with a complete deformation on every rank
	MS_RET(flof_allgather_slabs(ctx, vel, d.nt, vslice));
	*errOut = errCurr;
	return FLOF_OK;
}

}  // namespace

extern "C" int flof_optical_flow_multiscale4d(flof_ctx *ctx, float *vel, const float *i0, const float *i1,
                                              flof_dim4 d, const flof_multiscale_params *p,
                                              flof_multiscale_trace *tr, float *err_out)
{
	FLOF_ARG(p != NULL, "opticalFlowMultiscale4d: params is NULL");
	FLOF_ARG(d.nx >= 3 && d.ny >= 3 && d.nz >= 3 && d.nt >= 3, "opticalFlowMultiscale4d: grid too small");
	if (tr) memset(tr, 0, sizeof(*tr));
	cudaEventRecord(ctx->ev[0], ctx->stream);
	float e = 0.f;
	int rc = multiscale(ctx, vel, i0, i1, d, *p, 0, p->multiStep, p->doFinalProject != 0, tr, &e);
	cudaEventRecord(ctx->ev[1], ctx->stream);
	ctx->prof_cells = 0;
	if (rc != FLOF_OK) return rc;
	FLOF_CK(cudaEventSynchronize(ctx->ev[1]));
	if (tr) cudaEventElapsedTime(&tr->total_ms, ctx->ev[0], ctx->ev[1]);
	if (err_out) *err_out = e;
	if (ctx->nranks > 1) FLOF_RET(flof_comm_p2p_status(ctx, NULL, NULL));  // a timed-out peer wait invalidates the result
	return FLOF_OK;
}

// 3D entry of the same driver (flof_optical_flow_multiscale3d in flof_dim3.cu pads / unpads the Vec3 field around it)
int flof_multiscale_run3d(flof_ctx *ctx, float *vel4, const float *i0, const float *i1, flof_dim4 d,
                          const flof_multiscale_params *p, flof_multiscale_trace *tr, float *err_out)
{
	if (tr) memset(tr, 0, sizeof(*tr));
	cudaEventRecord(ctx->ev[0], ctx->stream);
	float e = 0.f;
	int rc = multiscale(ctx, vel4, i0, i1, d, *p, 0, p->multiStep, p->doFinalProject != 0, tr, &e);
	cudaEventRecord(ctx->ev[1], ctx->stream);
	ctx->prof_cells = 0;
	if (rc != FLOF_OK) return rc;
	FLOF_CK(cudaEventSynchronize(ctx->ev[1]));
	if (tr) cudaEventElapsedTime(&tr->total_ms, ctx->ev[0], ctx->ev[1]);
	if (err_out) *err_out = e;
	return FLOF_OK;
}

extern "C" int flof_optical_flow_multiscale4d_host(flof_ctx *ctx, float *vel_h, const float *i0_h,
                                                   const float *i1_h, flof_dim4 d,
                                                   const flof_multiscale_params *p,
                                                   flof_multiscale_trace *tr, float *err_out)
{
	const int64_t n = flof_cells(d);
	const size_t rb = sizeof(float) * (size_t)n, vb = rb * 4;
	DevBuf vel(ctx), i0(ctx), i1(ctx);
	MS_RET(vel.alloc(vb, false));
	MS_RET(i0.alloc(rb, false));
	MS_RET(i1.alloc(rb, false));
	if (ctx->comm && ctx->nranks > 1 && d.nt % ctx->nranks == 0) {
		// N ranks: every rank uploads only its own t-slab of the (replicated) host buffers -- 1/N of the bytes over PCIe --
		// and the slabs are all-gathered over NVLink (NCCL, in place)
		int ta, tb;
		flof_slab_range(d.nt, ctx->nranks, ctx->rank, &ta, &tb);
		const size_t s1 = sizeof(float) * (size_t)d.nx * d.ny * d.nz, o1 = s1 * (size_t)ta, n1 = s1 * (size_t)(tb - ta);
		MS_RET(flof_memcpy_h2d(ctx, (char *)i0.p + o1, (const char *)i0_h + o1, n1));
		MS_RET(flof_memcpy_h2d(ctx, (char *)i1.p + o1, (const char *)i1_h + o1, n1));
		MS_RET(flof_memcpy_h2d(ctx, (char *)vel.p + 4 * o1, (const char *)vel_h + 4 * o1, 4 * n1));
		MS_RET(flof_allgather_bytes(ctx, i0.p, o1, n1));
		MS_RET(flof_allgather_bytes(ctx, i1.p, o1, n1));
		MS_RET(flof_allgather_bytes(ctx, vel.p, 4 * o1, 4 * n1));
	} else {
		MS_RET(flof_memcpy_h2d(ctx, i0.p, i0_h, rb));
		MS_RET(flof_memcpy_h2d(ctx, i1.p, i1_h, rb));
		MS_RET(flof_memcpy_h2d(ctx, vel.p, vel_h, vb));
	}
	MS_RET(flof_optical_flow_multiscale4d(ctx, vel.f(), i0.f(), i1.f(), d, p, tr, err_out));
	// N ranks: by default every rank receives the deformation in its host buffer; with the option host_result_rank = r
	// only rank r downloads it (the rank that writes the .uni file) and the others skip their 16 B/cell over PCIe
	if (ctx->opt.host_result_rank < 0 || ctx->nranks <= 1 || ctx->rank == ctx->opt.host_result_rank)
		MS_RET(flof_memcpy_d2h(ctx, vel_h, vel.p, vb));
	return FLOF_OK;
}
