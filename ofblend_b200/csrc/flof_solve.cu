// flof_solve.cu -- 4D optical-flow system assembly and the matrix-free Jacobi-PCG.
// ref: FixedMatrixOF / applyMat / dotProd / getMaxNorm / addScaled / GridCGOptflow4d
//      optflow4d.cpp:179-355 and opticalFlowDim<.,.,4> :361-553.
//
// Design (B200): the reference stores the matrix explicitly (offd N + blockd 4N floats, 17 GB at
// 128^4).  Here the system is matrix-free: per cell only grad (float4) and rhs (float4) are kept;
// the 4x4 block  grad grad^T + diag*I  and the Jacobi preconditioner are recomputed in registers
// with the reference's fp32 operation order, so every matrix entry has the identical fp32 value.
// Unknown layout = Vec4 AoS (idx = cell*4 + d, ref :60-64), i.e. one 128-bit access per cell.
//
// One CG iteration = three streaming kernels over the cells (224 B/cell/iteration), each CTA walking one
// contiguous SEGMENT of the cell range (seq_make_part: all CTAs resident, segments tile the t-slices so that the t-1 / t+1
// stencil neighbours are the centre cells of another resident CTA at the same moment):
//   A: tmp = A*srch                       (+ partial fp64 dot(srch,tmp))
//   B: result += alpha*srch; res -= alpha*tmp   (+ partial signed max(res), fp64 dot(res*precond,res))
//   C: srch = res*precond + beta*srch     (the stop test and the state advance ran in the tail of B)
// Scalars (alpha, beta, sigma, residual) never leave the device: each reducing kernel ends with a
// "last block finishes" pass that sums the per-block partials in index order (deterministic) and
// updates the flof_cg_state; the host only polls the `done` flag every few iterations.
// On a t-sharded level the same tail also all-reduces the partials over the ranks through the NVLink
// peer mailboxes (flof_p2p.cuh), so reduction + collective are one kernel.
// Border rows are identity with zero rhs (ref :424-433): marked by grad.x = NaN so the solver
// kernels need no index arithmetic at all.
#include <math.h>
#include <stdlib.h>

#include "flof_common.cuh"
#include "flof_p2p.cuh"

int flof_reset_border_vec4(flof_ctx *ctx, float *vel, flof_dim4 d, int resetBnd);
int flof_optical_flow4d_ex(flof_ctx *ctx, float *vel, const float *i0, const float *i1, float *rhsT, flof_dim4 d,
                           float wSmooth, float wEnergy, float postVelBlur, float cgAccuracy,
                           float resetBndWidth, int vel_is_zero, int *cgIters, float *cgRes);

struct flof_of_consts {
	float offd;   // -wSmooth * mDx2Inv          ref :476
	float diag;   // 8*wSmooth*mDx2Inv + wEnergy ref :478-480
};

// dim: DIM of the reference's template instantiation (4; 3 for a 3D problem embedded as one slice, flof_dim3.cu)
static flof_of_consts of_consts(float wSmooth, float wEnergy, int dim = 4)
{
	flof_of_consts k;
	const float mDx2Inv = 1.f;
	k.offd = -wSmooth * mDx2Inv;
	volatile float diag = 0.f;
	diag += (float)(2 * dim) * wSmooth * mDx2Inv;
	diag += wEnergy;
	k.diag = diag;
	return k;
}

// ------------------------------------------------------------------ K1 assembly -----------
// ref :398-493.  Reads i0, i1 (+ the 6/8 stencil neighbours of i1 from cache) and, only if
// has_vel, the incoming vel with its clamped neighbours; writes grad and rhs (16 B each).
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_of_assemble(float4 *__restrict__ grad, float4 *__restrict__ rhs, const float *__restrict__ i0,
                  const float *__restrict__ i1, const float4 *__restrict__ vel, flof_kd d,
                  float wSmooth, float wEnergy, float mDx, float dxf, int has_vel)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t)) return;
	const int64_t c = flof_idx(d, i, j, k, t);
	if (!flof_in_bounds(d, i, j, k, t, 1)) {
		grad[c] = make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, 0.f);  // identity-row marker
		rhs[c] = make_float4(0.f, 0.f, 0.f, 0.f);
		return;
	}
	const int64_t sY = d.nx, sZ = (int64_t)d.nx * d.ny, sT = sZ * d.nz;
	const float mDt = 1.0f;
	const float tderiv = (__ldg(i1 + c) - __ldg(i0 + c)) / mDt;
	float g[4];
	g[0] = (__ldg(i1 + c + 1) - __ldg(i1 + c - 1)) * dxf;
	g[1] = (__ldg(i1 + c + sY) - __ldg(i1 + c - sY)) * dxf;
	g[2] = (__ldg(i1 + c + sZ) - __ldg(i1 + c - sZ)) * dxf;
	g[3] = (__ldg(i1 + c + sT) - __ldg(i1 + c - sT)) * dxf;
	float r[4];
#pragma unroll
	for (int dd = 0; dd < 4; ++dd) r[dd] = -g[dd] * tderiv;
	if (has_vel) {
		// smoothness + Tikhonov terms of the incoming field (ref :463-472, 491); neighbour
		// order t-1, z-1, y-1, x-1, x+1, y+1, z+1, t+1 with clamped indices
		const float4 vc4 = __ldg(vel + c);
		const float vc[4] = { vc4.x, vc4.y, vc4.z, vc4.w };
		const int nb[8][4] = { { 0, 0, 0, -1 }, { 0, 0, -1, 0 }, { 0, -1, 0, 0 }, { -1, 0, 0, 0 },
			                   { 1, 0, 0, 0 },  { 0, 1, 0, 0 },  { 0, 0, 1, 0 },  { 0, 0, 0, 1 } };
		float4 vn4[8];
#pragma unroll
		for (int m = 0; m < 8; ++m) {
			const int ti = max(0, min(d.nx - 1, i + nb[m][0])), tj = max(0, min(d.ny - 1, j + nb[m][1]));
			const int tk = max(0, min(d.nz - 1, k + nb[m][2])), tt = max(0, min(d.nt - 1, t + nb[m][3]));
			vn4[m] = __ldg(vel + flof_idx(d, ti, tj, tk, tt));
		}
#pragma unroll
		for (int dd = 0; dd < 4; ++dd) {
#pragma unroll
			for (int m = 0; m < 8; ++m) {
				const float vn = dd == 0 ? vn4[m].x : (dd == 1 ? vn4[m].y : (dd == 2 ? vn4[m].z : vn4[m].w));
				r[dd] -= wSmooth * (vc[dd] - vn) * mDx * 1.f;
			}
			r[dd] -= wEnergy * vc[dd] * mDx;
		}
	}
	grad[c] = make_float4(g[0], g[1], g[2], g[3]);
	rhs[c] = make_float4(r[0], r[1], r[2], r[3]);
}

extern "C" int flof_of_assemble(flof_ctx *ctx, float *grad, float *rhs, const float *i0,
                                const float *i1, const float *vel, flof_dim4 d, float wSmooth,
                                float wEnergy)
{
	FLOF_ARG(d.nx >= 3 && d.ny >= 3 && d.nz >= 3 && d.nt >= 3, "opticalFlow4d: grid too small");
	const float mDx = (float)(1. / d.nx);           // ref :369
	const float dxf = (float)(1. / (2. * mDx));     // ref :441
	dim3 g;
	const flof_kd kd = flof_kdim(ctx, d, &g);  // sharded: i1 (and vel) must be valid on the +-1 ghost slices
	FLOF_LAUNCH(k_of_assemble, g, FLOF_BLOCK, 0, (float4 *)grad, (float4 *)rhs, i0, i1, (const float4 *)vel, kd, wSmooth,
	            wEnergy, mDx, dxf, vel != NULL);
	return FLOF_OK;
}

// ------------------------------------------------------------------ CG kernels ------------
__device__ __forceinline__ bool is_border(const float4 &g) { return g.x != g.x; }

// reciprocal Jacobi diagonal, ref precondInit :331-343: precond = 1./diag, diag = blockd(i,i)
__device__ __forceinline__ float4 precond_of(const float4 &g, float diag)
{
	if (is_border(g)) return make_float4(1.f, 1.f, 1.f, 1.f);
	return make_float4(1.0f / (g.x * g.x + diag), 1.0f / (g.y * g.y + diag), 1.0f / (g.z * g.z + diag),
	                   1.0f / (g.w * g.w + diag));
}
__device__ __forceinline__ double dot4(const float4 &a, const float4 &b)
{  // ref dotProd :234-241: fp32 products accumulated in double
	double s = (double)(a.x * b.x);
	s += (double)(a.y * b.y);
	s += (double)(a.z * b.z);
	s += (double)(a.w * b.w);
	return s;
}
__device__ __forceinline__ float axpy1(float a, double b, float c)
{  // ref addScaled :253-258: a = float(double(a) + b*double(c))
	return (float)((double)a + b * (double)c);
}

// state after the initial residual / sigma are known (ref :286-301)
__device__ __forceinline__ void cg_init_finalize(flof_cg_state *st, double s, float m, float accuracy)
{
	const double residual = (double)m;  // ref :286
	st->iter = 0;
	st->done = 0;
	st->status = 0;
	st->residual = m;
	st->relResidual = 1e10f;  // cgRes initial value, ref :499
	st->resIni = residual;
	st->acc = (double)accuracy * residual;  // ref :292
	st->sigma[0] = s;
	st->sigma[1] = s;
	st->alpha1 = 0.;
	st->sigmaNew = 0.;
	if (residual < (double)FLOF_VECTOR_EPSILON) {  // ref :287-291
		st->done = 1;
		st->status = 2;
		st->relResidual = 0.f;
	} else if (s == 0. || s != s) {  // ref :298-301
		st->done = 1;
		st->status = 3;
	}
}

// end of an iteration's reductions (ref :311-317, 324): relative residual, stop test, sigma ring.  Runs in the tail
// of k_cg_update (one thread, after every block has read the state) -- or as its own launch after the NCCL calls.
__device__ __forceinline__ void cg_advance(flof_cg_state *st, double sigmaNew, float residual, int maxIter)
{
	st->sigmaNew = sigmaNew;
	st->residual = residual;
	st->relResidual = (float)((double)residual / st->resIni);  // ref :312
	const int it = st->iter;
	st->iter = it + 1;  // ret_iterations = iter + 1
	if ((double)residual <= st->acc) {
		st->done = 1;
		st->status = 1;
		return;
	}
	st->sigma[(it + 1) & 1] = sigmaNew;
	if (it + 1 >= maxIter) st->done = 1;  // ref: loop bound cgMaxIter, returns false
}

// fp32 products accumulated in double like dot4; af additionally collects their magnitudes (fp32, an upper bound after
// SEQ_SA_SLACK) for the error margin of the sequential-order dot products
__device__ __forceinline__ double dot4a(const float4 &a, const float4 &b, float &af)
{
	float q = a.x * b.x;
	double s = (double)q;
	af += fabsf(q);
	q = a.y * b.y;
	s += (double)q;
	af += fabsf(q);
	q = a.z * b.z;
	s += (double)q;
	af += fabsf(q);
	q = a.w * b.w;
	s += (double)q;
	af += fabsf(q);
	return s;
}

#include "flof_seqsum_kernels.cuh"

// what the reducing CG kernels need for the sequential-order dot products (pass 1, flof_seqsum.cuh)
struct cg_seq {
	int on;      // 1: leave per-segment prefixes + the work list for k_dot_seq and let k_seq_resolve advance the CG state
	seq_args A;
};
// tail of a reducing kernel in sequential-order mode (ONE block, every thread): per-segment prefixes, on a sharded
// level the all-reduce that also yields the lower ranks' share, classification + work list of pass 2;
// m: the block's max (thread 0), returned all-reduced
__device__ __forceinline__ float cg_seq_tail(const cg_seq &sq, float m, bool with_max, int multi, const flof_p2p_dev &pp,
                                             double *shd, int *shi, double *s_ar)
{
	double tx, ta;
	seq_tail_scan(sq.A, shd, tx, ta);
	double ox = 0., oa = 0.;
	if (multi == 2) {
		if (threadIdx.x == 0) {
			s_ar[0] = tx;
			s_ar[1] = ta;
			s_ar[2] = (double)m;
		}
		p2p_allreduce_block(pp, s_ar, with_max ? 3 : 2, 2, false, s_ar + 4);
		if (with_max) m = (float)s_ar[2];
		ox = s_ar[4];
		oa = s_ar[5];
	}
	seq_tail_classify(sq.A, ox, oa, shi);
	if (threadIdx.x == 0) {
		sq.A.ctl->tot[0] = tx;
		sq.A.ctl->tot[1] = ta;
		sq.A.ctl->off[0] = ox;
		sq.A.ctl->off[1] = oa;
	}
	return m;
}

// the streaming producers (init, update) walk `spc` consecutive segments per CTA
struct cg_walk {
	seq_part part;
	int spc;
};

// res = rhs, result = 0, tmp = res*precond, srch = tmp; residual0 = max(res), sigma = tmp.res
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_cg_init(float4 *__restrict__ x, float4 *__restrict__ res, float4 *__restrict__ srch,
              const float4 *__restrict__ grad, const float4 *__restrict__ rhs, cg_walk wk, float diag,
              float accuracy, int multi, cg_seq sq, flof_p2p_dev pp, flof_reduce_scratch *red, flof_cg_state *st)
{
	__shared__ double shd[32];
	__shared__ float shf[32];
	__shared__ int shi[8];
	__shared__ double s_ar[8];
	double csum = 0.;  // this CTA's tree partial (thread 0)
	float mx = -3.402823466e+38f;
	const int sg1 = min(((int)blockIdx.x + 1) * wk.spc, wk.part.nseg);
	for (int sg = (int)blockIdx.x * wk.spc; sg < sg1; ++sg) {
		double dsum = 0.;
		float af = 0.f;
		const int c0 = sg * wk.part.seg_cells, c1 = min(c0 + wk.part.seg_cells, wk.part.ncells);
		for (int c = c0 + (int)threadIdx.x; c < c1; c += FLOF_BLOCK) {
			const float4 g = __ldg(grad + c), r = __ldg(rhs + c);
			const float4 pc = precond_of(g, diag);
			const float4 z = make_float4(r.x * pc.x, r.y * pc.y, r.z * pc.z, r.w * pc.w);
			x[c] = make_float4(0.f, 0.f, 0.f, 0.f);
			res[c] = r;
			srch[c] = z;
			dsum += dot4a(z, r, af);
			mx = fmaxf(mx, fmaxf(fmaxf(r.x, r.y), fmaxf(r.z, r.w)));
		}
		double asum = (double)af;
		seq_block_sum2(dsum, asum, shd);
		if (threadIdx.x == 0) {
			csum += dsum;
			if (sq.on) {
				sq.A.aggx[sg] = dsum;
				sq.A.agga[sg] = asum;
			}
		}
	}
	mx = flof_block_max(mx, shf);
	if (threadIdx.x == 0) {
		red->dsum[0][blockIdx.x] = csum;
		red->fmax[blockIdx.x] = mx;
	}
	if (flof_last_block(&red->counter[1])) {
		double s = 0.;
		float m = -3.402823466e+38f;
		for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
			s += red->dsum[0][b];
			m = fmaxf(m, red->fmax[b]);
		}
		s = flof_block_sum(s, shd);
		m = flof_block_max(m, shf);
		if (sq.on) {
			// sigma comes from the sequential-order dot product (k_seq_resolve finalizes the state)
			m = cg_seq_tail(sq, m, true, multi, pp, shd, shi, s_ar);
			if (threadIdx.x == 0) {
				st->sigmaNew = s;
				st->residual = m;
				st->done = 0;
				st->seq_inexact = 0;
			}
			return;
		}
		if (multi == 2) {  // sharded, peer mailboxes: combine the slabs of all ranks right here
			if (threadIdx.x == 0) {
				s_ar[0] = s;
				s_ar[1] = (double)m;
			}
			p2p_allreduce_block(pp, s_ar, 2, 1, false);
			s = s_ar[0];
			m = (float)s_ar[1];
		}
		if (threadIdx.x == 0) {
			if (multi == 1) {
				// raw slab results; the ranks are combined by NCCL, then k_cg_init_finalize runs
				st->sigmaNew = s;
				st->residual = m;
				st->done = 0;
			} else {
				cg_init_finalize(st, s, m, accuracy);
			}
		}
	}
}
__global__ void k_cg_init_finalize(float accuracy, flof_cg_state *st) { cg_init_finalize(st, st->sigmaNew, st->residual, accuracy); }
__global__ void k_cg_update_finalize(int maxIter, flof_cg_state *st)
{
	if (st->done) return;
	cg_advance(st, st->sigmaNew, st->residual, maxIter);
}

// A: tmp = A * srch (ref applyMat :211-232), partial dot(srch, tmp)
// Grid-stride over LEAVES of 1024 consecutive cells: the resident CTAs sweep the grid as one compact window (a few
// t-slices), so every neighbour of the stencil has just been fetched by another CTA or is about to be its centre cell --
// srch comes from DRAM once.  (One contiguous segment per CTA, as in the purely streaming kernels, measured 1.5x slower
// here: the z+1 look-ahead of ~1200 independent streams does not fit into L2.)
// SEQ (sequential-order dot products): the per-leaf sums of the products and of their magnitudes go to the segment
// accumulators by one atomic per warp -- their order is irrelevant, they only steer the classification of pass 2.
#ifndef FLOF_APPLY_BLK
#define FLOF_APPLY_BLK 4
#endif
// STREAM: grad (read once) and tmp (written once) bypass the L2 residency competition with srch, which is read 9x
// Order in which the leaves are visited.  Large grids (a t-slice of srch beyond ~16 MB) are swept in z-CHUNKS: all
// t-slices of a block of z-planes, then the next block -- the t-1 / t+1 neighbours are then one chunk-slice (4 MB)
// away instead of one full slice (33.5 MB at 128^4, where the reuse distance with grad and tmp streaming through
// exceeded what L2 keeps and srch came from DRAM twice): 3.62 -> 2.45 ms at 128^4 (0.53 -> 0.80 of the HBM peak; chunks
// of 8 or 16 planes measured equal, 4 planes 2.96 ms, 32 planes 3.43 ms).  zc_leaves == 0: plain index order.
struct apply_order {
	int zc_leaves;  // leaves of one (z-chunk, t) block
	int lpt;        // leaves per t-slice
	int nT;         // t-slices of the range
};
__device__ __forceinline__ int apply_leaf(const apply_order &o, int q)
{
	if (o.zc_leaves == 0) return q;
	const int per_chunk = o.nT * o.zc_leaves;
	const int chunk = q / per_chunk, rem = q - chunk * per_chunk;
	const int t = rem / o.zc_leaves, l = rem - t * o.zc_leaves;
	return t * o.lpt + chunk * o.zc_leaves + l;
}
template <bool STREAM, int MINB, bool SEQ, int UNR>
__global__ void __launch_bounds__(FLOF_BLOCK, MINB)
    k_cg_apply(float4 *__restrict__ tmp, const float4 *__restrict__ srch, const float4 *__restrict__ grad,
               seq_part part, apply_order ord, int oY, int oZ, int oT, float offd, float diag, int multi, cg_seq sq,
               flof_p2p_dev pp, flof_reduce_scratch *red, flof_cg_state *st)
{
	if (st->done) return;
	__shared__ double shd[32];
	__shared__ int shi[8];
	__shared__ double s_ar[8];
	double dsum = 0.;
	// cells < 2^29 (checked by the caller): 32-bit cell indices keep the eight neighbour addresses out of registers
	// c walks the four 256-cell rows of a leaf, then jumps to this CTA's next leaf (one induction variable)
	// (the leaf loop is CTA-uniform, so that the full-mask shuffles below are safe when the range ends inside a warp)
	const int nleaf = (part.ncells + SEQ_LEAF_CELLS - 1) / SEQ_LEAF_CELLS;
	for (int lq = (int)blockIdx.x; lq < nleaf; lq += (int)gridDim.x) {
		const int c0l = apply_leaf(ord, lq) * SEQ_LEAF_CELLS;
		int c = c0l + (int)threadIdx.x;
		const int ce = min(c0l + SEQ_LEAF_CELLS, part.ncells);  // end of the leaf
		double lsum = 0.;
		float af = 0.f;
#pragma unroll UNR
		for (; c < ce; c += FLOF_BLOCK) {
			const float4 g = STREAM ? __ldcs(grad + c) : __ldg(grad + c);
			const float4 p = __ldg(srch + c);
			float4 v;
			if (is_border(g)) {
				v = p;  // identity row
			} else {
				v = make_float4(0.f, 0.f, 0.f, 0.f);
				if (offd != 0.f) {
					// neighbour order of nbx/nby/nbz/nbt (ref :389-392): t-1, z-1, y-1, x-1, x+1, y+1, z+1, t+1
					const int o[8] = { -oT, -oZ, -oY, -1, 1, oY, oZ, oT };
#pragma unroll
					for (int m = 0; m < 8; ++m) {
						const float4 q = __ldg(srch + (c + o[m]));
						v.x += offd * q.x; v.y += offd * q.y; v.z += offd * q.z; v.w += offd * q.w;
					}
				}
				// block row d: sum_m blockd(d,m) * x_m, blockd(d,m) = g_d*g_m (+ diag if d == m)  ref :483-488
				const float gg[4] = { g.x, g.y, g.z, g.w };
				const float pp[4] = { p.x, p.y, p.z, p.w };
				float vv[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
				for (int dd = 0; dd < 4; ++dd) {
#pragma unroll
					for (int m = 0; m < 4; ++m) {
						const float b = (dd == m) ? (gg[dd] * gg[m] + diag) : (gg[dd] * gg[m]);
						vv[dd] += b * pp[m];
					}
				}
				v = make_float4(vv[0], vv[1], vv[2], vv[3]);
			}
			if (STREAM)
				__stcs(tmp + c, v);
			else
				tmp[c] = v;
			if (SEQ)
				lsum += dot4a(p, v, af);
			else
				dsum += dot4(p, v);
		}
		if (SEQ) {
			// (every lane of the warp runs this: cells beyond the range contributed zero)
			double la = (double)af;
			for (int o = 16; o > 0; o >>= 1) {
				lsum += __shfl_down_sync(0xffffffffu, lsum, o);
				la += __shfl_down_sync(0xffffffffu, la, o);
			}
			if ((threadIdx.x & 31) == 0) {
				const int sg = (ce - 1) / part.seg_cells;
				atomicAdd(sq.A.aggx + sg, lsum);
				atomicAdd(sq.A.agga + sg, la);
			}
		}
	}
	if (!SEQ) {
		dsum = flof_block_sum(dsum, shd);
		if (threadIdx.x == 0) red->dsum[0][blockIdx.x] = dsum;
	}
	if (flof_last_block(&red->counter[1])) {
		if (SEQ) {  // alpha1 comes from the sequential-order dot product
			cg_seq_tail(sq, 0.f, false, multi, pp, shd, shi, s_ar);
			return;
		}
		double s = 0.;
		for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) s += red->dsum[0][b];
		s = flof_block_sum(s, shd);
		if (multi == 2) {  // dot(srch, A srch) over all ranks, exchanged through the peer mailboxes by this block
			if (threadIdx.x == 0) s_ar[0] = s;
			p2p_allreduce_block(pp, s_ar, 1, 1, false);
			s = s_ar[0];
		}
		if (threadIdx.x == 0) st->alpha1 = s;
	}
}

// B: alpha = sigma/alpha1; result += alpha*srch; res -= alpha*tmp; zv = res*precond; partial max(res), dot(zv, res)
// (zv is stored so that neither the sequential-order dot product nor the direction kernel repeats the divisions of the
// Jacobi preconditioner: 16 B/cell written here, 16 B/cell of grad not read there)
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_cg_update(float4 *__restrict__ x, float4 *__restrict__ res, float4 *__restrict__ zv, const float4 *__restrict__ srch,
                const float4 *__restrict__ tmp, const float4 *__restrict__ grad, cg_walk wk, float diag, int maxIter,
                int multi, cg_seq sq, flof_p2p_dev pp, flof_reduce_scratch *red, flof_cg_state *st)
{
	if (st->done) return;
	__shared__ double shd[32];
	__shared__ float shf[32];
	__shared__ int shi[8];
	__shared__ double s_ar[8];
	const double sigma = st->sigma[st->iter & 1];
	const double alpha = sigma / st->alpha1;  // ref :307-308
	const double nalpha = -alpha;
	double csum = 0.;
	float mx = -3.402823466e+38f;
	const int sg1 = min(((int)blockIdx.x + 1) * wk.spc, wk.part.nseg);
	for (int sg = (int)blockIdx.x * wk.spc; sg < sg1; ++sg) {
		double dsum = 0.;
		float af = 0.f;
		const int c0 = sg * wk.part.seg_cells, c1 = min(c0 + wk.part.seg_cells, wk.part.ncells);
		for (int c = c0 + (int)threadIdx.x; c < c1; c += FLOF_BLOCK) {
			const float4 p = __ldg(srch + c), ap = __ldg(tmp + c), g = __ldg(grad + c);
			float4 xv = x[c], r = res[c];
			xv.x = axpy1(xv.x, alpha, p.x); xv.y = axpy1(xv.y, alpha, p.y);
			xv.z = axpy1(xv.z, alpha, p.z); xv.w = axpy1(xv.w, alpha, p.w);
			r.x = axpy1(r.x, nalpha, ap.x); r.y = axpy1(r.y, nalpha, ap.y);
			r.z = axpy1(r.z, nalpha, ap.z); r.w = axpy1(r.w, nalpha, ap.w);
			x[c] = xv;
			res[c] = r;
			const float4 pc = precond_of(g, diag);
			const float4 z = make_float4(r.x * pc.x, r.y * pc.y, r.z * pc.z, r.w * pc.w);
			zv[c] = z;
			dsum += dot4a(z, r, af);
			mx = fmaxf(mx, fmaxf(fmaxf(r.x, r.y), fmaxf(r.z, r.w)));
		}
		double asum = (double)af;
		seq_block_sum2(dsum, asum, shd);
		if (threadIdx.x == 0) {
			csum += dsum;
			if (sq.on) {
				sq.A.aggx[sg] = dsum;
				sq.A.agga[sg] = asum;
			}
		}
	}
	mx = flof_block_max(mx, shf);
	if (threadIdx.x == 0) {
		red->dsum[1][blockIdx.x] = csum;
		red->fmax[blockIdx.x] = mx;
	}
	if (flof_last_block(&red->counter[2])) {
		double s = 0.;
		float m = -3.402823466e+38f;
		for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
			s += red->dsum[1][b];
			m = fmaxf(m, red->fmax[b]);
		}
		s = flof_block_sum(s, shd);
		m = flof_block_max(m, shf);
		if (sq.on) {
			// dot(tmp, res) comes from the sequential-order dot product, whose tail advances the state
			m = cg_seq_tail(sq, m, true, multi, pp, shd, shi, s_ar);
			if (threadIdx.x == 0) {
				st->sigmaNew = s;
				st->residual = m;
			}
			return;
		}
		if (multi == 2) {  // dot(z, res) summed and max(res) maximised over all ranks in one exchange
			if (threadIdx.x == 0) {
				s_ar[0] = s;
				s_ar[1] = (double)m;
			}
			p2p_allreduce_block(pp, s_ar, 2, 1, false);
			s = s_ar[0];
			m = (float)s_ar[1];
		}
		if (threadIdx.x == 0) {
			if (multi == 1) {
				// raw slab results; NCCL combines them, then k_cg_update_finalize advances the state
				st->sigmaNew = s;
				st->residual = m;
			} else {
				cg_advance(st, s, m, maxIter);
			}
		}
	}
}

// C: srch = res*precond + beta*srch (ref :318-323) with zv = res*precond stored by B; the stop test (ref :314-317)
// already ran in cg_advance
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_cg_direction(float4 *__restrict__ srch, const float4 *__restrict__ zv, int64_t cells, flof_cg_state *st)
{
	if (st->done) return;
	const int it = st->iter;  // iterations completed: sigma[it & 1] is the new sigma, the other slot the previous one
	const double beta = st->sigma[it & 1] / st->sigma[(it - 1) & 1];
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += stride) {
		const float4 z = __ldcs(zv + c);
		float4 p = srch[c];
		p.x = axpy1(z.x, beta, p.x); p.y = axpy1(z.y, beta, p.y);
		p.z = axpy1(z.z, beta, p.z); p.w = axpy1(z.w, beta, p.w);
		srch[c] = p;
	}
}

// ref :520-529 copy back: vel = result / mDx, optional rhsT = rhs[d=0]
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_of_copy_back(float4 *__restrict__ vel, const float4 *__restrict__ x, const float4 *__restrict__ rhs,
                   float *__restrict__ rhsT, int64_t cells, float mDx)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += stride) {
		const float4 v = __ldg(x + c);
		vel[c] = make_float4(v.x / mDx, v.y / mDx, v.z / mDx, v.w / mDx);
		if (rhsT) rhsT[c] = __ldg(rhs + c).x;
	}
}

// ---- host side of the sequential-order dot products ---------------------------------------------------------------
void flof_seq_release(flof_ctx *ctx)
{
	flof_seq *q = ctx->seq;
	if (!q) return;
	cudaFree(q->seg);
	cudaFree(q->cls);
	cudaFree(q->order);
	cudaFree(q->ent0);
	cudaFree(q->ent);
	cudaFree(q->ecnt);
	cudaFree(q->aggx);
	cudaFree(q->pool);
	cudaFree(q->gsteps);
	cudaFree(q->ctl);
	free(q);
	ctx->seq = NULL;
}
static int seq_ensure(flof_ctx *ctx)
{
	if (ctx->seq) return FLOF_OK;
	flof_seq *q = (flof_seq *)calloc(1, sizeof(flof_seq));
	if (!q) return flof_fail(ctx, FLOF_ERR_NOMEM, "flof_seq: out of host memory");
	ctx->seq = q;
	FLOF_CK(cudaMalloc((void **)&q->seg, sizeof(seq_seg) * SEQ_MAX_SEG));
	FLOF_CK(cudaMalloc((void **)&q->cls, sizeof(seq_cls) * SEQ_MAX_SEG));
	FLOF_CK(cudaMalloc((void **)&q->order, sizeof(int) * SEQ_MAX_SEG));
	FLOF_CK(cudaMalloc((void **)&q->ent0, sizeof(seq_rec) * (size_t)SEQ_MAX_SEG));
	FLOF_CK(cudaMalloc((void **)&q->ent, sizeof(seq_rec) * (size_t)SEQ_MAX_SEG * SEQ_ECAP));
	FLOF_CK(cudaMalloc((void **)&q->ecnt, sizeof(int) * SEQ_MAX_SEG));
	FLOF_CK(cudaMalloc((void **)&q->aggx, sizeof(double) * 2 * SEQ_MAX_SEG));
	FLOF_CK(cudaMemset(q->aggx, 0, sizeof(double) * 2 * SEQ_MAX_SEG));
	q->agga = q->aggx + SEQ_MAX_SEG;
	FLOF_CK(cudaMalloc((void **)&q->pool, sizeof(seq_rec) * (size_t)SEQ_POOL));
	FLOF_CK(cudaMalloc((void **)&q->gsteps, sizeof(seq_rec) * (size_t)SEQ_GSTEPS));
	FLOF_CK(cudaMalloc((void **)&q->ctl, sizeof(seq_ctl)));
	FLOF_CK(cudaMemset(q->ctl, 0, sizeof(seq_ctl)));
	FLOF_CK(cudaFuncSetAttribute(k_seq_resolve<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEQ_RESOLVE_SMEM));
	FLOF_CK(cudaFuncSetAttribute(k_seq_resolve<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEQ_RESOLVE_SMEM));
	FLOF_CK(cudaFuncSetAttribute(k_dot_seq<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEQ_DOT_SMEM));
	FLOF_CK(cudaFuncSetAttribute(k_dot_seq<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEQ_DOT_SMEM));
	return FLOF_OK;
}
static seq_args seq_make_args(flof_ctx *ctx, seq_part part, int64_t products_before)
{
	flof_seq *q = ctx->seq;
	seq_args A;
	A.seg = q->seg; A.cls = q->cls; A.order = q->order; A.aggx = q->aggx; A.agga = q->agga;
	A.ent0 = q->ent0; A.ent = q->ent; A.ecnt = q->ecnt; A.pool = q->pool; A.gsteps = q->gsteps; A.ctl = q->ctl;
	A.part = part;
	A.n0 = products_before;
	return A;
}
// pass 2 + resolve of one dot product whose pass 1 (per-segment prefixes in A.seg) has been enqueued
static int seq_dot(flof_ctx *ctx, int kind, const float4 *a, const float4 *b, float diag, const seq_args &A, int mode,
                   float accuracy, int maxIter, flof_cg_state *st, int multi)
{
	// segments are handed out by ticket, careful ones first: one resident wave of CTAs
	const int cap = ctx->sm_count * 4;
	const int blocks = A.part.nseg < cap ? A.part.nseg : cap;
	const flof_p2p_dev pp = ctx->p2p.dev;
	if (kind == 0) {
		FLOF_LAUNCH(k_dot_seq<0>, blocks, FLOF_BLOCK, SEQ_DOT_SMEM, a, b, diag, A, st);
		FLOF_LAUNCH(k_seq_resolve<0>, 1, FLOF_BLOCK, SEQ_RESOLVE_SMEM, a, b, diag, A, mode, accuracy, maxIter, st, multi, pp);
	} else {
		FLOF_LAUNCH(k_dot_seq<1>, blocks, FLOF_BLOCK, SEQ_DOT_SMEM, a, b, diag, A, st);
		FLOF_LAUNCH(k_seq_resolve<1>, 1, FLOF_BLOCK, SEQ_RESOLVE_SMEM, a, b, diag, A, mode, accuracy, maxIter, st, multi, pp);
	}
	return FLOF_OK;
}

static int seq_read_stats(flof_ctx *ctx, double *result, unsigned long long *stats, const char *who)
{
	seq_ctl *h = (seq_ctl *)malloc(sizeof(seq_ctl));
	if (!h) return flof_fail(ctx, FLOF_ERR_NOMEM, "%s: out of host memory", who);
	cudaError_t e = cudaMemcpyAsync(h, ctx->seq->ctl, sizeof(seq_ctl), cudaMemcpyDeviceToHost, ctx->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	if (e == cudaSuccess) {
		if (result) *result = h->result;
		if (stats) {
			stats[0] = h->n_dots; stats[1] = h->n_dirty; stats[2] = h->n_raw;
			stats[3] = h->n_pieces; stats[4] = h->n_fallback; stats[5] = h->n_inconsistent;
			stats[6] = h->n_slow_segments;
			stats[7] = h->n_inexact;
			stats[8] = h->why;
			stats[9] = h->n_rawleaves;
			for (int i = 0; i < 5; ++i) stats[10 + i] = h->prof[i];
		}
	}
	free(h);
	if (e != cudaSuccess) return flof_fail(ctx, FLOF_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e));
	return FLOF_OK;
}
// test / tool entry: the sequential-order sum of a[i]*b[i] (kind 0) or (a[i]*precond(b)[i])*a[i] (kind 1) over
// `cells` Vec4 cells.  stats (optional, 15 values; [10..14]: resolver cycles gather / compose / walk / finish and walk steps; [8]: OR of the reason flags of all fallbacks, [9]: leaves kept as plain products): dots, dirty leaves, raw products, pieces, fallbacks, inconsistencies,
// careful segments, inexact (tree-sum) fallbacks since the context was created.
extern "C" int flof_dot_seq(flof_ctx *ctx, const float *a, const float *b, int64_t cells, int kind, float diag,
                            double *result, unsigned long long *stats)
{
	FLOF_ARG(kind == 0 || kind == 1, "flof_dot_seq: kind must be 0 or 1");
	FLOF_ARG(cells > 0 && cells < ((int64_t)1 << 29), "flof_dot_seq: cell count out of range");
	FLOF_RET(seq_ensure(ctx));
	const seq_args A = seq_make_args(ctx, seq_make_part(cells, 0, ctx->sm_count), 0);
	if (kind == 0)
		FLOF_LAUNCH(k_seq_agg<0>, A.part.nseg, FLOF_BLOCK, 0, (const float4 *)a, (const float4 *)b, diag, A, ctx->red);
	else
		FLOF_LAUNCH(k_seq_agg<1>, A.part.nseg, FLOF_BLOCK, 0, (const float4 *)a, (const float4 *)b, diag, A, ctx->red);
	FLOF_RET(seq_dot(ctx, kind, (const float4 *)a, (const float4 *)b, diag, A, SEQ_MODE_NONE, 0.f, 0, NULL, 0));
	return seq_read_stats(ctx, result, stats, "flof_dot_seq");
}
extern "C" int flof_seq_stats(flof_ctx *ctx, unsigned long long *stats)
{
	FLOF_ARG(stats != NULL, "flof_seq_stats: stats is NULL");
	for (int i = 0; i < 15; ++i) stats[i] = 0;
	if (!ctx->seq) return FLOF_OK;
	return seq_read_stats(ctx, NULL, stats, "flof_seq_stats");
}

static int cg_run(flof_ctx *ctx, float *x, float *res, float *srch, float *tmp, float *zvec, const float *grad,
                  const float *rhs, flof_dim4 d, float wSmooth, float wEnergy, float accuracy,
                  int maxIter, int *iters, float *relRes, int *status, int dim = 4)
{
	const int64_t cells = flof_cells(d);
	const flof_of_consts k = of_consts(wSmooth, wEnergy, dim);
	const int64_t sY = d.nx, sZ = (int64_t)d.nx * d.ny, sT = sZ * d.nz;
	// sharded level: this rank iterates over its t-slab only; srch needs one ghost slice per side
	// before every apply, the three scalars of an iteration are all-reduced over the ranks
	int64_t c0, c1;
	flof_flat_range(ctx, cells, &c0, &c1);
	int multi = (c1 - c0 != cells && ctx->nranks > 1) ? 1 : 0;  // 1: scalars combined by NCCL, 2: inside the kernels (peer mailboxes)
	if (multi) {
		FLOF_RET(flof_p2p_ensure(ctx, sizeof(float4) * (size_t)sT));  // before the mailbox addresses are captured below
		if (ctx->p2p.enabled) multi = 2;
	}
	const flof_p2p_dev pp = ctx->p2p.dev;
	const int64_t n = c1 - c0;
	FLOF_ARG(n < ((int64_t)1 << 29), "opticalFlow4d: too many cells per rank");
	// one CTA per contiguous segment of the range (reducing kernels); the direction kernel stays a flat grid-stride loop
	const int apply_variant = ctx->opt.apply_variant;
	const seq_part part = seq_make_part(n, sT, ctx->sm_count);
	// streaming producers: consecutive segments per CTA, about 8 CTAs per SM
	cg_walk wk;
	wk.part = part;
	wk.spc = (part.nseg + ctx->sm_count * 8 - 1) / (ctx->sm_count * 8);
	const int blocks = (part.nseg + wk.spc - 1) / wk.spc;
	// the stencil kernel sweeps the grid leaf by leaf (grid-stride), at the occupancy of its variant
	const int64_t nleaf = (n + SEQ_LEAF_CELLS - 1) / SEQ_LEAF_CELLS;
	const int aper = apply_variant == 0 || apply_variant == 1 ? 4 : (apply_variant == 3 ? 5 : (apply_variant == 7 || apply_variant == 10 ? 8 : 6));
	const int ablocks = (int)(nleaf < (int64_t)ctx->sm_count * aper ? nleaf : (int64_t)ctx->sm_count * aper);
	const int dblocks = flof_flat_blocks(ctx, n, 8);
	// z-chunked leaf order for large t-slices (see apply_order); "apply_zchunk" 0 switches it off, > 0 forces that many planes
	apply_order ord = { 0, 0, 0 };
	{
		const int64_t slice_b = sT * 16;
		int zc = ctx->opt.apply_zchunk;
		if (zc < 0) zc = slice_b >= ((int64_t)16 << 20) ? (int)(((int64_t)4 << 20) / (sZ * 16)) : 0;  // 4 MB chunk-slices
		while (zc > 1 && d.nz % zc) --zc;
		if (zc >= 1 && zc < d.nz && (sZ * zc) % SEQ_LEAF_CELLS == 0 && n % sT == 0) {
			ord.zc_leaves = (int)(sZ * zc / SEQ_LEAF_CELLS);
			ord.lpt = (int)(sT / SEQ_LEAF_CELLS);
			ord.nT = (int)(n / sT);
		}
	}
	float4 *X = (float4 *)x + c0, *R = (float4 *)res + c0, *P = (float4 *)srch + c0, *AP = (float4 *)tmp + c0;
	float4 *Z = (float4 *)zvec + c0;
	const float4 *G = (const float4 *)grad + c0, *B = (const float4 *)rhs + c0;
	const size_t slice_bytes = sizeof(float4) * (size_t)sT;
	// dot products in the reference's sequential order (default): the reducing kernels leave per-segment prefixes
	// (pass 1), k_dot_seq + k_seq_resolve deliver the exact sums and advance the state.  On a sharded level this needs
	// the peer mailboxes (the exact running sum travels from rank to rank); the NCCL fallback keeps tree sums.
	const int seq = (ctx->opt.dot_mode == 1 && multi != 1) ? 1 : 0;
	cg_seq sq;
	memset(&sq, 0, sizeof(sq));
	if (seq) {
		FLOF_RET(seq_ensure(ctx));
		sq.A = seq_make_args(ctx, part, 4 * c0);  // products of the lower ranks' slabs come first in the sum
		sq.on = 1;
	}
	const seq_args &SA = sq.A;
	FLOF_LAUNCH(k_cg_init, blocks, FLOF_BLOCK, 0, X, R, P, G, B, wk, k.diag, accuracy, multi, sq, pp, ctx->red, ctx->cg);
	// (srch = res*precond after the init: the first sigma is dot(srch, res))
	if (seq) FLOF_RET(seq_dot(ctx, 0, P, R, k.diag, SA, SEQ_MODE_INIT, accuracy, maxIter, ctx->cg, multi));
	if (multi == 1) {
		FLOF_RET(flof_allreduce_f64_sum(ctx, &ctx->cg->sigmaNew, 1));
		FLOF_RET(flof_allreduce_f32_max(ctx, &ctx->cg->residual, 1));
		FLOF_LAUNCH(k_cg_init_finalize, 1, 1, 0, accuracy, ctx->cg);
	}
	flof_cg_state *h = (flof_cg_state *)ctx->pinned;
	// (exchanging the ghost slices of the new search direction on the side stream while k_cg_direction still updates the
	// interior slices was measured at 2 GPUs: bit-identical, no gain -- the push contends with the direction kernel for
	// the same HBM bandwidth: 178.8 vs 176.7-178.9 ms at 64^4 -- and is not kept)
	int launched = 0;
	// poll the device-side done flag every `chunk` iterations; iterations after convergence are
	// no-op kernels (early return on st->done)
	int chunk = 8;
	while (true) {
		FLOF_CK(cudaMemcpyAsync(h, ctx->cg, sizeof(flof_cg_state), cudaMemcpyDeviceToHost, ctx->stream));
		unsigned int *perr = (unsigned int *)((char *)ctx->pinned + 2048);
		*perr = 0;
		if (multi == 2) FLOF_CK(cudaMemcpyAsync(perr, ctx->p2p.dev.err, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
		FLOF_CK(cudaStreamSynchronize(ctx->stream));
		// a peer-mailbox wait that timed out leaves stale ghost slices / scalars: stop here instead of iterating on them
		if (*perr) return flof_fail(ctx, FLOF_ERR_CUDA, "opticalFlow4d: peer mailbox wait timed out during the CG (ranks out of step?)");
		if (h->done || launched >= maxIter) break;
		for (int q = 0; q < chunk && launched < maxIter; ++q, ++launched) {
			if (multi) FLOF_RET(flof_halo_exchange(ctx, srch, d.nt, slice_bytes, 1));
#define FLOF_APPLY_LAUNCH(STREAM, MINB, UNR)                                                                                  \
	do {                                                                                                                      \
		if (seq)                                                                                                              \
			FLOF_LAUNCH((k_cg_apply<STREAM, MINB, true, UNR>), ablocks, FLOF_BLOCK, 0, AP, (const float4 *)P, G, part, ord,   \
			            (int)sY, (int)sZ, (int)sT, k.offd, k.diag, multi, sq, pp, ctx->red, ctx->cg);                        \
		else                                                                                                                  \
			FLOF_LAUNCH((k_cg_apply<STREAM, MINB, false, UNR>), ablocks, FLOF_BLOCK, 0, AP, (const float4 *)P, G, part, ord,  \
			            (int)sY, (int)sZ, (int)sT, k.offd, k.diag, multi, sq, pp, ctx->red, ctx->cg);                        \
	} while (0)
			switch (apply_variant) {
			case 0: FLOF_APPLY_LAUNCH(false, 4, 4); break;
			case 1: FLOF_APPLY_LAUNCH(true, 4, 4); break;
			case 3: FLOF_APPLY_LAUNCH(true, 5, 4); break;
			case 5: FLOF_APPLY_LAUNCH(true, 6, 4); break;
			case 9: FLOF_APPLY_LAUNCH(true, 6, 2); break;
			case 10: FLOF_APPLY_LAUNCH(true, 8, 2); break;
			case 7: FLOF_APPLY_LAUNCH(true, 8, 1); break;
			default: FLOF_APPLY_LAUNCH(true, 6, 1); break;  // 11: 40 registers, no spills -- measured fastest (0.184 ms at 64^4)
			}
			if (multi == 1) FLOF_RET(flof_allreduce_f64_sum(ctx, &ctx->cg->alpha1, 1));
			if (seq) FLOF_RET(seq_dot(ctx, 0, P, AP, k.diag, SA, SEQ_MODE_ALPHA, accuracy, maxIter, ctx->cg, multi));
			FLOF_LAUNCH(k_cg_update, blocks, FLOF_BLOCK, 0, X, R, Z, (const float4 *)P, (const float4 *)AP, G, wk, k.diag, maxIter, multi,
			            sq, pp, ctx->red, ctx->cg);
			if (seq) FLOF_RET(seq_dot(ctx, 0, Z, R, k.diag, SA, SEQ_MODE_ADVANCE, accuracy, maxIter, ctx->cg, multi));
			if (multi == 1) {
				FLOF_RET(flof_allreduce_f64_sum(ctx, &ctx->cg->sigmaNew, 1));
				FLOF_RET(flof_allreduce_f32_max(ctx, &ctx->cg->residual, 1));
				FLOF_LAUNCH(k_cg_update_finalize, 1, 1, 0, maxIter, ctx->cg);
			}
			FLOF_LAUNCH(k_cg_direction, dblocks, FLOF_BLOCK, 0, P, (const float4 *)Z, n, ctx->cg);
		}
		if (launched >= 32) chunk = 16;
	}
	*iters = h->iter;
	*relRes = h->relResidual;
	*status = h->status;
	if (seq && h->seq_inexact)
		return flof_fail(ctx, FLOF_ERR_ARG, "opticalFlow4d: a sequential-order dot product exceeded its capacities (sum with heavy "
		                 "cancellation); the result would not be the reference's bits -- set dot_mode 0 for tree reductions");
	return FLOF_OK;
}

int flof_of_cg_dim(flof_ctx *ctx, float *x, const float *grad, const float *rhs, flof_dim4 d, float wSmooth, float wEnergy,
                   int dim, float accuracy, int maxIter, int *iters, float *relResidual);
extern "C" int flof_of_cg(flof_ctx *ctx, float *x, const float *grad, const float *rhs, flof_dim4 d,
                          float wSmooth, float wEnergy, float accuracy, int maxIter, int *iters,
                          float *relResidual)
{
	return flof_of_cg_dim(ctx, x, grad, rhs, d, wSmooth, wEnergy, 4, accuracy, maxIter, iters, relResidual);
}
// dim = 3: the diagonal constant of the DIM = 3 instantiation (ref :478-480) for an embedded 3D problem
int flof_of_cg_dim(flof_ctx *ctx, float *x, const float *grad, const float *rhs, flof_dim4 d, float wSmooth, float wEnergy,
                   int dim, float accuracy, int maxIter, int *iters, float *relResidual)
{
	const size_t vb = sizeof(float) * 4 * (size_t)flof_cells(d);
	void *res = NULL, *srch = NULL, *tmp = NULL, *zv = NULL;
	int rc = flof_tmp_alloc(ctx, &res, vb, false);
	if (rc == FLOF_OK) rc = flof_tmp_alloc(ctx, &srch, vb, false);
	if (rc == FLOF_OK) rc = flof_tmp_alloc(ctx, &tmp, vb, false);
	if (rc == FLOF_OK) rc = flof_tmp_alloc(ctx, &zv, vb, false);
	int st = 0, it = 0;
	float rr = 0.f;
	if (rc == FLOF_OK)
		rc = cg_run(ctx, x, (float *)res, (float *)srch, (float *)tmp, (float *)zv, grad, rhs, d, wSmooth, wEnergy,
		            accuracy, maxIter, &it, &rr, &st, dim);
	flof_tmp_free(ctx, res);
	flof_tmp_free(ctx, srch);
	flof_tmp_free(ctx, tmp);
	flof_tmp_free(ctx, zv);
	if (iters) *iters = it;
	if (relResidual) *relResidual = rr;
	return rc;
}

int flof_gaussian_blur4d_impl(flof_ctx *ctx, float *a, flof_dim4 d, int elem, float sigma, int iter);

// ctx-level timing hooks for the multi-scale trace (device time of the CG part of the last call)
float g_flof_last_cg_ms = 0.f;

extern "C" int flof_optical_flow4d(flof_ctx *ctx, float *vel, const float *i0, const float *i1,
                                   float *rhsT, flof_dim4 d, float wSmooth, float wEnergy,
                                   float postVelBlur, float cgAccuracy, float resetBndWidth,
                                   int *cgIters, float *cgRes)
{
	return flof_optical_flow4d_ex(ctx, vel, i0, i1, rhsT, d, wSmooth, wEnergy, postVelBlur, cgAccuracy,
	                              resetBndWidth, 0, cgIters, cgRes);
}

// vel_is_zero: the caller guarantees the incoming field is all zero (the multi-step driver solves
// for a fresh zero field every time, ref :1046, 1119).  Then every smoothness / Tikhonov rhs term
// is +-0 and `rhs -= 0` is the identity, so the assembly skips reading vel -- bit-identical.
int flof_optical_flow4d_ex(flof_ctx *ctx, float *vel, const float *i0, const float *i1, float *rhsT, flof_dim4 d,
                           float wSmooth, float wEnergy, float postVelBlur, float cgAccuracy,
                           float resetBndWidth, int vel_is_zero, int *cgIters, float *cgRes)
{
	FLOF_ARG(d.nx >= 3 && d.ny >= 3 && d.nz >= 3 && d.nt >= 3, "opticalFlow4d: grid too small");
	const int64_t cells = flof_cells(d);
	FLOF_ARG(cells < ((int64_t)1 << 29), "opticalFlow4d: N = cells*4 exceeds int range (ref :376)");
	const size_t vb = sizeof(float) * 4 * (size_t)cells;
	// (pool temporaries: a failed allocation releases the earlier ones below -- flof_tmp_free(NULL) is a no-op)
	void *grad = NULL, *rhs = NULL, *x = NULL, *res = NULL, *srch = NULL, *tmp = NULL, *zv = NULL;
	int rc = flof_tmp_alloc(ctx, &grad, vb, false);
	if (rc == FLOF_OK) rc = flof_tmp_alloc(ctx, &rhs, vb, false);
	if (rc == FLOF_OK) rc = flof_tmp_alloc(ctx, &x, vb, false);
	if (rc == FLOF_OK) rc = flof_tmp_alloc(ctx, &res, vb, false);
	if (rc == FLOF_OK) rc = flof_tmp_alloc(ctx, &srch, vb, false);
	if (rc == FLOF_OK) rc = flof_tmp_alloc(ctx, &tmp, vb, false);
	if (rc == FLOF_OK) rc = flof_tmp_alloc(ctx, &zv, vb, false);
	if (rc == FLOF_OK) rc = flof_of_assemble(ctx, (float *)grad, (float *)rhs, i0, i1, vel_is_zero ? NULL : vel, d, wSmooth, wEnergy);
	int it = 0, st = 0;
	float rr = 1e10f;
	if (rc == FLOF_OK) {
		cudaEventRecord(ctx->ev[2], ctx->stream);
		rc = cg_run(ctx, (float *)x, (float *)res, (float *)srch, (float *)tmp, (float *)zv, (const float *)grad,
		            (const float *)rhs, d, wSmooth, wEnergy, cgAccuracy, 1000, &it, &rr, &st);
		cudaEventRecord(ctx->ev[3], ctx->stream);
	}
	if (rc == FLOF_OK) {
		if (rr != rr) {  // ref :509-514 NaN -> zero the solution
			rc = flof_memset0(ctx, x, vb);
		}
	}
	if (rc == FLOF_OK) {
		const float mDx = (float)(1. / d.nx);
		int64_t c0, c1;
		flof_flat_range(ctx, cells, &c0, &c1);
		FLOF_LAUNCH(k_of_copy_back, flof_flat_blocks(ctx, c1 - c0, 8), FLOF_BLOCK, 0, (float4 *)vel + c0,
		            (const float4 *)x + c0, (const float4 *)rhs + c0, rhsT ? rhsT + c0 : NULL, c1 - c0, mDx);
	}
	flof_tmp_free(ctx, grad);
	flof_tmp_free(ctx, rhs);
	flof_tmp_free(ctx, x);
	flof_tmp_free(ctx, res);
	flof_tmp_free(ctx, srch);
	flof_tmp_free(ctx, tmp);
	flof_tmp_free(ctx, zv);
	if (rc != FLOF_OK) return rc;
	// ref :532-541 optional blur with half sigma
	if (postVelBlur > 0.f) FLOF_RET(flof_gaussian_blur4d_impl(ctx, vel, d, 4, (float)(0.5 * postVelBlur), 1));
	// ref :544-551 reset outer border
	if (resetBndWidth > 0.f) {
		const int resetBnd = (int)(resetBndWidth * d.nx) + 1;
		FLOF_RET(flof_reset_border_vec4(ctx, vel, d, resetBnd));
	}
	FLOF_CK(cudaEventSynchronize(ctx->ev[3]));
	cudaEventElapsedTime(&g_flof_last_cg_ms, ctx->ev[2], ctx->ev[3]);
	if (cgIters) *cgIters = it;
	if (cgRes) *cgRes = rr;
	return FLOF_OK;
}
