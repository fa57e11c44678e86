// flof_seqsum_kernels.cuh -- the CG dot products in the reference's sequential summation order, evaluated in parallel.
// ref: dotProd optflow4d.cpp:234-241 (call sites :296, :307, :319).  Arithmetic: flof_seqsum_core.h; organisation and
// data structures: flof_seqsum.cuh.  (included by flof_solve.cu after precond_of / cg_advance / cg_init_finalize)
//
// k_seq_agg<KIND>   stand-alone pass 1 (tests, tools): per-segment approximate sums + the tail scan.  Inside the CG the
//                   producing kernels do this on the fly (seq_tail_scan is their shared tail).
// k_dot_seq<KIND>   pass 2, one read of the two vectors (32 B/cell).
//                     safe segment (running sum provably inside one binade): eight independent warps; a warp folds
//                       128 cells per step -- coalesced loads, products through a warp-private shared-memory tile so that
//                       every lane gets 16 CONSECUTIVE products, per-lane rounding functions, ordered composition over
//                       the lanes by shuffles -- and only the eight warp results meet at the end of the segment.
//                     careful segment (around a binade crossing, or while the sum builds up from zero): leaf by leaf with
//                       a running prefix; clean leaves are folded by the whole CTA, dirty leaves are cut into function
//                       runs + raw products appended to the piece pool.
// k_seq_resolve     one CTA: composes the segment entries between dirty leaves in parallel, then one thread walks runs
//                   and raw products with real fp64 adds.  The result has the bits of the reference's loop; the tail
//                   feeds it to the CG state (alpha1 / cg_advance / cg_init_finalize).  If a capacity is exceeded or a
//                   consistency check fails it falls back to the plain loop on one thread (counted in
//                   seq_ctl::n_fallback; tests assert it stays 0).
//   KIND 0: sum_i a[i]*b[i]                         (srch . A srch)
//   KIND 1: sum_i (a[i]*precond(b)[i]) * a[i]       (a = res, b = grad: tmp = res*precond; dot(tmp, res))
#pragma once
#include "flof_seqsum.cuh"

struct seq_args {
	seq_seg *seg;
	seq_cls *cls;
	int *order;
	double *aggx, *agga;
	seq_rec *ent0;  // [SEQ_MAX_SEG] first entry of every segment (dense: the resolver gathers it with coalesced loads)
	seq_rec *ent;   // [SEQ_MAX_SEG * SEQ_ECAP] further entries of a segment (careful segments only)
	int *ecnt;
	seq_rec *pool;
	seq_rec *gsteps;
	seq_ctl *ctl;
	seq_part part;
	int64_t n0;  // products of the lower ranks' ranges (0 on a single GPU): global position of this range in the sum
};
// margin factor for everything up to the end of segment `seg` (the bound grows with the number of terms summed so far)
__device__ __forceinline__ double seq_kf_of(const seq_args &A, int seg)
{
	return seq_margin_factor(A.n0 + 4ll * min((int64_t)(seg + 1) * A.part.seg_cells, (int64_t)A.part.ncells));
}

#define SEQ_MODE_NONE 0     // result only (seq_ctl::result)
#define SEQ_MODE_ALPHA 1    // st->alpha1 = dot(srch, A srch)                  ref :307
#define SEQ_MODE_ADVANCE 2  // cg_advance(st, dot(tmp, res), st->residual)     ref :311-324
#define SEQ_MODE_INIT 3     // cg_init_finalize(st, dot(tmp, res), st->residual) ref :286-301

#define SEQ_LEAF_WILD 0
#define SEQ_LEAF_CLEAN 1
#define SEQ_LEAF_DIRTY 2
#define SEQ_XS 20  // floats per lane tile in shared memory: 16 products + 4 pad (conflict-free LDS.128 at 80 B lane stride)

template <int KIND>
__device__ __forceinline__ float4 seq_products(const float4 *__restrict__ a, const float4 *__restrict__ b, int c, float diag)
{
	const float4 p = __ldg(a + c), q = __ldg(b + c);
	if (KIND == 0) return make_float4(p.x * q.x, p.y * q.y, p.z * q.z, p.w * q.w);
	const float4 pc = precond_of(q, diag);
	const float4 z = make_float4(p.x * pc.x, p.y * pc.y, p.z * pc.z, p.w * pc.w);
	return make_float4(z.x * p.x, z.y * p.y, z.z * p.z, z.w * p.w);
}

__device__ __forceinline__ seq_fn seq_shfl_down(const seq_fn &f, int o)
{
	seq_fn g;
	g.d0 = __shfl_down_sync(0xffffffffu, f.d0, o);
	g.d1 = __shfl_down_sync(0xffffffffu, f.d1, o);
	g.q = __shfl_down_sync(0xffffffffu, f.q, o);
	return g;
}
// ordered composition over the lanes of a warp (lane 0 first); result valid in lane 0
__device__ __forceinline__ seq_fn seq_warp_compose(seq_fn f, int n)
{
	const int lane = threadIdx.x & 31;
	for (int o = 1; o < n; o <<= 1) {
		const seq_fn g = seq_shfl_down(f, o);
		if (lane + o < n) f = seq_compose(f, g);
	}
	return f;
}
// sixteen consecutive products (four float4 of one lane tile) -> one rounding function
__device__ __forceinline__ seq_fn seq_fold16(const float4 *xp, double C0, double C1)
{
	seq_fn f = seq_identity();
#pragma unroll
	for (int k = 0; k < SEQ_U; ++k) {
		const float4 v = xp[k];
		f = seq_compose(f, seq_elem((double)v.x, C0, C1));
		f = seq_compose(f, seq_elem((double)v.y, C0, C1));
		f = seq_compose(f, seq_elem((double)v.z, C0, C1));
		f = seq_compose(f, seq_elem((double)v.w, C0, C1));
	}
	return f;
}
// exclusive scan over the threads of the CTA (thread order); totals returned to every thread.  sh: >= 2 * 8 doubles
__device__ __forceinline__ void seq_block_exscan2(double &x, double &y, double *sh, double &totx, double &toty)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	double ix = x, iy = y;
	for (int o = 1; o < 32; o <<= 1) {
		const double ux = __shfl_up_sync(0xffffffffu, ix, o), uy = __shfl_up_sync(0xffffffffu, iy, o);
		if (lane >= o) {
			ix += ux;
			iy += uy;
		}
	}
	__syncthreads();
	if (lane == 31) {
		sh[wid] = ix;
		sh[8 + wid] = iy;
	}
	__syncthreads();
	double bx = 0., by = 0.;
	totx = 0.;
	toty = 0.;
	for (int w = 0; w < FLOF_BLOCK / 32; ++w) {
		if (w < wid) {
			bx += sh[w];
			by += sh[8 + w];
		}
		totx += sh[w];
		toty += sh[8 + w];
	}
	x = bx + (ix - x);
	y = by + (iy - y);
}
__device__ __forceinline__ int seq_block_exscan_int(int v, int *sh, int &tot)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	int iv = v;
	for (int o = 1; o < 32; o <<= 1) {
		const int u = __shfl_up_sync(0xffffffffu, iv, o);
		if (lane >= o) iv += u;
	}
	__syncthreads();
	if (lane == 31) sh[wid] = iv;
	__syncthreads();
	int b = 0;
	tot = 0;
	for (int w = 0; w < FLOF_BLOCK / 32; ++w) {
		if (w < wid) b += sh[w];
		tot += sh[w];
	}
	return b + iv - v;
}
// both sums of a CTA at once; result valid in thread 0.  sh: >= 16 doubles
__device__ __forceinline__ void seq_block_sum2(double &x, double &y, double *sh)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	for (int o = 16; o > 0; o >>= 1) {
		x += __shfl_down_sync(0xffffffffu, x, o);
		y += __shfl_down_sync(0xffffffffu, y, o);
	}
	__syncthreads();
	if (lane == 0) {
		sh[wid] = x;
		sh[8 + wid] = y;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		x = 0.;
		y = 0.;
		for (int w = 0; w < FLOF_BLOCK / 32; ++w) {
			x += sh[w];
			y += sh[8 + w];
		}
	}
}

// ---- pass 1 tail (ONE block, every thread): per-segment sums -> exclusive prefixes within this rank's range.
// aggx[b] / agga[b]: approximate sum of the products of segment b / fp32-accumulated sum of their magnitudes; both are
// zeroed for the next launch.  totals (every thread): tx, ta (ta already carries SEQ_SA_SLACK).
__device__ __forceinline__ void seq_tail_scan(const seq_args &A, double *sh, double &tx, double &ta)
{
	const int nseg = A.part.nseg;
	const int per = (nseg + FLOF_BLOCK - 1) / FLOF_BLOCK;  // <= 8
	const int b0 = min((int)threadIdx.x * per, nseg), b1 = min(b0 + per, nseg);
	double lx = 0., la = 0.;
	for (int b = b0; b < b1; ++b) {
		lx += A.aggx[b];
		la += A.agga[b];
	}
	double ex = lx, ea = la;
	seq_block_exscan2(ex, ea, sh, tx, ta);
	for (int b = b0; b < b1; ++b) {
		const double vx = A.aggx[b], va = A.agga[b];
		seq_seg s;
		s.sx = vx;
		s.sa = va * SEQ_SA_SLACK;
		s.px = ex;
		s.pa = ea * SEQ_SA_SLACK;
		A.seg[b] = s;
		ex += vx;
		ea += va;
		A.aggx[b] = 0.;
		A.agga[b] = 0.;
	}
	ta *= SEQ_SA_SLACK;
}
// ---- then (the lower ranks' share offx / offa known): classify every segment and build the work list of pass 2,
// careful segments first.  shi: >= 8 ints.
__device__ __forceinline__ void seq_tail_classify(const seq_args &A, double offx, double offa, int *shi)
{
	const int nseg = A.part.nseg;
	const int per = (nseg + FLOF_BLOCK - 1) / FLOF_BLOCK;
	const int b0 = min((int)threadIdx.x * per, nseg), b1 = min(b0 + per, nseg);
	int ncare = 0;
	unsigned bits = 0;
	for (int b = b0; b < b1; ++b) {
		const seq_seg s = A.seg[b];
		const double P = offx + s.px, T = offa + s.pa;
		seq_cls c;
		c.e = 0;
		if (!seq_finite(s.sx) || !seq_finite(s.sa) || !seq_finite(P) || !seq_finite(T)) {
			atomicOr(&A.ctl->flags, 1u);
			c.mode = SEQ_LEAF_WILD;
		} else if (s.sa == 0.)
			c.mode = SEQ_LEAF_WILD;
		else
			c.mode = seq_range_safe(P, T, s.sa, seq_kf_of(A, b), &c.e) ? SEQ_LEAF_CLEAN : SEQ_LEAF_DIRTY;
		A.cls[b] = c;
		if (c.mode == SEQ_LEAF_DIRTY) {
			++ncare;
			bits |= 1u << (b - b0);
		}
	}
	int total;
	int before = seq_block_exscan_int(ncare, shi, total);
	for (int b = b0; b < b1; ++b) {
		if ((bits >> (b - b0)) & 1u)
			A.order[before++] = b;
		else
			A.order[total + b - before] = b;  // safe / empty segments keep their order behind the careful ones
	}
}

// stand-alone pass 1: one CTA per segment
template <int KIND>
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_seq_agg(const float4 *__restrict__ a, const float4 *__restrict__ b, float diag, seq_args A, flof_reduce_scratch *red)
{
	__shared__ double shd[32];
	__shared__ int shi[8];
	const int seg = blockIdx.x;
	const int c0 = seg * A.part.seg_cells, c1 = min(c0 + A.part.seg_cells, A.part.ncells);
	double sx = 0.;
	float af = 0.f;
	for (int c = c0 + (int)threadIdx.x; c < c1; c += FLOF_BLOCK) {
		const float4 p = seq_products<KIND>(a, b, c, diag);
		sx += (double)p.x + (double)p.y + (double)p.z + (double)p.w;
		af += fabsf(p.x) + fabsf(p.y) + fabsf(p.z) + fabsf(p.w);
	}
	double sa = (double)af;
	seq_block_sum2(sx, sa, shd);
	if (threadIdx.x == 0) {
		A.aggx[seg] = sx;
		A.agga[seg] = sa;
	}
	if (flof_last_block(&red->counter[1])) {
		double tx, ta;
		seq_tail_scan(A, shd, tx, ta);
		__syncthreads();
		seq_tail_classify(A, 0., 0., shi);
		if (threadIdx.x == 0) {
			A.ctl->tot[0] = tx;
			A.ctl->tot[1] = ta;
			A.ctl->off[0] = 0.;
			A.ctl->off[1] = 0.;
		}
	}
}

struct seq_builder {  // consecutive safe products of one binade merge into one run; everything else becomes raw
	seq_rec *out;
	int n;
	bool have;
	int e;
	seq_fn f;
	__device__ __forceinline__ void flush()
	{
		if (!have) return;
		seq_rec r;
		r.d0 = f.d0; r.d1 = f.d1; r.e = e; r.q = f.q; r.pad[0] = r.pad[1] = 0;
		out[n++] = r;
		have = false;
	}
	__device__ __forceinline__ void push(double x, double P, double T, double kf)
	{
		if (x == 0.) return;
		int pe;
		if (seq_range_safe(P, T, seq_abs(x), kf, &pe)) {
			double C0, C1;
			seq_consts(pe, &C0, &C1);
			const seq_fn g = seq_elem(x, C0, C1);
			if (have && e == pe)
				f = seq_compose(f, g);
			else {
				flush();
				have = true;
				e = pe;
				f = g;
			}
		} else {
			flush();
			seq_rec r;
			r.d0 = x; r.d1 = x; r.e = SEQ_E_RAW; r.q = 0; r.pad[0] = r.pad[1] = 0;  // (d1 = d0: the walk adds by parity)
			out[n++] = r;
		}
	}
};

// entry list of one segment, kept by thread 0: consecutive clean runs of one binade merge into one entry
struct seq_emitter {
	seq_rec *out0;  // A.ent0 + seg: first entry
	seq_rec *out;   // A.ent + seg * SEQ_ECAP: the others
	int n;
	bool have;
	int e;
	seq_fn f;
	unsigned int *flags;
	__device__ __forceinline__ void put(const seq_rec &r)
	{
		if (n == 0)
			*out0 = r;
		else if (n <= SEQ_ECAP)
			out[n - 1] = r;
		else
			atomicOr(flags, 2u | 0x100u);   // segment entry list full
		++n;
	}
	__device__ __forceinline__ void flush()
	{
		if (!have) return;
		seq_rec r;
		r.d0 = f.d0; r.d1 = f.d1; r.e = e; r.q = f.q; r.pad[0] = r.pad[1] = 0;
		put(r);
		have = false;
	}
	__device__ __forceinline__ void run(const seq_fn &g, int ge)
	{
		if (have && e == ge)
			f = seq_compose(f, g);
		else {
			flush();
			have = true;
			e = ge;
			f = g;
		}
	}
};

template <int KIND>
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_dot_seq(const float4 *__restrict__ a, const float4 *__restrict__ b, float diag, seq_args A, const flof_cg_state *st)
{
	if (st && st->done) return;
	extern __shared__ __align__(16) float s_x[];  // SEQ_DOT_SMEM: eight warp tiles / one leaf (20 KB) / piece staging (48 KB)
	__shared__ double shd[32];
	__shared__ seq_fn s_fn[FLOF_BLOCK / 32];
	__shared__ int shi[8], s_wb[8];
	__shared__ int s_mode, s_e, s_seg, s_rawbase;
	__shared__ double s_P, s_T;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	for (;;) {
		// work list of the tail of pass 1: careful segments first, handed out by ticket
		if (tid == 0) s_seg = (int)atomicAdd(&A.ctl->ticket, 1u);
		__syncthreads();
		if (s_seg >= A.part.nseg) break;
		const int seg = A.order[s_seg];
		const int c0 = seg * A.part.seg_cells, c1 = min(c0 + A.part.seg_cells, A.part.ncells);
		seq_emitter em;
		em.out0 = A.ent0 + seg; em.out = A.ent + (size_t)seg * SEQ_ECAP; em.n = 0; em.have = false; em.e = 0; em.f = seq_identity();
		em.flags = &A.ctl->flags;
		double P = 0., T = 0.;  // running approximate prefix (thread 0)
		if (tid == 0) {
			const seq_seg s = A.seg[seg];
			const seq_cls c = A.cls[seg];
			P = A.ctl->off[0] + s.px;
			T = A.ctl->off[1] + s.pa;
			s_mode = c.mode;
			s_e = c.e;
		}
		__syncthreads();
		const int smode = s_mode;
		if (smode == SEQ_LEAF_CLEAN) {
			// ---- safe segment: eight independent warps, 128 cells per step
			double C0, C1;
			seq_consts(s_e, &C0, &C1);
			const int wc = A.part.seg_cells / (FLOF_BLOCK / 32);
			const int w0 = c0 + wid * wc, w1 = min(w0 + wc, c1);
			float *tile = s_x + wid * (32 * SEQ_XS);
			const float4 *xp = reinterpret_cast<const float4 *>(tile + lane * SEQ_XS);
			seq_fn F = seq_identity();
			for (int ch = w0; ch < w1; ch += SEQ_CHUNK_CELLS) {
#pragma unroll
				for (int i = 0; i < SEQ_U; ++i) {
					const int ci = i * 32 + lane, c = ch + ci;
					float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
					if (c < w1) p = seq_products<KIND>(a, b, c, diag);
					*reinterpret_cast<float4 *>(tile + (ci >> 2) * SEQ_XS + (ci & 3) * 4) = p;
				}
				__syncwarp();
				seq_fn g = seq_fold16(xp, C0, C1);
				__syncwarp();
				g = seq_warp_compose(g, 32);
				F = seq_compose(F, g);  // meaningful in lane 0
			}
			if (lane == 0) s_fn[wid] = F;
			__syncthreads();
			if (wid == 0) {
				seq_fn f = lane < FLOF_BLOCK / 32 ? s_fn[lane] : seq_identity();
				f = seq_warp_compose(f, FLOF_BLOCK / 32);
				if (lane == 0) em.run(f, s_e);
			}
		} else if (smode == SEQ_LEAF_DIRTY) {
			// ---- careful segment: leaf by leaf with a running prefix
			if (tid == 0) atomicAdd(&A.ctl->n_slow_segments, 1ull);
			const double kf = seq_kf_of(A, seg);
			for (int l0 = c0; l0 < c1; l0 += SEQ_LEAF_CELLS) {
				double sx = 0., sa = 0.;
#pragma unroll
				for (int s = 0; s < SEQ_U; ++s) {
					const int ci = s * FLOF_BLOCK + tid, c = l0 + ci;
					float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
					if (c < c1) p = seq_products<KIND>(a, b, c, diag);
					*reinterpret_cast<float4 *>(&s_x[(ci >> 2) * SEQ_XS + (ci & 3) * 4]) = p;
					sx += (double)p.x + (double)p.y + (double)p.z + (double)p.w;  // approximate: any order will do
					sa += (double)fabsf(p.x) + (double)fabsf(p.y) + (double)fabsf(p.z) + (double)fabsf(p.w);
				}
				seq_block_sum2(sx, sa, shd);
				if (tid == 0) {
					int mode, e = 0;
					if (!seq_finite(sx) || !seq_finite(sa)) {
						atomicOr(&A.ctl->flags, 1u);
						mode = SEQ_LEAF_WILD;
					} else if (sa == 0.)
						mode = SEQ_LEAF_WILD;
					else
						mode = seq_range_safe(P, T, sa, kf, &e) ? SEQ_LEAF_CLEAN : SEQ_LEAF_DIRTY;
					s_mode = mode;
					s_e = e;
					s_P = P;
					s_T = T;
					P += sx;
					T += sa;
				}
				__syncthreads();
				const int mode = s_mode;
				const float4 *xp = reinterpret_cast<const float4 *>(&s_x[tid * SEQ_XS]);
				if (mode == SEQ_LEAF_CLEAN) {
					double C0, C1;
					seq_consts(s_e, &C0, &C1);
					seq_fn f = seq_fold16(xp, C0, C1);
					f = seq_warp_compose(f, 32);
					if (lane == 0) s_fn[wid] = f;
					__syncthreads();
					if (wid == 0) {
						f = lane < FLOF_BLOCK / 32 ? s_fn[lane] : seq_identity();
						f = seq_warp_compose(f, FLOF_BLOCK / 32);
						if (lane == 0) em.run(f, s_e);
					}
				} else if (mode == SEQ_LEAF_DIRTY) {
					float xs[4 * SEQ_U];
#pragma unroll
					for (int k = 0; k < SEQ_U; ++k) {
						const float4 v = xp[k];
						xs[4 * k] = v.x; xs[4 * k + 1] = v.y; xs[4 * k + 2] = v.z; xs[4 * k + 3] = v.w;
					}
					double px = 0., pa = 0., totx, tota;
#pragma unroll
					for (int k = 0; k < 4 * SEQ_U; ++k) {
						px += (double)xs[k];
						pa += (double)fabsf(xs[k]);
					}
					seq_block_exscan2(px, pa, shd, totx, tota);  // (its barriers also end every thread's reads of s_x)
					seq_rec pc[4 * SEQ_U];
					seq_builder bd;
					bd.out = pc; bd.n = 0; bd.have = false; bd.e = 0; bd.f = seq_identity();
					double Pk = s_P + px, Tk = s_T + pa;
#pragma unroll 1
					for (int k = 0; k < 4 * SEQ_U; ++k) {
						const double x = (double)xs[k];
						bd.push(x, Pk, Tk, kf);
						Pk += x;
						Tk += seq_abs(x);
					}
					bd.flush();
					int total;
					const int off = seq_block_exscan_int(bd.n, shi, total);
					seq_rec *stage = reinterpret_cast<seq_rec *>(s_x);
					const int cap = SEQ_STAGE;
					if (total <= cap)
						for (int k = 0; k < bd.n; ++k) stage[off + k] = pc[k];
					__syncthreads();
					// merge neighbouring runs of one binade (pieces of different threads): lane 0 of every warp compacts
					// its warp's pieces in place, then thread 0 joins the eight compacted ranges
					if (total <= cap) {
						const int wbeg = __shfl_sync(0xffffffffu, off, 0), wend = __shfl_sync(0xffffffffu, off + bd.n, 31);
						if (lane == 0) {
							int m = wbeg;
							for (int k = wbeg; k < wend; ++k) {
								const seq_rec q = stage[k];
								if (m > wbeg && q.e > SEQ_E_WILD && stage[m - 1].e == q.e) {
									seq_fn f = { stage[m - 1].d0, stage[m - 1].d1, stage[m - 1].q };
									const seq_fn g = { q.d0, q.d1, q.q };
									f = seq_compose(f, g);
									stage[m - 1].d0 = f.d0; stage[m - 1].d1 = f.d1; stage[m - 1].q = f.q;
								} else
									stage[m++] = q;
							}
							shi[wid] = m - wbeg;
							s_wb[wid] = wbeg;
						}
					}
					__syncthreads();
					if (tid == 0) {
						int m = 0, nraw = 0;
						if (total <= cap) {
							for (int w = 0; w < FLOF_BLOCK / 32; ++w) {
								const int wb = s_wb[w], wn = shi[w];
								for (int k = wb; k < wb + wn; ++k) {
									const seq_rec q = stage[k];
									if (m > 0 && q.e > SEQ_E_WILD && stage[m - 1].e == q.e) {
										seq_fn f = { stage[m - 1].d0, stage[m - 1].d1, stage[m - 1].q };
										const seq_fn g = { q.d0, q.d1, q.q };
										f = seq_compose(f, g);
										stage[m - 1].d0 = f.d0; stage[m - 1].d1 = f.d1; stage[m - 1].q = f.q;
									} else {
										stage[m++] = q;
										nraw += q.e == SEQ_E_RAW;
									}
								}
							}
						}
						// a leaf that stays fragmented (the running sum hovers around zero or a binade boundary) is kept as
						// its plain products: one walk step, SEQ_LEAF_CELLS * 4 real adds in the resolver
						const bool rawleaf = total > cap || m > SEQ_RAWLEAF_MIN;
						const unsigned need = rawleaf ? 1u + SEQ_RAWLEAF_RECS : (unsigned)m;
						const unsigned id = atomicAdd(&A.ctl->ndirty, 1u);
						const unsigned base = atomicAdd(&A.ctl->pool_used, need);
						s_rawbase = -1;
						if (id >= SEQ_DMAX || base + need > SEQ_POOL)
							atomicOr(&A.ctl->flags, 2u | 0x400u);  // dirty-leaf list or piece pool full
						else {
							seq_rec r;
							if (rawleaf) {
								r.d0 = r.d1 = 0.; r.e = SEQ_E_RAWLEAF; r.q = 0; r.pad[0] = (int)base + 1; r.pad[1] = SEQ_LEAF_CELLS * 4;
								A.pool[base] = r;
								s_rawbase = (int)base + 1;
								m = 1;
								nraw = SEQ_LEAF_CELLS * 4;
								atomicAdd(&A.ctl->n_rawleaves, 1ull);
							} else
								for (int k = 0; k < m; ++k) A.pool[base + k] = stage[k];
							em.flush();
							r.d0 = r.d1 = 0.; r.e = SEQ_E_DIRTY; r.q = 0; r.pad[0] = (int)base; r.pad[1] = m;
							em.put(r);
							atomicAdd(&A.ctl->n_raw, (unsigned long long)nraw);
							atomicAdd(&A.ctl->n_pieces, (unsigned long long)m);
						}
					}
					__syncthreads();
					if (s_rawbase >= 0) {  // this thread's sixteen consecutive products, in order (cells beyond the range: zeros)
						float4 *fp = reinterpret_cast<float4 *>(A.pool + s_rawbase) + tid * SEQ_U;
#pragma unroll
						for (int k = 0; k < SEQ_U; ++k) fp[k] = make_float4(xs[4 * k], xs[4 * k + 1], xs[4 * k + 2], xs[4 * k + 3]);
					}
				}
				__syncthreads();  // s_x, s_fn and the staging area are reused by the next leaf
			}
		}
		if (tid == 0) {
			em.flush();
			A.ecnt[seg] = em.n <= SEQ_ECAP ? em.n : SEQ_ECAP + 1;
		}
		__syncthreads();  // s_mode / s_fn are rewritten for the next segment
	}
}

#define SEQ_DOT_SMEM ((size_t)SEQ_STAGE * sizeof(seq_rec))  // >= FLOF_BLOCK * SEQ_XS floats

// dynamic shared memory of the resolver: all segment entries, then the flat list of walk steps
#define SEQ_CTHREADS 64  // threads that compose entry chunks (one run step each + their dirty leaves)
#define SEQ_SMAX 4096    // walk steps (run steps + pieces of the dirty leaves) staged in shared memory
#define SEQ_RAWMAX 64    // leaves kept as plain products the walk takes per dot product
#define SEQ_RESOLVE_SMEM ((size_t)(SEQ_EMAX + SEQ_SMAX) * sizeof(seq_rec) + (size_t)SEQ_DMAX * 4 * sizeof(int))

// A walk step is a seq_rec whose e / q fields are re-used: S <- S + (mantissa of S odd ? d1 : d0), and a run step also
// checks that S lies in the binade the run was built for: ((hi32(S) >> 20) ^ e) & q must be 0 with e = biased exponent,
// q = 0x7ff (q = 0 for raw products, identities and the place holders of raw leaves: no check, d0 = d1).
#define SEQ_STEP(r)                                                                         \
	{                                                                                       \
		const int hi_ = __double2hiint(S), lo_ = __double2loint(S);                         \
		wrong |= (unsigned)(((hi_ >> 20) ^ (r).e) & (int)(r).q);                            \
		S = __dadd_rn(S, (lo_ & 1) ? (r).d1 : (r).d0);                                      \
	}
// the sequential part: one thread, ~10 instructions per step, the records of eight steps in flight
__device__ __forceinline__ void seq_walk(const seq_rec *st, int k0, int k1, double &S, unsigned &wrong)
{
	int k = k0;
	for (; k + 8 <= k1; k += 8) {
		seq_rec r[8];
#pragma unroll
		for (int u = 0; u < 8; ++u) r[u] = st[k + u];
#pragma unroll
		for (int u = 0; u < 8; ++u) SEQ_STEP(r[u]);
	}
	for (; k < k1; ++k) {
		const seq_rec r = st[k];
		SEQ_STEP(r);
	}
}
__device__ __forceinline__ seq_rec seq_run_step(const seq_fn &f, int e)
{
	seq_rec o;
	o.d0 = f.d0; o.d1 = f.d1; o.pad[0] = (int)f.q; o.pad[1] = 0;  // (pad[0]: the function's parities, for seq_precompose)
	o.e = e > SEQ_E_WILD ? e + 1023 : 0;
	o.q = e > SEQ_E_WILD ? 0x7ffu : 0u;
	return o;
}

// Sharded level, ranks above 0: while the exact running sum of the lower ranks is still on its way, neighbouring run
// steps of one binade are composed into one and identity steps dropped (in place; returns the new step count), so that
// the part of the walk that sits in the rank-to-rank chain shrinks to the binade changes and raw products of this rank.
__device__ __forceinline__ int seq_precompose(seq_rec *st, int N)
{
	int m = 0, e = 0;
	bool have = false;
	seq_fn f = seq_identity();
	for (int k = 0; k < N; ++k) {
		const seq_rec r = st[k];
		if (r.q == 0x7ffu) {
			const seq_fn g = { r.d0, r.d1, (unsigned)r.pad[0] };
			if (have && e == r.e)
				f = seq_compose(f, g);
			else {
				if (have) {
					seq_rec o = seq_run_step(f, e - 1023);
					st[m++] = o;
				}
				have = true;
				e = r.e;
				f = g;
			}
		} else if (r.d0 != 0. || r.d1 != 0.) {  // a raw product (identities add nothing: dropped)
			if (have) {
				seq_rec o = seq_run_step(f, e - 1023);
				st[m++] = o;
				have = false;
			}
			st[m++] = r;
		}
	}
	if (have) {
		seq_rec o = seq_run_step(f, e - 1023);
		st[m++] = o;
	}
	return m;
}

template <int KIND>
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_seq_resolve(const float4 *__restrict__ a, const float4 *__restrict__ b, float diag, seq_args A, int mode, float accuracy,
                  int maxIter, flof_cg_state *st, int multi, flof_p2p_dev pp)
{
	if (st && st->done) return;
	extern __shared__ __align__(16) unsigned char seq_smem[];
	seq_rec *s_in = reinterpret_cast<seq_rec *>(seq_smem);       // [SEQ_EMAX] all segment entries, in order
	seq_rec *s_st = s_in + SEQ_EMAX;                             // [SEQ_SMAX] walk steps, in order
	int *s_jsrc = reinterpret_cast<int *>(s_st + SEQ_SMAX);      // [SEQ_DMAX] copy jobs of the dirty leaves: pool offset,
	int *s_jcnt = s_jsrc + SEQ_DMAX;                             //            piece count,
	int *s_jdst = s_jcnt + SEQ_DMAX;                             //            first step,
	int *s_jpfx = s_jdst + SEQ_DMAX;                             //            pieces of all earlier leaves
	__shared__ int shi[8];
	__shared__ unsigned s_bad;
	__shared__ int s_E, s_D, s_N, s_NP, s_nraw, s_nx;
	__shared__ int s_rawpos[SEQ_RAWMAX], s_rawsrc[SEQ_RAWMAX], s_rawcnt[SEQ_RAWMAX];
	const int tid = threadIdx.x;
	seq_ctl *ctl = A.ctl;
	long long tp[5];  // phase time stamps of thread 0 (seq_ctl::prof)
	tp[0] = clock64();
	const unsigned flags = ctl->flags;
	const int nseg = A.part.nseg;
	if (tid == 0) {
		s_bad = (flags & ~1u) | (ctl->ndirty > SEQ_DMAX ? 2u | 0x400u : 0u);
		s_nraw = 0;
		s_nx = 0;
	}
	__syncthreads();
	// ---- A: gather the segment entries in order.  The first entry of every segment comes from the dense array (all
	// loads of a thread in flight at once); only careful segments have more.
	{
		const int per = (nseg + FLOF_BLOCK - 1) / FLOF_BLOCK;  // <= 8 (SEQ_MAX_SEG / FLOF_BLOCK)
		const int b0 = min(tid * per, nseg), b1 = min(b0 + per, nseg);
		int cnt[8], sum = 0;
		seq_rec first[8];
#pragma unroll
		for (int q = 0; q < 8; ++q) cnt[q] = b0 + q < b1 ? __ldcg(A.ecnt + b0 + q) : 0;
#pragma unroll
		for (int q = 0; q < 8; ++q)
			if (b0 + q < b1) {
				const double2 lo = __ldcg(reinterpret_cast<const double2 *>(A.ent0 + b0 + q));
				const int4 hi = __ldcg(reinterpret_cast<const int4 *>(A.ent0 + b0 + q) + 1);
				first[q].d0 = lo.x; first[q].d1 = lo.y; first[q].e = hi.x; first[q].q = (unsigned)hi.y;
				first[q].pad[0] = hi.z; first[q].pad[1] = hi.w;
			}
#pragma unroll
		for (int q = 0; q < 8; ++q) sum += cnt[q];
		int total;
		int off = seq_block_exscan_int(sum, shi, total);
		if (tid == 0) s_E = total;
		if (total <= SEQ_EMAX) {
#pragma unroll
			for (int q = 0; q < 8; ++q) {
				if (cnt[q] > 0) s_in[off] = first[q];
				if (cnt[q] > 1) {
					// further entries (careful segments; they cluster, one thread would fetch dozens one after the other):
					// left as a copy job for the whole CTA (the job arrays of the dirty leaves are still free here)
					const int jb = atomicAdd(&s_nx, 1);
					if (jb < SEQ_DMAX) {
						s_jsrc[jb] = b0 + q;
						s_jdst[jb] = off + 1;
						s_jcnt[jb] = cnt[q] - 1;
					} else
						for (int k = 1; k < cnt[q]; ++k) s_in[off + k] = A.ent[(size_t)(b0 + q) * SEQ_ECAP + k - 1];
				}
				off += cnt[q];
			}
		} else if (tid == 0)
			atomicOr(&s_bad, 2u | 0x800u);  // more segment entries than the resolver stages
	}
	__syncthreads();
	{  // the copy jobs: eight lanes per job, 32 jobs of the CTA in flight at once
		const int J = min(s_nx, SEQ_DMAX);
		for (int jb = tid >> 3; jb < J; jb += FLOF_BLOCK / 8) {
			const seq_rec *src = A.ent + (size_t)s_jsrc[jb] * SEQ_ECAP;
			seq_rec *dst = s_in + s_jdst[jb];
			const int c = s_jcnt[jb];
			for (int k = tid & 7; k < c; k += 8) dst[k] = src[k];
		}
	}
	__syncthreads();
	tp[1] = clock64();
	const int E = s_E <= SEQ_EMAX ? s_E : 0;
	// ---- B: SEQ_CTHREADS threads own consecutive chunks of entries: dirty leaves and pieces per chunk -> step offsets
	// (an odd chunk length keeps the 32-byte records of neighbouring threads out of the same shared-memory banks)
	const int chunk = ((E + SEQ_CTHREADS - 1) / SEQ_CTHREADS) | 1;
	const int e0 = tid < SEQ_CTHREADS ? min(tid * chunk, E) : E, e1 = tid < SEQ_CTHREADS ? min(e0 + chunk, E) : E;
	int dbefore, pbefore;
	{
		int nd = 0, np = 0;
		for (int k = e0; k < e1; ++k)
			if (s_in[k].e == SEQ_E_DIRTY) {
				++nd;
				np += s_in[k].pad[1];
			}
		int D, NP;
		dbefore = seq_block_exscan_int(nd, shi, D);
		pbefore = seq_block_exscan_int(np, shi, NP);
		if (tid == 0) {
			s_D = D;
			s_NP = NP;
			s_N = SEQ_CTHREADS + D + NP;
			if (D > SEQ_DMAX || SEQ_CTHREADS + D + NP > SEQ_GSTEPS) atomicOr(&s_bad, 2u | 0x1000u);  // more dirty leaves / walk steps than the resolver takes
		}
	}
	__syncthreads();
	const bool fits = s_D <= SEQ_DMAX && s_N <= SEQ_GSTEPS;
	const int D = fits ? s_D : 0, N = fits ? s_N : 0, NP = fits ? s_NP : 0;
	// the walk steps live in shared memory; a dot product with more of them (rare) takes the global list: slower, same result
	const bool in_smem = N <= SEQ_SMAX;
	seq_rec *const steps = in_smem ? s_st : A.gsteps;
	// ---- C: every composing thread folds the runs of its chunk into run steps, cutting at dirty leaves.
	// First step of thread t = t + (dirty leaves before its chunk) + (their pieces): dense and ordered.
	if (tid < SEQ_CTHREADS && fits) {
		int pos = tid + dbefore + pbefore, dk = dbefore, pk = pbefore;
		seq_fn f = seq_identity();
		int e = SEQ_E_WILD;
		unsigned bad = 0;
		for (int k = e0; k < e1; ++k) {
			const seq_rec r = s_in[k];
			if (r.e == SEQ_E_WILD) continue;
			if (r.e == SEQ_E_DIRTY) {
				steps[pos++] = seq_run_step(f, e);
				s_jsrc[dk] = r.pad[0];
				s_jcnt[dk] = r.pad[1];
				s_jdst[dk] = pos;
				s_jpfx[dk] = pk;
				pos += r.pad[1];
				pk += r.pad[1];
				++dk;
				f = seq_identity();
				e = SEQ_E_WILD;
				continue;
			}
			if (e != SEQ_E_WILD && e != r.e) bad = 4u;  // two clean neighbours in different binades: cannot happen
			e = r.e;
			const seq_fn g = { r.d0, r.d1, r.q };
			f = seq_compose(f, g);
		}
		steps[pos] = seq_run_step(f, e);
		if (bad) atomicOr(&s_bad, bad);
	}
	__syncthreads();
	// the pieces of all dirty leaves, flat over the CTA: piece -> its leaf by bisection of the piece prefixes
	for (int p = tid; p < NP; p += FLOF_BLOCK) {
		int lo = 0, hi = D - 1;
		while (lo < hi) {
			const int mid = (lo + hi + 1) >> 1;
			if (s_jpfx[mid] <= p)
				lo = mid;
			else
				hi = mid - 1;
		}
		const int j = p - s_jpfx[lo];
		seq_rec r = A.pool[s_jsrc[lo] + j];
		if (r.e == SEQ_E_RAWLEAF) {
			const int i = atomicAdd(&s_nraw, 1);
			if (i < SEQ_RAWMAX) {
				s_rawpos[i] = s_jdst[lo] + j;
				s_rawsrc[i] = r.pad[0];
				s_rawcnt[i] = r.pad[1];
			} else
				atomicOr(&s_bad, 2u | 0x2000u);  // more raw leaves than the walk's side list takes
			r.d0 = r.d1 = 0.;
		}
		r.pad[0] = (int)r.q;
		r.q = r.e > SEQ_E_WILD ? 0x7ffu : 0u;  // (raw products carry d0 = d1 = the product)
		r.e = r.e > SEQ_E_WILD ? r.e + 1023 : 0;
		steps[s_jdst[lo] + j] = r;
	}
	__syncthreads();
	if (tid != 0) return;
	tp[2] = clock64();
	// ---- D: the sequential walk (one thread)
	double S = 0.;
	unsigned int cseq = 0;
	int Nw = N;  // steps the walk takes
	if (multi && pp.rank > 0 && in_smem && s_nraw == 0 && !s_bad && !(flags & 1u)) Nw = seq_precompose(s_st, N);
	if (multi) {  // the running sum continues from the rank below (exact bits handed over through the mailboxes)
		cseq = ++(*pp.chain_seq);
		if (pp.rank > 0) {
			flof_mbox_hdr *me = (flof_mbox_hdr *)pp.peer[pp.rank];
			if (p2p_wait(&me->chain[cseq & 1u].seq, cseq, pp.err)) S = *(volatile double *)&me->chain[cseq & 1u].v;
		}
	}
	unsigned bad = s_bad;
	if (flags & 1u) {
		// a non-finite product: every summation order ends in the same Inf/NaN class; take the approximate sum
		S = S + ctl->tot[0];
	} else if (!bad) {
		unsigned wrong = 0;
		const int nraw = s_nraw;
		for (int i = 1; i < nraw; ++i) {  // raw leaves in walk order (a handful at most)
			const int p = s_rawpos[i], sr = s_rawsrc[i], c = s_rawcnt[i];
			int j = i - 1;
			for (; j >= 0 && s_rawpos[j] > p; --j) {
				s_rawpos[j + 1] = s_rawpos[j];
				s_rawsrc[j + 1] = s_rawsrc[j];
				s_rawcnt[j + 1] = s_rawcnt[j];
			}
			s_rawpos[j + 1] = p; s_rawsrc[j + 1] = sr; s_rawcnt[j + 1] = c;
		}
		int k = 0;
		for (int i = 0; i <= nraw; ++i) {
			const int kend = i < nraw ? s_rawpos[i] : Nw;
			if (in_smem)
				seq_walk(s_st, k, kend, S, wrong);
			else
				seq_walk(A.gsteps, k, kend, S, wrong);
			if (i == nraw) break;
			// a leaf kept as plain products: sixteen float4 per batch, the next batch in flight while this one is added
			const float4 *fp = reinterpret_cast<const float4 *>(A.pool + s_rawsrc[i]);
			const int n4 = s_rawcnt[i] >> 2;
			float4 cur[16], nxv[16];
#pragma unroll
			for (int u = 0; u < 16; ++u) cur[u] = __ldcg(fp + u);
			for (int j = 0; j < n4; j += 16) {
				if (j + 16 < n4) {
#pragma unroll
					for (int u = 0; u < 16; ++u) nxv[u] = __ldcg(fp + j + 16 + u);
				}
#pragma unroll
				for (int u = 0; u < 16; ++u) {
					S = __dadd_rn(S, (double)cur[u].x);
					S = __dadd_rn(S, (double)cur[u].y);
					S = __dadd_rn(S, (double)cur[u].z);
					S = __dadd_rn(S, (double)cur[u].w);
				}
#pragma unroll
				for (int u = 0; u < 16; ++u) cur[u] = nxv[u];
			}
			k = kend + 1;  // (the place holder step of the raw leaf adds nothing)
		}
		if (wrong) bad = 4u;
	}
	if (bad) {
		// capacity exceeded or inconsistent (never seen on CG data; tests assert the counters stay 0): the plain loop on
		// one thread -- slow, exact by definition -- while that is affordable, else the approximate (tree) sum of pass 1
		S = 0.;
		if (multi && pp.rank > 0) S = *(volatile double *)&((flof_mbox_hdr *)pp.peer[pp.rank])->chain[cseq & 1u].v;
		if (A.part.ncells <= SEQ_PLAIN_MAX) {
			for (int c = 0; c < A.part.ncells; ++c) {
				const float4 p = seq_products<KIND>(a, b, c, diag);
				S = __dadd_rn(S, (double)p.x);
				S = __dadd_rn(S, (double)p.y);
				S = __dadd_rn(S, (double)p.z);
				S = __dadd_rn(S, (double)p.w);
			}
		} else {
			S = S + ctl->tot[0];
			ctl->n_inexact++;
			if (st) st->seq_inexact = 1;  // the solve reports it (flof_solve.cu cg_run)
		}
		ctl->n_fallback++;
		ctl->why |= bad;
		if (bad & 4u) ctl->n_inconsistent++;
	}
	tp[3] = clock64();
	if (multi) {  // hand the running sum to the next rank; the last rank owns the total and tells everybody
		if (pp.rank < pp.nranks - 1) {
			flof_mbox_hdr *nx = (flof_mbox_hdr *)pp.peer[pp.rank + 1];
			*(volatile double *)&nx->chain[cseq & 1u].v = S;
			__threadfence_system();
			*(volatile unsigned int *)&nx->chain[cseq & 1u].seq = cseq;
		} else {
			for (int r = 0; r < pp.nranks; ++r) {
				flof_mbox_hdr *h = (flof_mbox_hdr *)pp.peer[r];
				*(volatile double *)&h->total[cseq & 1u].v = S;
			}
			__threadfence_system();
			for (int r = 0; r < pp.nranks; ++r)
				*(volatile unsigned int *)&((flof_mbox_hdr *)pp.peer[r])->total[cseq & 1u].seq = cseq;
		}
		flof_mbox_hdr *me = (flof_mbox_hdr *)pp.peer[pp.rank];
		if (p2p_wait(&me->total[cseq & 1u].seq, cseq, pp.err)) S = *(volatile double *)&me->total[cseq & 1u].v;
	}
	ctl->result = S;
	ctl->n_dots++;
	ctl->n_dirty += ctl->ndirty;
	ctl->ticket = 0;
	ctl->ndirty = 0;
	ctl->pool_used = 0;
	ctl->flags = 0;
	if (st) {
		if (mode == SEQ_MODE_ALPHA)
			st->alpha1 = S;
		else if (mode == SEQ_MODE_ADVANCE)
			cg_advance(st, S, st->residual, maxIter);
		else if (mode == SEQ_MODE_INIT)
			cg_init_finalize(st, S, st->residual, accuracy);
	}
	tp[4] = clock64();
	ctl->prof[0] += (unsigned long long)(tp[1] - tp[0]);  // gather of the segment entries
	ctl->prof[1] += (unsigned long long)(tp[2] - tp[1]);  // compose + piece copies
	ctl->prof[2] += (unsigned long long)(tp[3] - tp[2]);  // walk (on a sharded level including the wait for the rank below)
	ctl->prof[3] += (unsigned long long)(tp[4] - tp[3]);  // rank chain + state update
	ctl->prof[4] += (unsigned long long)N;                // walk steps
}
