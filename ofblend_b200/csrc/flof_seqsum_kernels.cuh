// flof_seqsum_kernels.cuh -- the CG dot products in the reference's sequential summation order, evaluated in parallel.
// ref: dotProd optflow4d.cpp:234-241 (call sites :296, :307, :319).  Arithmetic: flof_seqsum_core.h.
// (included by flof_solve.cu after precond_of / cg_advance / cg_init_finalize)
//
// k_dot_seq<KIND>   one pass over the two vectors (32 B/cell).  A CTA takes leaves of 1024 cells in ticket order;
//                   per leaf: products -> shared memory, approximate leaf sums, decoupled look-back over the leaf
//                   descriptors for the approximate prefix (single-pass scan), then
//                     clean leaf (running sum provably inside one binade): every thread folds its 16 consecutive
//                       products into a rounding function, ordered warp/block reduction -> one 32-byte leaf record;
//                     dirty leaf (around a binade crossing, or while the sum builds up from zero): per product
//                       classification with the thread's own approximate prefix -> runs + raw products, merged and
//                       appended to the piece pool.
// k_seq_resolve     one CTA: composes the clean leaf records between dirty leaves in parallel, then one thread
//                   walks runs and raw products with real fp64 adds.  The result has the bits of the reference's
//                   loop; the tail feeds it to the CG state (alpha1 / cg_advance / cg_init_finalize).
//                   If a capacity is exceeded or a consistency check fails it falls back to the plain loop on one
//                   thread (counted in seq_ctl::n_fallback; tests assert it stays 0).
//   KIND 0: sum_i a[i]*b[i]                         (srch . A srch)
//   KIND 1: sum_i (a[i]*precond(b)[i]) * a[i]       (a = res, b = grad: tmp = res*precond; dot(tmp, res))
#pragma once
#include "flof_seqsum.cuh"

struct seq_args {
	seq_desc *desc;
	seq_rec *leaf;
	seq_rec *pool;
	seq_ctl *ctl;
	const double *off;  // multi-GPU: approximate {sum, sum of magnitudes} of the lower ranks' slabs (device), else NULL
	unsigned int epoch;
	int nleaf;
	int ncells;
	double kf;          // seq_margin_factor(total number of products over all ranks)
};

#define SEQ_MODE_NONE 0     // result only (seq_ctl::result)
#define SEQ_MODE_ALPHA 1    // st->alpha1 = dot(srch, A srch)                  ref :307
#define SEQ_MODE_ADVANCE 2  // cg_advance(st, dot(tmp, res), st->residual)     ref :311-324
#define SEQ_MODE_INIT 3     // cg_init_finalize(st, dot(tmp, res), st->residual) ref :286-301

#define SEQ_LEAF_WILD 0
#define SEQ_LEAF_CLEAN 1
#define SEQ_LEAF_DIRTY 2
#define SEQ_XS 20  // floats per thread chunk in shared memory: 16 products + 4 pad (conflict-free LDS.128 at 80 B lane stride)

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

// approximate exclusive prefix of leaf L: decoupled look-back over the descriptors (executed by one full warp).
// Earlier leaves are owned by CTAs that drew their ticket before this one, i.e. are running and will publish.
__device__ __forceinline__ void seq_lookback(const seq_args &A, int L, double &Px, double &Pa)
{
	const int lane = threadIdx.x & 31;
	double accx = 0., acca = 0.;
	int base = L - 1;
	for (;;) {
		const int idx = base - lane;
		int state = 2;
		double vx = 0., va = 0.;
		if (idx >= 0) {
			volatile seq_desc *d = A.desc + idx;
			for (;;) {
				if (d->st_pre == A.epoch) { state = 2; break; }
				if (d->st_agg == A.epoch) { state = 1; break; }
			}
			__threadfence();
			vx = state == 2 ? d->px : d->ax;
			va = state == 2 ? d->pa : d->aa;
		} else if (idx == -1 && A.off) {  // the virtual leaf before the first one carries the lower ranks' slabs
			vx = A.off[0];
			va = A.off[1];
		}
		const unsigned pm = __ballot_sync(0xffffffffu, state == 2);
		const int first = pm ? __ffs(pm) - 1 : 32;
		if (lane > first) vx = va = 0.;
		for (int o = 16; o > 0; o >>= 1) {
			vx += __shfl_xor_sync(0xffffffffu, vx, o);
			va += __shfl_xor_sync(0xffffffffu, va, o);
		}
		accx += vx;
		acca += va;
		if (pm) break;
		base -= 32;
	}
	Px = accx;
	Pa = acca;
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
			r.d0 = x; r.d1 = 0.; r.e = SEQ_E_RAW; r.q = 0; r.pad[0] = r.pad[1] = 0;
			out[n++] = r;
		}
	}
};

template <int KIND>
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_dot_seq(const float4 *__restrict__ a, const float4 *__restrict__ b, float diag, seq_args A, const flof_cg_state *st)
{
	if (st && st->done) return;
	__shared__ __align__(16) float s_x[FLOF_BLOCK * SEQ_XS];  // 20 KB; reused as piece staging by dirty leaves
	__shared__ double shd[32];
	__shared__ seq_fn s_fn[FLOF_BLOCK / 32];
	__shared__ int shi[8];
	__shared__ int s_leaf, s_mode, s_e;
	__shared__ double s_P, s_T;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	for (;;) {
		if (tid == 0) s_leaf = (int)atomicAdd(&A.ctl->ticket, 1u);
		__syncthreads();
		const int L = s_leaf;
		if (L >= A.nleaf) break;
		// ---- products of the leaf (cells L*1024 + s*256 + tid), parked in shared memory in element order
		double sx = 0., sa = 0.;
#pragma unroll
		for (int s = 0; s < SEQ_U; ++s) {
			const int ci = s * FLOF_BLOCK + tid, c = L * SEQ_LEAF_CELLS + ci;
			float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
			if (c < A.ncells) p = seq_products<KIND>(a, b, c, diag);
			*reinterpret_cast<float4 *>(&s_x[(ci >> 2) * SEQ_XS + (ci & 3) * 4]) = p;
			sx += (double)p.x + (double)p.y + (double)p.z + (double)p.w;  // approximate: any order will do
			sa += (double)fabsf(p.x) + (double)fabsf(p.y) + (double)fabsf(p.z) + (double)fabsf(p.w);
		}
		sx = flof_block_sum(sx, shd);
		sa = flof_block_sum(sa, shd);
		if (wid == 0) {
			volatile seq_desc *d = A.desc + L;
			if (lane == 0) {
				d->ax = sx;
				d->aa = sa;
				__threadfence();
				d->st_agg = A.epoch;
			}
			double Px, Pa;
			seq_lookback(A, L, Px, Pa);
			if (lane == 0) {
				d->px = Px + sx;
				d->pa = Pa + sa;
				__threadfence();
				d->st_pre = A.epoch;
				int mode, e = 0;
				if (!seq_finite(sx) || !seq_finite(sa)) {
					atomicOr(&A.ctl->flags, 1u);
					mode = SEQ_LEAF_WILD;
				} else if (sa == 0.)
					mode = SEQ_LEAF_WILD;
				else
					mode = seq_range_safe(Px, Pa, sa, A.kf, &e) ? SEQ_LEAF_CLEAN : SEQ_LEAF_DIRTY;
				s_mode = mode;
				s_e = e;
				s_P = Px;
				s_T = Pa;
			}
		}
		__syncthreads();
		const int mode = s_mode;
		const float4 *xp = reinterpret_cast<const float4 *>(&s_x[tid * SEQ_XS]);
		if (mode == SEQ_LEAF_CLEAN) {
			double C0, C1;
			seq_consts(s_e, &C0, &C1);
			seq_fn f = seq_identity();
#pragma unroll
			for (int k = 0; k < SEQ_U; ++k) {
				const float4 v = xp[k];
				f = seq_compose(f, seq_elem((double)v.x, C0, C1));
				f = seq_compose(f, seq_elem((double)v.y, C0, C1));
				f = seq_compose(f, seq_elem((double)v.z, C0, C1));
				f = seq_compose(f, seq_elem((double)v.w, C0, C1));
			}
			f = seq_warp_compose(f, 32);
			if (lane == 0) s_fn[wid] = f;
			__syncthreads();
			if (wid == 0) {
				f = lane < FLOF_BLOCK / 32 ? s_fn[lane] : seq_identity();
				f = seq_warp_compose(f, FLOF_BLOCK / 32);
				if (lane == 0) {
					seq_rec r;
					r.d0 = f.d0; r.d1 = f.d1; r.e = s_e; r.q = f.q; r.pad[0] = r.pad[1] = 0;
					A.leaf[L] = r;
				}
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
			double P = s_P + px, T = s_T + pa;
#pragma unroll 1
			for (int k = 0; k < 4 * SEQ_U; ++k) {
				const double x = (double)xs[k];
				bd.push(x, P, T, A.kf);
				P += x;
				T += seq_abs(x);
			}
			bd.flush();
			int total;
			const int off = seq_block_exscan_int(bd.n, shi, total);
			seq_rec *stage = reinterpret_cast<seq_rec *>(s_x);
			const int cap = (int)(sizeof(s_x) / sizeof(seq_rec));
			if (total <= cap)
				for (int k = 0; k < bd.n; ++k) stage[off + k] = pc[k];
			__syncthreads();
			if (tid == 0) {
				seq_rec r;
				r.d0 = r.d1 = 0.; r.e = SEQ_E_WILD; r.q = 2u; r.pad[0] = r.pad[1] = 0;
				if (total > cap)
					atomicOr(&A.ctl->flags, 2u);
				else {
					// merge neighbouring runs of one binade (pieces of different threads), count raw products
					int m = 0, nraw = 0;
					for (int k = 0; k < total; ++k) {
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
					const unsigned id = atomicAdd(&A.ctl->ndirty, 1u);
					const unsigned base = atomicAdd(&A.ctl->pool_used, (unsigned)m);
					if (id >= SEQ_DMAX || base + (unsigned)m > SEQ_POOL)
						atomicOr(&A.ctl->flags, 2u);
					else {
						for (int k = 0; k < m; ++k) A.pool[base + k] = stage[k];
						A.ctl->dirty_leaf[id] = L;
						A.ctl->dirty_base[id] = base;
						A.ctl->dirty_cnt[id] = (unsigned)m;
						r.e = SEQ_E_DIRTY;
						r.q = id;
						atomicAdd(&A.ctl->n_raw, (unsigned long long)nraw);
						atomicAdd(&A.ctl->n_pieces, (unsigned long long)m);
					}
				}
				A.leaf[L] = r;
			}
		} else if (tid == 0) {
			seq_rec r;
			r.d0 = r.d1 = 0.; r.e = SEQ_E_WILD; r.q = 2u; r.pad[0] = r.pad[1] = 0;
			A.leaf[L] = r;
		}
		__syncthreads();  // s_leaf, s_x and the staging area are reused by the next leaf
	}
}

// dynamic shared memory of the resolver
#define SEQ_RESOLVE_SMEM ((size_t)(FLOF_BLOCK + SEQ_DMAX + SEQ_PIECE_SMEM) * sizeof(seq_rec) + (size_t)SEQ_DMAX * 4 * sizeof(int))

template <int KIND>
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_seq_resolve(const float4 *__restrict__ a, const float4 *__restrict__ b, float diag, seq_args A, int mode, float accuracy,
                  int maxIter, flof_cg_state *st, int multi, flof_p2p_dev pp)
{
	if (st && st->done) return;
	extern __shared__ __align__(16) unsigned char seq_smem[];
	seq_rec *s_ent = reinterpret_cast<seq_rec *>(seq_smem);  // [FLOF_BLOCK + SEQ_DMAX] runs between dirty leaves, in order
	seq_rec *s_pc = s_ent + FLOF_BLOCK + SEQ_DMAX;           // [SEQ_PIECE_SMEM] staged pieces of the dirty leaves
	int *s_dl = reinterpret_cast<int *>(s_pc + SEQ_PIECE_SMEM);  // [SEQ_DMAX] dirty leaves, sorted
	int *s_id = s_dl + SEQ_DMAX;                             // their ids in the dirty list
	int *s_po = s_id + SEQ_DMAX;                             // offset of their pieces in s_pc, -1 = read from the pool
	int *s_key = s_po + SEQ_DMAX;                            // unsorted keys (scratch)
	__shared__ int shi[8];
	__shared__ unsigned s_bad;
	const int tid = threadIdx.x;
	seq_ctl *ctl = A.ctl;
	const unsigned flags = ctl->flags;
	const unsigned nd = ctl->ndirty;
	const int D = (int)(nd < SEQ_DMAX ? nd : SEQ_DMAX);
	if (tid == 0) s_bad = (flags & 6u) | (nd > SEQ_DMAX ? 2u : 0u);
	// ---- A: dirty leaves in leaf order (rank sort; their number is small)
	for (int i = tid; i < D; i += FLOF_BLOCK) s_key[i] = ctl->dirty_leaf[i];
	__syncthreads();
	for (int i = tid; i < D; i += FLOF_BLOCK) {
		const int key = s_key[i];
		int rank = 0;
		for (int j = 0; j < D; ++j) rank += s_key[j] < key;
		s_dl[rank] = key;
		s_id[rank] = i;
	}
	__syncthreads();
	{  // piece offsets in sorted order (4 dirty leaves per thread), then stage the pieces
		int cnt[4], sum = 0;
		for (int q = 0; q < 4; ++q) {
			const int k = tid * 4 + q;
			cnt[q] = k < D ? (int)ctl->dirty_cnt[s_id[k]] : 0;
			sum += cnt[q];
		}
		int total;
		int off = seq_block_exscan_int(sum, shi, total);
		for (int q = 0; q < 4; ++q) {
			const int k = tid * 4 + q;
			if (k < D) s_po[k] = off + cnt[q] <= SEQ_PIECE_SMEM ? off : -1;
			off += cnt[q];
		}
	}
	__syncthreads();
	for (int k = tid; k < D; k += FLOF_BLOCK) {
		const int id = s_id[k], po = s_po[k];
		if (po < 0) continue;
		const seq_rec *src = A.pool + ctl->dirty_base[id];
		const int cnt = (int)ctl->dirty_cnt[id];
		for (int j = 0; j < cnt; ++j) s_pc[po + j] = src[j];
	}
	// ---- B: every thread composes the clean records of its chunk of leaves, cutting at dirty leaves.
	// Entry slot of thread t = t + (dirty leaves before its chunk) + (dirty leaves met so far): dense and ordered.
	{
		const int chunk = (A.nleaf + FLOF_BLOCK - 1) / FLOF_BLOCK;
		const int l0 = min(tid * chunk, A.nleaf), l1 = min(l0 + chunk, A.nleaf);
		int lo = 0, hi = D;
		while (lo < hi) {  // first sorted dirty leaf >= l0
			const int mid = (lo + hi) >> 1;
			if (s_dl[mid] < l0) lo = mid + 1; else hi = mid;
		}
		int slot = tid + lo, dk = lo;
		seq_fn f = seq_identity();
		int e = SEQ_E_WILD;
		unsigned bad = 0;
#pragma unroll 4
		for (int L = l0; L < l1; ++L) {
			const seq_rec r = A.leaf[L];
			if (r.e == SEQ_E_WILD) continue;
			if (r.e == SEQ_E_DIRTY) {
				if (dk >= D || s_dl[dk] != L) { bad = 4u; break; }
				seq_rec o;
				o.d0 = f.d0; o.d1 = f.d1; o.e = e; o.q = f.q; o.pad[0] = dk; o.pad[1] = 0;
				s_ent[slot++] = o;
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
		seq_rec o;
		o.d0 = f.d0; o.d1 = f.d1; o.e = e; o.q = f.q; o.pad[0] = -1; o.pad[1] = 0;
		s_ent[slot] = o;
		if (bad) atomicOr(&s_bad, bad);
	}
	__syncthreads();
	if (tid != 0) return;
	// ---- C: the sequential walk
	double S = 0.;
	unsigned int cseq = 0;
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
		volatile seq_desc *d = A.desc + (A.nleaf - 1);
		S = S + d->px - (A.off ? A.off[0] : 0.);
	} else if (!bad) {
		const int nent = FLOF_BLOCK + D;
		for (int k = 0; k < nent && !bad; ++k) {
			const seq_rec r = s_ent[k];
			if (r.e != SEQ_E_WILD) {
				if (seq_binade(S) != r.e) { bad = 4u; break; }
				const seq_fn f = { r.d0, r.d1, r.q };
				S = seq_apply(S, f);
			}
			const int dk = r.pad[0];
			if (dk >= 0) {
				const int id = s_id[dk], cnt = (int)ctl->dirty_cnt[id];
				const seq_rec *pc = s_po[dk] >= 0 ? s_pc + s_po[dk] : A.pool + ctl->dirty_base[id];
				for (int j = 0; j < cnt; ++j) {
					const seq_rec q = pc[j];
					if (q.e == SEQ_E_RAW)
						S = __dadd_rn(S, q.d0);
					else {
						if (seq_binade(S) != q.e) { bad = 4u; break; }
						const seq_fn f = { q.d0, q.d1, q.q };
						S = seq_apply(S, f);
					}
				}
			}
		}
	}
	if (bad) {  // capacity exceeded or inconsistent: the plain loop (slow, exact by definition)
		S = 0.;
		if (multi && pp.rank > 0) S = *(volatile double *)&((flof_mbox_hdr *)pp.peer[pp.rank])->chain[cseq & 1u].v;
		for (int c = 0; c < A.ncells; ++c) {
			const float4 p = seq_products<KIND>(a, b, c, diag);
			S = __dadd_rn(S, (double)p.x);
			S = __dadd_rn(S, (double)p.y);
			S = __dadd_rn(S, (double)p.z);
			S = __dadd_rn(S, (double)p.w);
		}
		ctl->n_fallback++;
		if (bad & 4u) ctl->n_inconsistent++;
	}
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
	ctl->n_dirty += nd;
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
}
