// Host model of the parallel evaluation of the reference's sequential dot product (test helper, CPU only).
// It includes the SAME arithmetic header the CUDA kernels use (ofblend_b200/csrc/flof_seqsum_core.h) and
// organises the work exactly like flof_seqsum.cu -- leaves, "threads" with consecutive products, dirty leaves
// cut into function runs and raw products, a final sequential walk -- so that tests/test_seqsum_model.py can
// check the scheme against a plain `for (i) d += a[i]*b[i]` loop without a GPU.
//   perturb: the approximate prefix handed to the classification is pushed to the edge of its allowed error
//            (fraction of the margin, sign alternating per leaf) to prove that the result does not depend on it.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../ofblend_b200/csrc/flof_seqsum_core.h"

extern "C" double seqsum_reference(const float *a, const float *b, long long n)
{
	volatile double d = 0.;
	for (long long i = 0; i < n; ++i) {
		volatile float p = a[i] * b[i];
		d += (double)p;
	}
	return d;
}

struct builder {  // run-length builder of pieces: consecutive safe products of one binade merge into one RUN
	std::vector<seq_rec> *out;
	bool have;
	int e;
	seq_fn f;
	void flush()
	{
		if (!have) return;
		seq_rec r;
		r.d0 = f.d0; r.d1 = f.d1; r.e = e; r.q = f.q; r.pad[0] = r.pad[1] = 0;
		out->push_back(r);
		have = false;
	}
	void push(double x, double P, double T, double kf)
	{
		if (x == 0.) return;  // identity in every binade
		int pe;
		if (seq_range_safe(P, T, seq_abs(x), kf, &pe)) {
			double C0, C1;
			seq_consts(pe, &C0, &C1);
			const seq_fn g = seq_elem(x, C0, C1);
			if (have && e == pe) {
				f = seq_compose(f, g);
			} else {
				flush();
				have = true;
				e = pe;
				f = g;
			}
		} else {
			flush();
			seq_rec r;
			r.d0 = x; r.d1 = 0.; r.e = SEQ_E_RAW; r.q = 0; r.pad[0] = r.pad[1] = 0;
			out->push_back(r);
		}
	}
};

// stats[0] = dirty leaves, stats[1] = raw products, stats[2] = pieces, stats[3] = consistency errors
extern "C" double seqsum_model(const float *a, const float *b, long long n, int leaf_elems, int thread_elems,
                               double perturb, double s_in, long long *stats)
{
	std::vector<double> x((size_t)n);
	for (long long i = 0; i < n; ++i) {
		volatile float p = a[i] * b[i];
		x[(size_t)i] = (double)p;
	}
	const double kf = seq_margin_factor(n);
	const long long nleaf = (n + leaf_elems - 1) / leaf_elems;
	std::vector<seq_rec> leaf((size_t)nleaf);
	std::vector<std::vector<seq_rec> > dirty;
	// approximate prefixes: long double running sums (any order would do), then perturbed
	long double Pq = (long double)s_in, Tq = fabsl((long double)s_in);
	long long ndirty = 0, nraw = 0, npieces = 0, nerr = 0;
	for (long long L = 0; L < nleaf; ++L) {
		const long long i0 = L * leaf_elems, i1 = (i0 + leaf_elems < n) ? i0 + leaf_elems : n;
		double sa = 0.;
		for (long long i = i0; i < i1; ++i) sa += fabs(x[(size_t)i]);
		const double T = (double)Tq;
		const double sgn = (L & 1) ? 1. : -1.;
		const double P = (double)Pq + sgn * perturb * 0.45 * kf * T;  // stays inside half the margin
		seq_rec r;
		r.pad[0] = r.pad[1] = 0;
		int e;
		if (sa == 0.) {
			r.d0 = r.d1 = 0.; r.q = 2u; r.e = SEQ_E_WILD;
		} else if (seq_range_safe(P, T, sa, kf, &e)) {
			double C0, C1;
			seq_consts(e, &C0, &C1);
			seq_fn f = seq_identity();
			for (long long i = i0; i < i1; ++i) f = seq_compose(f, seq_elem(x[(size_t)i], C0, C1));
			r.d0 = f.d0; r.d1 = f.d1; r.q = f.q; r.e = e;
		} else {
			// dirty leaf: per "thread" (thread_elems consecutive products) with its own approximate prefix
			std::vector<seq_rec> pieces;
			builder bd;
			bd.out = &pieces;
			bd.have = false;
			long double p2 = Pq, t2 = Tq;
			for (long long i = i0; i < i1; ++i) {
				const double Pi = (double)p2 + sgn * perturb * 0.45 * kf * (double)t2;
				bd.push(x[(size_t)i], Pi, (double)t2, kf);
				p2 += x[(size_t)i];
				t2 += fabs(x[(size_t)i]);
				if (thread_elems > 0 && ((i - i0 + 1) % thread_elems) == 0) bd.flush();  // thread boundary: pieces do not span threads...
			}
			bd.flush();
			// ... and are merged afterwards like the block-level merge of the kernel
			std::vector<seq_rec> merged;
			for (size_t k = 0; k < pieces.size(); ++k) {
				if (!merged.empty() && merged.back().e == pieces[k].e && pieces[k].e > SEQ_E_WILD) {
					seq_fn f = { merged.back().d0, merged.back().d1, merged.back().q };
					seq_fn g = { pieces[k].d0, pieces[k].d1, pieces[k].q };
					f = seq_compose(f, g);
					merged.back().d0 = f.d0; merged.back().d1 = f.d1; merged.back().q = f.q;
				} else
					merged.push_back(pieces[k]);
			}
			r.d0 = r.d1 = 0.; r.e = SEQ_E_DIRTY; r.q = (unsigned)dirty.size();
			for (size_t k = 0; k < merged.size(); ++k) nraw += merged[k].e == SEQ_E_RAW;
			npieces += (long long)merged.size();
			dirty.push_back(merged);
			++ndirty;
		}
		leaf[(size_t)L] = r;
		for (long long i = i0; i < i1; ++i) {
			Pq += x[(size_t)i];
			Tq += fabs(x[(size_t)i]);
		}
	}
	// sequential walk
	double S = s_in;
	for (long long L = 0; L < nleaf; ++L) {
		const seq_rec &r = leaf[(size_t)L];
		if (r.e == SEQ_E_WILD) continue;
		if (r.e == SEQ_E_DIRTY) {
			const std::vector<seq_rec> &pc = dirty[r.q];
			for (size_t k = 0; k < pc.size(); ++k) {
				if (pc[k].e == SEQ_E_RAW) {
					volatile double t = S + pc[k].d0;
					S = t;
				} else {
					if (seq_binade(S) != pc[k].e) ++nerr;
					const seq_fn f = { pc[k].d0, pc[k].d1, pc[k].q };
					S = seq_apply(S, f);
				}
			}
			continue;
		}
		if (seq_binade(S) != r.e) ++nerr;
		const seq_fn f = { r.d0, r.d1, r.q };
		S = seq_apply(S, f);
	}
	if (stats) {
		stats[0] = ndirty;
		stats[1] = nraw;
		stats[2] = npieces;
		stats[3] = nerr;
	}
	return S;
}
