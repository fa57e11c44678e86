// flof_seqsum_core.h -- the arithmetic behind the CG's "sequential-order" dot products.
// ref: dotProd optflow4d.cpp:234-241
//        double d = 0.;  for (i = 0; i < N; ++i) d += a[i] * b[i];     (fp32 product, fp64 running sum)
//
// The reference sums its 4*cells products one after the other.  Every add rounds, so the result depends on
// the order, and the mode-1 output depends on those last bits (the final SDF projection amplifies round-off
// level differences by ~4e4).  A tree reduction can therefore never reproduce the reference
// bit for bit.  This file restates the sequential sum in a form that CAN be evaluated in parallel:
//
//   While the running sum S stays inside one binade [2^e, 2^(e+1)), all representable values are multiples
//   of u = 2^(e-52) and  S <- fl(S + x)  is the integer update  M <- RN_even(M + x/u)  of the mantissa
//   M = S/u.  The rounded increment depends on M only through its parity, and only when x/u is an exact
//   tie.  So an element is a function  parity -> (increment, new parity), these functions compose
//   associatively, and the increments are multiples of u below 2^e, i.e. exact in fp64.
//   The increment for an even (odd) M is obtained with two fp64 adds each:  fl(x + C) - C  with
//   C = 1.5*2^e (mantissa even) resp. C + u (mantissa odd) rounds x to a multiple of u with exactly the
//   tie rule of the true sum; the low mantissa bit of fl(x + C) is the parity after the add.
//
// A leaf (a few thousand consecutive products) whose running sum provably stays inside one binade -- decided
// from an APPROXIMATE prefix sum with a rigorous error margin -- is reduced to one such function by an ordered
// parallel reduction; the few leaves around a binade crossing are cut into function runs and single "raw"
// products.  A short sequential walk over runs and raw products (real fp64 adds) then yields exactly the bits
// of the reference's loop.  Host and device code share this header (tests/test_seqsum_model.py runs the host
// side against a plain sequential loop).
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define SEQ_HD __host__ __device__ __forceinline__
#else
#define SEQ_HD static inline
#endif

#define SEQ_E_WILD (-100000)   // leaf/piece of zeros only: identity, compatible with any binade
#define SEQ_E_DIRTY (-100001)  // leaf record: see the dirty list
#define SEQ_E_RAW (-100002)    // piece: one product, added with a real fp64 add
#define SEQ_E_RAWLEAF (-100003)  // piece: a whole leaf of products kept as plain floats in the pool (pad[0] = first pool
                                 // record of the floats, pad[1] = number of products), added one by one with real fp64 adds
#define SEQ_E_MIN (-900)       // binades outside [MIN, MAX] (subnormal neighbourhood / overflow) are never "safe"
#define SEQ_E_MAX (900)

struct seq_fn {      // parity of the mantissa -> (exact increment, parity afterwards)
	double d0, d1;   // increment if the mantissa is even / odd
	unsigned q;      // bit 0: parity after, starting even; bit 1: parity after, starting odd
};
#ifdef __CUDACC__
#define SEQ_ALIGN16 __align__(16)
#else
#define SEQ_ALIGN16 alignas(16)
#endif
struct SEQ_ALIGN16 seq_rec {  // 32-byte leaf record / piece
	double d0, d1;   // function increments; RAW piece: d0 = the product
	int e;           // binade, or SEQ_E_WILD / SEQ_E_DIRTY / SEQ_E_RAW
	unsigned q;      // function parities; DIRTY leaf record: index into the dirty list
	int pad[2];
};

SEQ_HD uint64_t seq_bits(double v)
{
#ifdef __CUDA_ARCH__
	return (uint64_t)__double_as_longlong(v);
#else
	uint64_t b;
	memcpy(&b, &v, 8);
	return b;
#endif
}
SEQ_HD double seq_from_bits(uint64_t b)
{
#ifdef __CUDA_ARCH__
	return __longlong_as_double((long long)b);
#else
	double v;
	memcpy(&v, &b, 8);
	return v;
#endif
}
SEQ_HD double seq_abs(double v) { return seq_from_bits(seq_bits(v) & 0x7fffffffffffffffull); }
// unbiased exponent of a finite non-zero double (subnormals report -1023: always outside SEQ_E_MIN)
SEQ_HD int seq_binade(double v) { return (int)((seq_bits(v) >> 52) & 0x7ffu) - 1023; }
SEQ_HD double seq_pow2(int e) { return seq_from_bits((uint64_t)(e + 1023) << 52); }
SEQ_HD bool seq_finite(double v) { return ((seq_bits(v) >> 52) & 0x7ffu) != 0x7ffu; }

SEQ_HD seq_fn seq_identity()
{
	seq_fn f;
	f.d0 = 0.;
	f.d1 = 0.;
	f.q = 2u;
	return f;
}
// rounding constants of binade e: C0 = 1.5 * 2^e (even mantissa), C1 = C0 + ulp (odd mantissa)
SEQ_HD void seq_consts(int e, double *C0, double *C1)
{
	const uint64_t b = ((uint64_t)(e + 1023) << 52) | (1ull << 51);
	*C0 = seq_from_bits(b);
	*C1 = seq_from_bits(b | 1ull);
}
// function of one product x in binade e; needs |x| <= 2^(e-2)
SEQ_HD seq_fn seq_elem(double x, double C0, double C1)
{
	seq_fn f;
#ifdef __CUDA_ARCH__
	const double t0 = __dadd_rn(x, C0), t1 = __dadd_rn(x, C1);
	f.d0 = __dadd_rn(t0, -C0);
	f.d1 = __dadd_rn(t1, -C1);
#else
	volatile double t0 = x + C0, t1 = x + C1;  // volatile: no re-association by the host compiler
	f.d0 = t0 - C0;
	f.d1 = t1 - C1;
#endif
	f.q = (unsigned)(seq_bits(t0) & 1ull) | ((unsigned)(seq_bits(t1) & 1ull) << 1);
	return f;
}
// h = g after f
SEQ_HD seq_fn seq_compose(const seq_fn &f, const seq_fn &g)
{
	seq_fn h;
	const unsigned a = f.q & 1u, b = (f.q >> 1) & 1u;
	h.d0 = f.d0 + (a ? g.d1 : g.d0);
	h.d1 = f.d1 + (b ? g.d1 : g.d0);
	h.q = ((g.q >> a) & 1u) | (((g.q >> b) & 1u) << 1);
	return h;
}
// S <- f(S); S must lie in the function's binade (checked by the caller)
SEQ_HD double seq_apply(double S, const seq_fn &f) { return S + ((seq_bits(S) & 1ull) ? f.d1 : f.d0); }

// error margin of an approximate prefix after n products: |approx - sequential fp64 sum| <= Kf * (sum of |x| so far).
// Sequential summation of n terms is off the real sum by at most gamma_(n-1) * sum|x|, gamma_k = k u / (1 - k u),
// u = 2^-53; the approximate prefix (per-thread partial sums of <= 2^13 terms combined by trees) by less than 2^14 u sum|x|.
// (n + 2^20) u covers both for every n < 2^32: n u * n u < 2^20 u.
SEQ_HD double seq_margin_factor(int64_t n) { return ((double)n + 1048576.0) * 1.1102230246251565e-16; }

// Is a range of products (sum of magnitudes sa) whose running sum starts near P (approximate, |error| <= m)
// guaranteed to stay inside binade(P), with every product small enough for seq_elem?
// T = approximate sum of magnitudes before the range; kf = seq_margin_factor.
SEQ_HD bool seq_range_safe(double P, double T, double sa, double kf, int *e_out)
{
	const double ap = seq_abs(P);
	const int e = seq_binade(ap);
	*e_out = e;
	if (!(e >= SEQ_E_MIN && e <= SEQ_E_MAX)) return false;  // also catches P == 0, NaN, Inf
	const double lo = seq_pow2(e), hi = seq_pow2(e + 1);
	const double m = kf * (T + sa) + sa;
	// (written so that NaNs fail)
	return (ap - m >= lo) && (ap + m < hi) && (sa <= 0.25 * lo);
}
