// flof_p2p.cuh -- device side of the NVLink peer mailboxes (see flof_comm.cu): spin-wait with time-out and the
// in-kernel all-reduce that the CG kernels call from their "last block" tail, so that the reduction of an
// iteration's dot product and its exchange over NVLink are one kernel (no separate collective launch).
#pragma once
#include "flof_common.cuh"

#define FLOF_SPIN_LIMIT (20000000000ll)  // ~10 s of SM clocks: host-side skew between ranks (H2D of 6 GB inputs) stays far below

// The error word is sticky: after the first time-out every later wait returns at once, so a rank that lost its
// partner finishes its launch queue quickly and the host reports the failure (flof_comm_p2p_status).
__device__ __forceinline__ bool p2p_wait(volatile unsigned int *flag, unsigned int seq, unsigned int *err)
{
	const long long t0 = clock64();
	while (*flag != seq) {
		if (*(volatile unsigned int *)err) return false;
		if (clock64() - t0 > FLOF_SPIN_LIMIT) {
			atomicExch(err, 1u);
			return false;
		}
	}
	__threadfence_system();
	return true;
}
__device__ __forceinline__ size_t mbox_buf_off(size_t cap, int from, unsigned int par)
{
	return (size_t)FLOF_MBOX_HDR_BYTES + (size_t)(from * 2 + (int)par) * cap;
}

// all-reduce of n <= 4 doubles held in shared vals[] (called by every thread of ONE block with >= nranks threads;
// result in vals[], valid for every thread after the call).  vals[0..nsum) are summed in rank order -- the same
// order on every rank, so all ranks get the identical bits and stay in lock step --, the rest max (or min).
// excl (optional, shared, >= nsum doubles): the sum over the LOWER ranks only (exclusive prefix in rank order).
__device__ __forceinline__ void p2p_allreduce_block(const flof_p2p_dev &pp, double *vals, int n, int nsum, bool use_min,
                                                    double *excl = NULL)
{
	// the sequence number lives on the device and advances only when an all-reduce really runs (the CG kernels
	// return early once the solve is done), so consecutive exchanges always alternate the mailbox parity
	__shared__ unsigned int s_seq;
	if (threadIdx.x == 0) s_seq = ++(*pp.ar_seq);
	__syncthreads();
	const unsigned int seq = s_seq, par = seq & 1u;
	const int j = (int)threadIdx.x;
	flof_mbox_hdr *me = (flof_mbox_hdr *)pp.peer[pp.rank];
	if (j < pp.nranks) {
		flof_mbox_hdr *h = (flof_mbox_hdr *)pp.peer[j];
		volatile double *dv = h->ar[par][pp.rank].v;
		for (int q = 0; q < n; ++q) dv[q] = vals[q];
		__threadfence_system();
		*(volatile unsigned int *)&h->ar[par][pp.rank].seq = seq;
		p2p_wait(&me->ar[par][j].seq, seq, pp.err);
	}
	__syncthreads();
	if (j == 0) {
		for (int q = 0; q < n; ++q) {
			double acc = ((volatile double *)me->ar[par][0].v)[q];
			if (excl && q < nsum) excl[q] = 0.;
			for (int r = 1; r < pp.nranks; ++r) {
				if (excl && q < nsum && r == pp.rank) excl[q] = acc;
				const double x = ((volatile double *)me->ar[par][r].v)[q];
				acc = q < nsum ? acc + x : (use_min ? fmin(acc, x) : fmax(acc, x));
			}
			vals[q] = acc;
		}
	}
	__syncthreads();
}

