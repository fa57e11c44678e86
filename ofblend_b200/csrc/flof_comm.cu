// flof_comm.cu -- multi-GPU plumbing of the t-sharded solve: one process (and one flof_ctx) per
// GPU, NCCL over NVLink 5 / NVSwitch.
//
// The reference has no communication layer at all (single process).  Sharding scheme
// (SURVEY §8e): the 4D grid is cut along t, its slowest stride (grid4d.h:92-97), so a t-slab is one
// contiguous byte range.  Every rank keeps the grids of a sharded level in full size and global index
// space and owns the slices [ta, tb); exchanging a halo or gathering the slabs is therefore a plain
// contiguous NCCL send/recv or an in-place all-gather at identical offsets on every rank:
//   * halo(h):      stencil sweeps (CG apply h=1, extrapolation blur h=1, Gaussian blur h=S)
//   * all-gather:   sources of the semi-Lagrangian gathers (displacements exceed a slab) and the
//                   deformation when a level is left
//   * all-reduce:   fp64 dot products / error sums, fp32 max / min (device scalars, stream ordered)
// All calls are enqueued on the context stream; nothing here synchronises the host.
#include <nccl.h>
#include <stdlib.h>

#include "flof_common.cuh"

#define FLOF_NCCL(call)                                                                                     \
	do {                                                                                                    \
		ncclResult_t r__ = (call);                                                                          \
		if (r__ != ncclSuccess)                                                                             \
			return flof_fail(ctx, FLOF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, ncclGetErrorString(r__), __FILE__, __LINE__); \
	} while (0)

extern "C" int flof_comm_unique_id(char out[FLOF_COMM_ID_BYTES])
{
	static_assert(sizeof(ncclUniqueId) <= FLOF_COMM_ID_BYTES, "unique id size");
	ncclUniqueId id;
	if (ncclGetUniqueId(&id) != ncclSuccess) return FLOF_ERR_CUDA;
	memset(out, 0, FLOF_COMM_ID_BYTES);
	memcpy(out, &id, sizeof(id));
	return FLOF_OK;
}

extern "C" int flof_ctx_comm_init(flof_ctx *ctx, int nranks, int rank, const char id[FLOF_COMM_ID_BYTES])
{
	FLOF_ARG(nranks >= 1 && rank >= 0 && rank < nranks, "flof_ctx_comm_init: bad rank %d of %d", rank, nranks);
	FLOF_ARG(ctx->comm == NULL, "flof_ctx_comm_init: communicator already initialised");
	ncclUniqueId uid;
	memcpy(&uid, id, sizeof(uid));
	ncclComm_t c;
	FLOF_CK(cudaSetDevice(ctx->device));
	FLOF_NCCL(ncclCommInitRank(&c, nranks, uid, rank));
	ctx->comm = (void *)c;
	ctx->rank = rank;
	ctx->nranks = nranks;
	return FLOF_OK;
}

static void flof_p2p_release(flof_ctx *ctx);

extern "C" int flof_ctx_comm_destroy(flof_ctx *ctx)
{
	if (ctx->comm) {
		cudaStreamSynchronize(ctx->stream);
		flof_p2p_release(ctx);
		ncclCommDestroy((ncclComm_t)ctx->comm);
		ctx->comm = NULL;
	}
	ctx->nranks = 1;
	ctx->rank = 0;
	return FLOF_OK;
}

extern "C" int flof_ctx_rank(flof_ctx *ctx) { return ctx->rank; }
extern "C" int flof_ctx_nranks(flof_ctx *ctx) { return ctx->nranks > 0 ? ctx->nranks : 1; }

extern "C" void flof_slab_range(int nt, int nranks, int rank, int *ta, int *tb)
{  // equal contiguous slabs; callers shard only when nt % nranks == 0
	const int per = nt / nranks;
	*ta = rank * per;
	*tb = (rank == nranks - 1) ? nt : (rank + 1) * per;
}

// host-visible collectives for bench.py (barrier, max of a timing over ranks)
extern "C" int flof_comm_barrier(flof_ctx *ctx)
{
	FLOF_CK(cudaStreamSynchronize(ctx->stream));
	if (!ctx->comm || ctx->nranks <= 1) return FLOF_OK;
	FLOF_NCCL(ncclAllReduce(&ctx->red->out_i[2], &ctx->red->out_i[2], 1, ncclInt, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
	FLOF_CK(cudaStreamSynchronize(ctx->stream));
	return FLOF_OK;
}
extern "C" int flof_comm_allreduce_max_host(flof_ctx *ctx, double *v)
{
	if (!ctx->comm || ctx->nranks <= 1) return FLOF_OK;
	double *h = (double *)ctx->pinned;
	h[0] = *v;
	FLOF_CK(cudaMemcpyAsync(&ctx->red->out_d[3], h, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	FLOF_NCCL(ncclAllReduce(&ctx->red->out_d[3], &ctx->red->out_d[3], 1, ncclDouble, ncclMax, (ncclComm_t)ctx->comm, ctx->stream));
	FLOF_CK(cudaMemcpyAsync(h, &ctx->red->out_d[3], sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	FLOF_CK(cudaStreamSynchronize(ctx->stream));
	*v = h[0];
	return FLOF_OK;
}

// ---- NVLink peer mailboxes -------------------------------------------------------------------
// NCCL send/recv moved a 33 MB ghost slice in 0.45 ms and a scalar all-reduce cost 27-44 us (profiles/r1,
// 8 GPUs) -- a quarter of the sharded 128^4 step.  On an NVSwitch box every GPU can store straight into
// every peer's memory, so the two latency-critical exchanges are hand-written instead:
//   halo:       k_halo_push stores this rank's boundary slices into the neighbours' mailboxes and then raises
//               a system-scope flag; k_halo_pull waits for the flag and copies mailbox -> ghost slices.
//   all-reduce: one block stores its partial into every rank's mailbox slot, waits for all slots of the
//               same sequence number and reduces them in rank order (bit-identical on every rank).
// Mailboxes are double-buffered by the parity of the sequence number.  No acknowledgement is needed:
// a rank starts exchange s only after it has finished exchange s-1, which needed every partner's push of
// s-1, which that partner issued (stream order) after its own pull of s-2 -- so buffer s&1 is free again.
// Every spin has a clock64 time-out that raises p2p.dev.err instead of hanging the GPU.
// NCCL stays for bootstrap (handle exchange), barriers and the bulk all-gathers.
#include "flof_p2p.cuh"

__global__ void __launch_bounds__(FLOF_BLOCK)
    k_halo_push(flof_p2p_dev pp, const uint4 *__restrict__ lo_src, const uint4 *__restrict__ hi_src, size_t n16, size_t cap,
                unsigned int seq, unsigned int *counter)
{
	const unsigned int par = seq & 1u;
	const bool has_lo = pp.rank > 0, has_hi = pp.rank < pp.nranks - 1;
	// my first slices are the lower neighbour's "from rank+1" data, my last slices the upper neighbour's "from rank-1"
	uint4 *to_lo = has_lo ? (uint4 *)(pp.peer[pp.rank - 1] + mbox_buf_off(cap, 1, par)) : NULL;
	uint4 *to_hi = has_hi ? (uint4 *)(pp.peer[pp.rank + 1] + mbox_buf_off(cap, 0, par)) : NULL;
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
		if (has_lo) to_lo[i] = __ldg(lo_src + i);
		if (has_hi) to_hi[i] = __ldg(hi_src + i);
	}
	// "last block raises the flags": every thread fences its peer stores system-wide before the ticket
	__shared__ bool s_last;
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) {
		const unsigned int ticket = atomicAdd(counter, 1u);
		s_last = (ticket == gridDim.x - 1);
		if (s_last) *counter = 0;
	}
	__syncthreads();
	if (s_last && threadIdx.x == 0) {
		__threadfence_system();
		if (has_lo) *(volatile unsigned int *)&((flof_mbox_hdr *)pp.peer[pp.rank - 1])->halo_flag[1][par] = seq;
		if (has_hi) *(volatile unsigned int *)&((flof_mbox_hdr *)pp.peer[pp.rank + 1])->halo_flag[0][par] = seq;
	}
}

__global__ void __launch_bounds__(FLOF_BLOCK)
    k_halo_pull(flof_p2p_dev pp, uint4 *__restrict__ lo_ghost, uint4 *__restrict__ hi_ghost, size_t n16, size_t cap,
                unsigned int seq)
{
	const unsigned int par = seq & 1u;
	const bool has_lo = pp.rank > 0, has_hi = pp.rank < pp.nranks - 1;
	char *me = pp.peer[pp.rank];
	__shared__ bool s_ok;
	if (threadIdx.x == 0) {
		bool ok = true;
		flof_mbox_hdr *h = (flof_mbox_hdr *)me;
		if (has_lo) ok = p2p_wait(&h->halo_flag[0][par], seq, pp.err) && ok;
		if (has_hi) ok = p2p_wait(&h->halo_flag[1][par], seq, pp.err) && ok;
		s_ok = ok;
	}
	__syncthreads();
	if (!s_ok) return;
	const uint4 *from_lo = (const uint4 *)(me + mbox_buf_off(cap, 0, par));
	const uint4 *from_hi = (const uint4 *)(me + mbox_buf_off(cap, 1, par));
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
		if (has_lo) lo_ghost[i] = __ldcv(from_lo + i);  // written by a peer while this kernel may already run: no L1
		if (has_hi) hi_ghost[i] = __ldcv(from_hi + i);
	}
}

// stand-alone form on device scalars: kind 0 = n doubles (sum), 1 = n floats (max), 2 = n floats (min)
__global__ void k_p2p_allreduce(flof_p2p_dev pp, void *dev, int n, int kind)
{
	__shared__ double vals[4];
	if ((int)threadIdx.x < n) vals[threadIdx.x] = kind == 0 ? ((double *)dev)[threadIdx.x] : (double)((float *)dev)[threadIdx.x];
	p2p_allreduce_block(pp, vals, n, kind == 0 ? n : 0, kind == 2);
	if ((int)threadIdx.x < n) {
		if (kind == 0)
			((double *)dev)[threadIdx.x] = vals[threadIdx.x];
		else
			((float *)dev)[threadIdx.x] = (float)vals[threadIdx.x];
	}
}

static void flof_p2p_release(flof_ctx *ctx)
{
	if (!ctx->p2p.mbox) return;
	cudaStreamSynchronize(ctx->stream);  // my pulls are done => nobody still writes into my mailbox
	for (int r = 0; r < ctx->nranks && r < FLOF_P2P_MAX; ++r)
		if (r != ctx->rank && ctx->p2p.dev.peer[r]) cudaIpcCloseMemHandle(ctx->p2p.dev.peer[r]);
	// every rank must have unmapped this mailbox before its owner frees it
	if (ctx->comm && ctx->nranks > 1) {
		ncclAllReduce(&ctx->red->out_i[2], &ctx->red->out_i[2], 1, ncclInt, ncclSum, (ncclComm_t)ctx->comm, ctx->stream);
		cudaStreamSynchronize(ctx->stream);
	}
	cudaFree(ctx->p2p.mbox);
	if (ctx->p2p.counter) cudaFree(ctx->p2p.counter);
	memset(&ctx->p2p, 0, sizeof(ctx->p2p));
}

// (re)allocates the mailboxes so that one halo buffer holds `need` bytes; collective over all ranks
int flof_p2p_ensure(flof_ctx *ctx, size_t need)
{
	if (ctx->opt.no_p2p || ctx->nranks > FLOF_P2P_MAX) return FLOF_OK;  // "no_p2p": flof_ctx_set_option / FLOF_NO_P2P
	if (ctx->p2p.enabled && need <= ctx->p2p.cap) return FLOF_OK;
	if (ctx->p2p.mbox) flof_p2p_release(ctx);
	FLOF_CK(cudaStreamSynchronize(ctx->stream));
	size_t cap = need < ((size_t)1 << 20) ? ((size_t)1 << 20) : need;
	cap = (cap + 255) & ~(size_t)255;
	const size_t bytes = (size_t)FLOF_MBOX_HDR_BYTES + 4 * cap;
	FLOF_CK(cudaMalloc((void **)&ctx->p2p.mbox, bytes));
	FLOF_CK(cudaMemset(ctx->p2p.mbox, 0, FLOF_MBOX_HDR_BYTES));
	FLOF_CK(cudaMalloc((void **)&ctx->p2p.counter, 8 * sizeof(unsigned int)));
	FLOF_CK(cudaMemset(ctx->p2p.counter, 0, 8 * sizeof(unsigned int)));
	FLOF_CK(cudaDeviceSynchronize());
	// exchange the IPC handles through NCCL (device all-gather of 64-byte records)
	cudaIpcMemHandle_t mine;
	FLOF_CK(cudaIpcGetMemHandle(&mine, ctx->p2p.mbox));
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
	void *hd = NULL;
	FLOF_CK(cudaMalloc(&hd, 64 * (size_t)ctx->nranks));
	FLOF_CK(cudaMemcpyAsync((char *)hd + 64 * (size_t)ctx->rank, &mine, 64, cudaMemcpyHostToDevice, ctx->stream));
	FLOF_NCCL(ncclAllGather((char *)hd + 64 * (size_t)ctx->rank, hd, 64, ncclChar, (ncclComm_t)ctx->comm, ctx->stream));
	cudaIpcMemHandle_t all[FLOF_P2P_MAX];
	FLOF_CK(cudaMemcpyAsync(all, hd, 64 * (size_t)ctx->nranks, cudaMemcpyDeviceToHost, ctx->stream));
	FLOF_CK(cudaStreamSynchronize(ctx->stream));
	cudaFree(hd);
	memset(&ctx->p2p.dev, 0, sizeof(ctx->p2p.dev));
	int ok = 1;
	for (int r = 0; r < ctx->nranks; ++r) {
		if (r == ctx->rank) {
			ctx->p2p.dev.peer[r] = ctx->p2p.mbox;
			continue;
		}
		void *pm = NULL;
		if (cudaIpcOpenMemHandle(&pm, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
			cudaGetLastError();
			ok = 0;
			break;
		}
		ctx->p2p.dev.peer[r] = (char *)pm;
	}
	// all ranks must agree: one failed mapping disables the peer path everywhere (NCCL is used instead)
	int *flag = &ctx->red->out_i[3];
	FLOF_CK(cudaMemcpyAsync(flag, &ok, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
	FLOF_NCCL(ncclAllReduce(flag, flag, 1, ncclInt, ncclMin, (ncclComm_t)ctx->comm, ctx->stream));
	FLOF_CK(cudaMemcpyAsync(&ok, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	FLOF_CK(cudaStreamSynchronize(ctx->stream));
	ctx->p2p.dev.rank = ctx->rank;
	ctx->p2p.dev.nranks = ctx->nranks;
	ctx->p2p.dev.err = ctx->p2p.counter + 1;
	ctx->p2p.dev.ar_seq = ctx->p2p.counter + 2;
	ctx->p2p.dev.chain_seq = ctx->p2p.counter + 3;
	ctx->p2p.cap = cap;
	ctx->p2p.halo_seq = 0;
	ctx->p2p.enabled = ok;
	if (!ok) ctx->opt.no_p2p = 1;  // a mapping failed on some rank (agreed by all-reduce above): NCCL from now on
	return FLOF_OK;
}

// 1 if a spin-wait of the peer path timed out since the last check (host, synchronises)
extern "C" int flof_comm_p2p_status(flof_ctx *ctx, int *enabled, int *timed_out)
{
	if (enabled) *enabled = ctx->p2p.enabled;
	if (timed_out) *timed_out = 0;
	if (!ctx->p2p.enabled) return FLOF_OK;
	unsigned int e = 0;
	FLOF_CK(cudaMemcpyAsync(&e, ctx->p2p.dev.err, sizeof(e), cudaMemcpyDeviceToHost, ctx->stream));
	FLOF_CK(cudaStreamSynchronize(ctx->stream));
	if (timed_out) *timed_out = (int)e;
	if (e) return flof_fail(ctx, FLOF_ERR_CUDA, "peer mailbox wait timed out (ranks out of step?)");
	return FLOF_OK;
}

// ---- internal helpers used by the sharded operators -----------------------------------------
int flof_halo_exchange(flof_ctx *ctx, void *grid, int nt, size_t slice_bytes, int h)
{
	if (!flof_sharded(ctx, nt) || ctx->nranks <= 1 || h <= 0) return FLOF_OK;
	const int ta = ctx->sh.ta, tb = ctx->sh.tb;
	FLOF_ARG(tb - ta >= h, "halo of %d slices exceeds the slab thickness %d", h, tb - ta);
	char *g = (char *)grid;
	const size_t n = slice_bytes * (size_t)h;
	FLOF_RET(flof_p2p_ensure(ctx, n));
	if (ctx->p2p.enabled && n % 16 == 0) {
		const unsigned int seq = ++ctx->p2p.halo_seq;
		const size_t n16 = n / 16;
		size_t want = (n16 + FLOF_BLOCK - 1) / FLOF_BLOCK;
		const size_t cap_blocks = (size_t)ctx->sm_count * 4;
		const unsigned blocks = (unsigned)(want < cap_blocks ? (want ? want : 1) : cap_blocks);
		FLOF_LAUNCH(k_halo_push, blocks, FLOF_BLOCK, 0, ctx->p2p.dev, (const uint4 *)(g + slice_bytes * (size_t)ta),
		            (const uint4 *)(g + slice_bytes * (size_t)(tb - h)), n16, ctx->p2p.cap, seq, ctx->p2p.counter);
		FLOF_LAUNCH(k_halo_pull, blocks, FLOF_BLOCK, 0, ctx->p2p.dev, (uint4 *)(g + slice_bytes * (size_t)(ta - h)),
		            (uint4 *)(g + slice_bytes * (size_t)tb), n16, ctx->p2p.cap, seq);
		return FLOF_OK;
	}
	ncclComm_t c = (ncclComm_t)ctx->comm;
	const int pi = flof_prof_pre(ctx, "nccl_halo_sendrecv");
	FLOF_NCCL(ncclGroupStart());
	if (ctx->rank > 0) {
		FLOF_NCCL(ncclSend(g + slice_bytes * (size_t)ta, n, ncclChar, ctx->rank - 1, c, ctx->stream));
		FLOF_NCCL(ncclRecv(g + slice_bytes * (size_t)(ta - h), n, ncclChar, ctx->rank - 1, c, ctx->stream));
	}
	if (ctx->rank < ctx->nranks - 1) {
		FLOF_NCCL(ncclSend(g + slice_bytes * (size_t)(tb - h), n, ncclChar, ctx->rank + 1, c, ctx->stream));
		FLOF_NCCL(ncclRecv(g + slice_bytes * (size_t)tb, n, ncclChar, ctx->rank + 1, c, ctx->stream));
	}
	FLOF_NCCL(ncclGroupEnd());
	flof_prof_post(ctx, pi);
	return FLOF_OK;
}

int flof_allgather_slabs(flof_ctx *ctx, void *grid, int nt, size_t slice_bytes)
{
	if (!flof_sharded(ctx, nt) || ctx->nranks <= 1) return FLOF_OK;
	char *g = (char *)grid;
	const size_t n = slice_bytes * (size_t)(ctx->sh.tb - ctx->sh.ta);
	const int pi = flof_prof_pre(ctx, "nccl_allgather_slabs");
	FLOF_NCCL(ncclAllGather(g + slice_bytes * (size_t)ctx->sh.ta, g, n, ncclChar, (ncclComm_t)ctx->comm, ctx->stream));
	flof_prof_post(ctx, pi);
	return FLOF_OK;
}

// in-place all-gather of equal byte ranges: this rank's part sits at buf + off (off = rank * n)
int flof_allgather_bytes(flof_ctx *ctx, void *buf, size_t off, size_t n)
{
	if (!ctx->comm || ctx->nranks <= 1) return FLOF_OK;
	const int pi = flof_prof_pre(ctx, "nccl_allgather_inputs");
	FLOF_NCCL(ncclAllGather((char *)buf + off, buf, n, ncclChar, (ncclComm_t)ctx->comm, ctx->stream));
	flof_prof_post(ctx, pi);
	return FLOF_OK;
}

static int flof_allreduce_scalar(flof_ctx *ctx, void *dev, int n, int kind)
{
	if (!ctx->comm || ctx->nranks <= 1) return FLOF_OK;
	FLOF_RET(flof_p2p_ensure(ctx, 0));
	if (ctx->p2p.enabled && n <= 4) {
		FLOF_LAUNCH(k_p2p_allreduce, 1, 32, 0, ctx->p2p.dev, dev, n, kind);
		return FLOF_OK;
	}
	const int pi = flof_prof_pre(ctx, "nccl_allreduce_scalar");
	if (kind == 0)
		FLOF_NCCL(ncclAllReduce(dev, dev, n, ncclDouble, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
	else
		FLOF_NCCL(ncclAllReduce(dev, dev, n, ncclFloat, kind == 1 ? ncclMax : ncclMin, (ncclComm_t)ctx->comm, ctx->stream));
	flof_prof_post(ctx, pi);
	return FLOF_OK;
}
int flof_allreduce_f64_sum(flof_ctx *ctx, double *dev, int n) { return flof_allreduce_scalar(ctx, dev, n, 0); }
int flof_allreduce_f32_max(flof_ctx *ctx, float *dev, int n) { return flof_allreduce_scalar(ctx, dev, n, 1); }
int flof_allreduce_f32_min(flof_ctx *ctx, float *dev, int n) { return flof_allreduce_scalar(ctx, dev, n, 2); }
