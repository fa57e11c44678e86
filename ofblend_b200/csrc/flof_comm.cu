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

extern "C" int flof_ctx_comm_destroy(flof_ctx *ctx)
{
	if (ctx->comm) {
		cudaStreamSynchronize(ctx->stream);
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

// ---- internal helpers used by the sharded operators -----------------------------------------
int flof_halo_exchange(flof_ctx *ctx, void *grid, int nt, size_t slice_bytes, int h)
{
	if (!flof_sharded(ctx, nt) || ctx->nranks <= 1 || h <= 0) return FLOF_OK;
	const int ta = ctx->sh.ta, tb = ctx->sh.tb;
	FLOF_ARG(tb - ta >= h, "halo of %d slices exceeds the slab thickness %d", h, tb - ta);
	char *g = (char *)grid;
	ncclComm_t c = (ncclComm_t)ctx->comm;
	const size_t n = slice_bytes * (size_t)h;
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

int flof_allreduce_f64_sum(flof_ctx *ctx, double *dev, int n)
{
	if (!ctx->comm || ctx->nranks <= 1) return FLOF_OK;
	const int pi = flof_prof_pre(ctx, "nccl_allreduce_scalar");
	FLOF_NCCL(ncclAllReduce(dev, dev, n, ncclDouble, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
	flof_prof_post(ctx, pi);
	return FLOF_OK;
}
int flof_allreduce_f32_max(flof_ctx *ctx, float *dev, int n)
{
	if (!ctx->comm || ctx->nranks <= 1) return FLOF_OK;
	const int pi = flof_prof_pre(ctx, "nccl_allreduce_scalar");
	FLOF_NCCL(ncclAllReduce(dev, dev, n, ncclFloat, ncclMax, (ncclComm_t)ctx->comm, ctx->stream));
	flof_prof_post(ctx, pi);
	return FLOF_OK;
}
int flof_allreduce_f32_min(flof_ctx *ctx, float *dev, int n)
{
	if (!ctx->comm || ctx->nranks <= 1) return FLOF_OK;
	const int pi = flof_prof_pre(ctx, "nccl_allreduce_scalar");
	FLOF_NCCL(ncclAllReduce(dev, dev, n, ncclFloat, ncclMin, (ncclComm_t)ctx->comm, ctx->stream));
	flof_prof_post(ctx, pi);
	return FLOF_OK;
}
