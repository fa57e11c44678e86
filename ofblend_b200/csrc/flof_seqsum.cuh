// flof_seqsum.cuh -- device state of the sequential-order dot products (see flof_seqsum_core.h for the arithmetic,
// flof_seqsum.cu for the kernels).  ref: dotProd optflow4d.cpp:234-241.
#pragma once
#include "flof_common.cuh"
#include "flof_seqsum_core.h"

#define SEQ_U 4                          // cells per thread and leaf
#define SEQ_LEAF_CELLS (FLOF_BLOCK * SEQ_U)  // 1024 cells = 4096 products per leaf
#define SEQ_DMAX 1024                    // dirty leaves the resolver can take per dot product
#define SEQ_POOL (1 << 18)               // pieces (32 B each) all dirty leaves of one dot product may use
#define SEQ_PIECE_SMEM 1024              // pieces the resolver stages in shared memory

// decoupled look-back descriptor of one leaf: approximate sum / sum of magnitudes of the leaf (agg) and of
// everything up to and including it (pre).  A part is valid when its stamp equals the epoch of the launch.
struct seq_desc {
	unsigned int st_agg, st_pre;
	double ax, aa, px, pa;
	double pad;
};

struct seq_ctl {  // device-resident control block of one context
	unsigned int ticket;       // next leaf to hand out (reset by the resolver)
	unsigned int ndirty;       // dirty leaves of the running dot product
	unsigned int pool_used;    // pieces allocated from the pool
	unsigned int flags;        // bit 0: non-finite product seen, bit 1: capacity exceeded, bit 2: consistency check failed
	double result;             // last resolved sum (exact bits of the sequential loop)
	unsigned long long n_dots, n_dirty, n_raw, n_pieces, n_fallback, n_inconsistent;  // statistics since context creation
	int dirty_leaf[SEQ_DMAX];
	unsigned int dirty_base[SEQ_DMAX];
	unsigned int dirty_cnt[SEQ_DMAX];
};

struct flof_seq {  // host-side handle (ctx->seq)
	seq_desc *desc;
	seq_rec *leaf;
	seq_rec *pool;
	seq_ctl *ctl;
	int64_t cap_leaves;
	unsigned int epoch;
};
