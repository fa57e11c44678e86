// flof_seqsum.cuh -- device state of the sequential-order dot products (see flof_seqsum_core.h for the arithmetic,
// flof_seqsum_kernels.cuh for the kernels).  ref: dotProd optflow4d.cpp:234-241.
//
// Organisation (two passes, no spinning, identical on one GPU and on a t-sharded level):
//   pass 1  the kernel that PRODUCES the vectors (k_cg_init / k_cg_apply / k_cg_update; k_seq_agg for the stand-alone
//           entry) walks the cells in contiguous SEGMENTS, one CTA per segment, and leaves per segment an approximate
//           sum of its products and an upper bound of the sum of their magnitudes; its "last block" tail turns them into
//           exclusive prefixes (seq_seg) -- on a sharded level after the in-kernel all-reduce, which also yields the
//           lower ranks' share.
//   pass 2  k_dot_seq re-reads the two vectors (32 B/cell).  With the approximate prefix known, a segment whose running sum
//           provably stays inside one binade is folded into ONE rounding function by eight independent warps (no block
//           barrier); the few segments around a binade crossing (or where the sum still builds up from zero) take the
//           careful path: leaf by leaf, leaves that are still unsafe product by product (runs + raw products).
//   resolve k_seq_resolve (one CTA): composes the segment entries in parallel, then one thread walks runs and raw products
//           with real fp64 adds.  On a sharded level the exact running sum travels rank 0 -> 1 -> ... through the peer
//           mailboxes and the last rank broadcasts the total.
#pragma once
#include "flof_common.cuh"
#include "flof_seqsum_core.h"

#define SEQ_U 4                              // cells per thread and leaf
#define SEQ_LEAF_CELLS (FLOF_BLOCK * SEQ_U)  // 1024 cells = 4096 products per leaf
#define SEQ_CHUNK_CELLS 128                  // fast path: cells a warp folds per step (4 per lane)
#define SEQ_ECAP 520                         // entries a segment may emit (fast path: 1)
#define SEQ_DMAX 1024                        // dirty leaves the resolver can take per dot product
#define SEQ_POOL (1 << 18)                   // pieces (32 B each) all dirty leaves of one dot product may use
#define SEQ_STAGE 1536                       // pieces of ONE dirty leaf (k_dot_seq stages them in shared memory)
#define SEQ_RAWLEAF_MIN 1024                 // a dirty leaf that still has more pieces after merging is kept as plain products
#define SEQ_RAWLEAF_RECS (SEQ_LEAF_CELLS * 4 * 4 / 32)  // pool records that hold the 4096 floats of such a leaf
#define SEQ_GSTEPS (1 << 17)                 // walk steps of the resolver's global-memory list (used when SEQ_SMAX is exceeded)
#define SEQ_EMAX 2560                        // segment entries the resolver stages in shared memory
#define SEQ_MAX_SEG 2048                     // segments per rank
#define SEQ_PLAIN_MAX (1 << 21)                // cells up to which the resolver's fallback is the plain one-thread loop
#define SEQ_SA_SLACK 1.001                   // the producers sum |product| in fp32: inflate to a rigorous upper bound

// contiguous decomposition of one rank's cell range into segments (the same for the producing kernel and k_dot_seq)
struct seq_part {
	int ncells;     // cells of the range
	int seg_cells;  // cells per segment, a multiple of SEQ_LEAF_CELLS
	int nseg;       // <= FLOF_MAX_PARTIALS
};

struct seq_seg {  // per segment: approximate sum / magnitude bound of the segment and of everything before it in the range
	double sx, sa, px, pa;
};
struct seq_cls {  // classification of a segment by the tail of pass 1
	int mode;     // SEQ_LEAF_WILD (all zero) / SEQ_LEAF_CLEAN (safe: one binade) / SEQ_LEAF_DIRTY (careful path)
	int e;        // binade of a safe segment
};

struct seq_ctl {  // device-resident control block of one context
	unsigned int ticket;       // next entry of the work list (k_dot_seq), reset by the resolver
	unsigned int ndirty;       // dirty leaves of the running dot product
	unsigned int pool_used;    // pieces allocated from the pool
	unsigned int flags;        // bit 0: non-finite product seen, bit 1: capacity exceeded (bits 8..12: which), bit 2: consistency check failed
	unsigned int why;          // OR of the flags of every fallback since context creation
	double result;             // last resolved sum (exact bits of the sequential loop)
	double tot[2];             // approximate sum / magnitude bound of this rank's range (tail of pass 1)
	double off[2];             // the same for all lower ranks together (0 on a single GPU)
	unsigned long long n_dots, n_dirty, n_raw, n_pieces, n_fallback, n_inconsistent, n_slow_segments, n_inexact, n_rawleaves;  // statistics since context creation
	unsigned long long prof[5];  // resolver: cycles spent gathering / composing / walking / finishing, walk steps (thread 0, summed)
};

struct flof_seq {  // host-side handle (ctx->seq)
	seq_seg *seg;      // [SEQ_MAX_SEG]
	seq_cls *cls;      // [SEQ_MAX_SEG]
	int *order;        // [SEQ_MAX_SEG] work list of pass 2: careful segments first (they take longest), then the safe ones
	seq_rec *ent0;     // [SEQ_MAX_SEG] first entry of every segment
	seq_rec *ent;      // [SEQ_MAX_SEG * SEQ_ECAP] further entries of the segments, in order
	int *ecnt;         // [SEQ_MAX_SEG]
	double *aggx, *agga;  // [SEQ_MAX_SEG] each: per-segment sums of pass 1 (plain stores, or atomics of the stencil kernel); zero between launches
	seq_rec *pool;     // [SEQ_POOL]
	seq_rec *gsteps;   // [SEQ_GSTEPS] walk steps of a dot product that does not fit the resolver's shared memory
	seq_ctl *ctl;
};

// host: segment size for a range of `ncells` cells; slice_cells = cells of one t-slice (0 if unknown).  About SEQ_MAX_SEG
// segments: fine enough for the dynamic scheduling of pass 2, few enough for the one-CTA tail scan and the resolver.
static inline seq_part seq_make_part(int64_t ncells, int64_t slice_cells, int sm_count)
{
	(void)slice_cells;
	(void)sm_count;
	seq_part p;
	const int64_t leaves = (ncells + SEQ_LEAF_CELLS - 1) / SEQ_LEAF_CELLS;
	int64_t R = (leaves + SEQ_MAX_SEG - 1) / SEQ_MAX_SEG;
	if (R < 1) R = 1;
	p.ncells = (int)ncells;
	p.seg_cells = (int)(R * SEQ_LEAF_CELLS);
	p.nseg = (int)((leaves + R - 1) / R);
	return p;
}
