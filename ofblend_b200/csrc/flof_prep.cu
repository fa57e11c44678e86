// flof_prep.cu -- input pre-processing of the 4D SDFs: start-frame repeat and the simple
// layer-by-layer extrapolation with its integer marker grid.
// ref: repeatFrame4d :1254-1266, nbs4d :1323-1326, knExtrap4dLsSimple :1330-1352,
//      knSetRemaining4d :1354-1359, extrap4dLsSimple :1361-1406, extrapolateVec4Simple :1408-1439.
//
// The extrapolation is a BFS over Manhattan layers: sweep d gives every unmarked cell that has a
// neighbour in layer d the mean of those neighbours (+ direction) and marks it d+1.  A sweep only
// reads cells of layer d (never written in that sweep) and only writes unmarked cells, so the
// in-place update is order independent: the parallel result -- float values and the Grid4d<int>
// marker -- is bit-identical to the sequential reference.  39 sweeps of 16 B/cell.
#include <math.h>

#include "flof_common.cuh"

// ------------------------------------------------------------------ repeatFrame4d ----------
// phi(i,j,k,t) = phi.interp((i,j,k,srct)+0.5) for t in [int(srct-range+0.5), int(srct)].
// The interpolated source slice is evaluated once into a 3D scratch buffer, then copied to
// every target slice.  (The reference writes slice int(srct) last, in place; each cell then only
// reads its own not-yet-overwritten value with weight 1 and finite neighbours with weight 0.)
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_repeat_eval(const float *__restrict__ phi, float *__restrict__ slice, flof_dim4 d, float srct)
{
	const unsigned p = blockIdx.x * FLOF_BLOCK + threadIdx.x;
	if (p >= (unsigned)(d.nx * d.ny)) return;
	const int j = p / d.nx, i = p - j * d.nx, k = blockIdx.y;
	slice[(int64_t)i + (int64_t)d.nx * (j + (int64_t)d.ny * k)] =
	    flof_interpol4d<float>(phi, d, (float)i + 0.5f, (float)j + 0.5f, (float)k + 0.5f, srct + 0.5f);
}
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_repeat_store(float *__restrict__ phi, const float *__restrict__ slice, flof_dim4 d, int t0, int bnd)
{
	const unsigned p = blockIdx.x * FLOF_BLOCK + threadIdx.x;
	if (p >= (unsigned)(d.nx * d.ny)) return;
	const int j = p / d.nx, i = p - j * d.nx, k = blockIdx.y, t = t0 + (int)blockIdx.z;
	if (i < bnd || j < bnd || k < bnd || i >= d.nx - bnd || j >= d.ny - bnd || k >= d.nz - bnd) return;
	phi[flof_idx(d, i, j, k, t)] = slice[(int64_t)i + (int64_t)d.nx * (j + (int64_t)d.ny * k)];
}

extern "C" int flof_repeat_frame4d(flof_ctx *ctx, float *phi, flof_dim4 d, float srct, float range, int bnd)
{
	const float sr = srct - range;  // fp32 subtraction, then + 0.5 in double (ref :1258)
	const int t0 = (int)((double)sr + 0.5), t1 = (int)srct;
	if (t1 < t0) return FLOF_OK;
	FLOF_ARG(t0 >= 0 && t1 < d.nt, "repeatFrame4d: slice range [%d,%d] outside the grid (T=%d)", t0, t1, d.nt);
	void *slice = NULL;
	FLOF_RET(flof_tmp_alloc(ctx, &slice, sizeof(float) * (size_t)d.nx * d.ny * d.nz, false));
	dim3 g3((unsigned)(((int64_t)d.nx * d.ny + FLOF_BLOCK - 1) / FLOF_BLOCK), (unsigned)d.nz, 1);
	FLOF_LAUNCH(k_repeat_eval, g3, FLOF_BLOCK, 0, phi, (float *)slice, d, srct);
	dim3 g4(g3.x, g3.y, (unsigned)(t1 - t0 + 1));
	FLOF_LAUNCH(k_repeat_store, g4, FLOF_BLOCK, 0, phi, (const float *)slice, d, t0, bnd);
	return flof_tmp_free(ctx, slice);
}

// ------------------------------------------------------------------ extrapolation ----------
// MODE 0: tmp = 1 where phi < 0 ; MODE 1: tmp = 1 where phi > 0   (interior, bnd 1)  ref :1368-1385
template <int MODE>
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_extrap_mark_inside(const float *__restrict__ phi, int *__restrict__ tmp, flof_dim4 d)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t)) return;
	const int64_t c = flof_idx(d, i, j, k, t);
	int m = 0;
	if (flof_in_bounds(d, i, j, k, t, 1)) {
		const float v = __ldg(phi + c);
		if (MODE == 0 && v < 0.f) m = 1;
		if (MODE == 1 && v > 0.f) m = 1;
	}
	tmp[c] = m;  // also performs tmp.clear()
}
// first ring: unmarked interior cells with a neighbour marked 1 become 2   ref :1386-1397
__global__ void __launch_bounds__(FLOF_BLOCK) k_extrap_first_ring(int *tmp, flof_dim4 d)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t) || !flof_in_bounds(d, i, j, k, t, 1)) return;
	const int64_t c = flof_idx(d, i, j, k, t);
	// In-place update: neighbours are read while other threads may store a 2 into them.  A concurrently written 2 never
	// equals 1, so the outcome does not depend on the interleaving; the accesses are volatile (single, untorn 32-bit
	// loads / stores the compiler may neither cache nor split), which makes that argument hold formally as well.
	volatile int *vt = tmp;
	if (vt[c]) return;
	const int64_t sY = d.nx, sZ = (int64_t)d.nx * d.ny, sT = sZ * d.nz;
	if (vt[c - 1] == 1 || vt[c + 1] == 1 || vt[c - sY] == 1 || vt[c + sY] == 1 || vt[c - sZ] == 1 ||
	    vt[c + sZ] == 1 || vt[c - sT] == 1 || vt[c + sT] == 1)
		vt[c] = 2;
}
// ref knExtrap4dLsSimple :1330-1352; neighbour order nbs4d: -x,+x,-y,+y,-z,+z,-t,+t
template <class T>
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_extrap_sweep(T *val, int *tmp, flof_dim4 d, int dd, T direction)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t) || !flof_in_bounds(d, i, j, k, t, 1)) return;
	const int64_t c = flof_idx(d, i, j, k, t);
	// In-place sweep (like the reference's kernel): only cells with marker 0 are written (marker dd + 1, new value), only
	// neighbours with marker == dd are read -- those are not written in this sweep, and a concurrently stored dd + 1
	// never equals dd.  Marker accesses are volatile (untorn 32-bit loads / stores) so this holds formally.
	volatile int *vt = tmp;
	if (vt[c] != 0) return;
	const int64_t sY = d.nx, sZ = (int64_t)d.nx * d.ny, sT = sZ * d.nz;
	const int64_t o[8] = { -1, 1, -sY, sY, -sZ, sZ, -sT, sT };
	int nbs = 0;
	T avg;
	if constexpr (sizeof(T) == 4) avg = 0.f; else avg = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
	for (int n = 0; n < 8; ++n) {
		if (vt[c + o[n]] == dd) {
			const T q = val[c + o[n]];
			if constexpr (sizeof(T) == 4) { avg += q; } else { avg.x += q.x; avg.y += q.y; avg.z += q.z; avg.w += q.w; }
			nbs++;
		}
	}
	if (nbs > 0) {
		vt[c] = dd + 1;
		const float fn = (float)nbs;  // avg / nbs: int promoted to float
		if constexpr (sizeof(T) == 4) {
			val[c] = avg / fn + direction;
		} else {
			val[c] = make_float4(avg.x / fn + direction.x, avg.y / fn + direction.y, avg.z / fn + direction.z,
			                     avg.w / fn + direction.w);
		}
	}
}
// ref knSetRemaining4d :1354-1359
template <class T>
__global__ void __launch_bounds__(FLOF_BLOCK) k_extrap_set_remaining(T *__restrict__ phi, const int *__restrict__ tmp, flof_dim4 d, T set)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t) || !flof_in_bounds(d, i, j, k, t, 1)) return;
	const int64_t c = flof_idx(d, i, j, k, t);
	if (tmp[c] != 0) return;
	phi[c] = set;
}

extern "C" int flof_extrap4d_ls_simple(flof_ctx *ctx, float *phi, flof_dim4 d, int distance, int inside,
                                       int *marker)
{
	FLOF_ARG(d.nx >= 3 && d.ny >= 3 && d.nz >= 3 && d.nt >= 3, "extrap4dLsSimple: grid too small");
	const int64_t cells = flof_cells(d);
	int *tmp = marker;
	void *own = NULL;
	if (!tmp) {
		FLOF_RET(flof_tmp_alloc(ctx, &own, sizeof(int) * (size_t)cells, false));
		tmp = (int *)own;
	}
	const float direction = inside ? -1.f : 1.f;
	if (inside)
		FLOF_LAUNCH(k_extrap_mark_inside<1>, flof_grid4(d), FLOF_BLOCK, 0, phi, tmp, d);
	else
		FLOF_LAUNCH(k_extrap_mark_inside<0>, flof_grid4(d), FLOF_BLOCK, 0, phi, tmp, d);
	FLOF_LAUNCH(k_extrap_first_ring, flof_grid4(d), FLOF_BLOCK, 0, tmp, d);
	for (int dd = 2; dd < 1 + distance; ++dd)
		FLOF_LAUNCH(k_extrap_sweep<float>, flof_grid4(d), FLOF_BLOCK, 0, phi, tmp, d, dd, direction);
	FLOF_LAUNCH(k_extrap_set_remaining<float>, flof_grid4(d), FLOF_BLOCK, 0, phi, (const int *)tmp, d,
	            (float)(direction * (float)(distance + 2)));
	return flof_tmp_free(ctx, own);
}

extern "C" int flof_extrapolate_vec4_simple(flof_ctx *ctx, float *vel, const float *phi, flof_dim4 d, int distance)
{
	FLOF_ARG(d.nx >= 3 && d.ny >= 3 && d.nz >= 3 && d.nt >= 3, "extrapolateVec4Simple: grid too small");
	const int64_t cells = flof_cells(d);
	void *own = NULL;
	FLOF_RET(flof_tmp_alloc(ctx, &own, sizeof(int) * (size_t)cells, false));
	int *tmp = (int *)own;
	const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
	FLOF_LAUNCH(k_extrap_mark_inside<0>, flof_grid4(d), FLOF_BLOCK, 0, phi, tmp, d);
	FLOF_LAUNCH(k_extrap_first_ring, flof_grid4(d), FLOF_BLOCK, 0, tmp, d);
	for (int dd = 2; dd < distance + 1; ++dd)
		FLOF_LAUNCH(k_extrap_sweep<float4>, flof_grid4(d), FLOF_BLOCK, 0, (float4 *)vel, tmp, d, dd, zero);
	FLOF_LAUNCH(k_extrap_set_remaining<float4>, flof_grid4(d), FLOF_BLOCK, 0, (float4 *)vel, (const int *)tmp, d, zero);
	return flof_tmp_free(ctx, own);
}
