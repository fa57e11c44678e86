// flof_blur_tiled.cuh -- register-tiled versions of the two box-window blurs (Vec4 grids).
// ref: knGaussianBlur optflow4d.cpp:133-158, knCvExpolBlur4d :613-626.
// (included by flof_blur.cu after the definition of c_gauss_w: __constant__ symbols are per
// translation unit without relocatable device code)
//
// Same arithmetic as the generic kernels in flof_blur.cu -- every output cell accumulates its taps
// in the reference's order (vt, zk, yj, xi), fp32 multiply then add -- so results stay bit-identical
// to the CPU.  What changes is the data path.  The generic kernels issue one LDG.128 per tap
// ((2S+1)^4 per cell) and saturate the L1 data path (measured: 122 B/clk/SM, ncu profiles/r1).
// Here the lanes of a warp stay on consecutive x (every load is a unit-stride, fully coalesced
// 512 B request) and each thread produces PY output cells along y, keeping the source rows it
// needs in registers: a row (yj, zk, vt) is loaded once and feeds up to 2S+1 of the thread's
// outputs.  Loads per output drop from (2S+1)^4 to (2S+1)^3 * (PY+2S)/PY, which moves both
// kernels from L1-bandwidth bound to FP32-issue bound -- the floor of the bit-exact formulation
// (625 taps x 8 FMUL/FADD per cell for S = 2, no FMA because the reference rounds twice).
// The normalising weight of a cell depends only on how its window is clipped by the grid border;
// the (2S+1)^4 possible sums are accumulated on the host in the reference's order and looked up.

// weight sums per clip state (tiled kernels support S <= 2): [(st*NS + sz)*NS + sy]*NS + sx
__constant__ float c_gauss_wsum[625];

#define FLOF_TPY 4  // outputs per thread along y

// Blackwell packed fp32x2 arithmetic (PTX add.rn.f32x2 / mul.rn.f32x2, sm_100+): two independent
// IEEE round-to-nearest operations per instruction -- bit-identical to scalar fp32 multiplies and
// adds at half the issue slots.  Written as inline PTX with explicit .rn: the __fmul2_rn/__fadd2_rn
// intrinsics of CUDA 12.9 get contracted to FFMA2 by the compiler even under -fmad=false (checked in
// SASS), which would break parity with the reference's separately rounded multiply and add.
struct p4 { unsigned long long lo, hi; };  // a float4 as two f32x2 register pairs
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b)
{
	unsigned long long r;
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}
// w * (x, y) with two scalar multiplies, packed for the f32x2 add.  (A packed multiply cannot be
// used: ptxas 12.9 contracts mul.rn.f32x2 -- and even fma.rn.f32x2(a, b, -0.0) -- followed by
// add.rn.f32x2 into a single FFMA2 despite the explicit .rn and -fmad=false, verified in SASS and by
// the bit-exactness tests; it never contracts scalar mul.rn.f32 into a packed add.)
__device__ __forceinline__ unsigned long long f2_wmul(float w, unsigned long long q)
{
	float x, y;
	asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(q));
	const float px = __fmul_rn(w, x), py = __fmul_rn(w, y);
	unsigned long long r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(px), "f"(py));
	return r;
}
__device__ __forceinline__ unsigned long long f2_splat(float w)
{
	unsigned long long r;
	asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(w));
	return r;
}
__device__ __forceinline__ p4 p4_zero()
{
	p4 r;
	r.lo = 0ull;
	r.hi = 0ull;
	return r;
}
__device__ __forceinline__ p4 p4_load(const float4 *ptr)
{
	const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(ptr));
	p4 r;
	r.lo = v.x;
	r.hi = v.y;
	return r;
}
__device__ __forceinline__ float4 p4_unpack(const p4 &v)
{
	float4 r;
	asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v.lo));
	asm("mov.b64 {%0, %1}, %2;" : "=f"(r.z), "=f"(r.w) : "l"(v.hi));
	return r;
}
__device__ __forceinline__ void acc4(p4 &v, const p4 &q)
{
	v.lo = f2_add(v.lo, q.lo);
	v.hi = f2_add(v.hi, q.hi);
}
__device__ __forceinline__ void wacc4(p4 &v, float w, const p4 &q)
{
	v.lo = f2_add(v.lo, f2_wmul(w, q.lo));
	v.hi = f2_add(v.hi, f2_wmul(w, q.hi));
}
__device__ __forceinline__ void wacc4(float4 &v, float w, const float4 &q)  // scalar form (edge columns)
{
	v.x += w * q.x; v.y += w * q.y; v.z += w * q.z; v.w += w * q.w;
}

// thread -> (x, y0): consecutive threads = consecutive x; one y-patch of FLOF_TPY rows per thread
__device__ __forceinline__ bool tiled_xy(const flof_dim4 &d, int &x, int &y0)
{
	const int pty = (d.ny + FLOF_TPY - 1) / FLOF_TPY;
	const unsigned p = blockIdx.x * FLOF_BLOCK + threadIdx.x;
	if (p >= (unsigned)(d.nx * pty)) return false;
	const int py = (int)(p / (unsigned)d.nx);
	x = (int)(p - (unsigned)py * (unsigned)d.nx);
	y0 = py * FLOF_TPY;
	return true;
}
static inline flof_kd tiled_grid(const flof_ctx *ctx, flof_dim4 d, dim3 *g)
{
	const flof_kd kd = flof_kdim(ctx, d, g);  // z-dim = this rank's t-slices
	const int pty = (d.ny + FLOF_TPY - 1) / FLOF_TPY;
	g->x = (unsigned)(((int64_t)d.nx * pty + FLOF_BLOCK - 1) / FLOF_BLOCK);
	return kd;
}

// ------------------------------------------------------------------ 81-tap extrapolation ---
#ifndef FLOF_ETPY
#define FLOF_ETPY 4  // outputs per thread along y
#define FLOF_EBLK 2  // resident CTAs per SM the register budget is tuned for
#endif
#ifndef FLOF_EZCH
#define FLOF_EZCH 8   // z-planes walked by one CTA
#endif
template <int TPY> __device__ __forceinline__ bool tiled_xy_n(const flof_dim4 &d, int &x, int &y0)
{
	const int pty = (d.ny + TPY - 1) / TPY;
	const unsigned p = blockIdx.x * FLOF_BLOCK + threadIdx.x;
	if (p >= (unsigned)(d.nx * pty)) return false;
	const int py = (int)(p / (unsigned)d.nx);
	x = (int)(p - (unsigned)py * (unsigned)d.nx);
	y0 = py * TPY;
	return true;
}
// Per (vt, zk) pair the thread loads its FLOF_ETPY+2 rows x 3 columns (18 independent LDG.128 in
// flight) and then issues the 9 adds of each of its 4 outputs; row addresses are 32-bit offsets
// computed once per thread.  Lanes on the x border and patches whose cells are all marked skip
// the arithmetic and only copy.
__global__ void __launch_bounds__(FLOF_BLOCK, FLOF_EBLK)
    k_cv_expol_blur4d_tiled(const float4 *__restrict__ a, float4 *__restrict__ tmp, const float *__restrict__ mark,
                            flof_kd d)
{
	int x, y0;
	if (!tiled_xy_n<FLOF_ETPY>(d, x, y0)) return;
	const int t = (int)blockIdx.z + d.t0;
	// z-marching: a CTA walks FLOF_EZCH consecutive z-planes, so the rows of plane z+1 that it pulls from
	// L2 for output plane z are still in L1 when it computes planes z+1 and z+2 (ncu: the one-plane-per-CTA
	// version was L2->L1 bandwidth bound, every row being fetched by 9 different CTAs)
	for (int k = (int)blockIdx.y * FLOF_EZCH, kend = min(k + FLOF_EZCH, d.nz); k < kend; ++k) {
	const bool col_in = k >= 1 && k < d.nz - 1 && t >= 1 && t < d.nt - 1 && x >= 1 && x < d.nx - 1;
	const int64_t plane = flof_idx(d, 0, 0, k, t);
	bool need[FLOF_ETPY];
	bool any = false;
#pragma unroll
	for (int oy = 0; oy < FLOF_ETPY; ++oy) {
		const int y = y0 + oy;
		bool n = col_in && y >= 1 && y < d.ny - 1;
		if (n) n = __ldg(mark + plane + (int64_t)y * d.nx + x) == 0.f;
		need[oy] = n;
		any |= n;
	}
	p4 acc[FLOF_ETPY];
#pragma unroll
	for (int oy = 0; oy < FLOF_ETPY; ++oy) acc[oy] = p4_zero();
	if (any) {  // col_in holds: x-1 and x+1 are inside the grid
		int roff[FLOF_ETPY + 2];  // offset of (x, clamp(y0-1+r)) inside a z-t plane
#pragma unroll
		for (int r = 0; r < FLOF_ETPY + 2; ++r) roff[r] = min(max(y0 - 1 + r, 0), d.ny - 1) * d.nx + x;
		const int64_t sZ = (int64_t)d.nx * d.ny, sT = sZ * d.nz;
		for (int vt = t - 1; vt <= t + 1; ++vt) {
			const float4 *base = a + (sT * vt + sZ * (k - 1));
#pragma unroll 1
			for (int zk = 0; zk < 3; ++zk, base += sZ) {
				p4 L[FLOF_ETPY + 2][3];
#pragma unroll
				for (int r = 0; r < FLOF_ETPY + 2; ++r) {
					const float4 *row = base + roff[r];
					L[r][0] = p4_load(row - 1);
					L[r][1] = p4_load(row);
					L[r][2] = p4_load(row + 1);
				}
#pragma unroll
				for (int r = 0; r < FLOF_ETPY + 2; ++r)
#pragma unroll
					for (int oy = 0; oy < FLOF_ETPY; ++oy) {
						if (r < oy || r > oy + 2) continue;
						acc4(acc[oy], L[r][0]);
						acc4(acc[oy], L[r][1]);
						acc4(acc[oy], L[r][2]);
					}
			}
		}
	}
	const double f = 1. / 81.0;
#pragma unroll
	for (int oy = 0; oy < FLOF_ETPY; ++oy) {
		const int y = y0 + oy;
		if (y >= d.ny) continue;
		const int64_t c = plane + (int64_t)y * d.nx + x;
		if (need[oy]) {
			const float4 v = p4_unpack(acc[oy]);
			tmp[c] = make_float4((float)(v.x * f), (float)(v.y * f), (float)(v.z * f), (float)(v.w * f));
		} else {
			tmp[c] = __ldg(a + c);
		}
	}
	}  // z-march
}

int flof_launch_expol_tiled(flof_ctx *ctx, const float *a, float *tmp, const float *marker, flof_dim4 d)
{
	dim3 g;
	const flof_kd kd = tiled_grid(ctx, d, &g);
	g.x = (unsigned)(((int64_t)d.nx * ((d.ny + FLOF_ETPY - 1) / FLOF_ETPY) + FLOF_BLOCK - 1) / FLOF_BLOCK);
	g.y = (unsigned)((d.nz + FLOF_EZCH - 1) / FLOF_EZCH);
	FLOF_LAUNCH(k_cv_expol_blur4d_tiled, g, FLOF_BLOCK, 0, (const float4 *)a, (float4 *)tmp, marker, kd);
	return FLOF_OK;
}

// ------------------------------------------------------------------ 81-tap extrapolation, work list ---
// The marker grid is fixed for all sweeps of one corrVelsOf4d pass (ref :770-780) and only cells with
// marker == 0 ever change: on the synthetic pair ~23 % of the cells, scattered so that 94 % of the
// 32 x 4 warp patches of the dense kernel contain at least one (tools/expol_probe.py).  So the sweeps
// run over a compacted list of "items" -- one x, FLOF_ETPY consecutive y, one (z, t) -- built once per
// pass: every lane of a warp then has arithmetic to do, and the per-sweep `tmp.copyFrom(dst)` of the
// reference disappears (both ping-pong buffers already hold the cells that never change).
// item = linear id ((tl*nz + k)*nyb + yb)*nx + x in the low 28 bits, need-mask of the rows in the top 4.
#define FLOF_EITEM_BITS 28

__global__ void __launch_bounds__(FLOF_BLOCK)
    k_expol_build_items(const float *__restrict__ mark, uint32_t *__restrict__ items, unsigned int *__restrict__ count,
                        flof_kd d, int nyb)
{
	const unsigned p = blockIdx.x * FLOF_BLOCK + threadIdx.x;
	const int k = (int)blockIdx.y, tl = (int)blockIdx.z, t = tl + d.t0;
	unsigned mask = 0;
	uint32_t id = 0;
	if (p < (unsigned)(d.nx * nyb) && k >= 1 && k < d.nz - 1 && t >= 1 && t < d.nt - 1) {
		const int yb = (int)(p / (unsigned)d.nx), x = (int)(p - (unsigned)yb * (unsigned)d.nx);
		if (x >= 1 && x < d.nx - 1) {
			const int64_t plane = flof_idx(d, x, 0, k, t);
#pragma unroll
			for (int oy = 0; oy < FLOF_ETPY; ++oy) {
				const int y = yb * FLOF_ETPY + oy;
				if (y >= 1 && y < d.ny - 1 && __ldg(mark + plane + (int64_t)y * d.nx) == 0.f) mask |= 1u << oy;
			}
		}
		id = (uint32_t)(((tl * d.nz + k) * nyb + yb) * d.nx + x);
	}
	// block-aggregated append: one atomic per CTA, items of a CTA stay in x-fastest order
	__shared__ unsigned s_off[FLOF_BLOCK / 32], s_base;
	const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
	const unsigned b = __ballot_sync(0xffffffffu, mask != 0);
	if (lane == 0) s_off[wid] = __popc(b);
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned tot = 0;
		for (int w = 0; w < FLOF_BLOCK / 32; ++w) {
			const unsigned c = s_off[w];
			s_off[w] = tot;
			tot += c;
		}
		s_base = tot ? atomicAdd(count, tot) : 0u;
	}
	__syncthreads();
	if (mask) items[s_base + s_off[wid] + __popc(b & ((1u << lane) - 1u))] = id | (mask << FLOF_EITEM_BITS);
}

template <int MINB, int UNR>
__global__ void __launch_bounds__(FLOF_BLOCK, MINB)
    k_cv_expol_items(const float4 *__restrict__ a, float4 *__restrict__ out, const uint32_t *__restrict__ items, int n,
                     flof_kd d, int nyb)
{
	const int q = (int)(blockIdx.x * FLOF_BLOCK + threadIdx.x);
	if (q >= n) return;
	const uint32_t it = __ldg(items + q);
	const unsigned mask = it >> FLOF_EITEM_BITS;
	unsigned id = it & ((1u << FLOF_EITEM_BITS) - 1u);
	const int x = (int)(id % (unsigned)d.nx);
	id /= (unsigned)d.nx;
	const int y0 = (int)(id % (unsigned)nyb) * FLOF_ETPY;
	id /= (unsigned)nyb;
	const int k = (int)(id % (unsigned)d.nz), t = (int)(id / (unsigned)d.nz) + d.t0;

	p4 acc[FLOF_ETPY];
#pragma unroll
	for (int oy = 0; oy < FLOF_ETPY; ++oy) acc[oy] = p4_zero();
	int roff[FLOF_ETPY + 2];  // offset of (x, clamp(y0-1+r)) inside a z-t plane; clamped rows only feed unneeded outputs
#pragma unroll
	for (int r = 0; r < FLOF_ETPY + 2; ++r) roff[r] = min(max(y0 - 1 + r, 0), d.ny - 1) * d.nx + x;
	const int64_t sZ = (int64_t)d.nx * d.ny, sT = sZ * d.nz;
#pragma unroll(UNR >= 2 ? 3 : 1)
	for (int vt = t - 1; vt <= t + 1; ++vt) {
		const float4 *base = a + (sT * vt + sZ * (k - 1));
#pragma unroll(UNR >= 1 ? 3 : 1)
		for (int zk = 0; zk < 3; ++zk, base += sZ) {
			p4 L[FLOF_ETPY + 2][3];
#pragma unroll
			for (int r = 0; r < FLOF_ETPY + 2; ++r) {
				const float4 *row = base + roff[r];
				L[r][0] = p4_load(row - 1);
				L[r][1] = p4_load(row);
				L[r][2] = p4_load(row + 1);
			}
#pragma unroll
			for (int r = 0; r < FLOF_ETPY + 2; ++r)
#pragma unroll
				for (int oy = 0; oy < FLOF_ETPY; ++oy) {
					if (r < oy || r > oy + 2) continue;
					acc4(acc[oy], L[r][0]);
					acc4(acc[oy], L[r][1]);
					acc4(acc[oy], L[r][2]);
				}
		}
	}
	const double f = 1. / 81.0;
	float4 *o = out + (sT * t + sZ * k + (int64_t)y0 * d.nx + x);
#pragma unroll
	for (int oy = 0; oy < FLOF_ETPY; ++oy) {
		if (!((mask >> oy) & 1u)) continue;
		const float4 v = p4_unpack(acc[oy]);
		o[(int64_t)oy * d.nx] = make_float4((float)(v.x * f), (float)(v.y * f), (float)(v.z * f), (float)(v.w * f));
	}
}

// worst-case item count of this rank's slab; 0 = the linear id does not fit FLOF_EITEM_BITS (use the dense kernel)
static int64_t flof_expol_item_capacity(const flof_ctx *ctx, flof_dim4 d)
{
	int ta, tb;
	flof_slab(ctx, d.nt, &ta, &tb);
	const int64_t n = (int64_t)d.nx * ((d.ny + FLOF_ETPY - 1) / FLOF_ETPY) * d.nz * (tb - ta);
	return n < ((int64_t)1 << FLOF_EITEM_BITS) ? n : 0;
}
// builds the list into `items` (capacity entries), returns the number of items through *n (host, synchronises once)
static int flof_expol_build(flof_ctx *ctx, const float *marker, flof_dim4 d, uint32_t *items, unsigned int *count, int *n)
{
	dim3 g;
	const flof_kd kd = flof_kdim(ctx, d, &g);
	const int nyb = (d.ny + FLOF_ETPY - 1) / FLOF_ETPY;
	g.x = (unsigned)(((int64_t)d.nx * nyb + FLOF_BLOCK - 1) / FLOF_BLOCK);
	FLOF_CK(cudaMemsetAsync(count, 0, sizeof(unsigned int), ctx->stream));
	FLOF_LAUNCH(k_expol_build_items, g, FLOF_BLOCK, 0, marker, items, count, kd, nyb);
	unsigned int *h = (unsigned int *)ctx->pinned;
	FLOF_CK(cudaMemcpyAsync(h, count, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
	FLOF_CK(cudaStreamSynchronize(ctx->stream));
	*n = (int)h[0];
	return FLOF_OK;
}
static int flof_launch_expol_items(flof_ctx *ctx, const float *a, float *out, const uint32_t *items, int n, flof_dim4 d)
{
	if (n <= 0) return FLOF_OK;
	dim3 g;
	const flof_kd kd = flof_kdim(ctx, d, &g);
	const int nyb = (d.ny + FLOF_ETPY - 1) / FLOF_ETPY;
	const int variant = ctx->opt.expol_variant;
	const dim3 gi((unsigned)((n + FLOF_BLOCK - 1) / FLOF_BLOCK));
#define FLOF_EI_LAUNCH(MINB, UNR)                                                                                  \
	FLOF_LAUNCH((k_cv_expol_items<MINB, UNR>), gi, FLOF_BLOCK, 0, (const float4 *)a, (float4 *)out, items, n, kd, nyb)
	switch (variant) {
	case 1: FLOF_EI_LAUNCH(2, 1); break;
	case 2: FLOF_EI_LAUNCH(2, 2); break;
	case 3: FLOF_EI_LAUNCH(3, 0); break;
	case 4: FLOF_EI_LAUNCH(3, 1); break;
	case 5: FLOF_EI_LAUNCH(1, 2); break;
	case 6: FLOF_EI_LAUNCH(4, 0); break;
	default: FLOF_EI_LAUNCH(2, 0); break;
	}
	return FLOF_OK;
}

// ------------------------------------------------------------------ 81-tap extrapolation, work list, 4y x TZ z items ---
// Same work-list scheme with items of FLOF_ETPY rows x TZ z-planes (TZ = 2 or 4): the 18 row segments of a (vt, plane)
// pair feed the outputs of up to three z-planes, so an output costs 27 (TZ = 2) or 20.25 (TZ = 4) LDG.128 instead of
// 40.5 -- the kernel is bound by the L1 data pipe (~2 cycles per 128-byte line), not by its 81 packed adds per output.
// Per output the taps still arrive in the reference's order (vt, zk, yj, xi): planes ascend, rows ascend within a
// plane, and every accumulator only ever sees the planes of its own window.
// item = { linear id ((tl*nzb + kb)*nyb + yb)*nx + x , need-mask: bit (zo*FLOF_ETPY + oy) }
// sel: 0 = every item of the slab; 1 = only the items of its first and last slice (their outputs are what the
// neighbouring ranks need as ghost slices), 2 = only the others -- the two lists of the overlapped sweep (flof_blur.cu)
template <int TZ>
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_expol_build_items_zn(const float *__restrict__ mark, uint2 *__restrict__ items, unsigned int *__restrict__ count,
                           flof_kd d, int nyb, int sel)
{
	const unsigned p = blockIdx.x * FLOF_BLOCK + threadIdx.x;
	const int kb = (int)blockIdx.y, tl = (int)blockIdx.z, t = tl + d.t0;
	const int nzb = (d.nz + TZ - 1) / TZ;
	unsigned mask = 0;
	uint32_t id = 0;
	const bool edge = tl == 0 || tl == (int)gridDim.z - 1;
	if (p < (unsigned)(d.nx * nyb) && t >= 1 && t < d.nt - 1 && (sel == 0 || (sel == 1) == edge)) {
		const int yb = (int)(p / (unsigned)d.nx), x = (int)(p - (unsigned)yb * (unsigned)d.nx);
		if (x >= 1 && x < d.nx - 1) {
#pragma unroll
			for (int zo = 0; zo < TZ; ++zo) {
				const int k = kb * TZ + zo;
				if (k < 1 || k >= d.nz - 1) continue;
				const int64_t plane = flof_idx(d, x, 0, k, t);
#pragma unroll
				for (int oy = 0; oy < FLOF_ETPY; ++oy) {
					const int y = yb * FLOF_ETPY + oy;
					if (y >= 1 && y < d.ny - 1 && __ldg(mark + plane + (int64_t)y * d.nx) == 0.f) mask |= 1u << (zo * FLOF_ETPY + oy);
				}
			}
		}
		id = (uint32_t)(((tl * nzb + kb) * nyb + yb) * d.nx + x);
	}
	__shared__ unsigned s_off[FLOF_BLOCK / 32], s_base;
	const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
	const unsigned b = __ballot_sync(0xffffffffu, mask != 0);
	if (lane == 0) s_off[wid] = __popc(b);
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned tot = 0;
		for (int w = 0; w < FLOF_BLOCK / 32; ++w) {
			const unsigned c = s_off[w];
			s_off[w] = tot;
			tot += c;
		}
		s_base = tot ? atomicAdd(count, tot) : 0u;
	}
	__syncthreads();
	if (mask) items[s_base + s_off[wid] + __popc(b & ((1u << lane) - 1u))] = make_uint2(id, mask);
}

template <int TZ, int MINB>
__global__ void __launch_bounds__(FLOF_BLOCK, MINB)
    k_cv_expol_items_zn(const float4 *__restrict__ a, float4 *__restrict__ out, const uint2 *__restrict__ items, int n,
                        flof_kd d, int nyb)
{
	const int q = (int)(blockIdx.x * FLOF_BLOCK + threadIdx.x);
	if (q >= n) return;
	const uint2 it = __ldg(items + q);
	const unsigned mask = it.y;
	unsigned id = it.x;
	const int nzb = (d.nz + TZ - 1) / TZ;
	const int x = (int)(id % (unsigned)d.nx);
	id /= (unsigned)d.nx;
	const int y0 = (int)(id % (unsigned)nyb) * FLOF_ETPY;
	id /= (unsigned)nyb;
	const int k0 = (int)(id % (unsigned)nzb) * TZ, t = (int)(id / (unsigned)nzb) + d.t0;

	p4 acc[TZ][FLOF_ETPY];
#pragma unroll
	for (int zo = 0; zo < TZ; ++zo)
#pragma unroll
		for (int oy = 0; oy < FLOF_ETPY; ++oy) acc[zo][oy] = p4_zero();
	int roff[FLOF_ETPY + 2];  // offset of (x, clamp(y0-1+r)) inside a z-t plane; clamped rows only feed unneeded outputs
#pragma unroll
	for (int r = 0; r < FLOF_ETPY + 2; ++r) roff[r] = min(max(y0 - 1 + r, 0), d.ny - 1) * d.nx + x;
	const int64_t sZ = (int64_t)d.nx * d.ny, sT = sZ * d.nz;
#pragma unroll 1
	for (int vt = t - 1; vt <= t + 1; ++vt) {
		// planes k0 - 1 + pz, pz = 0 .. TZ+1 (clamped: out-of-range planes only feed outputs on the z border, which are
		// never needed).  The plane loop is unrolled so that every accumulator is a register, not an indexed array.
#pragma unroll
		for (int pz = 0; pz < TZ + 2; ++pz) {
			const float4 *base = a + (sT * vt + sZ * min(max(k0 - 1 + pz, 0), d.nz - 1));
			p4 L[FLOF_ETPY + 2][3];
#pragma unroll
			for (int r = 0; r < FLOF_ETPY + 2; ++r) {
				const float4 *row = base + roff[r];
				L[r][0] = p4_load(row - 1);
				L[r][1] = p4_load(row);
				L[r][2] = p4_load(row + 1);
			}
#pragma unroll
			for (int zo = 0; zo < TZ; ++zo) {
				if (pz < zo || pz > zo + 2) continue;
#pragma unroll
				for (int r = 0; r < FLOF_ETPY + 2; ++r)
#pragma unroll
					for (int oy = 0; oy < FLOF_ETPY; ++oy) {
						if (r < oy || r > oy + 2) continue;
						acc4(acc[zo][oy], L[r][0]);
						acc4(acc[zo][oy], L[r][1]);
						acc4(acc[zo][oy], L[r][2]);
					}
			}
		}
	}
	const double f = 1. / 81.0;
#pragma unroll
	for (int zo = 0; zo < TZ; ++zo) {
		if (!((mask >> (zo * FLOF_ETPY)) & ((1u << FLOF_ETPY) - 1u))) continue;
		float4 *o = out + (sT * t + sZ * (k0 + zo) + (int64_t)y0 * d.nx + x);
#pragma unroll
		for (int oy = 0; oy < FLOF_ETPY; ++oy) {
			if (!((mask >> (zo * FLOF_ETPY + oy)) & 1u)) continue;
			const float4 v = p4_unpack(acc[zo][oy]);
			o[(int64_t)oy * d.nx] = make_float4((float)(v.x * f), (float)(v.y * f), (float)(v.z * f), (float)(v.w * f));
		}
	}
}

// Shuffle variant of the same items: a lane loads only its OWN column of a row (one aligned, fully coalesced request per
// warp where the items are x-neighbours) and takes the x-1 / x+1 columns from the neighbouring lanes -- the list keeps the
// items of a row x-ordered and the cells that change form compact regions, so almost every lane has both neighbours next
// to it; a lane without one loads that column itself.  The L1 data pipe charges ~2 cycles per 128-byte line a request
// touches: three overlapping requests per row cost 4 + 5 + 5 lines, one aligned request 4 (plus the few lines of the
// predicated edge loads).  Arithmetic and tap order are unchanged.
__device__ __forceinline__ p4 p4_shfl_up(const p4 &v)
{
	p4 r;
	r.lo = __shfl_up_sync(0xffffffffu, v.lo, 1);
	r.hi = __shfl_up_sync(0xffffffffu, v.hi, 1);
	return r;
}
__device__ __forceinline__ p4 p4_shfl_down(const p4 &v)
{
	p4 r;
	r.lo = __shfl_down_sync(0xffffffffu, v.lo, 1);
	r.hi = __shfl_down_sync(0xffffffffu, v.hi, 1);
	return r;
}
template <int TZ, int MINB>
__global__ void __launch_bounds__(FLOF_BLOCK, MINB)
    k_cv_expol_items_zs(const float4 *__restrict__ a, float4 *__restrict__ out, const uint2 *__restrict__ items, int n,
                        flof_kd d, int nyb)
{
	const int q0 = (int)(blockIdx.x * FLOF_BLOCK + (threadIdx.x & ~31u));
	if (q0 >= n) return;  // warp-uniform
	const int lane = (int)(threadIdx.x & 31u), q = q0 + lane;
	const bool valid = q < n;
	const uint2 it = __ldg(items + (valid ? q : n - 1));
	const unsigned mask = valid ? it.y : 0u;
	unsigned id = it.x;
	const unsigned id_l = __shfl_up_sync(0xffffffffu, it.x, 1), id_r = __shfl_down_sync(0xffffffffu, it.x, 1);
	const int nzb = (d.nz + TZ - 1) / TZ;
	const int x = (int)(id % (unsigned)d.nx);
	id /= (unsigned)d.nx;
	const int y0 = (int)(id % (unsigned)nyb) * FLOF_ETPY;
	id /= (unsigned)nyb;
	const int k0 = (int)(id % (unsigned)nzb) * TZ, t = (int)(id / (unsigned)nzb) + d.t0;
	// the neighbouring lane holds the x-neighbour of the same rows exactly when the ids are consecutive (items have
	// 1 <= x <= nx-2, so id - 1 / id + 1 stay inside the row)
	const bool adj_l = lane > 0 && id_l + 1u == it.x;
	const bool adj_r = lane < 31 && id_r == it.x + 1u && q + 1 < n;

	p4 acc[TZ][FLOF_ETPY];
#pragma unroll
	for (int zo = 0; zo < TZ; ++zo)
#pragma unroll
		for (int oy = 0; oy < FLOF_ETPY; ++oy) acc[zo][oy] = p4_zero();
	int roff[FLOF_ETPY + 2];
#pragma unroll
	for (int r = 0; r < FLOF_ETPY + 2; ++r) roff[r] = min(max(y0 - 1 + r, 0), d.ny - 1) * d.nx + x;
	const int64_t sZ = (int64_t)d.nx * d.ny, sT = sZ * d.nz;
#pragma unroll 1
	for (int vt = t - 1; vt <= t + 1; ++vt) {
#pragma unroll
		for (int pz = 0; pz < TZ + 2; ++pz) {
			const float4 *base = a + (sT * vt + sZ * min(max(k0 - 1 + pz, 0), d.nz - 1));
			// row by row: the centre columns of all rows are requested first, each row's edge columns right before its adds
			p4 C[FLOF_ETPY + 2];
#pragma unroll
			for (int r = 0; r < FLOF_ETPY + 2; ++r) C[r] = p4_load(base + roff[r]);
#pragma unroll
			for (int r = 0; r < FLOF_ETPY + 2; ++r) {
				p4 L0 = p4_shfl_up(C[r]), L2 = p4_shfl_down(C[r]);
				if (!adj_l) L0 = p4_load(base + roff[r] - 1);
				if (!adj_r) L2 = p4_load(base + roff[r] + 1);
#pragma unroll
				for (int zo = 0; zo < TZ; ++zo) {
					if (pz < zo || pz > zo + 2) continue;
#pragma unroll
					for (int oy = 0; oy < FLOF_ETPY; ++oy) {
						if (r < oy || r > oy + 2) continue;
						acc4(acc[zo][oy], L0);
						acc4(acc[zo][oy], C[r]);
						acc4(acc[zo][oy], L2);
					}
				}
			}
		}
	}
	const double f = 1. / 81.0;
#pragma unroll
	for (int zo = 0; zo < TZ; ++zo) {
		if (!((mask >> (zo * FLOF_ETPY)) & ((1u << FLOF_ETPY) - 1u))) continue;
		float4 *o = out + (sT * t + sZ * (k0 + zo) + (int64_t)y0 * d.nx + x);
#pragma unroll
		for (int oy = 0; oy < FLOF_ETPY; ++oy) {
			if (!((mask >> (zo * FLOF_ETPY + oy)) & 1u)) continue;
			const float4 v = p4_unpack(acc[zo][oy]);
			o[(int64_t)oy * d.nx] = make_float4((float)(v.x * f), (float)(v.y * f), (float)(v.z * f), (float)(v.w * f));
		}
	}
}

static int64_t flof_expol_zn_capacity(const flof_ctx *ctx, flof_dim4 d, int tz)
{
	int ta, tb;
	flof_slab(ctx, d.nt, &ta, &tb);
	const int64_t n = (int64_t)d.nx * ((d.ny + FLOF_ETPY - 1) / FLOF_ETPY) * ((d.nz + tz - 1) / tz) * (tb - ta);
	return n < ((int64_t)1 << 31) ? n : 0;
}
static int flof_expol_zn_build(flof_ctx *ctx, const float *marker, flof_dim4 d, int tz, uint2 *items, unsigned int *count, int *n,
                               int sel = 0)
{
	dim3 g;
	const flof_kd kd = flof_kdim(ctx, d, &g);
	const int nyb = (d.ny + FLOF_ETPY - 1) / FLOF_ETPY;
	g.x = (unsigned)(((int64_t)d.nx * nyb + FLOF_BLOCK - 1) / FLOF_BLOCK);
	g.y = (unsigned)((d.nz + tz - 1) / tz);
	FLOF_CK(cudaMemsetAsync(count, 0, sizeof(unsigned int), ctx->stream));
	if (tz == 4)
		FLOF_LAUNCH(k_expol_build_items_zn<4>, g, FLOF_BLOCK, 0, marker, items, count, kd, nyb, sel);
	else
		FLOF_LAUNCH(k_expol_build_items_zn<2>, g, FLOF_BLOCK, 0, marker, items, count, kd, nyb, sel);
	unsigned int *h = (unsigned int *)ctx->pinned;
	FLOF_CK(cudaMemcpyAsync(h, count, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
	FLOF_CK(cudaStreamSynchronize(ctx->stream));
	*n = (int)h[0];
	return FLOF_OK;
}
static int flof_launch_expol_zn(flof_ctx *ctx, const float *a, float *out, const uint2 *items, int n, flof_dim4 d, int tz, int shfl)
{
	if (n <= 0) return FLOF_OK;
	dim3 g;
	const flof_kd kd = flof_kdim(ctx, d, &g);
	const int nyb = (d.ny + FLOF_ETPY - 1) / FLOF_ETPY;
	const dim3 gi((unsigned)((n + FLOF_BLOCK - 1) / FLOF_BLOCK));
#define FLOF_EZ_LAUNCH(TZ, MINB)                                                                                      \
	FLOF_LAUNCH((k_cv_expol_items_zn<TZ, MINB>), gi, FLOF_BLOCK, 0, (const float4 *)a, (float4 *)out, items, n, kd, nyb)
#define FLOF_ES_LAUNCH(TZ, MINB)                                                                                      \
	FLOF_LAUNCH((k_cv_expol_items_zs<TZ, MINB>), gi, FLOF_BLOCK, 0, (const float4 *)a, (float4 *)out, items, n, kd, nyb)
	if (shfl) {
		if (tz == 4) {
			if (ctx->opt.expol_variant == 1)
				FLOF_ES_LAUNCH(4, 1);
			else
				FLOF_ES_LAUNCH(4, 2);
		} else {
			if (ctx->opt.expol_variant == 1)
				FLOF_ES_LAUNCH(2, 1);
			else
				FLOF_ES_LAUNCH(2, 2);
		}
		return FLOF_OK;
	}
	if (tz == 4) {
		if (ctx->opt.expol_variant == 1)
			FLOF_EZ_LAUNCH(4, 1);
		else
			FLOF_EZ_LAUNCH(4, 2);
	} else {
		if (ctx->opt.expol_variant == 1)
			FLOF_EZ_LAUNCH(2, 1);
		else
			FLOF_EZ_LAUNCH(2, 2);
	}
	return FLOF_OK;
}

// ------------------------------------------------------------------ 81-tap extrapolation, component planes ---
// The work-list kernel above is bound by the L1 data pipe: 162 LDG.128 per item feed 648 packed adds, and every
// row costs three requests (x-1, x, x+1) that re-read the same lines (ncu: l1tex data-pipe 87 %, fma pipe 27 %).
// For the sweeps the field is therefore re-laid out as four component planes per t-slice,
//     S[(((t*4 + c)*nz + z)*ny + y)*nx + x]       (a t-slice keeps the byte size of the Vec4 layout)
// and an item becomes 4 consecutive x (one aligned LDG.128 of ONE component) times FLOF_ETPY rows: the x-reuse of
// the 3-wide window now happens in registers, the two edge values come from the neighbouring lanes (items of a
// warp are consecutive in x) with a predicated scalar load where the neighbour lane holds something else.
// One request per row instead of three; each output still adds its 81 taps in the reference's order.
// A CTA = 64 items x 4 components (warp w: component w & 3, items (w >> 2)*32 + lane).
struct flof_soa_dims { int nx, ny, nz, nt, t0, nxg, nyb; };

__global__ void __launch_bounds__(FLOF_BLOCK)
    k_expol_to_planes(const float4 *__restrict__ a, float *__restrict__ s, flof_kd d)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t)) return;
	const float4 v = __ldg(a + flof_idx(d, i, j, k, t));
	const int64_t n3 = (int64_t)d.nx * d.ny * d.nz;
	float *o = s + ((int64_t)t * 4) * n3 + ((int64_t)k * d.ny + j) * d.nx + i;
	o[0] = v.x; o[n3] = v.y; o[2 * n3] = v.z; o[3 * n3] = v.w;
}
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_expol_from_planes(float4 *__restrict__ a, const float *__restrict__ s, flof_kd d)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t)) return;
	const int64_t n3 = (int64_t)d.nx * d.ny * d.nz;
	const float *o = s + ((int64_t)t * 4) * n3 + ((int64_t)k * d.ny + j) * d.nx + i;
	a[flof_idx(d, i, j, k, t)] = make_float4(__ldg(o), __ldg(o + n3), __ldg(o + 2 * n3), __ldg(o + 3 * n3));
}

// item = { id = ((tl*nz + k)*nyb + yb)*nxg + xg,  mask bit (oy*4 + i) = cell (4*xg + i, FLOF_ETPY*yb + oy) is recomputed }
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_expol_build_items4(const float *__restrict__ mark, uint2 *__restrict__ items, unsigned int *__restrict__ count,
                         flof_soa_dims d)
{
	const unsigned p = blockIdx.x * FLOF_BLOCK + threadIdx.x;
	const int k = (int)blockIdx.y, tl = (int)blockIdx.z, t = tl + d.t0;
	unsigned mask = 0;
	uint32_t id = 0;
	if (p < (unsigned)(d.nxg * d.nyb) && k >= 1 && k < d.nz - 1 && t >= 1 && t < d.nt - 1) {
		const int yb = (int)(p / (unsigned)d.nxg), xg = (int)(p - (unsigned)yb * (unsigned)d.nxg);
		const int x0 = 4 * xg;
		const float *mp = mark + (((int64_t)t * d.nz + k) * d.ny) * d.nx + x0;
#pragma unroll
		for (int oy = 0; oy < FLOF_ETPY; ++oy) {
			const int y = yb * FLOF_ETPY + oy;
			if (y < 1 || y >= d.ny - 1) continue;
			const float4 m = __ldg(reinterpret_cast<const float4 *>(mp + (int64_t)y * d.nx));
			const float mv[4] = { m.x, m.y, m.z, m.w };
#pragma unroll
			for (int i = 0; i < 4; ++i)
				if (x0 + i >= 1 && x0 + i < d.nx - 1 && mv[i] == 0.f) mask |= 1u << (oy * 4 + i);
		}
		id = (uint32_t)(((tl * d.nz + k) * d.nyb + yb) * d.nxg + xg);
	}
	__shared__ unsigned s_off[FLOF_BLOCK / 32], s_base;
	const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
	const unsigned b = __ballot_sync(0xffffffffu, mask != 0);
	if (lane == 0) s_off[wid] = __popc(b);
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned tot = 0;
		for (int w = 0; w < FLOF_BLOCK / 32; ++w) {
			const unsigned c = s_off[w];
			s_off[w] = tot;
			tot += c;
		}
		s_base = tot ? atomicAdd(count, tot) : 0u;
	}
	__syncthreads();
	if (mask) items[s_base + s_off[wid] + __popc(b & ((1u << lane) - 1u))] = make_uint2(id, mask);
}

__device__ __forceinline__ unsigned long long f2_pack(float x, float y)
{
	unsigned long long r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
	return r;
}

template <int MINB>
__global__ void __launch_bounds__(FLOF_BLOCK, MINB)
    k_cv_expol_planes(const float *__restrict__ a, float *__restrict__ out, const uint2 *__restrict__ items, int n,
                      flof_soa_dims d)
{
	static_assert(FLOF_ETPY == 4, "mask layout assumes 4 rows per item");
	const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
	const int comp = (int)(wid & 3u);
	const int q0 = (int)(blockIdx.x * 64u + (wid >> 2) * 32u);
	if (q0 >= n) return;  // warp-uniform
	const int q = q0 + (int)lane;
	const bool valid = q < n;
	const uint2 it = __ldg(items + (valid ? q : n - 1));
	// neighbour lanes hold the x-adjacent groups of the same rows?  (ids are consecutive exactly then)
	const uint32_t id_l = __shfl_up_sync(0xffffffffu, it.x, 1), id_r = __shfl_down_sync(0xffffffffu, it.x, 1);
	unsigned id = it.x;
	const int xg = (int)(id % (unsigned)d.nxg);
	id /= (unsigned)d.nxg;
	const int y0 = (int)(id % (unsigned)d.nyb) * FLOF_ETPY;
	id /= (unsigned)d.nyb;
	const int k = (int)(id % (unsigned)d.nz), t = (int)(id / (unsigned)d.nz) + d.t0;
	const int x0 = 4 * xg;
	const bool adj_l = lane > 0 && id_l + 1u == it.x && xg > 0;
	const bool adj_r = lane < 31 && id_r == it.x + 1u && xg < d.nxg - 1 && q + 1 < n;
	const bool ld_l = !adj_l && x0 > 0, ld_r = !adj_r && x0 + 4 < d.nx;  // edge value needs its own load

	unsigned long long acc[FLOF_ETPY][2];
#pragma unroll
	for (int oy = 0; oy < FLOF_ETPY; ++oy) acc[oy][0] = acc[oy][1] = 0ull;
	int roff[FLOF_ETPY + 2];
#pragma unroll
	for (int r = 0; r < FLOF_ETPY + 2; ++r) roff[r] = min(max(y0 - 1 + r, 0), d.ny - 1) * d.nx + x0;
	const int64_t sZ = (int64_t)d.nx * d.ny, n3 = sZ * d.nz, sT = 4 * n3;
#pragma unroll 1
	for (int vt = t - 1; vt <= t + 1; ++vt) {
		const float *base = a + (sT * vt + n3 * comp + sZ * (k - 1));
#pragma unroll 1
		for (int zk = 0; zk < 3; ++zk, base += sZ) {
			float4 m[FLOF_ETPY + 2];
			float el[FLOF_ETPY + 2], er[FLOF_ETPY + 2];
#pragma unroll
			for (int r = 0; r < FLOF_ETPY + 2; ++r) {
				const float *row = base + roff[r];
				m[r] = __ldg(reinterpret_cast<const float4 *>(row));
				el[r] = ld_l ? __ldg(row - 1) : 0.f;
				er[r] = ld_r ? __ldg(row + 4) : 0.f;
			}
#pragma unroll
			for (int r = 0; r < FLOF_ETPY + 2; ++r) {
				const float sl = __shfl_up_sync(0xffffffffu, m[r].w, 1), sr = __shfl_down_sync(0xffffffffu, m[r].x, 1);
				const float l = adj_l ? sl : el[r], rr = adj_r ? sr : er[r];
				const unsigned long long A = f2_pack(l, m[r].x), B = f2_pack(m[r].x, m[r].y), C = f2_pack(m[r].y, m[r].z),
				                         D = f2_pack(m[r].z, m[r].w), E = f2_pack(m[r].w, rr);
#pragma unroll
				for (int oy = 0; oy < FLOF_ETPY; ++oy) {
					if (r < oy || r > oy + 2) continue;
					// outputs x0, x0+1 take (x-1, x, x+1) = A, B, C; outputs x0+2, x0+3 take C, D, E -- xi ascending
					acc[oy][0] = f2_add(f2_add(f2_add(acc[oy][0], A), B), C);
					acc[oy][1] = f2_add(f2_add(f2_add(acc[oy][1], C), D), E);
				}
			}
		}
	}
	if (!valid) return;
	const double f = 1. / 81.0;
	const int64_t cell0 = sT * t + n3 * comp + sZ * k + (int64_t)y0 * d.nx + x0;
#pragma unroll
	for (int oy = 0; oy < FLOF_ETPY; ++oy) {
		const unsigned mrow = (it.y >> (oy * 4)) & 15u;
		if (!mrow) continue;
		const int64_t c = cell0 + (int64_t)oy * d.nx;
		float4 o = __ldg(reinterpret_cast<const float4 *>(a + c));  // cells that are not recomputed keep their value
		float v0, v1, v2, v3;
		asm("mov.b64 {%0, %1}, %2;" : "=f"(v0), "=f"(v1) : "l"(acc[oy][0]));
		asm("mov.b64 {%0, %1}, %2;" : "=f"(v2), "=f"(v3) : "l"(acc[oy][1]));
		if (mrow & 1u) o.x = (float)(v0 * f);
		if (mrow & 2u) o.y = (float)(v1 * f);
		if (mrow & 4u) o.z = (float)(v2 * f);
		if (mrow & 8u) o.w = (float)(v3 * f);
		*reinterpret_cast<float4 *>(out + c) = o;
	}
}

static flof_soa_dims flof_soa_dims_of(const flof_ctx *ctx, flof_dim4 d)
{
	int ta, tb;
	flof_slab(ctx, d.nt, &ta, &tb);
	flof_soa_dims s = { d.nx, d.ny, d.nz, d.nt, ta, d.nx / 4, (d.ny + FLOF_ETPY - 1) / FLOF_ETPY };
	return s;
}
// worst-case item count of the plane layout; 0 = not applicable (nx not a multiple of 4, or ids exceed 32 bits)
static int64_t flof_expol_planes_capacity(const flof_ctx *ctx, flof_dim4 d)
{
	if (d.nx % 4 != 0 || d.nx < 8) return 0;
	int ta, tb;
	flof_slab(ctx, d.nt, &ta, &tb);
	const int64_t n = (int64_t)(d.nx / 4) * ((d.ny + FLOF_ETPY - 1) / FLOF_ETPY) * d.nz * (tb - ta);
	return n < ((int64_t)1 << 31) ? n : 0;
}
static int flof_expol_planes_build(flof_ctx *ctx, const float *marker, flof_dim4 d, uint2 *items, unsigned int *count, int *n)
{
	const flof_soa_dims sd = flof_soa_dims_of(ctx, d);
	int ta, tb;
	flof_slab(ctx, d.nt, &ta, &tb);
	const dim3 g((unsigned)(((int64_t)sd.nxg * sd.nyb + FLOF_BLOCK - 1) / FLOF_BLOCK), (unsigned)d.nz, (unsigned)(tb - ta));
	FLOF_CK(cudaMemsetAsync(count, 0, sizeof(unsigned int), ctx->stream));
	FLOF_LAUNCH(k_expol_build_items4, g, FLOF_BLOCK, 0, marker, items, count, sd);
	unsigned int *h = (unsigned int *)ctx->pinned;
	FLOF_CK(cudaMemcpyAsync(h, count, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
	FLOF_CK(cudaStreamSynchronize(ctx->stream));
	*n = (int)h[0];
	return FLOF_OK;
}
static int flof_expol_to_planes(flof_ctx *ctx, const float *a, float *s, flof_dim4 d)
{
	dim3 g;
	const flof_kd kd = flof_kdim(ctx, d, &g);
	FLOF_LAUNCH(k_expol_to_planes, g, FLOF_BLOCK, 0, (const float4 *)a, s, kd);
	return FLOF_OK;
}
static int flof_expol_from_planes(flof_ctx *ctx, float *a, const float *s, flof_dim4 d)
{
	dim3 g;
	const flof_kd kd = flof_kdim(ctx, d, &g);
	FLOF_LAUNCH(k_expol_from_planes, g, FLOF_BLOCK, 0, (float4 *)a, s, kd);
	return FLOF_OK;
}
static int flof_launch_expol_planes(flof_ctx *ctx, const float *a, float *out, const uint2 *items, int n, flof_dim4 d)
{
	if (n <= 0) return FLOF_OK;
	const flof_soa_dims sd = flof_soa_dims_of(ctx, d);
	const int variant = ctx->opt.expol_variant;
	const dim3 g((unsigned)((n + 63) / 64));
	switch (variant) {
	case 1: FLOF_LAUNCH((k_cv_expol_planes<2>), g, FLOF_BLOCK, 0, a, out, items, n, sd); break;
	case 2: FLOF_LAUNCH((k_cv_expol_planes<4>), g, FLOF_BLOCK, 0, a, out, items, n, sd); break;
	default: FLOF_LAUNCH((k_cv_expol_planes<3>), g, FLOF_BLOCK, 0, a, out, items, n, sd); break;
	}
	return FLOF_OK;
}

// ------------------------------------------------------------------ Gaussian ---------------
template <int S> __device__ __forceinline__ int clip_state(int i, int n)
{  // 0 = window not clipped, 1..S = clipped at the low side by that many taps, S+1..2S = high side
	const int lo = S - i, hi = i + S - (n - 1);
	return lo > 0 ? lo : (hi > 0 ? S + hi : 0);
}

// Fast path: outputs whose x window is complete (S <= x < nx-S), so the 2S+1 column loads of a row
// are plain unit-stride loads at immediate offsets and need no masking; rows outside the grid are
// skipped with a (warp-uniform) branch, t/z slabs outside with the loop `continue`.  The (S+1)^2
// distinct weights of a (vt, zk) pair are fetched once into registers.  The 2(S-1) remaining
// interior columns next to the x border are done by k_gauss_blur4d_cols (generic arithmetic).
template <int S>
__global__ void __launch_bounds__(FLOF_BLOCK, 3)
    k_gauss_blur4d_tiled(const float4 *__restrict__ a, float4 *__restrict__ tmp, flof_kd d)
{
	constexpr int NS = 2 * S + 1;
	int x, y0;
	if (!tiled_xy(d, x, y0)) return;
	const int k = (int)blockIdx.y, t = (int)blockIdx.z + d.t0;
	if (k < 1 || k >= d.nz - 1 || t < 1 || t >= d.nt - 1) return;  // KERNEL(fourd, bnd = 1)
	if (x < S || x >= d.nx - S) return;                            // x-border columns: other kernel
	p4 acc[FLOF_TPY];
#pragma unroll
	for (int oy = 0; oy < FLOF_TPY; ++oy) acc[oy] = p4_zero();
	const int64_t sZ = (int64_t)d.nx * d.ny, sT = sZ * d.nz;
	const int xoff = (y0 - S) * d.nx + x;  // may be negative: only dereferenced for rows inside the grid

	// both plane loops stay rolled: unrolling vt made the S = 2 kernel 4600 instructions (74 KB), and ncu showed
	// "no instruction" (i-cache) as its top stall; the body of one (vt, zk) plane is ~900 instructions.  3.38 -> 2.70 ms
	// at 64^4.  (S = 1 is small enough and measured faster with vt unrolled: 0.53 vs 0.62 ms.)
#pragma unroll(S >= 2 ? 1 : 3)
	for (int vt = t - S; vt <= t + S; ++vt) {
		if (vt < 0 || vt >= d.nt) continue;
		const int dt2 = (vt - t) * (vt - t);
#pragma unroll 1
		for (int zk = k - S; zk <= k + S; ++zk) {
			if (zk < 0 || zk >= d.nz) continue;
			const int dz2 = dt2 + (zk - k) * (zk - k);
			float w[S + 1][S + 1];
#pragma unroll
			for (int p = 0; p <= S; ++p)
#pragma unroll
				for (int q = 0; q <= S; ++q) w[p][q] = c_gauss_w[dz2 + p * p + q * q];
			const float4 *base = a + (sT * vt + sZ * zk) + xoff;
#pragma unroll
			for (int r = 0; r < FLOF_TPY + 2 * S; ++r) {
				const int yj = y0 - S + r;
				if (yj < 0 || yj >= d.ny) continue;
				const float4 *row = base + r * d.nx;
				p4 L[NS];
#pragma unroll
				for (int q = 0; q < NS; ++q) L[q] = p4_load(row + (q - S));
#pragma unroll
				for (int oy = 0; oy < FLOF_TPY; ++oy) {
					const int dy = r - S - oy;
					if (dy < -S || dy > S) continue;
#pragma unroll
					for (int q = 0; q < NS; ++q) wacc4(acc[oy], w[dy < 0 ? -dy : dy][q < S ? S - q : q - S], L[q]);
				}
			}
		}
	}
	const int st = clip_state<S>(t, d.nt), sz = clip_state<S>(k, d.nz);
#pragma unroll
	for (int oy = 0; oy < FLOF_TPY; ++oy) {
		const int y = y0 + oy;
		if (y < 1 || y >= d.ny - 1) continue;
		const int64_t c = flof_idx(d, x, y, k, t);
		const float weight = c_gauss_wsum[((st * NS + sz) * NS + clip_state<S>(y, d.ny)) * NS];  // sx = 0
		const float4 v = p4_unpack(acc[oy]);
		if (weight > FLOF_VECTOR_EPSILON)
			tmp[c] = make_float4(v.x / weight, v.y / weight, v.z / weight, v.w / weight);
		else
			tmp[c] = __ldg(a + c);
	}
}

// interior columns whose x window is clipped: 1 <= x < S and nx-S <= x < nx-1 (none for S = 1).
// One thread per cell, the generic tap loop of flof_blur.cu.
template <int S>
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_gauss_blur4d_cols(const float4 *__restrict__ a, float4 *__restrict__ tmp, flof_kd d)
{
	constexpr int NC = 2 * (S - 1);
	const unsigned p = blockIdx.x * FLOF_BLOCK + threadIdx.x;
	if (NC == 0 || p >= (unsigned)(NC * d.ny)) return;
	const int j = (int)(p / (unsigned)NC), ci = (int)(p - (unsigned)j * NC);
	const int i = ci < S - 1 ? 1 + ci : d.nx - S + (ci - (S - 1));
	const int k = (int)blockIdx.y, t = (int)blockIdx.z + d.t0;
	if (!flof_in_bounds(d, i, j, k, t, 1)) return;
	float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
	float weight = 0.f;
	const int x0 = max(i - S, 0), x1 = min(i + S, d.nx - 1);
#pragma unroll 1
	for (int vt = t - S; vt <= t + S; ++vt) {
		if (vt < 0 || vt >= d.nt) continue;
		const int dt2 = (vt - t) * (vt - t);
#pragma unroll 1
		for (int zk = k - S; zk <= k + S; ++zk) {
			if (zk < 0 || zk >= d.nz) continue;
			const int dz2 = dt2 + (zk - k) * (zk - k);
#pragma unroll 1
			for (int yj = j - S; yj <= j + S; ++yj) {
				if (yj < 0 || yj >= d.ny) continue;
				const int dy2 = dz2 + (yj - j) * (yj - j);
				const float4 *row = a + flof_idx(d, 0, yj, zk, vt);
				for (int xi = x0; xi <= x1; ++xi) {
					const float wcurr = c_gauss_w[dy2 + (xi - i) * (xi - i)];
					weight += wcurr;
					wacc4(val, wcurr, __ldg(row + xi));
				}
			}
		}
	}
	const int64_t c = flof_idx(d, i, j, k, t);
	if (weight > FLOF_VECTOR_EPSILON)
		tmp[c] = make_float4(val.x / weight, val.y / weight, val.z / weight, val.w / weight);
	else
		tmp[c] = __ldg(a + c);
}

// host: weight sums for every clip state, accumulated in the reference's tap order (:138-151)
static void build_wsum(const float *w, int S, float *out)
{
	const int NS = 2 * S + 1;
	for (int st = 0; st < NS; ++st)
		for (int sz = 0; sz < NS; ++sz)
			for (int sy = 0; sy < NS; ++sy)
				for (int sx = 0; sx < NS; ++sx) {
					const int s4[4] = { sx, sy, sz, st };
					int lo[4], hi[4];
					for (int c = 0; c < 4; ++c) {
						lo[c] = -S;
						hi[c] = S;
						if (s4[c] >= 1 && s4[c] <= S) lo[c] = -S + s4[c];
						if (s4[c] > S) hi[c] = S - (s4[c] - S);
					}
					volatile float weight = 0.f;
					for (int vt = lo[3]; vt <= hi[3]; ++vt)
						for (int zk = lo[2]; zk <= hi[2]; ++zk)
							for (int yj = lo[1]; yj <= hi[1]; ++yj)
								for (int xi = lo[0]; xi <= hi[0]; ++xi) weight += w[vt * vt + zk * zk + yj * yj + xi * xi];
					out[((st * NS + sz) * NS + sy) * NS + sx] = weight;
				}
}

// returns 1 if the tiled kernel handled the pass, 0 if the caller must use the generic kernel
int flof_launch_gauss_tiled(flof_ctx *ctx, const float *a, float *tmp, flof_dim4 d, int s, const float *w,
                            bool upload_tables)
{
	if (s > 2) return 0;
	if (d.nx < 2 * s + 1 || d.ny < 2 * s + 1 || d.nz < 2 * s + 1 || d.nt < 2 * s + 1) return 0;
	if (upload_tables) {
		float wsum[625];
		build_wsum(w, s, wsum);
		const int ns = 2 * s + 1;
		if (cudaMemcpyToSymbolAsync(c_gauss_wsum, wsum, sizeof(float) * ns * ns * ns * ns, 0, cudaMemcpyHostToDevice,
		                            ctx->stream) != cudaSuccess)
			return flof_fail(ctx, FLOF_ERR_CUDA, "gaussianBlur: weight table upload failed"), -1;
	}
	dim3 tg;
	const flof_kd kd = tiled_grid(ctx, d, &tg);
	const int pi = flof_prof_pre(ctx, s == 1 ? "k_gauss_blur4d_tiled<1>" : "k_gauss_blur4d_tiled<2>");
	if (s == 1)
		k_gauss_blur4d_tiled<1><<<tg, FLOF_BLOCK, 0, ctx->stream>>>((const float4 *)a, (float4 *)tmp, kd);
	else
		k_gauss_blur4d_tiled<2><<<tg, FLOF_BLOCK, 0, ctx->stream>>>((const float4 *)a, (float4 *)tmp, kd);
	flof_prof_post(ctx, pi);
	ctx->launches++;
	if (s == 2) {
		const int pj = flof_prof_pre(ctx, "k_gauss_blur4d_cols<2>");
		dim3 g((unsigned)((2 * d.ny + FLOF_BLOCK - 1) / FLOF_BLOCK), (unsigned)d.nz, tg.z);
		k_gauss_blur4d_cols<2><<<g, FLOF_BLOCK, 0, ctx->stream>>>((const float4 *)a, (float4 *)tmp, kd);
		flof_prof_post(ctx, pj);
		ctx->launches++;
	}
	if (cudaGetLastError() != cudaSuccess) return flof_fail(ctx, FLOF_ERR_CUDA, "k_gauss_blur4d_tiled launch failed"), -1;
	return 1;
}
