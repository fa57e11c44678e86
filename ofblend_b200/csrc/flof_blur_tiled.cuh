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

__device__ __forceinline__ void acc4(float4 &v, const float4 &q)
{
	v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
}
__device__ __forceinline__ void wacc4(float4 &v, float w, const float4 &q)
{
	v.x += w * q.x; v.y += w * q.y; v.z += w * q.z; v.w += w * q.w;
}

// thread -> (x, y0): consecutive threads = consecutive x; one y-patch of FLOF_TPY rows per thread
__device__ __forceinline__ bool tiled_xy(flof_dim4 d, int &x, int &y0)
{
	const int pty = (d.ny + FLOF_TPY - 1) / FLOF_TPY;
	const unsigned p = blockIdx.x * FLOF_BLOCK + threadIdx.x;
	if (p >= (unsigned)(d.nx * pty)) return false;
	const int py = (int)(p / (unsigned)d.nx);
	x = (int)(p - (unsigned)py * (unsigned)d.nx);
	y0 = py * FLOF_TPY;
	return true;
}
static inline dim3 tiled_grid(flof_dim4 d)
{
	const int pty = (d.ny + FLOF_TPY - 1) / FLOF_TPY;
	return dim3((unsigned)(((int64_t)d.nx * pty + FLOF_BLOCK - 1) / FLOF_BLOCK), (unsigned)d.nz, (unsigned)d.nt);
}

// ------------------------------------------------------------------ 81-tap extrapolation ---
__global__ void __launch_bounds__(FLOF_BLOCK, 3)
    k_cv_expol_blur4d_tiled(const float4 *__restrict__ a, float4 *__restrict__ tmp, const float *__restrict__ mark,
                            flof_dim4 d)
{
	int x, y0;
	if (!tiled_xy(d, x, y0)) return;
	const int k = (int)blockIdx.y, t = (int)blockIdx.z;
	const bool col_in = k >= 1 && k < d.nz - 1 && t >= 1 && t < d.nt - 1 && x >= 1 && x < d.nx - 1;
	const int64_t plane = flof_idx(d, 0, 0, k, t);
	bool need[FLOF_TPY];
	bool any = false;
#pragma unroll
	for (int oy = 0; oy < FLOF_TPY; ++oy) {
		const int y = y0 + oy;
		bool n = col_in && y >= 1 && y < d.ny - 1;
		if (n) n = __ldg(mark + plane + (int64_t)y * d.nx + x) == 0.f;
		need[oy] = n;
		any |= n;
	}
	float4 acc[FLOF_TPY];
#pragma unroll
	for (int oy = 0; oy < FLOF_TPY; ++oy) acc[oy] = make_float4(0.f, 0.f, 0.f, 0.f);
	if (any) {  // (whole warps of marked / border cells skip the arithmetic)
		const int xm = max(x - 1, 0), xp = min(x + 1, d.nx - 1);
		for (int vt = t - 1; vt <= t + 1; ++vt)
			for (int zk = k - 1; zk <= k + 1; ++zk) {
				const float4 *base = a + flof_idx(d, 0, 0, zk, vt);
#pragma unroll
				for (int r = 0; r < FLOF_TPY + 2; ++r) {
					const int yj = min(max(y0 - 1 + r, 0), d.ny - 1);
					const float4 *row = base + (int64_t)yj * d.nx;
					const float4 l0 = __ldg(row + xm), l1 = __ldg(row + x), l2 = __ldg(row + xp);
#pragma unroll
					for (int oy = 0; oy < FLOF_TPY; ++oy) {
						if (r < oy || r > oy + 2) continue;
						acc4(acc[oy], l0);
						acc4(acc[oy], l1);
						acc4(acc[oy], l2);
					}
				}
			}
	}
	const double f = 1. / 81.0;
#pragma unroll
	for (int oy = 0; oy < FLOF_TPY; ++oy) {
		const int y = y0 + oy;
		if (y >= d.ny) continue;
		const int64_t c = plane + (int64_t)y * d.nx + x;
		if (need[oy]) {
			const float4 v = acc[oy];
			tmp[c] = make_float4((float)(v.x * f), (float)(v.y * f), (float)(v.z * f), (float)(v.w * f));
		} else {
			tmp[c] = __ldg(a + c);
		}
	}
}

int flof_launch_expol_tiled(flof_ctx *ctx, const float *a, float *tmp, const float *marker, flof_dim4 d)
{
	FLOF_LAUNCH(k_cv_expol_blur4d_tiled, tiled_grid(d), FLOF_BLOCK, 0, (const float4 *)a, (float4 *)tmp, marker, d);
	return FLOF_OK;
}

// ------------------------------------------------------------------ Gaussian ---------------
template <int S> __device__ __forceinline__ int clip_state(int i, int n)
{  // 0 = window not clipped, 1..S = clipped at the low side by that many taps, S+1..2S = high side
	const int lo = S - i, hi = i + S - (n - 1);
	return lo > 0 ? lo : (hi > 0 ? S + hi : 0);
}

template <int S>
__global__ void __launch_bounds__(FLOF_BLOCK, 3)
    k_gauss_blur4d_tiled(const float4 *__restrict__ a, float4 *__restrict__ tmp, flof_dim4 d)
{
	constexpr int NS = 2 * S + 1;
	int x, y0;
	if (!tiled_xy(d, x, y0)) return;
	const int k = (int)blockIdx.y, t = (int)blockIdx.z;
	if (k < 1 || k >= d.nz - 1 || t < 1 || t >= d.nt - 1) return;  // KERNEL(fourd, bnd = 1)
	if (x < 1 || x >= d.nx - 1) return;
	int xs[NS];
	bool xin[NS];
#pragma unroll
	for (int q = 0; q < NS; ++q) {
		const int xi = x - S + q;
		xin[q] = xi >= 0 && xi < d.nx;
		xs[q] = min(max(xi, 0), d.nx - 1);
	}
	const bool edge = !(xin[0] && xin[NS - 1]);
	float4 acc[FLOF_TPY];
#pragma unroll
	for (int oy = 0; oy < FLOF_TPY; ++oy) acc[oy] = make_float4(0.f, 0.f, 0.f, 0.f);

	for (int vt = t - S; vt <= t + S; ++vt) {
		if (vt < 0 || vt >= d.nt) continue;
		const int dt2 = (vt - t) * (vt - t);
		for (int zk = k - S; zk <= k + S; ++zk) {
			if (zk < 0 || zk >= d.nz) continue;
			const int dz2 = dt2 + (zk - k) * (zk - k);
			const float4 *base = a + flof_idx(d, 0, 0, zk, vt);
#pragma unroll
			for (int r = 0; r < FLOF_TPY + 2 * S; ++r) {
				const int yj = y0 - S + r;
				if (yj < 0 || yj >= d.ny) continue;
				const float4 *row = base + (int64_t)yj * d.nx;
				float4 L[NS];
#pragma unroll
				for (int q = 0; q < NS; ++q) L[q] = __ldg(row + xs[q]);
#pragma unroll
				for (int oy = 0; oy < FLOF_TPY; ++oy) {
					const int dy = r - S - oy;
					if (dy < -S || dy > S) continue;
					const int by = dz2 + dy * dy;
#pragma unroll
					for (int q = 0; q < NS; ++q) {
						if (edge && !xin[q]) continue;
						wacc4(acc[oy], c_gauss_w[by + (q - S) * (q - S)], L[q]);
					}
				}
			}
		}
	}
	const int st = clip_state<S>(t, d.nt), sz = clip_state<S>(k, d.nz), sx = clip_state<S>(x, d.nx);
#pragma unroll
	for (int oy = 0; oy < FLOF_TPY; ++oy) {
		const int y = y0 + oy;
		if (y < 1 || y >= d.ny - 1) continue;
		const int64_t c = flof_idx(d, x, y, k, t);
		const float weight = c_gauss_wsum[((st * NS + sz) * NS + clip_state<S>(y, d.ny)) * NS + sx];
		const float4 v = acc[oy];
		if (weight > FLOF_VECTOR_EPSILON)
			tmp[c] = make_float4(v.x / weight, v.y / weight, v.z / weight, v.w / weight);
		else
			tmp[c] = __ldg(a + c);
	}
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
	const int pi = flof_prof_pre(ctx, s == 1 ? "k_gauss_blur4d_tiled<1>" : "k_gauss_blur4d_tiled<2>");
	if (s == 1)
		k_gauss_blur4d_tiled<1><<<tiled_grid(d), FLOF_BLOCK, 0, ctx->stream>>>((const float4 *)a, (float4 *)tmp, d);
	else
		k_gauss_blur4d_tiled<2><<<tiled_grid(d), FLOF_BLOCK, 0, ctx->stream>>>((const float4 *)a, (float4 *)tmp, d);
	flof_prof_post(ctx, pi);
	ctx->launches++;
	if (cudaGetLastError() != cudaSuccess) return flof_fail(ctx, FLOF_ERR_CUDA, "k_gauss_blur4d_tiled launch failed"), -1;
	return 1;
}
