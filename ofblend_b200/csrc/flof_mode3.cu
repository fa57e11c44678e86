// flof_mode3.cu -- mode 3: sliced hi-res SDF load into the 4D grid, per-frame deformation
// look-up into a 3D output grid, and the 3D output post-processing.
// ref: knLoadPlaceGridIpol :1445-1452, knLoadPlaceGridRescale :1454-1459, loadPlaceGrid4d
//      :1464-1595, shiftForwGrid4d :1598-1616, knSemiLagrangeLookupSlice4d_WithVel :1648-1665,
//      loadAdvectTimeSlice_OptRun :1951-2105 (optflow4d.cpp); simpleBlurSpecial test.cpp:93-131.
//
// B200 design: the deformation volume (32^3 x 48 Vec4 = 25 MB) and the hi-res 4D SDF
// (100^3 x 150 = 600 MB) are simply resident in HBM, so the reference's per-frame file reads and
// its +-10 % "defovol" streaming window disappear; one output frame is one gather kernel that
// writes 4 B/cell and reads phi through L2 (8 B per output cell).
#include <math.h>

#include "flof_common.cuh"

// ------------------------------------------------------------------ sliced load ------------
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_load_place_ipol(float *__restrict__ phi, flof_dim4 d, int slice, const float *__restrict__ src,
                      flof_dim3 sd, float3 fac, float3 off, float wc)
{
	const unsigned p = blockIdx.x * FLOF_BLOCK + threadIdx.x;
	if (p >= (unsigned)(d.nx * d.ny)) return;
	const int j = p / d.nx, i = p - j * d.nx, k = blockIdx.y;
	if (i < 1 || j < 1 || k < 1 || i >= d.nx - 1 || j >= d.ny - 1 || k >= d.nz - 1) return;  // KERNEL(bnd=1)
	const float x = (float)i * fac.x + off.x, y = (float)j * fac.y + off.y, z = (float)k * fac.z + off.z;
	const int64_t c = flof_idx(d, i, j, k, slice);
	phi[c] += wc * flof_interpol3d<float>(src, sd, x, y, z);
}
extern "C" int flof_load_place_slice(flof_ctx *ctx, float *phi, flof_dim4 d, int slice, const float *src3,
                                     flof_dim3 sd, const float srcFac3[3], const float off3[3], float wc)
{
	FLOF_ARG(slice >= 0 && slice < d.nt, "loadPlaceGrid4d: slice %d out of range", slice);
	FLOF_ARG(sd.nx >= 2 && sd.ny >= 2, "loadPlaceGrid4d: Invalid src size");
	dim3 g((unsigned)(((int64_t)d.nx * d.ny + FLOF_BLOCK - 1) / FLOF_BLOCK), (unsigned)d.nz, 1);
	FLOF_LAUNCH(k_load_place_ipol, g, FLOF_BLOCK, 0, phi, d, slice, src3, sd,
	            make_float3(srcFac3[0], srcFac3[1], srcFac3[2]), make_float3(off3[0], off3[1], off3[2]), wc);
	return FLOF_OK;
}

__global__ void __launch_bounds__(FLOF_BLOCK) k_load_place_rescale(float *__restrict__ phi, flof_dim4 d, const float *__restrict__ w)
{
	int i, j, k, t;
	if (!flof_cell_ijkt(d, i, j, k, t) || !flof_in_bounds(d, i, j, k, t, 1)) return;
	const float wt = w[t];
	if (wt <= 0.f) return;
	phi[flof_idx(d, i, j, k, t)] *= wt;
}
extern "C" int flof_load_place_rescale(flof_ctx *ctx, float *phi, flof_dim4 d, const float *w_host)
{
	void *w = NULL;
	FLOF_RET(flof_tmp_alloc(ctx, &w, sizeof(float) * (size_t)d.nt, false));
	// w_host is pageable: the runtime stages it before returning, so the caller may free it
	FLOF_CK(cudaMemcpyAsync(w, w_host, sizeof(float) * (size_t)d.nt, cudaMemcpyHostToDevice, ctx->stream));
	FLOF_LAUNCH(k_load_place_rescale, flof_grid4(d), FLOF_BLOCK, 0, phi, d, (const float *)w);
	return flof_tmp_free(ctx, w);
}

// ref :1569-1578: tmp(i,j,k) += sdfIsoOff on the interior (bnd 1), then tmp.setBound(1., 1)
__global__ void __launch_bounds__(FLOF_BLOCK) k_slice_fix(float *__restrict__ a, flof_dim3 d, float off)
{
	const unsigned p = blockIdx.x * FLOF_BLOCK + threadIdx.x;
	if (p >= (unsigned)(d.nx * d.ny)) return;
	const int j = p / d.nx, i = p - j * d.nx, k = blockIdx.y;
	const int64_t c = (int64_t)i + (int64_t)d.nx * (j + (int64_t)d.ny * k);
	const int w = 1;
	const bool bnd = (i <= w || i >= d.nx - 1 - w || j <= w || j >= d.ny - 1 - w ||
	                  (d.nz > 1 && (k <= w || k >= d.nz - 1 - w)));
	if (bnd)
		a[c] = 1.f;
	else
		a[c] += off;
}
extern "C" int flof_slice_fix_levelset(flof_ctx *ctx, float *src3, flof_dim3 sd, float sdfIsoOff)
{
	FLOF_LAUNCH(k_slice_fix, flof_grid3(sd), FLOF_BLOCK, 0, src3, sd, sdfIsoOff);
	return FLOF_OK;
}

extern "C" int flof_load_place_grid4d(flof_ctx *ctx, const float *slices, int nslices, flof_dim3 sd, float *phi,
                                      flof_dim4 d, const float offset[4], const float scale[4],
                                      int fileIdxStart, int fileIdxEnd, int debugSkipLoad, float spread,
                                      const float overrideSize[4], float overrideTimeOff,
                                      int overrideGoodRegion, float loadTimeScale, int rescaleSdfValues,
                                      float sdfIsoOff, float repeatStartFrame)
{
	FLOF_ARG(sd.nx >= 1 && sd.ny >= 1 && sd.nz >= 1, "Invalid src size");
	const int64_t n3 = (int64_t)d.nx * d.ny * d.nz, ns = flof_cells3(sd);
	const float v1[4] = { (float)sd.nx, (float)sd.ny, (float)sd.nz, 1.f };
	float defoSize[4] = { (float)d.nx, (float)d.ny, (float)d.nz, (float)d.nt };
	if (overrideSize[0] > 0.f) memcpy(defoSize, overrideSize, sizeof(defoSize));
	const float m1[4] = { -1.f, -1.f, -1.f, -1.f };
	float sourceFactor[4], off2[4] = { offset[0], offset[1], offset[2], offset[3] };
	flof_grid_factor4d(v1, defoSize, m1, scale, sourceFactor, off2);
	float valueScale = 1.f;
	if (rescaleSdfValues) valueScale = (float)(1. / sourceFactor[0]);
	float *w = (float *)calloc((size_t)d.nt, sizeof(float));
	if (!w) return flof_fail(ctx, FLOF_ERR_NOMEM, "loadPlaceGrid4d: out of host memory");

	int rc = FLOF_OK;
	if (overrideGoodRegion <= 0)
		rc = flof_memset0(ctx, phi, sizeof(float) * (size_t)n3 * d.nt);
	else
		rc = flof_memset0(ctx, phi + n3 * overrideGoodRegion, sizeof(float) * (size_t)n3 * (d.nt - overrideGoodRegion));
	if (fileIdxStart < 0) fileIdxStart = 0;
	if (fileIdxEnd < 0) fileIdxEnd = (int)defoSize[3];
	int repeatOff = 0;
	if (repeatStartFrame > 0.f) repeatOff = (int)((double)((float)(fileIdxEnd - fileIdxStart) * repeatStartFrame) + 1.0);
	void *tmp = NULL;
	if (rc == FLOF_OK) rc = flof_tmp_alloc(ctx, &tmp, sizeof(float) * (size_t)ns, false);

	const int fend = fileIdxEnd < debugSkipLoad ? fileIdxEnd : debugSkipLoad;
	for (int fileid = fileIdxStart - repeatOff; fileid < fend && rc == FLOF_OK; ++fileid) {
		const int fileidClamp = fileid > fileIdxStart ? fileid : fileIdxStart;
		bool didLoad = false;
		const float duration =
		    (float)(1. * (double)(defoSize[3] * scale[3] / loadTimeScale) / (double)(float)(fileIdxEnd - fileIdxStart));
		const float dstt = offset[3] + overrideTimeOff + duration * (float)(fileid - fileIdxStart);
		for (int slice = (int)(dstt - spread * duration) + 0; slice < (int)(dstt + spread * duration) + 2 && rc == FLOF_OK;
		     ++slice) {
			if (slice < overrideGoodRegion || slice >= d.nt) continue;
			const float wc = (float)(1. - (double)(fabsf((float)slice - dstt) / (duration * spread)));
			if (wc < FLOF_VECTOR_EPSILON) continue;
			w[slice] += wc;
			if (!didLoad) {
				if (fileidClamp - fileIdxStart >= nslices) {
					rc = flof_fail(ctx, FLOF_ERR_ARG, "loadPlaceGrid4d: slice %d not provided (have %d)",
					               fileidClamp - fileIdxStart, nslices);
					break;
				}
				const float *src = slices + ns * (fileidClamp - fileIdxStart);
				if (rescaleSdfValues && sdfIsoOff != 0.f) {
					rc = flof_memcpy_d2d(ctx, tmp, src, sizeof(float) * (size_t)ns);
					if (rc == FLOF_OK) rc = flof_slice_fix_levelset(ctx, (float *)tmp, sd, sdfIsoOff);
				}
				didLoad = true;
			}
			const float *use = (rescaleSdfValues && sdfIsoOff != 0.f) ? (const float *)tmp
			                                                          : slices + ns * (fileidClamp - fileIdxStart);
			if (rc == FLOF_OK) rc = flof_load_place_slice(ctx, phi, d, slice, use, sd, sourceFactor, off2, wc);
		}
	}
	if (rc == FLOF_OK) {
		for (int slice = 0; slice < d.nt; ++slice) {
			if (w[slice] < 1e-03f) {
				w[slice] = 0.f;
			} else {
				w[slice] = (float)(1. / w[slice]);
				w[slice] *= valueScale;
			}
		}
		rc = flof_load_place_rescale(ctx, phi, d, w);
	}
	flof_tmp_free(ctx, tmp);
	free(w);
	return rc;
}

extern "C" int flof_shift_forw_grid4d(flof_ctx *ctx, float *phi, flof_dim4 d, int overrideGoodRegion)
{  // ref :1598-1616: block-wise forward memcpy (non-overlapping per block since len <= shift)
	const int64_t n3 = (int64_t)d.nx * d.ny * d.nz;
	const int shift = d.nt - overrideGoodRegion;
	FLOF_ARG(shift > 0, "shiftForwGrid4d: overrideGoodRegion %d >= T %d", overrideGoodRegion, d.nt);
	for (int slice = 0; slice < overrideGoodRegion; slice += shift) {
		int len = shift;
		if (slice + shift >= overrideGoodRegion) len = overrideGoodRegion - slice;
		FLOF_RET(flof_memcpy_d2d(ctx, phi + n3 * slice, phi + n3 * (slice + shift), sizeof(float) * (size_t)n3 * len));
	}
	return FLOF_OK;
}

// ------------------------------------------------------------------ per-frame look-up ------
// v1, v2: two consecutive t-slices of the deformation; vt = (1-tw)*v1 + tw*v2 is formed per
// corner with the reference's fp32 sequence (vt = 0; vt += f1*v1; vt += f2*v2, ref :2011-2013)
// and then trilinearly up-sampled -- identical values to materialising vt first.
struct blend_slice {
	const float4 *v1, *v2;
	float f1, f2;
	__device__ __forceinline__ float4 at(int64_t o) const
	{
		const float4 a = __ldg(v1 + o), b = __ldg(v2 + o);
		float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
		r.x += f1 * a.x; r.y += f1 * a.y; r.z += f1 * a.z; r.w += f1 * a.w;
		r.x += f2 * b.x; r.y += f2 * b.y; r.z += f2 * b.z; r.w += f2 * b.w;
		return r;
	}
};

__device__ __forceinline__ float4 interp3_blend(const blend_slice &s, flof_dim3 d, float x, float y, float z)
{  // ref interpol<Vec4> util/interpol.h:101-116
	const float px = x - 0.5f, py = y - 0.5f, pz = z - 0.5f;
	int xi = (int)px, yi = (int)py, zi = (int)pz;
	float s1 = px - (float)xi, s0 = 1.0f - s1;
	float t1 = py - (float)yi, t0 = 1.0f - t1;
	float f1 = pz - (float)zi, f0 = 1.0f - f1;
	if (px < 0.f) { xi = 0; s0 = 1.f; s1 = 0.f; }
	if (py < 0.f) { yi = 0; t0 = 1.f; t1 = 0.f; }
	if (pz < 0.f) { zi = 0; f0 = 1.f; f1 = 0.f; }
	if (xi >= d.nx - 1) { xi = d.nx - 2; s0 = 0.f; s1 = 1.f; }
	if (yi >= d.ny - 1) { yi = d.ny - 2; t0 = 0.f; t1 = 1.f; }
	if (d.nz > 1) {
		if (zi >= d.nz - 1) { zi = d.nz - 2; f0 = 0.f; f1 = 1.f; }
	}
	const int64_t X = 1, Y = d.nx, Z = (int64_t)d.nx * d.ny;
	const int64_t o = (int64_t)xi + Y * yi + Z * zi;
	return ((s.at(o) * t0 + s.at(o + Y) * t1) * s0 + (s.at(o + X) * t0 + s.at(o + X + Y) * t1) * s1) * f0 +
	       ((s.at(o + Z) * t0 + s.at(o + Y + Z) * t1) * s0 + (s.at(o + X + Z) * t0 + s.at(o + X + Y + Z) * t1) * s1) * f1;
}

__global__ void __launch_bounds__(FLOF_BLOCK)
    k_lookup_slice4d(float *__restrict__ dst, flof_dim3 d3, const float *__restrict__ phi, flof_dim4 d, float time,
                     float dt, blend_slice vs, flof_dim3 vd, float3 fac3, float3 off3, float4 fac, int bord)
{
	const unsigned p = blockIdx.x * FLOF_BLOCK + threadIdx.x;
	if (p >= (unsigned)(d3.nx * d3.ny)) return;
	const int j = p / d3.nx, i = p - j * d3.nx, k = blockIdx.y;
	if (i < bord || j < bord || k < bord || i >= d3.nx - bord || j >= d3.ny - bord || k >= d3.nz - bord) return;
	const float x = (float)i * fac3.x + off3.x, y = (float)j * fac3.y + off3.y, z = (float)k * fac3.z + off3.z;
	float4 v = interp3_blend(vs, vd, x, y, z);
	v.x *= fac.x; v.y *= fac.y; v.z *= fac.z; v.w *= fac.w;  // vdst *= defoFactor
	const float qx = ((float)i + 0.5f) - v.x * dt, qy = ((float)j + 0.5f) - v.y * dt;
	const float qz = ((float)k + 0.5f) - v.z * dt, qt = (time + 0.5f) - v.w * dt;
	dst[(int64_t)i + (int64_t)d3.nx * (j + (int64_t)d3.ny * k)] = flof_interpol4d<float>(phi, d, qx, qy, qz, qt);
}

static int lookup_launch(flof_ctx *ctx, float *dst3, flof_dim3 d3, const float *phi, flof_dim4 d, float time,
                         float dt, const float *v1, const float *v2, float f1, float f2, flof_dim3 vd,
                         const float srcFac3[3], const float off3[3], const float fac[4], int bordSkip, int kernelBnd = 10)
{
	FLOF_ARG(vd.nx >= 2 && vd.ny >= 2, "loadAdvectTimeSlice: Invalid src size");
	FLOF_ARG(d.nx >= 2 && d.ny >= 2 && d.nz >= 2 && d.nt >= 2, "loadAdvectTimeSlice: phi grid too small");
	blend_slice vs;
	vs.v1 = (const float4 *)v1;
	vs.v2 = (const float4 *)v2;
	vs.f1 = f1;
	vs.f2 = f2;
	const int bord = bordSkip > kernelBnd ? bordSkip : kernelBnd;  // KERNEL(bnd = 10) and the bordSkip test (ref :1648, 1656)
	FLOF_LAUNCH(k_lookup_slice4d, flof_grid3(d3), FLOF_BLOCK, 0, dst3, d3, phi, d, time, dt, vs, vd,
	            make_float3(srcFac3[0], srcFac3[1], srcFac3[2]), make_float3(off3[0], off3[1], off3[2]),
	            make_float4(fac[0], fac[1], fac[2], fac[3]), bord);
	return FLOF_OK;
}

extern "C" int flof_lookup_slice4d_with_vel(flof_ctx *ctx, float *dst3, flof_dim3 d3, const float *phi, flof_dim4 d,
                                            float time, float dt, const float *vslice, flof_dim3 vd,
                                            const float srcFac3[3], const float off3[3], const float fac[4],
                                            int bordSkip)
{
	// a materialised slice: vt = 0 + 1*v + 0*v is not bit-neutral for -0/inf; use f1 = 1 on v and
	// a second term with weight 0 on the same finite data (exact for finite values)
	return lookup_launch(ctx, dst3, d3, phi, d, time, dt, vslice, vslice, 1.f, 0.f, vd, srcFac3, off3, fac, bordSkip);
}

extern "C" int flof_load_advect_time_slice(flof_ctx *ctx, const float *defo, flof_dim4 dd, float *dst3, flof_dim3 d3,
                                           const float *phi, flof_dim4 d, float time, float blendAlpha,
                                           float loadTimeScale, const float defoOffset[4], const float defoScale[4],
                                           const float defoFactor[4], const float overrideSize[4],
                                           float overrideTimeOff, int bordSkip, float defoAniFac)
{
	const int dimT = dd.nt;
	FLOF_ARG(dd.nx >= 1 && dd.ny >= 1 && dd.nz >= 1, "Invalid src size");
	const float dim4[4] = { (float)dd.nx, (float)dd.ny, (float)dd.nz, (float)dimT };
	float defoSize[4] = { (float)d.nx, (float)d.ny, (float)d.nz, (float)d.nt };
	if (overrideSize[0] > 0.f)
		for (int c = 0; c < 4; ++c) defoSize[c] = (float)(int)overrideSize[c];  // Vec4i truncation ref :1970-1973
	const float m1[4] = { -1.f, -1.f, -1.f, -1.f };
	float sourceFactor[4], off2[4] = { defoOffset[0], defoOffset[1], defoOffset[2], defoOffset[3] };
	flof_grid_factor4d(dim4, defoSize, m1, defoScale, sourceFactor, off2);
	// ref :1980-1985
	const volatile float a = time * sourceFactor[3];
	const volatile float b = a * loadTimeScale;
	const volatile float c2 = b + off2[3];
	const float srcTime = (float)((double)c2 - 0.5);
	int t = (int)srcTime;
	int tp1 = t + 1;
	const float tw = srcTime - (float)t;
	t = t < dimT - 1 ? t : dimT - 1;
	tp1 = tp1 < dimT - 1 ? tp1 : dimT - 1;
	FLOF_ARG(t >= 0, "loadAdvectTimeSlice: source time %g before the deformation volume", (double)srcTime);
	const int64_t nv = (int64_t)dd.nx * dd.ny * dd.nz;
	const float f1 = (float)(1. - (double)tw), f2 = tw;  // Vec4(1. - tw), Vec4(tw)
	float fac[4];
	for (int c = 0; c < 4; ++c) fac[c] = defoFactor[c] * defoAniFac;
	const flof_dim3 vd = { dd.nx, dd.ny, dd.nz };
	return lookup_launch(ctx, dst3, d3, phi, d, time + overrideTimeOff, blendAlpha, defo + nv * 4 * t,
	                     defo + nv * 4 * tp1, f1, f2, vd, sourceFactor, off2, fac, bordSkip);
}

// ------------------------------------------------------------------ deformation volumes ----
// ref: LoadAdvectData::updateDefoVol :1822-1863 and the defo-volume branches of loadAdvectTimeSlice_OptRun :2015-2089
// (flof.py `thirdload`: two or three deformations composed on the fly).  Every deformation volume stays resident on the
// device; the reference's window of Tw = int(dimT * max(0.2, partialLoadFac)) time slices around the current source
// time is materialised from it with the reference's own update rule -- including its start-up quirk: the very first
// request at t = 0 takes the "one slice further" branch (lastT + 1 == t) and fills only the last window slice.

// the three cases of updateDefoVol for one volume; win: Tw slices of nv cells, vol: dimT slices.
// Slices travel through the one-slice scratch grid vt like in the reference (lats.tmp, which is also the output of the
// composition), and the "one slice further" branch reads through the reference's re-used gz handle: the FIRST such read
// opens the file and seeks to the requested slice, every later one just takes the next slice of the file whatever index
// is asked for, and past the end of the file nothing is read, so that vt keeps its content (readGrid4dUni, fileio.cpp:
// 903-919).  *filepos: next slice of that handle, -1 = not open yet.  Identical to plain indexing whenever the caller
// advances one source slice at a time, which is what flof.py does.
extern "C" int flof_defovol_window_update(flof_ctx *ctx, float *win, flof_dim4 wd, const float *vol, int dimT, int t,
                                          int lastT, float *vt, int *filepos)
{
	FLOF_ARG(wd.nt >= 1 && dimT >= 1, "defo volume: empty window");
	FLOF_ARG(vt != NULL && filepos != NULL, "defo volume: scratch slice / file position missing");
	const int Tw = wd.nt, currt = Tw / 2;
	const size_t sb = sizeof(float) * 4 * (size_t)wd.nx * wd.ny * wd.nz;
	char *w = (char *)win;
	const char *v = (const char *)vol;
	if (lastT == t) return FLOF_OK;
	if (lastT + 1 == t) {
		// shift all back by one slice (ascending: every slice takes its successor's content), then read the last one
		for (int tl = 0; tl < Tw - 1; ++tl)
			FLOF_CK(cudaMemcpyAsync(w + sb * tl, w + sb * (tl + 1), sb, cudaMemcpyDeviceToDevice, ctx->stream));
		const int tl = Tw - 1;
		int it = min(max(t - currt + tl, 0), dimT - 1);
		if (*filepos >= 0) it = *filepos;  // re-used handle: sequential read
		if (it < dimT) FLOF_CK(cudaMemcpyAsync(vt, v + sb * it, sb, cudaMemcpyDeviceToDevice, ctx->stream));
		*filepos = it < dimT ? it + 1 : it;
		FLOF_CK(cudaMemcpyAsync(w + sb * tl, vt, sb, cudaMemcpyDeviceToDevice, ctx->stream));
		return FLOF_OK;
	}
	for (int tl = 0; tl < Tw; ++tl) {
		const int it = min(max(t - currt + tl, 0), dimT - 1);
		FLOF_CK(cudaMemcpyAsync(w + sb * tl, v + sb * it, sb, cudaMemcpyDeviceToDevice, ctx->stream));
	}
	// (the reference read these through vt as well: it ends up holding the last slice)
	FLOF_CK(cudaMemcpyAsync(vt, w + sb * (Tw - 1), sb, cudaMemcpyDeviceToDevice, ctx->stream));
	return FLOF_OK;
}

__device__ __forceinline__ float4 f4_scale(float s, const float4 &v) { return make_float4(s * v.x, s * v.y, s * v.z, s * v.w); }
__device__ __forceinline__ float4 f4_add(const float4 &a, const float4 &b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// MODE 0: two volumes, "aligned w/o backmove" :2019-2033      vt = bA * dvol1(q - v2) + tA * v2,  v2 = dvol2(q)
// MODE 1: three volumes :2063-2081                            vt = bA * v1 + tA * v2 + fA * v3 (each looked up behind the later ones)
// MODE 2: second stage of "align & move back" :2052-2060      vt = dvt(q - (1 - bA) * (-dvol1(q)))
// MODE 3: first stage of the same :2043-2051 over the WHOLE window grid (t = slice index): dvt = bA * dvol1(q - v2) + tA * v2
template <int MODE>
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_defovol_compose(float4 *__restrict__ out, const float4 *__restrict__ dv1, const float4 *__restrict__ dv2,
                      const float4 *__restrict__ dv3, flof_dim4 wd, float tcoord, float bA, float tA, float fA)
{
	const unsigned p = blockIdx.x * FLOF_BLOCK + threadIdx.x;
	if (p >= (unsigned)(wd.nx * wd.ny)) return;
	const int j = p / wd.nx, i = p - j * wd.nx, k = blockIdx.y, tl = blockIdx.z;
	const float tc = MODE == 3 ? (float)tl : tcoord;
	const float qx = (float)i + 0.5f, qy = (float)j + 0.5f, qz = (float)k + 0.5f, qt = tc + 0.5f;
	float4 r;
	if (MODE == 0 || MODE == 3) {
		const float4 v2 = flof_interpol4d<float4>(dv2, wd, qx, qy, qz, qt);
		const float4 a = flof_interpol4d<float4>(dv1, wd, qx - v2.x, qy - v2.y, qz - v2.z, qt - v2.w);
		r = f4_add(f4_scale(bA, a), f4_scale(tA, v2));
	} else if (MODE == 1) {
		const float4 v3 = flof_interpol4d<float4>(dv3, wd, qx, qy, qz, qt);
		const float rx = qx - v3.x, ry = qy - v3.y, rz = qz - v3.z, rt = qt - v3.w;
		const float4 v2 = flof_interpol4d<float4>(dv2, wd, rx, ry, rz, rt);
		const float4 v1 = flof_interpol4d<float4>(dv1, wd, rx - v2.x, ry - v2.y, rz - v2.z, rt - v2.w);
		r = f4_add(f4_add(f4_scale(bA, v1), f4_scale(tA, v2)), f4_scale(fA, v3));
	} else {
		// dv2 = dvt here.  v1 = dvol1(q) * -1; offset (1. - blendAlpha) * v1: a double scalar times a float vector,
		// rounded to float per component (Vector4D operator*(S2, Vector4D<S>), util/vector4d.h:244-248)
		const float4 a = flof_interpol4d<float4>(dv1, wd, qx, qy, qz, qt);
		const double w = 1. - (double)bA;
		const float sx = (float)(w * (double)(a.x * -1.f)), sy = (float)(w * (double)(a.y * -1.f));
		const float sz = (float)(w * (double)(a.z * -1.f)), st = (float)(w * (double)(a.w * -1.f));
		r = flof_interpol4d<float4>(dv2, wd, qx - sx, qy - sy, qz - sz, qt - st);
	}
	out[(int64_t)i + (int64_t)wd.nx * (j + (int64_t)wd.ny * (k + (int64_t)wd.nz * tl))] = r;
}

// vt (one slice, nx*ny*nz Vec4) from the windows of 2 or 3 deformation volumes.  dvt: scratch of the window's size, only
// used (and required) with doAligned.  tcoord = srcTime - defovolOff.
extern "C" int flof_defovol_compose(flof_ctx *ctx, float *vt, const float *dvol1, const float *dvol2, const float *dvol3,
                                    float *dvt, flof_dim4 wd, float tcoord, int doAligned, float blendAlpha, float thirdAlpha,
                                    float fourthAlpha)
{
	FLOF_ARG(wd.nx >= 2 && wd.ny >= 2 && wd.nz >= 2 && wd.nt >= 2, "defo volumes: window too small");
	FLOF_ARG(dvol1 && dvol2, "Code currently only supports 2 deformation volumes");  // ref :2083
	const dim3 g1((unsigned)((wd.nx * wd.ny + FLOF_BLOCK - 1) / FLOF_BLOCK), (unsigned)wd.nz, 1);
	const float4 *d1 = (const float4 *)dvol1, *d2 = (const float4 *)dvol2, *d3 = (const float4 *)dvol3;
	if (dvol3) {
		FLOF_LAUNCH(k_defovol_compose<1>, g1, FLOF_BLOCK, 0, (float4 *)vt, d1, d2, d3, wd, tcoord, blendAlpha, thirdAlpha, fourthAlpha);
	} else if (!doAligned) {
		FLOF_LAUNCH(k_defovol_compose<0>, g1, FLOF_BLOCK, 0, (float4 *)vt, d1, d2, d3, wd, tcoord, blendAlpha, thirdAlpha, fourthAlpha);
	} else {
		FLOF_ARG(dvt != NULL, "defo volumes: the aligned composition needs the window-sized scratch grid");
		const dim3 gw(g1.x, (unsigned)wd.nz, (unsigned)wd.nt);
		FLOF_LAUNCH(k_defovol_compose<3>, gw, FLOF_BLOCK, 0, (float4 *)dvt, d1, d2, d3, wd, 0.f, blendAlpha, thirdAlpha, fourthAlpha);
		FLOF_LAUNCH(k_defovol_compose<2>, g1, FLOF_BLOCK, 0, (float4 *)vt, d1, (const float4 *)dvt, d3, wd, tcoord, blendAlpha, thirdAlpha, fourthAlpha);
	}
	return FLOF_OK;
}

// the source-time arithmetic of loadAdvectTimeSlice[_OptRun] (ref :1962-1985) for callers that manage the deformation
// slices / windows themselves: srcTime, the two slice indices, the blend weight and the 3D look-up transform
extern "C" int flof_lats_source_time(flof_dim4 dd, flof_dim4 d, float time, float loadTimeScale, const float defoOffset[4],
                                     const float defoScale[4], const float overrideSize[4], float *srcTimeOut, int *tOut,
                                     int *tp1Out, float *twOut, float srcFac3[3], float off3[3])
{
	const int dimT = dd.nt;
	const float dim4[4] = { (float)dd.nx, (float)dd.ny, (float)dd.nz, (float)dimT };
	float defoSize[4] = { (float)d.nx, (float)d.ny, (float)d.nz, (float)d.nt };
	if (overrideSize[0] > 0.f)
		for (int c = 0; c < 4; ++c) defoSize[c] = (float)(int)overrideSize[c];
	const float m1[4] = { -1.f, -1.f, -1.f, -1.f };
	float sourceFactor[4], off2[4] = { defoOffset[0], defoOffset[1], defoOffset[2], defoOffset[3] };
	flof_grid_factor4d(dim4, defoSize, m1, defoScale, sourceFactor, off2);
	const volatile float a = time * sourceFactor[3];
	const volatile float b = a * loadTimeScale;
	const volatile float c2 = b + off2[3];
	const float srcTime = (float)((double)c2 - 0.5);
	int t = (int)srcTime;
	int tp1 = t + 1;
	const float tw = srcTime - (float)t;
	t = t < dimT - 1 ? t : dimT - 1;
	tp1 = tp1 < dimT - 1 ? tp1 : dimT - 1;
	if (srcTimeOut) *srcTimeOut = srcTime;
	if (tOut) *tOut = t;
	if (tp1Out) *tp1Out = tp1;
	if (twOut) *twOut = tw;
	for (int c = 0; c < 3; ++c) {
		if (srcFac3) srcFac3[c] = sourceFactor[c];
		if (off3) off3[c] = off2[c];
	}
	return 0;
}

// debug outputs of the unoptimised loadAdvectTimeSlice (ref :1731-1753): the time-blended deformation slice
// re-interpolated to the output size and scaled, split into a Vec3 grid (3 floats per cell) and a Real grid.
// Interior cells only (FOR_IJK_BND(vdst, 1) over a cleared grid): the outer shell is zero.
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_slice_vel_debug(float *__restrict__ dbgVel3, float *__restrict__ dbgVelT, flof_dim3 d3, blend_slice vs, flof_dim3 vd,
                      float3 fac3, float3 off3, float4 fac)
{
	const unsigned p = blockIdx.x * FLOF_BLOCK + threadIdx.x;
	if (p >= (unsigned)(d3.nx * d3.ny)) return;
	const int j = p / d3.nx, i = p - j * d3.nx, k = blockIdx.y;
	float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
	if (i >= 1 && j >= 1 && k >= 1 && i < d3.nx - 1 && j < d3.ny - 1 && k < d3.nz - 1) {
		const float x = (float)i * fac3.x + off3.x, y = (float)j * fac3.y + off3.y, z = (float)k * fac3.z + off3.z;
		v = interp3_blend(vs, vd, x, y, z);
		v.x *= fac.x; v.y *= fac.y; v.z *= fac.z; v.w *= fac.w;
	}
	const int64_t c = (int64_t)i + (int64_t)d3.nx * (j + (int64_t)d3.ny * k);
	if (dbgVel3) {
		dbgVel3[3 * c] = v.x;
		dbgVel3[3 * c + 1] = v.y;
		dbgVel3[3 * c + 2] = v.z;
	}
	if (dbgVelT) dbgVelT[c] = v.w;
}
// The unoptimised loadAdvectTimeSlice (ref :1671-1760): v1 *= (1 - tw); v2 *= tw; v1 += v2; the slice is re-interpolated
// to the output size for the interior cells (vdst, outer shell zero), scaled by defoFactor * defoAniFac, optionally
// returned as debugVel / debugVelT, and knSemiLagrangeLookupSlice4d (KERNEL(fourd, bnd = 1) on a one-slice grid: the
// generated code runs it as a 3D kernel with t = 0) looks every interior cell of dst up in phi with dt = blendAlpha.
// Same arithmetic as the optimised path with a border of 1 instead of max(10, bordSkip).
extern "C" int flof_load_advect_time_slice_unopt(flof_ctx *ctx, const float *defo, flof_dim4 dd, float *dst3, flof_dim3 d3,
                                                 const float *phi, flof_dim4 d, float time, float blendAlpha,
                                                 float loadTimeScale, const float defoOffset[4], const float defoScale[4],
                                                 const float defoFactor[4], const float overrideSize[4],
                                                 float overrideTimeOff, float defoAniFac, int zeroVel, float *dbgVel3,
                                                 float *dbgVelT)
{
	float srcTime, tw, sf3[3], off3[3];
	int t, tp1;
	flof_lats_source_time(dd, d, time, loadTimeScale, defoOffset, defoScale, overrideSize, &srcTime, &t, &tp1, &tw, sf3, off3);
	FLOF_ARG(t >= 0, "loadAdvectTimeSlice: source time %g before the deformation volume", (double)srcTime);
	FLOF_ARG(dd.nx >= 2 && dd.ny >= 2, "loadAdvectTimeSlice: Invalid src size");
	const int64_t nv = (int64_t)dd.nx * dd.ny * dd.nz;
	const float f1 = zeroVel ? 0.f : (float)(1. - (double)tw), f2 = zeroVel ? 0.f : tw;
	float fac[4];
	for (int c = 0; c < 4; ++c) fac[c] = defoFactor[c] * defoAniFac;
	const flof_dim3 vd = { dd.nx, dd.ny, dd.nz };
	if (dbgVel3 || dbgVelT) {
		blend_slice vs;
		vs.v1 = (const float4 *)(defo + nv * 4 * t);
		vs.v2 = (const float4 *)(defo + nv * 4 * tp1);
		vs.f1 = f1;
		vs.f2 = f2;
		FLOF_LAUNCH(k_slice_vel_debug, flof_grid3(d3), FLOF_BLOCK, 0, dbgVel3, dbgVelT, d3, vs, vd,
		            make_float3(sf3[0], sf3[1], sf3[2]), make_float3(off3[0], off3[1], off3[2]),
		            make_float4(fac[0], fac[1], fac[2], fac[3]));
	}
	return lookup_launch(ctx, dst3, d3, phi, d, time + overrideTimeOff, blendAlpha, defo + nv * 4 * t, defo + nv * 4 * tp1, f1,
	                     f2, vd, sf3, off3, fac, 1, 1);
}

// ------------------------------------------------------------------ 3D output blur ---------
// ref knSimpleBlurSpecial test.cpp:93-115: KERNEL(bnd=1); outside `bord` or below thresh copy,
// else 6-neighbour mean.  The 1-cell shell of tmp is never written (stays as allocated).
__global__ void __launch_bounds__(FLOF_BLOCK)
    k_simple_blur_special(const float *__restrict__ a, float *__restrict__ tmp, flof_dim3 d, float thresh, int bord)
{
	const unsigned p = blockIdx.x * FLOF_BLOCK + threadIdx.x;
	if (p >= (unsigned)(d.nx * d.ny)) return;
	const int j = p / d.nx, i = p - j * d.nx, k = blockIdx.y;
	if (i < 1 || j < 1 || k < 1 || i >= d.nx - 1 || j >= d.ny - 1 || k >= d.nz - 1) return;
	const int64_t sY = d.nx, sZ = (int64_t)d.nx * d.ny;
	const int64_t c = (int64_t)i + sY * j + sZ * k;
	const float v0 = __ldg(a + c);
	const bool inb = i >= bord && j >= bord && k >= bord && i < d.nx - bord && j < d.ny - bord && k < d.nz - bord;
	if (!inb || v0 < thresh) {
		tmp[c] = v0;
		return;
	}
	float v = __ldg(a + c + 1) + __ldg(a + c - 1) + __ldg(a + c + sY) + __ldg(a + c - sY);
	v += __ldg(a + c + sZ) + __ldg(a + c - sZ);
	tmp[c] = (float)((double)v * (1. / 6.));
}
extern "C" int flof_simple_blur_special(flof_ctx *ctx, float *a, flof_dim3 d, int iter, float thresh, int bord)
{
	if (iter <= 0) return FLOF_OK;
	FLOF_ARG(d.nz > 1, "simpleBlurSpecial: only 3D grids are on the FlOF path");
	const size_t bytes = sizeof(float) * (size_t)flof_cells3(d);
	void *tmp = NULL;
	FLOF_RET(flof_tmp_alloc(ctx, &tmp, bytes, true));  // Grid<S> tmp(parent): zero border
	float *cur = a, *oth = (float *)tmp;
	for (int it = 0; it < iter; ++it) {
		FLOF_LAUNCH(k_simple_blur_special, flof_grid3(d), FLOF_BLOCK, 0, (const float *)cur, oth, d, thresh, bord);
		float *sw = cur; cur = oth; oth = sw;
	}
	int rc = FLOF_OK;
	if (cur != a) rc = flof_memcpy_d2d(ctx, a, cur, bytes);
	flof_tmp_free(ctx, tmp);
	return rc;
}
