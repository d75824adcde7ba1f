// nyxus_b200 -- CUDA kernels (sm_100a): phase-1 label scan, ROI compaction, CSR packing, per-ROI family kernels.
#pragma once
#include <cuda_runtime.h>
#include "nyx_common.h"
#include "nyx_features.h"
#include "glcm_body.h"
#include "glrlm_body.h"
#include "zones_body.h"
#include "nbr_body.h"
#include "intensity_body.h"
#include "vol_body.h"

namespace nyx
{

// ------------------------------------------------------------------------------------------------
// Phase 1: one streaming pass over the slide -> per-label {area, min, max, AABB}
// (reference: gatherRoisMetricsInMemory phase1.cpp:373-407 + feed_pixel_2_metrics pixel_feed.cpp:19-43 +
//  init/update_label_record_3 features_calc_workflow.cpp:65-105: two hash probes per foreground pixel.)
// Here: a dense label-indexed table in HBM (32 B per label = one sector), 128-bit loads, and warp-aggregated
// updates: lanes holding the same label are grouped with match.any, reduced with redux.sync, and ONE lane issues
// the 7 atomics of the group.
// ------------------------------------------------------------------------------------------------
struct __align__(32) LabelStat
{
	uint32_t area, imin, imax, xmin, xmax, ymin, ymax, pad;
};

// info block shared between device and host (one D2H copy per slide)
constexpr int NCLASS = 5;		// tile classes by cell count
constexpr int NACLASS = 5;		// cloud classes by area
struct SlideInfo
{
	uint32_t max_label;		// largest label value seen by the scan
	uint32_t n_roi;
	unsigned long long n_fg;	// foreground pixels = sum of areas
	uint32_t max_inten;		// largest aux_max over ROIs (IBSI level bound)
	uint32_t class_count[NCLASS + 1];	// last = oversize
	uint32_t class_maxcells[NCLASS + 1];
	uint32_t class_maxwh[NCLASS + 1];
	uint32_t aclass_count[NACLASS + 1];
	uint32_t aclass_maxarea[NACLASS + 1];
	uint32_t inv_min_inten;		// ~(smallest aux_min over ROIs): the slide's foreground minimum (3D COVERED_IMAGE_INTENSITY_RANGE)
};

__host__ __device__ inline int tile_class (uint32_t ncell)
{
	return ncell <= 256 ? 0 : ncell <= 1024 ? 1 : ncell <= 4096 ? 2 : ncell <= 16384 ? 3 : ncell <= 40000 ? 4 : NCLASS;
}
__host__ __device__ inline int area_class (uint32_t area)
{
	return area <= 256 ? 0 : area <= 1024 ? 1 : area <= 4096 ? 2 : area <= 16384 ? 3 : area <= 32768 ? 4 : NACLASS;
}

__global__ void k_init_stats (LabelStat* tab, uint32_t cap, SlideInfo* info, uint2* ztab = nullptr)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < cap)
	{
		LabelStat s;
		s.area = 0; s.imin = 0xFFFFFFFFu; s.imax = 0; s.xmin = 0xFFFFFFFFu; s.xmax = 0; s.ymin = 0xFFFFFFFFu; s.ymax = 0; s.pad = 0;
		tab[i] = s;
		if (ztab)
			ztab[i] = make_uint2 (0xFFFFFFFFu, 0u);
	}
	if (i == 0)
		memset (info, 0, sizeof (SlideInfo));
}

__device__ __forceinline__ void stat_group_update (unsigned part, LabelStat* tab, uint32_t cap, uint32_t lbl, uint32_t cnt,
	uint32_t vmin, uint32_t vmax, uint32_t xmin, uint32_t xmax, uint32_t y, uint2* ztab = nullptr, uint32_t H = 0)
{
	// volumes: `y` arrives as the row index over all slices (z*H + y); lanes of one group may sit in different slices
	uint32_t z = 0;
	if (ztab)
	{
		z = y / H;
		y -= z * H;
	}
	// `part` = lanes taking part (same value in all of them); lanes with the same label reduce together, the lowest commits
	const unsigned mask = __match_any_sync (part, lbl);
	const int lane = threadIdx.x & 31;
	uint32_t area = __reduce_add_sync (mask, cnt);
	vmin = __reduce_min_sync (mask, vmin);
	vmax = __reduce_max_sync (mask, vmax);
	xmin = __reduce_min_sync (mask, xmin);
	xmax = __reduce_max_sync (mask, xmax);
	uint32_t ymin = __reduce_min_sync (mask, y);
	uint32_t ymax = __reduce_max_sync (mask, y);
	if (lane == __ffs (mask) - 1 && lbl < cap)
	{
		LabelStat* s = tab + lbl;
		atomicAdd (&s->area, area);
		atomicMin (&s->imin, vmin);
		atomicMax (&s->imax, vmax);
		atomicMin (&s->xmin, xmin);
		atomicMax (&s->xmax, xmax);
		atomicMin (&s->ymin, ymin);
		atomicMax (&s->ymax, ymax);
	}
	if (ztab)
	{
		const uint32_t zmin = __reduce_min_sync (mask, z), zmax = __reduce_max_sync (mask, z);
		if (lane == __ffs (mask) - 1 && lbl < cap)
		{
			atomicMin (&ztab[lbl].x, zmin);
			atomicMax (&ztab[lbl].y, zmax);
		}
	}
}

template <bool U16, bool VEC>
__global__ void __launch_bounds__(256) k_scan_stats (const void* __restrict__ inten_, const uint32_t* __restrict__ label,
	unsigned long long npix, uint32_t W, LabelStat* tab, uint32_t cap, SlideInfo* info, uint2* ztab = nullptr, uint32_t H = 0)
{
	const unsigned long long ngroups = (npix + 3) / 4;
	uint32_t warp_maxlbl = 0;
	for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; ; g += (unsigned long long)gridDim.x * blockDim.x)
	{
		// keep whole warps together: the loop bound is evaluated on the warp's first group
		const unsigned long long gw = g - (threadIdx.x & 31);
		if (gw >= ngroups)
			break;
		uint32_t l[4] = { 0, 0, 0, 0 }, v[4] = { 0, 0, 0, 0 };
		const unsigned long long p0 = g * 4;
		if (g < ngroups)
		{
			if (VEC && p0 + 3 < npix)
			{
				uint4 lv = *reinterpret_cast<const uint4*> (label + p0);
				l[0] = lv.x; l[1] = lv.y; l[2] = lv.z; l[3] = lv.w;
				if (l[0] | l[1] | l[2] | l[3])
				{
					if (U16)
					{
						ushort4 iv = *reinterpret_cast<const ushort4*> ((const uint16_t*)inten_ + p0);
						v[0] = iv.x; v[1] = iv.y; v[2] = iv.z; v[3] = iv.w;
					}
					else
					{
						uint4 iv = *reinterpret_cast<const uint4*> ((const uint32_t*)inten_ + p0);
						v[0] = iv.x; v[1] = iv.y; v[2] = iv.z; v[3] = iv.w;
					}
				}
			}
			else
				for (int k = 0; k < 4; k++)
					if (p0 + k < npix)
					{
						l[k] = label[p0 + k];
						v[k] = U16 ? (uint32_t)((const uint16_t*)inten_)[p0 + k] : ((const uint32_t*)inten_)[p0 + k];
					}
		}
		const uint32_t y0 = (uint32_t)(p0 / W), x0 = (uint32_t)(p0 - (unsigned long long)y0 * W);
		const bool uni = (l[0] == l[1]) && (l[1] == l[2]) && (l[2] == l[3]) && (x0 + 3 < W);
		uint32_t m = max (max (l[0], l[1]), max (l[2], l[3]));
		warp_maxlbl = max (warp_maxlbl, m);
		if (__all_sync (0xffffffffu, uni))
		{
			const unsigned part = __ballot_sync (0xffffffffu, l[0] != 0);
			if (l[0] != 0)
			{
				uint32_t vmin = min (min (v[0], v[1]), min (v[2], v[3])), vmax = max (max (v[0], v[1]), max (v[2], v[3]));
				stat_group_update (part, tab, cap, l[0], 4, vmin, vmax, x0, x0 + 3, y0, ztab, H);
			}
		}
		else
		{
#pragma unroll
			for (int k = 0; k < 4; k++)
			{
				uint32_t x = x0 + k, y = y0;
				if (x >= W)
				{
					x -= W;
					y += 1;
					if (x >= W)	// W < 4: fall back to exact division
					{
						y = (uint32_t)((p0 + k) / W);
						x = (uint32_t)((p0 + k) - (unsigned long long)y * W);
					}
				}
				const unsigned part = __ballot_sync (0xffffffffu, l[k] != 0);
				if (l[k] != 0)
					stat_group_update (part, tab, cap, l[k], 1, v[k], v[k], x, x, y, ztab, H);
			}
		}
	}
	warp_maxlbl = __reduce_max_sync (0xffffffffu, warp_maxlbl);
	if ((threadIdx.x & 31) == 0 && warp_maxlbl)
		atomicMax (&info->max_label, warp_maxlbl);
}

// ------------------------------------------------------------------------------------------------
// ROI compaction: dense label table -> RoiRec array in ascending label order (reference: rows sorted by label,
// output_2_buffer.cpp:305-306) + exclusive scan of the areas (CSR offsets) + size classification.
// Three small kernels over the table (count per 1024-label block, scan of block totals, scatter).
// ------------------------------------------------------------------------------------------------
constexpr int CB = 1024;

__device__ __forceinline__ void block_exscan_2 (uint32_t& a, unsigned long long& b, uint32_t* sa, unsigned long long* sb,
	uint32_t& tot_a, unsigned long long& tot_b)
{
	// exclusive scan over the block of (a, b); sa/sb: 32-entry shared scratch each
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t ia = a;
	unsigned long long ib = b;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		uint32_t ta = __shfl_up_sync (0xffffffffu, ia, o);
		unsigned long long tb = __shfl_up_sync (0xffffffffu, ib, o);
		if (lane >= o)
		{
			ia += ta;
			ib += tb;
		}
	}
	if (lane == 31)
	{
		sa[warp] = ia;
		sb[warp] = ib;
	}
	__syncthreads();
	if (warp == 0)
	{
		uint32_t wa = lane < (blockDim.x >> 5) ? sa[lane] : 0;
		unsigned long long wb = lane < (blockDim.x >> 5) ? sb[lane] : 0;
		uint32_t xa = wa;
		unsigned long long xb = wb;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			uint32_t ta = __shfl_up_sync (0xffffffffu, xa, o);
			unsigned long long tb = __shfl_up_sync (0xffffffffu, xb, o);
			if (lane >= o)
			{
				xa += ta;
				xb += tb;
			}
		}
		sa[lane] = xa - wa;	// exclusive warp offsets
		sb[lane] = xb - wb;
		if (lane == 31)
		{
			sa[32] = xa;
			sb[32] = xb;
		}
	}
	__syncthreads();
	a = sa[warp] + ia - a;
	b = sb[warp] + ib - b;
	tot_a = sa[32];
	tot_b = sb[32];
	__syncthreads();
}

__global__ void __launch_bounds__(CB) k_compact_count (const LabelStat* tab, uint32_t cap, const SlideInfo* info,
	uint32_t* blk_cnt, unsigned long long* blk_area)
{
	__shared__ uint32_t sa[33];
	__shared__ unsigned long long sb[33];
	const uint32_t i = blockIdx.x * CB + threadIdx.x;
	uint32_t a = 0;
	unsigned long long b = 0;
	if (i < cap && i <= info->max_label && i > 0)
	{
		uint32_t ar = tab[i].area;
		a = ar ? 1 : 0;
		b = ar;
	}
	uint32_t ta;
	unsigned long long tb;
	block_exscan_2 (a, b, sa, sb, ta, tb);
	if (threadIdx.x == 0)
	{
		blk_cnt[blockIdx.x] = ta;
		blk_area[blockIdx.x] = tb;
	}
}

__global__ void __launch_bounds__(CB) k_compact_scan (uint32_t* blk_cnt, unsigned long long* blk_area, uint32_t nblk, SlideInfo* info)
{
	__shared__ uint32_t sa[33];
	__shared__ unsigned long long sb[33];
	uint32_t run_a = 0;
	unsigned long long run_b = 0;
	for (uint32_t base = 0; base < nblk; base += CB)
	{
		uint32_t i = base + threadIdx.x;
		uint32_t a = i < nblk ? blk_cnt[i] : 0;
		unsigned long long b = i < nblk ? blk_area[i] : 0;
		uint32_t ta;
		unsigned long long tb;
		block_exscan_2 (a, b, sa, sb, ta, tb);
		if (i < nblk)
		{
			blk_cnt[i] = a + run_a;
			blk_area[i] = b + run_b;
		}
		run_a += ta;
		run_b += tb;
	}
	if (threadIdx.x == 0)
	{
		info->n_roi = run_a;
		info->n_fg = run_b;
	}
}

__global__ void __launch_bounds__(CB) k_compact_write (const LabelStat* tab, uint32_t cap, SlideInfo* info,
	const uint32_t* blk_cnt, const unsigned long long* blk_area, RoiRec* rois, uint32_t* class_slot, const uint2* ztab = nullptr)
{
	__shared__ uint32_t sa[33];
	__shared__ unsigned long long sb[33];
	const uint32_t i = blockIdx.x * CB + threadIdx.x;
	uint32_t a = 0;
	unsigned long long b = 0;
	LabelStat s;
	s.area = 0;
	if (i < cap && i <= info->max_label && i > 0)
	{
		s = tab[i];
		a = s.area ? 1 : 0;
		b = s.area;
	}
	uint32_t ta;
	unsigned long long tb;
	const uint32_t present = a;
	block_exscan_2 (a, b, sa, sb, ta, tb);
	if (present)
	{
		const uint32_t idx = blk_cnt[blockIdx.x] + a;
		RoiRec r;
		r.label = i;
		r.area = s.area;
		r.imin = s.imin;
		r.imax = s.imax;
		r.x0 = (int32_t)s.xmin;
		r.y0 = (int32_t)s.ymin;
		r.w = (int32_t)(s.xmax - s.xmin + 1);
		r.h = (int32_t)(s.ymax - s.ymin + 1);
		r.cloud_off = blk_area[blockIdx.x] + b;
		r.z0 = 0;
		r.d = 1;
		if (ztab)
		{
			const uint2 zz = ztab[i];
			r.z0 = (int32_t)zz.x;
			r.d = (int32_t)(zz.y - zz.x + 1);
		}
		rois[idx] = r;
		const unsigned long long ncell64 = (unsigned long long)r.w * (unsigned long long)r.h * (unsigned long long)r.d;
		const uint32_t ncell = ncell64 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)ncell64;
		const int tc = tile_class (ncell), ac = area_class (r.area);
		class_slot[idx * 2] = ((uint32_t)tc << 28) | atomicAdd (&info->class_count[tc], 1u);
		class_slot[idx * 2 + 1] = ((uint32_t)ac << 28) | atomicAdd (&info->aclass_count[ac], 1u);
		atomicMax (&info->class_maxcells[tc], ncell);
		atomicMax (&info->class_maxwh[tc], (uint32_t)max (max (r.w, r.h), r.d));
		atomicMax (&info->aclass_maxarea[ac], r.area);
		atomicMax (&info->max_inten, r.imax);
		atomicMax (&info->inv_min_inten, ~r.imin);
	}
}

// class lists: list[class_base[c] + slot] = roi index
__global__ void k_fill_class_lists (const uint32_t* class_slot, uint32_t n, const uint32_t* tbase, const uint32_t* abase,
	uint32_t* tlist, uint32_t* alist)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	uint32_t t = class_slot[i * 2], a = class_slot[i * 2 + 1];
	tlist[tbase[t >> 28] + (t & 0x0FFFFFFFu)] = i;
	alist[abase[a >> 28] + (a & 0x0FFFFFFFu)] = i;
}

// ------------------------------------------------------------------------------------------------
// CSR packing: the ROI's pixels in raster order, contiguous per ROI (reference: scanTrivialRoisInMemory
// phase2_2d.cpp:637-684 -- a full-slide re-scan with a binary search per pixel and a 24-byte push_back).
// One warp per ROI walks the ROI's AABB row by row; ballot + popc give every member pixel its slot, so the
// order is deterministic (raster) and the stores are contiguous.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack_cloud (ImageView im, const RoiRec* rois, const uint32_t* list, uint32_t n, uint32_t* cloud)
{
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (warp >= n)
		return;
	const RoiRec r = rois[list ? list[warp] : warp];
	uint32_t* dst = cloud + r.cloud_off;
	uint32_t run = 0;
	for (int yz = 0; yz < r.h * r.d; yz++)	// rows of the AABB, slice by slice (2D: d == 1)
	{
		const int z = yz / r.h, y = yz - z * r.h;
		const size_t rowbase = ((size_t)(r.z0 + z) * im.H + (size_t)(r.y0 + y)) * im.W + r.x0;
		for (int xb = 0; xb < r.w; xb += 32)
		{
			const int x = xb + lane;
			bool in = false;
			uint32_t v = 0;
			if (x < r.w && im.label[rowbase + x] == r.label)
			{
				in = true;
				v = load_inten (im, rowbase + x);
			}
			const unsigned b = __ballot_sync (0xffffffffu, in);
			if (in)
				dst[run + __popc (b & ((1u << lane) - 1))] = v;
			run += __popc (b);
		}
	}
}

// large ROIs: one CTA per ROI, per-row counts -> offsets -> write
__global__ void __launch_bounds__(256) k_pack_cloud_large (ImageView im, const RoiRec* rois, const uint32_t* list, uint32_t n,
	uint32_t* cloud, unsigned long long* rowoff_scratch, size_t rowoff_stride)
{
	// rowoff_scratch: per-CTA array of (h+1) row offsets
	for (uint32_t k = blockIdx.x; k < n; k += gridDim.x)
	{
		const RoiRec r = rois[list[k]];
		unsigned long long* rowoff = rowoff_scratch + (size_t)blockIdx.x * rowoff_stride;
		const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
		const int nrows = r.h * r.d;	// rows of the AABB, slice by slice (2D: d == 1)
		for (int y = warp; y < nrows; y += nw)
		{
			const int zz = y / r.h, yy = y - zz * r.h;
			const size_t rowbase = ((size_t)(r.z0 + zz) * im.H + (size_t)(r.y0 + yy)) * im.W + r.x0;
			uint32_t cnt = 0;
			for (int x = lane; x < r.w; x += 32)
				cnt += im.label[rowbase + x] == r.label ? 1 : 0;
			cnt = __reduce_add_sync (0xffffffffu, cnt);
			if (lane == 0)
				rowoff[y + 1] = cnt;
		}
		__syncthreads();
		if (threadIdx.x == 0)
		{
			rowoff[0] = 0;
			for (int y = 0; y < nrows; y++)
				rowoff[y + 1] += rowoff[y];
		}
		__syncthreads();
		uint32_t* dst = cloud + r.cloud_off;
		for (int y = warp; y < nrows; y += nw)
		{
			const int zz = y / r.h, yy = y - zz * r.h;
			const size_t rowbase = ((size_t)(r.z0 + zz) * im.H + (size_t)(r.y0 + yy)) * im.W + r.x0;
			unsigned long long run = rowoff[y];
			for (int xb = 0; xb < r.w; xb += 32)
			{
				const int x = xb + lane;
				bool in = false;
				uint32_t v = 0;
				if (x < r.w && im.label[rowbase + x] == r.label)
				{
					in = true;
					v = load_inten (im, rowbase + x);
				}
				const unsigned b = __ballot_sync (0xffffffffu, in);
				if (in)
					dst[run + __popc (b & ((1u << lane) - 1))] = v;
				run += __popc (b);
			}
		}
		__syncthreads();
	}
}

// ------------------------------------------------------------------------------------------------
// Per-ROI family kernels: persistent CTAs, one ROI per CTA iteration, tile + matrices in shared memory
// (or in a per-CTA global scratch slot when the matrices of the class do not fit).
// ------------------------------------------------------------------------------------------------
struct DebugOut
{
	uint32_t* u32;		// dense matrix dump / zone keys
	uint32_t* nkeys;
	double* f64;		// NGTDM S
};

constexpr int ROI_THREADS = 128;
constexpr int ROI_WARPS = ROI_THREADS / 32;

// tile load + one family body for one ROI, run by the context c (a CTA or a single warp)
template <int FAM, bool P16>
__device__ __forceinline__ void run_family (const Cta& c, const ImageView& im, const RoiRec& r, const Params& prm, uint32_t* cell,
	uint32_t* mem, int Lmax, uint32_t maxcells, int maxwh, double* o, const DebugOut& dbg)
{
	if (FAM == FAM_NGLDM)
		load_tile_ngldm (c, im, r, prm, cell);
	else
	{
		Binner B;
		B.init (prm.ibsi ? 0 : prm.greydepth, r.imin, r.imax);
		load_tile (c, im, r, B, cell);
	}
	c.sync();
	if (FAM == FAM_GLCM)
	{
		GlcmDebug d { dbg.u32 };
		glcm_body<P16> (c, r, prm, cell, mem, Lmax, o + COL_GLCM, dbg.u32 ? &d : nullptr);
	}
	else if (FAM == FAM_GLRLM)
	{
		GlrlmDebug d { dbg.u32 };
		glrlm_body (c, r, prm, cell, mem, Lmax, maxwh, (int)maxcells, o + COL_GLRLM, dbg.u32 ? &d : nullptr);
	}
	else if (FAM == FAM_GLDZM)
	{
		ZoneDebug d { dbg.u32, dbg.nkeys };
		gldzm_body (c, r, prm, cell, mem, Lmax, gldzm_nd_bound (maxcells, maxwh), maxcells, o + COL_GLDZM, dbg.u32 ? &d : nullptr);
	}
	else if (FAM == FAM_GLSZM)
	{
		ZoneDebug d { dbg.u32, dbg.nkeys };
		glszm_body (c, r, prm, cell, mem, maxcells, glszm_hash_slots (Lmax, maxcells), o + COL_GLSZM, dbg.u32 ? &d : nullptr);
	}
	else if (FAM == FAM_GLDM)
	{
		DenseDebug d { dbg.u32 };
		gldm_body (c, r, prm, cell, mem, Lmax, o + COL_GLDM, dbg.u32 ? &d : nullptr);
	}
	else if (FAM == FAM_NGLDM)
	{
		DenseDebug d { dbg.u32 };
		ngldm_body (c, r, prm, cell, mem, Lmax, o + COL_NGLDM, dbg.u32 ? &d : nullptr);
	}
	else if (FAM == FAM_NGTDM)
	{
		NgtdmDebug d { dbg.u32, dbg.f64 };
		ngtdm_body (c, r, prm, cell, mem, Lmax, o + COL_NGTDM, dbg.u32 ? &d : nullptr);
	}
	c.sync();
}

// CTA per ROI (GLCM / GLRLM: one warp per angle; every family for the larger tile classes)
template <int FAM, bool P16>
__global__ void __launch_bounds__(ROI_THREADS) k_roi_family (ImageView im, const RoiRec* __restrict__ rois,
	const uint32_t* __restrict__ list, uint32_t n, Params prm, int Lmax, uint32_t maxcells, int maxwh,
	double* out, int out_stride, uint32_t* gscratch, size_t gwords, DebugOut dbg)
{
	extern __shared__ __align__(16) uint32_t smem[];
	Cta c { (int)threadIdx.x, (int)blockDim.x, (double*)smem };
	uint32_t* cell = smem + RED_WORDS;
	uint32_t* mem = gscratch ? gscratch + (size_t)blockIdx.x * gwords : cell + ((maxcells + 1) & ~1u);
	// the family region is all zero on entry of every body and every body leaves it all zero again (owner-zeroing):
	// cleared here once per CTA (the global scratch slots are cleared by the host before the launch)
	if (!gscratch)
		for (size_t i = threadIdx.x; i < gwords; i += blockDim.x)
			mem[i] = 0;
	for (uint32_t k = blockIdx.x; k < n; k += gridDim.x)
	{
		const uint32_t ri = list ? list[k] : k;
		const RoiRec r = rois[ri];
		run_family<FAM, P16> (c, im, r, prm, cell, mem, Lmax, maxcells, maxwh, out + (size_t)ri * out_stride, dbg);
	}
}

// WARP per ROI (single-matrix families on small tiles): every warp of the CTA owns a private shared-memory slice
// (tile + family region) and walks its own ROIs -- no block-wide barrier anywhere, reductions are shuffle trees.
template <int FAM>
__global__ void __launch_bounds__(ROI_THREADS) k_roi_family_w (ImageView im, const RoiRec* __restrict__ rois,
	const uint32_t* __restrict__ list, uint32_t n, Params prm, int Lmax, uint32_t maxcells, int maxwh,
	double* out, int out_stride, size_t famw)
{
	extern __shared__ __align__(16) uint32_t smem[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const size_t tilew = (maxcells + 1) & ~1u;
	uint32_t* cell = smem + (size_t)warp * (tilew + famw);
	uint32_t* mem = cell + tilew;
	Cta c { lane, 32, nullptr, 1 };
	for (size_t i = lane; i < famw; i += 32)
		mem[i] = 0;
	__syncwarp();
	const DebugOut nodbg { nullptr, nullptr, nullptr };
	for (uint32_t k = blockIdx.x * ROI_WARPS + warp; k < n; k += gridDim.x * ROI_WARPS)
	{
		const uint32_t ri = list ? list[k] : k;
		const RoiRec r = rois[ri];
		run_family<FAM, false> (c, im, r, prm, cell, mem, Lmax, maxcells, maxwh, out + (size_t)ri * out_stride, nodbg);
	}
}

// ------------------------------------------------------------------------------------------------
// 3D (Nyxus3D) families: CTA per ROI, the ROI's AABB cube in shared memory, family region in shared memory or in a
// per-CTA global slot.  FAM3: 1 GLCM, 2 GLSZM, 3 NGTDM.
// ------------------------------------------------------------------------------------------------
enum Family3 : int { FAM3_INT = 0, FAM3_GLCM, FAM3_GLSZM, FAM3_NGTDM, FAM3_COUNT };
constexpr int NCOL3_INT = NCOL_INT, NCOL3_GLCM = NCOL_GLCM3, NCOL3_NGTDM = NCOL_NGTDM, NCOL3_GLSZM = NCOL_GLSZM;
constexpr int COL3_INT = 0, COL3_GLCM = COL3_INT + NCOL3_INT, COL3_NGTDM = COL3_GLCM + NCOL3_GLCM, COL3_GLSZM = COL3_NGTDM + NCOL3_NGTDM;
constexpr int NCOL3_ALL = COL3_GLSZM + NCOL3_GLSZM;

template <int FAM3, bool P16>
__global__ void __launch_bounds__(ROI_THREADS) k_roi_family3 (ImageView im, const RoiRec* __restrict__ rois,
	const uint32_t* __restrict__ list, uint32_t n, Params prm, int greyinfo, int Lmax, uint32_t maxcells,
	double* out, int out_stride, uint32_t* gscratch, size_t gwords)
{
	extern __shared__ __align__(16) uint32_t smem[];
	Cta c { (int)threadIdx.x, (int)blockDim.x, (double*)smem };
	uint32_t* cell = smem + RED_WORDS;
	uint32_t* mem = gscratch ? gscratch + (size_t)blockIdx.x * gwords : cell + ((maxcells + 1) & ~1u);
	if (!gscratch)
		for (size_t i = threadIdx.x; i < gwords; i += blockDim.x)
			mem[i] = 0;
	for (uint32_t k = blockIdx.x; k < n; k += gridDim.x)
	{
		const uint32_t ri = list ? list[k] : k;
		const RoiRec r = rois[ri];
		double* o = out + (size_t)ri * out_stride;
		Binner B;
		B.init (greyinfo, r.imin, r.imax);
		load_cube (c, im, r, B, cell);
		c.sync();
		if (FAM3 == FAM3_GLCM)
			glcm3d_body<P16> (c, r, prm, cell, mem, Lmax, o + COL3_GLCM);
		else if (FAM3 == FAM3_GLSZM)
			glszm3d_body (c, r, prm, cell, mem, maxcells, glszm_hash_slots (Lmax, maxcells), greyinfo > 0, o + COL3_GLSZM);
		else if (FAM3 == FAM3_NGTDM)
			ngtdm3d_body (c, r, prm, cell, mem, Lmax, greyinfo > 0, o + COL3_NGTDM);
		c.sync();
	}
}

// first-order: reads the ROI's CSR slice
__global__ void __launch_bounds__(ROI_THREADS) k_roi_intensity (const RoiRec* __restrict__ rois, const uint32_t* __restrict__ list,
	uint32_t n, Params prm, const uint32_t* __restrict__ cloud, uint32_t npad_max, double* out, int out_stride,
	uint32_t* gscratch, size_t gwords)
{
	extern __shared__ __align__(16) uint32_t smem[];
	Cta c { (int)threadIdx.x, (int)blockDim.x, (double*)smem };
	uint32_t* mem = gscratch ? gscratch + (size_t)blockIdx.x * gwords : smem + RED_WORDS;
	for (uint32_t k = blockIdx.x; k < n; k += gridDim.x)
	{
		const uint32_t ri = list ? list[k] : k;
		const RoiRec r = rois[ri];
		intensity_body (c, r, prm, cloud + r.cloud_off, r.area, next_pow2 (r.area), mem, out + (size_t)ri * out_stride + COL_INT);
		c.sync();
	}
}

// first-order, warp per ROI (small clouds)
__global__ void __launch_bounds__(ROI_THREADS) k_roi_intensity_w (const RoiRec* __restrict__ rois, const uint32_t* __restrict__ list,
	uint32_t n, Params prm, const uint32_t* __restrict__ cloud, double* out, int out_stride, size_t memw)
{
	extern __shared__ __align__(16) uint32_t smem[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint32_t* mem = smem + (size_t)warp * memw;
	Cta c { lane, 32, nullptr, 1 };
	for (uint32_t k = blockIdx.x * ROI_WARPS + warp; k < n; k += gridDim.x * ROI_WARPS)
	{
		const uint32_t ri = list ? list[k] : k;
		const RoiRec r = rois[ri];
		intensity_body (c, r, prm, cloud + r.cloud_off, r.area, next_pow2 (r.area), mem, out + (size_t)ri * out_stride + COL_INT);
		c.sync();
	}
}

// final gather: canonical table -> [label, t_index, selected columns], values through force_finite_number(v, noval)
// (reference output_2_buffer.cpp:490,577)
__global__ void k_select_columns (const double* __restrict__ full, int full_stride, const RoiRec* __restrict__ rois, uint32_t n,
	const int* __restrict__ sel, int nsel, double noval, double t_index, double* __restrict__ dst)
{
	const size_t total = (size_t)n * (nsel + 2);
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
	{
		const uint32_t r = (uint32_t)(i / (nsel + 2));
		const int cidx = (int)(i - (size_t)r * (nsel + 2));
		double v;
		if (cidx == 0)
			v = (double)(int32_t)rois[r].label;	// the reference keeps labels as int (pixel_feed.cpp:28)
		else if (cidx == 1)
			v = t_index;
		else
			v = finite_or (full[(size_t)r * full_stride + sel[cidx - 2]], noval);
		dst[i] = v;
	}
}

} // namespace nyx
