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
#include "morph_body.h"
#include "ih_body.h"
#include "geomom_body.h"
#include "contour_body.h"

namespace nyx
{

// ------------------------------------------------------------------------------------------------
// Phase 1: one streaming pass over the slide -> per-label {area, min, max, AABB}
// (reference: gatherRoisMetricsInMemory phase1.cpp:373-407 + feed_pixel_2_metrics pixel_feed.cpp:19-43 +
//  init/update_label_record_3 features_calc_workflow.cpp:65-105: two hash probes per foreground pixel.)
// Here: a dense label-indexed table in HBM (32 B per label = one sector), 128-bit loads, and run-aggregated
// updates: the pixels of a row run (and, through a per-lane cache, of the runs of a few consecutive rows) are
// reduced in registers / by warp shuffles and ONE lane issues the 7 atomics of the run.
// ------------------------------------------------------------------------------------------------
struct __align__(32) LabelStat
{
	uint32_t area, imin, imax, xmin, xmax, ymin, ymax, pad;
};

// info block shared between device and host (one D2H copy per slide)
constexpr int NCLASS = 5;		// tile classes by cell count
constexpr int NACLASS = 5;		// cloud classes by area
constexpr int NB_MAX = 16;		// ROI batches per slide (rows of batch b are copied to the host while batch b+1 computes)
constexpr int NTK = NB_MAX * (NCLASS + 1), NAK = NB_MAX * (NACLASS + 1);	// (batch, class) keys
struct SlideInfo
{
	uint32_t max_label;		// largest label value seen by the scan
	uint32_t n_roi;
	unsigned long long n_fg;	// foreground pixels = sum of areas
	uint32_t max_inten;		// largest aux_max over ROIs (IBSI level bound)
	// per (batch, class) key = batch * (NCLASS + 1) + class; the last class of a batch = oversize (giant path)
	uint32_t class_count[NTK];
	uint32_t class_maxcells[NTK];
	uint32_t class_maxwh[NTK];
	uint32_t aclass_count[NAK];
	uint32_t aclass_maxarea[NAK];
	uint32_t inv_min_inten;		// ~(smallest aux_min over ROIs): the slide's foreground minimum (3D COVERED_IMAGE_INTENSITY_RANGE)
	unsigned long long n_tile_cells;	// cells of the binned-tile store (sum of the padded AABBs of the small-class ROIs)
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

// One finished run of equal labels inside a row: the 7 (2D) or 9 (3D) commutative updates of its label record.
__device__ __forceinline__ void stat_commit (LabelStat* tab, uint32_t cap, uint32_t lbl, uint32_t cnt, uint32_t vmin, uint32_t vmax,
	uint32_t xmin, uint32_t xmax, uint32_t ymin, uint32_t ymax, uint2* ztab, uint32_t zmin, uint32_t zmax)
{
	if (lbl >= cap)		// table too small: the host sees info->max_label and retries with a larger one
		return;
	LabelStat* s = tab + lbl;
	atomicAdd (&s->area, cnt);
	atomicMin (&s->imin, vmin);
	atomicMax (&s->imax, vmax);
	atomicMin (&s->xmin, xmin);
	atomicMax (&s->xmax, xmax);
	atomicMin (&s->ymin, ymin);
	atomicMax (&s->ymax, ymax);
	if (ztab)
	{
		atomicMin (&ztab[lbl].x, zmin);
		atomicMax (&ztab[lbl].y, zmax);
	}
}

// (min, max) of a set of intensities.  16-bit intensities: ONE register, (65535 - min) in the high half and max in the
// low half, joined by a single VIMNMX.U16x2; 32-bit intensities: two registers.
template <bool U16> struct MinMax
{
	uint32_t a, b;
	__device__ __forceinline__ static MinMax of (uint32_t v)
	{
		MinMax m;
		if (U16) { m.a = ((v ^ 0xFFFFu) << 16) | v; m.b = 0; }
		else { m.a = v; m.b = v; }
		return m;
	}
	__device__ __forceinline__ void join (const MinMax& o)
	{
		if (U16) a = __vmaxu2 (a, o.a);
		else { a = min (a, o.a); b = max (b, o.b); }
	}
	__device__ __forceinline__ uint32_t vmin() const { return U16 ? (a >> 16) ^ 0xFFFFu : a; }
	__device__ __forceinline__ uint32_t vmax() const { return U16 ? (a & 0xFFFFu) : b; }
	__device__ __forceinline__ MinMax up (int o) const
	{
		MinMax m;
		m.a = __shfl_up_sync (0xffffffffu, a, o);
		m.b = U16 ? 0u : __shfl_up_sync (0xffffffffu, b, o);
		return m;
	}
};

constexpr int SCAN_PX = 8;			// pixels per lane
constexpr int SCAN_CHUNK = 32 * SCAN_PX;	// pixels per warp and row

// The 8 pixels of one lane by scalar loads (images whose width or alignment rules out the 16-byte path):
// labels l[], intensities as MinMax singletons p[]
template <bool U16>
__device__ __forceinline__ void scan_load (const void* __restrict__ inten_, const uint32_t* __restrict__ label, size_t rowbase, uint32_t x0,
	uint32_t W, uint32_t (&l)[SCAN_PX], MinMax<U16> (&p)[SCAN_PX])
{
	const size_t base = rowbase + x0;
#pragma unroll
	for (int k = 0; k < SCAN_PX; k++)
	{
		l[k] = 0;
		p[k] = MinMax<U16>::of (0);
		if (x0 + k < W)
		{
			l[k] = __ldg (label + base + k);
			p[k] = MinMax<U16>::of (U16 ? (uint32_t)__ldg ((const uint16_t*)inten_ + base + k) : __ldg ((const uint32_t*)inten_ + base + k));
		}
	}
}

constexpr int SCAN_DEPTH = 4;			// rows in flight per warp (cp.async ring in shared memory)
template <bool U16> __host__ __device__ constexpr int scan_narr() { return U16 ? 3 : 4; }	// 16-byte pieces per lane and row: 2 label + 1 or 2 intensity
template <bool U16> __host__ __device__ constexpr size_t scan_smem_bytes() { return (size_t)SCAN_DEPTH * scan_narr<U16>() * 256 * 16; }

__device__ __forceinline__ void cp_async16 (void* smem, const void* gmem)
{
	const unsigned sa = (unsigned)__cvta_generic_to_shared (smem);
	asm volatile ("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8 (void* smem, const void* gmem)
{
	const uint32_t sa = (uint32_t)__cvta_generic_to_shared (smem);
	asm volatile ("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4 (void* smem, const void* gmem)
{
	const uint32_t sa = (uint32_t)__cvta_generic_to_shared (smem);
	asm volatile ("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(sa), "l"(gmem) : "memory");
}

// the (unit, row) sequence of one warp: units warp, warp + nwarps, ...; inside a unit the rows top-down
struct ScanCursor
{
	unsigned long long unit;
	uint32_t row, row1, x0;
	bool valid;
	__device__ __forceinline__ void open (unsigned long long nunits, uint32_t nchunk, uint32_t rpu, uint32_t nrows, uint32_t lane)
	{
		valid = unit < nunits;
		if (valid)
		{
			const uint32_t rb = (uint32_t)(unit / nchunk);
			x0 = ((uint32_t)(unit - (unsigned long long)rb * nchunk) * SCAN_CHUNK) + lane * SCAN_PX;
			row = rb * rpu;
			row1 = min (nrows, row + rpu);
		}
	}
	__device__ __forceinline__ void advance (unsigned long long nunits, unsigned long long nwarps, uint32_t nchunk, uint32_t rpu, uint32_t nrows, uint32_t lane)
	{
		if (++row >= row1)
		{
			unit += nwarps;
			open (nunits, nchunk, rpu, nrows, lane);
		}
	}
};

// Work unit = (block of `rpu` consecutive rows, 256-pixel chunk): one warp, 8 consecutive pixels per lane.  The rows are
// STAGED THROUGH SHARED MEMORY by cp.async: every lane copies its own two 16-byte label pieces and one (uint16) or two
// (uint32) intensity pieces of the row SCAN_DEPTH-1 steps ahead into a lane-private ring slot, so a warp keeps 3 rows
// (4.5 KB) in flight without holding registers for them (60 registers, 4 CTAs / 32 warps per SM, ~140 KB in flight).
// Measured on configs[1] (B200): loads alone 0.109 ms (5.5 TB/s), + run analysis 0.155 ms, + L2 atomics 0.178 ms.
// Equal labels are contiguous along a row, so per row the warp does a SEGMENTED max-scan over lanes (shuffles) of the
// packed (min, max) and exactly one lane per run -- the one holding the run's last pixel -- closes it; the run length
// needs no scan (end position minus the start position held by the segment's head lane), y/z are warp-uniform.
// Every run that closes in a lane is first merged into a one-entry per-lane register cache: walking DOWN the rows of
// the unit, the right edge of a ROI usually stays in the same lane for several rows, so a ROI costs one commit (7 L2
// atomics) per few rows instead of one per row (a whole-slide ROI: one per unit).
template <bool U16, bool VEC, bool D3>
__global__ void __launch_bounds__(256, U16 ? 3 : 2) k_scan_stats (const void* __restrict__ inten_, const uint32_t* __restrict__ label,
	uint32_t W, uint32_t nrows, uint32_t H, uint32_t nchunk, uint32_t rpu, LabelStat* tab, uint32_t cap, SlideInfo* info, uint2* ztab)
{
	typedef MinMax<U16> MM;
	constexpr int NARR = scan_narr<U16>();
	extern __shared__ uint4 sring[];	// [SCAN_DEPTH][NARR][256]
	const uint32_t lane = threadIdx.x & 31;
	const unsigned long long nunits = (unsigned long long)((nrows + rpu - 1) / rpu) * nchunk;
	const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
	uint32_t maxlbl = 0;
	// per-lane run cache (c_lbl == 0: empty)
	uint32_t c_lbl = 0, c_cnt = 0, c_xmin = 0, c_xmax = 0, c_ymin = 0, c_ymax = 0, c_zmin = 0, c_zmax = 0;
	MM c_mm = MM::of (0);
	{
		// a finished run goes into the lane's one-entry cache: merged when the entry holds the same label (the edge of a
		// ROI stays in the same lane for several rows), else the entry is committed and replaced
		auto cache_put = [&](uint32_t lbl, uint32_t xs, uint32_t xe, const MM& mm, uint32_t y, uint32_t z)
		{
			const uint32_t cnt = xe - xs + 1;
			maxlbl = max (maxlbl, lbl);
			if (c_lbl == lbl)
			{
				c_cnt += cnt;
				c_mm.join (mm);
				c_xmin = min (c_xmin, xs);
				c_xmax = max (c_xmax, xe);
				c_ymin = min (c_ymin, y);
				c_ymax = max (c_ymax, y);
				c_zmin = min (c_zmin, z);
				c_zmax = max (c_zmax, z);
			}
			else
			{
				if (c_lbl)
					stat_commit (tab, cap, c_lbl, c_cnt, c_mm.vmin(), c_mm.vmax(), c_xmin, c_xmax, c_ymin, c_ymax, D3 ? ztab : nullptr, c_zmin, c_zmax);
				c_lbl = lbl; c_cnt = cnt; c_mm = mm; c_xmin = xs; c_xmax = xe;
				c_ymin = c_ymax = y;
				c_zmin = c_zmax = z;
			}
		};
		// one row of the unit: l[] / p[] = this lane's 8 pixels
		auto process = [&](const uint32_t (&l)[SCAN_PX], const MM (&p)[SCAN_PX], uint32_t row, uint32_t x0)
		{
			uint32_t any = l[0];
#pragma unroll
			for (int k = 1; k < SCAN_PX; k++)
				any |= l[k];
			if (__ballot_sync (0xffffffffu, any != 0) != 0)		// else: background chunk
			{
				uint32_t y = row, z = 0;
				if (D3)
				{
					z = row / H;
					y = row - z * H;
				}
				// ---- runs inside the lane: L = pixels [0, a) (closes here when b > 0), R = pixels [b, 8) (open to the right),
				// runs strictly between them are committed on the spot (they need nothing from other lanes)
				const uint32_t prev = __shfl_up_sync (0xffffffffu, l[SCAN_PX - 1], 1);
				const bool s0 = lane == 0 || l[0] != prev;	// pixel 0 starts a run
				// boundary mask: bit k-1 set when pixel k starts a new run.  Branch-free: L = join of the pixels before the first
				// boundary, R = join of the pixels from the last boundary on (the whole lane when there is none)
				uint32_t m = 0;
#pragma unroll
				for (int k = 1; k < SCAN_PX; k++)
					m |= (l[k] != l[k - 1] ? 1u : 0u) << (k - 1);
				const int a = m ? __ffs (m) : SCAN_PX, b = 32 - __clz (m);
				MM Lmm = p[0], cur = p[SCAN_PX - 1];
#pragma unroll
				for (int k = 1; k < SCAN_PX; k++)
				{
					if ((m & ((1u << k) - 1u)) == 0)		// no boundary up to pixel k: pixel k belongs to L
						Lmm.join (p[k]);
					if ((m >> (SCAN_PX - 1 - k)) == 0)		// no boundary after pixel 7-k: pixel 7-k belongs to R
						cur.join (p[SCAN_PX - 1 - k]);
				}
				if (__popc (m) >= 2)	// rare: runs strictly inside the lane
				{
					int st = 0;
					MM in = p[0];
#pragma unroll
					for (int k = 1; k < SCAN_PX; k++)
					{
						if (l[k] != l[k - 1])
						{
							if (st != 0 && l[k - 1] != 0)
							{
								maxlbl = max (maxlbl, l[k - 1]);
								stat_commit (tab, cap, l[k - 1], (uint32_t)(k - st), in.vmin(), in.vmax(), x0 + st, x0 + k - 1, y, y, D3 ? ztab : nullptr, z, z);
							}
							st = k;
							in = p[k];
						}
						else
							in.join (p[k]);
					}
				}
				// ---- inclusive segmented scan of R over the lanes; a segment starts where R starts inside the lane
				const bool head = b > 0 || s0;
				const unsigned heads = __ballot_sync (0xffffffffu, head);
				const int headlane = 31 - __clz (heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
				for (int o = 1; o < 32; o <<= 1)
				{
					const MM t = cur.up (o);
					if ((int)lane - o >= headlane)
						cur.join (t);
				}
				const uint32_t sx = __shfl_sync (0xffffffffu, x0 + (uint32_t)b, headlane);	// first pixel of my R segment
				// what flows in from the lanes to the left (meaningful when !s0)
				const MM kin = cur.up (1);
				const uint32_t ksx = __shfl_up_sync (0xffffffffu, sx, 1);
				const bool next_s0 = __shfl_down_sync (0xffffffffu, s0 ? 1 : 0, 1) != 0 || lane == 31;

				// ---- L closes inside this lane (b > 0): commit it, joined with the inflow
				if (b > 0 && l[0] != 0)
				{
					uint32_t xs = x0;
					if (!s0)
					{
						Lmm.join (kin);
						xs = ksx;
					}
					cache_put (l[0], xs, x0 + a - 1, Lmm, y, z);
				}
				// ---- R ends with this lane
				if (next_s0 && l[SCAN_PX - 1] != 0)
					cache_put (l[SCAN_PX - 1], sx, x0 + SCAN_PX - 1, cur, y, z);
			}
		};
		ScanCursor cu, pf;
		cu.unit = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
		cu.open (nunits, nchunk, rpu, nrows, lane);
		pf = cu;
		// the lane's pieces of the row under cursor c -> ring slot `stage` (one commit group per step, possibly empty)
		auto issue = [&](const ScanCursor& c, int stage)
		{
			if (VEC && c.valid && c.x0 < W)
			{
				const size_t base = (size_t)c.row * W + c.x0;
				uint4* slot = sring + (size_t)stage * NARR * 256 + threadIdx.x;
				cp_async16 (slot, label + base);
				cp_async16 (slot + 256, label + base + 4);
				if (U16)
					cp_async16 (slot + 512, (const uint16_t*)inten_ + base);
				else
				{
					cp_async16 (slot + 512, (const uint32_t*)inten_ + base);
					cp_async16 (slot + 768, (const uint32_t*)inten_ + base + 4);
				}
			}
			asm volatile ("cp.async.commit_group;" ::: "memory");
		};
#pragma unroll
		for (int i = 0; i < SCAN_DEPTH - 1; i++)
		{
			issue (pf, i);
			if (pf.valid)
				pf.advance (nunits, nwarps, nchunk, rpu, nrows, lane);
		}
		int stage = 0;
		while (cu.valid)
		{
			issue (pf, (stage + SCAN_DEPTH - 1) % SCAN_DEPTH);
			if (pf.valid)
				pf.advance (nunits, nwarps, nchunk, rpu, nrows, lane);
			asm volatile ("cp.async.wait_group %0;" :: "n"(SCAN_DEPTH - 1) : "memory");
			uint32_t l[SCAN_PX];
			MM p[SCAN_PX];
			if (VEC)
			{
#pragma unroll
				for (int k = 0; k < SCAN_PX; k++)
				{
					l[k] = 0;
					p[k] = MM::of (0);
				}
				if (cu.x0 < W)
				{
					const uint4* slot = sring + (size_t)stage * NARR * 256 + threadIdx.x;
					const uint4 l0 = slot[0], l1 = slot[256];
					l[0] = l0.x; l[1] = l0.y; l[2] = l0.z; l[3] = l0.w; l[4] = l1.x; l[5] = l1.y; l[6] = l1.z; l[7] = l1.w;
					if (U16)
					{
						const uint4 iv = slot[512];
						const uint32_t w4[4] = { iv.x, iv.y, iv.z, iv.w };
#pragma unroll
						for (int j = 0; j < 4; j++)
						{
							// two pixels per word -> ((65535 - v) << 16) | v of each half with two byte permutes
							p[2 * j].a = __byte_perm (w4[j], ~w4[j], 0x5410);
							p[2 * j + 1].a = __byte_perm (w4[j], ~w4[j], 0x7632);
						}
					}
					else
					{
						const uint4 i0 = slot[512], i1 = slot[768];
						p[0] = MM::of (i0.x); p[1] = MM::of (i0.y); p[2] = MM::of (i0.z); p[3] = MM::of (i0.w);
						p[4] = MM::of (i1.x); p[5] = MM::of (i1.y); p[6] = MM::of (i1.z); p[7] = MM::of (i1.w);
					}
				}
			}
			else
				scan_load<U16> (inten_, label, (size_t)cu.row * W, cu.x0, W, l, p);	// odd widths / unaligned images: direct scalar loads
			process (l, p, cu.row, cu.x0);
			cu.advance (nunits, nwarps, nchunk, rpu, nrows, lane);
			stage = (stage + 1) % SCAN_DEPTH;
		}
	}
	if (c_lbl)
		stat_commit (tab, cap, c_lbl, c_cnt, c_mm.vmin(), c_mm.vmax(), c_xmin, c_xmax, c_ymin, c_ymax, D3 ? ztab : nullptr, c_zmin, c_zmax);
	maxlbl = __reduce_max_sync (0xffffffffu, maxlbl);
	if (lane == 0 && maxlbl)
		atomicMax (&info->max_label, maxlbl);
}

// ------------------------------------------------------------------------------------------------
// ROI compaction: dense label table -> RoiRec array in ascending label order (reference: rows sorted by label,
// output_2_buffer.cpp:305-306) + exclusive scan of the areas (CSR offsets) + size classification.
// Three small kernels over the table (count per 1024-label block, scan of block totals, scatter).
// ------------------------------------------------------------------------------------------------
constexpr int CB = 1024;

// cells of the ROI's slot in the binned-tile store: the AABB padded to 8 cells (one 128-bit load of 16-bit codes);
// ROIs beyond the shared-memory classes (giant path) and 3D ROIs keep no stored tile
__host__ __device__ inline uint32_t tile_store_cells (uint32_t w, uint32_t h, bool d3)
{
	if (d3)
		return 0;
	const unsigned long long nc = (unsigned long long)w * h;
	if (nc > 40000ull)
		return 0;
	return ((uint32_t)nc + 7u) & ~7u;
}

__device__ __forceinline__ void block_exscan_3 (uint32_t& a, unsigned long long& b, unsigned long long& cc, uint32_t* sa,
	unsigned long long* sb, unsigned long long* sc, uint32_t& tot_a, unsigned long long& tot_b, unsigned long long& tot_c)
{
	// exclusive scan over the block of (a, b, cc); sa/sb/sc: 33-entry shared scratch each
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t ia = a;
	unsigned long long ib = b, ic = cc;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		uint32_t ta = __shfl_up_sync (0xffffffffu, ia, o);
		unsigned long long tb = __shfl_up_sync (0xffffffffu, ib, o);
		unsigned long long tc = __shfl_up_sync (0xffffffffu, ic, o);
		if (lane >= o)
		{
			ia += ta;
			ib += tb;
			ic += tc;
		}
	}
	if (lane == 31)
	{
		sa[warp] = ia;
		sb[warp] = ib;
		sc[warp] = ic;
	}
	__syncthreads();
	if (warp == 0)
	{
		uint32_t wa = lane < (blockDim.x >> 5) ? sa[lane] : 0;
		unsigned long long wb = lane < (blockDim.x >> 5) ? sb[lane] : 0;
		unsigned long long wc = lane < (blockDim.x >> 5) ? sc[lane] : 0;
		uint32_t xa = wa;
		unsigned long long xb = wb, xc = wc;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			uint32_t ta = __shfl_up_sync (0xffffffffu, xa, o);
			unsigned long long tb = __shfl_up_sync (0xffffffffu, xb, o);
			unsigned long long tc = __shfl_up_sync (0xffffffffu, xc, o);
			if (lane >= o)
			{
				xa += ta;
				xb += tb;
				xc += tc;
			}
		}
		sa[lane] = xa - wa;	// exclusive warp offsets
		sb[lane] = xb - wb;
		sc[lane] = xc - wc;
		if (lane == 31)
		{
			sa[32] = xa;
			sb[32] = xb;
			sc[32] = xc;
		}
	}
	__syncthreads();
	a = sa[warp] + ia - a;
	b = sb[warp] + ib - b;
	cc = sc[warp] + ic - cc;
	tot_a = sa[32];
	tot_b = sb[32];
	tot_c = sc[32];
	__syncthreads();
}

__global__ void __launch_bounds__(CB) k_compact_count (const LabelStat* tab, uint32_t cap, const SlideInfo* info,
	uint32_t* blk_cnt, unsigned long long* blk_area, unsigned long long* blk_cells, int d3)
{
	__shared__ uint32_t sa[33];
	__shared__ unsigned long long sb[33], sc[33];
	const uint32_t i = blockIdx.x * CB + threadIdx.x;
	uint32_t a = 0;
	unsigned long long b = 0, cc = 0;
	if (i < cap && i <= info->max_label && i > 0)
	{
		const LabelStat s = tab[i];
		a = s.area ? 1 : 0;
		b = s.area;
		if (s.area)
			cc = tile_store_cells (s.xmax - s.xmin + 1, s.ymax - s.ymin + 1, d3 != 0);
	}
	uint32_t ta;
	unsigned long long tb, tc;
	block_exscan_3 (a, b, cc, sa, sb, sc, ta, tb, tc);
	if (threadIdx.x == 0)
	{
		blk_cnt[blockIdx.x] = ta;
		blk_area[blockIdx.x] = tb;
		blk_cells[blockIdx.x] = tc;
	}
}

__global__ void __launch_bounds__(CB) k_compact_scan (uint32_t* blk_cnt, unsigned long long* blk_area, unsigned long long* blk_cells,
	uint32_t nblk, SlideInfo* info)
{
	__shared__ uint32_t sa[33];
	__shared__ unsigned long long sb[33], sc[33];
	uint32_t run_a = 0;
	unsigned long long run_b = 0, run_c = 0;
	for (uint32_t base = 0; base < nblk; base += CB)
	{
		uint32_t i = base + threadIdx.x;
		uint32_t a = i < nblk ? blk_cnt[i] : 0;
		unsigned long long b = i < nblk ? blk_area[i] : 0;
		unsigned long long cc = i < nblk ? blk_cells[i] : 0;
		uint32_t ta;
		unsigned long long tb, tc;
		block_exscan_3 (a, b, cc, sa, sb, sc, ta, tb, tc);
		if (i < nblk)
		{
			blk_cnt[i] = a + run_a;
			blk_area[i] = b + run_b;
			blk_cells[i] = cc + run_c;
		}
		run_a += ta;
		run_b += tb;
		run_c += tc;
	}
	if (threadIdx.x == 0)
	{
		info->n_roi = run_a;
		info->n_fg = run_b;
		info->n_tile_cells = run_c;
	}
}

__global__ void __launch_bounds__(CB) k_compact_write (const LabelStat* tab, uint32_t cap, SlideInfo* info,
	const uint32_t* blk_cnt, const unsigned long long* blk_area, const unsigned long long* blk_cells, RoiRec* rois,
	uint32_t* class_slot, uint32_t batch_rois, const uint2* ztab = nullptr, uint32_t tail_start = 0xFFFFFFFFu, uint32_t tail_q = 1)
{
	__shared__ uint32_t sa[33];
	__shared__ unsigned long long sb[33], sc[33];
	// per-block class tallies: slots are handed out inside the block first, ONE global atomic per (block, class) after
	__shared__ uint32_t s_cnt[NTK], s_acnt[NAK], s_base[NTK], s_abase[NAK];
	__shared__ uint32_t s_maxcells[NTK], s_maxwh[NTK], s_maxarea[NAK], s_maxinten, s_invmin;
	if (threadIdx.x < NTK)
	{
		s_cnt[threadIdx.x] = 0;
		s_maxcells[threadIdx.x] = 0;
		s_maxwh[threadIdx.x] = 0;
	}
	if (threadIdx.x < NAK)
	{
		s_acnt[threadIdx.x] = 0;
		s_maxarea[threadIdx.x] = 0;
	}
	if (threadIdx.x == 0)
	{
		s_maxinten = 0;
		s_invmin = 0;
	}
	const uint32_t i = blockIdx.x * CB + threadIdx.x;
	uint32_t a = 0;
	unsigned long long b = 0, cc = 0;
	LabelStat s;
	s.area = 0;
	if (i < cap && i <= info->max_label && i > 0)
	{
		s = tab[i];
		a = s.area ? 1 : 0;
		b = s.area;
		if (s.area)
			cc = tile_store_cells (s.xmax - s.xmin + 1, s.ymax - s.ymin + 1, ztab != nullptr);
	}
	uint32_t ta;
	unsigned long long tb, tc;
	const uint32_t present = a;
	block_exscan_3 (a, b, cc, sa, sb, sc, ta, tb, tc);		// (contains the barriers that publish the zeroed tallies)
	if (ta == 0)
		return;
	RoiRec r;
	uint32_t idx = 0, tslot = 0, aslot = 0;
	int tc_ = 0, ac = 0;
	if (present)
	{
		idx = blk_cnt[blockIdx.x] + a;
		r.label = i;
		r.area = s.area;
		r.imin = s.imin;
		r.imax = s.imax;
		r.x0 = (int32_t)s.xmin;
		r.y0 = (int32_t)s.ymin;
		r.w = (int32_t)(s.xmax - s.xmin + 1);
		r.h = (int32_t)(s.ymax - s.ymin + 1);
		r.cloud_off = blk_area[blockIdx.x] + b;
		r.tile_off = blk_cells[blockIdx.x] + cc;
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
		// uniform batches of batch_rois ROIs; from tail_start on, the last batch is cut into 2q | q | rest (the table copy of the
		// very last batch is the only one no later batch hides: see process_slide_body)
		int batch = (int)(idx / batch_rois);
		if (idx >= tail_start)
		{
			const uint32_t off = idx - tail_start;
			batch = (int)(tail_start / batch_rois) + (off < 2 * tail_q ? 0 : off < 3 * tail_q ? 1 : 2);
		}
		tc_ = batch * (NCLASS + 1) + tile_class (ncell);
		ac = batch * (NACLASS + 1) + area_class (r.area);
		tslot = atomicAdd (&s_cnt[tc_], 1u);
		aslot = atomicAdd (&s_acnt[ac], 1u);
		atomicMax (&s_maxcells[tc_], ncell);
		atomicMax (&s_maxwh[tc_], (uint32_t)max (max (r.w, r.h), r.d));
		atomicMax (&s_maxarea[ac], r.area);
		atomicMax (&s_maxinten, r.imax);
		atomicMax (&s_invmin, ~r.imin);
	}
	__syncthreads();
	if (threadIdx.x < NTK && s_cnt[threadIdx.x])
	{
		s_base[threadIdx.x] = atomicAdd (&info->class_count[threadIdx.x], s_cnt[threadIdx.x]);
		atomicMax (&info->class_maxcells[threadIdx.x], s_maxcells[threadIdx.x]);
		atomicMax (&info->class_maxwh[threadIdx.x], s_maxwh[threadIdx.x]);
	}
	if (threadIdx.x >= 64 && threadIdx.x - 64 < NAK && s_acnt[threadIdx.x - 64])
	{
		s_abase[threadIdx.x - 64] = atomicAdd (&info->aclass_count[threadIdx.x - 64], s_acnt[threadIdx.x - 64]);
		atomicMax (&info->aclass_maxarea[threadIdx.x - 64], s_maxarea[threadIdx.x - 64]);
	}
	if (threadIdx.x == 128)
	{
		atomicMax (&info->max_inten, s_maxinten);
		atomicMax (&info->inv_min_inten, s_invmin);
	}
	__syncthreads();
	if (present)
	{
		reinterpret_cast<uint4*> (class_slot)[idx] = make_uint4 ((uint32_t)tc_, s_base[tc_] + tslot, (uint32_t)ac, s_abase[ac] + aslot);
	}
}

// class lists: list[class_base[c] + slot] = roi index
__global__ void k_fill_class_lists (const uint32_t* class_slot, uint32_t n, const uint32_t* tbase, const uint32_t* abase,
	uint32_t* tlist, uint32_t* alist)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const uint4 k = reinterpret_cast<const uint4*> (class_slot)[i];	// (tile key, slot, area key, slot)
	tlist[tbase[k.x] + k.y] = i;
	alist[abase[k.z] + k.w] = i;
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

// Binning + packing in one pass (K2): one warp per ROI walks the ROI's AABB once in raster order and writes
//   * the CSR cloud slice (ballot/popc compaction, as k_pack_cloud),
//   * the ROI's slot of the binned-tile store: 16-bit codes (level + NZ/IN flags) of EVERY AABB cell, contiguous,
//     so the seven texture kernels load a coalesced pre-binned tile instead of re-gathering label + intensity rows and
//     re-binning them (reference: bin_intensities is redone per family and per angle, texture_feature.h:23-167).
// tileB: NGLDM's own grey scale (to_grayscale over [0, aux_max], ngldm.cpp:42,126).
__global__ void __launch_bounds__(256) k_pack_tiles (ImageView im, const RoiRec* __restrict__ rois, const uint32_t* __restrict__ list,
	uint32_t n, Params prm, uint32_t* __restrict__ cloud, uint16_t* __restrict__ tileA, uint16_t* __restrict__ tileB)
{
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (warp >= n)
		return;
	const RoiRec r = rois[list ? list[warp] : warp];
	const uint32_t ncell = (uint32_t)r.w * (uint32_t)r.h, npad = (ncell + 7u) & ~7u;
	Binner B;
	B.init (prm.ibsi ? 0 : prm.greydepth, r.imin, r.imax);
	FastDiv fw;
	fw.init ((uint32_t)r.w);
	uint32_t* dst = cloud ? cloud + r.cloud_off : nullptr;
	uint16_t* ta = tileA ? tileA + r.tile_off : nullptr;
	uint16_t* tb = tileB ? tileB + r.tile_off : nullptr;
	const size_t origin = (size_t)r.y0 * im.W + (size_t)r.x0;
	uint32_t run = 0;
	const bool gs_fast = !prm.ibsi && r.imax != 0 && prm.greydepth > 0 && prm.greydepth < 65536;
	const double inv_imax = gs_fast ? 1.0 / (double)r.imax : 0.0;
	constexpr int UN = 4;	// 32-cell chunks whose (independent) label and intensity loads are issued together
	for (uint32_t base0 = 0; base0 < npad; base0 += 32 * UN)
	{
		uint32_t lb[UN], vv[UN];
#pragma unroll
		for (int u = 0; u < UN; u++)
		{
			const uint32_t i = base0 + u * 32 + lane;
			lb[u] = 0;
			vv[u] = 0;
			if (i < ncell)
			{
				const uint32_t y = fw.div (i), x = i - y * (uint32_t)r.w;
				const size_t g = origin + (size_t)y * im.W + x;
				lb[u] = im.label[g];
				vv[u] = load_inten (im, g);
			}
		}
#pragma unroll
		for (int u = 0; u < UN; u++)
		{
			const uint32_t i = base0 + u * 32 + lane;
			if (base0 + u * 32 >= npad)
				break;
			const bool in = i < ncell && lb[u] == r.label;
			const uint32_t v = in ? vv[u] : 0u;
			if (i < npad)
			{
				if (ta)
				{
					uint32_t code = 0;
					if (i < ncell)
						code = (B.bin (v) & TILE16_LEVEL) | (v ? 0x8000u : 0u) | (in ? 0x4000u : 0u);
					ta[i] = (uint16_t)code;
				}
				if (tb)
				{
					// NGLDM's grey scale: to_grayscale over [0, aux_max] (exact, the division only where it could matter)
					const uint32_t lvb = gs_fast ? to_grayscale_fast (v, r.imax, (uint32_t)prm.greydepth, inv_imax)
						: to_grayscale (v, 0, r.imax, (uint32_t)prm.greydepth, prm.ibsi != 0);
					tb[i] = in ? (uint16_t)((lvb & TILE16_LEVEL) | 0x4000u) : (uint16_t)0;
				}
			}
			if (dst)
			{
				const unsigned b = __ballot_sync (0xffffffffu, in);
				if (in)
					dst[run + __popc (b & ((1u << lane) - 1))] = v;
				run += __popc (b);
			}
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
// CTA-per-ROI kernels whose shared-memory footprint leaves room for at most three CTAs on an SM (3D cubes, large first-order
// clouds) run with twice the threads: the same shared memory then feeds 8 instead of 4 warps (NW_MAX)
constexpr int ROI_THREADS_BIG = 256;
constexpr size_t ROI_BIG_SMEM_BYTES = 56 * 1024;

// per-LAUNCH constant of a family that is a function of the class bounds only (GLSZM: hash slots -- a loop with 64-bit
// divisions; GLDZM: distance bound): computed once per thread at kernel start, not once per ROI (it was 7 % of the
// instructions of the GLSZM kernel)
template <int FAM> __device__ __forceinline__ uint32_t family_aux (int Lmax, uint32_t maxcells, int maxwh)
{
	return FAM == FAM_GLSZM ? glszm_hash_slots (Lmax, maxcells) : FAM == FAM_GLDZM ? (uint32_t)gldzm_nd_bound (maxcells, maxwh) : 0u;
}

// tile load + one family body for one ROI, run by the context c (a CTA or a single warp)
template <int FAM, bool P16>
__device__ __forceinline__ void run_family (const Cta& c, const ImageView& im, const RoiRec& r, const Params& prm, uint32_t* cell,
	uint32_t* mem, int Lmax, uint32_t maxcells, int maxwh, double* o, const DebugOut& dbg, uint32_t aux, bool preloaded = false)
{
	if (preloaded)
		;	// the caller has already put the tile into cell[] (the pipelined CTA-per-ROI loop)
	else if (FAM == FAM_NGLDM ? im.tileB != nullptr : im.tileA != nullptr)
		load_tile_store (c, (FAM == FAM_NGLDM ? im.tileB : im.tileA) + r.tile_off, r.w * r.h, cell);
	else if (FAM == FAM_NGLDM)
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
		glcm_body<P16> (c, r, prm, cell, mem, Lmax, o + COL_GLCM, dbg.u32 ? &d : nullptr, maxcells);
	}
	else if (FAM == FAM_GLRLM)
	{
		GlrlmDebug d { dbg.u32 };
		glrlm_body (c, r, prm, cell, mem, Lmax, maxwh, (int)maxcells, o + COL_GLRLM, dbg.u32 ? &d : nullptr);
	}
	else if (FAM == FAM_GLDZM)
	{
		ZoneDebug d { dbg.u32, dbg.nkeys };
		gldzm_body (c, r, prm, cell, mem, Lmax, (int)aux, maxcells, o + COL_GLDZM, dbg.u32 ? &d : nullptr);
	}
	else if (FAM == FAM_GLSZM)
	{
		ZoneDebug d { dbg.u32, dbg.nkeys };
		glszm_body (c, r, prm, cell, mem, maxcells, aux, o + COL_GLSZM, dbg.u32 ? &d : nullptr, Lmax <= GLSZM_LMARG_MAX ? Lmax : 0);
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
// GSCR: the family region lives in a per-CTA global-memory slot (classes whose matrices exceed shared memory).  A
// template parameter, not a run-time select, so that in the common case the compiler sees a pure shared-memory pointer
// and emits LDS / STS / ATOMS with 32-bit addressing instead of generic LD / ATOM with 64-bit address arithmetic.
template <int FAM, bool P16, bool GSCR>
__global__ void __launch_bounds__((FAM == FAM_GLCM || FAM == FAM_GLRLM) ? ROI_THREADS : ROI_THREADS_BIG, FAM == FAM_GLCM ? 4 : FAM == FAM_GLRLM ? 6 : 1) k_roi_family (ImageView im, const RoiRec* __restrict__ rois,
	const uint32_t* __restrict__ list, uint32_t n, Params prm, int Lmax, uint32_t maxcells, int maxwh,
	double* out, int out_stride, uint32_t* gscratch, size_t gwords, DebugOut dbg)
{
	extern __shared__ __align__(16) uint32_t smem[];
	Cta c { (int)threadIdx.x, (int)blockDim.x, (double*)smem };
	const uint32_t aux = family_aux<FAM> (Lmax, maxcells, maxwh);
	uint32_t* cell = smem + RED_WORDS;
	uint32_t* mem = GSCR ? gscratch + (size_t)blockIdx.x * gwords : cell + tile_words (maxcells);
	// the family region is all zero on entry of every body and every body leaves it all zero again (owner-zeroing):
	// cleared here once per CTA (the global scratch slots are cleared by the host before the launch)
	if (!GSCR)
		for (size_t i = threadIdx.x; i < gwords; i += blockDim.x)
			mem[i] = 0;
	// ---- GLCM / GLRLM on the binned-tile store, tiles of at most 1024 cells: a software pipeline over the CTA's ROIs.  The
	// loop is a chain of three dependent global loads per ROI (list entry -> ROI record -> tile) in front of ~20 us of
	// shared-memory work, and a CTA has nothing else to do while it waits (4 - 6 CTAs per SM): ~10 % of the GLCM kernel's
	// stall samples sat on these loads.  Here every stage of the chain is a cp.async issued one ROI earlier than its consumer
	// needs it: while ROI i is computed, the 16-bit codes of tile i+1 land in the upper part of the reduction scratch (which
	// the angle warps of these two families never use), the record of ROI i+2 in a two-slot ring, the list entry of ROI i+3
	// in a four-slot ring.  One wait + barrier per ROI; no registers are held across the body.
	if ((FAM == FAM_GLCM || FAM == FAM_GLRLM) && im.tileA != nullptr && maxcells <= 1024)
	{
		static_assert (sizeof (RoiRec) == 56 && 4 + 2 * sizeof (RoiRec) / 4 <= PF_RING_WORDS, "ring sized for two 56-byte records");
		uint16_t* stage = (uint16_t*)(smem + PF_STAGE_OFF);				// [<= 1024] codes of the next tile
		uint32_t* ring = cell + tile_words (maxcells) + (GSCR ? 0 : gwords);		// PF_RING_WORDS after the family region
		uint32_t* idxr = ring;								// [4] ROI index of iteration it .. it + 3
		RoiRec* recr = (RoiRec*)(ring + 4);						// [2] record of iteration it, it + 1
		const uint32_t k0 = blockIdx.x, g = gridDim.x, tid = threadIdx.x;
		auto issue_tile = [&] (const RoiRec& rr) {
			const uint32_t chunks = (uint32_t)((rr.w * rr.h + 7) >> 3);		// 8 codes = 16 bytes (slots are padded to 8 cells)
			const uint16_t* src = im.tileA + rr.tile_off;
			for (uint32_t q = tid; q < chunks; q += blockDim.x)
				cp_async16 (stage + q * 8, src + q * 8);
		};
		// prologue (once per CTA, latencies exposed): list entries of the first three ROIs, records of the first two, tile of the first
		if (tid < 3 && k0 + tid * g < n)
			idxr[tid] = list ? list[k0 + tid * g] : k0 + tid * g;
		__syncthreads();
		if (tid < 2 && k0 + tid * g < n)
			recr[tid] = rois[idxr[tid]];
		__syncthreads();
		if (k0 < n)
			issue_tile (recr[0]);
		asm volatile ("cp.async.commit_group;" ::: "memory");
		uint32_t it = 0;
		for (uint32_t k = k0; k < n; k += g, it++)
		{
			asm volatile ("cp.async.wait_all;" ::: "memory");
			__syncthreads();	// tile it, record it + 1, list entry it + 2 have landed (and the previous body is done with cell[])
			const RoiRec r = recr[it & 1];
			const uint32_t ri = idxr[it & 3];
			{
				const uint32_t* s2 = (const uint32_t*)stage;
				const int n2 = (r.w * r.h + 1) >> 1;
				for (int q = tid; q < n2; q += blockDim.x)
				{
					const uint32_t u = s2[q];
					cell[2 * q] = tile16_decode (u & 0xFFFFu);
					cell[2 * q + 1] = tile16_decode (u >> 16);
				}
			}
			__syncthreads();	// cell[] complete; stage, recr[it & 1] and idxr[it & 3] are free again
			if (k + g < n)
				issue_tile (recr[(it + 1) & 1]);
			if (k + 2 * g < n && tid < sizeof (RoiRec) / 8)
				cp_async8 ((char*)&recr[it & 1] + tid * 8, (const char*)&rois[idxr[(it + 2) & 3]] + tid * 8);
			if (k + 3 * g < n && tid == 32)
			{
				if (list)
					cp_async4 (&idxr[(it + 3) & 3], list + k + 3 * g);
				else
					idxr[(it + 3) & 3] = k + 3 * g;
			}
			asm volatile ("cp.async.commit_group;" ::: "memory");
			run_family<FAM, P16> (c, im, r, prm, cell, mem, Lmax, maxcells, maxwh, out + (size_t)ri * out_stride, dbg, aux, true);
		}
		return;
	}
	for (uint32_t k = blockIdx.x; k < n; k += gridDim.x)
	{
		const uint32_t ri = list ? list[k] : k;
		const RoiRec r = rois[ri];
		run_family<FAM, P16> (c, im, r, prm, cell, mem, Lmax, maxcells, maxwh, out + (size_t)ri * out_stride, dbg, aux);
	}
}

// WARP per ROI (single-matrix families on small tiles): every warp of the CTA owns a private shared-memory slice
// (tile + family region) and walks its own ROIs -- no block-wide barrier anywhere, reductions are shuffle trees.
template <int FAM>
__global__ void __launch_bounds__(ROI_THREADS, (FAM == FAM_GLDM || FAM == FAM_NGLDM) ? 8 : FAM == FAM_NGTDM ? 7 : (FAM == FAM_GLDZM || FAM == FAM_GLSZM) ? 5 : 1) k_roi_family_w (ImageView im, const RoiRec* __restrict__ rois,
	const uint32_t* __restrict__ list, uint32_t n, Params prm, int Lmax, uint32_t maxcells, int maxwh,
	double* out, int out_stride, size_t famw)
{
	extern __shared__ __align__(16) uint32_t smem[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const size_t tilew = tile_words (maxcells);
	uint32_t* cell = smem + (size_t)warp * (tilew + famw);
	uint32_t* mem = cell + tilew;
	Cta c { lane, 32, nullptr, 1 };
	const uint32_t aux = family_aux<FAM> (Lmax, maxcells, maxwh);
	for (size_t i = lane; i < famw; i += 32)
		mem[i] = 0;
	__syncwarp();
	const DebugOut nodbg { nullptr, nullptr, nullptr };
	for (uint32_t k = blockIdx.x * ROI_WARPS + warp; k < n; k += gridDim.x * ROI_WARPS)
	{
		const uint32_t ri = list ? list[k] : k;
		const RoiRec r = rois[ri];
		run_family<FAM, false> (c, im, r, prm, cell, mem, Lmax, maxcells, maxwh, out + (size_t)ri * out_stride, nodbg, aux);
	}
}

// ------------------------------------------------------------------------------------------------
// 3D (Nyxus3D) families: CTA per ROI, the ROI's AABB cube in shared memory, family region in shared memory or in a
// per-CTA global slot.  FAM3: 1 GLCM, 2 GLSZM, 3 NGTDM.
// ------------------------------------------------------------------------------------------------
enum Family3 : int { FAM3_INT = 0, FAM3_GLCM, FAM3_GLSZM, FAM3_NGTDM, FAM3_GLDM, FAM3_GLRLM, FAM3_GLDZM, FAM3_NGLDM, FAM3_COUNT };
constexpr int NCOL3_INT = NCOL_INT, NCOL3_GLCM = NCOL_GLCM3, NCOL3_NGTDM = NCOL_NGTDM, NCOL3_GLSZM = NCOL_GLSZM, NCOL3_GLDM = NCOL_GLDM;
constexpr int NCOL3_GLRLM = NCOL_GLRLM3, NCOL3_GLDZM = NCOL_GLDZM, NCOL3_NGLDM = NCOL_NGLDM;
constexpr int COL3_INT = 0, COL3_GLCM = COL3_INT + NCOL3_INT, COL3_NGTDM = COL3_GLCM + NCOL3_GLCM, COL3_GLSZM = COL3_NGTDM + NCOL3_NGTDM;
constexpr int COL3_GLDM = COL3_GLSZM + NCOL3_GLSZM, COL3_GLRLM = COL3_GLDM + NCOL3_GLDM, COL3_GLDZM = COL3_GLRLM + NCOL3_GLRLM;
constexpr int COL3_NGLDM = COL3_GLDZM + NCOL3_GLDZM;
constexpr int NCOL3_ALL = COL3_NGLDM + NCOL3_NGLDM;

// per-family work of one cube (after load_cube); Nd / Nr bounds of the launch's ROI class come from prm.cube_maxdim
template <int FAM3, bool P16, bool BIG>
__device__ __forceinline__ void family3_dispatch (const Cta& c, const RoiRec& r, const Params& prm, uint32_t* cell, uint32_t* mem,
	int greyinfo, int Lmax, unsigned long long maxcells, uint32_t slots, double* o)
{
	if (FAM3 == FAM3_GLCM)
		glcm3d_body<P16> (c, r, prm, cell, mem, Lmax, o + COL3_GLCM, greyinfo);
	else if (FAM3 == FAM3_GLSZM)
	{
		if (BIG)
			glszm3d_body_big (c, r, prm, cell, mem, maxcells, slots, greyinfo > 0, o + COL3_GLSZM);
		else
			glszm3d_body (c, r, prm, cell, mem, (uint32_t)maxcells, slots, greyinfo > 0, o + COL3_GLSZM);
	}
	else if (FAM3 == FAM3_NGTDM)
		ngtdm3d_body (c, r, prm, cell, mem, Lmax, greyinfo > 0, o + COL3_NGTDM);
	else if (FAM3 == FAM3_GLDM)
		gldm3d_body (c, r, prm, cell, mem, Lmax, greyinfo > 0, o + COL3_GLDM);
	else if (FAM3 == FAM3_GLRLM)
		glrlm3d_body (c, r, prm, cell, mem, Lmax, prm.cube_maxdim, o + COL3_GLRLM);
	else if (FAM3 == FAM3_GLDZM)
		gldzm3d_body (c, r, prm, cell, mem, Lmax, gldzm_nd_bound ((uint32_t)maxcells, prm.cube_maxdim), maxcells, o + COL3_GLDZM);
	else if (FAM3 == FAM3_NGLDM)
		ngldm3d_body (c, r, prm, cell, mem, Lmax, o + COL3_NGLDM);
}

template <int FAM3, bool P16, bool GSCR>
__global__ void __launch_bounds__(ROI_THREADS_BIG) k_roi_family3 (ImageView im, const RoiRec* __restrict__ rois,
	const uint32_t* __restrict__ list, uint32_t n, Params prm, int greyinfo, int Lmax, uint32_t maxcells,
	double* out, int out_stride, uint32_t* gscratch, size_t gwords)
{
	extern __shared__ __align__(16) uint32_t smem[];
	Cta c { (int)threadIdx.x, (int)blockDim.x, (double*)smem };
	uint32_t* cell = smem + RED_WORDS;
	uint32_t* mem = GSCR ? gscratch + (size_t)blockIdx.x * gwords : cell + tile_words (maxcells);
	if (!GSCR)
		for (size_t i = threadIdx.x; i < gwords; i += blockDim.x)
			mem[i] = 0;
	const uint32_t slots3 = FAM3 == FAM3_GLSZM ? glszm_hash_slots (Lmax, maxcells) : 0u;	// per launch, not per ROI
	for (uint32_t k = blockIdx.x; k < n; k += gridDim.x)
	{
		const uint32_t ri = list ? list[k] : k;
		const RoiRec r = rois[ri];
		double* o = out + (size_t)ri * out_stride;
		Binner B;
		if (FAM3 == FAM3_NGLDM)
			B.init_tg (greyinfo, r.imax, prm.ibsi != 0);
		else
			B.init (greyinfo, r.imin, r.imax);
		load_cube (c, im, r, B, cell);
		c.sync();
		family3_dispatch<FAM3, P16, false> (c, r, prm, cell, mem, greyinfo, Lmax, maxcells, slots3, o);
		c.sync();
	}
}

// 3D cubes beyond the shared-memory size classes (organ-sized masks, whole-volume ROIs): one CTA per ROI, the cube AND the
// family's work arrays in a per-CTA slot of global memory (`slotwords` words: the cube first, `tilewords` words, then the
// family region, zero on entry and exit).  Throughput is that of one CTA per ROI walking L2 / HBM -- the job of this
// path is to give such ROIs their reference values at all (the reference walks them on one host thread).
template <int FAM3>
__global__ void __launch_bounds__(ROI_THREADS) k_roi_family3_big (ImageView im, const RoiRec* __restrict__ rois,
	const uint32_t* __restrict__ list, uint32_t n, Params prm, int greyinfo, int Lmax, unsigned long long maxcells,
	double* out, int out_stride, uint32_t* gscratch, size_t slotwords, size_t tilewords, uint32_t slots)
{
	extern __shared__ __align__(16) uint32_t smem[];
	Cta c { (int)threadIdx.x, (int)blockDim.x, (double*)smem };
	uint32_t* cell = gscratch + (size_t)blockIdx.x * slotwords;
	uint32_t* mem = cell + tilewords;
	for (uint32_t k = blockIdx.x; k < n; k += gridDim.x)
	{
		const uint32_t ri = list ? list[k] : k;
		const RoiRec r = rois[ri];
		double* o = out + (size_t)ri * out_stride;
		Binner B;
		if (FAM3 == FAM3_NGLDM)
			B.init_tg (greyinfo, r.imax, prm.ibsi != 0);
		else
			B.init (greyinfo, r.imin, r.imax);
		load_cube (c, im, r, B, cell);
		c.sync();
		family3_dispatch<FAM3, false, true> (c, r, prm, cell, mem, greyinfo, Lmax, maxcells, slots, o);
		c.sync();
	}
}

// first-order: reads the ROI's CSR slice
template <bool GSCR>
__global__ void __launch_bounds__(ROI_THREADS_BIG) k_roi_intensity (const RoiRec* __restrict__ rois, const uint32_t* __restrict__ list,
	uint32_t n, Params prm, const uint32_t* __restrict__ cloud, uint32_t npad_max, double* out, int out_stride,
	uint32_t* gscratch, size_t gwords)
{
	extern __shared__ __align__(16) uint32_t smem[];
	Cta c { (int)threadIdx.x, (int)blockDim.x, (double*)smem };
	uint32_t* mem = GSCR ? gscratch + (size_t)blockIdx.x * gwords : smem + RED_WORDS;
	for (uint32_t k = blockIdx.x; k < n; k += gridDim.x)
	{
		const uint32_t ri = list ? list[k] : k;
		const RoiRec r = rois[ri];
		intensity_body (c, r, prm, cloud + r.cloud_off, r.area, next_pow2 (r.area), mem, out + (size_t)ri * out_stride + COL_INT);
		c.sync();
	}
}

// first-order, warp per ROI (small clouds)
__global__ void __launch_bounds__(ROI_THREADS, 6) k_roi_intensity_w (const RoiRec* __restrict__ rois, const uint32_t* __restrict__ list,
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
		intensity_body (c, r, prm, cloud + r.cloud_off, r.area, max (next_pow2 (r.area), 64u), mem, out + (size_t)ri * out_stride + COL_INT);
		c.sync();
	}
}

// unweighted geometric moments: warp per ROI (shared-memory size classes) / CTA per ROI (larger); reads the slide
__global__ void __launch_bounds__(256) k_roi_geom_w (ImageView im, const RoiRec* __restrict__ rois, const uint32_t* __restrict__ list,
	uint32_t n, int variants, double* out, int out_stride)
{
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (warp >= n)
		return;
	const uint32_t ri = list ? list[warp] : warp;
	Cta c { (int)(threadIdx.x & 31), 32, nullptr, 1 };
	geomom_body (c, im, rois[ri], variants, out + (size_t)ri * out_stride + COL_GEOM);
}
__global__ void __launch_bounds__(ROI_THREADS) k_roi_geom (ImageView im, const RoiRec* __restrict__ rois, const uint32_t* __restrict__ list,
	uint32_t n, int variants, double* out, int out_stride)
{
	extern __shared__ __align__(16) uint32_t smem[];
	Cta c { (int)threadIdx.x, (int)blockDim.x, (double*)smem };
	for (uint32_t k = blockIdx.x; k < n; k += gridDim.x)
	{
		const uint32_t ri = list ? list[k] : k;
		geomom_body (c, im, rois[ri], variants, out + (size_t)ri * out_stride + COL_GEOM);
		c.sync();
	}
}

// distance-to-contour weighted geometric moments (contour_body.h): warp per ROI (one launch per size class, every warp owns a
// scratch slot of `slotw` words sized for the class) / CTA per ROI (oversize ROIs); reads the slide
__global__ void __launch_bounds__(ROI_THREADS) k_roi_wmom_w (ImageView im, const RoiRec* __restrict__ rois, const uint32_t* __restrict__ list,
	uint32_t n, int variants, uint32_t* scratch, size_t slotw, uint32_t maxcells, int maxwh, double* out, int out_stride)
{
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
	Cta c { (int)(threadIdx.x & 31), 32, nullptr, 1 };
	uint32_t* mem = scratch + (size_t)warp * slotw;
	for (uint32_t k = warp; k < n; k += nwarps)
	{
		const uint32_t ri = list ? list[k] : k;
		wmom_body (c, im, rois[ri], variants, mem, maxcells, maxwh, out + (size_t)ri * out_stride + COL_WMOM);
		c.sync();
	}
}
__global__ void __launch_bounds__(ROI_THREADS) k_roi_wmom (ImageView im, const RoiRec* __restrict__ rois, const uint32_t* __restrict__ list,
	uint32_t n, int variants, uint32_t* scratch, size_t slotw, unsigned long long maxcells, int maxwh, double* out, int out_stride)
{
	extern __shared__ __align__(16) uint32_t smem[];
	Cta c { (int)threadIdx.x, (int)blockDim.x, (double*)smem };
	uint32_t* mem = scratch + (size_t)blockIdx.x * slotw;
	for (uint32_t k = blockIdx.x; k < n; k += gridDim.x)
	{
		const uint32_t ri = list ? list[k] : k;
		wmom_body (c, im, rois[ri], variants, mem, maxcells, maxwh, out + (size_t)ri * out_stride + COL_WMOM);
		c.sync();
	}
}

// *ALL_IH*: warp per ROI (CSR slices of the shared-memory area classes) / CTA per ROI (larger clouds); reads the CSR cloud
__global__ void __launch_bounds__(ROI_THREADS) k_roi_ih_w (const RoiRec* __restrict__ rois, const uint32_t* __restrict__ list, uint32_t n,
	Params prm, const uint32_t* __restrict__ cloud, double* out, int out_stride, size_t memw)
{
	extern __shared__ __align__(16) uint32_t smem[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint32_t* mem = smem + (size_t)warp * memw;
	Cta c { lane, 32, nullptr, 1 };
	for (uint32_t k = blockIdx.x * ROI_WARPS + warp; k < n; k += gridDim.x * ROI_WARPS)
	{
		const uint32_t ri = list ? list[k] : k;
		const RoiRec r = rois[ri];
		ih_body (c, r, prm, cloud + r.cloud_off, r.area, mem, out + (size_t)ri * out_stride + COL_IH);
		c.sync();
	}
}
__global__ void __launch_bounds__(ROI_THREADS) k_roi_ih (const RoiRec* __restrict__ rois, const uint32_t* __restrict__ list, uint32_t n,
	Params prm, const uint32_t* __restrict__ cloud, double* out, int out_stride)
{
	extern __shared__ __align__(16) uint32_t smem[];
	Cta c { (int)threadIdx.x, (int)blockDim.x, (double*)smem };
	uint32_t* mem = smem + RED_WORDS;
	for (uint32_t k = blockIdx.x; k < n; k += gridDim.x)
	{
		const uint32_t ri = list ? list[k] : k;
		const RoiRec r = rois[ri];
		ih_body (c, r, prm, cloud + r.cloud_off, r.area, mem, out + (size_t)ri * out_stride + COL_IH);
		c.sync();
	}
}

// *BASIC_MORPHOLOGY*: warp per ROI (AABBs of the shared-memory size classes) / CTA per ROI (larger ones); reads the slide
__global__ void __launch_bounds__(256) k_roi_morph_w (ImageView im, const RoiRec* __restrict__ rois, const uint32_t* __restrict__ list,
	uint32_t n, Params prm, double* out, int out_stride)
{
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (warp >= n)
		return;
	const uint32_t ri = list ? list[warp] : warp;
	Cta c { (int)(threadIdx.x & 31), 32, nullptr, 1 };
	morph_body (c, im, rois[ri], prm, out + (size_t)ri * out_stride + COL_MORPH);
}
__global__ void __launch_bounds__(ROI_THREADS) k_roi_morph (ImageView im, const RoiRec* __restrict__ rois, const uint32_t* __restrict__ list,
	uint32_t n, Params prm, double* out, int out_stride)
{
	extern __shared__ __align__(16) uint32_t smem[];
	Cta c { (int)threadIdx.x, (int)blockDim.x, (double*)smem };
	for (uint32_t k = blockIdx.x; k < n; k += gridDim.x)
	{
		const uint32_t ri = list ? list[k] : k;
		morph_body (c, im, rois[ri], prm, out + (size_t)ri * out_stride + COL_MORPH);
		c.sync();
	}
}

// ------------------------------------------------------------------------------------------------
// Large / sparse label values (the reference keys its ROIs by a hash of the int label, pixel_feed.cpp:28 -- any value goes;
// the label table here is dense): masks whose largest label is beyond RELABEL_MIN are re-labelled on the device to
// 1 .. K in ascending label order before the scan, and the table's label column is mapped back at the end.
// A presence bitmap over [0, max_label] (one atomicOr per label RUN of a row), popcounts per 1024-word block, one
// single-CTA scan of the block sums, exclusive prefix per word: rank (label) = prefix[word] + popc (bits below) + 1.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t RELABEL_MIN = 1u << 24;	// below: the dense table simply grows (512 MB at the limit)
constexpr int RL_BLOCK_WORDS = 1024;

__global__ void k_rl_mark (const uint32_t* __restrict__ label, size_t npix, uint32_t* bitmap)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x)
	{
		const uint32_t l = label[i];
		if (l && (i == 0 || label[i - 1] != l))
			atomicOr (bitmap + (l >> 5), 1u << (l & 31u));
	}
}
// per block of RL_BLOCK_WORDS words: its popcount
__global__ void __launch_bounds__(256) k_rl_block_popc (const uint32_t* __restrict__ bitmap, size_t nwords, uint32_t* blocksum)
{
	__shared__ uint32_t red[8];
	const size_t w0 = (size_t)blockIdx.x * RL_BLOCK_WORDS;
	uint32_t c = 0;
	for (int k = threadIdx.x; k < RL_BLOCK_WORDS; k += 256)
		if (w0 + k < nwords)
			c += (uint32_t)__popc (bitmap[w0 + k]);
	c = __reduce_add_sync (0xffffffffu, c);
	if ((threadIdx.x & 31) == 0)
		red[threadIdx.x >> 5] = c;
	__syncthreads();
	if (threadIdx.x == 0)
	{
		uint32_t t = 0;
		for (int k = 0; k < 8; k++)
			t += red[k];
		blocksum[blockIdx.x] = t;
	}
}
// exclusive scan of the block sums in place (one CTA; nblocks <= 2^17), total -> *total
__global__ void __launch_bounds__(1024) k_rl_scan_blocks (uint32_t* blocksum, uint32_t nblocks, uint32_t* total)
{
	__shared__ uint32_t part[1024];
	const uint32_t per = (nblocks + 1023) / 1024, b0 = threadIdx.x * per;
	uint32_t c = 0;
	for (uint32_t k = b0; k < b0 + per && k < nblocks; k++)
		c += blocksum[k];
	part[threadIdx.x] = c;
	__syncthreads();
	if (threadIdx.x == 0)
	{
		uint32_t run = 0;
		for (int k = 0; k < 1024; k++)
		{
			const uint32_t t = part[k];
			part[k] = run;
			run += t;
		}
		*total = run;
	}
	__syncthreads();
	uint32_t run = part[threadIdx.x];
	for (uint32_t k = b0; k < b0 + per && k < nblocks; k++)
	{
		const uint32_t t = blocksum[k];
		blocksum[k] = run;
		run += t;
	}
}
// exclusive prefix of the popcounts per word (one thread walks 32 consecutive words: 32 threads per block of 1024 words)
__global__ void k_rl_word_prefix (const uint32_t* __restrict__ bitmap, size_t nwords, const uint32_t* __restrict__ blockbase, uint32_t* wprefix)
{
	__shared__ uint32_t seg[32];
	const size_t w0 = (size_t)blockIdx.x * RL_BLOCK_WORDS + (size_t)threadIdx.x * 32;
	uint32_t c = 0;
	for (int k = 0; k < 32; k++)
		if (w0 + k < nwords)
			c += (uint32_t)__popc (bitmap[w0 + k]);
	seg[threadIdx.x] = c;
	__syncwarp();
	uint32_t run = blockbase[blockIdx.x];
	for (int k = 0; k < (int)threadIdx.x; k++)
		run += seg[k];
	for (int k = 0; k < 32; k++)
		if (w0 + k < nwords)
		{
			wprefix[w0 + k] = run;
			run += (uint32_t)__popc (bitmap[w0 + k]);
		}
}
// the K original labels in ascending order: labmap[rank - 1]
__global__ void k_rl_list (const uint32_t* __restrict__ bitmap, size_t nwords, const uint32_t* __restrict__ wprefix, uint32_t* labmap)
{
	for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (size_t)gridDim.x * blockDim.x)
	{
		uint32_t bits = bitmap[w], pos = wprefix[w];
		while (bits)
		{
			const int b = __ffs (bits) - 1;
			labmap[pos++] = (uint32_t)(w << 5) + (uint32_t)b;
			bits &= bits - 1;
		}
	}
}
__global__ void k_rl_apply (const uint32_t* __restrict__ label, size_t npix, const uint32_t* __restrict__ bitmap,
	const uint32_t* __restrict__ wprefix, uint32_t* dense)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x)
	{
		const uint32_t l = label[i];
		dense[i] = l ? wprefix[l >> 5] + (uint32_t)__popc (bitmap[l >> 5] & ((1u << (l & 31u)) - 1u)) + 1u : 0u;
	}
}

// final gather: canonical table -> [label, t_index, selected columns], values through force_finite_number(v, noval)
// (reference output_2_buffer.cpp:490,577)
__global__ void k_select_columns (const double* __restrict__ full, int full_stride, const RoiRec* __restrict__ rois, uint32_t n,
	const int* __restrict__ sel, int nsel, double noval, double t_index, double* __restrict__ dst, const uint32_t* __restrict__ labmap = nullptr)
{
	const size_t total = (size_t)n * (nsel + 2);
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
	{
		const uint32_t r = (uint32_t)(i / (nsel + 2));
		const int cidx = (int)(i - (size_t)r * (nsel + 2));
		double v;
		if (cidx == 0)
			v = (double)(int32_t)(labmap ? labmap[rois[r].label - 1u] : rois[r].label);	// the reference keeps labels as int (pixel_feed.cpp:28); labmap: the mask was re-labelled (k_rl_*)
		else if (cidx == 1)
			v = t_index;
		else
			v = finite_or (full[(size_t)r * full_stride + sel[cidx - 2]], noval);
		dst[i] = v;
	}
}

} // namespace nyx
