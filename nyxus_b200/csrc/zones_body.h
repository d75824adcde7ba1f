// nyxus_b200 -- per-ROI zone families: GLSZM (size-zone) and GLDZM (distance-zone).
//
// GLSZM semantics follow the reference's GLSZMFeature (features/glszm.cpp):
//   blank test   :61-65     aux_min == aux_max -> 16 x softNaN
//   zones        :108-194   NOT 8-connected components: a depth-first flood that only ever steps E, SE, S, SW from the
//                           raster-first unvisited seed.  All steps go forward in raster order, so the zone of a cell is
//                           the raster-earliest seed that reaches it:
//                               label(p) = min( idx(p), label(q) : q in {W, NW, N, NE}, level(q) == level(p) )
//                           evaluated in raster order.  Here: rows top-down; per row first the {NW,N,NE} minimum of every
//                           cell in parallel, then a running minimum along each maximal same-level horizontal run.
//   Np           :199-204   non-zero cells of the BINNED tile (MATLAB binning: all w*h cells)
//   Ns = w*h, matrix P[Ng x Ns] :211-232  -- never materialised here; the zone list is sorted into unique
//                           (level,size) keys with counts, which is what every feature actually consumes.
//   features     :342-397 (calc_sums_of_P), :420-614
//
// GLDZM semantics follow GLDZMFeature (features/gldzm.cpp):
//   zones        :94-237    4-connected components of equal binned level (E,S,W,N with back-tracking => true components);
//                           zero level skipped only in IBSI mode (:104-106)
//   dist2border  :306-351   per direction: steps to the first zero cell OR to the tile edge, then +1; 0 -> 1
//   zone metric  = minimum over member pixels; matrix[level][metric] :239-258, :38-50
//   features     :353-478   Nv = aux_area; entropy with real log2
//
// B200 design: one WARP (small tiles) or one CTA per ROI, labels in shared memory.  GLSZM labels: per row a
// {NW,N,NE} minimum and a segmented min-scan along the same-level runs (warp shuffles in the warp-per-ROI form).
// GLDZM components by lock-free union-find (atomicMin hooking).  The sparse (level,size) matrix of GLSZM and its two
// marginals live in ONE open-addressing hash table of integer counters (three key kinds); GLDZM's small
// [level][distance] matrix is dense.  Reducers are "owner computes": every zone re-reads the count of its own matrix
// cell, so sum_cells f(P) == sum_zones f(P_zone)/P_zone and nothing is sorted or swept.  Integer counters make every
// matrix order-independent; FP sums use fixed lane assignments => deterministic.
#pragma once
#include "nyx_common.h"
#include "nyx_features.h"
#include "glcm_body.h"		// Counters<> (16-bit packed counters)

namespace nyx
{

NYX_D uint32_t atom_min_ret (const Cta&, uint32_t* p, uint32_t v)
{
#if NYX_DEVICE_BUILD
	return atomicMin (p, v);
#else
	uint32_t o = *p;
	if (v < o) *p = v;
	return o;
#endif
}
NYX_D uint32_t atom_inc_ret (const Cta&, uint32_t* p)
{
#if NYX_DEVICE_BUILD
	return atomicAdd (p, 1u);
#else
	return (*p)++;
#endif
}
NYX_D uint32_t atom_cas (const Cta&, uint32_t* p, uint32_t cmp, uint32_t val)
{
#if NYX_DEVICE_BUILD
	return atomicCAS (p, cmp, val);
#else
	uint32_t o = *p;
	if (o == cmp) *p = val;
	return o;
#endif
}

// ---- open-addressing table of integer counters: keys[slots] (0 = empty) + 16-bit cnts[slots] (slots / 2 words), slots = power of two
NYX_HD uint32_t htab_mix (uint32_t k)
{
	k *= 0x9E3779B1u;
	return k ^ (k >> 15);
}
NYX_D void htab_add (const Cta& c, uint32_t* keys, uint32_t* cnts, uint32_t mask, uint32_t key)
{
	uint32_t slot = htab_mix (key) & mask;
	for (;;)
	{
		uint32_t k = ((volatile uint32_t*)keys)[slot];
		if (k == 0)
			k = atom_cas (c, keys + slot, 0u, key);
		if (k == 0 || k == key)
			break;
		slot = (slot + 1) & mask;
	}
	atom_add (c, cnts + (slot >> 1), 1u << ((slot & 1u) * 16u));	// 16-bit packed counts (a count is at most the number of zones < 65536)
}
NYX_D uint32_t htab_get (const uint32_t* keys, const uint32_t* cnts, uint32_t mask, uint32_t key)
{
	uint32_t slot = htab_mix (key) & mask;
	while (keys[slot] != key)
		slot = (slot + 1) & mask;
	return ((const uint16_t*)cnts)[slot];
}

// upper bound of the distinct keys one GLSZM tile can produce: (level,size) keys -- sizes 1,2,3.. each usable by at
// most Lmax levels with sum of sizes <= cells -- plus one key per distinct level and per distinct size
NYX_HD uint32_t glszm_hash_slots (int Lmax, uint32_t maxcells)
{
	unsigned long long K = 0, cost = 0;
	for (unsigned long long m = 1; cost < maxcells; m++)
	{
		unsigned long long can = (maxcells - cost) / m;
		if (can > (unsigned long long)Lmax) can = (unsigned long long)Lmax;
		if (!can) break;
		K += can;
		cost += can * m;
	}
	unsigned long long nl = (unsigned long long)Lmax < maxcells ? (unsigned long long)Lmax : maxcells, ns = 2;
	while (ns * (ns + 1) / 2 <= maxcells) ns++;
	K += nl + ns;
	return next_pow2 ((uint32_t)(K + K / 8 + 2));
}

NYX_D uint32_t uf_find (const volatile uint32_t* parent, uint32_t x)
{
	uint32_t p = parent[x];
	while (p != x)
	{
		x = p;
		p = parent[x];
	}
	return x;
}

NYX_D void uf_unite (const Cta& c, uint32_t* parent, uint32_t a, uint32_t b)
{
	for (;;)
	{
		a = uf_find (parent, a);
		b = uf_find (parent, b);
		if (a == b)
			return;
		if (a < b)
		{
			uint32_t t = a;
			a = b;
			b = t;
		}
		uint32_t old = atom_min_ret (c, parent + a, b);	// hook the larger root under the smaller
		if (old == a)
			return;
		a = old;
	}
}

// 16-bit minimum into the halfword h16[i] of a packed array (CAS on the containing 32-bit word; a handful of calls per
// cell, the halved footprint buys a whole extra CTA per SM)
NYX_D void atom_min_u16 (const Cta& c, uint32_t* w32, uint32_t i, uint32_t v)
{
	uint32_t* p = w32 + (i >> 1);
	const uint32_t sh = (i & 1u) * 16u;
#if NYX_DEVICE_BUILD
	uint32_t old = *(volatile uint32_t*)p;
	for (;;)
	{
		if (((old >> sh) & 0xFFFFu) <= v)
			return;
		const uint32_t upd = (old & ~(0xFFFFu << sh)) | (v << sh);
		const uint32_t seen = atomicCAS (p, old, upd);
		if (seen == old)
			return;
		old = seen;
	}
#else
	(void)c;
	if (((*p >> sh) & 0xFFFFu) > v)
		*p = (*p & ~(0xFFFFu << sh)) | (v << sh);
#endif
}

typedef Counters<true> Counters16;

struct ZoneDebug
{
	uint32_t* keys;		// one (level<<16 | minor) key per zone of this ROI (unsorted), capacity = ncell
	uint32_t* nkeys;	// number of keys (zero-initialised by the host)
};

// ------------------------------------------------------------------------------------------------
// GLSZM
// ------------------------------------------------------------------------------------------------
// region: hash keys[slots] + 16-bit cnts[slots] (zero on entry/exit) + lab (u16 per cell) + zone sizes (u16 per cell)
// Lmarg > 0 (2D, at most GLSZM_LMARG_MAX levels): + DIRECT 16-bit arrays for the two marginals -- zones per level [Lmarg + 1]
// and zones per size for sizes below GLSZM_DIRECT_SIZES -- right behind the table (zero on entry / exit like it).  Four of
// the six table operations per zone were marginal keys; larger sizes (few zones) still go through the table.
constexpr int GLSZM_LMARG_MAX = 255, GLSZM_DIRECT_SIZES = 64;
NYX_HD size_t glszm_marg_words (int Lmarg)
{
	return Lmarg > 0 ? ((size_t)(Lmarg + 1) + GLSZM_DIRECT_SIZES + 1) / 2 : 0;
}
NYX_HD size_t glszm_smem_words (uint32_t maxcells, uint32_t slots, int Lmarg = 0)
{
	return (size_t)slots + slots / 2 + glszm_marg_words (Lmarg) + 2 * (((size_t)maxcells + 1) / 2);
}

NYX_D void glszm_reduce (const Cta& c, const Params& prm, const uint32_t* cell, uint32_t* keys, uint32_t* cnts, uint32_t mask,
	uint16_t* lab, uint32_t* zsw, int ncell, uint32_t skip, double np_fixed, double* out, const ZoneDebug* dbg, int Lmarg = 0);

NYX_D void glszm_body (const Cta& c, const RoiRec& r, const Params& prm, const uint32_t* cell, uint32_t* mem, uint32_t maxcells,
	uint32_t slots, double* out, const ZoneDebug* dbg, int Lmarg = 0)
{
	const int w = r.w, h = r.h, ncell = w * h;
	if (r.imin == r.imax)
	{
		for (int i = c.tid; i < NCOL_GLSZM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	uint32_t* keys = mem;						// [slots]
	uint32_t* cnts = keys + slots;					// [slots]
	const size_t mw = glszm_marg_words (Lmarg);			// direct marginal arrays behind the table (see glszm_smem_words)
	uint16_t* lab = (uint16_t*)(cnts + slots / 2 + mw);		// [ncell] zone label = raster index of the seed
	uint32_t* zsw = cnts + slots / 2 + mw + ((size_t)maxcells + 1) / 2;	// [ncell] zone size at the seed index, packed u16 pairs
	const uint16_t* zs = (const uint16_t*)zsw;
	const uint32_t mask = slots - 1;

	for (int i = c.tid; i < (ncell + 1) / 2; i += c.nthr)
		zsw[i] = 0;

	// ---- forward-reachability labels, row by row
#if NYX_DEVICE_BUILD
	if (c.wmode && w <= 32)
	{
		// tiles at most 32 cells wide (nearly all of them): a lane owns a COLUMN.  The levels / labels of the previous row stay
		// in registers, so the {NW, N, NE} candidates are shuffles, not dependent shared-memory loads; the segmented min-scan
		// along the row's same-level runs stops as soon as no run is longer than the distance already covered
		const int lane = c.tid;
		const bool inx = lane < w;
		uint32_t plv = 0, plab = 0;
		for (int y = 0; y < h; y++)
		{
			const int i = y * w + lane;
			uint32_t lv = 0, m = 0xFFFFu;
			if (inx)
			{
				lv = cell[i] & CELL_LEVEL;
				m = (uint32_t)i;
			}
			const uint32_t nw_lv = __shfl_up_sync (0xffffffffu, plv, 1), nw_lab = __shfl_up_sync (0xffffffffu, plab, 1);
			const uint32_t ne_lv = __shfl_down_sync (0xffffffffu, plv, 1), ne_lab = __shfl_down_sync (0xffffffffu, plab, 1);
			if (lv && y > 0)
			{
				if (lane > 0 && nw_lv == lv) m = min (m, nw_lab);
				if (plv == lv) m = min (m, plab);
				if (lane + 1 < w && ne_lv == lv) m = min (m, ne_lab);
			}
			const uint32_t left_lv = __shfl_up_sync (0xffffffffu, lv, 1);
			const bool head = !lv || lane == 0 || left_lv != lv;
			const unsigned heads = __ballot_sync (0xffffffffu, head);
			const int start = 31 - __clz (heads & (0xFFFFFFFFu >> (31 - lane)));	// lane of the head of my run (bit 0 is always set)
			unsigned pos = ~heads;		// lanes at distance >= o from their run head (o = 1 here)
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				if (!pos)
					break;
				const uint32_t t = __shfl_up_sync (0xffffffffu, m, o);
				if (lane - o >= start)
					m = min (m, t);
				pos &= pos << o;
			}
			if (inx)
				lab[i] = (uint16_t)m;
			plv = lv;
			plab = m;
		}
		__syncwarp();
	}
	else if (c.wmode)
	{
		const int lane = c.tid;
		for (int y = 0; y < h; y++)
		{
			uint32_t carry_lv = 0, carry_lab = 0;	// last cell of the previous 32-cell chunk of this row
			for (int x0 = 0; x0 < w; x0 += 32)
			{
				const int x = x0 + lane, i = y * w + x;
				uint32_t lv = 0, m = 0xFFFFu;
				if (x < w)
				{
					lv = cell[i] & CELL_LEVEL;
					m = (uint32_t)i;
					if (lv && y > 0)
					{
						const int up = i - w;
						if (x > 0 && (cell[up - 1] & CELL_LEVEL) == lv) m = min (m, (uint32_t)lab[up - 1]);
						if ((cell[up] & CELL_LEVEL) == lv) m = min (m, (uint32_t)lab[up]);
						if (x + 1 < w && (cell[up + 1] & CELL_LEVEL) == lv) m = min (m, (uint32_t)lab[up + 1]);
					}
				}
				// run heads: level differs from the left neighbour (the carry for lane 0)
				uint32_t left_lv = __shfl_up_sync (0xffffffffu, lv, 1);
				if (lane == 0)
					left_lv = carry_lv;
				const bool head = !lv || left_lv != lv;
				if (lane == 0 && !head)
					m = min (m, carry_lab);
				const unsigned heads = __ballot_sync (0xffffffffu, head);
				// lane of the head of my run (bit 0 is forced: a run continuing from the carry starts at lane 0)
				const int start = 31 - __clz ((heads | 1u) & (0xFFFFFFFFu >> (31 - lane)));
#pragma unroll
				for (int o = 1; o < 32; o <<= 1)
				{
					const uint32_t t = __shfl_up_sync (0xffffffffu, m, o);
					if (lane - o >= start)
						m = min (m, t);
				}
				if (x < w)
					lab[i] = (uint16_t)m;
				carry_lv = __shfl_sync (0xffffffffu, lv, 31);
				carry_lab = __shfl_sync (0xffffffffu, m, 31);
			}
			__syncwarp();
		}
	}
	else
#endif
	{
		c.sync();
		for (int y = 0; y < h; y++)
		{
			for (int x = c.tid; x < w; x += c.nthr)
			{
				int i = y * w + x;
				uint32_t lv = cell[i] & CELL_LEVEL;
				uint32_t m = (uint32_t)i;
				if (lv && y > 0)
				{
					const int up = i - w;
					if (x > 0 && (cell[up - 1] & CELL_LEVEL) == lv) m = lab[up - 1] < m ? lab[up - 1] : m;
					if ((cell[up] & CELL_LEVEL) == lv) m = lab[up] < m ? lab[up] : m;
					if (x + 1 < w && (cell[up + 1] & CELL_LEVEL) == lv) m = lab[up + 1] < m ? lab[up + 1] : m;
				}
				lab[i] = (uint16_t)m;
			}
			c.sync();
			for (int x = c.tid; x < w; x += c.nthr)
			{
				int i = y * w + x;
				uint32_t lv = cell[i] & CELL_LEVEL;
				if (!lv || (x > 0 && (cell[i - 1] & CELL_LEVEL) == lv))
					continue;	// not the start of a horizontal run
				uint32_t cur = lab[i];
				for (int x2 = x + 1; x2 < w && (cell[i + (x2 - x)] & CELL_LEVEL) == lv; x2++)
				{
					uint32_t v = lab[i + (x2 - x)];
					cur = v < cur ? v : cur;
					lab[i + (x2 - x)] = (uint16_t)cur;
				}
			}
			c.sync();
		}
	}
	c.sync();
	glszm_reduce (c, prm, cell, keys, cnts, mask, lab, zsw, ncell, 0u, -1.0, out, dbg, Lmarg);
}

// Common tail of 2D and 3D GLSZM.  lab[i] = zone id (cell index of the zone's representative) of every cell whose level is
// not `skip`; zsw: zeroed packed u16 size counters.  np_fixed >= 0 overrides Np (3D: the voxel-cloud size, 3d_glszm.cpp:534).
NYX_D void glszm_reduce (const Cta& c, const Params& prm, const uint32_t* cell, uint32_t* keys, uint32_t* cnts, uint32_t mask,
	uint16_t* lab, uint32_t* zsw, int ncell, uint32_t skip, double np_fixed, double* out, const ZoneDebug* dbg, int Lmarg)
{
	const uint16_t* zs = (const uint16_t*)zsw;
	const uint32_t slots = mask + 1;
	uint32_t* margw = Lmarg > 0 ? cnts + slots / 2 : nullptr;	// zones per level [Lmarg + 1], then zones per size < GLSZM_DIRECT_SIZES (u16 each)
	const uint32_t sz0 = (uint32_t)Lmarg + 1;			// index of size 0 in the packed array
	// ---- zone sizes (packed 16-bit counters: a size is at most the cell count < 65536), Np
	unsigned long long np[1] = { 0 };
	for (int i = c.tid; i < ncell; i += c.nthr)
		if ((cell[i] & CELL_LEVEL) != skip)
		{
			np[0]++;
			const uint32_t l = lab[i];
			atom_add (c, zsw + (l >> 1), 1u << ((l & 1) * 16));
		}
	c.sync();

	// ---- warp per ROI: compact the seeds (cells that are their own label) into a list, IN PLACE over the label array --
	// the labels are dead once the sizes are counted, a list slot is never above the cell it came from, and a warp walks
	// the cells 32 at a time (every lane has read its label before the ballot).  The two passes over the zones then run
	// with every lane busy instead of skipping the ~45 % of the cells that are not seeds.
	int nseed = -1;
#if NYX_DEVICE_BUILD
	if (c.wmode)
	{
		int run = 0;
		for (int base = 0; base < ncell; base += 32)
		{
			const int i = base + c.tid;
			const bool seed = i < ncell && (cell[i] & CELL_LEVEL) != skip && lab[i] == (uint32_t)i;
			const unsigned b = __ballot_sync (0xffffffffu, seed);
			__syncwarp();	// every lane has read its label: the list may overwrite the labels of this chunk (a vote orders no memory)
			if (seed)
				lab[run + __popc (b & ((1u << c.tid) - 1u))] = (uint16_t)i;
			run += __popc (b);	// (list positions never exceed the chunk's own cells: later chunks are still intact)
		}
		__syncwarp();
		nseed = run;
	}
#endif
	const int nloop = nseed >= 0 ? nseed : ncell;

	// ---- pass 1 over the zones (= seed cells): matrix cell (level,size) and both marginals into the hash; direct sums
	enum { Q_NZ, Q_LAE, Q_MUJ, Q_MUI, Q_HGL, Q_NP, Q_N };
	unsigned long long q[Q_N];
	for (int k = 0; k < Q_N; k++)
		q[k] = 0;
	q[Q_NP] = np[0];
	enum { A_SAE, A_LGL, A_SAL, A_SAH, A_LAL, A_LAH, A_N };
	double s[A_N];
	for (int k = 0; k < A_N; k++)
		s[k] = 0;
	for (int k = c.tid; k < nloop; k += c.nthr)
	{
		const int i = nseed >= 0 ? (int)lab[k] : k;
		const uint32_t lv = cell[i] & CELL_LEVEL;
		if (nseed < 0 && (lv == skip || lab[i] != (uint32_t)i))
			continue;
		const uint32_t sz = zs[i];
		htab_add (c, keys, cnts, mask, (lv << 16) | sz);
		if (margw)
		{
			atom_add (c, margw + (lv >> 1), 1u << ((lv & 1u) * 16u));
			if (sz < (uint32_t)GLSZM_DIRECT_SIZES)
				atom_add (c, margw + ((sz0 + sz) >> 1), 1u << (((sz0 + sz) & 1u) * 16u));
			else
				htab_add (c, keys, cnts, mask, sz);
		}
		else
		{
			htab_add (c, keys, cnts, mask, lv << 16);
			htab_add (c, keys, cnts, mask, sz);
		}
		if (dbg && dbg->keys)
			dbg->keys[atom_inc_ret (c, dbg->nkeys)] = (lv << 16) | sz;
		q[Q_NZ] += 1;
		q[Q_LAE] += (unsigned long long)sz * sz;			// :437-440
		q[Q_MUJ] += sz;							// :380
		q[Q_MUI] += lv;							// :381
		q[Q_HGL] += (unsigned long long)lv * lv;			// :574-578
		emphasis6_int (s, lv, sz);	// :372-375, 424-427 (sz < 65536: int j*j exact), 559-563
	}
	c.sync();
	cta_sum_u64<Q_N> (c, q);
	cta_sum<A_N> (c, s);
	const uint32_t nz = (uint32_t)q[Q_NZ];
	if (nz == 0)	// glszm.cpp:244-248 (nothing was inserted: the table is still all zero)
	{
		for (int i = c.tid; i < NCOL_GLSZM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	const double sum = (double)nz;
	const double mui = (double)q[Q_MUI] / sum, muj = (double)q[Q_MUJ] / sum;
	const double inv_sum = 1.0 / sum;

	// ---- pass 2: terms that need the counts (owner computes)
	unsigned long long q2[2] = { 0, 0 };	// sum n_level^2 (:450-454), sum n_size^2 (:481-485)
	double t[3] = { 0, 0, 0 };		// GLV, ZV, ZE
	for (int k = c.tid; k < nloop; k += c.nthr)
	{
		const int i = nseed >= 0 ? (int)lab[k] : k;
		const uint32_t lv = cell[i] & CELL_LEVEL;
		if (nseed < 0 && (lv == skip || lab[i] != (uint32_t)i))
			continue;
		const uint32_t sz = zs[i];
		if (margw)
		{
			const uint16_t* m16 = (const uint16_t*)margw;
			q2[0] += m16[lv];
			q2[1] += sz < (uint32_t)GLSZM_DIRECT_SIZES ? (uint32_t)m16[sz0 + sz] : htab_get (keys, cnts, mask, sz);
		}
		else
		{
			q2[0] += htab_get (keys, cnts, mask, lv << 16);
			q2[1] += htab_get (keys, cnts, mask, sz);
		}
		const double di = (double)lv - mui, dj = (double)sz - muj;
		t[0] += di * di;									// :518-526
		t[1] += dj * dj;									// :535-542
		const double pn = (double)htab_get (keys, cnts, mask, (lv << 16) | sz) * inv_sum;
		t[2] += fast_log10 (dadd (pn, EPS_ZONES));						// :377-378
	}
	cta_sum_u64<2> (c, q2);
	cta_sum<3> (c, t);
	c.sync();
	for (uint32_t i = c.tid; i < slots + slots / 2 + (uint32_t)glszm_marg_words (Lmarg); i += c.nthr)	// restore the all-zero table (keys + packed counts + direct marginals)
		keys[i] = 0;
	c.sync();
	if (c.tid == 0)
	{
		out[S_SAE] = s[A_SAE] * inv_sum;
		out[S_LAE] = (double)q[Q_LAE] * inv_sum;
		out[S_GLN] = (double)q2[0] * inv_sum;
		out[S_GLNN] = (double)q2[0] * (inv_sum * inv_sum);
		out[S_SZN] = (double)q2[1] * inv_sum;
		out[S_SZNN] = (double)q2[1] * (inv_sum * inv_sum);
		out[S_ZP] = sum / (np_fixed >= 0 ? np_fixed : (double)q[Q_NP]);	// :507-511
		out[S_GLV] = t[0] * inv_sum;
		out[S_ZV] = t[1] * inv_sum;
		out[S_ZE] = -(t[2] * inv_sum) * (1.0 / LOG10_2);
		out[S_LGLZE] = s[A_LGL] * inv_sum;
		out[S_HGLZE] = (double)q[Q_HGL] * inv_sum;
		out[S_SALGLE] = s[A_SAL] * inv_sum;
		out[S_SAHGLE] = s[A_SAH] * inv_sum;
		out[S_LALGLE] = s[A_LAL] * inv_sum;
		out[S_LAHGLE] = s[A_LAH] * inv_sum;
	}
}

// ------------------------------------------------------------------------------------------------
// GLDZM
// ------------------------------------------------------------------------------------------------
// largest zone metric of a tile: min over the 4 directions of (steps + 1) <= (min(w,h) + 1) / 2 + 1
NYX_HD int gldzm_nd_bound (uint32_t maxcells, int maxwh)
{
	int s = 1;
	while ((unsigned long long)s * s < maxcells) s++;	// min(w,h) <= sqrt(cells)
	if (s > maxwh) s = maxwh;
	return (s + 1) / 2 + 2;
}
// region: P[Lmax+1][Nd] (16-bit packed: a count is at most the number of zones < 65536) + Ri[Lmax+1] + Cd[Nd] (zero on
// entry/exit) + parent[ncell] + zmin[ncell] (16-bit packed)
NYX_HD size_t gldzm_p_words (int Lmax, int Nd)
{
	return ((size_t)(Lmax + 1) * Nd + 1) / 2;
}
NYX_HD size_t gldzm_smem_words (int Lmax, int Nd, uint32_t maxcells)
{
	return gldzm_p_words (Lmax, Nd) + (Lmax + 1) + Nd + (size_t)maxcells + ((size_t)maxcells + 1) / 2;
}

// distance of cell (x,y) to the ROI/tile border exactly as gldzm.cpp:306-351 (zeros = level-0 cells of the binned tile)
NYX_D int gldzm_dist2border (const uint32_t* cell, int w, int h, int x, int y, bool nozeros)
{
	int dl, dr, dt, db;
	if (nozeros)
	{
		dl = x + 1;
		dr = w - x;
		dt = y + 1;
		db = h - y;
	}
	else
	{
		dl = 0;
		for (int x0 = x - 1; x0 >= 0; x0--)
			if ((cell[y * w + x0] & CELL_LEVEL) == 0 || x0 == 0) { dl = x - x0; break; }
		dr = 0;
		for (int x0 = x + 1; x0 < w; x0++)
			if ((cell[y * w + x0] & CELL_LEVEL) == 0 || x0 == w - 1) { dr = x0 - x; break; }
		dt = 0;
		for (int y0 = y - 1; y0 >= 0; y0--)
			if ((cell[y0 * w + x] & CELL_LEVEL) == 0 || y0 == 0) { dt = y - y0; break; }
		db = 0;
		for (int y0 = y + 1; y0 < h; y0++)
			if ((cell[y0 * w + x] & CELL_LEVEL) == 0 || y0 == h - 1) { db = y0 - y; break; }
		dl++; dr++; dt++; db++;
	}
	int m = dl < dr ? dl : dr;
	m = dt < m ? dt : m;
	m = db < m ? db : m;
	return m == 0 ? 1 : m;
}

NYX_D void gldzm_body (const Cta& c, const RoiRec& r, const Params& prm, const uint32_t* cell, uint32_t* mem, int Lmax, int Nd,
	uint32_t maxcells, double* out, const ZoneDebug* dbg)
{
	const int w = r.w, h = r.h, ncell = w * h;
	if (r.imin == r.imax)	// gldzm.cpp:266-289
	{
		for (int i = c.tid; i < NCOL_GLDZM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	const bool ibsi = prm.ibsi != 0;
	const Counters16 P { mem };				// [Lmax+1][Nd]  zones per (level, metric), 16-bit packed
	uint32_t* Ri = mem + gldzm_p_words (Lmax, Nd);		// [Lmax+1]      zones per level
	uint32_t* Cd = Ri + (Lmax + 1);				// [Nd]          zones per metric
	uint32_t* parent = Cd + Nd;				// [ncell]
	uint32_t* zminw = parent + maxcells;			// [ncell] zone metric at the root, 16-bit packed
	const uint16_t* zmin = (const uint16_t*)zminw;
	FastDiv fw;
	fw.init ((uint32_t)w);

	// ---- 4-connected components of equal level + zone metric (= minimum member distance).
#if NYX_DEVICE_BUILD
	if (c.wmode)
	{
		// Warp per ROI, ONE pass over the cells in raster order, 32 at a time:
		//  * the head of a cell's horizontal same-level run is the nearest lane at or below it whose left neighbour differs
		//    (ballot + clz), or the head carried over from the previous 32 cells: parent[i] = head;
		//  * the cell's distance goes straight to its run head's metric (the head is in a register: no find);
		//  * one union per vertically adjacent pair of RUNS (a pair whose left neighbours already joined the same two runs is
		//    skipped) -- both cells have their parents by then (the upper one lies at least one row back).
		// Afterwards only the run heads that are not roots (runs hooked under an earlier run: few) look for their root and
		// hand their metric over.  A metric other than 0xFFFF marks a run head.
		for (int i = c.tid; i < (ncell + 1) / 2; i += c.nthr)
			zminw[i] = 0xFFFFFFFFu;
		__syncwarp();
		const int lane = c.tid;
		uint32_t carry = 0, carry_lv = 0xFFFFFFFFu;	// run head / level of the last cell of the previous chunk
		for (int base = 0; base < ncell; base += 32)
		{
			const int i = base + lane;
			const bool in = i < ncell;
			uint32_t lv = 0;
			int x = 0, y = 0;
			if (in)
			{
				lv = cell[i] & CELL_LEVEL;
				y = (int)fw.div ((uint32_t)i);
				x = i - y * w;
			}
			const bool skip = !in || (ibsi && !lv);
			uint32_t left_lv = __shfl_up_sync (0xffffffffu, lv, 1);
			if (lane == 0)
				left_lv = carry_lv;
			const bool left_same = !skip && x > 0 && left_lv == lv;
			const unsigned heads = __ballot_sync (0xffffffffu, !left_same);
			const unsigned below = heads & (0xFFFFFFFFu >> (31 - lane));
			const uint32_t hd = below ? (uint32_t)(base + 31 - __clz (below)) : carry;
			bool link = false;
			if (in)
			{
				parent[i] = hd;
				if (!skip)
				{
					atom_min_u16 (c, zminw, hd, (uint32_t)gldzm_dist2border (cell, w, h, x, y, !ibsi));
					if (y > 0 && (cell[i - w] & CELL_LEVEL) == lv)
						link = !(left_same && (cell[i - w - 1] & CELL_LEVEL) == lv);
				}
			}
			__syncwarp();	// the parents of this chunk are in place
			if (link)
				uf_unite (c, parent, (uint32_t)i, (uint32_t)(i - w));
			carry = __shfl_sync (0xffffffffu, hd, 31);
			carry_lv = __shfl_sync (0xffffffffu, in ? lv : 0xFFFFFFFFu, 31);
		}
		for (int base = 0; base < ncell; base += 32)
		{
			__syncwarp();	// (a metric is read plainly here and updated by CAS on its 32-bit word below: keep the two apart)
			const int i = base + lane;
			const uint32_t m = i < ncell ? zmin[i] : 0xFFFFu;
			const bool moves = m != 0xFFFFu && parent[i] != (uint32_t)i;
			__syncwarp();
			if (moves)
				atom_min_u16 (c, zminw, uf_find (parent, (uint32_t)i), m);
		}
		__syncwarp();
	}
	else
#endif
	{
	// (CTA per ROI / host emulation) every cell starts hooked to the head of its horizontal same-level run, then one union per
	// vertically adjacent pair of RUNS, then every cell's distance goes to its root
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const uint32_t lv = cell[i] & CELL_LEVEL;
		int hd = i;
		if (!(ibsi && !lv))
		{
			int x = i - (int)fw.div ((uint32_t)i) * w;
			while (x > 0 && (cell[hd - 1] & CELL_LEVEL) == lv)
			{
				hd--;
				x--;
			}
		}
		parent[i] = (uint32_t)hd;
	}
	for (int i = c.tid; i < (ncell + 1) / 2; i += c.nthr)
		zminw[i] = 0xFFFFFFFFu;
	c.sync();
	for (int i = c.tid + w; i < ncell; i += c.nthr)
	{
		const uint32_t lv = cell[i] & CELL_LEVEL;
		if ((ibsi && !lv) || (cell[i - w] & CELL_LEVEL) != lv)
			continue;
		const int x = i - (int)fw.div ((uint32_t)i) * w;
		if (x > 0 && (cell[i - 1] & CELL_LEVEL) == lv && (cell[i - w - 1] & CELL_LEVEL) == lv)
			continue;
		uf_unite (c, parent, (uint32_t)i, (uint32_t)(i - w));
	}
	c.sync();
	// ---- zone metric = min member distance
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		uint32_t lv = cell[i] & CELL_LEVEL;
		if (ibsi && !lv)
			continue;
		int y = (int)fw.div ((uint32_t)i), x = i - y * w;
		uint32_t root = uf_find (parent, (uint32_t)i);
		atom_min_u16 (c, zminw, root, (uint32_t)gldzm_dist2border (cell, w, h, x, y, !ibsi));
	}
	c.sync();

	}
	// ---- warp per ROI: the roots compacted into a list, in place over the parent array (dead from here on; see glszm_reduce)
	int nroot = -1;
#if NYX_DEVICE_BUILD
	if (c.wmode)
	{
		int run = 0;
		for (int base = 0; base < ncell; base += 32)
		{
			const int i = base + c.tid;
			const bool root = i < ncell && !(ibsi && !(cell[i] & CELL_LEVEL)) && parent[i] == (uint32_t)i;
			const unsigned b = __ballot_sync (0xffffffffu, root);
			__syncwarp();	// every lane has read its parent: the list may overwrite the parents of this chunk (a vote orders no memory)
			if (root)
				parent[run + __popc (b & ((1u << c.tid) - 1u))] = (uint32_t)i;
			run += __popc (b);	// (list positions never exceed the chunk's own cells: later chunks are still intact)
		}
		__syncwarp();
		nroot = run;
	}
#endif
	const int nloop = nroot >= 0 ? nroot : ncell;

	// ---- pass 1 over the zones (= roots)
	enum { Q_NS, Q_LDE, Q_ZDM, Q_GLM, Q_HGL, Q_N };
	unsigned long long q[Q_N];
	for (int k = 0; k < Q_N; k++)
		q[k] = 0;
	enum { A_SDE, A_LGL, A_SDL, A_SDH, A_LDL, A_LDH, A_N };
	double s[A_N];
	for (int k = 0; k < A_N; k++)
		s[k] = 0;
	for (int k = c.tid; k < nloop; k += c.nthr)
	{
		const int i = nroot >= 0 ? (int)parent[k] : k;
		const uint32_t lv = cell[i] & CELL_LEVEL;
		if (nroot < 0 && ((ibsi && !lv) || parent[i] != (uint32_t)i))
			continue;
		const uint32_t dd = zmin[i];
		P.inc (c, (int)(lv * (uint32_t)Nd + dd));
		atom_add (c, Ri + lv, 1u);
		atom_add (c, Cd + dd, 1u);
		if (dbg && dbg->keys)
			dbg->keys[atom_inc_ret (c, dbg->nkeys)] = (lv << 16) | dd;
		q[Q_NS] += 1;
		q[Q_LDE] += (unsigned long long)dd * dd;			// gldzm.cpp:403
		q[Q_ZDM] += dd;							// :445
		q[Q_GLM] += lv;							// :444
		q[Q_HGL] += (unsigned long long)lv * lv;			// :422
		emphasis6_int (s, lv, dd);	// :402, 421, 440-443
	}
	c.sync();
	cta_sum_u64<Q_N> (c, q);
	cta_sum<A_N> (c, s);
	const double Ns = (double)q[Q_NS], Nv = (double)r.area;
	const double glm = (double)q[Q_GLM] / Ns, zdm = (double)q[Q_ZDM] / Ns;

	// ---- pass 2: terms that need the counts (owner computes)
	unsigned long long q2[2] = { 0, 0 };	// sum n_level^2 (:423), sum n_metric^2 (:404)
	double t[3] = { 0, 0, 0 };		// GLV, ZDV, ZDE
	for (int k = c.tid; k < nloop; k += c.nthr)
	{
		const int i = nroot >= 0 ? (int)parent[k] : k;
		const uint32_t lv = cell[i] & CELL_LEVEL;
		if (nroot < 0 && ((ibsi && !lv) || parent[i] != (uint32_t)i))
			continue;
		const uint32_t dd = zmin[i];
		q2[0] += Ri[lv];
		q2[1] += Cd[dd];
		const double di = (double)lv - glm, dj = (double)dd - zdm;
		t[0] += di * di;						// :468-471
		t[1] += dj * dj;						// :474-476
		t[2] += ln_count (P.get ((int)(lv * (uint32_t)Nd + dd)));			// :446  log2(P/Ns + eps), summed as ln(P) and finished below
	}
	cta_sum_u64<2> (c, q2);
	cta_sum<3> (c, t);
	c.sync();
	// ---- pass 3: restore the all-zero region
	if (gldzm_p_words (Lmax, Nd) + (Lmax + 1) + Nd <= (size_t)ncell * 2)	// small matrix: wipe P, Ri, Cd (contiguous) with 128-bit stores
		wipe_words (c, mem, gldzm_p_words (Lmax, Nd) + (Lmax + 1) + Nd);
	else
	{
		for (int k = c.tid; k < nloop; k += c.nthr)
		{
			const int i = nroot >= 0 ? (int)parent[k] : k;
			const uint32_t lv = cell[i] & CELL_LEVEL;
			if (nroot < 0 && ((ibsi && !lv) || parent[i] != (uint32_t)i))
				continue;
			P.clear ((int)(lv * (uint32_t)Nd + zmin[i]));
		}
		for (int k = c.tid; k < Lmax + 1 + Nd; k += c.nthr)
			Ri[k] = 0;
	}
	c.sync();
	if (c.tid == 0)
	{
		const double inv_ns = 1.0 / Ns;		// one reciprocal instead of 17 divisions (features are held to 1e-9, not bit-exact)
		const double zdnu = (double)q2[1] * inv_ns, glnu = (double)q2[0] * inv_ns;
		out[Z_SDE] = s[A_SDE] * inv_ns;
		out[Z_LDE] = (double)q[Q_LDE] * inv_ns;
		out[Z_LGLZE] = s[A_LGL] * inv_ns;
		out[Z_HGLZE] = (double)q[Q_HGL] * inv_ns;
		out[Z_SDLGLE] = s[A_SDL] * inv_ns;
		out[Z_SDHGLE] = s[A_SDH] * inv_ns;
		out[Z_LDLGLE] = s[A_LDL] * inv_ns;
		out[Z_LDHGLE] = s[A_LDH] * inv_ns;
		out[Z_GLNU] = glnu;
		out[Z_GLNUN] = glnu * inv_ns;					// :428
		out[Z_ZDNU] = zdnu;
		out[Z_ZDNUN] = zdnu * inv_ns;					// :411
		out[Z_ZP] = Ns / Nv;						// :455
		out[Z_GLM] = glm;
		out[Z_GLV] = t[0] * inv_ns;
		out[Z_ZDM] = zdm;
		out[Z_ZDV] = t[1] * inv_ns;
		out[Z_ZDE] = -((t[2] - Ns * log (Ns)) * INV_LN2 * inv_ns);
	}
}

} // namespace nyx
