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
// B200 design: one CTA per ROI, labels in shared memory; GLDZM components by lock-free union-find (atomicMin
// hooking); sparse (level,minor) histograms by an in-smem bitonic sort of 32-bit keys -- independent of the number
// of grey levels, deterministic.
#pragma once
#include "nyx_common.h"
#include "nyx_features.h"

namespace nyx
{

constexpr uint32_t KEY_PAD = 0xFFFFFFFFu;

// visit every distinct key of a sorted list: f(major = key>>16, minor = key&0xFFFF, count)
template <class F> NYX_D void for_each_unique_key (const Cta& c, const uint32_t* keys, uint32_t n, F f)
{
	for (uint32_t i = c.tid; i < n; i += c.nthr)
	{
		uint32_t k = keys[i];
		if (i && keys[i - 1] == k)
			continue;
		uint32_t j = i + 1;
		while (j < n && keys[j] == k) j++;
		f (k >> 16, k & 0xFFFFu, j - i);
	}
}

// visit every distinct major (level) of a sorted list: f(level, number of list entries with that level)
template <class F> NYX_D void for_each_major (const Cta& c, const uint32_t* keys, uint32_t n, F f)
{
	for (uint32_t i = c.tid; i < n; i += c.nthr)
	{
		uint32_t m = keys[i] >> 16;
		if (i && (keys[i - 1] >> 16) == m)
			continue;
		uint32_t j = i + 1;
		while (j < n && (keys[j] >> 16) == m) j++;
		f (m, j - i);
	}
}

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

// ------------------------------------------------------------------------------------------------
// GLSZM
// ------------------------------------------------------------------------------------------------
NYX_HD size_t glszm_smem_words (int ncell)
{
	return (size_t)ncell * 2 + next_pow2 ((uint32_t)ncell) + (size_t)ncell + 1 + 4;
}

struct ZoneDebug
{
	uint32_t* keys;		// sorted (level<<16 | minor) keys of this ROI's zones, capacity = ncell
	uint32_t* nkeys;	// number of keys
};

NYX_D void glszm_body (const Cta& c, const RoiRec& r, const Params& prm, const uint32_t* cell, uint32_t* mem,
	double* out, const ZoneDebug* dbg)
{
	const int w = r.w, h = r.h, ncell = w * h;
	if (r.imin == r.imax)
	{
		for (int i = c.tid; i < NCOL_GLSZM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	const uint32_t npad = next_pow2 ((uint32_t)ncell);
	uint32_t* lab = mem;				// [ncell] zone label = raster index of the seed
	uint32_t* zs = lab + ncell;			// [ncell] zone size at the seed index
	uint32_t* keys = zs + ncell;			// [npad]
	uint32_t* sjh = keys + npad;			// [ncell+1] zones per size
	uint32_t* cnt = sjh + ncell + 1;		// [0] number of zones

	for (uint32_t i = c.tid; i < npad; i += c.nthr)
		keys[i] = KEY_PAD;
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		zs[i] = 0;
		sjh[i] = 0;
	}
	if (c.tid == 0)
	{
		sjh[ncell] = 0;
		cnt[0] = 0;
	}
	c.sync();

	// ---- forward-reachability labels, row by row
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
			lab[i] = m;
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
				lab[i + (x2 - x)] = cur;
			}
		}
		c.sync();
	}

	// ---- zone sizes, Np
	unsigned long long np[1] = { 0 };
	for (int i = c.tid; i < ncell; i += c.nthr)
		if (cell[i] & CELL_LEVEL)
		{
			np[0]++;
			atom_add (c, zs + lab[i], 1u);
		}
	c.sync();
	cta_sum_u64<1> (c, np);

	// ---- zone list -> keys (level<<16 | size), size histogram
	for (int i = c.tid; i < ncell; i += c.nthr)
		if ((cell[i] & CELL_LEVEL) && lab[i] == (uint32_t)i)
		{
			uint32_t sz = zs[i];
			uint32_t slot = atom_inc_ret (c, cnt);
			keys[slot] = ((cell[i] & CELL_LEVEL) << 16) | sz;
			atom_add (c, sjh + sz, 1u);
		}
	c.sync();
	const uint32_t nz = cnt[0];
	bitonic_sort (c, keys, npad);
	if (dbg && dbg->keys)
	{
		for (uint32_t i = c.tid; i < nz; i += c.nthr)
			dbg->keys[i] = keys[i];
		if (c.tid == 0)
			*dbg->nkeys = nz;
	}
	if (nz == 0)	// glszm.cpp:244-248
	{
		for (int i = c.tid; i < NCOL_GLSZM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	const double sum = (double)nz;

	// ---- integer sums
	enum { Q_GLN, Q_SZN, Q_LAE, Q_MUJ, Q_MUI, Q_HGL, Q_N };
	unsigned long long q[Q_N];
	for (int k = 0; k < Q_N; k++)
		q[k] = 0;
	for (int j = c.tid + 1; j <= ncell; j += c.nthr)
	{
		unsigned long long v = sjh[j];
		q[Q_SZN] += v * v;						// :481-485
		q[Q_LAE] += v * (unsigned long long)j * (unsigned long long)j;	// :437-440
		q[Q_MUJ] += v * (unsigned long long)j;				// :380
	}
	for_each_major (c, keys, nz, [&](uint32_t lv, uint32_t n) {
		q[Q_GLN] += (unsigned long long)n * n;				// :450-454
		q[Q_MUI] += (unsigned long long)n * lv;				// :381
		q[Q_HGL] += (unsigned long long)n * lv * lv;			// :574-578
	});
	cta_sum_u64<Q_N> (c, q);
	const double mui = (double)q[Q_MUI] / sum, muj = (double)q[Q_MUJ] / sum;

	// ---- FP sums
	enum { A_SAE, A_GLV, A_ZV, A_ZE, A_LGL, A_SAL, A_SAH, A_LAL, A_LAH, A_N };
	double s[A_N];
	for (int k = 0; k < A_N; k++)
		s[k] = 0;
	for (int j = c.tid + 1; j <= ncell; j += c.nthr)
	{
		uint32_t v = sjh[j];
		if (!v)
			continue;
		s[A_SAE] += (double)v / (double)(j * j);			// :424-427 (int j*j)
		double dj = (double)j - muj;
		s[A_ZV] += (double)v / sum * (dj * dj);				// :535-542
	}
	for_each_major (c, keys, nz, [&](uint32_t lv, uint32_t n) {
		double in = (double)lv, di = in - mui;
		s[A_GLV] += (double)n / sum * (di * di);			// :518-526
		s[A_LGL] += (double)n / (in * in);				// :559-563
	});
	for_each_unique_key (c, keys, nz, [&](uint32_t lv, uint32_t sz, uint32_t n) {
		double p = (double)n, i2 = (double)lv * (double)lv, j2 = (double)sz * (double)sz;
		s[A_LAH] += p * i2 * j2;					// :372-375
		s[A_LAL] += p * j2 / i2;
		s[A_SAH] += p * i2 / j2;
		s[A_SAL] += p / (i2 * j2);
		double pn = p / sum;
		s[A_ZE] += dmul (pn, ddiv (fast_log10 (dadd (pn, EPS_ZONES)), LOG10_2));	// :377-378
	});
	cta_sum<A_N> (c, s);
	if (c.tid == 0)
	{
		out[S_SAE] = s[A_SAE] / sum;
		out[S_LAE] = (double)q[Q_LAE] / sum;
		out[S_GLN] = (double)q[Q_GLN] / sum;
		out[S_GLNN] = (double)q[Q_GLN] / (sum * sum);
		out[S_SZN] = (double)q[Q_SZN] / sum;
		out[S_SZNN] = (double)q[Q_SZN] / (sum * sum);
		out[S_ZP] = sum / (double)np[0];				// :507-511
		out[S_GLV] = s[A_GLV];
		out[S_ZV] = s[A_ZV];
		out[S_ZE] = -s[A_ZE];
		out[S_LGLZE] = s[A_LGL] / sum;
		out[S_HGLZE] = (double)q[Q_HGL] / sum;
		out[S_SALGLE] = s[A_SAL] / sum;
		out[S_SAHGLE] = s[A_SAH] / sum;
		out[S_LALGLE] = s[A_LAL] / sum;
		out[S_LAHGLE] = s[A_LAH] / sum;
	}
}

// ------------------------------------------------------------------------------------------------
// GLDZM
// ------------------------------------------------------------------------------------------------
NYX_HD size_t gldzm_smem_words (int ncell, int maxwh)
{
	return (size_t)ncell * 2 + next_pow2 ((uint32_t)ncell) + (size_t)maxwh + 2 + 4;
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

NYX_D void gldzm_body (const Cta& c, const RoiRec& r, const Params& prm, const uint32_t* cell, uint32_t* mem,
	double* out, const ZoneDebug* dbg)
{
	const int w = r.w, h = r.h, ncell = w * h;
	const int maxwh = w > h ? w : h;
	if (r.imin == r.imax)	// gldzm.cpp:266-289
	{
		for (int i = c.tid; i < NCOL_GLDZM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	const bool ibsi = prm.ibsi != 0;
	const uint32_t npad = next_pow2 ((uint32_t)ncell);
	uint32_t* parent = mem;				// [ncell]
	uint32_t* zmin = parent + ncell;		// [ncell] zone metric at the root
	uint32_t* keys = zmin + ncell;			// [npad]
	uint32_t* sdh = keys + npad;			// [maxwh+2] zones per distance
	uint32_t* cnt = sdh + maxwh + 2;

	for (uint32_t i = c.tid; i < npad; i += c.nthr)
		keys[i] = KEY_PAD;
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		parent[i] = (uint32_t)i;
		zmin[i] = 0x7FFFFFFFu;
	}
	for (int i = c.tid; i < maxwh + 2; i += c.nthr)
		sdh[i] = 0;
	if (c.tid == 0)
		cnt[0] = 0;
	c.sync();

	// ---- 4-connected components of equal level
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		uint32_t lv = cell[i] & CELL_LEVEL;
		if (ibsi && !lv)
			continue;
		int y = i / w, x = i - y * w;
		if (x > 0 && (cell[i - 1] & CELL_LEVEL) == lv)
			uf_unite (c, parent, (uint32_t)i, (uint32_t)(i - 1));
		if (y > 0 && (cell[i - w] & CELL_LEVEL) == lv)
			uf_unite (c, parent, (uint32_t)i, (uint32_t)(i - w));
	}
	c.sync();
	// ---- zone metric = min member distance
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		uint32_t lv = cell[i] & CELL_LEVEL;
		if (ibsi && !lv)
			continue;
		int y = i / w, x = i - y * w;
		uint32_t root = uf_find (parent, (uint32_t)i);
		atom_min (c, zmin + root, (uint32_t)gldzm_dist2border (cell, w, h, x, y, !ibsi));
	}
	c.sync();
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		uint32_t lv = cell[i] & CELL_LEVEL;
		if ((ibsi && !lv) || parent[i] != (uint32_t)i)
			continue;
		uint32_t d = zmin[i];
		uint32_t slot = atom_inc_ret (c, cnt);
		keys[slot] = (lv << 16) | d;
		atom_add (c, sdh + d, 1u);
	}
	c.sync();
	const uint32_t nz = cnt[0];
	bitonic_sort (c, keys, npad);
	if (dbg && dbg->keys)
	{
		for (uint32_t i = c.tid; i < nz; i += c.nthr)
			dbg->keys[i] = keys[i];
		if (c.tid == 0)
			*dbg->nkeys = nz;
	}
	const double Ns = (double)nz, Nv = (double)r.area;

	enum { Q_ZDNU, Q_LDE, Q_ZDM, Q_GLNU, Q_GLM, Q_HGL, Q_N };
	unsigned long long q[Q_N];
	for (int k = 0; k < Q_N; k++)
		q[k] = 0;
	enum { A_SDE, A_LGL, A_SDL, A_SDH, A_LDL, A_LDH, A_ZDE, A_N };
	double s[A_N];
	for (int k = 0; k < A_N; k++)
		s[k] = 0;
	for (int d = c.tid + 1; d <= maxwh + 1; d += c.nthr)
	{
		unsigned long long v = sdh[d];
		if (!v)
			continue;
		q[Q_ZDNU] += v * v;						// gldzm.cpp:404
		q[Q_LDE] += v * (unsigned long long)d * (unsigned long long)d;	// :403
		q[Q_ZDM] += v * (unsigned long long)d;				// :445
		s[A_SDE] += (double)v / (double)d / (double)d;			// :402
	}
	for_each_major (c, keys, nz, [&](uint32_t lv, uint32_t n) {
		q[Q_GLNU] += (unsigned long long)n * n;				// :423
		q[Q_GLM] += (unsigned long long)n * lv;				// :444
		q[Q_HGL] += (unsigned long long)n * lv * lv;			// :422
		double g = (double)lv;
		s[A_LGL] += (double)n / (g * g);				// :421
	});
	for_each_unique_key (c, keys, nz, [&](uint32_t lv, uint32_t dd, uint32_t n) {
		double p = (double)n, g = (double)lv, d = (double)dd;
		s[A_SDL] += p / g / g / d / d;					// :440
		s[A_SDH] += g * g * p / d / d;					// :441
		s[A_LDL] += d * d * p / g / g;					// :442
		s[A_LDH] += g * g * d * d * p;					// :443
		s[A_ZDE] += p / Ns * log2 (p / Ns + EPS_ZONES);			// :446
	});
	cta_sum_u64<Q_N> (c, q);
	cta_sum<A_N> (c, s);
	const double glm = (double)q[Q_GLM] / Ns, zdm = (double)q[Q_ZDM] / Ns;
	double v2[2] = { 0, 0 };
	for_each_major (c, keys, nz, [&](uint32_t lv, uint32_t n) {
		double dif = (double)lv - glm;
		v2[0] += dif * dif * ((double)n / Ns);				// :468-471
	});
	for (int d = c.tid + 1; d <= maxwh + 1; d += c.nthr)
	{
		uint32_t v = sdh[d];
		if (!v)
			continue;
		double dif = (double)d - zdm;
		v2[1] += dif * dif * ((double)v / Ns);				// :474-476
	}
	cta_sum<2> (c, v2);
	if (c.tid == 0)
	{
		double zdnu = (double)q[Q_ZDNU] / Ns, glnu = (double)q[Q_GLNU] / Ns;
		out[Z_SDE] = s[A_SDE] / Ns;
		out[Z_LDE] = (double)q[Q_LDE] / Ns;
		out[Z_LGLZE] = s[A_LGL] / Ns;
		out[Z_HGLZE] = (double)q[Q_HGL] / Ns;
		out[Z_SDLGLE] = s[A_SDL] / Ns;
		out[Z_SDHGLE] = s[A_SDH] / Ns;
		out[Z_LDLGLE] = s[A_LDL] / Ns;
		out[Z_LDHGLE] = s[A_LDH] / Ns;
		out[Z_GLNU] = glnu;
		out[Z_GLNUN] = glnu / Ns;					// :428
		out[Z_ZDNU] = zdnu;
		out[Z_ZDNUN] = zdnu / Ns;					// :411
		out[Z_ZP] = Ns / Nv;						// :455
		out[Z_GLM] = glm;
		out[Z_GLV] = v2[0];
		out[Z_ZDM] = zdm;
		out[Z_ZDV] = v2[1];
		out[Z_ZDE] = -s[A_ZDE];
	}
}

} // namespace nyx
