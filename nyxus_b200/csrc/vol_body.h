// nyxus_b200 -- per-ROI 3D (Nyxus3D) families on the ROI's AABB cube: GLSZM and NGTDM.  (3D GLCM lives in glcm_body.h,
// 3D first-order re-uses intensity_body.h on the voxel cloud.)
//
// 3D GLSZM (reference features/3d_glszm.cpp):
//   blank test     :476-480   aux_min == aux_max -> 16 x softNaN
//   binning        :490-494   GLSZM_GREYDEPTH (0 unless set through "3glszm/greydepth": no binning), IBSI forces 0
//   background     :515       zeroI = 1 under MATLAB binning (background 0 is binned to level 1, so GENUINE level-1 voxels are
//                             dropped as well), else 0
//   zones          :44-469    flood over all 26 neighbours with back-tracking = 26-connected components of equal level
//   Ns = max zone size, Nz = #zones, Np = voxel-cloud size :527-534;  features :583-860 (SAE / LAE as direct double loops)
// 3D NGTDM (reference features/3d_ngtdm.cpp):
//   binning        :195-198   NGTDM_GREYDEPTH (0 unless set), IBSI forces 0
//   levels         :200-225   U = unique binned values of the WHOLE cube (background included); I = 0..max (IBSI flag) or
//                             sorted U; if I[0] == 0 every level and every voxel is shifted by +1
//   centres        :68-121    voxels whose (shifted) level != zeroI; neighbourhood = Chebyshev cube of NGTDM_RADIUS over all
//                             in-cube voxels INCLUDING background; nd == 0 -> dropped
//   |I| < 2 -> 5 x softNaN :229-237;  features :395-534
//
// B200 design: one CTA per ROI, cube in shared memory.  Components by lock-free union-find on 16-bit parents (a cube of
// the supported size class has < 65536 cells), every cell first hooked to the head of its x-run; reducers shared with 2D.
#pragma once
#include "nyx_common.h"
#include "nyx_features.h"
#include "zones_body.h"
#include "nbr_body.h"

namespace nyx
{

// ---- union-find on 16-bit parents (atomic min through a 16-bit compare-and-swap)
NYX_D uint32_t uf16_find (const volatile uint16_t* parent, uint32_t x)
{
	uint32_t p = parent[x];
	while (p != x)
	{
		x = p;
		p = parent[x];
	}
	return x;
}
NYX_D uint32_t atom_min16_ret (const Cta&, uint16_t* p, uint32_t v)
{
#if NYX_DEVICE_BUILD
	unsigned short old = *(volatile unsigned short*)p;
	while (v < old)
	{
		const unsigned short seen = atomicCAS ((unsigned short*)p, old, (unsigned short)v);
		if (seen == old)
			break;
		old = seen;
	}
	return old;
#else
	uint32_t o = *p;
	if (v < o) *p = (uint16_t)v;
	return o;
#endif
}
NYX_D void uf16_unite (const Cta& c, uint16_t* parent, uint32_t a, uint32_t b)
{
	for (;;)
	{
		a = uf16_find (parent, a);
		b = uf16_find (parent, b);
		if (a == b)
			return;
		if (a < b)
		{
			uint32_t t = a;
			a = b;
			b = t;
		}
		uint32_t old = atom_min16_ret (c, parent + a, b);	// hook the larger root under the smaller
		if (old == a)
			return;
		a = old;
	}
}

// region: hash keys[slots] + 16-bit cnts[slots] (zero on entry/exit) + parent/label (u16 per cell) + zone sizes (u16 per cell)
NYX_HD size_t glszm3d_smem_words (uint32_t maxcells, uint32_t slots)
{
	return glszm_smem_words (maxcells, slots);
}

// cell: cube binned with the GLSZM grey depth.  matlab: the binning was MATLAB (zeroI = 1).
NYX_D void glszm3d_body (const Cta& c, const RoiRec& r, const Params& prm, const uint32_t* cell, uint32_t* mem, uint32_t maxcells,
	uint32_t slots, bool matlab, double* out)
{
	const int w = r.w, h = r.h, d = r.d, ncell = w * h * d;
	if (r.imin == r.imax)
	{
		for (int i = c.tid; i < NCOL_GLSZM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	uint32_t* keys = mem;
	uint32_t* cnts = keys + slots;
	uint16_t* lab = (uint16_t*)(cnts + slots / 2);
	uint32_t* zsw = cnts + slots / 2 + ((size_t)maxcells + 1) / 2;
	const uint32_t mask = slots - 1;
	const uint32_t skip = matlab ? 1u : 0u;
	FastDiv fw, fh;
	fw.init ((uint32_t)w);
	fh.init ((uint32_t)h);

	for (int i = c.tid; i < (ncell + 1) / 2; i += c.nthr)
		zsw[i] = 0;
	// every cell hooked to the head of its same-level run along x
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const uint32_t lv = cell[i] & CELL_LEVEL;
		int hd = i;
		if (lv != skip)
		{
			int x = i - (int)fw.div ((uint32_t)i) * w;
			while (x > 0 && (cell[hd - 1] & CELL_LEVEL) == lv)
			{
				hd--;
				x--;
			}
		}
		lab[i] = (uint16_t)hd;
	}
	c.sync();
	// unions with the 12 remaining "backward" neighbours (previous rows of this slice, all of the previous slice)
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const uint32_t lv = cell[i] & CELL_LEVEL;
		if (lv == skip)
			continue;
		const int row = (int)fw.div ((uint32_t)i), x = i - row * w;
		const int z = (int)fh.div ((uint32_t)row), y = row - z * h;
		for (int dz = -1; dz <= 0; dz++)
		{
			const int zz = z + dz;
			if (zz < 0)
				continue;
			for (int dy = -1; dy <= (dz < 0 ? 1 : -1); dy++)
			{
				const int yy = y + dy;
				if (yy < 0 || yy >= h)
					continue;
				for (int dx = -1; dx <= 1; dx++)
				{
					const int xx = x + dx;
					if (xx < 0 || xx >= w)
						continue;
					const int j = (zz * h + yy) * w + xx;
					if ((cell[j] & CELL_LEVEL) == lv)
						uf16_unite (c, lab, (uint32_t)i, (uint32_t)j);
				}
			}
		}
	}
	c.sync();
	// flatten: lab[i] = root, staged through the (still all-zero) zone-size array so that no thread rewrites a parent
	// another thread may be chasing: reads of this phase see only the final forest (round 1's in-place rewrite was
	// correct -- a root keeps pointing to itself and parents only ever move towards it -- but it was a read/write
	// hazard in compute-sanitizer's racecheck).  Every thread stages and copies back its own cells.
	{
		uint16_t* stage = (uint16_t*)zsw;
		for (int i = c.tid; i < ncell; i += c.nthr)
			if ((cell[i] & CELL_LEVEL) != skip)
				stage[i] = (uint16_t)uf16_find (lab, (uint32_t)i);
		c.sync();
		for (int i = c.tid; i < ncell; i += c.nthr)
			if ((cell[i] & CELL_LEVEL) != skip)
			{
				lab[i] = stage[i];
				stage[i] = 0;
			}
		c.sync();
	}
	glszm_reduce (c, prm, cell, keys, cnts, mask, lab, zsw, ncell, skip, (double)r.area, out, nullptr);
}

// ------------------------------------------------------------------------------------------------
// 3D NGTDM, Chebyshev radius R (3d_ngtdm.cpp:69-94: the (2R+1)^3 - 1 in-cube neighbours, background included).
// L1 = level slots = (largest binned level) + 2 (room for the +1 shift)
// ------------------------------------------------------------------------------------------------
constexpr int NGTDM3_RMAX = 4;
NYX_HD int ngtdm3_ndmax (int R)
{
	const int s = 2 * (R < 1 ? 1 : R) + 1;
	return s * s * s - 1;
}
NYX_HD size_t ngtdm3d_smem_words (int Lmax, int R)
{
	return ngtdm_words (Lmax + 2, ngtdm3_ndmax (R)) + 2;
}

// cell: cube binned with the NGTDM grey depth; matlab: the binning was MATLAB (zeroI = 1); maxlevel: largest level in the cube
NYX_D void ngtdm3d_body (const Cta& c, const RoiRec& r, const Params& prm, const uint32_t* cell, uint32_t* mem, int Lmax,
	bool matlab, double* out)
{
	const int w = r.w, h = r.h, d = r.d, ncell = w * h * d;
	const int L1 = Lmax + 2;
	const int R = prm.ngtdm3_radius;
	if (R == 0)
	{
		// the default radius (env_features.cpp leaves NGTDM_RADIUS at 0): no voxel has a neighbour, the matrix is empty and
		// every feature is NaN (or softNaN when |I| < 2) -> `noval` at output either way (output_2_buffer.cpp:577)
		for (int i = c.tid; i < NCOL_NGTDM; i += c.nthr)
			out[i] = nan ("");
		return;
	}
	const int NDMAX = ngtdm3_ndmax (R);
	uint32_t* Nn = mem;
	uint32_t* SI = Nn + L1;
	uint32_t* PR = SI + (size_t)L1 * NDMAX;
	uint32_t* flags = mem + ngtdm_words (L1, NDMAX);	// [0] some cell has level 0, [1] largest level
	FastDiv fw, fh;
	fw.init ((uint32_t)w);
	fh.init ((uint32_t)h);

	// ---- pass 0: does the cube hold a zero level; largest level (3d_ngtdm.cpp:200-225)
	if (c.tid == 0)
	{
		flags[0] = 0;
		flags[1] = 0;
	}
	c.sync();
	{
		uint32_t z0 = 0, mx = 0;
		for (int i = c.tid; i < ncell; i += c.nthr)
		{
			const uint32_t lv = cell[i] & CELL_LEVEL;
			z0 |= lv == 0 ? 1u : 0u;
			mx = lv > mx ? lv : mx;
		}
		if (z0)
			flags[0] = 1;
		atom_max (c, flags + 1, mx);
	}
	c.sync();
	const bool ibsi = prm.ibsi != 0;
	const uint32_t shift = (ibsi || flags[0]) ? 1u : 0u;	// IBSI: I = 0..max so I[0] == 0 always
	const uint32_t maxlv = flags[1];
	const uint32_t zeroI = matlab ? 1u : 0u;
	c.sync();

	unsigned long long q[1] = { 0 };
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const uint32_t lv = (cell[i] & CELL_LEVEL) + shift;
		PR[lv] = 1;	// U holds every value of the cube
		if (lv == zeroI)
			continue;
		const int row = (int)fw.div ((uint32_t)i), x = i - row * w;
		const int z = (int)fh.div ((uint32_t)row), y = row - z * h;
		int nd = 0;
		uint32_t sum = 0;
		for (int dz = -R; dz <= R; dz++)
		{
			const int zz = z + dz;
			if (zz < 0 || zz >= d)
				continue;
			for (int dy = -R; dy <= R; dy++)
			{
				const int yy = y + dy;
				if (yy < 0 || yy >= h)
					continue;
				for (int dx = -R; dx <= R; dx++)
				{
					const int xx = x + dx;
					if ((!dx && !dy && !dz) || xx < 0 || xx >= w)
						continue;
					sum += (cell[(zz * h + yy) * w + xx] & CELL_LEVEL) + shift;
					nd++;
				}
			}
		}
		if (nd > 0)
		{
			long long df = (long long)lv * nd - (long long)sum;
			atom_add (c, Nn + lv, 1u);
			atom_add (c, SI + (size_t)lv * NDMAX + (nd - 1), (uint32_t)(df < 0 ? -df : df));
			q[0]++;
		}
	}
	c.sync();
	cta_sum_u64<1> (c, q);
	// |I|: IBSI flag -> 0..max (max + 1 levels); else the number of distinct values = Ngp, found by the sweep (-1)
	const int isize = ibsi ? (int)maxlv + 1 : -1;
	ngtdm_reduce (c, prm, mem, L1, NDMAX, 0, (double)q[0], isize, false, out, nullptr);
}

// ------------------------------------------------------------------------------------------------
// 3D GLDM (reference features/3d_gldm.cpp): blank :61-78; binning with GLDM_GREYDEPTH ("3gldm/greydepth", 0 unless set = no
// binning; IBSI forces 0) :93-97; background level zeroI = 1 under MATLAB binning, else 0 :117; every voxel whose level is
// not zeroI is a centre, its dependence = 1 + the number of its 26 in-cube neighbours of the same level :120-152;
// matrix[level][dependence 1..27]; features :271-530 -- the 2D formulas over level VALUES (gldm_body) with Nd = 27.
// region: P[(Lmax+1)][28] + Si[Lmax+1] + Sj[28] (zero on entry and exit) + one byte per cell (its dependence)
// ------------------------------------------------------------------------------------------------
constexpr int GLDM3_ND = 28;
NYX_HD size_t gldm3d_smem_words (int Lmax, uint32_t maxcells)
{
	return (size_t)(Lmax + 1) * (GLDM3_ND + 1) + GLDM3_ND + ((size_t)maxcells + 3) / 4;
}

NYX_D void gldm3d_body (const Cta& c, const RoiRec& r, const Params& prm, const uint32_t* cell, uint32_t* mem, int Lmax,
	bool matlab, double* out)
{
	const int w = r.w, h = r.h, d = r.d, ncell = w * h * d;
	if (r.imin == r.imax)
	{
		for (int i = c.tid; i < NCOL_GLDM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	uint32_t* P = mem;					// [level][nd], nd in 1..27
	uint32_t* Si = P + (size_t)(Lmax + 1) * GLDM3_ND;	// [level]
	uint32_t* Sj = Si + (Lmax + 1);				// [nd]
	uint8_t* aux = (uint8_t*)(Sj + GLDM3_ND);		// [ncell]
	const uint32_t zeroI = matlab ? 1u : 0u;
	FastDiv fw, fh;
	fw.init ((uint32_t)w);
	fh.init ((uint32_t)h);

	enum { Q_NZ, Q_LDE, Q_MUI, Q_MUJ, Q_HGL, Q_N };
	unsigned long long q[Q_N];
	for (int k = 0; k < Q_N; k++)
		q[k] = 0;
	enum { A_SDE, A_LGL, A_SDL, A_SDH, A_LDL, A_LDH, A_N };
	double s[A_N];
	for (int k = 0; k < A_N; k++)
		s[k] = 0;
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const uint32_t lv = cell[i] & CELL_LEVEL;
		if (lv == zeroI)
			continue;
		const int row = (int)fw.div ((uint32_t)i), x = i - row * w;
		const int z = (int)fh.div ((uint32_t)row), y = row - z * h;
		int j = 1;
		for (int dz = -1; dz <= 1; dz++)
		{
			const int zz = z + dz;
			if (zz < 0 || zz >= d)
				continue;
			for (int dy = -1; dy <= 1; dy++)
			{
				const int yy = y + dy;
				if (yy < 0 || yy >= h)
					continue;
				for (int dx = -1; dx <= 1; dx++)
				{
					const int xx = x + dx;
					if ((!dx && !dy && !dz) || xx < 0 || xx >= w)
						continue;
					j += (cell[(zz * h + yy) * w + xx] & CELL_LEVEL) == lv ? 1 : 0;
				}
			}
		}
		aux[i] = (uint8_t)j;
		atom_add (c, P + (size_t)lv * GLDM3_ND + j, 1u);
		atom_add (c, Si + lv, 1u);
		atom_add (c, Sj + j, 1u);
		q[Q_NZ] += 1;
		q[Q_LDE] += (unsigned long long)(j * j);
		q[Q_MUJ] += (unsigned long long)j;
		q[Q_MUI] += lv;
		q[Q_HGL] += (unsigned long long)lv * (unsigned long long)lv;
		emphasis6_int (s, lv, (uint32_t)j);
	}
	c.sync();
	cta_sum_u64<Q_N> (c, q);
	cta_sum<A_N> (c, s);
	if (q[Q_NZ] == 0)	// 3d_gldm.cpp:196-213
	{
		for (int i = c.tid; i < NCOL_GLDM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	const double Nz = (double)q[Q_NZ];
	const double mui = (double)q[Q_MUI] / Nz, muj = (double)q[Q_MUJ] / Nz;
	const double inv_nz = 1.0 / Nz;
	unsigned long long q2[2] = { 0, 0 };
	double t[3] = { 0, 0, 0 };
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const uint32_t lv = cell[i] & CELL_LEVEL;
		if (lv == zeroI)
			continue;
		const uint32_t j = aux[i];
		q2[0] += Si[lv];
		q2[1] += Sj[j];
		const double di = (double)lv - mui, dj = (double)j - muj;
		t[0] += di * di;
		t[1] += dj * dj;
		const double p = (double)P[(size_t)lv * GLDM3_ND + j] * inv_nz;
		t[2] += fast_log10 (dadd (p, EPS_ZONES));
	}
	cta_sum_u64<2> (c, q2);
	cta_sum<3> (c, t);
	c.sync();
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const uint32_t lv = cell[i] & CELL_LEVEL;
		if (lv == zeroI)
			continue;
		P[(size_t)lv * GLDM3_ND + aux[i]] = 0;
		Si[lv] = 0;
	}
	for (int k = c.tid; k < GLDM3_ND; k += c.nthr)
		Sj[k] = 0;
	c.sync();
	if (c.tid == 0)
	{
		out[D_SDE] = s[A_SDE] / Nz;
		out[D_LDE] = (double)q[Q_LDE] / Nz;
		out[D_GLN] = (double)q2[0] / Nz;
		out[D_DN] = (double)q2[1] / Nz;
		out[D_DNN] = (double)q2[1] / (Nz * Nz);
		out[D_GLV] = t[0] / Nz;
		out[D_DV] = t[1] / Nz;
		out[D_DE] = -(t[2] / Nz) * (1.0 / LOG10_2);
		out[D_LGLE] = s[A_LGL] / Nz;
		out[D_HGLE] = (double)q[Q_HGL] / Nz;
		out[D_SDLGLE] = s[A_SDL] / Nz;
		out[D_SDHGLE] = s[A_SDH] / Nz;
		out[D_LDLGLE] = s[A_LDL] / Nz;
		out[D_LDHGLE] = s[A_LDH] / Nz;
	}
}

// ------------------------------------------------------------------------------------------------
// 3D NGLDM (reference features/3d_ngldm.cpp): blank :173-198; levels = Nyxus::to_grayscale (v, 0, aux_max, GREYDEPTH, ibsi) of
// EVERY cube voxel, background included (:88-103, 219-223) -- the cube passed here carries those levels (Binner::init_tg); the
// sorted unique levels are the matrix rows.  Centres: the voxels with all 26 cube neighbours in range (margins skipped, off-ROI
// centres NOT skipped :128-147); dependence = number of the 24 neighbours listed at :12-41 ((0,0,+-1) are not among them)
// with the centre's level; Nr = max dependence + 1 :167.  Features :262-355: Ns sums ALL columns, every other sum skips the
// zero-dependence column; GLV takes the 1-based ROW INDEX against a mean of level VALUES (:331), reproduced.
// region: M[(Lmax+1)][25] + rank[Lmax+1] + Sr[25] + 256 words of scan scratch (zero on entry and exit)
// ------------------------------------------------------------------------------------------------
constexpr int NGLDM3_NC = 25;
NYX_HD size_t ngldm3d_smem_words (int Lmax)
{
	return (size_t)(Lmax + 1) * (NGLDM3_NC + 1) + NGLDM3_NC + 256;
}

NYX_D void ngldm3d_body (const Cta& c, const RoiRec& r, const Params& prm, const uint32_t* cell, uint32_t* mem, int Lmax, double* out)
{
	const int w = r.w, h = r.h, d = r.d, ncell = w * h * d;
	if (r.imin == r.imax)
	{
		for (int i = c.tid; i < NCOL_NGLDM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	const int L1 = Lmax + 1;
	uint32_t* M = mem;					// [level][dependence 0..24]
	uint32_t* rank = M + (size_t)L1 * NGLDM3_NC;		// [level]: voxel count, then the 1-based row of the level (0 = absent)
	uint32_t* Sr = rank + L1;				// [dependence]: centres per dependence
	uint32_t* part = Sr + NGLDM3_NC;			// [nthr <= 256]
	FastDiv fw, fh;
	fw.init ((uint32_t)w);
	fh.init ((uint32_t)h);
	for (int i = c.tid; i < ncell; i += c.nthr)
		atom_add (c, rank + (cell[i] & CELL_LEVEL), 1u);
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const int row = (int)fw.div ((uint32_t)i), x = i - row * w;
		const int z = (int)fh.div ((uint32_t)row), y = row - z * h;
		if (x < 1 || x >= w - 1 || y < 1 || y >= h - 1 || z < 1 || z >= d - 1)
			continue;
		const uint32_t lv = cell[i] & CELL_LEVEL;
		int n = 0;
		for (int dz = -1; dz <= 1; dz++)
			for (int dy = -1; dy <= 1; dy++)
				for (int dx = -1; dx <= 1; dx++)
					if (dx || dy)
						n += (cell[((z + dz) * h + (y + dy)) * w + (x + dx)] & CELL_LEVEL) == lv ? 1 : 0;
		atom_add (c, M + (size_t)lv * NGLDM3_NC + n, 1u);
		atom_add (c, Sr + n, 1u);
	}
	c.sync();
	// rows: 1-based rank of every level that occurs in the cube
	const int chunk = (L1 + c.nthr - 1) / c.nthr, lo = c.tid * chunk, hi = lo + chunk < L1 ? lo + chunk : L1;
	uint32_t np = 0;
	for (int lv = lo; lv < hi; lv++)
		np += rank[lv] ? 1u : 0u;
	part[c.tid] = np;
	c.sync();
	uint32_t base = 0;
	for (int t = 0; t < c.tid; t++)
		base += part[t];
	for (int lv = lo; lv < hi; lv++)
		rank[lv] = rank[lv] ? ++base : 0u;
	int Nr = 1;
	unsigned long long ns_all = 0, dcnu = 0;
	for (int j = 0; j < NGLDM3_NC; j++)
	{
		const unsigned long long sj = Sr[j];
		ns_all += sj;
		if (sj)
			Nr = j + 1;
		if (j)
			dcnu += sj * sj;		// :317-328 (columns above the largest dependence are empty)
	}
	c.sync();
	const double Ns = (double)ns_all;
	enum { Q_HDE, Q_HGL, Q_GLM, Q_DCM, Q_ENE, Q_GLNU, Q_N };
	unsigned long long q[Q_N];
	for (int k = 0; k < Q_N; k++)
		q[k] = 0;
	enum { A_LDE, A_LGL, A_LDL, A_LDH, A_HDL, A_HDH, A_ENT, A_N };
	double s[A_N];
	for (int k = 0; k < A_N; k++)
		s[k] = 0;
	for (int lv = c.tid; lv < L1; lv += c.nthr)
	{
		if (!rank[lv])
			continue;
		const double g = (double)lv, g2 = g * g;
		unsigned long long srow = 0;
		for (int j = 1; j < Nr; j++)
		{
			const unsigned long long sij = M[(size_t)lv * NGLDM3_NC + j];
			if (!sij)
				continue;
			const double sd = (double)sij, jd = (double)j, kd = (double)(j + 1), pij = sd / Ns;
			srow += sij;
			q[Q_HDE] += sij * (unsigned long long)(j * j);
			q[Q_HGL] += sij * (unsigned long long)lv * (unsigned long long)lv;
			q[Q_GLM] += sij * (unsigned long long)lv;
			q[Q_DCM] += sij * (unsigned long long)(j + 1);
			q[Q_ENE] += sij * sij;
			s[A_LDE] += sd / jd / jd;
			if (lv)
			{
				s[A_LGL] += sd / g / g;
				s[A_LDL] += sd / jd / jd / g / g;
				s[A_HDL] += sd * kd * kd / g2;
			}
			s[A_LDH] += sd * g2 / kd / kd;
			s[A_HDH] += sd * kd * kd * g2;
			s[A_ENT] += pij * log (pij);
		}
		q[Q_GLNU] += srow * srow;
	}
	cta_sum_u64<Q_N> (c, q);
	cta_sum<A_N> (c, s);
	const double glm = (double)q[Q_GLM] / Ns, dcm = (double)q[Q_DCM] / Ns;
	double t[2] = { 0, 0 };
	for (int lv = c.tid; lv < L1; lv += c.nthr)
	{
		const uint32_t rk = rank[lv];
		if (!rk)
			continue;
		for (int j = 1; j < Nr; j++)
		{
			const uint32_t sij = M[(size_t)lv * NGLDM3_NC + j];
			if (!sij)
				continue;
			const double pij = (double)sij / Ns, di = (double)rk - glm, dj = (double)(j + 1) - dcm;
			t[0] += di * di * pij;
			t[1] += dj * dj * pij;
		}
	}
	cta_sum<2> (c, t);
	c.sync();
	wipe_words (c, mem, (size_t)L1 * (NGLDM3_NC + 1) + NGLDM3_NC);
	c.sync();
	if (c.tid == 0)
	{
		out[N_LDE] = s[A_LDE] / Ns;
		out[N_HDE] = (double)q[Q_HDE] / Ns;
		out[N_LGLCE] = s[A_LGL] / Ns;
		out[N_HGLCE] = (double)q[Q_HGL] / Ns;
		out[N_LDLGLE] = s[A_LDL] / Ns;
		out[N_LDHGLE] = s[A_LDH] / Ns;
		out[N_HDLGLE] = s[A_HDL] / Ns;
		out[N_HDHGLE] = s[A_HDH] / Ns;
		out[N_GLNU] = (double)q[Q_GLNU] / Ns;
		out[N_GLNUN] = (double)q[Q_GLNU] / (Ns * Ns);
		out[N_DCNU] = (double)dcnu / Ns;
		out[N_DCNUN] = (double)dcnu / (Ns * Ns);
		out[N_DCP] = 1.0;
		out[N_GLM] = ns_all ? glm : 0.0;
		out[N_GLV] = t[0];
		out[N_DCM] = ns_all ? dcm : 0.0;
		out[N_DCV] = t[1];
		out[N_DCENT] = -s[A_ENT] * INV_LN2;
		out[N_DCENE] = ns_all ? (double)q[Q_ENE] / (Ns * Ns) : 0.0;
	}
}

// ------------------------------------------------------------------------------------------------
// 3D GLDZM (reference features/3d_gldzm.cpp): blank :523-546; binned with the coarse grey depth (IBSI: none) :65-72; zones =
// 6-connected components of equal level (E,S,W,N,up,down with back-tracking :94-283), level 0 skipped in IBSI mode only; the
// distance of a voxel is taken IN ITS SLICE (dist2border looks along x and y only :330-375); zone metric = minimum over the
// members; features :412-519 as in 2D.  A negative (PyRadiomics) depth makes the reference index its matrix out of range
// (level-0 zones have no row :294-297) -- refused by the host.
// region: P[Lmax+1][Nd] + Ri[Lmax+1] + Cd[Nd] (zero on entry and exit) + parent[ncell] + zmin[ncell]
// ------------------------------------------------------------------------------------------------
NYX_HD size_t gldzm3d_smem_words (int Lmax, int Nd, unsigned long long maxcells)
{
	return (size_t)(Lmax + 1) * Nd + (Lmax + 1) + Nd + 2 * (size_t)maxcells;
}

NYX_D void gldzm3d_body (const Cta& c, const RoiRec& r, const Params& prm, const uint32_t* cell, uint32_t* mem, int Lmax, int Nd,
	unsigned long long maxcells, double* out)
{
	const int w = r.w, h = r.h, d = r.d, wh = w * h, ncell = wh * d;
	if (r.imin == r.imax)
	{
		for (int i = c.tid; i < NCOL_GLDZM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	const bool ibsi = prm.ibsi != 0;
	uint32_t* P = mem;					// [Lmax+1][Nd]
	uint32_t* Ri = P + (size_t)(Lmax + 1) * Nd;		// [Lmax+1]
	uint32_t* Cd = Ri + (Lmax + 1);				// [Nd]
	uint32_t* parent = Cd + Nd;				// [ncell]
	uint32_t* zmin = parent + maxcells;			// [ncell]
	FastDiv fw, fh;
	fw.init ((uint32_t)w);
	fh.init ((uint32_t)h);
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		parent[i] = (uint32_t)i;
		zmin[i] = 0xFFFFFFFFu;
	}
	c.sync();
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const uint32_t lv = cell[i] & CELL_LEVEL;
		if (ibsi && !lv)
			continue;
		const int row = (int)fw.div ((uint32_t)i), x = i - row * w;
		const int z = (int)fh.div ((uint32_t)row), y = row - z * h;
		if (x > 0 && (cell[i - 1] & CELL_LEVEL) == lv)
			uf_unite (c, parent, (uint32_t)i, (uint32_t)(i - 1));
		if (y > 0 && (cell[i - w] & CELL_LEVEL) == lv)
			uf_unite (c, parent, (uint32_t)i, (uint32_t)(i - w));
		if (z > 0 && (cell[i - wh] & CELL_LEVEL) == lv)
			uf_unite (c, parent, (uint32_t)i, (uint32_t)(i - wh));
	}
	c.sync();
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const uint32_t lv = cell[i] & CELL_LEVEL;
		if (ibsi && !lv)
			continue;
		const int row = (int)fw.div ((uint32_t)i), x = i - row * w;
		const int z = (int)fh.div ((uint32_t)row), y = row - z * h;
		const uint32_t root = uf_find (parent, (uint32_t)i);
		atom_min_ret (c, zmin + root, (uint32_t)gldzm_dist2border (cell + (size_t)z * wh, w, h, x, y, !ibsi));
	}
	c.sync();
	enum { Q_NS, Q_LDE, Q_ZDM, Q_GLM, Q_HGL, Q_N };
	unsigned long long q[Q_N];
	for (int k = 0; k < Q_N; k++)
		q[k] = 0;
	enum { A_SDE, A_LGL, A_SDL, A_SDH, A_LDL, A_LDH, A_N };
	double s[A_N];
	for (int k = 0; k < A_N; k++)
		s[k] = 0;
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const uint32_t lv = cell[i] & CELL_LEVEL;
		if ((ibsi && !lv) || parent[i] != (uint32_t)i)
			continue;
		const uint32_t dd = zmin[i];
		atom_add (c, P + (size_t)lv * Nd + dd, 1u);
		atom_add (c, Ri + lv, 1u);
		atom_add (c, Cd + dd, 1u);
		q[Q_NS] += 1;
		q[Q_LDE] += (unsigned long long)dd * dd;
		q[Q_ZDM] += dd;
		q[Q_GLM] += lv;
		q[Q_HGL] += (unsigned long long)lv * lv;
		emphasis6_int (s, lv, dd);
	}
	c.sync();
	cta_sum_u64<Q_N> (c, q);
	cta_sum<A_N> (c, s);
	const double Ns = (double)q[Q_NS], Nv = (double)r.area;
	const double glm = (double)q[Q_GLM] / Ns, zdm = (double)q[Q_ZDM] / Ns;
	unsigned long long q2[2] = { 0, 0 };
	double t[3] = { 0, 0, 0 };
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const uint32_t lv = cell[i] & CELL_LEVEL;
		if ((ibsi && !lv) || parent[i] != (uint32_t)i)
			continue;
		const uint32_t dd = zmin[i];
		q2[0] += Ri[lv];
		q2[1] += Cd[dd];
		const double di = (double)lv - glm, dj = (double)dd - zdm;
		t[0] += di * di;
		t[1] += dj * dj;
		t[2] += ln_count (P[(size_t)lv * Nd + dd]);
	}
	cta_sum_u64<2> (c, q2);
	cta_sum<3> (c, t);
	c.sync();
	const size_t mw = (size_t)(Lmax + 1) * Nd + (Lmax + 1) + Nd;
	if (mw <= (size_t)ncell * 4)		// small matrix: wipe P, Ri, Cd (contiguous)
		wipe_words (c, mem, mw);
	else
	{
		for (int i = c.tid; i < ncell; i += c.nthr)
		{
			const uint32_t lv = cell[i] & CELL_LEVEL;
			if ((ibsi && !lv) || parent[i] != (uint32_t)i)
				continue;
			P[(size_t)lv * Nd + zmin[i]] = 0;
			Ri[lv] = 0;
		}
		for (int k = c.tid; k < Nd; k += c.nthr)
			Cd[k] = 0;
	}
	c.sync();
	if (c.tid == 0)
	{
		const double zdnu = (double)q2[1] / Ns, glnu = (double)q2[0] / Ns;
		out[Z_SDE] = s[A_SDE] / Ns;
		out[Z_LDE] = (double)q[Q_LDE] / Ns;
		out[Z_LGLZE] = s[A_LGL] / Ns;
		out[Z_HGLZE] = (double)q[Q_HGL] / Ns;
		out[Z_SDLGLE] = s[A_SDL] / Ns;
		out[Z_SDHGLE] = s[A_SDH] / Ns;
		out[Z_LDLGLE] = s[A_LDL] / Ns;
		out[Z_LDHGLE] = s[A_LDH] / Ns;
		out[Z_GLNU] = glnu;
		out[Z_GLNUN] = glnu / Ns;
		out[Z_ZDNU] = zdnu;
		out[Z_ZDNUN] = zdnu / Ns;
		out[Z_ZP] = Ns / Nv;
		out[Z_GLM] = glm;
		out[Z_GLV] = t[0] / Ns;
		out[Z_ZDM] = zdm;
		out[Z_ZDV] = t[1] / Ns;
		out[Z_ZDE] = -((t[2] - Ns * log (Ns)) * INV_LN2 / Ns);
	}
}

// ------------------------------------------------------------------------------------------------
// 3D GLRLM (reference features/3d_glrlm.cpp): blank :113-136 (softNaN in every angle, so in the _AVE too); binned with
// GLRLM_GREYDEPTH ("3glrlm/greydepth", 0 unless set = no binning; IBSI forces 0) :148-152; 13 directions :17-32 (dz, dy, dx), all
// raster-forward, so the flood of :34-106 finds the maximal same-level segments of every line; level 0 neither starts nor joins
// a run; Np = the voxel-cloud size :196; 16 features per direction :211-226 with the 32-bit products of the 2D family
// (:556, 603, 627, 651, 675); the table keeps direction 0 only (save_value :266-281 stores vectors, the output takes [0])
// plus the 13-direction averages :284-299.
// region: P[Lmax][Nr] + Ri[Lmax] + Cj[Nr] (zero on entry and exit) + run length per cell
// ------------------------------------------------------------------------------------------------
constexpr int GLRLM3_NA = 13;
constexpr int NCOL_GLRLM3 = 2 * R_NANGLED;
NYX_HD size_t glrlm3d_smem_words (int Lmax, int Nr, unsigned long long maxcells)
{
	return (size_t)Lmax * Nr + Lmax + Nr + (size_t)maxcells;
}

NYX_D void glrlm3d_body (const Cta& c, const RoiRec& r, const Params& prm, const uint32_t* cell, uint32_t* mem, int Lmax, int Nr,
	double* out)
{
	const int w = r.w, h = r.h, d = r.d, wh = w * h, ncell = wh * d;
	if (r.imin == r.imax)
	{
		for (int i = c.tid; i < NCOL_GLRLM3; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	uint32_t* P = mem;				// [Lmax][Nr]
	uint32_t* Ri = P + (size_t)Lmax * Nr;		// [Lmax]
	uint32_t* Cj = Ri + Lmax;			// [Nr]
	uint32_t* rl = Cj + Nr;				// [ncell] length of the run starting at the cell (0: not a start)
	FastDiv fw, fh;
	fw.init ((uint32_t)w);
	fh.init ((uint32_t)h);
	double ave[R_NANGLED];
	for (int k = 0; k < R_NANGLED; k++)
		ave[k] = 0;
	for (int a = 0; a < GLRLM3_NA; a++)
	{
		// :17-32: dz = 1 for a < 9, dy = 1, 0, -1 by threes; then dz = 0: (dy, dx) = (1,1), (1,0), (1,-1), (0,1)
		const int dz = a < 9 ? 1 : 0;
		const int dy = a < 9 ? 1 - a / 3 : (a < 12 ? 1 : 0);
		const int dx = a < 9 ? 1 - a % 3 : (a < 12 ? 10 - a : 1);
		enum { Q_SUM, Q_LRE, Q_MUI, Q_MUJ, Q_HGL, Q_N };
		unsigned long long q[Q_N];
		for (int k = 0; k < Q_N; k++)
			q[k] = 0;
		enum { A_SRE, A_LGL, A_SRL, A_SRH, A_LRL, A_LRH, A_N };
		double s[A_N];
		for (int k = 0; k < A_N; k++)
			s[k] = 0;
		const int stride = dz * wh + dy * w + dx;
		for (int i = c.tid; i < ncell; i += c.nthr)
		{
			const uint32_t lv = cell[i] & CELL_LEVEL;
			int len = 0;
			if (lv)
			{
				const int row = (int)fw.div ((uint32_t)i), x = i - row * w;
				const int z = (int)fh.div ((uint32_t)row), y = row - z * h;
				const int px = x - dx, py = y - dy, pz = z - dz;
				if (!(px >= 0 && px < w && py >= 0 && py < h && pz >= 0 && (cell[i - stride] & CELL_LEVEL) == lv))
				{
					int xx = x + dx, yy = y + dy, zz = z + dz, j = i + stride;
					len = 1;
					while (xx >= 0 && xx < w && yy >= 0 && yy < h && zz < d && (cell[j] & CELL_LEVEL) == lv)
					{
						len++;
						xx += dx;
						yy += dy;
						zz += dz;
						j += stride;
					}
					atom_add (c, P + (size_t)(lv - 1) * Nr + (len - 1), 1u);
					atom_add (c, Ri + (lv - 1), 1u);
					atom_add (c, Cj + (len - 1), 1u);
					const uint32_t i2 = lv * lv, j2 = (uint32_t)(len * len), i2j2 = lv * lv * (uint32_t)len * (uint32_t)len;	// 32-bit wrap
					q[Q_SUM] += 1;
					q[Q_LRE] += (unsigned long long)len * (unsigned long long)len;
					q[Q_MUI] += lv;
					q[Q_MUJ] += (unsigned long long)len;
					q[Q_HGL] += (unsigned long long)lv * (unsigned long long)lv;	// :603 double(inten) * double(inten): no wrap
					if (lv < 65536u && (unsigned long long)i2 * (unsigned long long)j2 < 0x100000000ull)
						emphasis6_int (s, lv, (uint32_t)len);
					else
					{
						s[A_SRE] += 1.0 / (double)(len * len);
						s[A_LGL] += 1.0 / (double)i2;
						s[A_SRL] += 1.0 / (double)i2j2;
						s[A_SRH] += (double)i2 / (double)(int)j2;
						s[A_LRL] += (double)(int)j2 / (double)i2;
						s[A_LRH] += (double)i2j2;
					}
				}
			}
			rl[i] = (uint32_t)len;
		}
		c.sync();
		cta_sum_u64<Q_N> (c, q);
		cta_sum<A_N> (c, s);
		double f[R_NANGLED];
		for (int k = 0; k < R_NANGLED; k++)
			f[k] = 0.0;
		if (q[Q_SUM])		// no run: every feature 0 (:331-333 ...), RP = 0 / Np
		{
			const double sum = (double)q[Q_SUM], Np = (double)r.area;
			const double mui = (double)q[Q_MUI] / sum, muj = (double)q[Q_MUJ] / sum;
			const double inv_sum = 1.0 / sum;
			unsigned long long q2[2] = { 0, 0 };
			double t[3] = { 0, 0, 0 };
			for (int i = c.tid; i < ncell; i += c.nthr)
			{
				const int len = (int)rl[i];
				if (!len)
					continue;
				const uint32_t lv = cell[i] & CELL_LEVEL;
				q2[0] += Ri[lv - 1];
				q2[1] += Cj[len - 1];
				const double di = (double)lv - mui, dj = (double)len - muj;
				t[0] += di * di;
				t[1] += dj * dj;
				const double p = (double)P[(size_t)(lv - 1) * Nr + (len - 1)] * inv_sum;
				t[2] += fast_log10 (dadd (p, EPS_ZONES));
			}
			cta_sum_u64<2> (c, q2);
			cta_sum<3> (c, t);
			c.sync();
			const size_t mw = (size_t)Lmax * Nr + Lmax + Nr;
			if (mw <= (size_t)ncell * 4)		// small matrix: wipe P, Ri, Cj (contiguous)
				wipe_words (c, mem, mw);
			else
				for (int i = c.tid; i < ncell; i += c.nthr)
				{
					const int len = (int)rl[i];
					if (!len)
						continue;
					const uint32_t lv = cell[i] & CELL_LEVEL;
					P[(size_t)(lv - 1) * Nr + (len - 1)] = 0;
					Ri[lv - 1] = 0;
					Cj[len - 1] = 0;
				}
			f[R_SRE] = s[A_SRE] / sum;
			f[R_LRE] = (double)q[Q_LRE] / sum;
			f[R_GLN] = (double)q2[0] / sum;
			f[R_GLNN] = (double)q2[0] / (sum * sum);
			f[R_RLN] = (double)q2[1] / sum;
			f[R_RLNN] = (double)q2[1] / (sum * sum);
			f[R_RP] = sum / Np;
			f[R_GLV] = t[0] / sum;
			f[R_RV] = t[1] / sum;
			f[R_RE] = -(t[2] / sum) * (1.0 / LOG10_2);
			f[R_LGLRE] = s[A_LGL] / sum;
			f[R_HGLRE] = (double)q[Q_HGL] / sum;
			f[R_SRLGLE] = s[A_SRL] / sum;
			f[R_SRHGLE] = s[A_SRH] / sum;
			f[R_LRLGLE] = s[A_LRL] / sum;
			f[R_LRHGLE] = s[A_LRH] / sum;
		}
		c.sync();
		for (int k = 0; k < R_NANGLED; k++)
			ave[k] += f[k];
		if (a == 0 && c.tid == 0)
			for (int k = 0; k < R_NANGLED; k++)
				out[k] = f[k];
	}
	if (c.tid == 0)
		for (int k = 0; k < R_NANGLED; k++)
			out[R_NANGLED + k] = ave[k] / (double)GLRLM3_NA;
}

// ------------------------------------------------------------------------------------------------
// 3D GLSZM of a LARGE cube (beyond the shared-memory size classes: organ-sized masks, whole-volume ROIs).  Same
// semantics as glszm3d_body; everything lives in global memory: 32-bit union-find parents and zone sizes, and the
// sparse (level, size) matrix + marginals in one open-addressing table of 64-bit keys with 32-bit counts.
// SAE / LAE take `j * j` in int like the reference (3d_glszm.cpp:calc_SAE / calc_LAE: wraps for zones of >= 46341 voxels);
// the other emphasis sums use double j^2 (calc_sums_of_P).
// ------------------------------------------------------------------------------------------------
NYX_HD uint32_t glszm3d_big_slots (int Lmax, unsigned long long maxcells)
{
	// distinct sizes d: 1 + 2 + .. + d <= cells; distinct (level, size) keys <= min (cells, Lmax * d)
	unsigned long long d = 2;
	while (d * (d + 1) / 2 <= maxcells) d++;
	unsigned long long K = (unsigned long long)Lmax * d;
	if (K > maxcells) K = maxcells;
	K += (unsigned long long)Lmax + d + 2;
	K += K / 2;
	if (K > 0x40000000ull) K = 0x40000000ull;
	return next_pow2 ((uint32_t)K);
}
// words: parent[cells] + size[cells] + keys[slots] (64-bit) + cnts[slots]
NYX_HD size_t glszm3d_big_words (unsigned long long maxcells, uint32_t slots)
{
	return (size_t)(2 * maxcells + 3ull * slots + 2);
}

NYX_D void htab64_add (const Cta& c, unsigned long long* keys, uint32_t* cnts, uint32_t mask, unsigned long long key)
{
	uint32_t slot = htab_mix ((uint32_t)key ^ (uint32_t)(key >> 29)) & mask;
	for (;;)
	{
		unsigned long long k = ((volatile unsigned long long*)keys)[slot];
		if (k == 0)
		{
#if NYX_DEVICE_BUILD
			k = atomicCAS (keys + slot, 0ull, key);
#else
			keys[slot] = key;
#endif
		}
		if (k == 0 || k == key)
			break;
		slot = (slot + 1) & mask;
	}
	atom_add (c, cnts + slot, 1u);
}
NYX_D uint32_t htab64_get (const unsigned long long* keys, const uint32_t* cnts, uint32_t mask, unsigned long long key)
{
	uint32_t slot = htab_mix ((uint32_t)key ^ (uint32_t)(key >> 29)) & mask;
	while (keys[slot] != key)
		slot = (slot + 1) & mask;
	return cnts[slot];
}

// mem: glszm3d_big_words (maxcells, slots) words of global memory, keys / cnts all zero on entry and exit
NYX_D void glszm3d_body_big (const Cta& c, const RoiRec& r, const Params& prm, const uint32_t* cell, uint32_t* mem,
	unsigned long long maxcells, uint32_t slots, bool matlab, double* out)
{
	const int w = r.w, h = r.h, d = r.d, ncell = w * h * d;
	if (r.imin == r.imax)
	{
		for (int i = c.tid; i < NCOL_GLSZM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	uint32_t* lab = mem;
	uint32_t* zs = lab + maxcells;
	unsigned long long* keys = (unsigned long long*)(zs + maxcells);	// 2 * maxcells words past the (8-byte aligned) region start
	uint32_t* cnts = (uint32_t*)(keys + slots);
	const uint32_t mask = slots - 1;
	const uint32_t skip = matlab ? 1u : 0u;
	FastDiv fw, fh;
	fw.init ((uint32_t)w);
	fh.init ((uint32_t)h);
	const unsigned long long K_PAIR = 1ull << 63, K_LEVEL = 1ull << 62, K_SIZE = 1ull << 61;

	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		zs[i] = 0;
		const uint32_t lv = cell[i] & CELL_LEVEL;
		int hd = i;
		if (lv != skip)
		{
			int x = i - (int)fw.div ((uint32_t)i) * w;
			while (x > 0 && (cell[hd - 1] & CELL_LEVEL) == lv)
			{
				hd--;
				x--;
			}
		}
		lab[i] = (uint32_t)hd;
	}
	c.sync();
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const uint32_t lv = cell[i] & CELL_LEVEL;
		if (lv == skip)
			continue;
		const int row = (int)fw.div ((uint32_t)i), x = i - row * w;
		const int z = (int)fh.div ((uint32_t)row), y = row - z * h;
		for (int dz = -1; dz <= 0; dz++)
		{
			const int zz = z + dz;
			if (zz < 0)
				continue;
			for (int dy = -1; dy <= (dz < 0 ? 1 : -1); dy++)
			{
				const int yy = y + dy;
				if (yy < 0 || yy >= h)
					continue;
				for (int dx = -1; dx <= 1; dx++)
				{
					const int xx = x + dx;
					if (xx < 0 || xx >= w)
						continue;
					const int j = (zz * h + yy) * w + xx;
					if ((cell[j] & CELL_LEVEL) == lv)
						uf_unite (c, lab, (uint32_t)i, (uint32_t)j);
				}
			}
		}
	}
	c.sync();
	unsigned long long np[1] = { 0 };
	for (int i = c.tid; i < ncell; i += c.nthr)
		if ((cell[i] & CELL_LEVEL) != skip)
		{
			const uint32_t root = uf_find (lab, (uint32_t)i);
			if (root != (uint32_t)i)
				lab[i] = root;
			atom_add (c, zs + root, 1u);
			np[0]++;
		}
	c.sync();

	// ---- pass 1 over the zones (= root cells)
	enum { Q_NZ, Q_MUJ, Q_MUI, Q_HGL, Q_N };
	unsigned long long q[Q_N];
	for (int k = 0; k < Q_N; k++)
		q[k] = 0;
	long long lae[1] = { 0 };
	enum { A_SAE, A_LGL, A_SAL, A_SAH, A_LAL, A_LAH, A_N };
	double s[A_N];
	for (int k = 0; k < A_N; k++)
		s[k] = 0;
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const uint32_t lv = cell[i] & CELL_LEVEL;
		if (lv == skip || lab[i] != (uint32_t)i)
			continue;
		const uint32_t sz = zs[i];
		htab64_add (c, keys, cnts, mask, K_PAIR | ((unsigned long long)lv << 36) | sz);
		htab64_add (c, keys, cnts, mask, K_LEVEL | lv);
		htab64_add (c, keys, cnts, mask, K_SIZE | sz);
		q[Q_NZ] += 1;
		q[Q_MUJ] += sz;
		q[Q_MUI] += lv;
		q[Q_HGL] += (unsigned long long)lv * lv;
		const int jj = (int)(sz * sz);		// `j * j` in int
		lae[0] += (long long)jj;
		const double i2 = (double)lv * (double)lv, j2 = (double)sz * (double)sz;
		s[A_SAE] += 1.0 / (double)jj;
		s[A_LGL] += 1.0 / i2;
		s[A_SAL] += 1.0 / (i2 * j2);
		s[A_SAH] += i2 / j2;
		s[A_LAL] += j2 / i2;
		s[A_LAH] += i2 * j2;
	}
	c.sync();
	cta_sum_u64<Q_N> (c, q);
	{
		unsigned long long t[2] = { (unsigned long long)(lae[0] < 0 ? -lae[0] : 0), (unsigned long long)(lae[0] > 0 ? lae[0] : 0) };
		cta_sum_u64<2> (c, t);
		lae[0] = (long long)t[1] - (long long)t[0];
	}
	cta_sum<A_N> (c, s);
	const uint32_t nz = (uint32_t)q[Q_NZ];
	if (nz == 0)
	{
		for (int i = c.tid; i < NCOL_GLSZM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	const double sum = (double)nz;
	const double mui = (double)q[Q_MUI] / sum, muj = (double)q[Q_MUJ] / sum;
	const double inv_sum = 1.0 / sum;

	// ---- pass 2: terms that need the counts (owner computes)
	unsigned long long q2[2] = { 0, 0 };
	double t[3] = { 0, 0, 0 };
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const uint32_t lv = cell[i] & CELL_LEVEL;
		if (lv == skip || lab[i] != (uint32_t)i)
			continue;
		const uint32_t sz = zs[i];
		q2[0] += htab64_get (keys, cnts, mask, K_LEVEL | lv);
		q2[1] += htab64_get (keys, cnts, mask, K_SIZE | sz);
		const double di = (double)lv - mui, dj = (double)sz - muj;
		t[0] += di * di;
		t[1] += dj * dj;
		const double pn = (double)htab64_get (keys, cnts, mask, K_PAIR | ((unsigned long long)lv << 36) | sz) * inv_sum;
		t[2] += fast_log10 (dadd (pn, EPS_ZONES));
	}
	cta_sum_u64<2> (c, q2);
	cta_sum<3> (c, t);
	c.sync();
	for (size_t i = c.tid; i < (size_t)slots; i += c.nthr)	// restore the all-zero table
	{
		keys[i] = 0;
		cnts[i] = 0;
	}
	c.sync();
	if (c.tid == 0)
	{
		out[S_SAE] = s[A_SAE] / sum;
		out[S_LAE] = (double)lae[0] / sum;
		out[S_GLN] = (double)q2[0] / sum;
		out[S_GLNN] = (double)q2[0] / (sum * sum);
		out[S_SZN] = (double)q2[1] / sum;
		out[S_SZNN] = (double)q2[1] / (sum * sum);
		out[S_ZP] = sum / (double)r.area;
		out[S_GLV] = t[0] / sum;
		out[S_ZV] = t[1] / sum;
		out[S_ZE] = -(t[2] / sum) * (1.0 / LOG10_2);
		out[S_LGLZE] = s[A_LGL] / sum;
		out[S_HGLZE] = (double)q[Q_HGL] / sum;
		out[S_SALGLE] = s[A_SAL] / sum;
		out[S_SAHGLE] = s[A_SAH] / sum;
		out[S_LALGLE] = s[A_LAL] / sum;
		out[S_LAHGLE] = s[A_LAH] / sum;
	}
	(void)np;
}

} // namespace nyx
