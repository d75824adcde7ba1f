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
