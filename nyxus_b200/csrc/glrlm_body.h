// nyxus_b200 -- per-ROI GLRLM: 4 run-length matrices + 16 features x (4 angles + _AVE).
//
// Semantics follow the reference's GLRLMFeature (features/glrlm.cpp):
//   blank test   :29-52     aux_min == aux_max -> 4 x softNaN per feature (and the _AVE of those)
//   runs         :111-195   raster scan of the BINNED tile; a run starts at any unvisited non-zero cell and extends
//                           forward only (E / SE / S / SW) while the next cell has the same level.  Raster order makes
//                           every run maximal, so "run start" == "the backward neighbour differs or is outside".
//                           MATLAB binning maps background 0 -> level 1, so in-AABB background forms runs too.
//   Np           :198-204   number of non-zero ORIGINAL tile pixels
//   features     :357-882   sum_p == 0 -> 0.0 (not softNaN); SRLGLE/LRHGLE/... evaluate inten*inten*j*j in 32-bit
//                           unsigned arithmetic before widening (:717,749,781,813,844,876)
//
// B200 design: one CTA per ROI, ONE WARP PER ANGLE.  Every cell tests "am I a run start" for the warp's direction and
// walks its run; the run-length matrix and both marginals are integer shared-memory histograms (order independent,
// deterministic).  The reducers never sweep the Ng x Nr matrix: every feature is a sum over RUNS ("owner computes":
// sum_cells f(P) == sum_runs f(P_run)/P_run), and each run zeroes its own matrix cell afterwards, so the matrix region
// is all zero on entry and on exit.
#pragma once
#include "nyx_common.h"
#include "nyx_features.h"
#include "glcm_body.h"		// Counters<> (16-bit packed counters)

namespace nyx
{

// words of ONE angle's region: P[Lmax][Nr] (16-bit packed: a count is at most the number of runs <= cells of a tile, and the
// shared-memory tile classes end at 40000 cells) + Ri[Lmax] + Cj[Nr] (zero on entry/exit) + run length per cell (u16)
NYX_HD size_t glrlm_p_words (int Lmax, int Nr)
{
	return ((size_t)Lmax * Nr + 1) / 2;
}
NYX_HD size_t glrlm_angle_words (int Lmax, int Nr, int maxcells)
{
	return glrlm_p_words (Lmax, Nr) + Lmax + Nr + ((size_t)maxcells + 1) / 2;
}
NYX_HD size_t glrlm_smem_words (int Lmax, int Nr, int maxcells)
{
	return 4 * glrlm_angle_words (Lmax, Nr, maxcells);
}

struct GlrlmDebug
{
	uint32_t* P;	// [4][Lmax*Nr] dense dump, may be null
};

NYX_D void glrlm_angle (const Cta& g, int a, const RoiRec& r, const Params& prm, const uint32_t* cell, uint32_t* reg, int Lmax,
	int Nr, double* out, const GlrlmDebug* dbg)
{
	const int w = r.w, h = r.h, ncell = w * h;
	const size_t pn = (size_t)Lmax * Nr, pw = glrlm_p_words (Lmax, Nr);
	const Counters<true> P { reg };		// [Lmax][Nr], 16-bit packed
	uint32_t* Ri = reg + pw;		// [Lmax]  runs per level
	uint32_t* Cj = Ri + Lmax;		// [Nr]    runs per length
	uint16_t* rl = (uint16_t*)(Cj + Nr);	// [ncell] length of the run recorded at ONE of its cells (device: the last, host: the first; 0: none)
	const int dx = a == 2 ? 0 : (a == 3 ? -1 : 1), dy = a == 0 ? 0 : 1;	// E, SE, S, SW (glrlm.cpp:129-175)
	FastDiv fw;
	fw.init ((uint32_t)w);

	// ---- pass 1: runs -> matrix + marginals + per-run integer / FP sums
	enum { Q_SUM, Q_LRE, Q_MUI, Q_MUJ, Q_HGL, Q_NP, Q_N };
	unsigned long long q[Q_N];
	for (int k = 0; k < Q_N; k++)
		q[k] = 0;
	enum { A_SRE, A_LGL, A_SRL, A_SRH, A_LRL, A_LRH, A_N };
	double s[A_N];
	for (int k = 0; k < A_N; k++)
		s[k] = 0;
#if NYX_DEVICE_BUILD
	// Runs without walking them.  The cells are visited LINE BY LINE along the warp's direction -- rows (E), columns (S), or
	// the tile's diagonals wrapped around its left / right edge (SE: column (c0 + row) mod w, SW: (c0 - row) mod w), so that
	// every direction is one sequence of ncell cells in which a run is a maximal stretch of equal levels between BREAKS
	// (first cell of a line, the cell after a wrap, a level change).  32 consecutive cells per step: a ballot of the breaks
	// gives every break the position of the break before it, and the lane of a break closes the run that ended on the cell
	// before it (level and tile index of that cell come over from the lane below, or from the previous step for lane 0).
	// A virtual break at position ncell closes the last run.  rl[] is written at the run's LAST cell (same level as its head).
	{
		uint32_t* rlw = (uint32_t*)rl;
		for (int k = g.tid; k < (ncell + 1) / 2; k += g.nthr)
			rlw[k] = 0;
		g.sync();
		const int lane = g.tid, L = a == 0 ? w : h;
		FastDiv fl;
		fl.init ((uint32_t)L);
		uint32_t carry_lv = 0;
		int carry_idx = 0, carry_head = 0;
		for (int base = 0; base <= ncell; base += 32)
		{
			const int sq = base + lane;
			int idx = 0;
			uint32_t lv0 = 0;
			bool border = true;
			if (sq < ncell)
			{
				const int line = (int)fl.div ((uint32_t)sq), pos = sq - line * L;
				int row = pos, col = line;
				if (a == 0)
				{
					row = line;
					col = pos;
				}
				else if (a != 2)
				{
					const uint32_t t = (uint32_t)line + (uint32_t)pos * (a == 1 ? 1u : (uint32_t)(w - 1));
					col = (int)(t - fw.div (t) * (uint32_t)w);
				}
				border = pos == 0 || (a == 1 && col == 0) || (a == 3 && col == w - 1);
				idx = row * w + col;
				const uint32_t cc = cell[idx];
				if (cc & CELL_NZ)
					q[Q_NP]++;
				lv0 = cc & CELL_LEVEL;
			}
			uint32_t lv = __shfl_up_sync (0xffffffffu, lv0, 1);
			int pidx = __shfl_up_sync (0xffffffffu, idx, 1);
			if (lane == 0)
			{
				lv = carry_lv;
				pidx = carry_idx;
			}
			const bool brk = sq <= ncell && (border || lv != lv0);
			const unsigned bm = __ballot_sync (0xffffffffu, brk);
			if (brk && lv)
			{
				const unsigned below = bm & ((1u << lane) - 1u);
				const int len = sq - (below ? base + 31 - __clz (below) : carry_head);
				rl[pidx] = (uint16_t)len;
				P.inc (g, (int)((lv - 1) * (uint32_t)Nr + (uint32_t)(len - 1)));
				atom_add (g, Ri + (lv - 1), 1u);
				atom_add (g, Cj + (len - 1), 1u);
				const uint32_t i2 = lv * lv, j2 = (uint32_t)(len * len), i2j2 = lv * lv * (uint32_t)len * (uint32_t)len;	// 32-bit wrap
				q[Q_SUM] += 1;
				q[Q_LRE] += (unsigned long long)len * (unsigned long long)len;	// :409-415
				q[Q_MUI] += lv;							// :598-605
				q[Q_MUJ] += (unsigned long long)len;				// :640-646
				q[Q_HGL] += (unsigned long long)i2;				// :744-751
				if (lv < 65536u && (unsigned long long)i2 * (unsigned long long)j2 < 0x100000000ull)
					emphasis6_int (s, lv, (uint32_t)len);			// :381-383,712-719,781,813,844,876 (no 32-bit wrap: i2j2 == i2 * j2)
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
			carry_lv = __shfl_sync (0xffffffffu, lv0, 31);
			carry_idx = __shfl_sync (0xffffffffu, idx, 31);
			if (bm)
				carry_head = base + 31 - __clz (bm);
		}
	}
#else
	for (int i = g.tid; i < ncell; i += g.nthr)
	{
		const uint32_t cc = cell[i];
		if (cc & CELL_NZ)
			q[Q_NP]++;
		const uint32_t lv = cc & CELL_LEVEL;
		int len = 0;
		if (lv)
		{
			const int row = (int)fw.div ((uint32_t)i), col = i - row * w;
			const int pr = row - dy, pc = col - dx;
			if (!(pr >= 0 && pc >= 0 && pc < w && (cell[pr * w + pc] & CELL_LEVEL) == lv))
			{
				// walk the run: at most `room` more cells before the tile border, linear index stride dy * w + dx
				int room = dy ? h - 1 - row : w;
				if (dx > 0)
					room = room < w - 1 - col ? room : w - 1 - col;
				else if (dx < 0)
					room = room < col ? room : col;
				const int stride = dy * w + dx;
				int j = i + stride;
				len = 1;
				while (len <= room && (cell[j] & CELL_LEVEL) == lv)
				{
					len++;
					j += stride;
				}
				P.inc (g, (int)((lv - 1) * (uint32_t)Nr + (uint32_t)(len - 1)));
				atom_add (g, Ri + (lv - 1), 1u);
				atom_add (g, Cj + (len - 1), 1u);
				const uint32_t i2 = lv * lv, j2 = (uint32_t)(len * len), i2j2 = lv * lv * (uint32_t)len * (uint32_t)len;	// 32-bit wrap
				q[Q_SUM] += 1;
				q[Q_LRE] += (unsigned long long)len * (unsigned long long)len;	// :409-415
				q[Q_MUI] += lv;							// :598-605
				q[Q_MUJ] += (unsigned long long)len;				// :640-646
				q[Q_HGL] += (unsigned long long)i2;				// :744-751
				if (lv < 65536u && (unsigned long long)i2 * (unsigned long long)j2 < 0x100000000ull)
					emphasis6_int (s, lv, (uint32_t)len);			// :381-383,712-719,781,813,844,876 (no 32-bit wrap: i2j2 == i2 * j2)
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
		rl[i] = (uint16_t)len;
	}
#endif
	g.sync();
	cta_sum_u64<Q_N> (g, q);
	cta_sum<A_N> (g, s);

	if (dbg && dbg->P)
		for (size_t i = g.tid; i < pn; i += g.nthr)
			dbg->P[(size_t)a * pn + i] = P.get ((int)i);

	if (q[Q_SUM] == 0)	// :363-366 etc. (no run: the region is still all zero)
	{
		for (int f = g.tid; f < R_NANGLED; f += g.nthr)
			out[f * 4 + a] = 0.0;
		return;
	}
	const double sum = (double)q[Q_SUM], Np = (double)q[Q_NP];
	const double mui = (double)q[Q_MUI] / sum, muj = (double)q[Q_MUJ] / sum;
	const double inv_sum = 1.0 / sum;

	// ---- pass 2: terms that need the counts (owner computes)
	unsigned long long q2[2] = { 0, 0 };	// sum Ri^2 (:441-448), sum Cj^2 (:509-517)
	double t[3] = { 0, 0, 0 };		// GLV, RV, RE
	for (int i = g.tid; i < ncell; i += g.nthr)
	{
		const int len = rl[i];
		if (!len)
			continue;
		const uint32_t lv = cell[i] & CELL_LEVEL;
		q2[0] += Ri[lv - 1];
		q2[1] += Cj[len - 1];
		const double di = (double)lv - mui, dj = (double)len - muj;
		t[0] += di * di;								// :608-616
		t[1] += dj * dj;								// :649-656
		const double p = (double)P.get ((int)((lv - 1) * (uint32_t)Nr + (uint32_t)(len - 1))) * inv_sum;
		t[2] += fast_log10 (dadd (p, EPS_ZONES));					// :684-685
	}
	cta_sum_u64<2> (g, q2);
	cta_sum<3> (g, t);
	g.sync();

	// ---- pass 3: restore the all-zero region
	if (pw + Lmax + Nr <= (size_t)ncell * 2)	// small matrix: wipe P, Ri, Cj (contiguous) with 128-bit stores
		wipe_words (g, reg, pw + Lmax + Nr);
	else
	{
		for (int i = g.tid; i < ncell; i += g.nthr)
		{
			const int len = rl[i];
			if (len)
				P.clear ((int)(((cell[i] & CELL_LEVEL) - 1) * (uint32_t)Nr + (uint32_t)(len - 1)));
		}
		for (int k = g.tid; k < Lmax + Nr; k += g.nthr)
			Ri[k] = 0;
	}
	g.sync();

	if (g.tid == 0)
	{
		double f[R_NANGLED];
		f[R_SRE] = s[A_SRE] * inv_sum;
		f[R_LRE] = (double)q[Q_LRE] * inv_sum;
		f[R_GLN] = (double)q2[0] * inv_sum;
		f[R_GLNN] = (double)q2[0] * (inv_sum * inv_sum);
		f[R_RLN] = (double)q2[1] * inv_sum;
		f[R_RLNN] = (double)q2[1] * (inv_sum * inv_sum);
		f[R_RP] = sum / Np;					// :573
		f[R_GLV] = t[0] * inv_sum;
		f[R_RV] = t[1] * inv_sum;
		f[R_RE] = -(t[2] * inv_sum) * (1.0 / LOG10_2);
		f[R_LGLRE] = s[A_LGL] * inv_sum;
		f[R_HGLRE] = (double)q[Q_HGL] * inv_sum;
		f[R_SRLGLE] = s[A_SRL] * inv_sum;
		f[R_SRHGLE] = s[A_SRH] * inv_sum;
		f[R_LRLGLE] = s[A_LRL] * inv_sum;
		f[R_LRHGLE] = s[A_LRH] * inv_sum;
		for (int k = 0; k < R_NANGLED; k++)
			out[k * 4 + a] = f[k];
	}
}

// cell: binned tile (level | flags).  Lmax = largest possible level, Nr = max(w,h) bound of the launch, maxcells = cell
// bound of the launch.  mem: 4 angle regions (zero on entry and exit).  out: NCOL_GLRLM doubles.
NYX_D void glrlm_body (const Cta& c, const RoiRec& r, const Params& prm, const uint32_t* cell, uint32_t* mem, int Lmax, int Nr,
	int maxcells, double* out, const GlrlmDebug* dbg)
{
	constexpr int NA = 4;
	if (r.imin == r.imax)
	{
		for (int i = c.tid; i < NCOL_GLRLM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	const size_t aw = glrlm_angle_words (Lmax, Nr, maxcells);
	const Cta g = c.sub();
	double* st = c.red ? c.red : out;	// angled values staged in the reduction scratch (unused by the angle warps), see glcm_body
	for (int a = c.sub_id(); a < NA; a += c.sub_count())
		glrlm_angle (g, a, r, prm, cell, mem + (size_t)a * aw, Lmax, Nr, st, dbg);
	c.sync();
	if (st != out)
		for (int k = c.tid; k < R_NANGLED * 4; k += c.nthr)
			out[k] = st[k];
	for (int k = c.tid; k < R_NANGLED; k += c.nthr)		// glrlm.cpp:336-351, 902-909
	{
		const double* v = st + k * 4;
		out[R_NANGLED * 4 + k] = (((v[0] + v[1]) + v[2]) + v[3]) / 4.0;
	}
}

} // namespace nyx
