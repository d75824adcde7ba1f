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
// B200 design: one CTA per ROI; every cell tests "am I a run start" for the 4 directions and walks its run; the
// run-length matrix and both marginals are integer shared-memory histograms (order independent, deterministic).
#pragma once
#include "nyx_common.h"
#include "nyx_features.h"

namespace nyx
{

NYX_HD size_t glrlm_smem_words (int Lmax, int Nr)
{
	return 4 * ((size_t)Lmax * Nr + Lmax + Nr);
}

struct GlrlmDebug
{
	uint32_t* P;	// [4][Lmax*Nr] dense dump, may be null
};

// cell: binned tile (level | flags).  Lmax = largest possible level, Nr = max(w,h).  out: NCOL_GLRLM doubles.
NYX_D void glrlm_body (const Cta& c, const RoiRec& r, const Params& prm, const uint32_t* cell, uint32_t* mem, int Lmax, int Nr,
	double* out, const GlrlmDebug* dbg)
{
	constexpr int NA = 4;
	const int w = r.w, h = r.h, ncell = w * h;
	if (r.imin == r.imax)
	{
		for (int i = c.tid; i < NCOL_GLRLM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	const size_t pw = (size_t)Lmax * Nr;
	uint32_t* P = mem;			// [NA][Lmax][Nr]
	uint32_t* Ri = P + NA * pw;		// [NA][Lmax]  runs per level
	uint32_t* Cj = Ri + NA * Lmax;		// [NA][Nr]    runs per length
	const size_t nwords = glrlm_smem_words (Lmax, Nr);
	for (size_t i = c.tid; i < nwords; i += c.nthr)
		mem[i] = 0;
	c.sync();

	const int dxs[NA] = { 1, 1, 0, -1 }, dys[NA] = { 0, 1, 1, 1 };
	unsigned long long cnt[1] = { 0 };
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		uint32_t cc = cell[i];
		if (cc & CELL_NZ)
			cnt[0]++;
		uint32_t lv = cc & CELL_LEVEL;
		if (!lv)
			continue;
		int row = i / w, col = i - row * w;
#pragma unroll
		for (int a = 0; a < NA; a++)
		{
			int pr = row - dys[a], pc = col - dxs[a];
			if (pr >= 0 && pc >= 0 && pc < w && (cell[pr * w + pc] & CELL_LEVEL) == lv)
				continue;	// not a run start
			int len = 1, y = row + dys[a], x = col + dxs[a];
			while (y < h && x >= 0 && x < w && (cell[y * w + x] & CELL_LEVEL) == lv)
			{
				len++;
				y += dys[a];
				x += dxs[a];
			}
			atom_add (c, P + a * pw + (size_t)(lv - 1) * Nr + (len - 1), 1u);
			atom_add (c, Ri + a * Lmax + (lv - 1), 1u);
			atom_add (c, Cj + a * Nr + (len - 1), 1u);
		}
	}
	c.sync();
	cta_sum_u64<1> (c, cnt);
	const double Np = (double)cnt[0];

	if (dbg && dbg->P)
		for (size_t i = c.tid; i < NA * pw; i += c.nthr)
			dbg->P[i] = P[i];

	for (int a = 0; a < NA; a++)
	{
		// ---- integer marginal sums
		enum { Q_SUM, Q_LRE, Q_GLN, Q_RLN, Q_MUI, Q_MUJ, Q_HGL, Q_N };
		unsigned long long q[Q_N];
		for (int k = 0; k < Q_N; k++)
			q[k] = 0;
		for (int k = c.tid; k < Lmax; k += c.nthr)
		{
			unsigned long long v = Ri[a * Lmax + k];
			uint32_t lv = (uint32_t)(k + 1);
			q[Q_SUM] += v;
			q[Q_GLN] += v * v;				// :441-448
			q[Q_MUI] += v * lv;				// :598-605
			q[Q_HGL] += v * (unsigned long long)(uint32_t)(lv * lv);	// :744-751
		}
		for (int k = c.tid; k < Nr; k += c.nthr)
		{
			unsigned long long v = Cj[a * Nr + k], j = (unsigned long long)(k + 1);
			q[Q_LRE] += v * j * j;				// :409-415
			q[Q_RLN] += v * v;				// :509-517
			q[Q_MUJ] += v * j;				// :640-646
		}
		cta_sum_u64<Q_N> (c, q);
		if (q[Q_SUM] == 0)
		{
			for (int f = c.tid; f < R_NANGLED; f += c.nthr)	// :363-366 etc.
				out[f * 4 + a] = 0.0;
			continue;
		}
		const double sum = (double)q[Q_SUM];
		const double mui = (double)q[Q_MUI] / sum, muj = (double)q[Q_MUJ] / sum;

		// ---- FP marginal + cell sums
		enum { A_SRE, A_GLV, A_RV, A_RE, A_LGL, A_SRL, A_SRH, A_LRL, A_LRH, A_N };
		double s[A_N];
		for (int k = 0; k < A_N; k++)
			s[k] = 0;
		for (int k = c.tid; k < Nr; k += c.nthr)
		{
			uint32_t v = Cj[a * Nr + k];
			if (!v)
				continue;
			int j = k + 1;
			s[A_SRE] += (double)v / (double)(j * j);		// :381-383
			double dj = (double)j - muj;
			s[A_RV] += (double)v / sum * (dj * dj);			// :649-656
		}
		for (int k = c.tid; k < Lmax; k += c.nthr)
		{
			uint32_t v = Ri[a * Lmax + k];
			if (!v)
				continue;
			uint32_t lv = (uint32_t)(k + 1);
			double di = (double)lv - mui;
			s[A_GLV] += (double)v / sum * (di * di);		// :608-616
			s[A_LGL] += (double)v / (double)(uint32_t)(lv * lv);	// :712-719
		}
		for (size_t i = c.tid; i < pw; i += c.nthr)
		{
			uint32_t v = P[a * pw + i];
			if (!v)
				continue;
			uint32_t lv = (uint32_t)(i / Nr) + 1;
			int j = (int)(i % Nr) + 1;
			double p = (double)v / sum;
			s[A_RE] += dmul (p, ddiv (fast_log10 (dadd (p, EPS_ZONES)), LOG10_2));	// :684-685
			uint32_t i2 = lv * lv, j2 = (uint32_t)(j * j), i2j2 = lv * lv * (uint32_t)j * (uint32_t)j;
			s[A_SRL] += (double)v / (double)i2j2;					// :781
			s[A_SRH] += (double)v * (double)i2 / (double)(int)j2;			// :813
			s[A_LRL] += (double)v * (double)(int)j2 / (double)i2;			// :844
			s[A_LRH] += (double)v * (double)i2j2;					// :876
		}
		cta_sum<A_N> (c, s);
		if (c.tid == 0)
		{
			double f[R_NANGLED];
			f[R_SRE] = s[A_SRE] / sum;
			f[R_LRE] = (double)q[Q_LRE] / sum;
			f[R_GLN] = (double)q[Q_GLN] / sum;
			f[R_GLNN] = (double)q[Q_GLN] / (sum * sum);
			f[R_RLN] = (double)q[Q_RLN] / sum;
			f[R_RLNN] = (double)q[Q_RLN] / (sum * sum);
			f[R_RP] = sum / Np;					// :573
			f[R_GLV] = s[A_GLV];
			f[R_RV] = s[A_RV];
			f[R_RE] = -s[A_RE];
			f[R_LGLRE] = s[A_LGL] / sum;
			f[R_HGLRE] = (double)q[Q_HGL] / sum;
			f[R_SRLGLE] = s[A_SRL] / sum;
			f[R_SRHGLE] = s[A_SRH] / sum;
			f[R_LRLGLE] = s[A_LRL] / sum;
			f[R_LRHGLE] = s[A_LRH] / sum;
			for (int k = 0; k < R_NANGLED; k++)
				out[k * 4 + a] = f[k];
		}
	}
	c.sync();
	for (int k = c.tid; k < R_NANGLED; k += c.nthr)		// glrlm.cpp:336-351, 902-909
	{
		const double* v = out + k * 4;
		out[R_NANGLED * 4 + k] = (((v[0] + v[1]) + v[2]) + v[3]) / 4.0;
	}
}

} // namespace nyx
