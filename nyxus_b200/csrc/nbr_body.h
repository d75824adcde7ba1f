// nyxus_b200 -- per-ROI 8-neighbourhood families: GLDM, NGLDM, NGTDM.
//
// GLDM  (reference features/gldm.cpp): blank :22-40; centres and neighbours are ROI pixels with ORIGINAL intensity != 0
//        (:83,:91...), dependence nd = 1 + #equal-level 8-neighbours; matrix[level][nd] :170-205; features :295-554.
// NGLDM (reference features/ngldm.cpp): blank :163-185; mask from the pixel CLOUD (intensity-0 ROI pixels count, :224-231);
//        levels by Nyxus::to_grayscale(i, 0, aux_max, Ng, ibsi) (:42,126,144); matrix[level][matches 0..8] :116-154;
//        features :267-355 (std::log entropy, DCP = 1).
// NGTDM (reference features/ngtdm.cpp): no blank test, |I| < 2 -> softNaN (:70-78); centres/neighbours by BINNED level != 0
//        (:88-146); N[level]++, S[level] += |level - mean of neighbours| (:159-183); features :228-347.
//
// B200 design: one CTA per ROI, one 3x3 stencil pass over the shared-memory tile, integer shared-memory histograms.
// NGTDM's S is accumulated as exact integers |level*nd - sum| per neighbour count nd and divided once per (level, nd),
// which makes it independent of the accumulation order (deterministic) and within a few ulp of the reference.
#pragma once
#include "nyx_common.h"
#include "nyx_features.h"

namespace nyx
{

struct DenseDebug
{
	uint32_t* P;	// dense dump of the family's matrix, may be null
};

// ------------------------------------------------------------------------------------------------
// GLDM: hist[(Lmax+1)][10]
// ------------------------------------------------------------------------------------------------
NYX_HD size_t gldm_smem_words (int Lmax)
{
	return (size_t)(Lmax + 1) * 10;
}

NYX_D void gldm_body (const Cta& c, const RoiRec& r, const Params& prm, const uint32_t* cell, uint32_t* mem, int Lmax,
	double* out, const DenseDebug* dbg)
{
	const int w = r.w, h = r.h, ncell = w * h;
	if (r.imin == r.imax)
	{
		for (int i = c.tid; i < NCOL_GLDM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	uint32_t* P = mem;	// [level][nd], nd in 1..9
	const size_t nw = gldm_smem_words (Lmax);
	for (size_t i = c.tid; i < nw; i += c.nthr)
		P[i] = 0;
	c.sync();
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		uint32_t cc = cell[i];
		if (!(cc & CELL_NZ))
			continue;
		uint32_t lv = cc & CELL_LEVEL;
		int y = i / w, x = i - y * w, nd = 1;
		for (int dy = -1; dy <= 1; dy++)
			for (int dx = -1; dx <= 1; dx++)
			{
				if (!dx && !dy)
					continue;
				int yy = y + dy, xx = x + dx;
				if (yy < 0 || yy >= h || xx < 0 || xx >= w)
					continue;
				uint32_t q = cell[yy * w + xx];
				if ((q & CELL_NZ) && (q & CELL_LEVEL) == lv)
					nd++;
			}
		atom_add (c, P + (size_t)lv * 10 + nd, 1u);
	}
	c.sync();
	if (dbg && dbg->P)
		for (size_t i = c.tid; i < nw; i += c.nthr)
			dbg->P[i] = P[i];

	// per-level and per-dependence marginals are recomputed on the fly (10 columns)
	enum { Q_NZ, Q_LDE, Q_GLN, Q_MUI, Q_MUJ, Q_HGL, Q_N };
	unsigned long long q[Q_N];
	for (int k = 0; k < Q_N; k++)
		q[k] = 0;
	unsigned long long sj[10];
	for (int k = 0; k < 10; k++)
		sj[k] = 0;
	enum { A_SDE, A_LGL, A_DE0, A_SDL, A_SDH, A_LDL, A_LDH, A_N };
	double s[A_N];
	for (int k = 0; k < A_N; k++)
		s[k] = 0;
	for (int lv = c.tid + 1; lv <= Lmax; lv += c.nthr)
	{
		unsigned long long si = 0;
		const double in = (double)lv;
		for (int j = 1; j <= 9; j++)
		{
			unsigned long long v = P[(size_t)lv * 10 + j];
			if (!v)
				continue;
			si += v;
			sj[j] += v;
			q[Q_LDE] += v * (unsigned long long)(j * j);			// gldm.cpp:314-316
			q[Q_MUJ] += v * (unsigned long long)j;				// :410
			s[A_SDE] += (double)v / ((double)j * (double)j);		// :301-302
			s[A_SDL] += (double)v / (double)(in * in * j * j);		// :498-505
			s[A_SDH] += (double)v * (in * in) / (double)(j * j);		// :513-520
			s[A_LDL] += (double)v * (double)(j * j) / (in * in);		// :528-535
			s[A_LDH] += (double)v * (double)(in * in * j * j);		// :543-550
		}
		q[Q_NZ] += si;
		q[Q_GLN] += si * si;							// :325-338
		q[Q_MUI] += si * (unsigned long long)lv;				// :388-395
		q[Q_HGL] += si * (unsigned long long)lv * (unsigned long long)lv;	// :470-484
		if (si)
			s[A_LGL] += (double)si / (in * in);				// :452-464
	}
	cta_sum_u64<Q_N> (c, q);
	cta_sum_u64<10> (c, sj);
	cta_sum<A_N> (c, s);
	if (q[Q_NZ] == 0)	// gldm.cpp:213-231
	{
		for (int i = c.tid; i < NCOL_GLDM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	const double Nz = (double)q[Q_NZ];
	const double mui = (double)q[Q_MUI] / Nz, muj = (double)q[Q_MUJ] / Nz;
	double t[2] = { 0, 0 };	// GLV, DE
	for (int lv = c.tid + 1; lv <= Lmax; lv += c.nthr)
	{
		unsigned long long si = 0;
		for (int j = 1; j <= 9; j++)
		{
			uint32_t v = P[(size_t)lv * 10 + j];
			if (!v)
				continue;
			si += v;
			double p = (double)v / Nz;
			t[1] += dmul (p, ddiv (fast_log10 (dadd (p, EPS_ZONES)), LOG10_2));	// :437-446
		}
		if (si)
		{
			double di = (double)lv - mui;
			t[0] += (double)si / Nz * (di * di);				// :398-406
		}
	}
	cta_sum<2> (c, t);
	if (c.tid == 0)
	{
		unsigned long long dn = 0;
		double dv = 0;
		for (int j = 1; j <= 9; j++)
		{
			dn += sj[j] * sj[j];						// :343-357
			double d = (double)j - muj;
			dv += (double)sj[j] / Nz * (d * d);				// :418-429
		}
		out[D_SDE] = s[A_SDE] / Nz;
		out[D_LDE] = (double)q[Q_LDE] / Nz;
		out[D_GLN] = (double)q[Q_GLN] / Nz;
		out[D_DN] = (double)dn / Nz;
		out[D_DNN] = (double)dn / (Nz * Nz);
		out[D_GLV] = t[0];
		out[D_DV] = dv;
		out[D_DE] = -t[1];
		out[D_LGLE] = s[A_LGL] / Nz;
		out[D_HGLE] = (double)q[Q_HGL] / Nz;
		out[D_SDLGLE] = s[A_SDL] / Nz;
		out[D_SDHGLE] = s[A_SDH] / Nz;
		out[D_LDLGLE] = s[A_LDL] / Nz;
		out[D_LDHGLE] = s[A_LDH] / Nz;
	}
}

// ------------------------------------------------------------------------------------------------
// NGLDM: the tile passed here carries to_grayscale() levels (see load_tile_ngldm); hist[(Lmax+1)][9]
// ------------------------------------------------------------------------------------------------
NYX_HD size_t ngldm_smem_words (int Lmax)
{
	return (size_t)(Lmax + 1) * 9;
}

// tile for NGLDM: level = to_grayscale(i, 0, aux_max, greydepth, ibsi) (ngldm.cpp:42,126), CELL_IN = cloud mask
NYX_D void load_tile_ngldm (const Cta& c, const ImageView& im, const RoiRec& r, const Params& prm, uint32_t* cell)
{
	const int n = r.w * r.h;
	for (int i = c.tid; i < n; i += c.nthr)
	{
		int y = i / r.w, x = i - y * r.w;
		size_t g = (size_t)(r.y0 + y) * im.W + (size_t)(r.x0 + x);
		bool in = im.label[g] == r.label;
		uint32_t code = 0;
		if (in)
			code = (to_grayscale (load_inten (im, g), 0, r.imax, (uint32_t)prm.greydepth, prm.ibsi != 0) & CELL_LEVEL) | CELL_IN;
		cell[i] = code;
	}
}

NYX_D void ngldm_body (const Cta& c, const RoiRec& r, const Params& prm, const uint32_t* cell, uint32_t* mem, int Lmax,
	double* out, const DenseDebug* dbg)
{
	const int w = r.w, h = r.h, ncell = w * h;
	if (r.imin == r.imax)
	{
		for (int i = c.tid; i < NCOL_NGLDM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	uint32_t* P = mem;	// [level 0..Lmax][matches 0..8]
	const size_t nw = ngldm_smem_words (Lmax);
	for (size_t i = c.tid; i < nw; i += c.nthr)
		P[i] = 0;
	c.sync();
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		uint32_t cc = cell[i];
		if (!(cc & CELL_IN))
			continue;
		uint32_t lv = cc & CELL_LEVEL;
		int y = i / w, x = i - y * w, nm = 0;
		for (int dy = -1; dy <= 1; dy++)
			for (int dx = -1; dx <= 1; dx++)
			{
				if (!dx && !dy)
					continue;
				int yy = y + dy, xx = x + dx;
				if (yy < 0 || yy >= h || xx < 0 || xx >= w)
					continue;
				uint32_t q = cell[yy * w + xx];
				if ((q & CELL_IN) && (q & CELL_LEVEL) == lv)
					nm++;
			}
		atom_add (c, P + (size_t)lv * 9 + nm, 1u);
	}
	c.sync();
	if (dbg && dbg->P)
		for (size_t i = c.tid; i < nw; i += c.nthr)
			dbg->P[i] = P[i];

	unsigned long long q[2] = { 0, 0 };	// Ns, sum Sg^2
	unsigned long long sr[9];
	for (int k = 0; k < 9; k++)
		sr[k] = 0;
	for (int lv = c.tid; lv <= Lmax; lv += c.nthr)
	{
		unsigned long long sg = 0;
		for (int j = 0; j < 9; j++)
		{
			unsigned long long v = P[(size_t)lv * 9 + j];
			sg += v;
			sr[j] += v;
		}
		q[0] += sg;
		q[1] += sg * sg;
	}
	cta_sum_u64<2> (c, q);
	cta_sum_u64<9> (c, sr);
	const double Ns = (double)q[0];
	enum { A_LDE, A_HDE, A_LGLCE, A_HGLCE, A_LDLGLE, A_LDHGLE, A_HDLGLE, A_HDHGLE, A_GLM, A_DCM, A_DCENT, A_DCENE, A_N };
	double s[A_N];
	for (int k = 0; k < A_N; k++)
		s[k] = 0;
	for (int lv = c.tid; lv <= Lmax; lv += c.nthr)
	{
		const double g = (double)lv;
		for (int j = 0; j < 9; j++)
		{
			uint32_t v = P[(size_t)lv * 9 + j];
			if (!v)
				continue;
			double dc = (double)(j + 1), sij = (double)v, pij = sij / Ns;	// ngldm.cpp:283-305
			s[A_LDE] += sij / dc / dc;
			s[A_HDE] += sij * dc * dc;
			if (g != 0.0)
			{
				s[A_LGLCE] += sij / g / g;
				s[A_LDLGLE] += sij / dc / dc / g / g;
				s[A_HDLGLE] += sij * dc * dc / g / g;
			}
			s[A_HGLCE] += sij * g * g;
			s[A_LDHGLE] += sij * g * g / dc / dc;
			s[A_HDHGLE] += sij * dc * dc * g * g;
			s[A_GLM] += g * pij;
			s[A_DCM] += dc * pij;
			if (pij > 0.0)
				s[A_DCENT] -= pij * log (pij) / log (2.0);
			s[A_DCENE] += pij * pij;
		}
	}
	cta_sum<A_N> (c, s);
	double t[2] = { 0, 0 };
	for (int lv = c.tid; lv <= Lmax; lv += c.nthr)
	{
		const double g = (double)lv;
		for (int j = 0; j < 9; j++)
		{
			uint32_t v = P[(size_t)lv * 9 + j];
			if (!v)
				continue;
			double dc = (double)(j + 1), pij = (double)v / Ns;
			t[0] += (g - s[A_GLM]) * (g - s[A_GLM]) * pij;			// :330
			t[1] += (dc - s[A_DCM]) * (dc - s[A_DCM]) * pij;		// :337
		}
	}
	cta_sum<2> (c, t);
	if (c.tid == 0)
	{
		unsigned long long dcn = 0;
		for (int j = 0; j < 9; j++)
			dcn += sr[j] * sr[j];
		out[N_LDE] = s[A_LDE] / Ns;
		out[N_HDE] = s[A_HDE] / Ns;
		out[N_LGLCE] = s[A_LGLCE] / Ns;
		out[N_HGLCE] = s[A_HGLCE] / Ns;
		out[N_LDLGLE] = s[A_LDLGLE] / Ns;
		out[N_LDHGLE] = s[A_LDHGLE] / Ns;
		out[N_HDLGLE] = s[A_HDLGLE] / Ns;
		out[N_HDHGLE] = s[A_HDHGLE] / Ns;
		out[N_GLNU] = (double)q[1] / Ns;
		out[N_GLNUN] = (double)q[1] / (Ns * Ns);
		out[N_DCNU] = (double)dcn / Ns;
		out[N_DCNUN] = (double)dcn / (Ns * Ns);
		out[N_DCP] = 1.0;						// :354
		out[N_GLM] = s[A_GLM];
		out[N_GLV] = t[0];
		out[N_DCM] = s[A_DCM];
		out[N_DCV] = t[1];
		out[N_DCENT] = s[A_DCENT];
		out[N_DCENE] = s[A_DCENE];
	}
}

// ------------------------------------------------------------------------------------------------
// NGTDM: N[(Lmax+1)], SI[(Lmax+1)][8] exact integer |level*nd - sum|, PR[(Lmax+1)] presence, then P,S doubles
// ------------------------------------------------------------------------------------------------
NYX_HD size_t ngtdm_smem_words (int Lmax)
{
	return (size_t)(Lmax + 1) * (1 + 8 + 1) + 2 + (size_t)(Lmax + 1) * 4;
}

struct NgtdmDebug
{
	uint32_t* N;	// [Lmax+1]
	double* S;	// [Lmax+1]
};

NYX_D void ngtdm_body (const Cta& c, const RoiRec& r, const Params& prm, const uint32_t* cell, uint32_t* mem, int Lmax,
	double* out, const NgtdmDebug* dbg)
{
	const int w = r.w, h = r.h, ncell = w * h;
	const int L1 = Lmax + 1;
	uint32_t* Nn = mem;			// [L1]
	uint32_t* SI = Nn + L1;			// [L1][8]
	uint32_t* PR = SI + (size_t)L1 * 8;	// [L1] level present anywhere in the binned tile
	size_t off = (size_t)L1 * 10;
	off += off & 1;
	double* Pd = (double*)(mem + off);	// [L1]
	double* Sd = Pd + L1;			// [L1]
	for (size_t i = c.tid; i < (size_t)L1 * 10; i += c.nthr)
		mem[i] = 0;
	c.sync();
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		uint32_t lv = cell[i] & CELL_LEVEL;
		if (!lv)
			continue;
		PR[lv] = 1;
		int y = i / w, x = i - y * w, nd = 0;
		uint32_t sum = 0;
		for (int dy = -1; dy <= 1; dy++)
			for (int dx = -1; dx <= 1; dx++)
			{
				if (!dx && !dy)
					continue;
				int yy = y + dy, xx = x + dx;
				if (yy < 0 || yy >= h || xx < 0 || xx >= w)
					continue;
				uint32_t q = cell[yy * w + xx] & CELL_LEVEL;
				if (q)
				{
					sum += q;
					nd++;
				}
			}
		if (nd > 0)
		{
			long long df = (long long)lv * nd - (long long)sum;
			atom_add (c, Nn + lv, 1u);
			atom_add (c, SI + (size_t)lv * 8 + (nd - 1), (uint32_t)(df < 0 ? -df : df));
		}
	}
	c.sync();
	// ---- counts: Ngp = unique non-zero levels in the tile (ngtdm.cpp:53-55,152), Nvc = sum N
	unsigned long long q[2] = { 0, 0 };
	for (int lv = c.tid + 1; lv <= Lmax; lv += c.nthr)
	{
		q[0] += PR[lv] ? 1 : 0;
		q[1] += Nn[lv];
	}
	cta_sum_u64<2> (c, q);
	const int Ngp = (int)q[0];
	const int Isize = prm.ibsi ? (Ngp ? Lmax + 1 : 0) : Ngp;	// |I| (:57-67); IBSI: 0..max(D)
	if (Isize < 2)
	{
		for (int i = c.tid; i < NCOL_NGTDM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	const double Nvc = (double)q[1], Nvp = (double)q[1];	// every kept pixel has a positive neighbourhood mean (:181-182)
	for (int lv = c.tid; lv <= Lmax; lv += c.nthr)
	{
		double sv = 0;
		for (int k = 0; k < 8; k++)
			if (SI[(size_t)lv * 8 + k])
				sv += (double)SI[(size_t)lv * 8 + k] / (double)(k + 1);
		Sd[lv] = sv;
		Pd[lv] = (double)Nn[lv] / Nvc;		// :191-192
	}
	c.sync();
	if (dbg && dbg->N)
		for (int lv = c.tid; lv <= Lmax; lv += c.nthr)
		{
			dbg->N[lv] = Nn[lv];
			dbg->S[lv] = Sd[lv];
		}
	enum { A_PS, A_S, A_CON, A_BUSY, A_CPLX, A_STR, A_N };
	double s[A_N];
	for (int k = 0; k < A_N; k++)
		s[k] = 0;
	for (int lv = c.tid + 1; lv <= Lmax; lv += c.nthr)
	{
		s[A_PS] += Pd[lv] * Sd[lv];		// :231-233
		s[A_S] += Sd[lv];			// :256-258
	}
	const long long npair = (long long)Lmax * Lmax;
	for (long long k = c.tid; k < npair; k += c.nthr)
	{
		int i = (int)(k / Lmax) + 1, j = (int)(k % Lmax) + 1;
		double pi = Pd[i], pj = Pd[j];
		if (pi == 0 || pj == 0)
			continue;
		double iv = (double)i, jv = (double)j;
		s[A_CON] += pi * pj * (iv - jv) * (iv - jv);						// :245-251
		s[A_BUSY] += fabs (pi * iv - pj * jv);							// :283-287
		s[A_CPLX] += fabs (iv - jv) * (pi * Sd[i] + pj * Sd[j]) / (pi + pj);			// :311-314
		s[A_STR] += (pi + pj) * (iv - jv) * (iv - jv);						// :333-336
	}
	cta_sum<A_N> (c, s);
	if (c.tid == 0)
	{
		out[T_COARSENESS] = 1.0 / s[A_PS];
		int ngp2 = Ngp > 1 ? Ngp * (Ngp - 1) : Ngp;
		out[T_CONTRAST] = (s[A_CON] / (double)ngp2) * (s[A_S] / Nvc);
		out[T_BUSYNESS] = (Ngp == 1 || s[A_BUSY] == 0) ? 0.0 : s[A_PS] / s[A_BUSY];
		out[T_COMPLEXITY] = s[A_CPLX] / Nvp;
		out[T_STRENGTH] = s[A_STR] / s[A_S];
	}
}

} // namespace nyx
