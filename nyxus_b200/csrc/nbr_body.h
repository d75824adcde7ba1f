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
// B200 design: one WARP (small tiles) or one CTA per ROI, one 3x3 stencil pass over the shared-memory tile, integer
// shared-memory histograms.  GLDM / NGLDM reducers are "owner computes": every pixel re-reads the count of its own
// matrix cell (its dependence is stashed in spare bits of the tile word), so no [level][dependence] sweep per ROI and the
// matrix region is all zero on entry and exit.  NGTDM's S is accumulated as exact integers |level*nd - sum| per
// neighbour count nd and divided once per (level, nd) -- independent of the accumulation order (deterministic) and
// within a few ulp of the reference; its O(Ng^2) pair features run over the compacted list of present levels only.
#pragma once
#include "nyx_common.h"
#include "nyx_features.h"

namespace nyx
{

struct DenseDebug
{
	uint32_t* P;	// dense dump of the family's matrix, may be null
};


// number of 8-neighbours q of cell (x,y) with (cell[q] & flag) and the same level (flag: ONE flag bit).  Branch-free: a
// neighbour outside the tile is redirected to the centre cell and masked out, so all lanes of a warp run the same eight
// load / compare / add groups whatever their position in the tile
NYX_D int nbr_same_level (const uint32_t* cell, int w, int h, int x, int y, uint32_t lv, uint32_t flag)
{
	const int i = y * w + x;
	const bool xl = x > 0, xr = x < w - 1, yu = y > 0, yd = y < h - 1;
	const uint32_t msk = flag | CELL_LEVEL, want = flag | lv;
	int n = 0;
#define NYX_NB(ok, off) { const bool k_ = (ok); const uint32_t q_ = cell[k_ ? i + (off) : i]; n += (k_ && (q_ & msk) == want) ? 1 : 0; }
	NYX_NB (yu && xl, -w - 1)
	NYX_NB (yu, -w)
	NYX_NB (yu && xr, -w + 1)
	NYX_NB (xl, -1)
	NYX_NB (xr, 1)
	NYX_NB (yd && xl, w - 1)
	NYX_NB (yd, w)
	NYX_NB (yd && xr, w + 1)
#undef NYX_NB
	return n;
}

// ------------------------------------------------------------------------------------------------
// GLDM: P[(Lmax+1)][10] + Si[Lmax+1] + Sj[10], all zero on entry and exit
// ------------------------------------------------------------------------------------------------
// (+ one byte per cell: the cell's dependence count, kept between the passes)
NYX_HD size_t gldm_smem_words (int Lmax, uint32_t maxcells = 0)
{
	return (size_t)(Lmax + 1) * 11 + 10 + ((size_t)maxcells + 3) / 4;
}

NYX_D void gldm_body (const Cta& c, const RoiRec& r, const Params& prm, uint32_t* cell, uint32_t* mem, int Lmax,
	double* out, const DenseDebug* dbg)
{
	const int w = r.w, h = r.h, ncell = w * h;
	if (r.imin == r.imax)
	{
		for (int i = c.tid; i < NCOL_GLDM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	uint32_t* P = mem;				// [level][nd], nd in 1..9
	uint32_t* Si = P + (size_t)(Lmax + 1) * 10;	// [level]
	uint32_t* Sj = Si + (Lmax + 1);			// [nd]
	uint8_t* aux = (uint8_t*)(Sj + 10);		// [ncell] dependence count of the cell (written in pass 1, read in passes 2-3)
	FastDiv fw;
	fw.init ((uint32_t)w);

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
		const uint32_t cc = cell[i];
		if (!(cc & CELL_NZ))
			continue;
		const uint32_t lv = cc & CELL_LEVEL;
		const int y = (int)fw.div ((uint32_t)i), x = i - y * w;
		const int j = 1 + nbr_same_level (cell, w, h, x, y, lv, CELL_NZ);	// gldm.cpp:74-166
		aux[i] = (uint8_t)j;
		atom_add (c, P + (size_t)lv * 10 + j, 1u);
		atom_add (c, Si + lv, 1u);
		atom_add (c, Sj + j, 1u);
		const double in = (double)lv;
		q[Q_NZ] += 1;
		q[Q_LDE] += (unsigned long long)(j * j);			// gldm.cpp:314-316
		q[Q_MUJ] += (unsigned long long)j;				// :410
		q[Q_MUI] += lv;							// :388-395
		q[Q_HGL] += (unsigned long long)lv * (unsigned long long)lv;	// :470-484
		emphasis6_int (s, lv, (uint32_t)j);			// :301-302, 452-464, 498-550
	}
	c.sync();
	if (dbg && dbg->P)
		for (size_t i = c.tid; i < (size_t)(Lmax + 1) * 10; i += c.nthr)
			dbg->P[i] = P[i];
	cta_sum_u64<Q_N> (c, q);
	cta_sum<A_N> (c, s);
	if (q[Q_NZ] == 0)	// gldm.cpp:213-231 (nothing was counted: the region is still all zero)
	{
		for (int i = c.tid; i < NCOL_GLDM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	const double Nz = (double)q[Q_NZ];
	const double mui = (double)q[Q_MUI] / Nz, muj = (double)q[Q_MUJ] / Nz;
	const double inv_nz = 1.0 / Nz;

	// ---- pass 2: terms that need the counts (owner computes)
	unsigned long long q2[2] = { 0, 0 };	// sum Si^2 (:325-338), sum Sj^2 (:343-357)
	double t[3] = { 0, 0, 0 };		// GLV, DV, DE
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const uint32_t cc = cell[i];
		if (!(cc & CELL_NZ))
			continue;
		const uint32_t lv = cc & CELL_LEVEL, j = aux[i];
		q2[0] += Si[lv];
		q2[1] += Sj[j];
		const double di = (double)lv - mui, dj = (double)j - muj;
		t[0] += di * di;							// :398-406
		t[1] += dj * dj;							// :418-429
		const double p = (double)P[(size_t)lv * 10 + j] * inv_nz;
		t[2] += fast_log10 (dadd (p, EPS_ZONES));				// :437-446
	}
	cta_sum_u64<2> (c, q2);
	cta_sum<3> (c, t);
	c.sync();
	// ---- pass 3: restore the all-zero region
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const uint32_t cc = cell[i];
		if (!(cc & CELL_NZ))
			continue;
		const uint32_t lv = cc & CELL_LEVEL, j = aux[i];
		P[(size_t)lv * 10 + j] = 0;
		Si[lv] = 0;
	}
	for (int k = c.tid; k < 10; k += c.nthr)
		Sj[k] = 0;
	c.sync();
	if (c.tid == 0)
	{
		out[D_SDE] = s[A_SDE] * inv_nz;
		out[D_LDE] = (double)q[Q_LDE] * inv_nz;
		out[D_GLN] = (double)q2[0] * inv_nz;
		out[D_DN] = (double)q2[1] * inv_nz;
		out[D_DNN] = (double)q2[1] * (inv_nz * inv_nz);
		out[D_GLV] = t[0] * inv_nz;
		out[D_DV] = t[1] * inv_nz;
		out[D_DE] = -(t[2] * inv_nz) * (1.0 / LOG10_2);
		out[D_LGLE] = s[A_LGL] * inv_nz;
		out[D_HGLE] = (double)q[Q_HGL] * inv_nz;
		out[D_SDLGLE] = s[A_SDL] * inv_nz;
		out[D_SDHGLE] = s[A_SDH] * inv_nz;
		out[D_LDLGLE] = s[A_LDL] * inv_nz;
		out[D_LDHGLE] = s[A_LDH] * inv_nz;
	}
}

// ------------------------------------------------------------------------------------------------
// NGLDM: the tile passed here carries to_grayscale() levels (see load_tile_ngldm);
// P[(Lmax+1)][9] + Sg[Lmax+1] + Sr[9], all zero on entry and exit
// ------------------------------------------------------------------------------------------------
NYX_HD size_t ngldm_smem_words (int Lmax, uint32_t maxcells = 0)
{
	return (size_t)(Lmax + 1) * 10 + 9 + ((size_t)maxcells + 3) / 4;
}

// tile for NGLDM: level = to_grayscale(i, 0, aux_max, greydepth, ibsi) (ngldm.cpp:42,126), CELL_IN = cloud mask
NYX_D void load_tile_ngldm (const Cta& c, const ImageView& im, const RoiRec& r, const Params& prm, uint32_t* cell)
{
	const int n = r.w * r.h;
	FastDiv fw;
	fw.init ((uint32_t)r.w);
	for (int i = c.tid; i < n; i += c.nthr)
	{
		int y = (int)fw.div ((uint32_t)i), x = i - y * r.w;
		size_t g = (size_t)(r.y0 + y) * im.W + (size_t)(r.x0 + x);
		bool in = im.label[g] == r.label;
		uint32_t code = 0;
		if (in)
			code = (to_grayscale (load_inten (im, g), 0, r.imax, (uint32_t)prm.greydepth, prm.ibsi != 0) & CELL_LEVEL) | CELL_IN;
		cell[i] = code;
	}
}

NYX_D void ngldm_body (const Cta& c, const RoiRec& r, const Params& prm, uint32_t* cell, uint32_t* mem, int Lmax,
	double* out, const DenseDebug* dbg)
{
	const int w = r.w, h = r.h, ncell = w * h;
	if (r.imin == r.imax)
	{
		for (int i = c.tid; i < NCOL_NGLDM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	uint32_t* P = mem;				// [level 0..Lmax][matches 0..8]
	uint32_t* Sg = P + (size_t)(Lmax + 1) * 9;	// [level]
	uint32_t* Sr = Sg + (Lmax + 1);			// [matches]
	uint8_t* aux = (uint8_t*)(Sr + 9);		// [ncell] match count of the cell (written in pass 1, read in passes 2-3)
	FastDiv fw;
	fw.init ((uint32_t)w);

	enum { Q_NS, Q_G, Q_DC, Q_N };
	unsigned long long q[Q_N];
	for (int k = 0; k < Q_N; k++)
		q[k] = 0;
	enum { A_LDE, A_HDE, A_LGLCE, A_HGLCE, A_LDLGLE, A_LDHGLE, A_HDLGLE, A_HDHGLE, A_N };
	double s[A_N];
	for (int k = 0; k < A_N; k++)
		s[k] = 0;
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const uint32_t cc = cell[i];
		if (!(cc & CELL_IN))
			continue;
		const uint32_t lv = cc & CELL_LEVEL;
		const int y = (int)fw.div ((uint32_t)i), x = i - y * w;
		const int nm = nbr_same_level (cell, w, h, x, y, lv, CELL_IN);	// ngldm.cpp:116-154
		aux[i] = (uint8_t)nm;
		atom_add (c, P + (size_t)lv * 9 + nm, 1u);
		atom_add (c, Sg + lv, 1u);
		atom_add (c, Sr + nm, 1u);
		const double g = (double)lv, dc = (double)(nm + 1);		// ngldm.cpp:283-305
		q[Q_NS] += 1;
		q[Q_G] += lv;
		q[Q_DC] += (unsigned long long)(nm + 1);
		const double d2 = dc * dc, g2 = g * g, rd = inv_sq_int ((uint32_t)(nm + 1));	// reciprocals (from the table) instead of the reference's chained divisions (<= 2 ulp)
		s[A_LDE] += rd;
		s[A_HDE] += d2;
		if (g != 0.0)
		{
			const double rg = inv_sq_int (lv);
			s[A_LGLCE] += rg;
			s[A_LDLGLE] += rd * rg;
			s[A_HDLGLE] += d2 * rg;
		}
		s[A_HGLCE] += g2;
		s[A_LDHGLE] += g2 * rd;
		s[A_HDHGLE] += d2 * g2;
	}
	c.sync();
	if (dbg && dbg->P)
		for (size_t i = c.tid; i < (size_t)(Lmax + 1) * 9; i += c.nthr)
			dbg->P[i] = P[i];
	cta_sum_u64<Q_N> (c, q);
	cta_sum<A_N> (c, s);
	const double Ns = (double)q[Q_NS];
	const double glm = (double)q[Q_G] / Ns, dcm = (double)q[Q_DC] / Ns;	// sum g*p, sum dc*p

	// ---- pass 2 (owner computes)
	unsigned long long q2[3] = { 0, 0, 0 };	// sum Sg^2, sum Sr^2, sum P^2
	double t[3] = { 0, 0, 0 };		// GLV (:330), DCV (:337), sum log(p) (:302-303)
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const uint32_t cc = cell[i];
		if (!(cc & CELL_IN))
			continue;
		const uint32_t lv = cc & CELL_LEVEL, nm = aux[i];
		const uint32_t v = P[(size_t)lv * 9 + nm];
		q2[0] += Sg[lv];
		q2[1] += Sr[nm];
		q2[2] += v;
		const double dg = (double)lv - glm, dd = (double)(nm + 1) - dcm;
		t[0] += dg * dg;
		t[1] += dd * dd;
		t[2] += ln_count (v);		// log(v / Ns), summed as ln(v) and finished below
	}
	cta_sum_u64<3> (c, q2);
	cta_sum<3> (c, t);
	c.sync();
	// ---- pass 3: restore the all-zero region
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const uint32_t cc = cell[i];
		if (!(cc & CELL_IN))
			continue;
		const uint32_t lv = cc & CELL_LEVEL, nm = aux[i];
		P[(size_t)lv * 9 + nm] = 0;
		Sg[lv] = 0;
	}
	for (int k = c.tid; k < 9; k += c.nthr)
		Sr[k] = 0;
	c.sync();
	if (c.tid == 0)
	{
		const double inv_ns = 1.0 / Ns;		// one reciprocal instead of 16 divisions (features are held to 1e-9, not bit-exact)
		out[N_LDE] = s[A_LDE] * inv_ns;
		out[N_HDE] = s[A_HDE] * inv_ns;
		out[N_LGLCE] = s[A_LGLCE] * inv_ns;
		out[N_HGLCE] = s[A_HGLCE] * inv_ns;
		out[N_LDLGLE] = s[A_LDLGLE] * inv_ns;
		out[N_LDHGLE] = s[A_LDHGLE] * inv_ns;
		out[N_HDLGLE] = s[A_HDLGLE] * inv_ns;
		out[N_HDHGLE] = s[A_HDHGLE] * inv_ns;
		out[N_GLNU] = (double)q2[0] * inv_ns;
		out[N_GLNUN] = (double)q2[0] * (inv_ns * inv_ns);
		out[N_DCNU] = (double)q2[1] * inv_ns;
		out[N_DCNUN] = (double)q2[1] * (inv_ns * inv_ns);
		out[N_DCP] = 1.0;						// :354
		out[N_GLM] = glm;
		out[N_GLV] = t[0] * inv_ns;
		out[N_DCM] = dcm;
		out[N_DCV] = t[1] * inv_ns;
		out[N_DCENT] = -((t[2] - Ns * log (Ns)) * inv_ns) * INV_LN2;
		out[N_DCENE] = (double)q2[2] * (inv_ns * inv_ns);
	}
}

// ------------------------------------------------------------------------------------------------
// NGTDM: N[L1], SI[L1][8] exact integer |level*nd - sum|, PR[L1] presence (zero on entry/exit), then the compacted
// list of present levels with their P and S
// ------------------------------------------------------------------------------------------------
// NDMAX = largest neighbour count (2D: 8; 3D radius 1: 26).  L1 = number of level slots.
NYX_HD size_t ngtdm_words (int L1, int ndmax)
{
	size_t ints = (size_t)L1 * (ndmax + 2) + ((size_t)L1 + 1) / 2;	// N, SI[ndmax], PR + level list (u16)
	ints += ints & 1;
	return ints + (size_t)L1 * 4;					// + Pl, Sl doubles
}
NYX_HD size_t ngtdm_smem_words (int Lmax)
{
	return ngtdm_words (Lmax + 1, 8);
}

struct NgtdmDebug
{
	uint32_t* N;	// [Lmax+1]
	double* S;	// [Lmax+1]
};

// Common tail of 2D and 3D NGTDM.  mem = N[L1] | SI[L1][ndmax] | PR[L1] (zero on exit) | list | Pl | Sl.
// lv0: first level that may be a centre level (levels below only count for presence).  nvc: number of kept centres.
// isize_fixed >= 0: |I| given by the caller (3D); else derived like ngtdm.cpp:57-67.
NYX_D void ngtdm_reduce (const Cta& c, const Params& prm, uint32_t* mem, int L1, int ndmax, int lv0, double nvc, int isize_fixed,
	bool ibsi_list, double* out, const NgtdmDebug* dbg)
{
	uint32_t* Nn = mem;				// [L1]
	uint32_t* SI = Nn + L1;				// [L1][ndmax]
	uint32_t* PR = SI + (size_t)L1 * ndmax;		// [L1] level present anywhere in the binned tile
	uint16_t* lst = (uint16_t*)(PR + L1);		// [L1] levels with N > 0, ascending
	size_t off = (size_t)L1 * (ndmax + 2) + ((size_t)L1 + 1) / 2;
	off += off & 1;
	double* Pl = (double*)(mem + off);		// [L1] P of the listed levels (:191-192)
	double* Sl = Pl + L1;				// [L1] S of the listed levels
	const double Nvc = nvc, Nvp = nvc;		// every kept pixel has a positive neighbourhood mean (:181-182)

	// ---- one sub-group sweeps the levels: Ngp, compaction of the present levels, S and P, and restores the zero region
	int* shared_cnt = (int*)c.red;	// [0] Ngp, [1] list length (block mode); a warp-wide context keeps them in registers
	int Ngp = 0, nl = 0;
	if (c.sub_id() == 0)
	{
		const Cta g = c.sub();
		for (int base = 0; base < L1; base += g.nthr)
		{
			const int lv = base + g.tid;
			uint32_t n = 0, pr = 0;
			double sv = 0;
			if (lv < L1)
			{
				n = Nn[lv];
				pr = PR[lv];
				for (int k = 0; k < ndmax; k++)
				{
					const uint32_t si = SI[(size_t)lv * ndmax + k];
					if (si)
						sv += ndmax <= 8 ? (double)si * k_weights ((uint32_t)k).inv1pk : (double)si / (double)(k + 1);	// 2D: 1 / (k + 1) from the table (<= 1 ulp)
					SI[(size_t)lv * ndmax + k] = 0;
				}
				Nn[lv] = 0;
				PR[lv] = 0;
				if (dbg && dbg->N)
				{
					dbg->N[lv] = n;
					dbg->S[lv] = sv;
				}
			}
			const bool keep = n != 0 && lv >= lv0;
			const unsigned bk = g.ballot (keep), bp = g.ballot (pr != 0);
			if (keep)
			{
				const int pos = nl + popc32 (bk & g.lanemask_lt());
				lst[pos] = (uint16_t)lv;
				Pl[pos] = (double)n / Nvc;
				Sl[pos] = sv;
			}
			nl += popc32 (bk);
			Ngp += popc32 (bp);	// unique levels in the tile (ngtdm.cpp:53-55,152)
		}
		if (!c.wmode && g.tid == 0)
		{
			shared_cnt[0] = Ngp;
			shared_cnt[1] = nl;
		}
	}
	c.sync();
#if NYX_DEVICE_BUILD
	if (!c.wmode)
	{
		Ngp = shared_cnt[0];
		nl = shared_cnt[1];
		__syncthreads();
	}
#endif
	const int Isize = isize_fixed >= 0 ? isize_fixed : (ibsi_list ? (Ngp ? L1 : 0) : Ngp);	// |I| (:57-67); IBSI: 0..max(D)
	if (Isize < 2)
	{
		for (int i = c.tid; i < NCOL_NGTDM; i += c.nthr)
			out[i] = prm.softnan;
		if (dbg && dbg->N)	// the reference returns before building the matrix (ngtdm.cpp:70-78)
			for (int lv = c.tid; lv < L1; lv += c.nthr)
			{
				dbg->N[lv] = 0;
				dbg->S[lv] = 0;
			}
		return;
	}
	enum { A_PS, A_S, A_CON, A_BUSY, A_CPLX, A_STR, A_N };
	double s[A_N];
	for (int k = 0; k < A_N; k++)
		s[k] = 0;
	for (int k = c.tid; k < nl; k += c.nthr)
	{
		s[A_PS] += Pl[k] * Sl[k];		// :231-233
		s[A_S] += Sl[k];			// :256-258
	}
	{
		// every term is symmetric in (a, b) and vanishes for a == b: sum over the pairs a < b only and double.  The
		// triangular index walks forward by nthr positions per step (row a holds the nl - 1 - a pairs (a, a+1 .. nl-1)).
		double t[4] = { 0, 0, 0, 0 };
		int a = 0, b = 1 + c.tid;
		while (a < nl - 1 && b >= nl)
		{
			b -= nl;
			a++;
			b += a + 1;
		}
		while (a < nl - 1)
		{
			const double pi = Pl[a], pj = Pl[b];
			const double iv = (double)lst[a], jv = (double)lst[b];
			const double d = iv - jv, d2 = d * d;
			t[0] += pi * pj * d2;								// :245-251
			t[1] += fabs (pi * iv - pj * jv);						// :283-287
			t[2] += fabs (d) * (pi * Sl[a] + pj * Sl[b]) / (pi + pj);			// :311-314
			t[3] += (pi + pj) * d2;								// :333-336
			b += c.nthr;
			while (a < nl - 1 && b >= nl)
			{
				b -= nl;
				a++;
				b += a + 1;
			}
		}
		s[A_CON] += 2.0 * t[0];
		s[A_BUSY] += 2.0 * t[1];
		s[A_CPLX] += 2.0 * t[2];
		s[A_STR] += 2.0 * t[3];
	}
	cta_sum<A_N> (c, s);
	c.sync();	// the lists may be rewritten by the next ROI
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

NYX_D void ngtdm_body (const Cta& c, const RoiRec& r, const Params& prm, const uint32_t* cell, uint32_t* mem, int Lmax,
	double* out, const NgtdmDebug* dbg)
{
	const int w = r.w, h = r.h, ncell = w * h;
	const int L1 = Lmax + 1;
	uint32_t* Nn = mem;				// [L1]
	uint32_t* SI = Nn + L1;				// [L1][8]
	uint32_t* PR = SI + (size_t)L1 * 8;		// [L1] level present anywhere in the binned tile
	FastDiv fw;
	fw.init ((uint32_t)w);

	unsigned long long q[1] = { 0 };		// Nvc
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const uint32_t lv = cell[i] & CELL_LEVEL;
		if (!lv)
			continue;
		PR[lv] = 1;
		const int y = (int)fw.div ((uint32_t)i), x = i - y * w;
		int nd = 0;
		uint32_t sum = 0;
		{
			// branch-free 8-neighbourhood (see nbr_same_level): a neighbour outside the tile reads the centre and counts as level 0
			const bool xl = x > 0, xr = x < w - 1, yu = y > 0, yd = y < h - 1;
#define NYX_NB(ok, off) { const bool k_ = (ok); const uint32_t q_ = k_ ? cell[i + (off)] & CELL_LEVEL : 0u; sum += q_; nd += q_ ? 1 : 0; }
			NYX_NB (yu && xl, -w - 1)
			NYX_NB (yu, -w)
			NYX_NB (yu && xr, -w + 1)
			NYX_NB (xl, -1)
			NYX_NB (xr, 1)
			NYX_NB (yd && xl, w - 1)
			NYX_NB (yd, w)
			NYX_NB (yd && xr, w + 1)
#undef NYX_NB
		}
		if (nd > 0)	// ngtdm.cpp:83-146
		{
			long long df = (long long)lv * nd - (long long)sum;
			atom_add (c, Nn + lv, 1u);
			atom_add (c, SI + (size_t)lv * 8 + (nd - 1), (uint32_t)(df < 0 ? -df : df));
			q[0]++;
		}
	}
	c.sync();
	cta_sum_u64<1> (c, q);
	ngtdm_reduce (c, prm, mem, L1, 8, 1, (double)q[0], -1, prm.ibsi != 0, out, dbg);
}

} // namespace nyx
