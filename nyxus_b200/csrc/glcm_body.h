// nyxus_b200 -- per-ROI GLCM: 4 co-occurrence matrices + 30 Haralick features x (4 angles + _AVE).
//
// Semantics follow the reference's GLCMFeature (features/glcm.cpp):
//   blank test           :23-95      bin_pixel(min) == bin_pixel(max) with GLCM_GREYDEPTH (MATLAB binning even in IBSI mode)
//   offsets              :234-251    (dx,dy) = (o,0) (o,o) (0,o) (-o,o), y pointing down
//   matrix build         :431-478    skip if either ORIGINAL intensity is 0; GLCM.xy(a,b)++ i.e. row = centre level,
//                                    column = neighbour level; mirrored too iff symmetric / IBSI
//   marginals            :487-537    Pxpy[k=x+y], Pxmy[k=|x-y|], by_row_mean over the COLUMN marginal
//   features             :546-1202   (each formula is annotated below)
//   _AVE                 :1205-1214
//
// B200 design: one CTA per ROI, ONE WARP PER ANGLE (the four directions are independent after the tile load, so the
// only block-wide barriers are the ones around the tile).  Per angle:
//   pass 1  every (cell, neighbour) pair bumps the joint counter and the four integer marginals (row, column, sum,
//           difference) with shared-memory integer atomics -- integers make the matrix order-independent, hence
//           run-to-run deterministic; the integer first/second moments are accumulated per pair in registers.
//   margin  one sweep over <= 2*Ng marginal entries gives px/py, every marginal-only feature, and (by ballot
//           compaction) the lists of non-empty marginal levels.
//   pass 2  "owner computes": every pair re-reads the count of ITS joint cell and adds that cell's term divided by
//           the count, so  sum_cells f(cnt) == sum_pairs f(cnt)/cnt  needs no sweep over the Ng x Ng matrix
//           (ASM and autocorrelation stay exact integers; the entropies become (1/N) sum_pairs log(p_cell)).
//   pass 3  every pair zeroes its cell again: the matrix region is zero on entry AND on exit, so nothing of size
//           Ng^2 is ever cleared or scanned per ROI (the reference's cost is ~35 dense Ng^2 loops per angle).
//   HXY2    the outer product of the NON-EMPTY marginals only (reference :905-911).
// All FP64 sums use fixed lane assignments and shuffle trees => deterministic.
#pragma once
#include "nyx_common.h"
#include "nyx_features.h"

namespace nyx
{

// Counter storage: 16-bit packed pairs (tiles < 32768 cells can never overflow a cell) or plain 32-bit.
template <bool P16> struct Counters
{
	uint32_t* w;
	NYX_D void inc (const Cta& c, int idx) const
	{
		if (P16)
			atom_add (c, w + (idx >> 1), 1u << ((idx & 1) * 16));
		else
			atom_add (c, w + idx, 1u);
	}
	NYX_D uint32_t get (int idx) const
	{
		if (P16)
			return ((const uint16_t*)w)[idx];
		return w[idx];
	}
	NYX_D void clear (int idx) const
	{
		if (P16)
			((uint16_t*)w)[idx] = 0;
		else
			w[idx] = 0;
	}
	static NYX_HD size_t words (size_t ncells) { return P16 ? (ncells + 1) / 2 : ncells; }
};

// 32-bit words of one angle's region: joint + R,C,SPM (Ng each) + SPP (2Ng) [+pad] + px,py (doubles, by level) +
// the compacted non-empty px / py values (doubles)
template <bool P16> NYX_HD size_t glcm_angle_words (int Ng)
{
	size_t ints = Counters<P16>::words ((size_t)Ng * Ng) + (size_t)Ng * 5;
	ints += ints & 1;
	return ints + (size_t)Ng * 8;
}

// shared-memory words needed by glcm_body for `na` angles held at once
// (rad: + the level -> rank map [Ng + 1] and the rank -> level list [Ng] of PyRadiomics binning)
// (listcells > 0, 2D: + the 16-bit list of the tile's non-zero cells, the only cells that can be the centre of a pair)
template <bool P16> NYX_HD size_t glcm_smem_words (int Ng, int na, bool rad = false, uint32_t listcells = 0)
{
	return (size_t)na * glcm_angle_words<P16> (Ng) + (rad ? (size_t)2 * Ng + 2 : 0) + ((size_t)listcells + 1) / 2;
}

// PyRadiomics binning (glcm.cpp:388-398, 3d_glcm.cpp:286-296): I = the sorted distinct non-zero bins of the tile / cube;
// the matrix is indexed by rank in I.  Rewrites every non-zero cell level to rank + 1, fills Ilev[rank] = level and
// returns |I|.  n = |grey depth| (largest possible bin); map: n + 1 words of scratch.
NYX_D int glcm_rank_levels (const Cta& c, uint32_t* cell, int ncell, int n, uint32_t* map, uint32_t* Ilev)
{
	for (int l = c.tid; l <= n; l += c.nthr)
		map[l] = 0;
	c.sync();
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const uint32_t lv = cell[i] & CELL_LEVEL;
		if (lv)
			map[lv] = 1;
	}
	c.sync();
	if (c.sub_id() == 0)
	{
		const Cta g = c.sub();
		int base = 0;
		for (int l0 = 1; l0 <= n; l0 += g.nthr)
		{
			const int l = l0 + g.tid;
			const bool present = l <= n && map[l] != 0;
			const unsigned b = g.ballot (present);
			if (present)
			{
				const int rank = base + popc32 (b & g.lanemask_lt());
				map[l] = (uint32_t)rank + 1u;
				Ilev[rank] = (uint32_t)l;
			}
			base += popc32 (b);
		}
		if (g.tid == 0)
			map[0] = (uint32_t)base;
	}
	c.sync();
	const int cnt = (int)map[0];
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const uint32_t cc = cell[i], lv = cc & CELL_LEVEL;
		if (lv)
			cell[i] = (cc & ~CELL_LEVEL) | map[lv];
	}
	c.sync();
	return cnt;
}

// angled feature behind every _AVE column, in output order
#if NYX_DEVICE_BUILD
__constant__ int g_glcm_ave_order[G_NAVE] = NYX_GLCM_AVE_ORDER;
#else
static const int g_glcm_ave_order[G_NAVE] = NYX_GLCM_AVE_ORDER;
#endif

struct GlcmDebug
{
	uint32_t* joint;	// [4][Ng*Ng] dense dump (row = centre level-1, col = neighbour level-1), may be null
};

constexpr double INV_LOG10_2 = 1.0 / LOG10_2;

// sums over the marginals of one angle (shared by the per-pair ROI kernels and the dense large-ROI reducer)
enum GlcmM { M_S2R, M_S2C, M_JVAR, M_HX, M_IDM, M_ID, M_IDN, M_IDMN, M_IV, M_DENT, M_DVAR, M_SENT, M_CP, M_CS, M_CT, M_N };

// contribution of marginal index k (k < 2Ng): cv = C[k], rv = R[k], dv = SPM[k] (k < Ng), sv = SPP[k].
// lv = grey level of index k (k + 1 unless the levels are the ROI's sorted unique bins: PyRadiomics binning), kdl =
// kValuesDiff[k] (k likewise); clusters: CLUPROM / CLUSHADE / CLUTEND from the sum marginal (valid when level == index + 1)
NYX_D void glcm_marginal_terms (int k, int Ng, unsigned long long cv, unsigned long long rv, unsigned long long dv, unsigned long long sv,
	double N, double mr, double mc, double diffavg, double NgD, double* m, double* px, double* py, double lv, double kdl, bool clusters)
{
	const double invN = 1.0 / N;	// one reciprocal per marginal index instead of four divisions (features are held to 1e-9, not bit-exact)
	if (k < Ng)
	{
		const double pc = (double)cv * invN, pr = (double)rv * invN;
		px[k] = pc;
		py[k] = pr;
		const double dr = lv - mr, dc = lv - mc, dj = (double)(k + 1) - mr;
		m[M_S2R] += pr * dr * dr;			// :612-618, f_var :667-675
		m[M_S2C] += pc * dc * dc;			// :621-627
		m[M_JVAR] += dj * dj * pc;			// :1194-1200 (column marginal, row-marginal mean, INDEX + 1 not level)
		if (cv)
			m[M_HX] += dmul (pc, fast_log10 (dadd (pc, GLCM_EPSILON)));	// :873-876
		if (dv)
		{
			const double pm = (double)dv * invN, kd = (double)k;
			const KWeights kw = k_weights ((uint32_t)k);	// (reciprocal products instead of the reference's divisions: <= 2 ulp)
			m[M_IDM] += pm * kw.inv1pk2;			// pm / (1 + k^2)		:685-687, HOM2 :1069
			m[M_ID] += pm * kw.inv1pk;			// pm / (1 + k)			:1096-1097, HOM1 :942
			m[M_IDN] += pm * (NgD / (NgD + kd));		// pm / (1 + k / Ng)		:1110-1111
			m[M_IDMN] += pm * ((NgD * NgD) / (NgD * NgD + kd * kd));	// pm / (1 + k^2 / Ng^2)	:1082-1084
			if (k >= 1)
				m[M_IV] += kdl == kd ? pm * kw.invk2 : pm / (kdl * kdl);	// pm / kValuesDiff[k]^2	:1123-1128
			m[M_DENT] += dmul (pm, fast_log10 (dadd (pm, GLCM_EPSILON)));	// :778-781
			const double dd = kd - diffavg;
			m[M_DVAR] += dd * dd * pm;			// :753-765
		}
	}
	if (k < 2 * Ng && sv)
	{
		const double pp = (double)sv * invN;
		m[M_SENT] += dmul (pp, fast_log10 (dadd (pp, GLCM_EPSILON)));	// :712-716
		if (clusters)
		{
			const double mm = (double)(k + 2) - mc - mc;
			m[M_CT] += mm * mm * pp;				// :1028-1036
			m[M_CS] += mm * mm * mm * pp;				// :1001-1009
			m[M_CP] += mm * mm * mm * mm * pp;			// :979-987
		}
	}
}

// the 30 values of one angle from its sums.  sum_cnt2 = sum cnt^2, sum_rc = sum cnt*I_r*I_c, sum_d2 = sum k^2 Pxmy*N,
// sum_s = sum (I_x+I_y) Pxpy*N (integers held in doubles are exact below 2^53);  HXY.. already scaled by 1/LOG10_2;
// corr = sum (r-mr)(c-mc) p.  guard: the 2D softNaN guards (glcm.cpp:637-639,880-883; absent from 3d_glcm.cpp).
NYX_D void glcm_finish (double* f, const double* m, double N, double mr, double diffavg, double sum_cnt2, double sum_rc, double sum_d2,
	double sum_s, double jmax, double HXY, double HXY1, double HXY2, double HX, double corr, bool degenerate, bool guard, double softnan,
	double dis)
{
	// A single non-empty row or column makes every p(i,j) EXACTLY px(i)*py(j): the reference then sums identical terms
	// for HXY, HXY1 and HXY2 and gets HXY - HXY2 == 0 exactly (INFOMEAS2 = 0, not sqrt(1e-16)); keep that identity.
	if (degenerate)
		HXY1 = HXY2 = HXY;
	f[G_ASM] = sum_cnt2 / (N * N);					// :546-558
	f[G_ENERGY] = f[G_ASM];						// :916-932
	f[G_ACOR] = sum_rc / N;						// :947-966
	f[G_CLUPROM] = m[M_CP];
	f[G_CLUSHADE] = m[M_CS];
	f[G_CLUTEND] = m[M_CT];
	f[G_SUMVARIANCE] = m[M_CT];					// :324-326
	f[G_CONTRAST] = sum_d2 / N;					// :566-584
	const double sr = sqrt (m[M_S2R]), sc = sqrt (m[M_S2C]), den = sr * sc;
	f[G_CORRELATION] = (!guard || den > 0.0) ? corr / den : softnan;	// :637-643 (3d_glcm.cpp: unguarded)
	f[G_DIFAVE] = diffavg;
	f[G_DIFENTRO] = -m[M_DENT] * INV_LOG10_2;
	f[G_DIFVAR] = m[M_DVAR];
	f[G_DIS] = dis;						// :1050-1052 (index difference; == DIFAVE unless the levels are a sorted list)
	f[G_ENTROPY] = -HXY;						// :722-739
	f[G_JE] = -HXY;							// :1149-1165
	f[G_HOM1] = m[M_ID];
	f[G_HOM2] = m[M_IDM];
	f[G_ID] = m[M_ID];
	f[G_IDN] = m[M_IDN];
	f[G_IDM] = m[M_IDM];
	f[G_IDMN] = m[M_IDMN];
	const double im1 = (HXY - HXY1) / HX;				// :880-883
	f[G_INFOMEAS1] = (guard && (isnan (im1) || isinf (im1))) ? softnan : im1;
	f[G_INFOMEAS2] = sqrt (fabs (1.0 - exp (-2.0 * (-HXY2 + HXY))));	// :913
	f[G_IV] = m[M_IV];
	f[G_JAVE] = mr;
	f[G_JMAX] = jmax / N;						// :1167-1183
	f[G_JVAR] = m[M_JVAR];
	f[G_SUMAVERAGE] = sum_s / N;					// :693-704
	f[G_SUMENTROPY] = -m[M_SENT] * INV_LOG10_2;
	f[G_VARIANCE] = m[M_S2R];					// :647-676
}


// Pair enumeration.  2D (glcm.cpp:431-478): centre and neighbour must both have a non-zero ORIGINAL intensity.
// 3D (3d_glcm.cpp:333-372): every in-cube pair counts (MATLAB binning turns background into level 1) unless the
// binning keeps zeros (IBSI), where zero levels are skipped.
template <bool D3> struct GlcmGeom
{
	int w, h, d, dx, dy, dz;
	bool skipzero;
	bool identity;	// the levels are the intensities themselves (grey info 0: IBSI flag, or a 3D GLCM grey depth of 0)
	FastDiv fw, fh;
	NYX_D void init (const RoiRec& r, int dx_, int dy_, int dz_, bool skipzero_, bool identity_)
	{
		w = r.w; h = r.h; d = D3 ? r.d : 1;
		dx = dx_; dy = dy_; dz = dz_;
		skipzero = skipzero_;
		identity = identity_;
		fw.init ((uint32_t)w);
		fh.init ((uint32_t)h);
	}
	NYX_D int ncell() const { return w * h * d; }
	NYX_D bool pair (const uint32_t* cell, int i, int& b, int& av) const
	{
		const uint32_t cc = cell[i];
		uint32_t nc;
		if (!D3)
		{
			if (!(cc & CELL_NZ))
				return false;
			const int row = (int)fw.div ((uint32_t)i), col = i - row * w;
			const int r2 = row + dy, c2 = col + dx;
			if (r2 < 0 || r2 >= h || c2 < 0 || c2 >= w)
				return false;
			nc = cell[r2 * w + c2];
			if (!(nc & CELL_NZ))
				return false;
		}
		else
		{
			if (skipzero && !(cc & CELL_LEVEL))
				return false;
			const int row = (int)fw.div ((uint32_t)i), x = i - row * w;
			const int z = (int)fh.div ((uint32_t)row), y = row - z * h;
			const int x2 = x + dx, y2 = y + dy, z2 = z + dz;
			if (x2 < 0 || x2 >= w || y2 < 0 || y2 >= h || z2 < 0 || z2 >= d)
				return false;
			nc = cell[(z2 * h + y2) * w + x2];
			if (skipzero && !(nc & CELL_LEVEL))
				return false;
		}
		b = (int)(cc & CELL_LEVEL) - 1;
		av = (int)(nc & CELL_LEVEL) - 1;
		return true;
	}
};

// One angle, one warp-wide context g.  reg: this angle's region (all integer words zero on entry and on exit).
// The angle's 30 values go to out[f * ostride + a].  guard: 2D softNaN guards of CORRELATION / INFOMEAS1
// (glcm.cpp:637-639,880-883), absent from 3d_glcm.cpp.
// RAD (PyRadiomics binning, grey depth < 0; glcm.cpp:388-398,455-466): the matrix is indexed by the RANK of a level in the
// ROI's sorted list of distinct non-zero bins I[] (the caller has already rewritten the cell levels to rank + 1: see
// glcm_rank_levels), NgRoi = |I|; formulas that use I[] take the level from Ilev[rank], the index-based ones do not change.
template <bool P16, bool D3, bool RAD = false>
NYX_D void glcm_angle (const Cta& g, int a, const GlcmGeom<D3>& G, const RoiRec& r, const Params& prm, const uint32_t* cell,
	uint32_t* reg, int Ng, double* out, int ostride, const GlcmDebug* dbg, const uint32_t* Ilev = nullptr, int NgRoi = 0,
	const uint16_t* act = nullptr, int nact = 0)
{
	// act / nact (2D): the tile's non-zero cells in raster order -- a pair needs a non-zero centre, so the three
	// pair loops walk this list (every lane busy) instead of the whole bounding box (~25 % of it lies outside the ROI)
	const int ncell = act ? nact : G.ncell();
	const double softnan = prm.softnan;
	const size_t jw = Counters<P16>::words ((size_t)Ng * Ng);
	Counters<P16> J { reg };
	uint32_t* Rm = reg + jw;		// [Ng]   row marginal (centre level)
	uint32_t* Cm = Rm + Ng;			// [Ng]   column marginal (neighbour level)
	uint32_t* SPM = Cm + Ng;		// [Ng]   |r-c| marginal
	uint32_t* SPP = SPM + Ng;		// [2Ng]  r+c marginal
	size_t ints = jw + (size_t)Ng * 5;
	ints += ints & 1;
	double* px = (double*)(reg + ints);	// [Ng] = C/N	(glcm.cpp:859-864: px[i] += P.xy(i,j) -> column marginal)
	double* py = px + Ng;			// [Ng] = R/N
	double* pxc = py + Ng;			// [nx] px of the non-empty columns, ascending level
	double* pyc = pxc + Ng;			// [ny] py of the non-empty rows

	const bool sym = prm.glcm_symmetric || G.identity || RAD;	// glcm.cpp:472-476, 3d_glcm.cpp:365-367

	// ---- pass 1: build (glcm.cpp:431-478) + integer moments per pair
	enum { Q_N, Q_R, Q_C, Q_D, Q_D2, Q_S, Q_RC, Q_DL2, Q_CNT_ALL };
	constexpr int Q_CNT = RAD ? Q_CNT_ALL : Q_DL2;
	unsigned long long q[Q_CNT];
	for (int k = 0; k < Q_CNT; k++)
		q[k] = 0;
	for (int k = g.tid; k < ncell; k += g.nthr)
	{
		const int i = act ? (int)act[k] : k;
		int b, av;
		if (!G.pair (cell, i, b, av))
			continue;
		const int d = av > b ? av - b : b - av;
		const unsigned long long lb = RAD ? Ilev[b] : (unsigned long long)(b + 1), la = RAD ? Ilev[av] : (unsigned long long)(av + 1);
		const unsigned long long ld = (unsigned long long)d;
		if constexpr (RAD)
		{
			const unsigned long long dl = la > lb ? la - lb : lb - la;	// CONTRAST uses the LEVEL difference (:566-584)
			q[Q_DL2] += 2 * dl * dl;
		}
		J.inc (g, b * Ng + av);
		atom_add (g, Rm + b, 1u);
		atom_add (g, Cm + av, 1u);
		if (sym)
		{
			J.inc (g, av * Ng + b);
			atom_add (g, Rm + av, 1u);
			atom_add (g, Cm + b, 1u);
			atom_add (g, SPM + d, 2u);
			atom_add (g, SPP + av + b, 2u);
			q[Q_N] += 2;
			q[Q_R] += lb + la;
			q[Q_C] += lb + la;
			q[Q_D] += 2 * ld;
			q[Q_D2] += 2 * ld * ld;
			q[Q_S] += 2 * (lb + la);
			q[Q_RC] += 2 * lb * la;
		}
		else
		{
			atom_add (g, SPM + d, 1u);
			atom_add (g, SPP + av + b, 1u);
			q[Q_N] += 1;
			q[Q_R] += lb;		// sum_r R[r] * I[r]
			q[Q_C] += la;		// sum_c C[c] * I[c]
			q[Q_D] += ld;		// sum_k k * Pxmy[k]
			q[Q_D2] += ld * ld;	// sum_k k^2 * Pxmy[k]
			q[Q_S] += lb + la;	// sum_k (I_x + I_y) * Pxpy[k]
			q[Q_RC] += lb * la;	// sum cnt * I_r * I_c
		}
	}
	g.sync();
	cta_sum_u64<Q_CNT> (g, q);	// q[Q_N] = sum_p of this angle (glcm.cpp:480-484)

	if (dbg && dbg->joint)
		for (int i = g.tid; i < Ng * Ng; i += g.nthr)
		{
			if constexpr (RAD)
			{
				// the dump is by grey LEVEL (the oracle reports I[row], I[col]); the matrix is by rank
				const int row = i / Ng, col = i % Ng;
				if (row < NgRoi && col < NgRoi)
					dbg->joint[(size_t)a * Ng * Ng + (size_t)(Ilev[row] - 1) * Ng + (Ilev[col] - 1)] = J.get (i);
			}
			else
				dbg->joint[(size_t)a * Ng * Ng + i] = J.get (i);
		}

	if (!q[Q_N])	// glcm.cpp:260-295 (nothing was counted: the region is still all zero)
	{
		for (int f = g.tid; f < G_NANGLED; f += g.nthr)
			out[f * ostride + a] = softnan;
		return;
	}
	const double N = (double)q[Q_N];
	const double invN = 1.0 / N;			// every probability of this angle is count * invN (marginals, their lists, joint cells)
	const double mr = (double)q[Q_R] / N;		// f_corr mr (:598-603), f_var mean (:657-665), JAVE (:1140-1146)
	const double mc = (double)q[Q_C] / N;		// f_corr mc (:605-610), by_row_mean (:531-536)
	const double diffavg = (double)q[Q_D] / N;	// f_difference_avg (:786-795), DIS (:1050-1052)
	// Ng of THIS ROI: MATLAB = grey depth; IBSI = max level of the tile = aux_max (glcm.cpp:412-419).
	// (the `Ng` argument may be a class-wide storage bound in IBSI mode); PyRadiomics = number of distinct bins of the ROI
	const double NgD = RAD ? (double)NgRoi : G.identity ? (double)r.imax : (double)Ng;
	// kValuesDiff[k] / kValuesSum[k] (glcm.cpp:487-514): the LAST (x, y) of the x-outer / y-inner double loop with
	// |x - y| == k resp. x + y == k wins, i.e. x = NgRoi - 1 resp. x = min (k, NgRoi - 1); with I[i] = i + 1 these are k and k + 2
	double difavg_lv = diffavg, sumavg_lv = 0;
	if (RAD)
	{
		double pre[2] = { 0, 0 };
		for (int k = g.tid; k < 2 * NgRoi - 1; k += g.nthr)
		{
			if (k < NgRoi)
				pre[0] += (double)(Ilev[NgRoi - 1] - Ilev[NgRoi - 1 - k]) * ((double)SPM[k] / N);	// f_difference_avg :786-795
			const int x = k < NgRoi - 1 ? k : NgRoi - 1;
			pre[1] += (double)(Ilev[x] + Ilev[k - x]) * ((double)SPP[k] / N);			// f_savg :693-704
		}
		cta_sum<2> (g, pre);
		difavg_lv = pre[0];
		sumavg_lv = pre[1];
	}

	// ---- marginal sweep: px/py, marginal-only features, lists of the non-empty levels
	double m[M_N];
	for (int k = 0; k < M_N; k++)
		m[k] = 0;
	int nx = 0, ny = 0;
	unsigned long long cmax = 0, rmax = 0;
	for (int base = 0; base < 2 * Ng; base += g.nthr)
	{
		const int k = base + g.tid;
		uint32_t cv = 0, rv = 0, dv = 0, sv = 0;
		if (k < Ng)
		{
			cv = Cm[k];
			rv = Rm[k];
			dv = SPM[k];
		}
		if (k < 2 * Ng)
			sv = SPP[k];
		if (RAD)
			glcm_marginal_terms (k, Ng, cv, rv, dv, sv, N, mr, mc, difavg_lv, NgD, m, px, py, k < NgRoi ? (double)Ilev[k] : 0.0,
				k < NgRoi ? (double)(Ilev[NgRoi - 1] - Ilev[NgRoi - 1 - k]) : 0.0, false);
		else
			glcm_marginal_terms (k, Ng, cv, rv, dv, sv, N, mr, mc, diffavg, NgD, m, px, py, (double)(k + 1), (double)k, true);
		nx += popc32 (g.ballot (cv != 0));
		ny += popc32 (g.ballot (rv != 0));
		cmax = cv > cmax ? cv : cmax;
		rmax = rv > rmax ? rv : rmax;
	}
	cmax = cta_max_u64 (g, cmax);
	rmax = cta_max_u64 (g, rmax);
	// ---- the marginals as lists of DISTINCT values with multiplicities (for HXY2): px[i] = C[i] / N takes few distinct
	// values (a ROI of some hundred pixels over ~50 levels has column counts 1..~20), and the HXY2 term of a pair of levels
	// depends on (C[i], R[j]) only.  Counts below 2 Ng are bucketed in two tables that live in the list region itself;
	// lists: pxc[i] / pyc[j] = value, mxl[i] / myl[j] = number of levels that have it (in the marginal region, which is
	// dead after this point and zeroed again after HXY2).
	uint32_t* mxl = Rm;			// [<= Ng]
	uint32_t* myl = Rm + Ng;		// [<= Ng]
	uint32_t* cxl = Rm + 2 * Ng;		// [<= Ng] the distinct column counts
	uint32_t* cyl = Rm + 3 * Ng;		// [<= Ng] the distinct row counts
	int ndx = 0, ndy = 0;
	const int T = 2 * Ng;
	uint32_t* TX = (uint32_t*)pxc;		// [2 Ng] levels per column count
	uint32_t* TY = (uint32_t*)pyc;		// [2 Ng] levels per row count
	if (cmax < (unsigned long long)T && rmax < (unsigned long long)T)
	{
		for (int k = g.tid; k < 2 * T; k += g.nthr)
			TX[k] = 0;
		g.sync();
		for (int k = g.tid; k < Ng; k += g.nthr)
		{
			const uint32_t cv = Cm[k], rv = Rm[k];
			if (cv)
				atom_add (g, TX + cv, 1u);
			if (rv)
				atom_add (g, TY + rv, 1u);
		}
		g.sync();
		for (int base = 0; base < T; base += g.nthr)	// ascending count order: deterministic lists
		{
			const int cval = base + g.tid;
			const uint32_t tx = cval < T ? TX[cval] : 0u, ty = cval < T ? TY[cval] : 0u;
			const unsigned bx = g.ballot (tx != 0), by = g.ballot (ty != 0);
			g.sync();	// every lane has read Rm / Cm (previous loop) before the lists overwrite them
			if (tx)
			{
				const int pos = ndx + popc32 (bx & g.lanemask_lt());
				cxl[pos] = (uint32_t)cval;
				mxl[pos] = tx;
			}
			if (ty)
			{
				const int pos = ndy + popc32 (by & g.lanemask_lt());
				cyl[pos] = (uint32_t)cval;
				myl[pos] = ty;
			}
			ndx += popc32 (bx);
			ndy += popc32 (by);
		}
		g.sync();
		for (int i = g.tid; i < ndx; i += g.nthr)	// (the tables are dead: their region becomes the value lists)
			pxc[i] = (double)cxl[i] * invN;
		g.sync();
		for (int j = g.tid; j < ndy; j += g.nthr)
			pyc[j] = (double)cyl[j] * invN;
	}
	else
	{
		// large counts: one list entry per non-empty level, multiplicity 1
		uint32_t* tmpx = cxl;	// counts first (the marginals are overwritten by the lists)
		uint32_t* tmpy = cyl;
		for (int base = 0; base < Ng; base += g.nthr)
		{
			const int k = base + g.tid;
			const uint32_t cv = k < Ng ? Cm[k] : 0u, rv = k < Ng ? Rm[k] : 0u;
			const unsigned bx = g.ballot (cv != 0), by = g.ballot (rv != 0);
			g.sync();
			if (cv)
				tmpx[ndx + popc32 (bx & g.lanemask_lt())] = cv;
			if (rv)
				tmpy[ndy + popc32 (by & g.lanemask_lt())] = rv;
			ndx += popc32 (bx);
			ndy += popc32 (by);
		}
		g.sync();
		for (int i = g.tid; i < ndx; i += g.nthr)
		{
			pxc[i] = (double)tmpx[i] * invN;
			mxl[i] = 1;
		}
		for (int j = g.tid; j < ndy; j += g.nthr)
		{
			pyc[j] = (double)tmpy[j] * invN;
			myl[j] = 1;
		}
	}
	g.sync();	// px/py and the lists visible
	cta_sum<M_N> (g, m);

	// ---- pass 2: owner-computes terms of the joint cells
	unsigned long long s_asm = 0, jmax = 0;
	constexpr int NJD = RAD ? 6 : 3;
	double jd[NJD] = {};	// sum log(p_cell), sum log(px*py), sum (r-mr)(c-mc) [, RAD: cluster sums m^2, m^3, m^4 per pair]
	for (int k = g.tid; k < ncell; k += g.nthr)
	{
		const int i = act ? (int)act[k] : k;
		int b, av;
		if (!G.pair (cell, i, b, av))
			continue;
		const unsigned long long cnt = J.get (b * Ng + av);
		const double p = (double)cnt * invN;
		const double l1 = fast_log10 (dadd (p, GLCM_EPSILON));					// HXY (glcm.cpp:868, 734-735, 1160-1161)
		const double l2 = fast_log10 (dadd (dmul (px[av], py[b]), GLCM_EPSILON));		// HXY1 (:869)
		const double vb = RAD ? (double)Ilev[b] : (double)(b + 1), va = RAD ? (double)Ilev[av] : (double)(av + 1);
		double cr = (vb - mr) * (va - mc);								// (:633)
		if constexpr (RAD)
		{
			// CLUTEND / CLUSHADE / CLUPROM over the joint cells (:979-1036): the level sum is not a function of x + y
			const double mm = vb + va - mc - mc;
			jd[3] += 2.0 * mm * mm;
			jd[4] += 2.0 * mm * mm * mm;
			jd[5] += 2.0 * mm * mm * mm * mm;
		}
		if (sym)
		{
			// the mirrored cell (av,b) holds the same count; px == py
			s_asm += 2 * cnt;
			jd[0] += l1 + l1;
			jd[1] += l2 + fast_log10 (dadd (dmul (px[b], py[av]), GLCM_EPSILON));
			cr += (va - mr) * (vb - mc);
		}
		else
		{
			s_asm += cnt;
			jd[0] += l1;
			jd[1] += l2;
		}
		jd[2] += cr;
		jmax = cnt > jmax ? cnt : jmax;
	}
	{
		unsigned long long t[1] = { s_asm };
		cta_sum_u64<1> (g, t);	// (its shuffles/barrier also order pass 2's reads before pass 3's stores)
		s_asm = t[0];
	}
	cta_sum<NJD> (g, jd);
	jmax = cta_max_u64 (g, jmax);
	g.sync();

	// ---- pass 3: restore the all-zero region: a small joint matrix is cheaper to wipe with 128-bit stores than to
	// re-enumerate the pairs (Ng = 64, packed counters: 16 stores per lane), a large one is cleared cell by cell
#if NYX_DEVICE_BUILD
	if (jw * 4 <= (size_t)ncell * 24 && (jw & 3) == 0 && ((size_t)reg & 15) == 0)
	{
		uint4* j4 = reinterpret_cast<uint4*> (reg);
		for (int i = g.tid; i < (int)(jw >> 2); i += g.nthr)
			j4[i] = make_uint4 (0u, 0u, 0u, 0u);
	}
	else
#endif
	for (int k = g.tid; k < ncell; k += g.nthr)
	{
		const int i = act ? (int)act[k] : k;
		int b, av;
		if (!G.pair (cell, i, b, av))
			continue;
		J.clear (b * Ng + av);
		if (sym)
			J.clear (av * Ng + b);
	}

	// ---- HXY2 over the outer product of the non-empty marginals (glcm.cpp:905-911), one term per pair of DISTINCT
	// marginal values weighted by the number of level pairs that share it
	double hxy2[1] = { 0 };
	{
		// flat index k = ii * ndy + jj, advanced by nthr per step without a division
		const int tot = ndx * ndy, nyy = ndy > 0 ? ndy : 1;
		int ii = g.tid / nyy, jj = g.tid - ii * nyy;
		const int di = g.nthr / nyy, dj = g.nthr - di * nyy;
		for (int k = g.tid; k < tot; k += g.nthr)
		{
			const double qv = dmul (pxc[ii], pyc[jj]);
			hxy2[0] += (double)(mxl[ii] * myl[jj]) * dmul (qv, fast_log10 (dadd (qv, GLCM_EPSILON)));
			ii += di;
			jj += dj;
			while (jj >= nyy)
			{
				jj -= nyy;
				ii++;
			}
		}
	}
	cta_sum<1> (g, hxy2);
	g.sync();	// the lists have been read: the marginal region goes back to all zero
	for (int k = g.tid; k < 5 * Ng; k += g.nthr)
		Rm[k] = 0;
	g.sync();	// the region (incl. px/py/lists) may be reused by the next ROI

	// ---- features of this angle
	double f[G_NANGLED];
	double sum_d2 = (double)q[Q_D2];
	if constexpr (RAD)
	{
		m[M_CT] = jd[3] / N;
		m[M_CS] = jd[4] / N;
		m[M_CP] = jd[5] / N;
		sum_d2 = (double)q[Q_DL2];
	}
	glcm_finish (f, m, N, mr, difavg_lv, (double)s_asm, (double)q[Q_RC], sum_d2,
		RAD ? sumavg_lv * N : (double)q[Q_S], (double)jmax,
		jd[0] / N * INV_LOG10_2, jd[1] / N * INV_LOG10_2, hxy2[0] * INV_LOG10_2, m[M_HX] * INV_LOG10_2, jd[2] / N,
		nx == 1 || ny == 1, !D3, softnan, diffavg);
	if (g.tid == 0)
		for (int k = 0; k < G_NANGLED; k++)
			out[k * ostride + a] = f[k];
}

// out: NCOL_GLCM doubles of this ROI (angled block then _AVE block).  mem: 4 angle regions, zero on entry and exit.
template <bool P16>
NYX_D void glcm_body (const Cta& c, const RoiRec& r, const Params& prm, uint32_t* cell, uint32_t* mem, int Ng,
	double* out, const GlcmDebug* dbg, uint32_t listcells = 0)
{
	constexpr int NA = 4;
	const int* ave_order = g_glcm_ave_order;	// (a table in constant memory: a local array was re-initialised on the stack for every ROI)

	// ---- blank ROI (glcm.cpp:22-28): the ONLY use of GLCM_GREYDEPTH ("glcm/greydepth"); the matrix bins with GREYDEPTH (:354)
	{
		Binner Bm;
		Bm.init (prm.glcm_greydepth, r.imin, r.imax);
		if (Bm.bin (r.imin) == Bm.bin (r.imax))
		{
			for (int i = c.tid; i < NCOL_GLCM; i += c.nthr)
				out[i] = prm.softnan;
			return;
		}
	}
	const size_t aw = glcm_angle_words<P16> (Ng);
	const Cta g = c.sub();
	const int o = prm.glcm_offset;
	const bool rad = !prm.ibsi && prm.greydepth < 0;
	uint32_t* Ilev = mem + NA * aw + Ng + 1;
	const int ngroi = rad ? glcm_rank_levels (c, cell, r.w * r.h, Ng, mem + NA * aw, Ilev) : 0;
	// ---- the non-zero cells of the tile, raster order (one warp, ballot compaction: a deterministic list)
	uint16_t* act = nullptr;
	int nact = 0;
	const int ncell = r.w * r.h;
	if (listcells && ncell < 65536 && c.red)
	{
		act = (uint16_t*)(mem + NA * aw + (rad ? (size_t)2 * Ng + 2 : 0));
#if NYX_DEVICE_BUILD
		const int nchunk = (ncell + 31) >> 5;
		if (nchunk <= 32)
		{
			// up to 1024 cells: ALL warps build the list (one warp alone kept the other three at the barrier).  Chunk ch = 32
			// consecutive cells belongs to warp ch mod nw: count, barrier, every warp scans the <= 32 counts itself, then
			// writes its chunks at their offsets -- raster order, deterministic
			uint32_t* cnt = (uint32_t*)c.red;
			const int nw = c.sub_count(), wi = c.sub_id(), lane = g.tid;
			for (int ch = wi; ch < nchunk; ch += nw)
			{
				const int i = ch * 32 + lane;
				const unsigned b = g.ballot (i < ncell && (cell[i] & CELL_NZ) != 0);
				if (lane == 0)
					cnt[ch] = (uint32_t)popc32 (b);
			}
			c.sync();
			const uint32_t mine = lane < nchunk ? cnt[lane] : 0u;
			uint32_t inc = mine;
#pragma unroll
			for (int o2 = 1; o2 < 32; o2 <<= 1)
			{
				const uint32_t t = __shfl_up_sync (0xffffffffu, inc, o2);
				if (lane >= o2)
					inc += t;
			}
			nact = (int)__shfl_sync (0xffffffffu, inc, 31);
			for (int ch = wi; ch < nchunk; ch += nw)
			{
				const int i = ch * 32 + lane;
				const bool nz = i < ncell && (cell[i] & CELL_NZ) != 0;
				const unsigned b = g.ballot (nz);
				const uint32_t off = __shfl_sync (0xffffffffu, inc - mine, ch);
				if (nz)
					act[off + popc32 (b & g.lanemask_lt())] = (uint16_t)i;
			}
			c.sync();
		}
		else
		{
		if (c.sub_id() == 0)
		{
			int run = 0;
			for (int base = 0; base < ncell; base += g.nthr)
			{
				const int i = base + g.tid;
				const bool nz = i < ncell && (cell[i] & CELL_NZ) != 0;
				const unsigned b = g.ballot (nz);
				if (nz)
					act[run + popc32 (b & g.lanemask_lt())] = (uint16_t)i;
				run += popc32 (b);
			}
			if (g.tid == 0)
				c.red[0] = (double)run;
		}
		c.sync();
		nact = (int)c.red[0];
		c.sync();
		}
#else
		for (int i = 0; i < ncell; i++)
			if (cell[i] & CELL_NZ)
				act[nact++] = (uint16_t)i;
#endif
	}
	// the 4 x 30 angled values are staged in the reduction scratch (the angle warps never use it) instead of being written
	// to the table and read back for the averages (a global-memory round trip at the end of every ROI)
	double* st = c.red ? c.red : out;
	for (int a = c.sub_id(); a < NA; a += c.sub_count())
	{
		GlcmGeom<false> G;
		G.init (r, a == 2 ? 0 : (a == 3 ? -o : o), a == 0 ? 0 : o, 0, false, prm.ibsi != 0);	// glcm.cpp:234-251
		if (rad)
			glcm_angle<P16, false, true> (g, a, G, r, prm, cell, mem + (size_t)a * aw, Ng, st, 4, dbg, Ilev, ngroi, act, nact);
		else
			glcm_angle<P16, false> (g, a, G, r, prm, cell, mem + (size_t)a * aw, Ng, st, 4, dbg, nullptr, 0, act, nact);
	}
	c.sync();

	// ---- the angled block leaves as one coalesced store; _AVE = ((v0+v1)+v2)+v3)/4 (glcm.cpp:1205-1214) from the staged values
	if (st != out)
		for (int k = c.tid; k < G_NANGLED * 4; k += c.nthr)
			out[k] = st[k];
	for (int k = c.tid; k < G_NAVE; k += c.nthr)
	{
		const double* v = st + ave_order[k] * 4;
		out[G_NANGLED * 4 + k] = (((v[0] + v[1]) + v[2]) + v[3]) / 4.0;
	}
}

// ------------------------------------------------------------------------------------------------
// 3D GLCM (reference features/3d_glcm.cpp): 13 directions (:16-31) x static offset 1, all formulas as in 2D but without
// the CORRELATION / INFOMEAS1 guards; the table exports direction 0 of every angled code plus the 13-direction _AVE.
// out: the ROI's 3D GLCM block: [30][13] per-direction values, then 29 _AVE.
// ------------------------------------------------------------------------------------------------
constexpr int G3_NDIR = 13;
constexpr int NCOL_GLCM3 = G_NANGLED * G3_NDIR + G_NAVE;

// greyinfo: what the cube was binned with (GLCM_GREYDEPTH, 0 under the IBSI flag): > 0 MATLAB levels, 0 none,
// < 0 PyRadiomics bin count ("3glcm/greydepth=-N")
template <bool P16>
NYX_D void glcm3d_body (const Cta& c, const RoiRec& r, const Params& prm, uint32_t* cell, uint32_t* mem, int Ng,
	double* out, int greyinfo)
{
	const bool rad = greyinfo < 0;
	const int* ave_order = g_glcm_ave_order;	// (a table in constant memory: a local array was re-initialised on the stack for every ROI)
	const int sx[G3_NDIR] = { 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0 };
	const int sy[G3_NDIR] = { 1, 1, 1, 0, 0, 0, -1, -1, -1, 1, 1, 1, 0 };
	const int sz[G3_NDIR] = { 1, 0, -1, 1, 0, -1, 1, 0, -1, 1, 0, -1, 1 };
	const size_t aw = glcm_angle_words<P16> (Ng);
	const Cta g = c.sub();
	const int nsub = c.sub_count() < 4 ? c.sub_count() : 4;	// at most 4 regions
	uint32_t* Ilev = mem + 4 * aw + Ng + 1;
	const int ngroi = rad ? glcm_rank_levels (c, cell, r.w * r.h * r.d, Ng, mem + 4 * aw, Ilev) : 0;
	if (c.sub_id() < nsub)
		for (int a = c.sub_id(); a < G3_NDIR; a += nsub)
		{
			GlcmGeom<true> G;
			G.init (r, sx[a], sy[a], sz[a], greyinfo <= 0, greyinfo == 0);	// zero levels are skipped unless the binning is MATLAB's (3d_glcm.cpp:333-350)
			if (rad)
				glcm_angle<P16, true, true> (g, a, G, r, prm, cell, mem + (size_t)c.sub_id() * aw, Ng, out, G3_NDIR, nullptr, Ilev, ngroi);
			else
				glcm_angle<P16, true> (g, a, G, r, prm, cell, mem + (size_t)c.sub_id() * aw, Ng, out, G3_NDIR, nullptr);
		}
	c.sync();
	for (int k = c.tid; k < G_NAVE; k += c.nthr)	// calc_ave (texture_feature.h:121-130)
	{
		const double* v = out + ave_order[k] * G3_NDIR;
		double sum = 0;
		for (int a = 0; a < G3_NDIR; a++)
			sum += v[a];
		out[G_NANGLED * G3_NDIR + k] = sum / (double)G3_NDIR;
	}
}

} // namespace nyx
