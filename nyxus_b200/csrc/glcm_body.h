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
// B200 design: one CTA per ROI.  The joint counts and the four integer marginals (row, column, sum, difference) are
// built by shared-memory integer atomics in ONE pass over the tile -- integers make the result order-independent and
// therefore run-to-run deterministic.  Every feature that only needs a marginal is evaluated over <= 2*Ng entries;
// only ASM / entropy / HXY1 / correlation's cross term / autocorrelation / max need the joint cells, and HXY2 needs
// the outer product of the non-empty marginals (reference :905-911).  All sums are FP64 with a fixed reduction tree.
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
			return (w[idx >> 1] >> ((idx & 1) * 16)) & 0xFFFFu;
		return w[idx];
	}
	static NYX_HD size_t words (size_t ncells) { return P16 ? (ncells + 1) / 2 : ncells; }
};

// shared-memory words needed by glcm_body for `na` angles held at once
template <bool P16> NYX_HD size_t glcm_smem_words (int Ng, int na)
{
	size_t joint = Counters<P16>::words ((size_t)Ng * Ng);
	size_t marg = (size_t)Ng * 5;			// R, C, SPM (Ng each) + SPP (2Ng)
	size_t pxy = (size_t)Ng * 2 * 2;		// px, py as doubles
	return (size_t)na * (joint + marg + pxy);
}

struct GlcmDebug
{
	uint32_t* joint;	// [4][Ng*Ng] dense dump (row = centre level-1, col = neighbour level-1), may be null
};

// out: NCOL_GLCM doubles of this ROI (angled block then _AVE block)
template <bool P16>
NYX_D void glcm_body (const Cta& c, const RoiRec& r, const Params& prm, const uint32_t* cell, uint32_t* mem, int Ng,
	double* out, const GlcmDebug* dbg)
{
	const int w = r.w, h = r.h, ncell = w * h;
	const double softnan = prm.softnan;
	constexpr int NA = 4;
	const int ave_order[G_NAVE] = NYX_GLCM_AVE_ORDER;

	// ---- blank ROI (glcm.cpp:27-95)
	{
		Binner Bm;
		Bm.init (prm.greydepth, r.imin, r.imax);
		if (Bm.bin (r.imin) == Bm.bin (r.imax))
		{
			for (int i = c.tid; i < NCOL_GLCM; i += c.nthr)
				out[i] = softnan;
			return;
		}
	}

	// ---- carve shared memory
	const size_t jw = Counters<P16>::words ((size_t)Ng * Ng);
	uint32_t* jointw = mem;				// [NA][jw]
	uint32_t* Rm = jointw + NA * jw;		// [NA][Ng]   row marginal (centre level)
	uint32_t* Cm = Rm + NA * Ng;			// [NA][Ng]   column marginal (neighbour level)
	uint32_t* SPM = Cm + NA * Ng;			// [NA][Ng]   |r-c| marginal
	uint32_t* SPP = SPM + NA * Ng;			// [NA][2Ng]  r+c marginal
	double* px = (double*)(SPP + NA * 2 * Ng);	// [NA][Ng] = C/N	(glcm.cpp:859-864: px[i] += P.xy(i,j) -> column marginal)
	double* py = px + NA * Ng;			// [NA][Ng] = R/N
	const size_t nwords = NA * (jw + 5 * (size_t)Ng);
	for (size_t i = c.tid; i < nwords; i += c.nthr)
		mem[i] = 0;
	c.sync();

	// ---- build (glcm.cpp:431-478)
	const int o = prm.glcm_offset;
	const int dxs[NA] = { o, o, 0, -o }, dys[NA] = { 0, o, o, o };
	const bool sym = prm.glcm_symmetric || prm.ibsi;
	unsigned long long npair[NA] = { 0, 0, 0, 0 };
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		uint32_t cc = cell[i];
		if (!(cc & CELL_NZ))
			continue;
		int row = i / w, col = i - row * w;
		int b = (int)(cc & CELL_LEVEL) - 1;
#pragma unroll
		for (int a = 0; a < NA; a++)
		{
			int r2 = row + dys[a], c2 = col + dxs[a];
			if (r2 < 0 || r2 >= h || c2 < 0 || c2 >= w)
				continue;
			uint32_t nc = cell[r2 * w + c2];
			if (!(nc & CELL_NZ))
				continue;
			int av = (int)(nc & CELL_LEVEL) - 1;
			Counters<P16> J { jointw + a * jw };
			int d = av > b ? av - b : b - av;
			J.inc (c, b * Ng + av);
			atom_add (c, Rm + a * Ng + b, 1u);
			atom_add (c, Cm + a * Ng + av, 1u);
			if (sym)
			{
				J.inc (c, av * Ng + b);
				atom_add (c, Rm + a * Ng + av, 1u);
				atom_add (c, Cm + a * Ng + b, 1u);
				atom_add (c, SPM + a * Ng + d, 2u);
				atom_add (c, SPP + a * 2 * Ng + av + b, 2u);
				npair[a] += 2;
			}
			else
			{
				atom_add (c, SPM + a * Ng + d, 1u);
				atom_add (c, SPP + a * 2 * Ng + av + b, 1u);
				npair[a] += 1;
			}
		}
	}
	c.sync();
	cta_sum_u64<NA> (c, npair);	// sum_p per angle (glcm.cpp:480-484)

	if (dbg && dbg->joint)
		for (int a = 0; a < NA; a++)
		{
			Counters<P16> J { jointw + a * jw };
			for (int i = c.tid; i < Ng * Ng; i += c.nthr)
				dbg->joint[(size_t)a * Ng * Ng + i] = J.get (i);
		}

	// ---- integer first moments of the marginals + px/py
	unsigned long long im[NA * 5];
	for (int k = 0; k < NA * 5; k++)
		im[k] = 0;
	for (int a = 0; a < NA; a++)
	{
		const double N = (double)npair[a];
		for (int k = c.tid; k < 2 * Ng; k += c.nthr)
		{
			if (k < Ng)
			{
				unsigned long long rv = Rm[a * Ng + k], cv = Cm[a * Ng + k], dv = SPM[a * Ng + k];
				im[a * 5 + 0] += rv * (unsigned long long)(k + 1);		// sum R[r]*I[r]
				im[a * 5 + 1] += cv * (unsigned long long)(k + 1);		// sum C[c]*I[c]
				im[a * 5 + 2] += dv * (unsigned long long)k;			// sum k*Pxmy
				im[a * 5 + 3] += dv * (unsigned long long)k * (unsigned long long)k;	// sum k^2*Pxmy
				px[a * Ng + k] = npair[a] ? (double)cv / N : 0.0;
				py[a * Ng + k] = npair[a] ? (double)rv / N : 0.0;
			}
			im[a * 5 + 4] += (unsigned long long)SPP[a * 2 * Ng + k] * (unsigned long long)(k + 2);	// sum (I_x+I_y)*Pxpy
		}
	}
	cta_sum_u64<NA * 5> (c, im);	// (also orders the px/py writes before the reads below)

	// ---- joint-cell pass: ASM, ACOR, JMAX (exact integers), HXY, HXY1, correlation cross term (FP64)
	unsigned long long ji[NA * 2];
	double jd[NA * 3];
	unsigned long long jmax[NA];
	for (int a = 0; a < NA; a++)
	{
		ji[a * 2] = ji[a * 2 + 1] = 0;
		jd[a * 3] = jd[a * 3 + 1] = jd[a * 3 + 2] = 0;
		jmax[a] = 0;
		if (!npair[a])
			continue;
		const double N = (double)npair[a];
		const double mr = (double)im[a * 5 + 0] / N, mc = (double)im[a * 5 + 1] / N;
		Counters<P16> J { jointw + a * jw };
		for (int i = c.tid; i < Ng * Ng; i += c.nthr)
		{
			unsigned long long cnt = J.get (i);
			if (!cnt)
				continue;
			int rr = i / Ng, cl = i - rr * Ng;
			ji[a * 2] += cnt * cnt;
			ji[a * 2 + 1] += cnt * (unsigned long long)(rr + 1) * (unsigned long long)(cl + 1);
			jmax[a] = cnt > jmax[a] ? cnt : jmax[a];
			double p = (double)cnt / N;
			jd[a * 3] += plogp (p, GLCM_EPSILON);									// HXY (glcm.cpp:868, 734-735, 1160-1161)
			jd[a * 3 + 1] += ddiv (dmul (p, fast_log10 (dadd (dmul (px[a * Ng + cl], py[a * Ng + rr]), GLCM_EPSILON))), LOG10_2);	// HXY1 (:869)
			jd[a * 3 + 2] += ((double)(rr + 1) - mr) * ((double)(cl + 1) - mc) * p;				// (:633)
		}
	}
	cta_sum_u64<NA * 2> (c, ji);
	cta_sum<NA * 3> (c, jd);
	for (int a = 0; a < NA; a++)
		jmax[a] = cta_max_u64 (c, jmax[a]);

	// ---- HXY2 over the outer product of the marginals (glcm.cpp:905-911)
	double hxy2[NA] = { 0, 0, 0, 0 };
	for (int a = 0; a < NA; a++)
	{
		if (!npair[a])
			continue;
		for (int i = c.tid; i < Ng * Ng; i += c.nthr)
		{
			int ii = i / Ng, jj = i - ii * Ng;
			double pxi = px[a * Ng + ii], pyj = py[a * Ng + jj];
			if (pxi == 0 || pyj == 0)
				continue;
			hxy2[a] += plogp (dmul (pxi, pyj), GLCM_EPSILON);
		}
	}
	cta_sum<NA> (c, hxy2);

	// ---- marginal pass, one angle at a time
	for (int a = 0; a < NA; a++)
	{
		if (!npair[a])
		{
			for (int f = c.tid; f < G_NANGLED; f += c.nthr)	// glcm.cpp:260-295
				out[f * 4 + a] = softnan;
			continue;
		}
		const double N = (double)npair[a];
		const double mr = (double)im[a * 5 + 0] / N;		// f_corr mr (:598-603), f_var mean (:657-665), JAVE (:1140-1146)
		const double mc = (double)im[a * 5 + 1] / N;		// f_corr mc (:605-610), by_row_mean (:531-536)
		const double diffavg = (double)im[a * 5 + 2] / N;	// f_difference_avg (:786-795), DIS (:1050-1052)
		// Ng of THIS ROI: MATLAB = grey depth; IBSI = max level of the tile = aux_max (glcm.cpp:412-419).
		// (the `Ng` argument may be a class-wide storage bound in IBSI mode)
		const double NgD = prm.ibsi ? (double)r.imax : (double)Ng;
		enum { M_S2R, M_S2C, M_JVAR, M_HX, M_IDM, M_ID, M_IDN, M_IDMN, M_IV, M_DENT, M_DVAR, M_SENT, M_CP, M_CS, M_CT, M_N };
		double m[M_N];
		for (int k = 0; k < M_N; k++)
			m[k] = 0;
		for (int k = c.tid; k < 2 * Ng; k += c.nthr)
		{
			if (k < Ng)
			{
				double lv = (double)(k + 1);
				double pr = py[a * Ng + k], pc = px[a * Ng + k];
				double dr = lv - mr, dc = lv - mc, dj = lv - mr;
				m[M_S2R] += pr * dr * dr;			// :612-618, f_var :667-675
				m[M_S2C] += pc * dc * dc;			// :621-627
				m[M_JVAR] += dj * dj * pc;			// :1194-1200 (column marginal, row-marginal mean)
				if (pc > 0)
					m[M_HX] += plogp (pc, GLCM_EPSILON);	// :873-876
				uint32_t dv = SPM[a * Ng + k];
				if (dv)
				{
					double pm = (double)dv / N, kd = (double)k;
					m[M_IDM] += pm / (1.0 + kd * kd);		// :685-687, HOM2 :1069
					m[M_ID] += pm / (1.0 + kd);			// :1096-1097, HOM1 :942
					m[M_IDN] += pm / (1.0 + kd / NgD);		// :1110-1111
					m[M_IDMN] += pm / (1.0 + (kd * kd) / (NgD * NgD));	// :1082-1084
					if (k >= 1)
						m[M_IV] += pm / (kd * kd);		// :1123-1128
					m[M_DENT] += plogp (pm, GLCM_EPSILON);		// :778-781
					double dd = kd - diffavg;
					m[M_DVAR] += dd * dd * pm;			// :753-765
				}
			}
			uint32_t sv = SPP[a * 2 * Ng + k];
			if (sv)
			{
				double pp = (double)sv / N;
				m[M_SENT] += plogp (pp, GLCM_EPSILON);			// :712-716
				double mm = (double)(k + 2) - mc - mc;
				m[M_CT] += mm * mm * pp;				// :1028-1036
				m[M_CS] += mm * mm * mm * pp;				// :1001-1009
				m[M_CP] += mm * mm * mm * mm * pp;			// :979-987
			}
		}
		cta_sum<M_N> (c, m);

		double f[G_NANGLED];
		f[G_ASM] = (double)ji[a * 2] / (N * N);				// :546-558
		f[G_ENERGY] = f[G_ASM];						// :916-932
		f[G_ACOR] = (double)ji[a * 2 + 1] / N;				// :947-966
		f[G_CLUPROM] = m[M_CP];
		f[G_CLUSHADE] = m[M_CS];
		f[G_CLUTEND] = m[M_CT];
		f[G_SUMVARIANCE] = m[M_CT];					// :324-326
		f[G_CONTRAST] = (double)im[a * 5 + 3] / N;			// :566-584
		double sr = sqrt (m[M_S2R]), sc = sqrt (m[M_S2C]), den = sr * sc;
		f[G_CORRELATION] = (den > 0.0) ? jd[a * 3 + 2] / den : softnan;	// :637-643
		f[G_DIFAVE] = diffavg;
		f[G_DIFENTRO] = -m[M_DENT];
		f[G_DIFVAR] = m[M_DVAR];
		f[G_DIS] = diffavg;
		f[G_ENTROPY] = -jd[a * 3];					// :722-739
		f[G_JE] = -jd[a * 3];						// :1149-1165
		f[G_HOM1] = m[M_ID];
		f[G_HOM2] = m[M_IDM];
		f[G_ID] = m[M_ID];
		f[G_IDN] = m[M_IDN];
		f[G_IDM] = m[M_IDM];
		f[G_IDMN] = m[M_IDMN];
		double im1 = (jd[a * 3] - jd[a * 3 + 1]) / m[M_HX];		// :880-883
		f[G_INFOMEAS1] = (isnan (im1) || isinf (im1)) ? softnan : im1;
		f[G_INFOMEAS2] = sqrt (fabs (1.0 - exp (-2.0 * (-hxy2[a] + jd[a * 3]))));	// :913
		f[G_IV] = m[M_IV];
		f[G_JAVE] = mr;
		f[G_JMAX] = (double)jmax[a] / N;				// :1167-1183
		f[G_JVAR] = m[M_JVAR];
		f[G_SUMAVERAGE] = (double)im[a * 5 + 4] / N;			// :693-704
		f[G_SUMENTROPY] = -m[M_SENT];
		f[G_VARIANCE] = m[M_S2R];					// :647-676
		if (c.tid == 0)
			for (int k = 0; k < G_NANGLED; k++)
				out[k * 4 + a] = f[k];
	}
	c.sync();

	// ---- _AVE = ((v0+v1)+v2)+v3)/4 (glcm.cpp:1205-1214)
	for (int k = c.tid; k < G_NAVE; k += c.nthr)
	{
		const double* v = out + ave_order[k] * 4;
		out[G_NANGLED * 4 + k] = (((v[0] + v[1]) + v[2]) + v[3]) / 4.0;
	}
}

} // namespace nyx
