// nyxus_b200 -- 2D geometric moments, the UNWEIGHTED half of *SGEOMOMS* (shape: every ROI pixel weighs 1) and *IGEOMOMS*
// (intensity-weighted) -- reference features/2d_geomoments_basic.cpp:104-376, 2d_geomoments.cpp:14-330:
//   raw        m_pq  = sum I x^p y^q,  x, y relative to the ROI's bounding-box origin           (:136-143, :268-283)
//   origin     (m10 / m00, m01 / m00)                                                           (:157-164)
//   central    mu_pq = sum I (x - ox)^p (y - oy)^q, p, q in 0..3                                 (:180-190, :300-318)
//   normalised raw     w_pq  = m_pq  / m00^((p + q) / 2 + 1), p, q in 0..3                       (:206-211, :352-370)
//   normalised central nu_pq = mu_pq / m00^((p + q) / 2 + 1) for 02 03 11 12 20 21 30            (:214-219, :329-338)
//   Hu's seven invariants of nu                                                                  (:231-256)
// The distance-to-contour WEIGHTED moments of both groups need the reference's contour tracer (contour.cpp) and are not
// built: their feature names, and the groups that contain them, are refused by name.
//
// B200 form: warp (CTA for oversize ROIs) per ROI, two walks over the bounding box on the resident slide: the 16 raw
// sums as exact 64-bit integers (doubles when x^3 y^3 I n could overflow), then the 16 central sums in FP64 with a fixed
// reduction tree.  Everything after that is a few dozen scalar operations.
#pragma once
#include "nyx_common.h"
#include "nyx_features.h"

namespace nyx
{

// one variant (shape or intensity): GM_NCOL doubles at `out`: RM[13] | CM[16] | NRM[16] | NCM[7] | HU[7]
template <bool INTEN>
NYX_D void geomom_variant (const Cta& c, const ImageView& im, const RoiRec& r, double* out)
{
	const uint32_t w = (uint32_t)r.w, h = (uint32_t)r.h, ncell = w * h;
	FastDiv fw;
	fw.init (w);
	const size_t origin = (size_t)r.y0 * im.W + (size_t)r.x0;
	const bool wide = (unsigned long long)ncell >= (1ull << 32) / (w > 1 ? w : 1);
	auto rowcol = [&](uint32_t i, uint32_t& y, uint32_t& x)
	{
		y = wide ? i / w : fw.div (i);
		x = i - y * w;
	};
	const double wmax = (double)(w > 1 ? w - 1 : 1), hmax = (double)(h > 1 ? h - 1 : 1);
	const bool exact = wmax * wmax * wmax * hmax * hmax * hmax * (INTEN ? (double)(r.imax ? r.imax : 1u) : 1.0) * (double)r.area < 1.8e19;

	// ---- raw moments m[p][q] = sum I x^p y^q
	double m[4][4];
	{
		unsigned long long qi[16];
		double qd[16];
		for (int k = 0; k < 16; k++)
		{
			qi[k] = 0;
			qd[k] = 0;
		}
		for (uint32_t i = c.tid; i < ncell; i += c.nthr)
		{
			uint32_t y, x;
			rowcol (i, y, x);
			const size_t g = origin + (size_t)y * im.W + x;
			if (im.label[g] != r.label)
				continue;
			const unsigned long long v = INTEN ? (unsigned long long)load_inten (im, g) : 1ull;
			if (exact)
			{
				unsigned long long xp = v;
#pragma unroll
				for (int p = 0; p < 4; p++)
				{
					unsigned long long t = xp;
#pragma unroll
					for (int q = 0; q < 4; q++)
					{
						qi[p * 4 + q] += t;
						t *= y;
					}
					xp *= x;
				}
			}
			else
			{
				double xp = (double)v;
#pragma unroll
				for (int p = 0; p < 4; p++)
				{
					double t = xp;
#pragma unroll
					for (int q = 0; q < 4; q++)
					{
						qd[p * 4 + q] += t;
						t *= (double)y;
					}
					xp *= (double)x;
				}
			}
		}
		if (exact)
		{
			cta_sum_u64<16> (c, qi);
			for (int k = 0; k < 16; k++)
				m[k >> 2][k & 3] = (double)qi[k];
		}
		else
		{
			cta_sum<16> (c, qd);
			for (int k = 0; k < 16; k++)
				m[k >> 2][k & 3] = qd[k];
		}
	}
	const double m00 = m[0][0];
	const double ox = m[1][0] / m00, oy = m[0][1] / m00;

	// ---- central moments mu[p][q]
	double mu[16];
	for (int k = 0; k < 16; k++)
		mu[k] = 0;
	for (uint32_t i = c.tid; i < ncell; i += c.nthr)
	{
		uint32_t y, x;
		rowcol (i, y, x);
		const size_t g = origin + (size_t)y * im.W + x;
		if (im.label[g] != r.label)
			continue;
		const double v = INTEN ? (double)load_inten (im, g) : 1.0;
		const double dx = (double)x - ox, dy = (double)y - oy;
		double xp = v;
#pragma unroll
		for (int p = 0; p < 4; p++)
		{
			double t = xp;
#pragma unroll
			for (int q = 0; q < 4; q++)
			{
				mu[p * 4 + q] += t;
				t *= dy;
			}
			xp *= dx;
		}
	}
	cta_sum<16> (c, mu);
	if (c.tid != 0)
		return;

	// m00^((p + q) / 2 + 1) for p + q = 0..6: powers of sqrt (m00) (pow() within 2 ulp)
	const double rt = sqrt (m00);
	double pw[7];
	pw[0] = m00;
	for (int k = 1; k < 7; k++)
		pw[k] = pw[k - 1] * rt;
	double* RM = out;
	double* CM = RM + 13;
	double* NRM = CM + 16;
	double* NCM = NRM + 16;
	double* HU = NCM + 7;
	// raw: 00 01 02 03 10 11 12 13 20 21 22 23 30
	{
		int k = 0;
		for (int p = 0; p < 4; p++)
			for (int q = 0; q < 4; q++)
				if (p < 3 || q == 0)
					RM[k++] = m[p][q];
	}
	for (int p = 0; p < 4; p++)
		for (int q = 0; q < 4; q++)
		{
			CM[p * 4 + q] = mu[p * 4 + q];
			NRM[p * 4 + q] = m[p][q] / pw[p + q];
		}
	const double n02 = mu[0 * 4 + 2] / pw[2], n03 = mu[0 * 4 + 3] / pw[3], n11 = mu[1 * 4 + 1] / pw[2], n12 = mu[1 * 4 + 2] / pw[3];
	const double n20 = mu[2 * 4 + 0] / pw[2], n21 = mu[2 * 4 + 1] / pw[3], n30 = mu[3 * 4 + 0] / pw[3];
	NCM[0] = n02; NCM[1] = n03; NCM[2] = n11; NCM[3] = n12; NCM[4] = n20; NCM[5] = n21; NCM[6] = n30;
	// Hu 1962 (2d_geomoments_basic.cpp:231-256)
	const double a = n30 + n12, b = n21 + n03, e = n30 - 3.0 * n12, f = 3.0 * n21 - n03;
	HU[0] = n20 + n02;
	HU[1] = (n20 - n02) * (n20 - n02) + 4.0 * (n11 * n11);
	HU[2] = e * e + f * f;
	HU[3] = a * a + b * b;
	HU[4] = e * a * (a * a - 3.0 * (b * b)) + f * b * (3.0 * (a * a) - b * b);
	HU[5] = (n20 - n02) * (a * a - b * b) + 4.0 * n11 * a * b;
	HU[6] = f * a * (a * a - 3.0 * (b * b)) - e * b * (3.0 * (a * a) - b * b);
}

// out: NCOL_GEOMOM doubles: the shape variant's 59 columns, then the intensity variant's
NYX_D void geomom_body (const Cta& c, const ImageView& im, const RoiRec& r, int variants, double* out)
{
	if (variants & 1)
		geomom_variant<false> (c, im, r, out);
	c.sync();
	if (variants & 2)
		geomom_variant<true> (c, im, r, out + GM_NCOL);
}

} // namespace nyx
