// nyxus_b200 -- *BASIC_MORPHOLOGY* (reference features/basic_morphology.cpp:20-95, featureset.h:46-60): area, bounding box,
// centroid, intensity-weighted centroid, mass displacement, compactness, equal-area diameter, extent, aspect ratio.
//
// Everything except COMPACTNESS and the weighted centroid already sits in the ROI record the phase-1 scan produced
// (area, AABB).  The rest is three walks over the ROI's AABB on the resident slide (label compare), warp or CTA wide:
//   A  exact integer sums  sum x, sum y, sum I, sum x I, sum y I   (coordinates relative to the AABB origin)
//   B  sum d,   d = distance of a pixel to the centroid             (fixed-tree FP64)
//   C  sum (d - mean d)^2                                           (Moments2::std, moments.h:13-43: sqrt (M2 / (n - 1)), 0 for n <= 2)
// The reference accumulates the same sums sequentially in double; features are held to 1e-9 relative.
//
// Reference quirk kept: AREA_UM2 is only written when settings[XYRES] > 0, and Environment::compile_feature_settings
// (env_features.cpp:712-736) never fills XYRES, so the column is 0 whatever pixels_per_micron is.
#pragma once
#include "nyx_common.h"
#include "nyx_features.h"

namespace nyx
{

// out: NCOL_MORPH doubles in Feature2D code order
NYX_D void morph_body (const Cta& c, const ImageView& im, const RoiRec& r, const Params& prm, double* out)
{
	(void)prm;
	const uint32_t w = (uint32_t)r.w, h = (uint32_t)r.h, ncell = w * h;
	FastDiv fw;
	fw.init (w);
	const size_t origin = (size_t)r.y0 * im.W + (size_t)r.x0;
	const bool wide = (unsigned long long)ncell >= (1ull << 32) / (w > 1 ? w : 1);	// FastDiv needs n * d < 2^32
	auto rowcol = [&](uint32_t i, uint32_t& y, uint32_t& x)
	{
		y = wide ? i / w : fw.div (i);
		x = i - y * w;
	};

	// ---- A: integer sums
	enum { Q_X, Q_Y, Q_I, Q_XI, Q_YI, Q_N };
	unsigned long long q[Q_N] = { 0, 0, 0, 0, 0 };
	// 64-bit sums hold w * Imax * area; beyond that (giant ROI of 32-bit intensities) fall back to FP64 accumulation
	const bool exact = (double)(w > h ? w : h) * (double)(r.imax ? r.imax : 1u) * (double)r.area < 1.8e19;
	double qd[2] = { 0, 0 };
	for (uint32_t i = c.tid; i < ncell; i += c.nthr)
	{
		uint32_t y, x;
		rowcol (i, y, x);
		const size_t g = origin + (size_t)y * im.W + x;
		if (im.label[g] != r.label)
			continue;
		const unsigned long long v = load_inten (im, g);
		q[Q_X] += x;
		q[Q_Y] += y;
		q[Q_I] += v;
		if (exact)
		{
			q[Q_XI] += (unsigned long long)x * v;
			q[Q_YI] += (unsigned long long)y * v;
		}
		else
		{
			qd[0] += (double)x * (double)v;
			qd[1] += (double)y * (double)v;
		}
	}
	cta_sum_u64<Q_N> (c, q);
	cta_sum<2> (c, qd);
	const double n = (double)r.area;
	// absolute pixel coordinates: x = column, y = row (pixel.h)
	const double cx = (double)r.x0 + (double)q[Q_X] / n, cy = (double)r.y0 + (double)q[Q_Y] / n;
	const double mass = (double)q[Q_I];
	const double xm = exact ? (double)q[Q_XI] : qd[0], ym = exact ? (double)q[Q_YI] : qd[1];
	double wcx = 0.0, wcy = 0.0;
	if (mass > 0)
	{
		// sum (x0 + x) I = x0 * mass + sum x I
		wcx = ((double)r.x0 * mass + xm) / mass;
		wcy = ((double)r.y0 * mass + ym) / mass;
	}

	// ---- B, C: distances to the centroid
	const double rcx = cx - (double)r.x0, rcy = cy - (double)r.y0;
	double sd[1] = { 0 };
	for (uint32_t i = c.tid; i < ncell; i += c.nthr)
	{
		uint32_t y, x;
		rowcol (i, y, x);
		if (im.label[origin + (size_t)y * im.W + x] != r.label)
			continue;
		const double dx = (double)x - rcx, dy = (double)y - rcy;
		sd[0] += sqrt (dx * dx + dy * dy);
	}
	cta_sum<1> (c, sd);
	const double dmean = sd[0] / n;
	double m2[1] = { 0 };
	for (uint32_t i = c.tid; i < ncell; i += c.nthr)
	{
		uint32_t y, x;
		rowcol (i, y, x);
		if (im.label[origin + (size_t)y * im.W + x] != r.label)
			continue;
		const double dx = (double)x - rcx, dy = (double)y - rcy;
		const double t = sqrt (dx * dx + dy * dy) - dmean;
		m2[0] += t * t;
	}
	cta_sum<1> (c, m2);

	if (c.tid == 0)
	{
		const double sdev = r.area > 2 ? sqrt (m2[0] / (n - 1.0)) : 0.0;
		const double ddx = wcx - cx, ddy = wcy - cy;
		out[MO_AREA_PIXELS_COUNT] = n;
		out[MO_AREA_UM2] = 0.0;		// see the header note (XYRES is never compiled into the settings)
		out[MO_CENTROID_X] = cx;
		out[MO_CENTROID_Y] = cy;
		out[MO_WEIGHTED_CENTROID_Y] = wcy;
		out[MO_WEIGHTED_CENTROID_X] = wcx;
		out[MO_MASS_DISPLACEMENT] = sqrt (ddx * ddx + ddy * ddy);
		out[MO_COMPACTNESS] = sdev / n;
		out[MO_BBOX_YMIN] = (double)r.y0;
		out[MO_BBOX_XMIN] = (double)r.x0;
		out[MO_BBOX_HEIGHT] = (double)r.h;
		out[MO_BBOX_WIDTH] = (double)r.w;
		out[MO_DIAMETER_EQUAL_AREA] = 2.0 * sqrt (n / 3.14159265358979323846);
		out[MO_EXTENT] = n / ((double)r.w * (double)r.h);
		out[MO_ASPECT_RATIO] = (double)r.w / (double)r.h;
	}
}

} // namespace nyx
