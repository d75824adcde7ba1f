// nyxus_b200 -- the distance-to-contour WEIGHTED 2D geometric moments (the second half of *SGEOMOMS* / *IGEOMOMS*) and the
// contour they need.  Reference: features/contour.cpp (the tracer), features/pixel.cpp:13-71 (the distance), features/
// 2d_geomoments_basic.cpp:33-54,104-133,145-156,193-229,262-346 (weights and moments).
//
// The contour has to be the reference's own, ORDER included: a pixel's weight is I * log (dist + 0.001) with dist taken by
// Pixel2::min_sqdist, a hill descent over the ORDERED contour that can settle next to the true minimum (pixel.cpp:36-70).
// So the tracer is restated step by step:
//   padded ROI image, border image                                   contour.cpp:634-655, 394-404
//   raster scan with the inside / outside state and the Moore trace  :408-502  (neighbour table :436-446)
//   border pixels that have a border neighbour -- with the reference's bounds x < w - 1, y < h - 1 on PADDED coordinates :509-577
//   X-crossing fix (a pixel whose N, S, W, E are all still in the set leaves it; sequential, raster order) :590-610
//   loops: from the first remaining pixel follow the best candidate -- straight neighbours W > N > E > S first, else diagonal
//          NW > NE > SE > SW (get_dial_pos :222-268); a tip with several candidates is stacked; a dead end closes the loop if
//          |dx| == 1 OR |dy| == 1 to the origin, else backtracks to the stacked tip, else the chain is dropped :298-380, 613-640
//   contour coordinates = PADDED coordinates + bounding-box origin, i.e. shifted by (+1, +1) against the pixel cloud :658-664
// The reference's list searches (std::find over the contour, std::list::remove) are replaced by flag bytes on the padded grid:
// the same decisions in O(cells + contour) instead of O(contour^2).
//
// B200 form: warp (CTA for oversize ROIs) per ROI, scratch in a global-memory slot.  Flags and the raster-ordered border list
// are built by all lanes (ballot compaction); the trace, the X-crossing fix and the loop walk are sequential by nature and run
// on one lane; the per-pixel distances, weights (float, like the reference's RealPixIntens) and the two moment passes are
// parallel again.
#pragma once
#include "nyx_common.h"
#include "nyx_features.h"

namespace nyx
{

constexpr int WM_NCOL = 10 + 7 + 7 + 7;	// one variant: WRM[10] | WCM[7] | WNCM[7] | WHU[7]
constexpr double WM_EPSILON = 0.001;	// 2d_geomoments.h:49

enum : uint32_t { CF_ROI = 1u, CF_BORDER = 2u, CF_U = 1u };

// padded cells of the largest ROI of a launch: (w + 2)(h + 2) <= cells + 4 max(w,h) + 4
NYX_HD unsigned long long wmom_padded_bound (unsigned long long maxcells, int maxwh)
{
	return maxcells + 4ull * (unsigned long long)maxwh + 4ull;
}
// scratch words of one ROI: F, G (one byte per padded cell each) + LC, K, stack (one word per cell each; LC and stack hold
// the two float weights of every bounding-box cell afterwards)
NYX_HD size_t wmom_scratch_words (unsigned long long maxcells, int maxwh)
{
	const unsigned long long pc = wmom_padded_bound (maxcells, maxwh);
	return (size_t)(2 * ((pc + 3) / 4) + 3 * maxcells);
}

// Sequential part (one lane): Moore trace into F, then -- after the parallel border-list step -- X-crossing fix and loops.
NYX_D void contour_trace_border (uint8_t* F, int w, int h)
{
	const int W2 = w + 2, pc = W2 * (h + 2);
	const int nbo[8] = { -1, -W2 - 1, -W2, -W2 + 1, 1, W2 + 1, W2, W2 - 1 };	// contour.cpp:436-446
	const int nbn[8] = { 7, 7, 1, 1, 3, 3, 5, 5 };
	bool inside = false;
	for (int p0 = 0; p0 < pc; p0++)
	{
		const uint32_t f = F[p0];
		const bool bi = (f & CF_BORDER) != 0, pi = (f & CF_ROI) != 0;
		if (bi && !inside)
			inside = true;
		else if (pi && inside)
			continue;
		else if (!pi && inside)
			inside = false;
		else if (pi && !inside)
		{
			F[p0] = (uint8_t)(f | CF_BORDER);
			int pos = p0, cl = 1, counter = 0, counter2 = 0;
			for (;;)
			{
				const int cp = pos + nbo[cl - 1], ncl = nbn[cl - 1];
				if (cp >= pc)
					break;
				if (F[cp] & CF_ROI)
				{
					if (cp == p0)
					{
						counter++;
						if (ncl == 1 || counter >= 3)
						{
							inside = true;
							break;
						}
					}
					cl = ncl;
					pos = cp;
					counter2 = 0;
					F[cp] = (uint8_t)(F[cp] | CF_BORDER);
				}
				else
				{
					cl = 1 + (cl % 8);
					if (counter2 > 8)
						break;
					counter2++;
				}
			}
		}
	}
}

// LC[0..nC): padded cell indices of the candidate contour pixels in raster order, G[cell] & CF_U set for each.  Returns the
// number of ordered contour pixels written to K (padded cell indices, the reference's order).
NYX_D int contour_loops (uint8_t* G, const uint32_t* LC, int nC, uint32_t* K, uint32_t* stack, int w)
{
	const int W2 = w + 2;
	int nU = nC;
	for (int k = 0; k < nC; k++)	// X-crossing fix, contour.cpp:590-610
	{
		const int p = (int)LC[k];
		if ((G[p - W2] & CF_U) && (G[p + W2] & CF_U) && (G[p - 1] & CF_U) && (G[p + 1] & CF_U))
		{
			G[p] = 0;
			nU--;
		}
	}
	int nk = 0, fp = 0;
	while (nU > 0)
	{
		while (!(G[LC[fp]] & CF_U))
			fp++;
		const int origin = (int)LC[fp], kstart = nk;
		const int ox = origin % W2, oy = origin / W2;
		int looplen = 0, sp = 0, tip = origin;
		K[nk++] = (uint32_t)origin;
		G[origin] = 0;
		nU--;
		while (nU > 0)
		{
			int cnt = 0, best = -1;
			// straight neighbours, ascending dial position (S -2, E 1, N 3, W 5): the last hit is the best step
			if (G[tip + W2] & CF_U) { cnt++; best = tip + W2; }
			if (G[tip + 1] & CF_U) { cnt++; best = tip + 1; }
			if (G[tip - W2] & CF_U) { cnt++; best = tip - W2; }
			if (G[tip - 1] & CF_U) { cnt++; best = tip - 1; }
			if (!cnt)	// diagonal neighbours (SW -3, SE -1, NE 2, NW 4)
			{
				if (G[tip + W2 - 1] & CF_U) { cnt++; best = tip + W2 - 1; }
				if (G[tip + W2 + 1] & CF_U) { cnt++; best = tip + W2 + 1; }
				if (G[tip - W2 + 1] & CF_U) { cnt++; best = tip - W2 + 1; }
				if (G[tip - W2 - 1] & CF_U) { cnt++; best = tip - W2 - 1; }
			}
			if (cnt > 1)
				stack[sp++] = (uint32_t)tip;
			if (!cnt)
			{
				int dx = tip % W2 - ox, dy = tip / W2 - oy;
				dx = dx < 0 ? -dx : dx;
				dy = dy < 0 ? -dy : dy;
				if (dx == 1 || dy == 1)		// contour.cpp:349-354
				{
					looplen++;
					break;
				}
				if (!sp)
				{
					looplen = 0;
					break;
				}
				tip = (int)stack[--sp];
			}
			else
			{
				looplen++;
				tip = best;
				K[nk++] = (uint32_t)tip;
				G[tip] = 0;
				nU--;
			}
		}
		if (!looplen)
			nk = kstart;	// a chain that did not close: dropped (its pixels stay consumed), contour.cpp:629-639
	}
	return nk;
}

// Pixel2::min_sqdist (pixel.cpp:36-70) of the point (px, py) against the ordered contour K (packed x | y << 16, n entries)
NYX_D double contour_min_sqdist (const uint32_t* K, long long n, int px, int py)
{
	if (n <= 0)
		return 0.0;
	auto sqd = [&](long long i)
	{
		const uint32_t k = K[i];
		const double dx = (double)(int)(k & 0xFFFFu) - (double)px, dy = (double)(int)(k >> 16) - (double)py;
		return dx * dx + dy * dy;
	};
	double ext = sqd (0);
	if (n == 1)
		return ext;	// (the reference's step is (int)(1 / log 1): nothing else is visited)
	long long a = 0, b = n, ei = 0;
	long long step = (long long)(int)((double)(b - a) / log ((double)(b - a)));
	do
	{
		for (long long i = a + step; i < b; i += step)
		{
			const double d = sqd (i);
			if (ext > d)
			{
				ext = d;
				ei = i;
			}
		}
		const long long stepL = ei >= step ? step : ei, stepR = ei + step < n ? step : n - ei;
		a = ei - stepL;
		b = ei + stepR;
		step = b - a <= 10 ? 1 : (long long)(int)((double)(b - a) / log ((double)(b - a)));
	}
	while (b - a > 2);
	return ext;
}

// out: 2 * WM_NCOL doubles -- the shape variant's 31 columns (variants & 1), then the intensity variant's (variants & 2).
// mem: wmom_scratch_words (cells bound, edge bound) words of scratch; maxcells: the bound the slot was sized with.
NYX_D void wmom_body (const Cta& c, const ImageView& im, const RoiRec& r, int variants, uint32_t* mem, unsigned long long maxcells,
	int maxwh, double* out)
{
	const int w = r.w, h = r.h, W2 = w + 2, H2 = h + 2, pc = W2 * H2, ncell = w * h;
	const unsigned long long pcb = wmom_padded_bound (maxcells, maxwh);
	uint8_t* F = (uint8_t*)mem;
	uint8_t* G = (uint8_t*)(mem + (pcb + 3) / 4);
	uint32_t* LC = mem + 2 * ((pcb + 3) / 4);
	uint32_t* K = LC + maxcells;
	uint32_t* stack = K + maxcells;
	const size_t origin = (size_t)r.y0 * im.W + (size_t)r.x0;
	FastDiv f2, fw;
	f2.init ((uint32_t)W2);
	fw.init ((uint32_t)w);

	// ---- padded ROI image (contour.cpp:634-655); every ROI pixel is non-blank there (intensity + 1)
	for (int p = c.tid; p < pc; p += c.nthr)
	{
		const int y = (int)f2.div ((uint32_t)p), x = p - y * W2;
		bool in = false;
		if (x >= 1 && x <= w && y >= 1 && y <= h)
			in = im.label[origin + (size_t)(y - 1) * im.W + (size_t)(x - 1)] == r.label;
		F[p] = in ? (uint8_t)CF_ROI : (uint8_t)0;
		G[p] = 0;
	}
	c.sync();
	if (c.tid == 0)
		contour_trace_border (F, w, h);
	c.sync();
	// ---- border pixels with a border neighbour, raster order (contour.cpp:509-577: the bounds are the reference's)
	int nC = 0;
	{
		const int nthr = c.nthr < 32 ? c.nthr : 32;	// one warp compacts (a CTA's other warps wait)
		if (c.tid < nthr)
			for (int base = 0; base < pc; base += nthr)
			{
				const int p = base + c.tid;
				bool keep = false;
				if (p < pc && (F[p] & CF_BORDER))
				{
					const int y = (int)f2.div ((uint32_t)p), x = p - y * W2;
					const bool xl = x > 0, xr = x < w - 1, yu = y > 0, yd = y < h - 1;
					keep = (xl && (F[p - 1] & CF_BORDER)) || (xr && (F[p + 1] & CF_BORDER)) || (yu && (F[p - W2] & CF_BORDER)) || (yd && (F[p + W2] & CF_BORDER))
						|| (xl && yu && (F[p - W2 - 1] & CF_BORDER)) || (xr && yu && (F[p - W2 + 1] & CF_BORDER))
						|| (xl && yd && (F[p + W2 - 1] & CF_BORDER)) || (xr && yd && (F[p + W2 + 1] & CF_BORDER));
				}
#if NYX_DEVICE_BUILD
				const unsigned bal = __ballot_sync (0xffffffffu, keep);
				if (keep)
				{
					LC[nC + __popc (bal & ((1u << c.tid) - 1u))] = (uint32_t)p;
					G[p] = (uint8_t)CF_U;
				}
				nC += __popc (bal);
#else
				if (keep)
				{
					LC[nC++] = (uint32_t)p;
					G[p] = (uint8_t)CF_U;
				}
#endif
			}
	}
	c.sync();
	// ---- X-crossing fix and loops on one lane; the count travels through K[maxcells - 1] ... the last word of the stack
	uint32_t* nk_slot = stack + (maxcells - 1);
	if (c.tid == 0)
	{
		const int nk0 = contour_loops (G, LC, nC, K, stack, w);
		*nk_slot = (uint32_t)nk0;
	}
	c.sync();
	const long long nk = (long long)*nk_slot;
	c.sync();
	for (long long k = c.tid; k < nk; k += c.nthr)	// padded cell index -> packed padded coordinates
	{
		const uint32_t p = K[k];
		const uint32_t y = f2.div (p), x = p - y * (uint32_t)W2;
		K[k] = x | (y << 16);
	}
	c.sync();

	// ---- weights (2d_geomoments_basic.cpp:33-54) and raw sums (:145-156, 262-274): LC / stack now hold the float weights
	float* wS = (float*)LC;
	float* wI = (float*)stack;
	double s[20];
	for (int k = 0; k < 20; k++)
		s[k] = 0;
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const int y = (int)fw.div ((uint32_t)i), x = i - y * w;
		const size_t g = origin + (size_t)y * im.W + (size_t)x;
		if (im.label[g] != r.label)
			continue;
		const double d2 = contour_min_sqdist (K, nk, x, y);
		const double lg = log (sqrt (d2) + WM_EPSILON);
		const float ws = (float)(1.0 * lg), wi = (float)((double)load_inten (im, g) * lg);
		wS[i] = ws;
		wI[i] = wi;
		const double xd = (double)x, yd = (double)y, x2 = xd * xd, x3 = x2 * xd, y2 = yd * yd, y3 = y2 * yd;
		for (int v = 0; v < 2; v++)
		{
			const double I = v ? (double)wi : (double)ws;
			double* q = s + v * 10;
			q[0] += I;			// 00
			q[1] += I * yd;			// 01
			q[2] += I * y2;			// 02
			q[3] += I * y3;			// 03
			q[4] += I * xd;			// 10
			q[5] += (I * xd) * yd;		// 11
			q[6] += (I * xd) * y2;		// 12
			q[7] += I * x2;			// 20
			q[8] += (I * x2) * yd;		// 21
			q[9] += I * x3;			// 30
		}
	}
	c.sync();
	cta_sum<20> (c, s);
	const double oxS = s[4] / s[0], oyS = s[1] / s[0], oxI = s[14] / s[10], oyI = s[11] / s[10];	// :166-171

	// ---- weighted central sums (:193-203, 313-322)
	double m[14];
	for (int k = 0; k < 14; k++)
		m[k] = 0;
	for (int i = c.tid; i < ncell; i += c.nthr)
	{
		const int y = (int)fw.div ((uint32_t)i), x = i - y * w;
		const size_t g = origin + (size_t)y * im.W + (size_t)x;
		if (im.label[g] != r.label)
			continue;
		for (int v = 0; v < 2; v++)
		{
			const double I = v ? (double)wI[i] : (double)wS[i];
			const double dx = (double)x - (v ? oxI : oxS), dy = (double)y - (v ? oyI : oyS);
			const double dx2 = dx * dx, dx3 = dx2 * dx, dy2 = dy * dy, dy3 = dy2 * dy;
			double* q = m + v * 7;
			q[0] += I * dy2;		// 02
			q[1] += I * dy3;		// 03
			q[2] += (I * dx) * dy;		// 11
			q[3] += (I * dx) * dy2;		// 12
			q[4] += I * dx2;		// 20
			q[5] += (I * dx2) * dy;		// 21
			q[6] += I * dx3;		// 30
		}
	}
	cta_sum<14> (c, m);
	c.sync();
	if (c.tid != 0)
		return;
	for (int v = 0; v < 2; v++)
	{
		if (!(variants & (1 << v)))
			continue;
		double* o = out + v * WM_NCOL;
		const double* q = s + v * 10;
		const double* mu = m + v * 7;
		for (int k = 0; k < 10; k++)
			o[k] = q[k];
		for (int k = 0; k < 7; k++)
			o[10 + k] = mu[k];
		// mu_pq / pow (m00, (p + q) / 2 + 1) (:221-229): exponent 2 for the second-order moments, 2.5 for the third-order
		// ones -- NaN for a negative weighted mass, like pow()
		const double m00 = q[0], p2 = m00 * m00, p25 = p2 * sqrt (m00);
		const double n02 = mu[0] / p2, n03 = mu[1] / p25, n11 = mu[2] / p2, n12 = mu[3] / p25, n20 = mu[4] / p2, n21 = mu[5] / p25, n30 = mu[6] / p25;
		o[17] = n02; o[18] = n03; o[19] = n11; o[20] = n12; o[21] = n20; o[22] = n21; o[23] = n30;
		const double a = n30 + n12, b = n21 + n03, e = n30 - 3.0 * n12, f = 3.0 * n21 - n03;	// :231-256
		o[24] = n20 + n02;
		o[25] = (n20 - n02) * (n20 - n02) + 4.0 * (n11 * n11);
		o[26] = e * e + f * f;
		o[27] = a * a + b * b;
		o[28] = e * a * (a * a - 3.0 * (b * b)) + f * b * (3.0 * (a * a) - b * b);
		o[29] = (n20 - n02) * (a * a - b * b) + 4.0 * n11 * a * b;
		o[30] = f * a * (a * a - 3.0 * (b * b)) - e * b * (3.0 * (a * a) - b * b);
	}
}

} // namespace nyx
