// nyxus_b200 -- IBSI intensity-histogram family *ALL_IH* (46 features; reference features/intensity_histogram.cpp:29-277,
// featureset.h:590-637).  Everything derives from ONE N-bin histogram over the ROI's [min, max], N = grey depth:
//   bin of a value   floor ((v - min) / binW) clamped to [0, N-1], binW = (max - min) / N            (:44-50, :64-70)
//   median           centre of the bin where the running count first EXCEEDS count / 2 (integer division) (:114-127)
//   p10/p25/p75/p90  quantiles interpolated inside a bin, walked from the low end (p < 0.5) or the high end (:73-110)
//   "..._IDX"        the same statistics over bin indices (reported 1-based where the reference adds 1)  (:236-257)
//   gradients        central differences of the bin counts, extremes seeded with DBL_MIN / DBL_MAX      (:166-170,209-227)
// Gate: IBSI mode only (env_features.cpp:513-527 strips the family otherwise); blank / degenerate ROI (max <= min) and
// N < 2 give `noval` everywhere (:285-296).  Integer images only (pscale = 1, poffset = 0, :312-350).
//
// B200 form: warp (or CTA) per ROI on the ROI's CSR slice: the histogram is built by all lanes with shared-memory atomics
// (exact bin index: the reference's double division, taken without dividing unless the quotient sits next to an
// integer), the N-bin passes that follow are serial by nature (running sums, first-exceeds searches) and run on one lane.
#pragma once
#include "nyx_common.h"
#include "nyx_features.h"

namespace nyx
{

// shared-memory words: N bin counts
NYX_HD size_t ih_smem_words (int nb)
{
	nb = nb < 0 ? -nb : nb;
	return (size_t)nb + 2;
}

NYX_D void ih_body (const Cta& c, const RoiRec& r, const Params& prm, const uint32_t* vals, uint32_t n, uint32_t* mem, double* out)
{
	const int N = prm.greydepth > 0 ? prm.greydepth : -prm.greydepth;
	const double mn = (double)r.imin, mx = (double)r.imax;
	if (!prm.ibsi || !(mx > mn) || N < 2 || n == 0)
	{
		for (int i = c.tid; i < NCOL_IH; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	uint32_t* freq = mem;
	for (int i = c.tid; i < N; i += c.nthr)
		freq[i] = 0;
	c.sync();
	const double binW = ddiv (mx - mn, (double)N);
	const double inv_binW = 1.0 / binW;
	for (uint32_t i = c.tid; i < n; i += c.nthr)
	{
		int idx = trunc_div_exact ((double)(uint32_t)(vals[i] - r.imin), binW, inv_binW);	// v >= min: floor == trunc
		if (idx >= N)
			idx = N - 1;
		atom_add (c, freq + idx, 1u);
	}
	c.sync();
	if (c.tid != 0)
		return;

	const double tot = (double)n;
	auto bin_of = [&](double value) -> int
	{
		int idx = (int)floor (ddiv (value - mn, binW));
		return idx < 0 ? 0 : (idx >= N ? N - 1 : idx);
	};
	auto centre = [&](int i) { return mn + ((double)i + 0.5) * binW; };
	// quantile interpolated within its bin
	auto quantile = [&](double p) -> double
	{
		double cum = 0.0, prev = 0.0, pn, f = 0.0;
		if (p < 0.5)
		{
			int k = 0;
			pn = 0.0;
			do
			{
				f = (double)freq[k];
				cum += f;
				prev = pn;
				pn = cum / tot;
				++k;
			} while (k < N && pn < p);
			const double lo = mn + (double)(k - 1) * binW, hi = mn + (double)k * binW;
			return lo + ((p - prev) / (f / tot)) * (hi - lo);
		}
		int k = N - 1, m = 0;
		pn = 1.0;
		do
		{
			f = (double)freq[k];
			cum += f;
			prev = pn;
			pn = 1.0 - cum / tot;
			--k;
			++m;
		} while (m < N && pn > p);
		const double lo = mn + (double)(k + 1) * binW, hi = mn + (double)(k + 2) * binW;
		return hi - ((prev - p) / (f / tot)) * (hi - lo);
	};
	// bin-centre median
	double med;
	{
		double run = 0.0;
		const double half = (double)(n / 2);
		int b = 0;
		while (run <= half && b < N)
		{
			run += (double)freq[b];
			++b;
		}
		med = centre (b - 1);
	}
	const int i_med = bin_of (med), i_min = bin_of (mn), i_max = bin_of (mx);
	const double p10 = quantile (0.10), p25 = quantile (0.25), p75 = quantile (0.75), p90 = quantile (0.90);
	const int i10 = bin_of (p10), i25 = bin_of (p25), i75 = bin_of (p75), i90 = bin_of (p90);

	// ---- pass 1: means, robust means over the bins [i10, i90]
	double mean_v = 0, mean_i = 0, rmean_v = 0, rmean_i = 0, rcnt = 0;
	for (int i = 0; i < N; i++)
	{
		const double f = (double)freq[i], pr = f / tot, v = centre (i);
		mean_v += pr * v;
		mean_i += pr * (double)i;
		if (i >= i10 && i <= i90)
		{
			rmean_v += f * v;
			rmean_i += f * (double)i;
			rcnt += f;
		}
	}
	rmean_v /= rcnt;
	rmean_i /= rcnt;

	// ---- pass 2: central sums, mode, deviations, entropy, uniformity, gradients
	double var_v = 0, var_i = 0, sk_v = 0, sk_i = 0, ku_v = 0, ku_i = 0, mode_v = 0, mode_i = 0, mode_f = 0;
	double mad_v = 0, mad_i = 0, rmad_v = 0, rmad_i = 0, medad_v = 0, medad_i = 0, ent = 0, uni = 0;
	double gmax = 2.2250738585072014e-308, gmax_i = 0, gmin = 1.7976931348623157e308, gmin_i = 0;
	const double LN2 = log (2.0);
	for (int i = 0; i < N; i++)
	{
		const double f = (double)freq[i], pr = f / tot, v = centre (i);
		const double dv = v - mean_v, di = (double)i - mean_i;
		var_v += pr * dv * dv;
		var_i += pr * di * di;
		sk_v += pr * dv * dv * dv;
		sk_i += pr * di * di * di;
		ku_v += pr * dv * dv * dv * dv;
		ku_i += pr * di * di * di * di;
		if (mode_f < f)
		{
			mode_f = f;
			mode_v = v;
			mode_i = (double)i;
		}
		mad_v += pr * fabs (dv);
		mad_i += pr * fabs (di);
		if (i >= i10 && i <= i90)
		{
			rmad_v += f * fabs (v - rmean_v);
			rmad_i += f * fabs ((double)i - rmean_i);
		}
		medad_v += pr * fabs (v - med);
		medad_i += pr * fabs ((double)i - (double)i_med);
		if (pr > 0.0000001)
			ent -= pr * log (pr) / LN2;
		uni += pr * pr;
		double g;
		if (i == 0)
			g = (double)freq[1] - (double)freq[0];
		else if (i == N - 1)
			g = (double)freq[i] - (double)freq[i - 1];
		else
			g = ((double)freq[i + 1] - (double)freq[i - 1]) / 2.0;
		if (g > gmax)
		{
			gmax = g;
			gmax_i = (double)(i + 1);
		}
		if (g < gmin)
		{
			gmin = g;
			gmin_i = (double)(i + 1);
		}
	}
	const double skew_v = sk_v / (var_v * sqrt (var_v)), skew_i = sk_i / (var_i * sqrt (var_i));
	const double kurt_v = ku_v / (var_v * var_v) - 3.0, kurt_i = ku_i / (var_i * var_i) - 3.0;
	rmad_v /= rcnt;
	rmad_i /= rcnt;

	// ---- the 46 columns in the order of the reference's feature list (== Feature2D code order)
	out[IH_MEAN_VAL] = mean_v;
	out[IH_VARIANCE_VAL] = var_v;
	out[IH_SKEWNESS_VAL] = skew_v;
	out[IH_EXCESS_KURTOSIS_VAL] = kurt_v;
	out[IH_MEDIAN_VAL] = med;
	out[IH_MINIMUM_VAL] = mn;
	out[IH_P10_VAL] = p10;
	out[IH_P90_VAL] = p90;
	out[IH_MAXIMUM_VAL] = mx;
	out[IH_MODE_VAL] = mode_v;
	out[IH_INTERQUANTILE_RANGE_VAL] = p75 - p25;
	out[IH_RANGE_VAL] = mx - mn;
	out[IH_MEAN_ABSOLUTE_DEVIATION_VAL] = mad_v;
	out[IH_ROBUST_MEAN_ABSOLUTE_DEVIATION_VAL] = rmad_v;
	out[IH_MEDIAN_ABSOLUTE_DEVIATION_VAL] = medad_v;
	out[IH_COEFFICIENT_OF_VARIATION_VAL] = sqrt (var_v) / mean_v;
	out[IH_QUANTILE_COEFFICIENT_OF_DISPERSION_VAL] = (p75 - p25) / (p75 + p25);
	out[IH_ENTROPY_VAL] = ent;
	out[IH_UNIFORMITY_VAL] = uni;
	out[IH_ROBUST_MEAN_VAL] = rmean_v;
	out[IH_MEAN_IDX] = mean_i + 1.0;
	out[IH_VARIANCE_IDX] = var_i;
	out[IH_SKEWNESS_IDX] = skew_i;
	out[IH_EXCESS_KURTOSIS_IDX] = kurt_i;
	out[IH_MEDIAN_IDX] = (double)i_med + 1.0;
	out[IH_MINIMUM_IDX] = (double)i_min + 1.0;
	out[IH_P10_IDX] = (double)i10 + 1.0;
	out[IH_P90_IDX] = (double)i90 + 1.0;
	out[IH_MAXIMUM_IDX] = (double)i_max + 1.0;
	out[IH_MODE_IDX] = mode_i + 1.0;
	out[IH_INTERQUANTILE_RANGE_IDX] = (double)i75 - (double)i25;
	out[IH_RANGE_IDX] = (double)i_max - (double)i_min;
	out[IH_MEAN_ABSOLUTE_DEVIATION_IDX] = mad_i;
	out[IH_ROBUST_MEAN_ABSOLUTE_DEVIATION_IDX] = rmad_i;
	out[IH_MEDIAN_ABSOLUTE_DEVIATION_IDX] = medad_i;
	out[IH_COEFFICIENT_OF_VARIATION_IDX] = sqrt (var_i) / (mean_i + 1.0);
	out[IH_QUANTILE_COEFFICIENT_OF_DISPERSION_IDX] = ((double)i75 - (double)i25) / ((double)i75 + 1.0 + (double)i25 + 1.0);
	out[IH_ENTROPY_IDX] = ent;
	out[IH_UNIFORMITY_IDX] = uni;
	out[IH_MAX_GRADIENT] = gmax;
	out[IH_MAX_GRADIENT_IDX] = gmax_i;
	out[IH_MIN_GRADIENT] = gmin;
	out[IH_MIN_GRADIENT_IDX] = gmin_i;
	out[IH_ROBUST_MEAN_IDX] = rmean_i + 1.0;
	out[IH_NUM_BINS] = (double)N;
	out[IH_BIN_SIZE] = binW;
}

} // namespace nyx
