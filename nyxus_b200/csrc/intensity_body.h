// nyxus_b200 -- per-ROI first-order intensity statistics (36 features of *ALL_INTENSITY*).
//
// Semantics follow the reference's PixelIntensityFeatures::calculate (features/intensity.cpp:57-192),
// TrivialHistogram (features/histogram.h:28-309) and Moments4 (features/moments.h:48-109):
//   ENERGY accumulates inten*inten in 32-bit unsigned arithmetic (intensity.cpp:90)
//   100-bin "percentile" histogram with NaN->0 index and bin 100 folded into 99 (histogram.h:55-66)
//   custom (grey-depth) histogram via to_grayscale, last bin folded (:68-77)
//   percentiles by linear interpolation inside the 100-bin histogram, later bins overwrite (:214-243)
//   exact median by full sort (:268-287), mode = smallest of the most frequent values (:289-309)
//   skewness / kurtosis from the central sums M2,M3,M4 (moments.h:83-109)
//
// B200 design: one CTA per ROI reads the ROI's contiguous slice of the CSR pixel cloud (coalesced), keeps it in
// shared memory, and derives everything from a few passes over that slice: exact integer sums for the raw
// moments, FP64 fixed-tree reductions for the central sums, integer shared-memory histograms, and an in-smem
// bitonic sort for median / mode.
#pragma once
#include "nyx_common.h"
#include "nyx_features.h"

namespace nyx
{

// shared-memory words: sorted buffer (pow2 >= n) + 101 + (nb+1) bins + 8 doubles
NYX_HD size_t intensity_smem_words (uint32_t npad, int nb)
{
	nb = nb < 0 ? -nb : nb;		// TrivialHistogram::initialize takes |n_cust_bins| (histogram.h:31)
	return (size_t)npad + 101 + (size_t)nb + 1 + 1 /*align*/ + 16;
}

// percentile interpolation exactly as TrivialHistogram::calc_percentiles (histogram.h:214-243)
NYX_HD double hist_percentile (const uint32_t* bins100, double cnt, double binW, uint32_t minv)
{
	double p = 0;
	unsigned long long run = 0;
	for (int i = 0; i < 100; i++)
	{
		uint32_t b = bins100[i];
		if ((double)run <= cnt && cnt <= (double)(run + b))
			p = dadd (dadd (ddiv (dmul (cnt - (double)run, binW), (double)b), (double)minv), dmul (binW, (double)i));
		run += b;
	}
	return p;
}

// vals: the ROI's n intensities (global memory).  mem: intensity_smem_words(npad, nb) words.  out: NCOL_INT doubles.
NYX_D void intensity_body (const Cta& c, const RoiRec& r, const Params& prm, const uint32_t* vals, uint32_t n,
	uint32_t npad, uint32_t* mem, double* out)
{
	const int nb = prm.greydepth > 0 ? prm.greydepth : -prm.greydepth;
	uint32_t* srt = mem;
	uint32_t* b100 = srt + npad;
	uint32_t* bc = b100 + 101;
	size_t off = (size_t)npad + 101 + nb + 1;
	off += off & 1;
	double* sh = (double*)(mem + off);	// [0..5] percentiles

	const uint32_t mn = r.imin, mx = r.imax, range = mx - mn;
	const double nd = (double)n;

	for (int i = c.tid; i < NCOL_INT; i += c.nthr)
		out[i] = 0.0;
	for (uint32_t i = c.tid; i < 101 + (uint32_t)nb + 1; i += c.nthr)
		b100[i] = 0;

	// ---- pass 1: load + exact raw sums (intensity.cpp:86-98)
	unsigned long long s12[2] = { 0, 0 };
	for (uint32_t i = c.tid; i < npad; i += c.nthr)
	{
		uint32_t v = 0xFFFFFFFFu;
		if (i < n)
		{
			v = vals[i];
			s12[0] += v;
			s12[1] += (uint32_t)(v * v);
		}
		srt[i] = v;
	}
	c.sync();
	cta_sum_u64<2> (c, s12);
	const double mean = (double)s12[0] / nd;
	const double energy = (double)s12[1];

	// ---- pass 2: central sums + histograms (intensity.cpp:101-110,169-182; histogram.h:50-77)
	const double binW100 = ddiv ((double)range, 100.0);
	const double inv_binW100 = range ? 1.0 / binW100 : 0.0, inv_range = range ? 1.0 / (double)range : 0.0;
	const bool fast_gs = range != 0 && nb < 65536;
	enum { C_MAD, C_M2, C_M3, C_M4, C_P5, C_P6, C_N };
	double cs[C_N] = { 0, 0, 0, 0, 0, 0 };
	const bool blank = (mn == 0 && mx == 0);
	for (uint32_t i = c.tid; i < n; i += c.nthr)
	{
		uint32_t v = srt[i];
		double d = (double)v - mean, d2 = d * d;
		cs[C_MAD] += fabs (d);
		cs[C_M2] += d2;
		cs[C_M3] += d2 * d;
		cs[C_M4] += d2 * d2;
		cs[C_P5] += d2 * d2 * d;
		cs[C_P6] += d2 * d2 * d2;
		// bin of the 100-bin histogram: (int)((v - min) / binW), NaN (range 0: 0 / 0) -> 0 (histogram.h:55-66)
		const int idx = range ? trunc_div_exact ((double)(uint32_t)(v - mn), binW100, inv_binW100) : 0;
		atom_add (c, b100 + idx, 1u);
		// custom-bin histogram (histogram.h:68-77): to_grayscale, NaN (range 0) -> 0
		atom_add (c, bc + (fast_gs ? to_grayscale_fast (v - mn, range, (uint32_t)nb, inv_range) : to_grayscale (v, mn, range, (uint32_t)nb, false)), 1u);
	}
	c.sync();
	cta_sum<C_N> (c, cs);

	const double var = cs[C_M2];
	const double variance = n > 1 ? var / (nd - 1.0) : 0.0;
	const double variance_b = n > 1 ? var / nd : 0.0;
	const double sd = sqrt (variance), sd_b = sqrt (variance_b);

	if (c.tid == 0)
	{
		out[I_MIN] = (double)mn;
		out[I_MAX] = (double)mx;
		out[I_RANGE] = (double)mx - (double)mn;
		// in-memory featurize(): slide extrema are the sentinels lowest()/max() (slideprops.cpp:29-30,74-75)
		// -> range / (lowest - max) = range / -inf = -0.0
		// file-based volumes: the scanned slide range over the labelled voxels (slideprops.cpp:305-327, 3d_intensity.cpp:56-61)
		out[I_COVERED_IMAGE_INTENSITY_RANGE] = prm.covered_den != 0.0 ? (double)(uint32_t)(mx - mn) / prm.covered_den : -0.0;
		out[I_MEAN] = mean;
		out[I_ENERGY] = energy;
		out[I_ROOT_MEAN_SQUARED] = sqrt (energy / nd);
		out[I_INTEGRATED_INTENSITY] = (double)s12[0];
		out[I_MEAN_ABSOLUTE_DEVIATION] = cs[C_MAD] / nd;
		out[I_VARIANCE] = variance;
		out[I_VARIANCE_BIASED] = variance_b;
		out[I_STANDARD_DEVIATION] = sd;
		out[I_STANDARD_DEVIATION_BIASED] = sd_b;
		out[I_COV] = sd / mean;
		out[I_STANDARD_ERROR] = sd / sqrt (nd);
	}
	// HISTOGRAM (intensity.cpp:129, histogram.h:190-198; blank ROIs: padded with zeros, output_2_buffer.cpp:283-293): `out` is the
	// ROI's first-order block, the per-bin block starts hist_col - COL_INT columns further
	double* hist = prm.hist_col ? out + (prm.hist_col - COL_INT) : nullptr;
	if (blank)	// intensity.cpp:121-122
	{
		if (hist)
			for (int k = c.tid; k < nb; k += c.nthr)
				hist[k] = 0.0;
		return;
	}

	// ---- fold the special last bins, percentiles (one thread per percentile)
	if (c.tid == 0)
	{
		b100[99] += b100[100];
		b100[100] = 0;
		bc[nb - 1] += bc[nb];
		bc[nb] = 0;
	}
	c.sync();
	{
		const double q[6] = { 0.01, 0.1, 0.25, 0.75, 0.9, 0.99 };
#if NYX_DEVICE_BUILD
		if (c.wmode)
		{
			// warp form of hist_percentile: 4 bins per lane, exclusive prefix of the bin counts by a shuffle scan, then per
			// percentile the LAST bin whose [run, run + b] holds the target count evaluates the reference's formula
			const int lane = c.tid;
			uint32_t b4[4], run4[4], tot = 0;
#pragma unroll
			for (int k = 0; k < 4; k++)
			{
				b4[k] = lane * 4 + k < 100 ? b100[lane * 4 + k] : 0u;
				run4[k] = tot;
				tot += b4[k];
			}
			uint32_t inc = tot;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				const uint32_t t = __shfl_up_sync (0xffffffffu, inc, o);
				if (lane >= o)
					inc += t;
			}
			const uint32_t excl = inc - tot;
#pragma unroll
			for (int qi = 0; qi < 6; qi++)
			{
				const double cnt = dmul (nd, q[qi]);
				int last = 0;	// 1 + index of my last matching bin
#pragma unroll
				for (int k = 0; k < 4; k++)
				{
					const double r0 = (double)(excl + run4[k]), r1 = (double)((unsigned long long)excl + run4[k] + b4[k]);
					if (lane * 4 + k < 100 && r0 <= cnt && cnt <= r1)
						last = lane * 4 + k + 1;
				}
				const int glast = (int)__reduce_max_sync (0xffffffffu, (unsigned)last);
				if (glast == 0)
				{
					if (lane == 0)
						sh[qi] = 0.0;
				}
				else if (last == glast)
				{
					const int i = glast - 1, k = i & 3;
					const uint32_t run = excl + (k == 0 ? run4[0] : k == 1 ? run4[1] : k == 2 ? run4[2] : run4[3]);
					const uint32_t b = k == 0 ? b4[0] : k == 1 ? b4[1] : k == 2 ? b4[2] : b4[3];
					sh[qi] = dadd (dadd (ddiv (dmul (cnt - (double)run, binW100), (double)b), (double)mn), dmul (binW100, (double)i));
				}
			}
		}
		else
#endif
		for (int k = c.tid; k < 6; k += c.nthr)
			sh[k] = hist_percentile (b100, dmul (nd, q[k]), binW100, mn);
	}

	// ---- entropy / uniformity over the custom bins (histogram.h:145-151)
	double eu[2] = { 0, 0 };
#if NYX_DEVICE_BUILD
	if (c.wmode && n < LN_TABLE_N)
	{
		// log2 (cnt / n + 2.2e-16) = (ln cnt - ln n) / ln 2 up to 2.2e-16 * n / cnt <= 3e-13 (n < 1024): table loads
		// instead of a double-precision logarithm per bin; empty bins contribute 0 * log2 (2.2e-16) = 0
		const double ln_n = g_ln_table[n], inv_n = 1.0 / nd;
		for (int k = c.tid; k <= nb; k += c.nthr)
		{
			const uint32_t cnt = bc[k];
			if (cnt)
			{
				const double p = (double)cnt * inv_n;
				eu[0] += p * ((g_ln_table[cnt] - ln_n) * INV_LN2);
				eu[1] += p * p;
			}
		}
	}
	else
#endif
	for (int k = c.tid; k <= nb; k += c.nthr)
	{
		double p = (double)bc[k] / nd;
		eu[0] += p * log2 (p + 2.2e-16);
		eu[1] += p * p;
	}
	c.sync();
	cta_sum<2> (c, eu);
	const double p01 = sh[0], p10 = sh[1], p25 = sh[2], p75 = sh[3], p90 = sh[4], p99 = sh[5];
	if (hist)	// the folded grey-depth histogram itself: raw counts of bins 0 .. nb-1
		for (int k = c.tid; k < nb; k += c.nthr)
			hist[k] = (double)bc[k];

	// ---- sort -> median (histogram.h:268-287), mode (:289-309)
	sort_keys (c, srt, npad);
	double median;
	if (n & 1)
		median = (double)srt[n / 2];
	else
		median = (double)(uint32_t)(srt[n / 2] + srt[n / 2 - 1]) / 2.0;
	unsigned long long best = 0;	// (run length << 32) | ~value : longest run wins, ties -> smallest value
#if NYX_DEVICE_BUILD
	if (c.wmode)
	{
		// chunks of 32 consecutive sorted keys, last chunk first: a run head's length is the distance to the next head,
		// found in the chunk's head mask or, when the run leaves the chunk, in the carried index of the next head
		const int lane = c.tid;
		uint32_t next_head = n;		// index of the first run head after the current chunk
		for (int base = (int)((n - 1) & ~31u); base >= 0; base -= 32)
		{
			const uint32_t i = (uint32_t)base + lane;
			const bool in = i < n;
			const uint32_t v = in ? srt[i] : 0u;
			const bool head = in && (i == 0 || srt[i - 1] != v);
			const unsigned heads = __ballot_sync (0xffffffffu, head);
			if (head)
			{
				const unsigned later = lane < 31 ? heads >> (lane + 1) : 0u;
				const uint32_t nxt = later ? i + (uint32_t)__ffs (later) : next_head;
				const unsigned long long key = ((unsigned long long)(nxt - i) << 32) | (unsigned long long)(~v);
				best = key > best ? key : best;
			}
			if (heads)
				next_head = (uint32_t)base + (uint32_t)__ffs (heads) - 1u;
		}
	}
	else
#endif
	for (uint32_t i = c.tid; i < n; i += c.nthr)
	{
		uint32_t v = srt[i];
		if (i == 0 || srt[i - 1] != v)
		{
			uint32_t j = i + 1;
			while (j < n && srt[j] == v) j++;
			unsigned long long key = ((unsigned long long)(j - i) << 32) | (unsigned long long)(~v);
			best = key > best ? key : best;
		}
	}
	best = cta_max_u64 (c, best);
	const double mode = (double)(uint32_t)(~(uint32_t)(best & 0xFFFFFFFFull));

	// ---- pass 3: robust mean (intensity.cpp:139-149, histogram.h:89-100), median abs deviation (:156-159)
	unsigned long long rb[2] = { 0, 0 };
	double md[1] = { 0 };
	for (uint32_t i = c.tid; i < n; i += c.nthr)
	{
		double a = (double)srt[i];
		if (a >= p10 && a <= p90)
		{
			rb[0] += srt[i];
			rb[1] += 1;
		}
		md[0] += fabs (a - median);
	}
	cta_sum_u64<2> (c, rb);
	cta_sum<1> (c, md);
	const double robust_mean = rb[1] ? (double)rb[0] / (double)rb[1] : 0.0;

	// ---- pass 4: robust mean absolute deviation (histogram.h:101-112)
	double rm[1] = { 0 };
	if (rb[1])
		for (uint32_t i = c.tid; i < n; i += c.nthr)
		{
			double a = (double)srt[i];
			if (a >= p10 && a <= p90)
				rm[0] += fabs (a - robust_mean);
		}
	cta_sum<1> (c, rm);
	const double rmad = rb[1] ? rm[0] / (double)rb[1] : 0.0;

	if (c.tid == 0)
	{
		out[I_MEDIAN] = median;
		out[I_MODE] = mode;
		out[I_P01] = p01;
		out[I_P10] = p10;
		out[I_P25] = p25;
		out[I_P75] = p75;
		out[I_P90] = p90;
		out[I_P99] = p99;
		out[I_QCOD] = (p75 - p25) / (p75 + p25);
		out[I_INTERQUARTILE_RANGE] = p75 - p25;
		out[I_ROBUST_MEAN] = robust_mean;
		out[I_ROBUST_MEAN_ABSOLUTE_DEVIATION] = rmad;
		out[I_ENTROPY] = -eu[0];
		out[I_UNIFORMITY] = eu[1];
		out[I_MEDIAN_ABSOLUTE_DEVIATION] = md[0] / nd;
		out[I_UNIFORMITY_PIU] = (1.0 - (double)(uint32_t)(mx - mn) / (double)(uint32_t)(mx + mn)) * 100.0;	// :162
		// Moments4 (moments.h:83-109)
		const double M2 = cs[C_M2], M3 = cs[C_M3], M4 = cs[C_M4];
		double skew = 0, kurt = 0, ekurt = 0;
		if (M2 != 0.0)
		{
			skew = n > 3 ? (sqrt (nd) * M3) / (M2 * sqrt (M2)) : 0.0;	// pow (M2, 1.5) within 2 ulp, without the generic pow()
			kurt = n > 4 ? (nd * M4) / (M2 * M2) : 0.0;
			ekurt = n > 4 ? (nd * M4) / (M2 * M2) - 3.0 : 0.0;
		}
		out[I_SKEWNESS] = skew;
		out[I_KURTOSIS] = kurt;
		out[I_EXCESS_KURTOSIS] = ekurt;
		const double sd2 = sd * sd;
		double den5 = nd * (sd2 * sd2 * sd), den6 = nd * (sd2 * sd2 * sd2);	// nd * pow (sd, 5 | 6) (intensity.cpp:185-191) within 3 ulp
		out[I_HYPERSKEWNESS] = den5 == 0.0 ? 0.0 : cs[C_P5] / den5;
		out[I_HYPERFLATNESS] = den6 == 0.0 ? 0.0 : cs[C_P6] / den6;
	}
}

} // namespace nyx
