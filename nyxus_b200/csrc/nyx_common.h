// nyxus_b200 -- shared scalar numerics and the per-CTA execution context used by every per-ROI feature body.
//
// The feature bodies (glcm_body.h, glrlm_body.h, ...) are written against the small `Cta` interface below:
// strided loops over work items (`for (i = c.tid; i < n; i += c.nthr)`), `c.sync()`, block reductions and
// integer shared-memory atomics.  On the device a CTA of 128-256 threads runs them; tests/emu compiles the very
// same bodies for the host with nthr == 1 so their indexing/arithmetic can be checked against the oracle on a
// machine without a GPU (test harness only; the product library contains only the CUDA instantiation).
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__) && !defined(NYX_HOST_EMU)
#define NYX_DEVICE_BUILD 1
#define NYX_HD __host__ __device__ __forceinline__
#define NYX_D __device__ __forceinline__
#else
#define NYX_DEVICE_BUILD 0
#define NYX_HD inline
#define NYX_D inline
#endif

namespace nyx
{

// ------------------------------------------------------------------------------------------------
// Exactly-rounded, never-contracted FP ops.  The reference is built for generic x86-64 at -O2 (no FMA),
// so every bit-exact path (binning, fast_log10, histogram bin indices) must avoid FMA contraction.
// ------------------------------------------------------------------------------------------------
NYX_HD double dmul (double a, double b)
{
#ifdef __CUDA_ARCH__
	return __dmul_rn (a, b);
#else
	return a * b;
#endif
}
NYX_HD double dadd (double a, double b)
{
#ifdef __CUDA_ARCH__
	return __dadd_rn (a, b);
#else
	return a + b;
#endif
}
NYX_HD double ddiv (double a, double b)
{
#ifdef __CUDA_ARCH__
	return __ddiv_rn (a, b);
#else
	return a / b;
#endif
}
NYX_HD float fmul (float a, float b)
{
#ifdef __CUDA_ARCH__
	return __fmul_rn (a, b);
#else
	return a * b;
#endif
}
NYX_HD float fadd (float a, float b)
{
#ifdef __CUDA_ARCH__
	return __fadd_rn (a, b);
#else
	return a + b;
#endif
}
NYX_HD uint32_t f2u (float f)
{
#ifdef __CUDA_ARCH__
	return __float_as_uint (f);
#else
	uint32_t u;
	memcpy (&u, &f, 4);
	return u;
#endif
}
NYX_HD float u2f (uint32_t u)
{
#ifdef __CUDA_ARCH__
	return __uint_as_float (u);
#else
	float f;
	memcpy (&f, &u, 4);
	return f;
#endif
}

constexpr double LOG10_2 = 0.30102999566;	// reference glcm.h:241, glrlm.h:171, glszm.h:140, gldm.h:106
constexpr double GLCM_EPSILON = 0.000000001;	// reference glcm.h:249
constexpr double EPS_ZONES = 2.2e-16;		// reference glrlm.h:169, glszm.h:138, gldm.h:105, gldzm.h:68

// The reference's fast_log10() (helpers/helpers.h:283-330): a float32 quadratic approximation of log2(x)
// scaled by log10(2).  Restated: split exponent / mantissa, reduce the significand to [0.75,1.5),
// evaluate  fexp + a*s*s + b*s  left to right in float32, widen, scale.  Must be bit-identical.
NYX_HD double fast_log10 (double xd)
{
	const float a = -.6296735f, b = 1.466967f;
	float x = (float)xd;
	uint32_t bits = f2u (x);
	int e = (int)((bits & 0x7F800000u) >> 23);
	bool upper = (bits & 0x00400000u) != 0;		// significand >= 1.5 -> halve it
	float s = u2f ((bits & 0x007FFFFFu) | (upper ? 0x3f000000u : 0x3f800000u));
	float fexp = (float)(upper ? e - 126 : e - 127);
	s = fadd (s, -1.0f);
	float lg2 = fadd (fadd (fexp, fmul (fmul (a, s), s)), fmul (b, s));
	return dmul ((double)lg2, 0.30102999566);
}

// x * fast_log10(x + eps) / LOG10_2  -- the entropy term every family uses
NYX_HD double plogp (double p, double eps)
{
	return ddiv (dmul (p, fast_log10 (dadd (p, eps))), LOG10_2);
}

// ln(k) of a small integer count.  Entropies over integer matrices, sum_cells P log(P/N), become
// sum_elements ln(count) - n ln(N): the per-element log is a table load (the table is filled by the host when the
// context is created).  |error| ~ 1e-16 relative to the reference's log(count/N + 2.2e-16).
constexpr uint32_t LN_TABLE_N = 1024;
#if NYX_DEVICE_BUILD
__device__ double g_ln_table[LN_TABLE_N];
#endif
NYX_D double ln_count (uint32_t k)
{
#if NYX_DEVICE_BUILD
	return k < LN_TABLE_N ? g_ln_table[k] : log ((double)k);
#else
	return log ((double)k);
#endif
}
constexpr double INV_LN2 = 1.4426950408889634074;

// weights that depend on a small integer k only: 1/(1+k^2), 1/(1+k), 1/k^2 (GLCM difference-marginal features), from a
// table filled by the host at context creation instead of three FP64 divisions per marginal entry
constexpr uint32_t KW_TABLE_N = 1024;
struct KWeights { double inv1pk2, inv1pk, invk2; };
#if NYX_DEVICE_BUILD
__device__ KWeights g_kw_table[KW_TABLE_N];
#endif
NYX_D KWeights k_weights (uint32_t k)
{
#if NYX_DEVICE_BUILD
	if (k < KW_TABLE_N)
		return g_kw_table[k];
#endif
	const double kd = (double)k;
	KWeights w;
	w.inv1pk2 = 1.0 / (1.0 + kd * kd);
	w.inv1pk = 1.0 / (1.0 + kd);
	w.invk2 = k ? 1.0 / (kd * kd) : 0.0;
	return w;
}

// 1 / k^2 of a positive integer (table for small k: no division)
NYX_D double inv_sq_int (uint32_t k)
{
#if NYX_DEVICE_BUILD
	if (k - 1u < KW_TABLE_N - 1u)
		return g_kw_table[k].invk2;
#endif
	return 1.0 / ((double)k * (double)k);
}

// The six "emphasis" sums every run / zone / dependence family takes over its elements (level i, minor j):
//   s[0] += 1/j^2   s[1] += 1/i^2   s[2] += 1/(i^2 j^2)   s[3] += i^2/j^2   s[4] += j^2/i^2   s[5] += i^2 j^2
// (GLRLM glrlm.cpp:381-383,712-719,781,813,844,876; GLSZM glszm.cpp:372-375,424-427,559-563; GLDZM gldzm.cpp:402,421,
// 440-443; GLDM gldm.cpp:301-302,452-550).  Two reciprocals instead of the reference's five-to-thirteen divisions per
// element: each term differs from the reference's by at most 2 ulp (features are held to 1e-9, not bit-exact).
NYX_HD void emphasis6 (double* s, double i2, double j2)
{
	const double ri = 1.0 / i2, rj = 1.0 / j2;
	s[0] += rj;
	s[1] += ri;
	s[2] += ri * rj;
	s[3] += i2 * rj;
	s[4] += j2 * ri;
	s[5] += i2 * j2;
}

// the same for small integer (i, j): 1/i^2 and 1/j^2 come from the table (no division at all)
NYX_D void emphasis6_int (double* s, uint32_t i, uint32_t j)
{
#if NYX_DEVICE_BUILD
	if (i - 1u < KW_TABLE_N - 1u && j - 1u < KW_TABLE_N - 1u)	// 1 <= i, j < KW_TABLE_N (0 keeps the division's inf)
	{
		const double ri = g_kw_table[i].invk2, rj = g_kw_table[j].invk2;
		const double i2 = (double)(i * i), j2 = (double)(j * j);
		s[0] += rj;
		s[1] += ri;
		s[2] += ri * rj;
		s[3] += i2 * rj;
		s[4] += j2 * ri;
		s[5] += i2 * j2;
		return;
	}
#endif
	emphasis6 (s, (double)i * (double)i, (double)j * (double)j);
}


// ------------------------------------------------------------------------------------------------
// Grey-level binning (reference features/texture_feature.h:23-50,138-167; helpers/helpers.h:337-345)
// ------------------------------------------------------------------------------------------------
NYX_HD uint32_t to_grayscale (uint32_t i, uint32_t mn, uint32_t range, uint32_t n, bool disable);
struct Binner
{
	int greyinfo;		// >0 MATLAB levels, 0 IBSI identity, <0 PyRadiomics bin count
	double slope;		// MATLAB: g / max;  PyRadiomics: the bin width (max - min) / |g|
	uint32_t n;		// |g|
	uint32_t mn;		// PyRadiomics: aux_min
	uint32_t tg;		// 0, or 1 + disable: Nyxus::to_grayscale (v, 0, range = mn, n, disable) instead (3D NGLDM)

	// the NGLDM grey scale (3d_ngldm.cpp:88-103): to_grayscale over [0, aux_max]
	NYX_HD void init_tg (int ngreys, uint32_t imax, bool disable)
	{
		greyinfo = 0;
		slope = 0.0;
		n = (uint32_t)ngreys;
		mn = imax;
		tg = disable ? 2u : 1u;
	}
	NYX_HD void init (int greyinfo_, uint32_t imin, uint32_t imax)
	{
		tg = 0;
		greyinfo = greyinfo_;
		n = (uint32_t)(greyinfo_ > 0 ? greyinfo_ : -greyinfo_);
		mn = imin;
		if (greyinfo_ > 0)
			slope = ddiv ((double)greyinfo_, (double)imax);
		else if (greyinfo_ < 0)
			slope = ddiv ((double)(uint32_t)(imax - imin), (double)(int)n);	// texture_feature.h:110
		else
			slope = 0.0;
	}
	NYX_HD uint32_t bin (uint32_t i) const
	{
		if (tg)
			return to_grayscale (i, 0u, mn, n, tg == 2u);
		if (greyinfo == 0)
			return i;
		if (greyinfo < 0)
		{
			// to_grayscale_radiomix (texture_feature.h:106-118): 0 stays 0; 1-based bins, the last one a unit wider.
			// (PixIntens)(double): x86-64 gcc converts through a 64-bit cvttsd2si and keeps the low 32 bits, so NaN
			// (min == max: 0/0) and anything >= 2^63 give 0
			if (i == 0)
				return 0;
			const double q = dadd (ddiv ((double)(uint32_t)(i - mn), slope), 1.0);
			uint32_t y = 0;
			if (q == q && q > -9223372036854775808.0 && q < 9223372036854775808.0)
				y = (uint32_t)(unsigned long long)(long long)q;
			if (y > n)
				y = n;
			return y;
		}
		if (i == 0)
			return 1;
		double r = floor (dadd (dmul (slope, (double)i), 1.0));
		uint32_t s = (uint32_t)r;
		if (s > n) s = n;
		if (s < 1) s = 1;
		return s;
	}
};

// Nyxus::to_grayscale (helpers.h:337-345): (uint)(double(i-min)/double(range)*double(n)); NaN (range==0) -> 0 like x86-64 gcc
NYX_HD uint32_t to_grayscale (uint32_t i, uint32_t mn, uint32_t range, uint32_t n, bool disable)
{
	if (disable)
		return i;
	double pi = dmul (ddiv ((double)(uint32_t)(i - mn), (double)range), (double)n);
	if (!(pi == pi))
		return 0;
	if (pi >= 4294967296.0)
		return 0;
	return (uint32_t)pi;
}

// (int)(x / w) for x >= 0, w > 0 exactly as the reference's double division would give it, mostly without dividing:
// q = x * (1/w) is within a few ulp of x / w, so both truncate alike unless q sits next to an integer; only then the
// correctly rounded division decides.  inv_w = 1 / w (rounded).
NYX_HD int trunc_div_exact (double x, double w, double inv_w)
{
	if (x == 0.0)
		return 0;
	const double q = x * inv_w;
	const int k = (int)q;
	const double fr = q - (double)k;
	if (fr > 1e-7 && fr < 1.0 - 1e-7)
		return k;
	return (int)ddiv (x, w);
}

// Nyxus::to_grayscale (helpers.h:337-345) = (uint)(double(x) / double(range) * double(n)) with x = i - min <= range:
// the two roundings move the product by < 2^-35 (n < 2^16), so it truncates like the exact x * n / range unless that is
// (next to) an integer -- only then the reference's own operation sequence is evaluated.  inv_range = 1 / range.
NYX_HD uint32_t to_grayscale_fast (uint32_t x, uint32_t range, uint32_t n, double inv_range)
{
	if (x == 0)	// 0 / range * n == 0 exactly (every cell outside the ROI takes this exit)
		return 0;
	const double q = (double)x * (double)n * inv_range;
	const uint32_t k = (uint32_t)q;
	const double fr = q - (double)k;
	if (fr > 1e-7 && fr < 1.0 - 1e-7)
		return k;
	return to_grayscale (x, 0u, range, n, false);
}

NYX_HD double finite_or (double v, double sub)
{
	return (isnan (v) || isinf (v)) ? sub : v;
}

// ------------------------------------------------------------------------------------------------
// Device-side records
// ------------------------------------------------------------------------------------------------
struct RoiRec
{
	uint32_t label;
	uint32_t area;
	uint32_t imin, imax;
	int32_t x0, y0, w, h;	// AABB origin and extent (w = xmax-xmin+1)
	uint64_t cloud_off;	// offset of this ROI's pixels in the CSR cloud
	int32_t z0, d;		// 3D only
	uint64_t tile_off;	// first cell of this ROI's slot in the binned-tile store (multiple of 8)
};

struct ImageView
{
	const void* inten;	// uint32 or uint16 pixels, row-major
	const uint32_t* label;
	int32_t H, W;
	int32_t inten_u16;
	int32_t D = 1;		// slices (1: 2D image); voxel (z,y,x) lives at (z*H + y)*W + x
	// binned-tile store written by k_pack_tiles (null: the family kernels bin straight from the slide):
	// tileA = Binner levels + NZ/IN flags, tileB = NGLDM's to_grayscale levels + IN flag; 16-bit codes, ROI slot at RoiRec::tile_off
	const uint16_t* tileA = nullptr;
	const uint16_t* tileB = nullptr;
};

struct Params
{
	int32_t greydepth;	// NyxSetting::GREYDEPTH = coarse_gray_depth (> 0 MATLAB levels; < 0 PyRadiomics bin count, C ABI only)
	int32_t glcm_greydepth;	// NyxSetting::GLCM_GREYDEPTH: 2D = the GLCM blank-ROI test only (glcm.cpp:22-28); 3D = the GLCM binning (3d_glcm.cpp:53)
	int32_t ibsi;
	int32_t glcm_offset;
	int32_t glcm_symmetric;
	double softnan;
	// 3D (Nyxus3D) settings: per-family grey depths (env_features.cpp:712-736: only GLCM gets a default, the rest are 0
	// = no binning unless set through set_metaparam) and the NGTDM neighbourhood radius
	int32_t glszm3_greydepth = 0, ngtdm3_greydepth = 0, ngtdm3_radius = 0;
	int32_t cube_maxdim = 0;	// 3D launches: the longest cube edge of the launch's ROI class (run-length / distance bounds)
	double covered_den = 0;	// first-order COVERED_IMAGE_INTENSITY_RANGE denominator (slide max - min); 0 = in-memory sentinel
	int32_t hist_col = 0;	// canonical column of HISTOGRAM_BIN_0 when the opt-in HISTOGRAM feature is selected (grey depth > 0), else 0
};

NYX_HD uint32_t load_inten (const ImageView& im, size_t idx)
{
	return im.inten_u16 ? (uint32_t)((const uint16_t*)im.inten)[idx] : ((const uint32_t*)im.inten)[idx];
}

// tile cell code: level in the low 24 bits + flags
constexpr uint32_t CELL_NZ = 0x80000000u;	// original intensity != 0 (and pixel belongs to the ROI)
constexpr uint32_t CELL_IN = 0x40000000u;	// pixel belongs to the ROI (may have intensity 0)
constexpr uint32_t CELL_LEVEL = 0x00FFFFFFu;

// ------------------------------------------------------------------------------------------------
// CTA context
// ------------------------------------------------------------------------------------------------
struct Cta
{
	int tid, nthr;
	double* red;	// reduction scratch: RED_MAX * NW_MAX doubles (unused by a warp-wide context)
	int wmode = 0;	// 1: this context is ONE full warp (nthr == 32): sync = __syncwarp, reductions by shuffles only

#if NYX_DEVICE_BUILD
	NYX_D void sync() const
	{
		if (wmode)
			__syncwarp();
		else
			__syncthreads();
	}
	// sub-groups: the warps of this context (a warp-wide context is its own single sub-group)
	NYX_D int sub_id() const { return wmode ? 0 : (tid >> 5); }
	NYX_D int sub_count() const { return wmode ? 1 : ((nthr + 31) >> 5); }
	NYX_D Cta sub() const { return Cta { tid & 31, 32, red, 1 }; }
	NYX_D unsigned ballot (bool p) const { return __ballot_sync (0xffffffffu, p); }	// sub-group (warp) contexts only
	NYX_D unsigned lanemask_lt() const { return (1u << (tid & 31)) - 1u; }
#else
	inline void sync() const {}
	inline int sub_id() const { return 0; }
	inline int sub_count() const { return 1; }
	inline Cta sub() const { return *this; }
	inline unsigned ballot (bool p) const { return p ? 1u : 0u; }
	inline unsigned lanemask_lt() const { return 0u; }
#endif
};

NYX_HD int popc32 (unsigned v)
{
#ifdef __CUDA_ARCH__
	return __popc (v);
#else
	return __builtin_popcount (v);
#endif
}

// exact n / d for n * d < 2^32 by one multiply-high (tile indices: n < cells of a tile, d = tile width)
struct FastDiv
{
	uint32_t d, m;
	NYX_HD void init (uint32_t d_)
	{
		d = d_;
		m = d_ > 1 ? (uint32_t)((0x100000000ull + d_ - 1) / d_) : 0;
	}
	NYX_HD uint32_t div (uint32_t n) const
	{
#ifdef __CUDA_ARCH__
		return d > 1 ? __umulhi (n, m) : n;
#else
		return n / d;
#endif
	}
};

constexpr int RED_MAX = 48;
constexpr int NW_MAX = 8;	// CTAs have at most 256 threads
constexpr int RED_WORDS = RED_MAX * NW_MAX * 2;	// 32-bit words of reduction scratch
// the pipelined GLCM / GLRLM loop (k_roi_family): words [0, PF_STAGE_OFF) of the scratch hold the staged angle values
// (4 x 30 doubles), words [PF_STAGE_OFF, PF_STAGE_OFF + 512) the 16-bit codes of the next tile (<= 1024 cells);
// PF_RING_WORDS after the family region: 4 list entries + 2 ROI records in flight
constexpr int PF_STAGE_OFF = 240, PF_RING_WORDS = 32;
static_assert (PF_STAGE_OFF + 512 <= RED_WORDS, "tile stage does not fit the reduction scratch");

NYX_D void atom_add (const Cta&, uint32_t* p, uint32_t v)
{
#if NYX_DEVICE_BUILD
	atomicAdd (p, v);
#else
	*p += v;
#endif
}
NYX_D void atom_max (const Cta&, uint32_t* p, uint32_t v)
{
#if NYX_DEVICE_BUILD
	atomicMax (p, v);
#else
	if (v > *p) *p = v;
#endif
}
NYX_D void atom_min (const Cta&, uint32_t* p, uint32_t v)
{
#if NYX_DEVICE_BUILD
	atomicMin (p, v);
#else
	if (v < *p) *p = v;
#endif
}
NYX_D void atom_add64 (const Cta&, unsigned long long* p, unsigned long long v)
{
#if NYX_DEVICE_BUILD
	atomicAdd (p, v);
#else
	*p += v;
#endif
}

// zero n 32-bit words (128-bit stores when the range is 16-byte aligned): small dense matrices are cheaper to wipe
// than to clear cell by cell through a re-enumeration of their owners
template <class C> NYX_D void wipe_words (const C& c, uint32_t* p, size_t n)
{
#if NYX_DEVICE_BUILD
	if (((size_t)p & 15) == 0)
	{
		uint4* p4 = reinterpret_cast<uint4*> (p);
		const int n4 = (int)(n >> 2);
		for (int i = c.tid; i < n4; i += c.nthr)
			p4[i] = make_uint4 (0u, 0u, 0u, 0u);
		for (int i = (n4 << 2) + c.tid; i < (int)n; i += c.nthr)
			p[i] = 0;
		return;
	}
#endif
	for (size_t i = c.tid; i < n; i += c.nthr)
		p[i] = 0;
}

#if NYX_DEVICE_BUILD
// Warp-wide sums of N doubles, every lane gets all N totals.  N independent 5-step butterflies cost 10 N shuffles and 5 N
// additions; here the first log2(M) steps (M = N rounded up to a power of two) are a REDUCE-SCATTER -- at the step with
// lane distance o a lane keeps the half of its values selected by its bit o and sends the other half -- so the work halves
// from step to step (M - 1 exchanges in all), the remaining steps are a butterfly of the ONE value a lane is left with
// (value k ends up, complete, in the lanes whose top bits spell k), and N broadcasts hand the totals round:
// 2 (M - 1) + 2 (5 - log2 M) + 2 N shuffles and M + 4 - log2 M additions (N = 15: 66 instead of 150, 16 instead of 75).
// A fixed tree: deterministic.
template <int N> NYX_D void warp_allsum (double (&v)[N], int lane)
{
	constexpr int M = N <= 2 ? 2 : N <= 4 ? 4 : N <= 8 ? 8 : N <= 16 ? 16 : 32;
	constexpr int L = M == 2 ? 1 : M == 4 ? 2 : M == 8 ? 3 : M == 16 ? 4 : 5;
	double w[M];
#pragma unroll
	for (int i = 0; i < M; i++)
		w[i] = i < N ? v[i] : 0.0;
#pragma unroll
	for (int st = 0; st < L; st++)
	{
		const int o = 16 >> st, half = M >> (st + 1);
		const bool up = (lane & o) != 0;
#pragma unroll
		for (int i = 0; i < half; i++)
		{
			const double send = up ? w[i] : w[i + half], keep = up ? w[i + half] : w[i];
			w[i] = keep + __shfl_xor_sync (0xffffffffu, send, o);
		}
	}
#pragma unroll
	for (int o = 16 >> L; o > 0; o >>= 1)
		w[0] += __shfl_xor_sync (0xffffffffu, w[0], o);
#pragma unroll
	for (int k = 0; k < N; k++)
		v[k] = __shfl_sync (0xffffffffu, w[0], k << (5 - L));
}
#endif

// Block-wide sum of N doubles; the result is broadcast into v[] of every thread.  Fixed tree => deterministic.
template <int N> NYX_D void cta_sum (const Cta& c, double (&v)[N])
{
#if NYX_DEVICE_BUILD
	static_assert (N <= RED_MAX, "too many values");
	if (c.wmode)
	{
		if constexpr (N >= 2 && N <= 32)
			warp_allsum<N> (v, c.tid);
		else
		{
#pragma unroll
			for (int k = 0; k < N; k++)
			{
				double x = v[k];
#pragma unroll
				for (int o = 16; o > 0; o >>= 1)
					x += __shfl_xor_sync (0xffffffffu, x, o);
				v[k] = x;
			}
		}
		return;
	}
	const int lane = c.tid & 31, warp = c.tid >> 5, nw = (c.nthr + 31) >> 5;
#pragma unroll
	for (int k = 0; k < N; k++)
	{
		double x = v[k];
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
			x += __shfl_xor_sync (0xffffffffu, x, o);
		if (lane == 0)
			c.red[k * NW_MAX + warp] = x;
	}
	__syncthreads();
#pragma unroll
	for (int k = 0; k < N; k++)
	{
		double x = 0;
		for (int w = 0; w < nw; w++)
			x += c.red[k * NW_MAX + w];
		v[k] = x;
	}
	__syncthreads();
#else
	(void)c; (void)v;
#endif
}

#if NYX_DEVICE_BUILD
// exact warp-wide sum / max of 64-bit integers with redux.sync (one instruction per 32-bit piece instead of a
// 5-step shuffle butterfly): 26 + 26 + 12 bit pieces cannot overflow a 32-bit lane sum
NYX_D unsigned long long warp_sum_u64 (unsigned long long x)
{
	const uint32_t lo = (uint32_t)x & 0x3FFFFFFu, mid = (uint32_t)(x >> 26) & 0x3FFFFFFu, hi = (uint32_t)(x >> 52);
	return (unsigned long long)__reduce_add_sync (0xffffffffu, lo) + ((unsigned long long)__reduce_add_sync (0xffffffffu, mid) << 26)
		+ ((unsigned long long)__reduce_add_sync (0xffffffffu, hi) << 52);
}
NYX_D unsigned long long warp_max_u64 (unsigned long long x)
{
	const uint32_t hi = (uint32_t)(x >> 32), mh = __reduce_max_sync (0xffffffffu, hi);
	const uint32_t ml = __reduce_max_sync (0xffffffffu, hi == mh ? (uint32_t)x : 0u);
	return ((unsigned long long)mh << 32) | ml;
}
#endif

template <int N> NYX_D void cta_sum_u64 (const Cta& c, unsigned long long (&v)[N])
{
#if NYX_DEVICE_BUILD
	static_assert (N <= RED_MAX, "too many values");
	if (c.wmode)
	{
#pragma unroll
		for (int k = 0; k < N; k++)
			v[k] = warp_sum_u64 (v[k]);
		return;
	}
	const int lane = c.tid & 31, warp = c.tid >> 5, nw = (c.nthr + 31) >> 5;
	unsigned long long* red = (unsigned long long*)c.red;
#pragma unroll
	for (int k = 0; k < N; k++)
	{
		const unsigned long long x = warp_sum_u64 (v[k]);
		if (lane == 0)
			red[k * NW_MAX + warp] = x;
	}
	__syncthreads();
#pragma unroll
	for (int k = 0; k < N; k++)
	{
		unsigned long long x = 0;
		for (int w = 0; w < nw; w++)
			x += red[k * NW_MAX + w];
		v[k] = x;
	}
	__syncthreads();
#else
	(void)c; (void)v;
#endif
}

NYX_D unsigned long long cta_max_u64 (const Cta& c, unsigned long long x)
{
#if NYX_DEVICE_BUILD
	const int lane = c.tid & 31, warp = c.tid >> 5, nw = (c.nthr + 31) >> 5;
	unsigned long long* red = (unsigned long long*)c.red;
	x = warp_max_u64 (x);
	if (c.wmode)
		return x;
	if (lane == 0)
		red[warp] = x;
	__syncthreads();
	x = 0;
	for (int w = 0; w < nw; w++)
		x = red[w] > x ? red[w] : x;
	__syncthreads();
	return x;
#else
	(void)c;
	return x;
#endif
}

// ------------------------------------------------------------------------------------------------
// In-shared-memory bitonic sort of npad (power of two) keys, ascending
// ------------------------------------------------------------------------------------------------
NYX_HD uint32_t next_pow2 (uint32_t n)
{
	uint32_t p = 1;
	while (p < n) p <<= 1;
	return p;
}

NYX_D void bitonic_sort (const Cta& c, uint32_t* a, uint32_t npad)
{
	for (uint32_t k = 2; k <= npad; k <<= 1)
		for (uint32_t j = k >> 1; j > 0; j >>= 1)
		{
			for (uint32_t i = c.tid; i < npad; i += c.nthr)
			{
				uint32_t ixj = i ^ j;
				if (ixj > i)
				{
					uint32_t x = a[i], y = a[ixj];
					bool up = (i & k) == 0;
					if ((x > y) == up)
					{
						a[i] = y;
						a[ixj] = x;
					}
				}
			}
			c.sync();
		}
}

#if NYX_DEVICE_BUILD
// One warp sorts 32*R keys held in REGISTERS (R per lane): a bitonic network over the element index e = lane*R + r.
// Exchanges at distance j < R stay inside a lane (a min/max pair), j >= R are one shuffle per register; nothing goes
// through shared memory until the sorted keys are stored.  In: a[] filled from srt[r*32 + lane] (any order is a valid
// input).  Out: srt[0 .. 32R) ascending.
template <int R> NYX_D void warp_sort_regs (uint32_t* srt, int lane)
{
	uint32_t a[R];
#pragma unroll
	for (int r = 0; r < R; r++)
		a[r] = srt[r * 32 + lane];
	__syncwarp();
#pragma unroll
	for (int k = 2; k <= 32 * R; k <<= 1)
	{
		// direction of the k-block this element is in: ascending when (e & k) == 0
		const bool lane_up = k >= R ? (lane & (k / R)) == 0 : true;	// k >= R: the bit lives in the lane part of e
#pragma unroll
		for (int j = k >> 1; j > 0; j >>= 1)
		{
			if (j >= R)
			{
				const bool lower = (lane & (j / R)) == 0;
				const bool keep_min = lower == lane_up;
#pragma unroll
				for (int r = 0; r < R; r++)
				{
					const uint32_t t = __shfl_xor_sync (0xffffffffu, a[r], j / R);
					a[r] = keep_min ? min (a[r], t) : max (a[r], t);
				}
			}
			else
			{
#pragma unroll
				for (int r = 0; r < R; r++)
					if ((r & j) == 0)
					{
						const bool up = k >= R ? lane_up : (r & k) == 0;
						const uint32_t x = a[r], y = a[r | j];
						const uint32_t mn = min (x, y), mx = max (x, y);
						a[r] = up ? mn : mx;
						a[r | j] = up ? mx : mn;
					}
			}
		}
	}
	if (R >= 4)
	{
		uint4* d4 = reinterpret_cast<uint4*> (srt + lane * R);	// 16-byte aligned: srt is, and R is a multiple of 4
#pragma unroll
		for (int r = 0; r + 3 < R; r += 4)
			d4[r >> 2] = make_uint4 (a[r], a[r + 1], a[r + 2], a[r + 3]);
	}
	else
	{
#pragma unroll
		for (int r = 0; r < R; r++)
			srt[lane * R + r] = a[r];
	}
	__syncwarp();
}
#endif

// ascending sort of npad keys (power of two, >= 64 for a warp-wide context) in shared memory
NYX_D void sort_keys (const Cta& c, uint32_t* a, uint32_t npad)
{
#if NYX_DEVICE_BUILD
	if (c.wmode && npad >= 64 && npad <= 1024)
	{
		switch (npad)
		{
		case 64: warp_sort_regs<2> (a, c.tid); break;
		case 128: warp_sort_regs<4> (a, c.tid); break;
		case 256: warp_sort_regs<8> (a, c.tid); break;
		case 512: warp_sort_regs<16> (a, c.tid); break;
		default: warp_sort_regs<32> (a, c.tid); break;
		}
		return;
	}
#endif
	bitonic_sort (c, a, npad);
}

// ------------------------------------------------------------------------------------------------
// Tile materialisation: the ROI's AABB read straight from the resident slide (reference:
// ImageMatrix::calculate_from_pixelcloud, image_matrix.h:284-304 -- 0 outside the ROI).
// cell[i] = binned level | CELL_NZ (orig != 0) | CELL_IN (label match)
// ------------------------------------------------------------------------------------------------
NYX_D void load_tile (const Cta& c, const ImageView& im, const RoiRec& r, const Binner& B, uint32_t* cell)
{
	const int n = r.w * r.h;
	FastDiv fw;
	fw.init ((uint32_t)r.w);
	for (int i = c.tid; i < n; i += c.nthr)
	{
		int y = (int)fw.div ((uint32_t)i), x = i - y * r.w;
		size_t g = (size_t)(r.y0 + y) * im.W + (size_t)(r.x0 + x);
		bool in = im.label[g] == r.label;
		uint32_t v = in ? load_inten (im, g) : 0u;
		uint32_t code = B.bin (v) & CELL_LEVEL;
		if (v) code |= CELL_NZ;
		if (in) code |= CELL_IN;
		cell[i] = code;
	}
}

// 16-bit code of the binned-tile store: level in the low 14 bits, CELL_IN -> bit 14, CELL_NZ -> bit 15
constexpr uint32_t TILE16_LEVEL = 0x3FFFu;
NYX_HD uint32_t tile16_encode (uint32_t cellcode)
{
	return (cellcode & TILE16_LEVEL) | ((cellcode >> 16) & 0xC000u);
}
NYX_HD uint32_t tile16_decode (uint32_t code16)
{
	return (code16 & TILE16_LEVEL) | ((code16 & 0xC000u) << 16);
}
// words of the shared-memory tile of a class (room for the 8-cell padding of a stored tile)
NYX_HD size_t tile_words (uint32_t maxcells)
{
	return ((size_t)maxcells + 7) & ~(size_t)7;
}

// tile from the binned-tile store: coalesced 32-bit loads of two codes each (the slot is padded to 8 cells and 16-byte
// aligned; cell[] has tile_words() room)
NYX_D void load_tile_store (const Cta& c, const uint16_t* slot, int ncell, uint32_t* cell)
{
	const uint32_t* s2 = (const uint32_t*)slot;
	const int n2 = (ncell + 1) >> 1;
	for (int k = c.tid; k < n2; k += c.nthr)
	{
		const uint32_t u = s2[k];
		cell[2 * k] = tile16_decode (u & 0xFFFFu);
		cell[2 * k + 1] = tile16_decode (u >> 16);
	}
}

// 3D: the ROI's AABB cube, cell index (z*h + y)*w + x (reference: SimpleCube::calculate_from_pixelcloud, image_cube.h:43-66)
NYX_D void load_cube (const Cta& c, const ImageView& im, const RoiRec& r, const Binner& B, uint32_t* cell)
{
	const int n = r.w * r.h * r.d;
	FastDiv fw, fh;
	fw.init ((uint32_t)r.w);
	fh.init ((uint32_t)r.h);
	for (int i = c.tid; i < n; i += c.nthr)
	{
		const int row = (int)fw.div ((uint32_t)i), x = i - row * r.w;
		const int z = (int)fh.div ((uint32_t)row), y = row - z * r.h;
		size_t g = ((size_t)(r.z0 + z) * im.H + (size_t)(r.y0 + y)) * im.W + (size_t)(r.x0 + x);
		bool in = im.label[g] == r.label;
		uint32_t v = in ? load_inten (im, g) : 0u;
		uint32_t code = B.bin (v) & CELL_LEVEL;
		if (v) code |= CELL_NZ;
		if (in) code |= CELL_IN;
		cell[i] = code;
	}
}

} // namespace nyx
