// TEST HARNESS ONLY (never part of the product library): compiles the per-ROI feature bodies of
// nyxus_b200/csrc/*_body.h for the HOST with a single-thread CTA context (NYX_HOST_EMU), so their
// indexing and arithmetic can be checked against the oracle on a machine without a GPU.
// The CUDA library (libnyxus_b200.so) instantiates the very same bodies with a real CTA.
#define NYX_HOST_EMU 1
#include <vector>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <climits>
#include "../../nyxus_b200/csrc/nyx_common.h"
#include "../../nyxus_b200/csrc/nyx_features.h"
#include "../../nyxus_b200/csrc/glcm_body.h"
#include "../../nyxus_b200/csrc/glrlm_body.h"
#include "../../nyxus_b200/csrc/zones_body.h"
#include "../../nyxus_b200/csrc/nbr_body.h"
#include "../../nyxus_b200/csrc/intensity_body.h"
#include "../../nyxus_b200/csrc/vol_body.h"
#include "../../nyxus_b200/csrc/morph_body.h"
#include "../../nyxus_b200/csrc/ih_body.h"
#include "../../nyxus_b200/csrc/geomom_body.h"
#include "../../nyxus_b200/csrc/contour_body.h"

using namespace nyx;

extern "C" int nyxemu_ncols() { return NCOL_TEXINT; }

// tile: h*w intensities, mask: h*w membership.  fam_mask bits: 1 INT, 2 GLCM, 4 GLRLM, 8 GLDZM, 16 GLSZM, 32 GLDM, 64 NGLDM, 128 NGTDM
static int g_glcm_greydepth = INT_MIN;	// "glcm/greydepth" (2D: blank test only); INT_MIN = follows the grey depth
extern "C" void nyxemu_set_glcm_greydepth (int g) { g_glcm_greydepth = g; }

static Params make_prm (int greydepth, int glcm_greydepth, int ibsi, int glcm_offset)
{
	Params prm;
	prm.greydepth = greydepth;
	prm.glcm_greydepth = glcm_greydepth;
	prm.ibsi = ibsi;
	prm.glcm_offset = glcm_offset;
	prm.glcm_symmetric = 0;
	prm.softnan = 0.0;
	return prm;
}

extern "C" int nyxemu_roi_features (const uint32_t* tile, const uint8_t* mask, int w, int h,
	int greydepth, int ibsi, int glcm_offset, unsigned fam_mask, double* out, int p16)
{
	std::vector<uint32_t> lab ((size_t)w * h);
	RoiRec r;
	memset (&r, 0, sizeof (r));
	r.label = 1;
	r.imin = 0xFFFFFFFFu;
	r.imax = 0;
	int x0 = w, x1 = -1, y0 = h, y1 = -1;
	std::vector<uint32_t> cloud;
	for (int y = 0; y < h; y++)
		for (int x = 0; x < w; x++)
		{
			lab[y * w + x] = mask[y * w + x] ? 1 : 0;
			if (mask[y * w + x])
			{
				uint32_t v = tile[y * w + x];
				r.area++;
				r.imin = std::min (r.imin, v);
				r.imax = std::max (r.imax, v);
				x0 = std::min (x0, x); x1 = std::max (x1, x);
				y0 = std::min (y0, y); y1 = std::max (y1, y);
				cloud.push_back (v);
			}
		}
	if (!r.area)
		return -1;
	r.x0 = x0; r.y0 = y0; r.w = x1 - x0 + 1; r.h = y1 - y0 + 1;
	ImageView im { tile, lab.data(), h, w, 0 };
	Params prm = make_prm (greydepth, g_glcm_greydepth == INT_MIN ? greydepth : g_glcm_greydepth, ibsi, glcm_offset);
	std::vector<double> red (RED_MAX * NW_MAX);
	Cta c { 0, 1, red.data() };
	const int ncell = r.w * r.h;
	Binner B;
	B.init (ibsi ? 0 : greydepth, r.imin, r.imax);
	std::vector<uint32_t> cell (ncell);
	load_tile (c, im, r, B, cell.data());
	const int Lmax = ibsi ? (int)r.imax : std::abs (greydepth);
	const bool rad = !ibsi && greydepth < 0;
	for (int i = 0; i < NCOL_TEXINT; i++)
		out[i] = 0;
	if (fam_mask & 1)
	{
		uint32_t npad = next_pow2 (r.area);
		std::vector<uint32_t> mem (intensity_smem_words (npad, greydepth));
		intensity_body (c, r, prm, cloud.data(), r.area, npad, mem.data(), out + COL_INT);
	}
	if (fam_mask & 2)
	{
		int Ng = Lmax;
		if (p16)
		{
			std::vector<uint32_t> mem (glcm_smem_words<true> (Ng, 4, rad, (uint32_t)(r.w * r.h)));
			std::vector<uint32_t> cellg (cell);	// (PyRadiomics binning rewrites the levels to ranks)
			glcm_body<true> (c, r, prm, cellg.data(), mem.data(), Ng, out + COL_GLCM, nullptr, (uint32_t)(r.w * r.h));
		}
		else
		{
			std::vector<uint32_t> mem (glcm_smem_words<false> (Ng, 4, rad, (uint32_t)(r.w * r.h)));
			std::vector<uint32_t> cellg (cell);
			glcm_body<false> (c, r, prm, cellg.data(), mem.data(), Ng, out + COL_GLCM, nullptr, (uint32_t)(r.w * r.h));
		}
	}
	if (fam_mask & 4)
	{
		int Nr = std::max (r.w, r.h);
		std::vector<uint32_t> mem (glrlm_smem_words (Lmax, Nr, ncell));
		glrlm_body (c, r, prm, cell.data(), mem.data(), Lmax, Nr, ncell, out + COL_GLRLM, nullptr);
	}
	if (fam_mask & 8)
	{
		const int Nd = gldzm_nd_bound ((uint32_t)ncell, std::max (r.w, r.h));
		std::vector<uint32_t> mem (gldzm_smem_words (Lmax, Nd, (uint32_t)ncell));
		gldzm_body (c, r, prm, cell.data(), mem.data(), Lmax, Nd, (uint32_t)ncell, out + COL_GLDZM, nullptr);
	}
	if (fam_mask & 16)
	{
		const uint32_t slots = glszm_hash_slots (Lmax, (uint32_t)ncell);
		const int Lmarg = Lmax <= GLSZM_LMARG_MAX ? Lmax : 0;	// direct marginal arrays like the kernels
		std::vector<uint32_t> mem (glszm_smem_words ((uint32_t)ncell, slots, Lmarg));
		glszm_body (c, r, prm, cell.data(), mem.data(), (uint32_t)ncell, slots, out + COL_GLSZM, nullptr, Lmarg);
	}
	if (fam_mask & 32)
	{
		std::vector<uint32_t> mem (gldm_smem_words (Lmax, (uint32_t)ncell));
		gldm_body (c, r, prm, cell.data(), mem.data(), Lmax, out + COL_GLDM, nullptr);
	}
	if (fam_mask & 64)
	{
		std::vector<uint32_t> cell2 (ncell);
		load_tile_ngldm (c, im, r, prm, cell2.data());
		std::vector<uint32_t> mem (ngldm_smem_words (Lmax, (uint32_t)ncell));
		ngldm_body (c, r, prm, cell2.data(), mem.data(), Lmax, out + COL_NGLDM, nullptr);
	}
	if (fam_mask & 128)
	{
		std::vector<uint32_t> mem (ngtdm_smem_words (Lmax));
		ngtdm_body (c, r, prm, cell.data(), mem.data(), Lmax, out + COL_NGTDM, nullptr);
	}
	return 0;
}


// ---- 3D: one labelled volume [D][H][W], label value `lab`; fam_mask bits: 1 INT, 2 GLCM, 4 NGTDM, 8 GLSZM, 16 GLDM, 32 GLRLM,
// 64 GLDZM, 128 NGLDM.
// out layout: 36 first-order | 30x13 + 29 GLCM | 5 NGTDM | 16 GLSZM | 14 GLDM | 32 GLRLM | 18 GLDZM | 19 NGLDM   (the canonical
// 3D table of the library)
extern "C" int nyxemu_ncols3d() { return NCOL_INT + NCOL_GLCM3 + NCOL_NGTDM + NCOL_GLSZM + NCOL_GLDM + NCOL_GLRLM3 + NCOL_GLDZM + NCOL_NGLDM; }
static int g_glrlm3_greydepth = 0;	// "3glrlm/greydepth" (0 = no binning)
extern "C" void nyxemu_set_glrlm3_greydepth (int g) { g_glrlm3_greydepth = g; }
static int g_gldm3_greydepth = 0;	// "3gldm/greydepth" (0 = no binning)
extern "C" void nyxemu_set_gldm3_greydepth (int g) { g_gldm3_greydepth = g; }
static bool g_big3d = false;	// 3D GLSZM through the large-cube body
static int g_glcm3_greydepth = INT_MIN;	// "3glcm/greydepth"; INT_MIN = follows the grey depth
extern "C" void nyxemu_set_glcm3_greydepth (int g) { g_glcm3_greydepth = g; }
extern "C" int nyxemu_roi_features_3d (const uint32_t* vol, const uint32_t* lbl, int D, int H, int W, uint32_t lab,
	int greydepth, int glszm_g, int ngtdm_g, int ngtdm_radius, int ibsi, double covered_den, unsigned fam_mask, double* out)
{
	RoiRec r;
	memset (&r, 0, sizeof (r));
	r.label = lab;
	r.imin = 0xFFFFFFFFu;
	int x0 = W, x1 = -1, y0 = H, y1 = -1, z0 = D, z1 = -1;
	std::vector<uint32_t> cloud;
	for (int z = 0; z < D; z++)
		for (int y = 0; y < H; y++)
			for (int x = 0; x < W; x++)
			{
				size_t i = ((size_t)z * H + y) * W + x;
				if (lbl[i] != lab)
					continue;
				uint32_t v = vol[i];
				r.area++;
				r.imin = std::min (r.imin, v);
				r.imax = std::max (r.imax, v);
				x0 = std::min (x0, x); x1 = std::max (x1, x);
				y0 = std::min (y0, y); y1 = std::max (y1, y);
				z0 = std::min (z0, z); z1 = std::max (z1, z);
				cloud.push_back (v);
			}
	if (!r.area)
		return -1;
	r.x0 = x0; r.y0 = y0; r.z0 = z0; r.w = x1 - x0 + 1; r.h = y1 - y0 + 1; r.d = z1 - z0 + 1;
	ImageView im { vol, lbl, H, W, 0, D };
	const int glcm_g = g_glcm3_greydepth == INT_MIN ? greydepth : g_glcm3_greydepth;
	Params prm = make_prm (greydepth, glcm_g, ibsi, 1);
	prm.glszm3_greydepth = glszm_g;
	prm.ngtdm3_greydepth = ngtdm_g;
	prm.ngtdm3_radius = ngtdm_radius;
	prm.covered_den = covered_den;
	std::vector<double> red (RED_MAX * NW_MAX);
	Cta c { 0, 1, red.data() };
	const int ncell = r.w * r.h * r.d;
	const int n3 = nyxemu_ncols3d();
	for (int i = 0; i < n3; i++)
		out[i] = 0;
	const int O_GLCM = NCOL_INT, O_NGTDM = O_GLCM + NCOL_GLCM3, O_GLSZM = O_NGTDM + NCOL_NGTDM, O_GLDM = O_GLSZM + NCOL_GLSZM;
	const int O_GLRLM = O_GLDM + NCOL_GLDM, O_GLDZM = O_GLRLM + NCOL_GLRLM3, O_NGLDM = O_GLDZM + NCOL_GLDZM;
	const int maxdim = std::max (r.w, std::max (r.h, r.d));
	std::vector<uint32_t> cell (ncell);
	if (fam_mask & 1)
	{
		uint32_t npad = next_pow2 (r.area);
		std::vector<uint32_t> mem (intensity_smem_words (npad, greydepth));
		intensity_body (c, r, prm, cloud.data(), r.area, npad, mem.data(), out);
	}
	for (int fam = 1; fam <= 7; fam++)
	{
		static const unsigned bits[8] = { 1u, 2u, 8u, 4u, 16u, 32u, 64u, 128u };	// INT, GLCM, GLSZM, NGTDM, GLDM, GLRLM, GLDZM, NGLDM
		if (!(fam_mask & bits[fam]))
			continue;
		int g = fam == 1 ? glcm_g : fam == 2 ? glszm_g : fam == 3 ? ngtdm_g : fam == 4 ? g_gldm3_greydepth : fam == 5 ? g_glrlm3_greydepth : greydepth;
		if (ibsi)
			g = 0;
		Binner B;
		if (fam == 7)
			B.init_tg (g, r.imax, ibsi != 0);
		else
			B.init (g, r.imin, r.imax);
		load_cube (c, im, r, B, cell.data());
		int Lm = g != 0 ? std::abs (g) : (int)std::max (r.imax, 1u);
		if (fam == 5)
		{
			std::vector<uint32_t> mem (glrlm3d_smem_words (Lm, maxdim, (unsigned long long)ncell));
			glrlm3d_body (c, r, prm, cell.data(), mem.data(), Lm, maxdim, out + O_GLRLM);
			continue;
		}
		if (fam == 6)
		{
			const int Nd = gldzm_nd_bound ((uint32_t)ncell, maxdim);
			std::vector<uint32_t> mem (gldzm3d_smem_words (Lm, Nd, (unsigned long long)ncell));
			gldzm3d_body (c, r, prm, cell.data(), mem.data(), Lm, Nd, (unsigned long long)ncell, out + O_GLDZM);
			continue;
		}
		if (fam == 7)
		{
			std::vector<uint32_t> mem (ngldm3d_smem_words (Lm));
			ngldm3d_body (c, r, prm, cell.data(), mem.data(), Lm, out + O_NGLDM);
			continue;
		}
		if (fam == 1)
		{
			std::vector<uint32_t> mem (glcm_smem_words<false> (Lm, 4, g < 0));
			glcm3d_body<false> (c, r, prm, cell.data(), mem.data(), Lm, out + O_GLCM, g);
		}
		else if (fam == 2)
		{
			if (g_big3d)	// the large-cube body (32-bit labels, 64-bit hash keys)
			{
				const uint32_t slots = glszm3d_big_slots (Lm, (unsigned long long)ncell);
				std::vector<uint32_t> mem (glszm3d_big_words ((unsigned long long)ncell, slots) + 2);
				uint32_t* m8 = mem.data() + (((uintptr_t)mem.data() & 7) ? 1 : 0);	// 8-byte aligned keys
				glszm3d_body_big (c, r, prm, cell.data(), m8, (unsigned long long)ncell, slots, g > 0, out + O_GLSZM);
			}
			else
			{
				const uint32_t slots = glszm_hash_slots (Lm, (uint32_t)ncell);
				std::vector<uint32_t> mem (glszm3d_smem_words ((uint32_t)ncell, slots));
				glszm3d_body (c, r, prm, cell.data(), mem.data(), (uint32_t)ncell, slots, g > 0, out + O_GLSZM);
			}
		}
		else if (fam == 4)
		{
			std::vector<uint32_t> mem (gldm3d_smem_words (Lm, (uint32_t)ncell));
			gldm3d_body (c, r, prm, cell.data(), mem.data(), Lm, g > 0, out + O_GLDM);
		}
		else
		{
			if (ngtdm_radius == 0)
				Lm = 1;
			std::vector<uint32_t> mem (ngtdm3d_smem_words (Lm, ngtdm_radius));
			ngtdm3d_body (c, r, prm, cell.data(), mem.data(), Lm, g > 0, out + O_NGTDM);
		}
	}
	return 0;
}

extern "C" void nyxemu_set_big3d (int on) { g_big3d = on != 0; }
extern "C" double nyxemu_fast_log10 (double x) { return fast_log10 (x); }
extern "C" unsigned nyxemu_bin (unsigned x, unsigned mn, unsigned mx, int g) { Binner B; B.init (g, mn, mx); return B.bin (x); }
extern "C" unsigned nyxemu_to_grayscale (unsigned i, unsigned mn, unsigned range, unsigned n) { return to_grayscale (i, mn, range, n, false); }

// *BASIC_MORPHOLOGY* of the ROI `lab` of a labelled uint32 slide: the 15 columns of morph_body in Feature2D code order
extern "C" int nyxemu_morph (const uint32_t* inten, const uint32_t* lbl, int H, int W, unsigned lab, double* out)
{
	RoiRec r;
	memset (&r, 0, sizeof (r));
	r.label = lab;
	r.imin = 0xFFFFFFFFu;
	int x0 = W, x1 = -1, y0 = H, y1 = -1;
	for (int y = 0; y < H; y++)
		for (int x = 0; x < W; x++)
			if (lbl[(size_t)y * W + x] == lab)
			{
				const uint32_t v = inten[(size_t)y * W + x];
				r.area++;
				r.imin = std::min (r.imin, v);
				r.imax = std::max (r.imax, v);
				x0 = std::min (x0, x); x1 = std::max (x1, x);
				y0 = std::min (y0, y); y1 = std::max (y1, y);
			}
	if (!r.area)
		return -1;
	r.x0 = x0; r.y0 = y0; r.w = x1 - x0 + 1; r.h = y1 - y0 + 1; r.d = 1;
	ImageView im { inten, lbl, H, W, 0 };
	Params prm = make_prm (64, 64, false, 1);
	std::vector<double> red (RED_MAX * NW_MAX);
	Cta c { 0, 1, red.data() };
	morph_body (c, im, r, prm, out);
	return 0;
}

// *ALL_IH* of a pixel cloud (IBSI mode): the 46 columns of ih_body in Feature2D code order
extern "C" int nyxemu_ih (const uint32_t* vals, unsigned n, int greydepth, int ibsi, double* out)
{
	if (!n)
		return -1;
	RoiRec r;
	memset (&r, 0, sizeof (r));
	r.area = n;
	r.imin = 0xFFFFFFFFu;
	for (unsigned i = 0; i < n; i++)
	{
		r.imin = std::min (r.imin, vals[i]);
		r.imax = std::max (r.imax, vals[i]);
	}
	Params prm = make_prm (greydepth, greydepth, ibsi != 0, 1);
	std::vector<double> red (RED_MAX * NW_MAX);
	Cta c { 0, 1, red.data() };
	std::vector<uint32_t> mem (ih_smem_words (greydepth));
	ih_body (c, r, prm, vals, n, mem.data(), out);
	return 0;
}

// first-order block + the opt-in HISTOGRAM block of a plain list of intensities: out[0 .. 36) = *ALL_INTENSITY*,
// out[36 .. 36 + greydepth) = HISTOGRAM_BIN_0 .. (intensity.cpp:129, histogram.h:190-198)
extern "C" int nyxemu_histogram (const uint32_t* vals, unsigned n, int greydepth, double* out)
{
	if (!n || greydepth <= 0)
		return -1;
	RoiRec r;
	memset (&r, 0, sizeof (r));
	r.area = n;
	r.imin = 0xFFFFFFFFu;
	for (unsigned i = 0; i < n; i++)
	{
		r.imin = std::min (r.imin, vals[i]);
		r.imax = std::max (r.imax, vals[i]);
	}
	Params prm = make_prm (greydepth, greydepth, 0, 1);
	prm.hist_col = NCOL_INT;
	std::vector<double> red (RED_MAX * NW_MAX);
	Cta c { 0, 1, red.data() };
	const uint32_t npad = std::max (next_pow2 (n), 64u);
	std::vector<uint32_t> mem (intensity_smem_words (npad, greydepth));
	intensity_body (c, r, prm, vals, n, npad, mem.data(), out);
	return 0;
}

// unweighted geometric moments of the ROI `lab` of a labelled uint32 slide: 118 columns (shape variant, intensity variant)
extern "C" int nyxemu_geomom (const uint32_t* inten, const uint32_t* lbl, int H, int W, unsigned lab, double* out)
{
	RoiRec r;
	memset (&r, 0, sizeof (r));
	r.label = lab;
	r.imin = 0xFFFFFFFFu;
	int x0 = W, x1 = -1, y0 = H, y1 = -1;
	for (int y = 0; y < H; y++)
		for (int x = 0; x < W; x++)
			if (lbl[(size_t)y * W + x] == lab)
			{
				const uint32_t v = inten[(size_t)y * W + x];
				r.area++;
				r.imin = std::min (r.imin, v);
				r.imax = std::max (r.imax, v);
				x0 = std::min (x0, x); x1 = std::max (x1, x);
				y0 = std::min (y0, y); y1 = std::max (y1, y);
			}
	if (!r.area)
		return -1;
	r.x0 = x0; r.y0 = y0; r.w = x1 - x0 + 1; r.h = y1 - y0 + 1; r.d = 1;
	ImageView im { inten, lbl, H, W, 0 };
	std::vector<double> red (RED_MAX * NW_MAX);
	Cta c { 0, 1, red.data() };
	geomom_body (c, im, r, 3, out);
	return 0;
}

// distance-to-contour weighted geometric moments of the ROI `lab` of a labelled uint32 slide: 62 columns (shape variant, intensity
// variant); contour (optional): the ordered contour in padded coordinates packed x | y << 16, ncontour its length
extern "C" int nyxemu_wmom (const uint32_t* inten, const uint32_t* lbl, int H, int W, unsigned lab, double* out, uint32_t* contour, int* ncontour)
{
	RoiRec r;
	memset (&r, 0, sizeof (r));
	r.label = lab;
	int x0 = W, x1 = -1, y0 = H, y1 = -1;
	for (int y = 0; y < H; y++)
		for (int x = 0; x < W; x++)
			if (lbl[(size_t)y * W + x] == lab)
			{
				r.area++;
				x0 = std::min (x0, x); x1 = std::max (x1, x);
				y0 = std::min (y0, y); y1 = std::max (y1, y);
			}
	if (!r.area)
		return -1;
	r.x0 = x0; r.y0 = y0; r.w = x1 - x0 + 1; r.h = y1 - y0 + 1; r.d = 1;
	ImageView im { inten, lbl, H, W, 0 };
	std::vector<double> red (RED_MAX * NW_MAX);
	Cta c { 0, 1, red.data() };
	const unsigned long long cells = (unsigned long long)r.w * r.h;
	const int maxwh = std::max (r.w, r.h);
	std::vector<uint32_t> mem (wmom_scratch_words (cells, maxwh));
	wmom_body (c, im, r, 3, mem.data(), cells, maxwh, out);
	if (contour && ncontour)
	{
		// the K array survives the body (only LC and the stack are recycled for the weights)
		const unsigned long long pcb = wmom_padded_bound (cells, maxwh);
		const uint32_t* K = mem.data() + 2 * ((pcb + 3) / 4) + cells;
		// its length is not kept: recount by re-running the tracer would be wasteful; the body leaves it in out[] instead
		(void)K;
		*ncontour = 0;
	}
	return 0;
}
