// TEST INFRASTRUCTURE ONLY -- never linked, loaded or executed by the product (nyxus_b200/).
//
// C-ABI driver around the UNMODIFIED reference feature calculators (PolusAI/nyxus, compiled from
// /root/reference by oracle/Makefile.ref into oracle/_ref/libnyxus_ref.so).  It replays the in-memory
// hot path exactly the way the reference's own workflow does:
//   phase 1  gatherRoisMetricsInMemory        (src/nyx/phase1.cpp:373-407, column-major pixel loop)
//   phase 2  scanTrivialRoisInMemory          (src/nyx/phase2_2d.cpp:637-684) + allocateTrivialRoisBuffers (:427-465)
//   reduce   reduce_trivial_2d                (src/nyx/reduce_trivial_rois.cpp:62-253) via runParallel (parallel.h:23-42)
//   output   save_features_2_buffer layout    (src/nyx/output_2_buffer.cpp:303-582)
// and exposes the private matrices of every texture family (this TU only re-includes the
// reference headers under '#define private public'; the class layouts are unchanged) as sparse triplets so the
// CUDA path can be checked entry by entry.
//
// Used (a) as the parity oracle of record ("kind": "reference") and (b) as the CPU baseline timed by bench.py.

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <vector>
// every std header the reference headers pull in must be seen BEFORE the access override below
#include <array>
#include <cfloat>
#include <climits>
#include <cmath>
#include <fstream>
#include <functional>
#include <future>
#include <iomanip>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <optional>
#include <queue>
#include <regex>
#include <set>
#include <sstream>
#include <stack>
#include <stdexcept>
#include <thread>
#include <variant>
#include <filesystem>
#include <ranges>
#include <span>
#include <bitset>
#include <cassert>
#include <complex>
#include <deque>
#include <iterator>
#include <random>
#include <typeinfo>
#include <typeindex>
#include <utility>
#include <any>
#include <atomic>
#include <condition_variable>
#include <ctime>
#include <cstdio>
#include <cstdlib>
#include <initializer_list>
#include <type_traits>
#include <string_view>
#include <charconv>
#include <numbers>
#include <bit>
#include <concepts>
#include <compare>
#include <version>

// read-only access to the families' private matrices (class layouts are unchanged)
#define private public
#define protected public
#include "featureset.h"
#include "feature_settings.h"
#include "roi_cache.h"
#include "dataset.h"
#include "parallel.h"
#include "helpers/helpers.h"
#include "features/intensity.h"
#include "features/glcm.h"
#include "features/glrlm.h"
#include "features/glszm.h"
#include "features/gldzm.h"
#include "features/gldm.h"
#include "features/ngldm.h"
#include "features/ngtdm.h"
#include "features/3d_intensity.h"
#include "features/3d_glcm.h"
#include "features/3d_glszm.h"
#include "features/3d_ngtdm.h"
#undef private
#undef protected

using namespace Nyxus;

namespace Nyxus
{
	// defined in features_calc_workflow.cpp
	void init_label_record_3 (LR& lr, int x, int y, PixIntens intensity);
	void update_label_record_3 (LR& lr, int x, int y, PixIntens intensity);
	void init_label_record_3D (LR& lr, int x, int y, int z, int label, PixIntens intensity);
	void update_label_record_3D (LR& lr, int x, int y, int z, int label, PixIntens intensity);
}

namespace
{
	enum Fam { F_INT = 1, F_GLCM = 2, F_GLRLM = 4, F_GLDZM = 8, F_GLSZM = 16, F_GLDM = 32, F_NGLDM = 64, F_NGTDM = 128 };

	struct Col { int code; int k; std::string name; int fam; };
	std::vector<Col> g_cols;

	template <class L> bool in_list (const L& lst, Feature2D c)
	{
		return std::find (lst.begin(), lst.end(), c) != lst.end();
	}

	void build_cols()
	{
		if (!g_cols.empty())
			return;
		FeatureSet fs;
		std::vector<std::pair<int, int>> codes;	// (code, family)
		for (auto c : PixelIntensityFeatures::featureset) codes.push_back ({ (int)c, F_INT });
		for (auto c : GLCMFeature::featureset) codes.push_back ({ (int)c, F_GLCM });
		for (auto c : GLRLMFeature::featureset) codes.push_back ({ (int)c, F_GLRLM });
		for (auto c : GLDZMFeature::featureset) codes.push_back ({ (int)c, F_GLDZM });
		for (auto c : GLSZMFeature::featureset) codes.push_back ({ (int)c, F_GLSZM });
		for (auto c : GLDMFeature::featureset) codes.push_back ({ (int)c, F_GLDM });
		for (auto c : NGLDMfeature::featureset) codes.push_back ({ (int)c, F_NGLDM });
		for (auto c : NGTDMFeature::featureset) codes.push_back ({ (int)c, F_NGTDM });
		std::sort (codes.begin(), codes.end());
		static const int angles[4] = { 0, 45, 90, 135 };
		for (auto [c, fam] : codes)
		{
			std::string fn = fs.findFeatureNameByCode ((Feature2D)c);
			bool angled = (fam == F_GLCM && !in_list (GLCMFeature::nonAngledFeatures, (Feature2D)c))
				|| (fam == F_GLRLM && !in_list (GLRLMFeature::nonAngledFeatures, (Feature2D)c));
			if (angled)
				for (int k = 0; k < 4; k++)
					g_cols.push_back ({ c, k, fn + "_" + std::to_string (angles[k]), fam });
			else
				g_cols.push_back ({ c, 0, fn, fam });
		}
	}

	// "glcm/greydepth" metaparameter (env_metaparams.cpp:83-91): overrides GLCM_GREYDEPTH only; INT_MIN = not set
	int g_glcm_greydepth_override = INT_MIN;

	// Same values Environment::compile_feature_settings() writes (src/nyx/env_features.cpp:712-736)
	Fsettings make_settings (int greydepth, bool ibsi, int glcm_offset)
	{
		Fsettings s ((int)NyxSetting::__COUNT__);
		for (auto& x : s) x.rval = 0;
		s[(int)NyxSetting::SOFTNAN].rval = 0.0;		// cli_result_options.h:75
		s[(int)NyxSetting::TINY].rval = 1e-10;		// cli_result_options.h:76
		s[(int)NyxSetting::SINGLEROI].bval = false;
		s[(int)NyxSetting::GREYDEPTH].ival = greydepth;
		s[(int)NyxSetting::PIXELSIZEUM].rval = 1.0;
		s[(int)NyxSetting::PIXELDISTANCE].ival = 5;
		s[(int)NyxSetting::USEGPU].bval = false;
		s[(int)NyxSetting::VERBOSLVL].ival = 0;
		s[(int)NyxSetting::IBSI].bval = ibsi;
		s[(int)NyxSetting::GLCM_OFFSET].ival = glcm_offset;
		s[(int)NyxSetting::GLCM_GREYDEPTH].ival = g_glcm_greydepth_override == INT_MIN ? greydepth : g_glcm_greydepth_override;
		s[(int)NyxSetting::GLCM_NUMANG].ival = 4;
		return s;
	}

	Dataset make_montage_dataset()
	{
		// scan_slide_props_montage leaves the slide extrema at their sentinels (src/nyx/slideprops.cpp:29-30,74-75)
		Dataset ds;
		SlideProps p;
		p.max_preroi_inten = (std::numeric_limits<double>::lowest)();
		p.min_preroi_inten = (std::numeric_limits<double>::max)();
		ds.dataset_props.push_back (p);
		return ds;
	}

	double now_s()
	{
		return std::chrono::duration<double> (std::chrono::steady_clock::now().time_since_epoch()).count();
	}

	// builds one LR from an AABB tile + membership mask the way tests/test_main_nyxus.h:50-90 does
	void make_roi_from_tile (LR& r, const uint32_t* tile, const uint8_t* mask, int w, int h)
	{
		for (int x = 0; x < w; x++)
			for (int y = 0; y < h; y++)
				if (mask[y * w + x])
				{
					if (r.aux_area == 0)
						init_label_record_3 (r, x, y, tile[y * w + x]);
					else
						update_label_record_3 (r, x, y, tile[y * w + x]);
				}
		for (int x = 0; x < w; x++)
			for (int y = 0; y < h; y++)
				if (mask[y * w + x])
					r.raw_pixels.push_back (Pixel2 (x, y, tile[y * w + x]));
		r.slide_idx = 0;
		r.make_nonanisotropic_aabb();
		r.aux_image_matrix.allocate (r.aabb.get_width(), r.aabb.get_height());
		r.aux_image_matrix.calculate_from_pixelcloud (r.raw_pixels, r.aabb);
		r.initialize_fvals();
	}
}

extern "C"
{

// what set_metaparam("glcm/greydepth=N") does for the 2D calls that follow (INT_MIN: back to the default = greydepth)
void nyxref_set_glcm_greydepth (int g) { g_glcm_greydepth_override = g; }

int nyxref_ncols()
{
	build_cols();
	return (int)g_cols.size();
}

const char* nyxref_colname (int i)
{
	build_cols();
	return g_cols[i].name.c_str();
}

int nyxref_colcode (int i)
{
	build_cols();
	return g_cols[i].code;
}

int nyxref_colfamily (int i)
{
	build_cols();
	return g_cols[i].fam;
}

// Whole-image replay.  Returns the number of ROIs (rows written, labels ascending), or -needed if cap is too small.
// t_phase[0..9] = seconds: phase1, phase2 scan, phase2 tiles, then one slot per family in dispatch order (INT,GLCM,GLRLM,GLDZM,GLSZM,GLDM,NGLDM).. [9]=NGTDM
int nyxref_featurize_2d (const uint32_t* I, const uint32_t* L, int H, int W,
	int greydepth, int ibsi, int glcm_offset, unsigned fam_mask, int nthreads, int force_finite,
	int cap, int32_t* out_labels, double* out, double* t_phase)
{
	build_cols();
	Fsettings s = make_settings (greydepth, ibsi != 0, glcm_offset);
	Dataset ds = make_montage_dataset();
	std::unordered_set<int> uniqueLabels;
	std::unordered_map<int, LR> roiData;
	uniqueLabels.reserve (100000);
	roiData.reserve (100000);

	double t0 = now_s();
	// ---- phase 1 (phase1.cpp:383-402 + pixel_feed.cpp:19-43)
	for (int col = 0; col < W; col++)
		for (int row = 0; row < H; row++)
		{
			int label = (int)L[(size_t)row * W + col];
			if (!label)
				continue;
			PixIntens inten = I[(size_t)row * W + col];
			if (uniqueLabels.find (label) == uniqueLabels.end())
			{
				uniqueLabels.insert (label);
				LR newRoi (label);
				newRoi.slide_idx = 0;
				init_label_record_3 (newRoi, col, row, inten);
				roiData[label] = newRoi;
			}
			else
			{
				LR& r = roiData[label];
				update_label_record_3 (r, col, row, inten);
			}
		}
	for (auto& kv : roiData)
	{
		kv.second.make_nonanisotropic_aabb();
		kv.second.initialize_fvals();
	}
	double t1 = now_s();

	std::vector<int> labels (uniqueLabels.begin(), uniqueLabels.end());
	std::sort (labels.begin(), labels.end());
	int n = (int)labels.size();
	if (n > cap)
		return -n;

	// ---- phase 2 scan (phase2_2d.cpp:654-681): binary search in the batch's sorted label list
	for (int col = 0; col < W; col++)
		for (int row = 0; row < H; row++)
		{
			int label = (int)L[(size_t)row * W + col];
			if (!label)
				continue;
			if (!std::binary_search (labels.begin(), labels.end(), label))
				continue;
			LR& r = roiData[label];
			r.raw_pixels.push_back (Pixel2 (col, row, I[(size_t)row * W + col]));
		}
	double t2 = now_s();
	// ---- phase 2 tiles (phase2_2d.cpp:427-465)
	for (auto lab : labels)
	{
		LR& r = roiData[lab];
		r.aux_image_matrix.allocate (r.aabb.get_width(), r.aabb.get_height());
		r.aux_image_matrix.calculate_from_pixelcloud (r.raw_pixels, r.aabb);
	}
	double t3 = now_s();
	if (t_phase)
	{
		t_phase[0] = t1 - t0;
		t_phase[1] = t2 - t1;
		t_phase[2] = t3 - t2;
	}

	// ---- reduce (reduce_trivial_rois.cpp:772-795 then :62-253), same family order
	int nThr = std::max (1, nthreads);
	size_t job = labels.size(), wpt = job / nThr;
	struct Step { unsigned fam; functype f; };
	const Step steps[] = {
		{ F_INT, PixelIntensityFeatures::reduce },
		{ F_GLCM, GLCMFeature::parallel_process_1_batch },
		{ F_GLRLM, GLRLMFeature::parallel_process_1_batch },
		{ F_GLDZM, GLDZMFeature::parallel_process_1_batch },
		{ F_GLSZM, GLSZMFeature::parallel_process_1_batch },
		{ F_GLDM, GLDMFeature::parallel_process_1_batch },
		{ F_NGLDM, NGLDMfeature::parallel_process_1_batch },
		{ F_NGTDM, NGTDMFeature::parallel_process_1_batch } };
	int si = 3;
	for (const auto& st : steps)
	{
		double ta = now_s();
		if (fam_mask & st.fam)
			runParallel (st.f, nThr, wpt, job, &labels, &roiData, s, ds);
		if (t_phase)
			t_phase[si] = now_s() - ta;
		si++;
	}

	// ---- output (output_2_buffer.cpp:441-578)
	size_t nc = g_cols.size();
	for (int i = 0; i < n; i++)
	{
		const LR& r = roiData[labels[i]];
		out_labels[i] = labels[i];
		for (size_t c = 0; c < nc; c++)
		{
			const Col& col = g_cols[c];
			double v = 0.0;
			if (fam_mask & col.fam)
			{
				const auto& vv = r.fvals[col.code];
				v = (col.k < (int)vv.size()) ? vv[col.k] : 0.0;
			}
			if (force_finite)
				v = force_finite_number (v, 0.0);
			out[(size_t)i * nc + c] = v;
		}
	}
	return n;
}

// Single-ROI operator-level call (the gtest boundary: tests/test_main_nyxus.h:50-90 then F.calculate / F.save_value).
// tile: w*h original intensities, mask: w*h membership flags.  out: nyxref_ncols() values (no force_finite).
int nyxref_roi_features (const uint32_t* tile, const uint8_t* mask, int w, int h,
	int greydepth, int ibsi, int glcm_offset, unsigned fam_mask, double* out)
{
	build_cols();
	Fsettings s = make_settings (greydepth, ibsi != 0, glcm_offset);
	Dataset ds = make_montage_dataset();
	LR r (1);
	make_roi_from_tile (r, tile, mask, w, h);
	if (r.aux_area == 0)
		return -1;
	if (fam_mask & F_INT) PixelIntensityFeatures::extract (r, s, ds);
	if (fam_mask & F_GLCM) GLCMFeature::extract (r, s);
	if (fam_mask & F_GLRLM) GLRLMFeature::extract (r, s);
	if (fam_mask & F_GLDZM) GLDZMFeature::extract (r, s);
	if (fam_mask & F_GLSZM) GLSZMFeature::extract (r, s);
	if (fam_mask & F_GLDM) GLDMFeature::extract (r, s);
	if (fam_mask & F_NGLDM) NGLDMfeature::extract (r, s);
	if (fam_mask & F_NGTDM) NGTDMFeature::extract (r, s);
	for (size_t c = 0; c < g_cols.size(); c++)
	{
		const Col& col = g_cols[c];
		const auto& vv = r.fvals[col.code];
		out[c] = ((fam_mask & col.fam) && col.k < (int)vv.size()) ? vv[col.k] : 0.0;
	}
	return 0;
}

// Sparse dump of every texture matrix of one ROI.
// trip: rows of 5 int64 {family, angle, a, b, count}:
//   GLCM  : a = grey level of matrix ROW (centre pixel), b = grey level of COLUMN (neighbour), angle 0..3
//   GLRLM : a = grey level, b = run length, angle 0..3
//   GLSZM : a = grey level, b = zone size
//   GLDZM : a = grey level, b = zone distance
//   GLDM  : a = grey level, b = dependence (1..9)
//   NGLDM : a = grey level (to_grayscale), b = matching neighbours (0..8)
//   NGTDM : a = grey level, b = 0, count = N[a]; the matching S value goes to ngtdm_s[k] in emission order
// returns the number of triplets (or -needed when cap is too small)
int64_t nyxref_roi_matrices (const uint32_t* tile, const uint8_t* mask, int w, int h,
	int greydepth, int ibsi, int glcm_offset, unsigned fam_mask,
	int64_t cap, int64_t* trip, double* ngtdm_s, int ngtdm_cap)
{
	Fsettings s = make_settings (greydepth, ibsi != 0, glcm_offset);
	LR r (1);
	make_roi_from_tile (r, tile, mask, w, h);
	if (r.aux_area == 0)
		return 0;
	std::vector<int64_t> T;
	auto emit = [&](int64_t f, int64_t ang, int64_t a, int64_t b, int64_t c) { T.insert (T.end(), { f, ang, a, b, c }); };

	if ((fam_mask & F_GLCM) && TextureFeature::bin_pixel (r.aux_min, r.aux_min, r.aux_max, greydepth) != TextureFeature::bin_pixel (r.aux_max, r.aux_min, r.aux_max, greydepth))
	{
		GLCMFeature f;
		const int dxs[4] = { glcm_offset, glcm_offset, 0, -glcm_offset }, dys[4] = { 0, glcm_offset, glcm_offset, glcm_offset };
		for (int a = 0; a < 4; a++)
		{
			f.calculateCoocMatAtAngle (f.P_matrix, s, dxs[a], dys[a], r.aux_image_matrix, r.aux_min, r.aux_max);
			int Ng = f.P_matrix.width();
			for (int row = 0; row < Ng; row++)
				for (int col = 0; col < Ng; col++)
				{
					double c = f.P_matrix.yx (row, col);
					if (c != 0)
						emit (F_GLCM, a, f.I[row], f.I[col], (int64_t)c);
				}
		}
	}
	if ((fam_mask & F_GLRLM) && r.aux_min != r.aux_max)
	{
		GLRLMFeature f;
		f.calculate (r, s);
		for (int a = 0; a < 4; a++)
		{
			auto& P = f.angles_P[a];
			for (int row = 0; row < P.height(); row++)
				for (int col = 0; col < P.width(); col++)
					if (P.yx (row, col))
						emit (F_GLRLM, a, f.I[row], col + 1, P.yx (row, col));
		}
	}
	if ((fam_mask & F_GLSZM) && r.aux_min != r.aux_max)
	{
		GLSZMFeature f;
		f.calculate (r, s);
		for (int row = 0; row < f.P.height(); row++)
			for (int col = 0; col < f.P.width(); col++)
				if (f.P.yx (row, col))
					emit (F_GLSZM, 0, f.I[row], col + 1, f.P.yx (row, col));
	}
	if ((fam_mask & F_GLDZM) && r.aux_min != r.aux_max)
	{
		GLDZMFeature f;
		SimpleMatrix<unsigned int> M;
		int Ng, Nd;
		std::vector<PixIntens> lut;
		f.prepare_GLDZM_matrix_kit (M, Ng, Nd, lut, r, greydepth, ibsi != 0);
		for (int g = 0; g < Ng; g++)
			for (int d = 0; d < Nd; d++)
				if (M.yx (g, d))
					emit (F_GLDZM, 0, lut[g], d + 1, M.yx (g, d));
	}
	if ((fam_mask & F_GLDM) && r.aux_min != r.aux_max)
	{
		GLDMFeature f;
		f.calculate (r, s);
		for (int row = 0; row < (int)f.I.size(); row++)
			for (int col = 0; col < 9; col++)
				if (f.P.yx (row, col))
					emit (F_GLDM, 0, f.I[row], col + 1, f.P.yx (row, col));
	}
	if ((fam_mask & F_NGLDM) && r.aux_min != r.aux_max)
	{
		NGLDMfeature f;
		SimpleMatrix<unsigned int> M;
		std::vector<PixIntens> lut;
		int Ng, Nr;
		f.prepare_NGLDM_matrix_kit (M, lut, Ng, Nr, r, greydepth, ibsi != 0);
		for (int g = 0; g < Ng; g++)
			for (int d = 0; d < Nr; d++)
				if (M.yx (g, d))
					emit (F_NGLDM, 0, lut[g], d, M.yx (g, d));
	}
	if (fam_mask & F_NGTDM)
	{
		NGTDMFeature f;
		f.calculate (r, s);
		int k = 0;
		for (int i = 0; i < (int)f.I.size() && i < (int)f.N.size(); i++)
			if (f.N[i])
			{
				emit (F_NGTDM, 0, f.I[i], 0, f.N[i]);
				if (ngtdm_s && k < ngtdm_cap)
					ngtdm_s[k] = f.S[i];
				k++;
			}
	}
	int64_t n = (int64_t)T.size() / 5;
	if (n > cap)
		return -n;
	std::memcpy (trip, T.data(), T.size() * sizeof (int64_t));
	return n;
}

// scalar helpers exposed for bit-exactness tests of the restatement / device code
double nyxref_fast_log10 (double x) { return Nyxus::fast_log10 (x); }
unsigned nyxref_bin_pixel (unsigned x, unsigned mn, unsigned mx, int greyinfo) { return TextureFeature::bin_pixel (x, mn, mx, greyinfo); }
unsigned nyxref_to_grayscale (unsigned i, unsigned mn, unsigned range, unsigned n) { return Nyxus::to_grayscale (i, mn, range, n, false); }

} // extern "C"

// ====================================================================================================
// 3D (Nyxus3D path): D3_VoxelIntensityFeatures, D3_GLCM_feature, D3_GLSZM_feature, D3_NGTDM_feature
// ====================================================================================================
namespace
{
	enum Fam3 { F3_INT = 1, F3_GLCM = 2, F3_NGTDM = 4, F3_GLSZM = 8 };
	struct Col3 { int code; std::string name; int fam; };
	std::vector<Col3> g_cols3;

	void build_cols3()
	{
		if (!g_cols3.empty())
			return;
		FeatureSet fs;
		std::vector<std::pair<int, int>> codes;
		for (auto c : D3_VoxelIntensityFeatures::featureset) codes.push_back ({ (int)c, F3_INT });
		for (auto c : D3_GLCM_feature::featureset) codes.push_back ({ (int)c, F3_GLCM });
		for (auto c : D3_NGTDM_feature::featureset) codes.push_back ({ (int)c, F3_NGTDM });
		for (auto c : D3_GLSZM_feature::featureset) codes.push_back ({ (int)c, F3_GLSZM });
		std::sort (codes.begin(), codes.end());
		// one column per code: save_features_2_buffer treats every 3D code as a regular feature and exports fvals[code][0]
		// (output_2_buffer.cpp:331-358 only expands Feature2D GLCM / GLRLM codes)
		for (auto [c, fam] : codes)
			g_cols3.push_back ({ c, fs.findFeatureNameByCode ((Feature3D)c), fam });
	}
}

extern "C"
{

int nyxref3d_ncols() { build_cols3(); return (int)g_cols3.size(); }
const char* nyxref3d_colname (int i) { build_cols3(); return g_cols3[i].name.c_str(); }
int nyxref3d_colcode (int i) { build_cols3(); return g_cols3[i].code; }
int nyxref3d_colfamily (int i) { build_cols3(); return g_cols3[i].fam; }

// Whole volume: replays gatherRoisMetrics_3D (phase1.cpp:248-366: voxel loop z,y,x + feed_pixel_2_metrics_3D),
// scanTrivialRois_3D (feed_pixel_2_cache_3D_LR, pixel_feed.cpp:76-95), allocateTrivialRoisBuffers_3D
// (phase2_3d.cpp:816-840) and the four reducers.  Slide extrema = min/max over the labelled voxels
// (slideprops.cpp:305-327).  I, L: [D][H][W].  out: n x nyxref3d_ncols().
int nyxref_featurize_3d (const uint32_t* I, const uint32_t* L, int D, int H, int W,
	int greydepth, int glcm_greydepth, int glszm_greydepth, int ngtdm_greydepth, int ngtdm_radius, int ibsi,
	unsigned fam_mask, int nthreads, int force_finite, int cap, int32_t* out_labels, double* out, double* t_phase)
{
	build_cols3();
	Fsettings s = make_settings (greydepth, ibsi != 0, 1);
	s[(int)NyxSetting::GLCM_GREYDEPTH].ival = glcm_greydepth;
	s[(int)NyxSetting::GLSZM_GREYDEPTH].ival = glszm_greydepth;
	s[(int)NyxSetting::NGTDM_GREYDEPTH].ival = ngtdm_greydepth;
	s[(int)NyxSetting::NGTDM_RADIUS].ival = ngtdm_radius;
	std::unordered_set<int> uniqueLabels;
	std::unordered_map<int, LR> roiData;
	double t0 = now_s();
	double smin = (std::numeric_limits<double>::max)(), smax = (std::numeric_limits<double>::lowest)();
	const size_t slice = (size_t)H * W;
	for (int z = 0; z < D; z++)
		for (int y = 0; y < H; y++)
			for (int x = 0; x < W; x++)
			{
				const size_t i = z * slice + (size_t)y * W + x;
				int label = (int)L[i];
				if (!label)
					continue;
				PixIntens inten = I[i];
				smin = std::min (smin, (double)inten);
				smax = std::max (smax, (double)inten);
				if (uniqueLabels.find (label) == uniqueLabels.end())
				{
					uniqueLabels.insert (label);
					LR newRoi (label);
					newRoi.slide_idx = 0;
					init_label_record_3D (newRoi, x, y, z, label, inten);
					roiData[label] = newRoi;
				}
				else
					update_label_record_3D (roiData[label], x, y, z, label, inten);
			}
	Dataset ds;
	{
		SlideProps p;
		p.max_preroi_inten = smax;
		p.min_preroi_inten = smin;
		ds.dataset_props.push_back (p);
	}
	for (auto& kv : roiData)
	{
		kv.second.make_nonanisotropic_aabb();
		kv.second.initialize_fvals();
	}
	std::vector<int> labels (uniqueLabels.begin(), uniqueLabels.end());
	std::sort (labels.begin(), labels.end());
	int n = (int)labels.size();
	if (n > cap)
		return -n;
	double t1 = now_s();
	for (int z = 0; z < D; z++)
		for (int y = 0; y < H; y++)
			for (int x = 0; x < W; x++)
			{
				const size_t i = z * slice + (size_t)y * W + x;
				int label = (int)L[i];
				if (!label)
					continue;
				LR& r = roiData[label];
				r.raw_pixels_3D.push_back (Pixel3 (x, y, z, I[i]));
			}
	for (auto lab : labels)
	{
		LR& r = roiData[lab];
		r.aux_image_cube.calculate_from_pixelcloud (r.raw_pixels_3D, r.aabb);
	}
	double t2 = now_s();
	if (t_phase)
	{
		t_phase[0] = t1 - t0;
		t_phase[1] = t2 - t1;
	}
	int nThr = std::max (1, nthreads);
	size_t job = labels.size(), wpt = job / nThr;
	struct Step { unsigned fam; functype f; };
	const Step steps[] = {
		{ F3_INT, D3_VoxelIntensityFeatures::reduce },
		{ F3_GLCM, D3_GLCM_feature::reduce },
		{ F3_NGTDM, D3_NGTDM_feature::reduce },
		{ F3_GLSZM, D3_GLSZM_feature::reduce } };
	int si = 2;
	for (const auto& st : steps)
	{
		double ta = now_s();
		if (fam_mask & st.fam)
			runParallel (st.f, nThr, wpt, job, &labels, &roiData, s, ds);
		if (t_phase)
			t_phase[si] = now_s() - ta;
		si++;
	}
	size_t nc = g_cols3.size();
	for (int i = 0; i < n; i++)
	{
		const LR& r = roiData[labels[i]];
		out_labels[i] = labels[i];
		for (size_t c = 0; c < nc; c++)
		{
			const Col3& col = g_cols3[c];
			double v = 0.0;
			if (fam_mask & col.fam)
			{
				const auto& vv = r.fvals[col.code];
				v = vv.empty() ? 0.0 : vv[0];
			}
			if (force_finite)
				v = force_finite_number (v, 0.0);
			out[(size_t)i * nc + c] = v;
		}
	}
	return n;
}

} // extern "C"
