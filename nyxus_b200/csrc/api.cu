// nyxus_b200 -- host side of the C ABI (include/nyxus_b200.h): context, feature selection, the per-slide pipeline
// (scan -> compact -> pack -> per-family ROI kernels -> column select -> D2H) and the test/measurement hooks.
// Mirrors, for this path only, what the reference does in featurize_montage_imp -> processMontage ->
// processIntSegImagePairInMemory (new_bindings_py.cpp:562-645, workflow_pythonapi.cpp:39-192).
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/nyxus_b200.h"
#include "kernels.cuh"

using namespace nyx;

namespace
{

struct FeatureDef { const char* name; int code; int family; int col0; int ncols; };
const FeatureDef g_features2d[] = {
#include "nyx_feature_table.inc"
};
struct FeatureDef3 { const char* name; int code; int family; int col0; };
const FeatureDef3 g_features3d_src[] = {
#include "nyx_feature_table3d.inc"
};
constexpr int NFEATURES_MAX = 256;

struct GroupDef { const char* name; int fam_mask; };
// reference featureset.cpp:650-668 (2D group names); *ALL_TEXTURE* is defined HERE as the union of the seven 2D
// texture groups (the reference has no 2D *ALL_TEXTURE*, only *3D_ALL_TEXTURE*: env_features.cpp:199-337,383-393)
const GroupDef g_groups2d[] = {
	{ "*ALL_INTENSITY*", 1 << FAM_INT },
	{ "*ALL_GLCM*", 1 << FAM_GLCM },
	{ "*ALL_GLRLM*", 1 << FAM_GLRLM },
	{ "*ALL_GLDZM*", 1 << FAM_GLDZM },
	{ "*ALL_GLSZM*", 1 << FAM_GLSZM },
	{ "*ALL_GLDM*", 1 << FAM_GLDM },
	{ "*ALL_NGLDM*", 1 << FAM_NGLDM },
	{ "*ALL_NGTDM*", 1 << FAM_NGTDM },
	{ "*ALL_TEXTURE*", (1 << FAM_GLCM) | (1 << FAM_GLRLM) | (1 << FAM_GLDZM) | (1 << FAM_GLSZM) | (1 << FAM_GLDM) | (1 << FAM_NGLDM) | (1 << FAM_NGTDM) },
};
// reference featureset.cpp:940-953 (3D group names) for the 3D families on this path
const GroupDef g_groups3d[] = {
	{ "*3D_ALL_INTENSITY*", 1 << FAM3_INT },
	{ "*3D_GLCM*", 1 << FAM3_GLCM },
	{ "*3D_GLSZM*", 1 << FAM3_GLSZM },
	{ "*3D_NGTDM*", 1 << FAM3_NGTDM },
};
// 3D groups of the reference that contain families outside this path (3D GLRLM / GLDM / GLDZM / NGLDM, morphology, moments)
const char* g_groups3d_unbuilt[] = { "*3D_ALL*", "*3D_ALL_TEXTURE*", "*3D_ALL_MORPHOLOGY*", "*3D_GLDM*", "*3D_GLDZM*", "*3D_GLRLM*",
	"*3D_NGLDM*", "*3D_ALL_NEIGHBOR*", "*3D_MOMENTS*" };

const char* g_stage_names[NYXB200_NSTAGES] = { "h2d", "init", "scan", "compact", "pack", "intensity", "glcm", "glrlm", "gldzm",
	"glszm", "gldm", "ngldm", "ngtdm", "select", "d2h", "total" };
enum Stage { ST_H2D = 0, ST_INIT, ST_SCAN, ST_COMPACT, ST_PACK, ST_INT, ST_GLCM, ST_GLRLM, ST_GLDZM, ST_GLSZM, ST_GLDM, ST_NGLDM,
	ST_NGTDM, ST_SELECT, ST_D2H, ST_TOTAL };

std::string g_create_error;

template <class T> struct DevBuf
{
	T* p = nullptr;
	size_t cap = 0;
	cudaError_t reserve (size_t n)
	{
		if (n <= cap)
			return cudaSuccess;
		if (p)
			cudaFree (p);
		p = nullptr;
		cap = 0;
		size_t want = n + n / 4 + 16;
		cudaError_t e = cudaMalloc ((void**)&p, want * sizeof (T));
		if (e == cudaSuccess)
			cap = want;
		return e;
	}
	void release()
	{
		if (p)
			cudaFree (p);
		p = nullptr;
		cap = 0;
	}
};

constexpr size_t SMEM_LIMIT_WORDS = (227 * 1024) / 4;

} // namespace

struct nyxb200_ctx
{
	nyxb200_settings st;
	std::string err;
	int sm_count = 148;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev[NYXB200_NSTAGES + 2];
	bool ev_used[NYXB200_NSTAGES + 2];

	// selection
	std::vector<FeatureDef> feats;	// the context's feature table (2D or 3D codes)
	bool feat_on[NFEATURES_MAX];
	int fam_mask = 0;
	std::vector<int> sel_cols;		// canonical column of every output column
	std::vector<std::string> sel_names;
	std::vector<int> sel_codes;
	DevBuf<int> d_sel;
	bool sel_dirty = true;

	// device buffers (grow-only)
	DevBuf<uint8_t> d_inten;
	DevBuf<uint32_t> d_label;
	DevBuf<LabelStat> d_stats;
	DevBuf<uint2> d_ztab;
	uint32_t stats_cap = 1u << 20;
	DevBuf<SlideInfo> d_info;
	DevBuf<uint32_t> d_blkcnt;
	DevBuf<unsigned long long> d_blkarea;
	DevBuf<RoiRec> d_rois;
	DevBuf<uint32_t> d_classslot, d_tlist, d_alist, d_bases;
	DevBuf<uint32_t> d_cloud;
	DevBuf<double> d_full, d_sel_out;
	DevBuf<uint32_t> d_scratch;
	DevBuf<unsigned long long> d_rowoff;
	DevBuf<uint32_t> d_dbg;
	DevBuf<double> d_dbgf;

	// state of the last processed image (for the debug hooks)
	ImageView last_im {};
	SlideInfo last_info {};
	uint32_t tbase[NCLASS + 2] {}, abase[NACLASS + 2] {};
	std::vector<RoiRec> h_rois;
	bool h_rois_valid = false;

	double* h_out = nullptr;		// pinned result buffer (grow-only)
	size_t h_out_cap = 0, h_out_used = 0;

	int glszm3_greydepth = 0, ngtdm3_greydepth = 0, ngtdm3_radius = 0;	// env_features.cpp:712-736: zero unless set

	float stage_ms[NYXB200_NSTAGES] {};
	int64_t counts[4] {};
	std::vector<int32_t> row_image;
};

namespace
{

int fail (nyxb200_ctx* c, int code, const char* fmt, ...)
{
	char buf[1024];
	va_list ap;
	va_start (ap, fmt);
	vsnprintf (buf, sizeof (buf), fmt, ap);
	va_end (ap);
	if (c)
		c->err = buf;
	else
		g_create_error = buf;
	return code;
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail (ctx, NYXB200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString (e_), __FILE__, __LINE__); } while (0)

void rebuild_selection (nyxb200_ctx* ctx)
{
	ctx->sel_cols.clear();
	ctx->sel_names.clear();
	ctx->sel_codes.clear();
	ctx->fam_mask = 0;
	static const int angles[4] = { 0, 45, 90, 135 };
	// enabled features in ascending code order (reference featureset.cpp getEnabledFeatures; the table is already sorted)
	for (int i = 0; i < (int)ctx->feats.size(); i++)
	{
		if (!ctx->feat_on[i])
			continue;
		const FeatureDef& f = ctx->feats[i];
		ctx->fam_mask |= 1 << f.family;
		for (int k = 0; k < f.ncols; k++)
		{
			ctx->sel_cols.push_back (f.col0 + k);
			ctx->sel_codes.push_back (f.code);
			ctx->sel_names.push_back (f.ncols == 4 ? std::string (f.name) + "_" + std::to_string (angles[k]) : std::string (f.name));
		}
	}
	ctx->sel_dirty = true;
}

struct Timer
{
	nyxb200_ctx* ctx;
	void begin (int st)
	{
		cudaEventRecord (ctx->ev[st], ctx->stream);
	}
};

} // namespace

extern "C"
{

const char* nyxb200_version (void) { return "nyxus_b200 0.1.0 (sm_100a)"; }
const char* nyxb200_stage_name (int i) { return (i >= 0 && i < NYXB200_NSTAGES) ? g_stage_names[i] : ""; }

int nyxb200_family_id (const char* name)
{
	static const char* names[FAM_COUNT] = { "INT", "GLCM", "GLRLM", "GLDZM", "GLSZM", "GLDM", "NGLDM", "NGTDM" };
	for (int i = 0; i < FAM_COUNT; i++)
		if (!strcmp (names[i], name))
			return i;
	return -1;
}

void nyxb200_default_settings (nyxb200_settings* s)
{
	s->device = 0;
	s->coarse_gray_depth = 64;	// nyxus.py:181, environment.cpp:106
	s->ibsi = 0;
	s->glcm_offset = 1;		// env_features.cpp:727
	s->glcm_symmetric = 0;		// glcm.cpp:8
	s->verbose = 0;
	s->soft_nan = 0.0;		// cli_result_options.h:75
	s->dim = 2;
	s->reserved = 0;
}

const char* nyxb200_last_error (const nyxb200_ctx* ctx)
{
	return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

int nyxb200_create (const nyxb200_settings* s, nyxb200_ctx** out)
{
	if (!s || !out)
		return fail (nullptr, NYXB200_EINVAL, "null argument");
	*out = nullptr;
	if (s->coarse_gray_depth <= 0)	// nyxus.py:206-207
		return fail (nullptr, NYXB200_EINVAL, "Custom number of grayscale levels (parameter coarse_gray_depth, default=64) must be non-negative.");
	if (s->coarse_gray_depth > 65535)
		return fail (nullptr, NYXB200_EUNSUPPORTED, "coarse_gray_depth > 65535 is not supported");
	if (s->dim != 2 && s->dim != 3)
		return fail (nullptr, NYXB200_EINVAL, "dim must be 2 or 3");
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount (&ndev);
	if (e != cudaSuccess || ndev == 0)
		return fail (nullptr, NYXB200_ECUDA, "no CUDA device available (%s); this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString (e) : "device count 0");
	if (s->device < 0 || s->device >= ndev)
		return fail (nullptr, NYXB200_EINVAL, "device %d out of range (have %d)", s->device, ndev);
	cudaDeviceProp prop;
	if ((e = cudaGetDeviceProperties (&prop, s->device)) != cudaSuccess)
		return fail (nullptr, NYXB200_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString (e));
	if (prop.major != 10)
		return fail (nullptr, NYXB200_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", s->device, prop.major, prop.minor);
	if ((e = cudaSetDevice (s->device)) != cudaSuccess)
		return fail (nullptr, NYXB200_ECUDA, "cudaSetDevice: %s", cudaGetErrorString (e));
	nyxb200_ctx* ctx = new nyxb200_ctx();
	ctx->st = *s;
	ctx->sm_count = prop.multiProcessorCount;
	if (s->dim == 3)
		for (const auto& f : g_features3d_src)
			ctx->feats.push_back ({ f.name, f.code, f.family, f.col0, 1 });
	else
		for (const auto& f : g_features2d)
			ctx->feats.push_back (f);
	for (int i = 0; i < NFEATURES_MAX; i++)
		ctx->feat_on[i] = false;
	if (cudaStreamCreateWithFlags (&ctx->stream, cudaStreamNonBlocking) != cudaSuccess)
	{
		delete ctx;
		return fail (nullptr, NYXB200_ECUDA, "cudaStreamCreate failed");
	}
	for (auto& ev : ctx->ev)
		cudaEventCreate (&ev);
	rebuild_selection (ctx);
	*out = ctx;
	return NYXB200_OK;
}

void nyxb200_destroy (nyxb200_ctx* ctx)
{
	if (!ctx)
		return;
	cudaSetDevice (ctx->st.device);
	cudaStreamSynchronize (ctx->stream);
	ctx->d_ztab.release(); ctx->d_sel.release(); ctx->d_inten.release(); ctx->d_label.release(); ctx->d_stats.release(); ctx->d_info.release();
	ctx->d_blkcnt.release(); ctx->d_blkarea.release(); ctx->d_rois.release(); ctx->d_classslot.release(); ctx->d_tlist.release();
	ctx->d_alist.release(); ctx->d_bases.release(); ctx->d_cloud.release(); ctx->d_full.release(); ctx->d_sel_out.release();
	ctx->d_scratch.release(); ctx->d_rowoff.release(); ctx->d_dbg.release(); ctx->d_dbgf.release();
	if (ctx->h_out)
		cudaFreeHost (ctx->h_out);
	for (auto& ev : ctx->ev)
		cudaEventDestroy (ev);
	cudaStreamDestroy (ctx->stream);
	delete ctx;
}

int nyxb200_select_features (nyxb200_ctx* ctx, const char* const* names, int n)
{
	if (!ctx || (!names && n))
		return fail (ctx, NYXB200_EINVAL, "null argument");
	bool on[NFEATURES_MAX];
	const int nf = (int)ctx->feats.size();
	const bool d3 = ctx->st.dim == 3;
	for (int i = 0; i < NFEATURES_MAX; i++)
		on[i] = false;	// env_features.cpp:459 "initially, no features are enabled"
	for (int k = 0; k < n; k++)
	{
		const char* s = names[k];
		bool found = false;
		const GroupDef* groups = d3 ? g_groups3d : g_groups2d;
		const int ngroups = d3 ? (int)(sizeof (g_groups3d) / sizeof (g_groups3d[0])) : (int)(sizeof (g_groups2d) / sizeof (g_groups2d[0]));
		for (int gi = 0; gi < ngroups && !found; gi++)
			if (!strcmp (groups[gi].name, s))
			{
				for (int i = 0; i < nf; i++)
					if (groups[gi].fam_mask & (1 << ctx->feats[i].family))
						on[i] = true;
				found = true;
			}
		if (found)
			continue;
		if (d3)
			for (const char* g : g_groups3d_unbuilt)
				if (!strcmp (g, s))
					return fail (ctx, NYXB200_EUNSUPPORTED, "feature group '%s' contains 3D families outside this path "
						"(built: *3D_ALL_INTENSITY*, *3D_GLCM*, *3D_GLSZM*, *3D_NGTDM*)", s);
		for (int i = 0; i < nf; i++)
			if (!strcmp (ctx->feats[i].name, s))
			{
				on[i] = true;
				found = true;
				break;
			}
		if (!found)	// env_features.cpp:496,507
			return fail (ctx, NYXB200_EINVAL, "Error: '%s' is not a valid %s feature name ", s, d3 ? "3D" : "2D");
	}
	memcpy (ctx->feat_on, on, sizeof (on));
	rebuild_selection (ctx);
	return NYXB200_OK;
}

int nyxb200_num_columns (const nyxb200_ctx* ctx) { return ctx ? (int)ctx->sel_cols.size() : 0; }
const char* nyxb200_column_name (const nyxb200_ctx* ctx, int i)
{
	return (ctx && i >= 0 && i < (int)ctx->sel_names.size()) ? ctx->sel_names[i].c_str() : "";
}
int nyxb200_column_code (const nyxb200_ctx* ctx, int i)
{
	return (ctx && i >= 0 && i < (int)ctx->sel_codes.size()) ? ctx->sel_codes[i] : -1;
}

int nyxb200_set_param (nyxb200_ctx* ctx, const char* key, double value)
{
	if (!ctx || !key)
		return fail (ctx, NYXB200_EINVAL, "null argument");
	int iv = (int)value;
	if (!strcmp (key, "glcm/offset"))		// env_metaparams.cpp (glcm/offset)
	{
		if (iv < 0)
			return fail (ctx, NYXB200_EINVAL, "glcm/offset must be >= 0");
		ctx->st.glcm_offset = iv;
	}
	else if (!strcmp (key, "glcm/greydepth") || !strcmp (key, "coarse_gray_depth"))
	{
		if (iv <= 0 || iv > 65535)
			return fail (ctx, NYXB200_EINVAL, "grey depth must be in 1..65535");
		ctx->st.coarse_gray_depth = iv;
	}
	else if (!strcmp (key, "ibsi"))
		ctx->st.ibsi = iv != 0;
	else if (!strcmp (key, "glcm/symmetric"))
		ctx->st.glcm_symmetric = iv != 0;
	else if (!strcmp (key, "soft_nan"))
		ctx->st.soft_nan = value;
	else if (!strcmp (key, "3glcm/greydepth"))		// env_metaparams.cpp:109-120
	{
		if (iv <= 0 || iv > 65535)
			return fail (ctx, NYXB200_EINVAL, "grey depth must be in 1..65535");
		ctx->st.coarse_gray_depth = iv;
	}
	else if (!strcmp (key, "3glszm/greydepth"))		// env_metaparams.cpp:222-233
	{
		if (iv < 0 || iv > 65535)
			return fail (ctx, NYXB200_EINVAL, "grey depth must be in 0..65535");
		ctx->glszm3_greydepth = iv;
	}
	else if (!strcmp (key, "3ngtdm/greydepth"))		// env_metaparams.cpp:175-186
	{
		if (iv < 0 || iv > 65535)
			return fail (ctx, NYXB200_EINVAL, "grey depth must be in 0..65535");
		ctx->ngtdm3_greydepth = iv;
	}
	else if (!strcmp (key, "3ngtdm/radius"))		// env_metaparams.cpp:187-197
	{
		if (iv < 0)
			return fail (ctx, NYXB200_EINVAL, "3ngtdm/radius: expecting a positive integer");
		if (iv > 1)
			return fail (ctx, NYXB200_EUNSUPPORTED, "3ngtdm/radius > 1 is not built (26-neighbourhood only)");
		ctx->ngtdm3_radius = iv;
	}
	else
		return fail (ctx, NYXB200_EINVAL, "unknown parameter '%s'", key);
	return NYXB200_OK;
}

} // extern "C"

// ====================================================================================================
// pipeline
// ====================================================================================================
namespace
{

template <class K> int set_smem (nyxb200_ctx* ctx, K kernel, size_t bytes)
{
	if (bytes > 48 * 1024)
		CU (cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
	return 0;
}

// words of family matrices for a class
size_t family_words (int fam, bool p16, int Lmax, uint32_t maxcells, int maxwh)
{
	switch (fam)
	{
	case FAM_GLCM: return p16 ? glcm_smem_words<true> (Lmax, 4) : glcm_smem_words<false> (Lmax, 4);
	case FAM_GLRLM: return glrlm_smem_words (Lmax, maxwh, (int)maxcells);
	case FAM_GLDZM: return gldzm_smem_words (Lmax, gldzm_nd_bound (maxcells, maxwh), maxcells);
	case FAM_GLSZM: return glszm_smem_words (maxcells, glszm_hash_slots (Lmax, maxcells));
	case FAM_GLDM: return gldm_smem_words (Lmax);
	case FAM_NGLDM: return ngldm_smem_words (Lmax);
	case FAM_NGTDM: return ngtdm_smem_words (Lmax);
	}
	return 0;
}

constexpr size_t WARP_SLICE_LIMIT_WORDS = (28 * 1024) / 4;	// per-warp shared memory above which the CTA-per-ROI form is used

template <int FAM, bool P16>
int launch_family_t (nyxb200_ctx* ctx, const ImageView& im, const uint32_t* list, uint32_t n, const Params& prm, int Lmax,
	uint32_t maxcells, int maxwh, double* out, DebugOut dbg)
{
	if (!n)
		return 0;
	const size_t tilew = (maxcells + 1) & ~(size_t)1;
	const size_t famw = (family_words (FAM, P16, Lmax, maxcells, maxwh) + 1) & ~(size_t)1;
	// small tiles of the single-matrix families: one warp per ROI
	if (FAM != FAM_GLCM && FAM != FAM_GLRLM && !dbg.u32 && maxcells <= 1024 && tilew + famw <= WARP_SLICE_LIMIT_WORDS)
	{
		const size_t bytes = (tilew + famw) * 4 * ROI_WARPS;
		auto kw = k_roi_family_w<FAM>;
		if (set_smem (ctx, kw, bytes))
			return NYXB200_ECUDA;
		const int per_sm = (int)std::max<size_t> (1, std::min<size_t> (16, (size_t)(227 * 1024) / (bytes + 1024)));
		const int grid = (int)std::min<uint32_t> ((n + ROI_WARPS - 1) / ROI_WARPS, (uint32_t)(ctx->sm_count * per_sm));
		kw<<<grid, ROI_THREADS, bytes, ctx->stream>>> (im, ctx->d_rois.p, list, n, prm, Lmax, maxcells, maxwh, out, NCOL_ALL, famw);
		CU (cudaGetLastError());
		ctx->counts[0]++;
		return 0;
	}
	size_t smem_words = RED_WORDS + tilew + famw;
	uint32_t* gs = nullptr;
	int grid = (int)std::min<uint32_t> (n, (uint32_t)ctx->sm_count * 16);
	if (smem_words > SMEM_LIMIT_WORDS)
	{
		// matrices go to a per-CTA slot in global memory (L2 resident), the tile stays in shared memory
		smem_words = RED_WORDS + tilew;
		grid = (int)std::min<uint32_t> (n, (uint32_t)ctx->sm_count * 2);
		CU (ctx->d_scratch.reserve ((size_t)grid * famw));
		gs = ctx->d_scratch.p;
		CU (cudaMemsetAsync (gs, 0, (size_t)grid * famw * 4, ctx->stream));	// bodies expect (and leave) an all-zero region
	}
	auto kern = k_roi_family<FAM, P16>;
	if (set_smem (ctx, kern, smem_words * 4))
		return NYXB200_ECUDA;
	kern<<<grid, ROI_THREADS, smem_words * 4, ctx->stream>>> (im, ctx->d_rois.p, list, n, prm, Lmax, maxcells, maxwh, out,
		NCOL_ALL, gs, famw, dbg);
	CU (cudaGetLastError());
	ctx->counts[0]++;
	return 0;
}

int launch_family (nyxb200_ctx* ctx, int fam, const ImageView& im, const uint32_t* list, uint32_t n, const Params& prm, int Lmax,
	uint32_t maxcells, int maxwh, double* out, DebugOut dbg)
{
	switch (fam)
	{
	case FAM_GLCM:
		// 16-bit packed joint counters: a cell count is at most 2 * ncell (symmetric)
		if (maxcells < 32768)
			return launch_family_t<FAM_GLCM, true> (ctx, im, list, n, prm, Lmax, maxcells, maxwh, out, dbg);
		return launch_family_t<FAM_GLCM, false> (ctx, im, list, n, prm, Lmax, maxcells, maxwh, out, dbg);
	case FAM_GLRLM: return launch_family_t<FAM_GLRLM, false> (ctx, im, list, n, prm, Lmax, maxcells, maxwh, out, dbg);
	case FAM_GLDZM: return launch_family_t<FAM_GLDZM, false> (ctx, im, list, n, prm, Lmax, maxcells, maxwh, out, dbg);
	case FAM_GLSZM: return launch_family_t<FAM_GLSZM, false> (ctx, im, list, n, prm, Lmax, maxcells, maxwh, out, dbg);
	case FAM_GLDM: return launch_family_t<FAM_GLDM, false> (ctx, im, list, n, prm, Lmax, maxcells, maxwh, out, dbg);
	case FAM_NGLDM: return launch_family_t<FAM_NGLDM, false> (ctx, im, list, n, prm, Lmax, maxcells, maxwh, out, dbg);
	case FAM_NGTDM: return launch_family_t<FAM_NGTDM, false> (ctx, im, list, n, prm, Lmax, maxcells, maxwh, out, dbg);
	}
	return 0;
}

int launch_intensity (nyxb200_ctx* ctx, const uint32_t* list, uint32_t n, const Params& prm, uint32_t maxarea, double* out, int ostride)
{
	if (!n)
		return 0;
	const uint32_t npad = next_pow2 (maxarea);
	const size_t memw = (intensity_smem_words (npad, prm.greydepth) + 1) & ~(size_t)1;
	if (npad <= 1024 && memw <= WARP_SLICE_LIMIT_WORDS)	// small clouds: one warp per ROI
	{
		const size_t bytes = memw * 4 * ROI_WARPS;
		if (set_smem (ctx, k_roi_intensity_w, bytes))
			return NYXB200_ECUDA;
		const int per_sm = (int)std::max<size_t> (1, std::min<size_t> (16, (size_t)(227 * 1024) / (bytes + 1024)));
		const int grid = (int)std::min<uint32_t> ((n + ROI_WARPS - 1) / ROI_WARPS, (uint32_t)(ctx->sm_count * per_sm));
		k_roi_intensity_w<<<grid, ROI_THREADS, bytes, ctx->stream>>> (ctx->d_rois.p, list, n, prm, ctx->d_cloud.p, out, ostride, memw);
		CU (cudaGetLastError());
		ctx->counts[0]++;
		return 0;
	}
	size_t smem_words = RED_WORDS + memw;
	uint32_t* gs = nullptr;
	int grid = (int)std::min<uint32_t> (n, (uint32_t)ctx->sm_count * 16);
	if (smem_words > SMEM_LIMIT_WORDS)
	{
		smem_words = RED_WORDS;
		grid = (int)std::min<uint32_t> (n, (uint32_t)ctx->sm_count * 2);
		CU (ctx->d_scratch.reserve ((size_t)grid * memw));
		gs = ctx->d_scratch.p;
	}
	if (set_smem (ctx, k_roi_intensity, smem_words * 4))
		return NYXB200_ECUDA;
	k_roi_intensity<<<grid, ROI_THREADS, smem_words * 4, ctx->stream>>> (ctx->d_rois.p, list, n, prm, ctx->d_cloud.p, npad, out,
		ostride, gs, memw);
	CU (cudaGetLastError());
	ctx->counts[0]++;
	return 0;
}

// ---- 3D families: CTA per ROI on the AABB cube
template <int FAM3, bool P16>
int launch_family3_t (nyxb200_ctx* ctx, const ImageView& im, const uint32_t* list, uint32_t n, const Params& prm, int greyinfo,
	int Lmax, uint32_t maxcells, double* out)
{
	if (!n)
		return 0;
	const size_t tilew = (maxcells + 1) & ~(size_t)1;
	size_t famw = 0;
	if (FAM3 == FAM3_GLCM) famw = P16 ? glcm_smem_words<true> (Lmax, 4) : glcm_smem_words<false> (Lmax, 4);
	if (FAM3 == FAM3_GLSZM) famw = glszm3d_smem_words (maxcells, glszm_hash_slots (Lmax, maxcells));
	if (FAM3 == FAM3_NGTDM) famw = ngtdm3d_smem_words (Lmax);
	famw = (famw + 1) & ~(size_t)1;
	size_t smem_words = RED_WORDS + tilew + famw;
	uint32_t* gs = nullptr;
	int grid = (int)std::min<uint32_t> (n, (uint32_t)ctx->sm_count * 8);
	if (smem_words > SMEM_LIMIT_WORDS)
	{
		smem_words = RED_WORDS + tilew;
		grid = (int)std::min<uint32_t> (n, (uint32_t)ctx->sm_count * 2);
		CU (ctx->d_scratch.reserve ((size_t)grid * famw));
		gs = ctx->d_scratch.p;
		CU (cudaMemsetAsync (gs, 0, (size_t)grid * famw * 4, ctx->stream));
	}
	auto kern = k_roi_family3<FAM3, P16>;
	if (set_smem (ctx, kern, smem_words * 4))
		return NYXB200_ECUDA;
	kern<<<grid, ROI_THREADS, smem_words * 4, ctx->stream>>> (im, ctx->d_rois.p, list, n, prm, greyinfo, Lmax, maxcells, out,
		NCOL3_ALL, gs, famw);
	CU (cudaGetLastError());
	ctx->counts[0]++;
	return 0;
}

Params make_params (const nyxb200_ctx* ctx)
{
	Params p;
	p.greydepth = ctx->st.coarse_gray_depth;
	p.ibsi = ctx->st.ibsi;
	p.glcm_offset = ctx->st.glcm_offset;
	p.glcm_symmetric = ctx->st.glcm_symmetric;
	p.softnan = ctx->st.soft_nan;
	p.glszm3_greydepth = ctx->glszm3_greydepth;
	p.ngtdm3_greydepth = ctx->ngtdm3_greydepth;
	p.ngtdm3_radius = ctx->ngtdm3_radius;
	return p;
}

void stage_begin (nyxb200_ctx* ctx, int st)
{
	cudaEventRecord (ctx->ev[st], ctx->stream);
	ctx->ev_used[st] = true;
}

// grow-only pinned host buffer for the result table (keeps its content when it grows)
int reserve_pinned_out (nyxb200_ctx* ctx, size_t n)
{
	if (n <= ctx->h_out_cap)
		return 0;
	size_t want = n + n / 2 + 1024;
	double* p = nullptr;
	CU (cudaMallocHost ((void**)&p, want * sizeof (double)));
	if (ctx->h_out)
	{
		memcpy (p, ctx->h_out, ctx->h_out_used * sizeof (double));
		cudaFreeHost (ctx->h_out);
	}
	ctx->h_out = p;
	ctx->h_out_cap = want;
	return 0;
}

// One slide already resident on the device.  Appends its rows to the context's pinned table buffer.
int process_slide (nyxb200_ctx* ctx, const ImageView& im, int img_index, int64_t* nrows_out)
{
	const unsigned long long npix = (unsigned long long)im.D * (unsigned long long)im.H * (unsigned long long)im.W;
	const bool d3 = ctx->st.dim == 3;
	const int ncol_all = d3 ? NCOL3_ALL : NCOL_ALL;
	Params prm = make_params (ctx);
	cudaStream_t s = ctx->stream;
	CU (ctx->d_info.reserve (1));
	SlideInfo info;

	// ---- scan (retry once with a larger table when a label exceeds the table)
	for (int attempt = 0; ; attempt++)
	{
		CU (ctx->d_stats.reserve (ctx->stats_cap));
		uint2* ztab = nullptr;
		if (d3)
		{
			CU (ctx->d_ztab.reserve (ctx->stats_cap));
			ztab = ctx->d_ztab.p;
		}
		stage_begin (ctx, ST_INIT);
		k_init_stats<<<(ctx->stats_cap + 255) / 256, 256, 0, s>>> (ctx->d_stats.p, ctx->stats_cap, ctx->d_info.p, ztab);
		CU (cudaGetLastError());
		stage_begin (ctx, ST_SCAN);
		const bool vec = (im.W % 4 == 0) && ((uintptr_t)im.label % 16 == 0) && ((uintptr_t)im.inten % (im.inten_u16 ? 8 : 16) == 0);
		const uint32_t nrows = (uint32_t)im.D * (uint32_t)im.H, nchunk = ((uint32_t)im.W + 127) / 128;
		const unsigned long long nitems = (unsigned long long)nrows * nchunk;
		const int grid = (int)std::max<unsigned long long> (1, std::min<unsigned long long> ((nitems + 7) / 8, (unsigned long long)ctx->sm_count * 8));
#define NYX_SCAN(U16, VEC, D3) k_scan_stats<U16, VEC, D3><<<grid, 256, 0, s>>> (im.inten, im.label, (uint32_t)im.W, nrows, (uint32_t)im.H, nchunk, \
			ctx->d_stats.p, ctx->stats_cap, ctx->d_info.p, ztab)
		if (d3)
		{
			if (im.inten_u16) { if (vec) NYX_SCAN (true, true, true); else NYX_SCAN (true, false, true); }
			else { if (vec) NYX_SCAN (false, true, true); else NYX_SCAN (false, false, true); }
		}
		else
		{
			if (im.inten_u16) { if (vec) NYX_SCAN (true, true, false); else NYX_SCAN (true, false, false); }
			else { if (vec) NYX_SCAN (false, true, false); else NYX_SCAN (false, false, false); }
		}
#undef NYX_SCAN
		CU (cudaGetLastError());
		ctx->counts[0] += 2;
		stage_begin (ctx, ST_COMPACT);
		const uint32_t nblk = (ctx->stats_cap + CB - 1) / CB;
		CU (ctx->d_blkcnt.reserve (nblk));
		CU (ctx->d_blkarea.reserve (nblk));
		k_compact_count<<<nblk, CB, 0, s>>> (ctx->d_stats.p, ctx->stats_cap, ctx->d_info.p, ctx->d_blkcnt.p, ctx->d_blkarea.p);
		k_compact_scan<<<1, CB, 0, s>>> (ctx->d_blkcnt.p, ctx->d_blkarea.p, nblk, ctx->d_info.p);
		CU (cudaGetLastError());
		ctx->counts[0] += 2;
		CU (cudaMemcpyAsync (&info, ctx->d_info.p, sizeof (info), cudaMemcpyDeviceToHost, s));
		CU (cudaStreamSynchronize (s));
		if (info.max_label >= ctx->stats_cap)
		{
			if (attempt > 0)
				return fail (ctx, NYXB200_ECUDA, "label table overflow persisted");
			if (info.max_label >= (1u << 27))
				return fail (ctx, NYXB200_EUNSUPPORTED, "label value %u exceeds the dense label table limit (2^27); relabel the mask", info.max_label);
			ctx->stats_cap = info.max_label + 1;
			continue;
		}
		break;
	}
	const uint32_t nroi = info.n_roi;
	*nrows_out = nroi;
	ctx->counts[1] += nroi;
	ctx->counts[2] += (int64_t)info.n_fg;
	ctx->counts[3] += (int64_t)npix;
	ctx->last_im = im;
	ctx->h_rois_valid = false;
	if (nroi == 0)
	{
		ctx->last_info = info;
		return 0;
	}
	if (ctx->st.ibsi && info.max_inten > CELL_LEVEL)
		return fail (ctx, NYXB200_EUNSUPPORTED, "IBSI mode needs intensities < 2^24 (max %u)", info.max_inten);

	// ---- ROI records + class lists
	CU (ctx->d_rois.reserve (nroi));
	CU (ctx->d_classslot.reserve ((size_t)nroi * 2));
	CU (ctx->d_tlist.reserve (nroi));
	CU (ctx->d_alist.reserve (nroi));
	CU (ctx->d_bases.reserve (NCLASS + NACLASS + 4));
	{
		const uint32_t nblk = (ctx->stats_cap + CB - 1) / CB;
		k_compact_write<<<nblk, CB, 0, s>>> (ctx->d_stats.p, ctx->stats_cap, ctx->d_info.p, ctx->d_blkcnt.p, ctx->d_blkarea.p,
			ctx->d_rois.p, ctx->d_classslot.p, d3 ? ctx->d_ztab.p : nullptr);
		CU (cudaGetLastError());
		ctx->counts[0]++;
		CU (cudaMemcpyAsync (&info, ctx->d_info.p, sizeof (info), cudaMemcpyDeviceToHost, s));
		CU (cudaStreamSynchronize (s));
		uint32_t bases[NCLASS + NACLASS + 4];
		uint32_t acc = 0;
		for (int c = 0; c <= NCLASS; c++)
		{
			ctx->tbase[c] = bases[c] = acc;
			acc += info.class_count[c];
		}
		ctx->tbase[NCLASS + 1] = acc;
		acc = 0;
		for (int c = 0; c <= NACLASS; c++)
		{
			ctx->abase[c] = bases[NCLASS + 1 + c] = acc;
			acc += info.aclass_count[c];
		}
		ctx->abase[NACLASS + 1] = acc;
		CU (cudaMemcpyAsync (ctx->d_bases.p, bases, sizeof (uint32_t) * (NCLASS + NACLASS + 2), cudaMemcpyHostToDevice, s));
		k_fill_class_lists<<<(nroi + 255) / 256, 256, 0, s>>> (ctx->d_classslot.p, nroi, ctx->d_bases.p, ctx->d_bases.p + NCLASS + 1,
			ctx->d_tlist.p, ctx->d_alist.p);
		CU (cudaGetLastError());
		ctx->counts[0]++;
	}
	ctx->last_info = info;

	// ---- canonical table
	CU (ctx->d_full.reserve ((size_t)nroi * ncol_all));
	CU (cudaMemsetAsync (ctx->d_full.p, 0, (size_t)nroi * ncol_all * sizeof (double), s));
	if (d3)	// 3d_intensity.cpp:56-61: ROI range over the scanned range of the labelled voxels (slideprops.cpp:305-327)
		prm.covered_den = (double)info.max_inten - (double)(~info.inv_min_inten);

	// ---- CSR cloud (first-order only)
	stage_begin (ctx, ST_PACK);
	if (ctx->fam_mask & (1 << FAM_INT))
	{
		CU (ctx->d_cloud.reserve ((size_t)info.n_fg + 4));
		const uint32_t nsmall = ctx->tbase[NCLASS];
		if (nsmall)
		{
			k_pack_cloud<<<(nsmall * 32 + 255) / 256, 256, 0, s>>> (im, ctx->d_rois.p, ctx->d_tlist.p, nsmall, ctx->d_cloud.p);
			CU (cudaGetLastError());
			ctx->counts[0]++;
		}
		const uint32_t nlarge = info.class_count[NCLASS];
		if (nlarge)
		{
			const int grid = (int)std::min<uint32_t> (nlarge, (uint32_t)ctx->sm_count * 4);
			const size_t stride = (size_t)im.H * (size_t)im.D + 2;
			CU (ctx->d_rowoff.reserve ((size_t)grid * stride));
			k_pack_cloud_large<<<grid, 256, 0, s>>> (im, ctx->d_rois.p, ctx->d_tlist.p + ctx->tbase[NCLASS], nlarge, ctx->d_cloud.p,
				ctx->d_rowoff.p, stride);
			CU (cudaGetLastError());
			ctx->counts[0]++;
		}
	}

	// ---- families
	const int Lmax = ctx->st.ibsi ? (int)std::max<uint32_t> (info.max_inten, 1u) : ctx->st.coarse_gray_depth;
	DebugOut nodbg { nullptr, nullptr, nullptr };
	stage_begin (ctx, ST_INT);
	if (ctx->fam_mask & (1 << FAM_INT))
	{
		for (int c = 0; c < NACLASS; c++)
			if (int rc = launch_intensity (ctx, ctx->d_alist.p + ctx->abase[c], info.aclass_count[c], prm, info.aclass_maxarea[c], ctx->d_full.p, ncol_all))
				return rc;
		if (info.aclass_count[NACLASS])
			if (int rc = launch_intensity (ctx, ctx->d_alist.p + ctx->abase[NACLASS], info.aclass_count[NACLASS], prm, info.aclass_maxarea[NACLASS], ctx->d_full.p, ncol_all))
				return rc;
	}
	static const int fam_stage[FAM_COUNT] = { ST_INT, ST_GLCM, ST_GLRLM, ST_GLDZM, ST_GLSZM, ST_GLDM, ST_NGLDM, ST_NGTDM };
	if (d3)
	{
		if ((ctx->fam_mask & ~(1 << FAM3_INT)) && info.class_count[NCLASS])
			return fail (ctx, NYXB200_EUNSUPPORTED, "a 3D ROI whose bounding cube exceeds 40000 voxels needs the large-ROI texture path (not in this build)");
		const int maxI = (int)std::max<uint32_t> (info.max_inten, 1u);
		for (int fam = FAM3_GLCM; fam < FAM3_COUNT; fam++)
		{
			static const int st3[FAM3_COUNT] = { ST_INT, ST_GLCM, ST_GLSZM, ST_NGTDM };
			stage_begin (ctx, st3[fam]);
			if (!(ctx->fam_mask & (1 << fam)))
				continue;
			// grey-binning info per family (3d_glcm.cpp:53-55, 3d_glszm.cpp:490-492, 3d_ngtdm.cpp:195-197); 0 = no binning
			int g = fam == FAM3_GLCM ? ctx->st.coarse_gray_depth : fam == FAM3_GLSZM ? ctx->glszm3_greydepth : ctx->ngtdm3_greydepth;
			if (ctx->st.ibsi)
				g = 0;
			int Lm = g > 0 ? g : maxI;
			if (fam == FAM3_NGTDM && ctx->ngtdm3_radius == 0)
				Lm = 1;		// nothing is accumulated (see ngtdm3d_body)
			for (int c = 0; c < NCLASS; c++)
			{
				const uint32_t n = info.class_count[c], mc = info.class_maxcells[c];
				const uint32_t* lst = ctx->d_tlist.p + ctx->tbase[c];
				int rc = 0;
				if (fam == FAM3_GLCM)
				{
					if (g == 0 && Lm > 4096)
						return fail (ctx, NYXB200_EUNSUPPORTED, "3D GLCM without grey binning needs intensities <= 4096 (max %d)", Lm);
					rc = mc < 32768 ? launch_family3_t<FAM3_GLCM, true> (ctx, im, lst, n, prm, g, Lm, mc, ctx->d_full.p)
						: launch_family3_t<FAM3_GLCM, false> (ctx, im, lst, n, prm, g, Lm, mc, ctx->d_full.p);
				}
				else if (fam == FAM3_GLSZM)
				{
					if (Lm > 65535)
						return fail (ctx, NYXB200_EUNSUPPORTED, "3D GLSZM needs grey levels < 65536 (max %d)", Lm);
					rc = launch_family3_t<FAM3_GLSZM, false> (ctx, im, lst, n, prm, g, Lm, mc, ctx->d_full.p);
				}
				else
				{
					if ((unsigned long long)(Lm + 2) * NGTDM3_NDMAX * (unsigned long long)mc >= 0xFFFFFFFFull || Lm + 2 > 65535)
						return fail (ctx, NYXB200_EUNSUPPORTED, "3D NGTDM: level range %d too large for the exact integer accumulators; set 3ngtdm/greydepth", Lm);
					rc = launch_family3_t<FAM3_NGTDM, false> (ctx, im, lst, n, prm, g, Lm, mc, ctx->d_full.p);
				}
				if (rc)
					return rc;
			}
		}
	}
	else
	for (int fam = FAM_GLCM; fam < FAM_COUNT; fam++)
	{
		stage_begin (ctx, fam_stage[fam]);
		if (!(ctx->fam_mask & (1 << fam)))
			continue;
		if (info.class_count[NCLASS])
			return fail (ctx, NYXB200_EUNSUPPORTED, "a ROI with an AABB larger than 40000 pixels needs the large-ROI texture path (not in this build)");
		for (int c = 0; c < NCLASS; c++)
			if (int rc = launch_family (ctx, fam, im, ctx->d_tlist.p + ctx->tbase[c], info.class_count[c], prm, Lmax,
				info.class_maxcells[c], (int)info.class_maxwh[c], ctx->d_full.p, nodbg))
				return rc;
	}

	// ---- select + D2H
	stage_begin (ctx, ST_SELECT);
	const int nsel = (int)ctx->sel_cols.size();
	if (ctx->sel_dirty)
	{
		CU (ctx->d_sel.reserve (std::max (nsel, 1)));
		if (nsel)
			CU (cudaMemcpyAsync (ctx->d_sel.p, ctx->sel_cols.data(), sizeof (int) * nsel, cudaMemcpyHostToDevice, s));
		ctx->sel_dirty = false;
	}
	const size_t outn = (size_t)nroi * (nsel + 2);
	CU (ctx->d_sel_out.reserve (outn));
	{
		int grid = (int)std::min<size_t> ((outn + 255) / 256, (size_t)ctx->sm_count * 16);
		k_select_columns<<<grid, 256, 0, s>>> (ctx->d_full.p, ncol_all, ctx->d_rois.p, nroi, ctx->d_sel.p, nsel, ctx->st.soft_nan, 0.0, ctx->d_sel_out.p);
		CU (cudaGetLastError());
		ctx->counts[0]++;
	}
	stage_begin (ctx, ST_D2H);
	if (int rc = reserve_pinned_out (ctx, ctx->h_out_used + outn))
		return rc;
	CU (cudaMemcpyAsync (ctx->h_out + ctx->h_out_used, ctx->d_sel_out.p, outn * sizeof (double), cudaMemcpyDeviceToHost, s));
	ctx->h_out_used += outn;
	stage_begin (ctx, ST_TOTAL);
	CU (cudaStreamSynchronize (s));
	(void)img_index;
	return 0;
}

void collect_timings (nyxb200_ctx* ctx)
{
	// stage i lasts from its own event to the next recorded event
	int order[NYXB200_NSTAGES + 1];
	int n = 0;
	for (int i = 0; i <= ST_TOTAL; i++)
		if (ctx->ev_used[i])
			order[n++] = i;
	for (int k = 0; k + 1 < n; k++)
	{
		float ms = 0;
		if (cudaEventElapsedTime (&ms, ctx->ev[order[k]], ctx->ev[order[k + 1]]) == cudaSuccess)
			ctx->stage_ms[order[k]] += ms;
	}
}

int featurize_common (nyxb200_ctx* ctx, const void* inten, int inten_u16, const uint32_t* label, bool on_device,
	int n_img, int64_t D, int64_t H, int64_t W, nyxb200_table* out, const int32_t** row_image)
{
	if (!ctx || !inten || !label || !out)
		return fail (ctx, NYXB200_EINVAL, "null argument");
	if (n_img <= 0 || H <= 0 || W <= 0 || D <= 0)
		return fail (ctx, NYXB200_EINVAL, "bad image dimensions");
	if (D * H > 0x7FFFFFFF)
		return fail (ctx, NYXB200_EUNSUPPORTED, "volume has more than 2^31-1 rows");
	if (H > 0x7FFFFFFF || W > 0x7FFFFFFF)
		return fail (ctx, NYXB200_EUNSUPPORTED, "image side exceeds 2^31-1");
	out->data = nullptr;
	out->nrows = 0;
	out->ncols = (int)ctx->sel_cols.size() + 2;
	out->reserved = 0;
	CU (cudaSetDevice (ctx->st.device));
	for (auto& m : ctx->stage_ms) m = 0;
	for (auto& v : ctx->counts) v = 0;
	ctx->row_image.clear();
	const size_t npix = (size_t)D * (size_t)H * (size_t)W;
	const size_t ibytes = npix * (inten_u16 ? 2 : 4), lbytes = npix * 4;
	ctx->h_out_used = 0;
	for (int k = 0; k < n_img; k++)
	{
		for (auto& u : ctx->ev_used) u = false;
		ImageView im;
		im.H = (int32_t)H;
		im.W = (int32_t)W;
		im.D = (int32_t)D;
		im.inten_u16 = inten_u16;
		const uint8_t* ip = (const uint8_t*)inten + (size_t)k * ibytes;
		const uint32_t* lp = label + (size_t)k * npix;
		stage_begin (ctx, ST_H2D);
		if (on_device)
		{
			im.inten = ip;
			im.label = lp;
		}
		else
		{
			CU (ctx->d_inten.reserve (ibytes));
			CU (ctx->d_label.reserve (npix));
			CU (cudaMemcpyAsync (ctx->d_inten.p, ip, ibytes, cudaMemcpyHostToDevice, ctx->stream));
			CU (cudaMemcpyAsync (ctx->d_label.p, lp, lbytes, cudaMemcpyHostToDevice, ctx->stream));
			im.inten = ctx->d_inten.p;
			im.label = ctx->d_label.p;
		}
		int64_t nr = 0;
		int rc = process_slide (ctx, im, k, &nr);
		if (rc)
			return rc;
		collect_timings (ctx);
		for (int64_t i = 0; i < nr; i++)
			ctx->row_image.push_back (k);
	}
	out->nrows = (int64_t)ctx->row_image.size();
	if (out->nrows)
		out->data = ctx->h_out;
	if (row_image)
		*row_image = ctx->row_image.data();
	float tot = 0;
	for (int i = 0; i < ST_TOTAL; i++)
		tot += ctx->stage_ms[i];
	ctx->stage_ms[ST_TOTAL] = tot;
	return NYXB200_OK;
}

} // namespace

extern "C"
{

int nyxb200_featurize_2d_u32 (nyxb200_ctx* ctx, const uint32_t* inten, const uint32_t* label, int n_img, int64_t H, int64_t W,
	nyxb200_table* out, const int32_t** row_image)
{
	if (ctx && ctx->st.dim != 2)
		return fail (ctx, NYXB200_EINVAL, "this context was created for 3D features (dim = 3)");
	return featurize_common (ctx, inten, 0, label, false, n_img, 1, H, W, out, row_image);
}

int nyxb200_featurize_2d_u16 (nyxb200_ctx* ctx, const uint16_t* inten, const uint32_t* label, int n_img, int64_t H, int64_t W,
	nyxb200_table* out, const int32_t** row_image)
{
	if (ctx && ctx->st.dim != 2)
		return fail (ctx, NYXB200_EINVAL, "this context was created for 3D features (dim = 3)");
	return featurize_common (ctx, inten, 1, label, false, n_img, 1, H, W, out, row_image);
}

int nyxb200_featurize_2d_device (nyxb200_ctx* ctx, const void* inten, int inten_is_u16, const uint32_t* label, int n_img,
	int64_t H, int64_t W, nyxb200_table* out, const int32_t** row_image)
{
	if (ctx && ctx->st.dim != 2)
		return fail (ctx, NYXB200_EINVAL, "this context was created for 3D features (dim = 3)");
	return featurize_common (ctx, inten, inten_is_u16 ? 1 : 0, label, true, n_img, 1, H, W, out, row_image);
}

int nyxb200_featurize_3d_u32 (nyxb200_ctx* ctx, const uint32_t* inten, const uint32_t* label, int n_vol, int64_t D, int64_t H, int64_t W,
	nyxb200_table* out, const int32_t** row_image)
{
	if (ctx && ctx->st.dim != 3)
		return fail (ctx, NYXB200_EINVAL, "this context was created for 2D features (dim = 2)");
	return featurize_common (ctx, inten, 0, label, false, n_vol, D, H, W, out, row_image);
}

int nyxb200_featurize_3d_u16 (nyxb200_ctx* ctx, const uint16_t* inten, const uint32_t* label, int n_vol, int64_t D, int64_t H, int64_t W,
	nyxb200_table* out, const int32_t** row_image)
{
	if (ctx && ctx->st.dim != 3)
		return fail (ctx, NYXB200_EINVAL, "this context was created for 2D features (dim = 2)");
	return featurize_common (ctx, inten, 1, label, false, n_vol, D, H, W, out, row_image);
}

int nyxb200_featurize_3d_device (nyxb200_ctx* ctx, const void* inten, int inten_is_u16, const uint32_t* label, int n_vol,
	int64_t D, int64_t H, int64_t W, nyxb200_table* out, const int32_t** row_image)
{
	if (ctx && ctx->st.dim != 3)
		return fail (ctx, NYXB200_EINVAL, "this context was created for 2D features (dim = 2)");
	return featurize_common (ctx, inten, inten_is_u16 ? 1 : 0, label, true, n_vol, D, H, W, out, row_image);
}

void nyxb200_table_free (nyxb200_ctx* ctx, nyxb200_table* t)
{
	(void)ctx;	// the rows live in the context's grow-only pinned buffer, reused by the next featurize call
	if (t)
	{
		t->data = nullptr;
		t->nrows = 0;
	}
}

int nyxb200_last_timings (const nyxb200_ctx* ctx, float* ms, int n)
{
	if (!ctx || !ms)
		return NYXB200_EINVAL;
	for (int i = 0; i < n && i < NYXB200_NSTAGES; i++)
		ms[i] = ctx->stage_ms[i];
	return NYXB200_OK;
}

int nyxb200_last_counts (const nyxb200_ctx* ctx, int64_t* v, int n)
{
	if (!ctx || !v)
		return NYXB200_EINVAL;
	for (int i = 0; i < n && i < 4; i++)
		v[i] = ctx->counts[i];
	return NYXB200_OK;
}

// ---- test hooks -------------------------------------------------------------------------------------
static int fetch_rois (nyxb200_ctx* ctx)
{
	if (ctx->h_rois_valid)
		return 0;
	ctx->h_rois.resize (ctx->last_info.n_roi);
	if (ctx->last_info.n_roi)
		CU (cudaMemcpy (ctx->h_rois.data(), ctx->d_rois.p, sizeof (RoiRec) * ctx->last_info.n_roi, cudaMemcpyDeviceToHost));
	ctx->h_rois_valid = true;
	return 0;
}

int nyxb200_debug_rois (nyxb200_ctx* ctx, int64_t* out, int64_t cap, int64_t* n)
{
	if (!ctx || !n)
		return fail (ctx, NYXB200_EINVAL, "null argument");
	CU (cudaSetDevice (ctx->st.device));
	if (int rc = fetch_rois (ctx))
		return rc;
	*n = (int64_t)ctx->h_rois.size();
	if (!out)
		return NYXB200_OK;
	for (int64_t i = 0; i < *n && i < cap; i++)
	{
		const RoiRec& r = ctx->h_rois[i];
		int64_t* o = out + i * 8;
		o[0] = r.label; o[1] = r.area; o[2] = r.imin; o[3] = r.imax; o[4] = r.x0; o[5] = r.y0; o[6] = r.w; o[7] = r.h;
	}
	return NYXB200_OK;
}

static int find_roi (nyxb200_ctx* ctx, uint32_t label)
{
	if (fetch_rois (ctx))
		return -1;
	auto it = std::lower_bound (ctx->h_rois.begin(), ctx->h_rois.end(), label, [](const RoiRec& r, uint32_t l) { return r.label < l; });
	if (it == ctx->h_rois.end() || it->label != label)
		return -1;
	return (int)(it - ctx->h_rois.begin());
}

int nyxb200_debug_cloud (nyxb200_ctx* ctx, uint32_t label, uint32_t* out, int64_t cap, int64_t* n)
{
	if (!ctx || !n)
		return fail (ctx, NYXB200_EINVAL, "null argument");
	CU (cudaSetDevice (ctx->st.device));
	int ri = find_roi (ctx, label);
	if (ri < 0)
		return fail (ctx, NYXB200_EINVAL, "label %u not found in the last slide", label);
	if (!(ctx->fam_mask & (1 << FAM_INT)))
		return fail (ctx, NYXB200_EINVAL, "the pixel cloud is only packed when an intensity feature is selected");
	const RoiRec& r = ctx->h_rois[ri];
	*n = r.area;
	if (out && cap >= (int64_t)r.area)
		CU (cudaMemcpy (out, ctx->d_cloud.p + r.cloud_off, sizeof (uint32_t) * r.area, cudaMemcpyDeviceToHost));
	return NYXB200_OK;
}

int nyxb200_debug_matrix (nyxb200_ctx* ctx, uint32_t label, int family, int64_t* trip, int64_t cap, int64_t* n, double* svals)
{
	if (!ctx || !n || family <= FAM_INT || family >= FAM_COUNT)
		return fail (ctx, NYXB200_EINVAL, "bad argument");
	CU (cudaSetDevice (ctx->st.device));
	int ri = find_roi (ctx, label);
	if (ri < 0)
		return fail (ctx, NYXB200_EINVAL, "label %u not found in the last slide", label);
	const RoiRec r = ctx->h_rois[ri];
	const uint32_t ncell = (uint32_t)r.w * (uint32_t)r.h;
	if (tile_class (ncell) >= NCLASS)
		return fail (ctx, NYXB200_EUNSUPPORTED, "ROI too large for the debug dump");
	const Params prm = make_params (ctx);
	const int Lmax = ctx->st.ibsi ? (int)std::max<uint32_t> (r.imax, 1u) : ctx->st.coarse_gray_depth;
	const int maxwh = std::max (r.w, r.h);
	size_t words = 0;
	switch (family)
	{
	case FAM_GLCM: words = (size_t)4 * Lmax * Lmax; break;
	case FAM_GLRLM: words = (size_t)4 * Lmax * maxwh; break;
	case FAM_GLDZM: case FAM_GLSZM: words = ncell + 1; break;
	case FAM_GLDM: words = gldm_smem_words (Lmax); break;
	case FAM_NGLDM: words = ngldm_smem_words (Lmax); break;
	case FAM_NGTDM: words = Lmax + 1; break;
	}
	CU (ctx->d_dbg.reserve (words + 4));
	CU (ctx->d_dbgf.reserve ((size_t)Lmax + 2));
	CU (cudaMemsetAsync (ctx->d_dbg.p, 0, (words + 4) * 4, ctx->stream));
	CU (cudaMemsetAsync (ctx->d_dbgf.p, 0, ((size_t)Lmax + 2) * 8, ctx->stream));
	DebugOut dbg { ctx->d_dbg.p, ctx->d_dbg.p + words, ctx->d_dbgf.p };
	// run the family kernel on this one ROI (list == nullptr -> index k of a 1-element view)
	DevBuf<double> tmp;
	CU (tmp.reserve (NCOL_ALL));
	RoiRec* saved = ctx->d_rois.p;
	ctx->d_rois.p = saved + ri;
	int rc = launch_family (ctx, family, ctx->last_im, nullptr, 1, prm, Lmax, ncell, maxwh, tmp.p, dbg);
	ctx->d_rois.p = saved;
	if (rc)
	{
		tmp.release();
		return rc;
	}
	CU (cudaStreamSynchronize (ctx->stream));
	std::vector<uint32_t> h (words + 4);
	CU (cudaMemcpy (h.data(), ctx->d_dbg.p, (words + 4) * 4, cudaMemcpyDeviceToHost));
	std::vector<double> hs ((size_t)Lmax + 2);
	CU (cudaMemcpy (hs.data(), ctx->d_dbgf.p, ((size_t)Lmax + 2) * 8, cudaMemcpyDeviceToHost));
	tmp.release();
	int64_t cnt = 0;
	auto emit = [&](int64_t ang, int64_t a, int64_t b, int64_t c, double sv) {
		if (trip && cnt < cap)
		{
			trip[cnt * 4] = ang; trip[cnt * 4 + 1] = a; trip[cnt * 4 + 2] = b; trip[cnt * 4 + 3] = c;
			if (svals)
				svals[cnt] = sv;
		}
		cnt++;
	};
	switch (family)
	{
	case FAM_GLCM:
		for (int a = 0; a < 4; a++)
			for (int row = 0; row < Lmax; row++)
				for (int col = 0; col < Lmax; col++)
					if (uint32_t v = h[((size_t)a * Lmax + row) * Lmax + col])
						emit (a, row + 1, col + 1, v, 0);
		break;
	case FAM_GLRLM:
		for (int a = 0; a < 4; a++)
			for (int lv = 0; lv < Lmax; lv++)
				for (int j = 0; j < maxwh; j++)
					if (uint32_t v = h[((size_t)a * Lmax + lv) * maxwh + j])
						emit (a, lv + 1, j + 1, v, 0);
		break;
	case FAM_GLDZM:
	case FAM_GLSZM:
	{
		uint32_t nk = h[words];
		std::sort (h.begin(), h.begin() + nk);	// one key per zone, unsorted on the device
		for (uint32_t i = 0; i < nk; )
		{
			uint32_t j = i;
			while (j < nk && h[j] == h[i]) j++;
			emit (0, h[i] >> 16, h[i] & 0xFFFF, j - i, 0);
			i = j;
		}
		break;
	}
	case FAM_GLDM:
		for (int lv = 1; lv <= Lmax; lv++)
			for (int j = 1; j <= 9; j++)
				if (uint32_t v = h[(size_t)lv * 10 + j])
					emit (0, lv, j, v, 0);
		break;
	case FAM_NGLDM:
		for (int lv = 0; lv <= Lmax; lv++)
			for (int j = 0; j < 9; j++)
				if (uint32_t v = h[(size_t)lv * 9 + j])
					emit (0, lv, j, v, 0);
		break;
	case FAM_NGTDM:
		for (int lv = 0; lv <= Lmax; lv++)
			if (h[lv])
				emit (0, lv, 0, h[lv], hs[lv]);
		break;
	}
	*n = cnt;
	return NYXB200_OK;
}

} // extern "C"
