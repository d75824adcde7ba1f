// nyxus_b200 -- host side of the C ABI (include/nyxus_b200.h): context, feature selection, the per-slide pipeline
// (scan -> compact -> pack -> per-family ROI kernels -> column select -> D2H) and the test/measurement hooks.
// Mirrors, for this path only, what the reference does in featurize_montage_imp -> processMontage ->
// processIntSegImagePairInMemory (new_bindings_py.cpp:562-645, workflow_pythonapi.cpp:39-192).
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <condition_variable>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <string>
#include <vector>
#include "../../include/nyxus_b200.h"
#include "kernels.cuh"
#include "giant.cuh"

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
constexpr int NFEATURES_MAX = 512;

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
	{ "*BASIC_MORPHOLOGY*", 1 << FAM_MORPH },		// featureset.cpp:654
	{ "*ALL_IH*", 1 << FAM_IH },				// featureset.cpp:665: IBSI mode only
	{ "*ALL_TEXTURE*", (1 << FAM_GLCM) | (1 << FAM_GLRLM) | (1 << FAM_GLDZM) | (1 << FAM_GLSZM) | (1 << FAM_GLDM) | (1 << FAM_NGLDM) | (1 << FAM_NGTDM) },
};
// reference featureset.cpp:940-953 (3D group names) for the 3D families on this path
const GroupDef g_groups3d[] = {
	{ "*3D_ALL_INTENSITY*", 1 << FAM3_INT },
	{ "*3D_GLCM*", 1 << FAM3_GLCM },
	{ "*3D_GLSZM*", 1 << FAM3_GLSZM },
	{ "*3D_NGTDM*", 1 << FAM3_NGTDM },
	{ "*3D_GLDM*", 1 << FAM3_GLDM },
	{ "*3D_GLRLM*", 1 << FAM3_GLRLM },
	{ "*3D_GLDZM*", 1 << FAM3_GLDZM },
	{ "*3D_NGLDM*", 1 << FAM3_NGLDM },
	// featureset.h: the texture group = the seven 3D texture families, the neighbour group = GLDM, NGLDM, NGTDM
	{ "*3D_ALL_TEXTURE*", (1 << FAM3_GLCM) | (1 << FAM3_GLSZM) | (1 << FAM3_NGTDM) | (1 << FAM3_GLDM) | (1 << FAM3_GLRLM) | (1 << FAM3_GLDZM) | (1 << FAM3_NGLDM) },
};
// 3D groups of the reference that contain families outside this path (surface morphology, moments)
const char* g_groups3d_unbuilt[] = { "*3D_ALL*", "*3D_ALL_MORPHOLOGY*", "*3D_ALL_NEIGHBOR*", "*3D_MOMENTS*" };

const char* g_stage_names[NYXB200_NSTAGES] = { "h2d", "init", "scan", "compact", "pack", "intensity", "glcm", "glrlm", "gldzm",
	"glszm", "gldm", "ngldm", "ngtdm", "select", "d2h", "total", "families", "giant" };
enum Stage { ST_H2D = 0, ST_INIT, ST_SCAN, ST_COMPACT, ST_PACK, ST_INT, ST_GLCM, ST_GLRLM, ST_GLDZM, ST_GLSZM, ST_GLDM, ST_NGLDM,
	ST_NGTDM, ST_SELECT, ST_D2H, ST_TOTAL, ST_FAMILIES, ST_GIANT };

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

// Process-wide pool of pinned (portable) host buffers for result tables.  A table whose ownership went to the caller
// (nyxb200_table_detach: the pybind module wraps it in a numpy capsule, like the reference's as_pyarray hand-over,
// new_bindings_py.cpp:63-73) comes back through nyxb200_host_release and is reused by the next call: pinning hundreds
// of megabytes per call would cost more than the features.
struct PinnedBuf { void* p; size_t bytes; bool in_use; };
std::mutex g_pool_mtx;
std::vector<PinnedBuf> g_pool;
constexpr size_t POOL_MAX_FREE = 4;

void* pool_acquire (size_t bytes, size_t* got)
{
	{
		std::lock_guard<std::mutex> lk (g_pool_mtx);
		int best = -1;
		for (int i = 0; i < (int)g_pool.size(); i++)
			if (!g_pool[i].in_use && g_pool[i].bytes >= bytes && (best < 0 || g_pool[i].bytes < g_pool[best].bytes))
				best = i;
		if (best >= 0)
		{
			g_pool[best].in_use = true;
			*got = g_pool[best].bytes;
			return g_pool[best].p;
		}
	}
	void* p = nullptr;
	if (cudaHostAlloc (&p, bytes, cudaHostAllocPortable) != cudaSuccess)
	{
		cudaGetLastError();
		return nullptr;
	}
	std::lock_guard<std::mutex> lk (g_pool_mtx);
	g_pool.push_back ({ p, bytes, true });
	*got = bytes;
	return p;
}

void pool_release (void* p)
{
	if (!p)
		return;
	std::vector<void*> drop;
	{
		std::lock_guard<std::mutex> lk (g_pool_mtx);
		for (auto& b : g_pool)
			if (b.p == p)
				b.in_use = false;
		// keep the POOL_MAX_FREE largest free buffers
		for (;;)
		{
			size_t nfree = 0;
			int smallest = -1;
			for (int i = 0; i < (int)g_pool.size(); i++)
				if (!g_pool[i].in_use)
				{
					nfree++;
					if (smallest < 0 || g_pool[i].bytes < g_pool[smallest].bytes)
						smallest = i;
				}
			if (nfree <= POOL_MAX_FREE)
				break;
			drop.push_back (g_pool[smallest].p);
			g_pool.erase (g_pool.begin() + smallest);
		}
	}
	for (void* d : drop)
		cudaFreeHost (d);
}

} // namespace

// One result table shared by several contexts (nyxb200_featurize_2d_list_multi): slide k claims its row range as soon as
// its ROI count is known, in slide order, and its rows are copied device -> host straight to their final place.
struct SharedOut
{
	std::mutex m;
	std::condition_variable cv;
	double* base = nullptr;
	size_t cap_rows = 0, rowlen = 0;
	int n_img = 0;
	int next_slide = 0;		// the slide whose claim is due (claims are made in slide order: a deterministic table)
	int64_t cursor = 0;		// rows claimed so far
	std::vector<int64_t> row0, nrows;	// per slide
	std::vector<char> deferred;		// slides that did not fit the first estimate (second pass)
	bool failed = false;
};

struct nyxb200_ctx
{
	nyxb200_settings st;
	SharedOut* share = nullptr;	// set while the context works for a multi-device call
	// pageable host slides: worker threads copy chunks into pinned slots and send them on their own streams
	// (the driver's own pageable path stages through one thread); 0 = leave it to cudaMemcpyAsync
	int staging_threads = 4;
	int covered_from_scan = 0;	// 2D: COVERED_IMAGE_INTENSITY_RANGE against the scanned slide range (file-based workflows)
	int single_roi = 0;		// the caller runs the reference's whole-slide (single ROI) mode on this call
	int multi_row_pad = 65536;	// slack rows of the shared table's first estimate (test knob: 0 forces the second pass)
	static constexpr int STG_MAX = 16;
	static constexpr size_t STG_CHUNK = (size_t)8 << 20;
	uint8_t* stg_slot[STG_MAX * 2] {};
	cudaStream_t stg_stream[STG_MAX] {};
	cudaEvent_t stg_ev[STG_MAX * 2] {};
	std::vector<std::thread> stg_workers;
	int stg_rc = 0;
	std::string err;
	int sm_count = 148;
	cudaStream_t stream = nullptr;
	// stage begin / end events, one set per ROI batch (stage times are summed over the batches)
	static constexpr int NEV = NYXB200_NSTAGES + 2;
	cudaEvent_t ev[NB_MAX * NEV];
	cudaEvent_t ev_end[NB_MAX * NEV];
	bool ev_used[NB_MAX * NEV];
	bool ev_end_used[NB_MAX * NEV];
	int cur_b = 0;			// batch whose event set stage_begin / stage_end record into
	int roi_batches = 0;		// 0 = automatic
	int big_cta_threads = 1;	// measurement knob: 0 = the CTA-per-ROI kernels always run 128 threads
	cudaStream_t copy_stream = nullptr;	// D2H of a finished batch's rows
	cudaEvent_t ev_sel;
	// host slides travel on their own stream into two alternating device buffers: slide k+1 is uploaded while slide k computes
	cudaStream_t up_stream = nullptr;
	cudaEvent_t ev_up[2];
	// the eight families are independent once the ROI records exist: each runs on its own stream (fork/join by events)
	// so their low-occupancy, issue-bound kernels share the SMs; `concurrent = 0` serialises them (per-stage timing)
	static constexpr int NFSTREAM = 8;
	cudaStream_t fstream[NFSTREAM];
	cudaEvent_t ev_fork, ev_join[NFSTREAM];
	int concurrent = 1;
	cudaStream_t ls = nullptr;	// stream the family launchers submit to (the main stream or a family stream)
	int lsi = 0;			// scratch slot that goes with it
	int open_stage = -1;		// main-stream stage whose end event is still to be recorded

	// selection
	std::vector<FeatureDef> feats;	// the context's feature table (2D or 3D codes)
	bool feat_on[NFEATURES_MAX];
	int fam_mask = 0;		// first-order + texture families (bits of Family)
	bool want_morph = false;	// *BASIC_MORPHOLOGY* columns selected
	bool want_ih = false;		// IH_* columns selected (and IBSI mode on)
	int row_stride = NCOL_TEXINT;	// doubles per ROI of the canonical table: the appended families only when one of them is selected
	int geom_variants = 0;		// geometric moments selected: bit 0 shape (SPAT_ / CENTRAL_ / NORM_ / HU_M), bit 1 intensity (IMOM_)
	bool want_hist = false;	// the opt-in HISTOGRAM feature (one column per grey-depth bin) is selected
	int wmom_variants = 0;		// distance-to-contour weighted moments selected: bit 0 shape (WEIGHTED_ / WT_NORM_CTR_MOM_), bit 1 intensity (IMOM_W)
	DevBuf<uint32_t> d_wmom;	// their per-warp / per-CTA scratch slots
	// masks with large / sparse label values: presence bitmap, per-word ranks, block sums, rank -> label, the re-labelled mask
	DevBuf<uint32_t> d_rl_bitmap, d_rl_wprefix, d_rl_blocks, d_rl_labmap, d_rl_mask;
	bool relabelled = false;	// the current slide runs on d_rl_mask (labels 1 .. K); the label column is mapped back through d_rl_labmap
	std::vector<int> sel_cols;		// canonical column of every output column
	std::vector<std::string> sel_names;
	std::vector<int> sel_codes;
	DevBuf<int> d_sel;
	bool sel_dirty = true;

	// device buffers (grow-only)
	DevBuf<uint8_t> d_inten[2];
	DevBuf<uint32_t> d_label[2];
	DevBuf<LabelStat> d_stats;
	DevBuf<uint2> d_ztab;
	uint32_t stats_cap = 1u << 20;
	DevBuf<SlideInfo> d_info;
	DevBuf<uint32_t> d_blkcnt;
	DevBuf<unsigned long long> d_blkarea, d_blkcells;
	DevBuf<uint16_t> d_tiles;	// binned-tile store (k_pack_tiles)
	int use_tile_store = 1;
	DevBuf<RoiRec> d_rois;
	DevBuf<uint32_t> d_classslot, d_tlist, d_alist, d_bases;
	DevBuf<uint32_t> d_cloud;
	DevBuf<double> d_full, d_sel_out;
	DevBuf<uint32_t> d_scratch[NFSTREAM];	// per-family global matrix slots (families run concurrently)
	DevBuf<unsigned long long> d_rowoff;
	DevBuf<uint32_t> d_dbg;
	DevBuf<double> d_dbgf;
	// large-ROI (giant) path: global-memory tile + work arrays of ONE ROI at a time (giant.cuh)
	DevBuf<uint32_t> g_cell, g_cell2, g_auxA, g_auxB, g_auxC, g_mat, g_hcnts, g_flags, g_idx;
	DevBuf<unsigned long long> g_acc, g_hkeys;
	DevBuf<double> g_part;
	DevBuf<RoiRec> g_recs;
	int64_t giant_rois = 0;		// ROIs of the last call that went through the giant path

	// state of the last processed image (for the debug hooks)
	ImageView last_im {};
	SlideInfo last_info {};
	uint32_t tbase[NTK] {}, abase[NAK] {};	// list offsets per (batch, class) key
	uint32_t giant_tbase = 0, giant_tcount = 0, giant_abase = 0, giant_acount = 0;	// oversize ROIs of all batches (contiguous)
	std::vector<RoiRec> h_rois;
	bool h_rois_valid = false;

	double* h_out = nullptr;		// pinned result buffer (grow-only)
	size_t h_out_cap = 0, h_out_used = 0;

	int glszm3_greydepth = 0, ngtdm3_greydepth = 0, ngtdm3_radius = 0;	// env_features.cpp:712-736: zero unless set
	int gldm3_greydepth = 0;	// "3gldm/greydepth" (env_metaparams.cpp:158-173)
	int glrlm3_greydepth = 0;	// "3glrlm/greydepth" (env_metaparams.cpp:204-220)
	int glcm_greydepth = 0;		// GLCM_GREYDEPTH when set through "glcm/greydepth" / "3glcm/greydepth" (env_metaparams.cpp:83-91,112-120)
	bool glcm_greydepth_set = false;	// else it follows coarse_gray_depth (env_features.cpp:730)

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
	ctx->want_morph = false;
	ctx->want_ih = false;
	ctx->geom_variants = 0;
	ctx->wmom_variants = 0;
	ctx->want_hist = false;
	static const int angles[4] = { 0, 45, 90, 135 };
	// enabled features in ascending code order (reference featureset.cpp getEnabledFeatures; the table is already sorted)
	for (int i = 0; i < (int)ctx->feats.size(); i++)
	{
		if (!ctx->feat_on[i])
			continue;
		const FeatureDef& f = ctx->feats[i];
		if (f.family == FAM_IH)
		{
			// the reference strips the IH family when IBSI compliance is off (env_features.cpp:513-527): no columns at all
			if (!ctx->st.ibsi)
				continue;
			ctx->want_ih = true;
		}
		else if (f.family == FAM_HIST)
		{
			// HISTOGRAM_BIN_0 .. _N-1, N = coarse grey depth (output_2_buffer.cpp:158-168; a negative depth gives no column)
			const int nbins = ctx->st.dim == 3 ? 0 : ctx->st.coarse_gray_depth;
			ctx->want_hist = nbins > 0;
			for (int k = 0; k < nbins; k++)
			{
				ctx->sel_cols.push_back (COL_HIST + k);
				ctx->sel_codes.push_back (f.code);
				ctx->sel_names.push_back (std::string (f.name) + "_BIN_" + std::to_string (k));
			}
			continue;
		}
		else if (f.family == FAM_MORPH)
			ctx->want_morph = true;
		else if (f.family == FAM_GEOM)
			ctx->geom_variants |= f.col0 < COL_GEOM + GM_NCOL ? 1 : 2;
		else if (f.family == FAM_WMOM)
			ctx->wmom_variants |= f.col0 < COL_WMOM + WMOM_NCOL ? 1 : 2;
		else
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

// the feature table of a 2D / 3D context
std::vector<FeatureDef> feature_table (int dim)
{
	std::vector<FeatureDef> v;
	if (dim == 3)
		for (const auto& f : g_features3d_src)
			v.push_back ({ f.name, f.code, f.family, f.col0, 1 });
	else
		for (const auto& f : g_features2d)
			v.push_back (f);
	return v;
}

// Group / feature names -> enabled flags, exactly as Environment::expand_featuregroups does (env_features.cpp:456-528);
// needs no device (nyxb200_validate_features)
int resolve_features (int dim, const std::vector<FeatureDef>& feats, const char* const* names, int n, bool* on, std::string& msg)
{
	const int nf = (int)feats.size();
	const bool d3 = dim == 3;
	for (int i = 0; i < NFEATURES_MAX; i++)
		on[i] = false;	// env_features.cpp:459 "initially, no features are enabled"
	for (int k = 0; k < n; k++)
	{
		const char* s = names[k];
		if (!s)
		{
			msg = "null feature name";
			return NYXB200_EINVAL;
		}
		bool found = false;
		const GroupDef* groups = d3 ? g_groups3d : g_groups2d;
		const int ngroups = d3 ? (int)(sizeof (g_groups3d) / sizeof (g_groups3d[0])) : (int)(sizeof (g_groups2d) / sizeof (g_groups2d[0]));
		for (int gi = 0; gi < ngroups && !found; gi++)
			if (!strcmp (groups[gi].name, s))
			{
				for (int i = 0; i < nf; i++)
					if (groups[gi].fam_mask & (1 << feats[i].family))
						on[i] = true;
				found = true;
			}
		if (found)
			continue;
		if (!d3)
		{
			// env_features.cpp:316-333: the shape moments (Smoms2D, 90 features), the intensity moments (Imoms2D, 90), or both --
			// each the unweighted block and the distance-to-contour weighted block of its variant
			const int gm = !strcmp (s, "*GEOMOMS*") ? 3 : !strcmp (s, "*SGEOMOMS*") ? 1 : !strcmp (s, "*IGEOMOMS*") ? 2 : 0;
			if (gm)
			{
				for (int i = 0; i < nf; i++)
				{
					int v = 0;
					if (feats[i].family == FAM_GEOM)
						v = feats[i].col0 < COL_GEOM + GM_NCOL ? 1 : 2;
					else if (feats[i].family == FAM_WMOM)
						v = feats[i].col0 < COL_WMOM + WMOM_NCOL ? 1 : 2;
					if (v & gm)
						on[i] = true;
				}
				continue;
			}
		}
		if (d3)
			for (const char* g : g_groups3d_unbuilt)
				if (!strcmp (g, s))
				{
					msg = std::string ("feature group '") + s + "' contains 3D families outside this path "
						"(built: *3D_ALL_INTENSITY*, *3D_ALL_TEXTURE* = *3D_GLCM*, *3D_GLSZM*, *3D_NGTDM*, *3D_GLDM*, *3D_GLRLM*, *3D_GLDZM*, *3D_NGLDM*)";
					return NYXB200_EUNSUPPORTED;
				}
		for (int i = 0; i < nf; i++)
			if (!strcmp (feats[i].name, s))
			{
				on[i] = true;
				found = true;
				break;
			}
		if (!found)	// env_features.cpp:496,507
		{
			msg = std::string ("Error: '") + s + "' is not a valid " + (d3 ? "3D" : "2D") + " feature name ";
			return NYXB200_EINVAL;
		}
	}
	return NYXB200_OK;
}

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
	{
		// ln(k) of small integer counts (nyx_common.h ln_count), per device
		static double lnt[LN_TABLE_N];
		for (uint32_t k = 0; k < LN_TABLE_N; k++)
			lnt[k] = k ? std::log ((double)k) : -INFINITY;
		if ((e = cudaMemcpyToSymbol (g_ln_table, lnt, sizeof (lnt))) != cudaSuccess)
			return fail (nullptr, NYXB200_ECUDA, "cudaMemcpyToSymbol: %s", cudaGetErrorString (e));
		static KWeights kwt[KW_TABLE_N];
		for (uint32_t k = 0; k < KW_TABLE_N; k++)
		{
			const double kd = (double)k;
			kwt[k].inv1pk2 = 1.0 / (1.0 + kd * kd);
			kwt[k].inv1pk = 1.0 / (1.0 + kd);
			kwt[k].invk2 = k ? 1.0 / (kd * kd) : 0.0;
		}
		if ((e = cudaMemcpyToSymbol (g_kw_table, kwt, sizeof (kwt))) != cudaSuccess)
			return fail (nullptr, NYXB200_ECUDA, "cudaMemcpyToSymbol: %s", cudaGetErrorString (e));
	}
	nyxb200_ctx* ctx = new nyxb200_ctx();
	ctx->st = *s;
	ctx->sm_count = prop.multiProcessorCount;
	ctx->feats = feature_table (s->dim);
	for (int i = 0; i < NFEATURES_MAX; i++)
		ctx->feat_on[i] = false;
	if (cudaStreamCreateWithFlags (&ctx->stream, cudaStreamNonBlocking) != cudaSuccess)
	{
		delete ctx;
		return fail (nullptr, NYXB200_ECUDA, "cudaStreamCreate failed");
	}
	for (auto& ev : ctx->ev)
		cudaEventCreate (&ev);
	for (auto& ev : ctx->ev_end)
		cudaEventCreate (&ev);
	for (auto& fs : ctx->fstream)
		cudaStreamCreateWithFlags (&fs, cudaStreamNonBlocking);
	cudaStreamCreateWithFlags (&ctx->copy_stream, cudaStreamNonBlocking);
	cudaStreamCreateWithFlags (&ctx->up_stream, cudaStreamNonBlocking);
	for (auto& ev : ctx->ev_up)
		cudaEventCreateWithFlags (&ev, cudaEventDisableTiming);
	cudaEventCreateWithFlags (&ctx->ev_sel, cudaEventDisableTiming);
	cudaEventCreateWithFlags (&ctx->ev_fork, cudaEventDisableTiming);
	for (auto& ev : ctx->ev_join)
		cudaEventCreateWithFlags (&ev, cudaEventDisableTiming);
	ctx->ls = ctx->stream;
	rebuild_selection (ctx);
	*out = ctx;
	return NYXB200_OK;
}

void nyxb200_destroy (nyxb200_ctx* ctx)
{
	if (!ctx)
		return;
	cudaSetDevice (ctx->st.device);
	cudaDeviceSynchronize();
	ctx->d_ztab.release(); ctx->d_sel.release(); for (int i = 0; i < 2; i++) { ctx->d_inten[i].release(); ctx->d_label[i].release(); } ctx->d_stats.release(); ctx->d_info.release();
	ctx->d_blkcnt.release(); ctx->d_blkarea.release(); ctx->d_rois.release(); ctx->d_classslot.release(); ctx->d_tlist.release();
	ctx->d_alist.release(); ctx->d_bases.release(); ctx->d_cloud.release(); ctx->d_full.release(); ctx->d_sel_out.release();
	for (auto& b : ctx->d_scratch) b.release();
	ctx->d_wmom.release();
	ctx->d_rl_bitmap.release(); ctx->d_rl_wprefix.release(); ctx->d_rl_blocks.release(); ctx->d_rl_labmap.release(); ctx->d_rl_mask.release();
	ctx->g_cell.release(); ctx->g_cell2.release(); ctx->g_auxA.release(); ctx->g_auxB.release(); ctx->g_auxC.release(); ctx->g_mat.release();
	ctx->g_hcnts.release(); ctx->g_flags.release(); ctx->g_idx.release(); ctx->g_acc.release(); ctx->g_hkeys.release(); ctx->g_part.release();
	ctx->g_recs.release();
	ctx->d_rowoff.release(); ctx->d_dbg.release(); ctx->d_dbgf.release(); ctx->d_blkcells.release(); ctx->d_tiles.release();
	pool_release (ctx->h_out);
	for (auto& t : ctx->stg_workers)
		if (t.joinable())
			t.join();
	for (auto& p : ctx->stg_slot)
		if (p)
			cudaFreeHost (p);
	for (auto& st : ctx->stg_stream)
		if (st)
			cudaStreamDestroy (st);
	for (auto& ev : ctx->stg_ev)
		if (ev)
			cudaEventDestroy (ev);
	for (auto& ev : ctx->ev)
		cudaEventDestroy (ev);
	for (auto& ev : ctx->ev_end)
		cudaEventDestroy (ev);
	for (auto& ev : ctx->ev_join)
		cudaEventDestroy (ev);
	cudaEventDestroy (ctx->ev_fork);
	cudaEventDestroy (ctx->ev_sel);
	for (auto& fs : ctx->fstream)
		cudaStreamDestroy (fs);
	cudaStreamDestroy (ctx->copy_stream);
	cudaStreamDestroy (ctx->up_stream);
	for (auto& ev : ctx->ev_up)
		cudaEventDestroy (ev);
	cudaStreamDestroy (ctx->stream);
	delete ctx;
}

int nyxb200_select_features (nyxb200_ctx* ctx, const char* const* names, int n)
{
	if (!ctx || (!names && n))
		return fail (ctx, NYXB200_EINVAL, "null argument");
	bool on[NFEATURES_MAX];
	std::string msg;
	if (int rc = resolve_features (ctx->st.dim, ctx->feats, names, n, on, msg))
		return fail (ctx, rc, "%s", msg.c_str());
	memcpy (ctx->feat_on, on, sizeof (on));
	rebuild_selection (ctx);
	return NYXB200_OK;
}

int nyxb200_validate_features (int dim, const char* const* names, int n, char* msg, size_t msgcap)
{
	if (msg && msgcap)
		msg[0] = 0;
	std::string m;
	int rc = NYXB200_OK;
	if ((dim != 2 && dim != 3) || (!names && n))
	{
		rc = NYXB200_EINVAL;
		m = "bad argument";
	}
	else
	{
		bool on[NFEATURES_MAX];
		rc = resolve_features (dim, feature_table (dim), names, n, on, m);
	}
	if (rc && msg && msgcap)
		snprintf (msg, msgcap, "%s", m.c_str());
	return rc;
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
	else if (!strcmp (key, "coarse_gray_depth"))
	{
		// NyxSetting::GREYDEPTH.  Negative = PyRadiomics bin count (texture_feature.h:23-50,106-118): not reachable from
		// the reference's Python API (coarse_gray_depth is unsigned there, new_bindings_py.cpp:81,167) but it is what the
		// reference's own gtests set on Fsettings, so the C ABI takes it.  Setting it re-derives GLCM_GREYDEPTH like
		// Environment::compile_feature_settings does (env_features.cpp:720,730).
		if (iv == 0 || iv > 65535 || iv < -65535)
			return fail (ctx, NYXB200_EINVAL, "grey depth must be in 1..65535 (or -65535..-1: PyRadiomics bin count)");
		ctx->st.coarse_gray_depth = iv;
		ctx->glcm_greydepth_set = false;
		rebuild_selection (ctx);	// the HISTOGRAM block has one column per bin
	}
	else if (!strcmp (key, "glcm/greydepth") || !strcmp (key, "3glcm/greydepth"))
	{
		// env_metaparams.cpp:83-91 / :112-120: writes ONLY GLCM_GREYDEPTH (any integer).  2D: feeds the blank-ROI test of
		// GLCMFeature::calculate (glcm.cpp:22-28), the matrix keeps binning with GREYDEPTH (glcm.cpp:354).  3D: the
		// binning of D3_GLCM_feature (3d_glcm.cpp:53-57); 0 = no binning, negative = PyRadiomics bin count.
		if (!strcmp (key, "3glcm/greydepth") != (ctx->st.dim == 3))
			return fail (ctx, NYXB200_EINVAL, "'%s' does not apply to a %dD context", key, ctx->st.dim);
		if (ctx->st.dim == 3 && (iv > 65535 || iv < -65535))	// 2D: any integer (it only feeds the blank-ROI comparison)
			return fail (ctx, NYXB200_EUNSUPPORTED, "grey depth must be in -65535..65535");
		ctx->glcm_greydepth = iv;
		ctx->glcm_greydepth_set = true;
	}
	else if (!strcmp (key, "ibsi"))
	{
		ctx->st.ibsi = iv != 0;
		rebuild_selection (ctx);	// the IH_* columns exist in IBSI mode only
	}
	else if (!strcmp (key, "glcm/symmetric"))
		ctx->st.glcm_symmetric = iv != 0;
	else if (!strcmp (key, "soft_nan"))
		ctx->st.soft_nan = value;
	else if (!strcmp (key, "concurrent_families"))	// measurement knob: 0 = one family after the other (clean per-stage times)
		ctx->concurrent = iv != 0;
	else if (!strcmp (key, "big_cta_threads"))
		ctx->big_cta_threads = iv != 0;
	else if (!strcmp (key, "roi_batches"))			// 0 = automatic, else 1..NB_MAX
	{
		if (iv < 0 || iv > NB_MAX)
			return fail (ctx, NYXB200_EINVAL, "roi_batches must be in 0..%d", NB_MAX);
		ctx->roi_batches = iv;
	}
	else if (!strcmp (key, "tile_store"))			// test knob: 0 = family kernels bin straight from the slide
		ctx->use_tile_store = iv != 0;
	else if (!strcmp (key, "single_roi"))	// 1: whole-slide mode (SINGLEROI): features whose reference code branches on it can tell
		ctx->single_roi = iv != 0;
	else if (!strcmp (key, "covered_range_from_scan"))	// 1: the slide came from a file (featurize_directory / featurize_files)
		ctx->covered_from_scan = iv != 0;
	else if (!strcmp (key, "multi_row_pad"))
		ctx->multi_row_pad = iv < 0 ? 0 : iv;
	else if (!strcmp (key, "staging_threads"))		// pageable host slides: copy threads (0 = the driver's pageable path)
	{
		if (iv < 0 || iv > nyxb200_ctx::STG_MAX)
			return fail (ctx, NYXB200_EINVAL, "staging_threads must be in 0..%d", nyxb200_ctx::STG_MAX);
		ctx->staging_threads = iv;
	}
	else if (!strcmp (key, "3glszm/greydepth"))		// env_metaparams.cpp:222-233 (negative = PyRadiomics bin count)
	{
		if (iv < -65535 || iv > 65535)
			return fail (ctx, NYXB200_EUNSUPPORTED, "grey depth must be in -65535..65535");
		ctx->glszm3_greydepth = iv;
	}
	else if (!strcmp (key, "3gldm/greydepth"))		// env_metaparams.cpp:158-173 (negative = PyRadiomics bin count)
	{
		if (iv < -65535 || iv > 65535)
			return fail (ctx, NYXB200_EUNSUPPORTED, "grey depth must be in -65535..65535");
		ctx->gldm3_greydepth = iv;
	}
	else if (!strcmp (key, "3glrlm/greydepth"))		// env_metaparams.cpp:204-220 (negative = PyRadiomics bin count)
	{
		if (iv < -65535 || iv > 65535)
			return fail (ctx, NYXB200_EUNSUPPORTED, "grey depth must be in -65535..65535");
		ctx->glrlm3_greydepth = iv;
	}
	else if (!strcmp (key, "3ngtdm/greydepth"))		// env_metaparams.cpp:175-186
	{
		if (iv < -65535 || iv > 65535)
			return fail (ctx, NYXB200_EUNSUPPORTED, "grey depth must be in -65535..65535");
		ctx->ngtdm3_greydepth = iv;
	}
	else if (!strcmp (key, "3ngtdm/radius"))		// env_metaparams.cpp:187-197
	{
		if (iv < 0)
			return fail (ctx, NYXB200_EINVAL, "3ngtdm/radius: expecting a positive integer");
		if (iv > NGTDM3_RMAX)
			return fail (ctx, NYXB200_EUNSUPPORTED, "3ngtdm/radius > %d is not supported", NGTDM3_RMAX);
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

// The dynamic shared-memory limit of a kernel is per-device state shared by every context (and host thread) that uses
// the device: it is only ever RAISED, under a lock -- two contexts on one device (or two threads of a multi-device call
// given contexts of the same device) must not lower each other's limit between a set and the launch that relies on it.
std::mutex g_attr_mtx;
std::map<std::pair<const void*, int>, size_t> g_attr_bytes;
template <class K> int set_smem (nyxb200_ctx* ctx, K kernel, size_t bytes)
{
	if (bytes <= 48 * 1024)
		return 0;
	std::lock_guard<std::mutex> lk (g_attr_mtx);
	size_t& cur = g_attr_bytes[{ reinterpret_cast<const void*> (kernel), ctx->st.device }];
	if (bytes > cur)
	{
		CU (cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
		cur = bytes;
	}
	return 0;
}

// words of family matrices for a class
size_t family_words (int fam, bool p16, int Lmax, uint32_t maxcells, int maxwh, bool rad)
{
	switch (fam)
	{
	case FAM_GLCM: return p16 ? glcm_smem_words<true> (Lmax, 4, rad, maxcells) : glcm_smem_words<false> (Lmax, 4, rad, maxcells);
	case FAM_GLRLM: return glrlm_smem_words (Lmax, maxwh, (int)maxcells);
	case FAM_GLDZM: return gldzm_smem_words (Lmax, gldzm_nd_bound (maxcells, maxwh), maxcells);
	case FAM_GLSZM: return glszm_smem_words (maxcells, glszm_hash_slots (Lmax, maxcells), Lmax <= GLSZM_LMARG_MAX ? Lmax : 0);
	case FAM_GLDM: return gldm_smem_words (Lmax, maxcells);
	case FAM_NGLDM: return ngldm_smem_words (Lmax, maxcells);
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
	const size_t tilew = tile_words (maxcells);
	const size_t famw = (family_words (FAM, P16, Lmax, maxcells, maxwh, !prm.ibsi && prm.greydepth < 0) + 1) & ~(size_t)1;
	// small tiles of the single-matrix families: one warp per ROI
	if (FAM != FAM_GLCM && FAM != FAM_GLRLM && !dbg.u32 && maxcells <= 1024 && tilew + famw <= WARP_SLICE_LIMIT_WORDS)
	{
		const size_t bytes = (tilew + famw) * 4 * ROI_WARPS;
		auto kw = k_roi_family_w<FAM>;
		if (set_smem (ctx, kw, bytes))
			return NYXB200_ECUDA;
		const int per_sm = (int)std::max<size_t> (1, std::min<size_t> (16, (size_t)(227 * 1024) / (bytes + 1024)));
		const int grid = (int)std::min<uint32_t> ((n + ROI_WARPS - 1) / ROI_WARPS, (uint32_t)(ctx->sm_count * per_sm));
		kw<<<grid, ROI_THREADS, bytes, ctx->ls>>> (im, ctx->d_rois.p, list, n, prm, Lmax, maxcells, maxwh, out, ctx->row_stride, famw);
		CU (cudaGetLastError());
		ctx->counts[0]++;
		return 0;
	}
	const size_t ringw = FAM == FAM_GLCM || FAM == FAM_GLRLM ? PF_RING_WORDS : 0;	// in-flight list entries / records of the pipelined loop
	size_t smem_words = RED_WORDS + tilew + famw + ringw;
	uint32_t* gs = nullptr;
	int grid = (int)std::min<uint32_t> (n, (uint32_t)ctx->sm_count * 16);
	if (smem_words > SMEM_LIMIT_WORDS)
	{
		// matrices go to a per-CTA slot in global memory (L2 resident), the tile stays in shared memory
		smem_words = RED_WORDS + tilew + ringw;
		grid = (int)std::min<uint32_t> (n, (uint32_t)ctx->sm_count * 2);
		CU (ctx->d_scratch[ctx->lsi].reserve ((size_t)grid * famw));
		gs = ctx->d_scratch[ctx->lsi].p;
		CU (cudaMemsetAsync (gs, 0, (size_t)grid * famw * 4, ctx->ls));	// bodies expect (and leave) an all-zero region
	}
	auto kern = gs ? k_roi_family<FAM, P16, true> : k_roi_family<FAM, P16, false>;
	if (set_smem (ctx, kern, smem_words * 4))
		return NYXB200_ECUDA;
	// single-matrix families on tiles that leave room for at most three CTAs per SM: twice the threads on the same shared memory
	const int threads = ctx->big_cta_threads && FAM != FAM_GLCM && FAM != FAM_GLRLM && smem_words * 4 > ROI_BIG_SMEM_BYTES ? ROI_THREADS_BIG : ROI_THREADS;
	kern<<<grid, threads, smem_words * 4, ctx->ls>>> (im, ctx->d_rois.p, list, n, prm, Lmax, maxcells, maxwh, out,
		ctx->row_stride, gs, famw, dbg);
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
	const uint32_t npad = std::max (next_pow2 (maxarea), 64u);	// the warp-per-ROI register sort works on >= 64 keys
	const size_t memw = (intensity_smem_words (npad, prm.greydepth) + 3) & ~(size_t)3;	// 16-byte aligned warp slices
	if (npad <= 1024 && memw <= WARP_SLICE_LIMIT_WORDS)	// small clouds: one warp per ROI
	{
		const size_t bytes = memw * 4 * ROI_WARPS;
		if (set_smem (ctx, k_roi_intensity_w, bytes))
			return NYXB200_ECUDA;
		const int per_sm = (int)std::max<size_t> (1, std::min<size_t> (16, (size_t)(227 * 1024) / (bytes + 1024)));
		const int grid = (int)std::min<uint32_t> ((n + ROI_WARPS - 1) / ROI_WARPS, (uint32_t)(ctx->sm_count * per_sm));
		k_roi_intensity_w<<<grid, ROI_THREADS, bytes, ctx->ls>>> (ctx->d_rois.p, list, n, prm, ctx->d_cloud.p, out, ostride, memw);
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
		CU (ctx->d_scratch[ctx->lsi].reserve ((size_t)grid * memw));
		gs = ctx->d_scratch[ctx->lsi].p;
	}
	auto kern = gs ? k_roi_intensity<true> : k_roi_intensity<false>;
	if (set_smem (ctx, kern, smem_words * 4))
		return NYXB200_ECUDA;
	kern<<<grid, ctx->big_cta_threads && (smem_words * 4 > ROI_BIG_SMEM_BYTES || gs) ? ROI_THREADS_BIG : ROI_THREADS, smem_words * 4, ctx->ls>>> (ctx->d_rois.p, list, n, prm, ctx->d_cloud.p, npad, out,
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
	const size_t tilew = tile_words (maxcells);
	size_t famw = 0;
	if (FAM3 == FAM3_GLCM) famw = P16 ? glcm_smem_words<true> (Lmax, 4, greyinfo < 0) : glcm_smem_words<false> (Lmax, 4, greyinfo < 0);
	if (FAM3 == FAM3_GLSZM) famw = glszm3d_smem_words (maxcells, glszm_hash_slots (Lmax, maxcells));
	if (FAM3 == FAM3_NGTDM) famw = ngtdm3d_smem_words (Lmax, prm.ngtdm3_radius);
	if (FAM3 == FAM3_GLDM) famw = gldm3d_smem_words (Lmax, maxcells);
	if (FAM3 == FAM3_GLRLM) famw = glrlm3d_smem_words (Lmax, prm.cube_maxdim, maxcells);
	if (FAM3 == FAM3_GLDZM) famw = gldzm3d_smem_words (Lmax, gldzm_nd_bound (maxcells, prm.cube_maxdim), maxcells);
	if (FAM3 == FAM3_NGLDM) famw = ngldm3d_smem_words (Lmax);
	famw = (famw + 1) & ~(size_t)1;
	size_t smem_words = RED_WORDS + tilew + famw;
	uint32_t* gs = nullptr;
	int grid = (int)std::min<uint32_t> (n, (uint32_t)ctx->sm_count * 8);
	if (smem_words > SMEM_LIMIT_WORDS)
	{
		smem_words = RED_WORDS + tilew;
		grid = (int)std::min<uint32_t> (n, (uint32_t)ctx->sm_count * 2);
		CU (ctx->d_scratch[ctx->lsi].reserve ((size_t)grid * famw));
		gs = ctx->d_scratch[ctx->lsi].p;
		CU (cudaMemsetAsync (gs, 0, (size_t)grid * famw * 4, ctx->ls));
	}
	auto kern = gs ? k_roi_family3<FAM3, P16, true> : k_roi_family3<FAM3, P16, false>;
	if (set_smem (ctx, kern, smem_words * 4))
		return NYXB200_ECUDA;
	kern<<<grid, ctx->big_cta_threads && smem_words * 4 > ROI_BIG_SMEM_BYTES ? ROI_THREADS_BIG : ROI_THREADS, smem_words * 4, ctx->ls>>> (im, ctx->d_rois.p, list, n, prm, greyinfo, Lmax, maxcells, out,
		NCOL3_ALL, gs, famw);
	CU (cudaGetLastError());
	ctx->counts[0]++;
	return 0;
}

// 3D cubes beyond the shared-memory classes: cube + work arrays in per-CTA slots of global memory (k_roi_family3_big)
template <int FAM3>
int launch_family3_big (nyxb200_ctx* ctx, const ImageView& im, const uint32_t* list, uint32_t n, const Params& prm, int greyinfo,
	int Lmax, unsigned long long maxcells, double* out)
{
	if (!n)
		return 0;
	if (maxcells > 0x40000000ull)
		return fail (ctx, NYXB200_EUNSUPPORTED, "a 3D ROI whose bounding cube exceeds 2^30 voxels is not supported");
	const size_t tilew = ((size_t)maxcells + 7) & ~(size_t)7;
	uint32_t slots = 0;
	size_t famw = 0;
	if (FAM3 == FAM3_GLCM) famw = glcm_smem_words<false> (Lmax, 4, greyinfo < 0);
	if (FAM3 == FAM3_GLSZM)
	{
		slots = glszm3d_big_slots (Lmax, maxcells);
		famw = glszm3d_big_words (maxcells, slots);
	}
	if (FAM3 == FAM3_NGTDM) famw = ngtdm3d_smem_words (Lmax, prm.ngtdm3_radius);
	if (FAM3 == FAM3_GLDM) famw = gldm3d_smem_words (Lmax, (uint32_t)maxcells);
	if (FAM3 == FAM3_GLRLM) famw = glrlm3d_smem_words (Lmax, prm.cube_maxdim, maxcells);
	if (FAM3 == FAM3_GLDZM) famw = gldzm3d_smem_words (Lmax, gldzm_nd_bound ((uint32_t)maxcells, prm.cube_maxdim), maxcells);
	if (FAM3 == FAM3_NGLDM) famw = ngldm3d_smem_words (Lmax);
	famw = (famw + 3) & ~(size_t)3;
	const size_t slotw = tilew + famw;
	size_t freeb = 0, totalb = 0;
	CU (cudaMemGetInfo (&freeb, &totalb));
	const size_t budget = std::min<size_t> ((size_t)4 << 30, freeb / 2 + ctx->d_scratch[ctx->lsi].cap * 4);
	if (slotw * 4 > budget)
		return fail (ctx, NYXB200_ENOMEM, "a 3D ROI of %llu voxels needs %zu MB of work space", maxcells, slotw * 4 >> 20);
	const int grid = (int)std::max<size_t> (1, std::min<size_t> (std::min<size_t> (n, (size_t)ctx->sm_count), budget / (slotw * 4)));
	CU (ctx->d_scratch[ctx->lsi].reserve ((size_t)grid * slotw));
	uint32_t* gs = ctx->d_scratch[ctx->lsi].p;
	CU (cudaMemsetAsync (gs, 0, (size_t)grid * slotw * 4, ctx->ls));
	auto kern = k_roi_family3_big<FAM3>;
	kern<<<grid, ROI_THREADS, RED_WORDS * 4, ctx->ls>>> (im, ctx->d_rois.p, list, n, prm, greyinfo, Lmax, maxcells, out, NCOL3_ALL,
		gs, slotw, tilew, slots);
	CU (cudaGetLastError());
	ctx->counts[0]++;
	return 0;
}

Params make_params (const nyxb200_ctx* ctx)
{
	Params p;
	p.greydepth = ctx->st.coarse_gray_depth;
	p.glcm_greydepth = ctx->glcm_greydepth_set ? ctx->glcm_greydepth : ctx->st.coarse_gray_depth;
	p.ibsi = ctx->st.ibsi;
	p.glcm_offset = ctx->st.glcm_offset;
	p.glcm_symmetric = ctx->st.glcm_symmetric;
	p.softnan = ctx->st.soft_nan;
	p.glszm3_greydepth = ctx->glszm3_greydepth;
	p.ngtdm3_greydepth = ctx->ngtdm3_greydepth;
	p.ngtdm3_radius = ctx->ngtdm3_radius;
	p.hist_col = ctx->want_hist && ctx->st.dim != 3 ? COL_HIST : 0;
	return p;
}


// ====================================================================================================
// large-ROI (giant) path: host loop over the oversize ROIs of the slide, whole-GPU kernels per ROI (giant.cuh)
// ====================================================================================================
constexpr unsigned long long GIANT_MAX_CELLS = 0xFFFFFFF0ull;	// 32-bit cell indices
constexpr size_t GIANT_MAX_MATRIX_WORDS = (size_t)1 << 28;	// 1 GiB per dense family matrix

int fetch_class_rois (nyxb200_ctx* ctx, const uint32_t* dlist, uint32_t n, std::vector<RoiRec>& recs, std::vector<uint32_t>& idx)
{
	recs.resize (n);
	idx.resize (n);
	if (!n)
		return 0;
	CU (ctx->g_recs.reserve (n));
	CU (ctx->g_idx.reserve (n));
	k_gather_rois<<<(n + 255) / 256, 256, 0, ctx->stream>>> (ctx->d_rois.p, dlist, n, ctx->g_recs.p, ctx->g_idx.p);
	CU (cudaGetLastError());
	ctx->counts[0]++;
	CU (cudaMemcpyAsync (recs.data(), ctx->g_recs.p, sizeof (RoiRec) * n, cudaMemcpyDeviceToHost, ctx->stream));
	CU (cudaMemcpyAsync (idx.data(), ctx->g_idx.p, sizeof (uint32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
	CU (cudaStreamSynchronize (ctx->stream));
	return 0;
}

// optional dumps for the matrix-exact tests (nyxb200_debug_matrix on a giant ROI)
struct GiantDebug
{
	int family = -1;
	std::vector<uint32_t>* dense = nullptr;		// GLCM joint [4][Ng][Ng], GLRLM P [4][Lm][Nr], GLDM/NGLDM P [L1][10], NGTDM N [L1]
	std::vector<uint32_t>* zone_keys = nullptr;	// GLSZM / GLDZM: (level, minor) per zone
	std::vector<double>* ngtdm_s = nullptr;
	int dim0 = 0, dim1 = 0;
};

int giant_texture_roi (nyxb200_ctx* ctx, const ImageView& im, const RoiRec& r, const Params& prm, int fam_mask, double* orow, GiantDebug* dbg)
{
	cudaStream_t s = ctx->stream;
	const unsigned long long ncell = (unsigned long long)r.w * (unsigned long long)r.h;
	if (ncell > GIANT_MAX_CELLS)
		return fail (ctx, NYXB200_EUNSUPPORTED, "ROI %u: bounding box of %llu pixels exceeds the 2^32 cell limit of the large-ROI path", r.label, ncell);
	const int GB = ctx->sm_count * 4;
	const bool ibsi = prm.ibsi != 0;
	const int Lm = ibsi ? (int)std::max<uint32_t> (r.imax, 1u) : prm.greydepth;	// largest level of this ROI's binned tile
	const int L1 = Lm + 1;
	const bool blank = r.imin == r.imax;
	auto fill = [&](int col, int n, double v) { k_big_fill<<<1, 128, 0, s>>> (orow + col, n, v); ctx->counts[0]++; };
	auto need_smem = [&](auto kern, size_t bytes) -> int { return set_smem (ctx, kern, bytes); };

	// ---- tile(s)
	const bool want_ngldm = (fam_mask & (1 << FAM_NGLDM)) && !blank;
	const bool want_tile = (fam_mask & ~((1 << FAM_NGLDM) | (1 << FAM_INT))) != 0;
	if (want_tile)
		CU (ctx->g_cell.reserve (ncell));
	if (want_ngldm)
		CU (ctx->g_cell2.reserve (ncell));
	k_big_tile<<<GB, BG_T, 0, s>>> (im, r, prm, want_tile ? ctx->g_cell.p : nullptr, want_ngldm ? ctx->g_cell2.p : nullptr);
	CU (cudaGetLastError());
	ctx->counts[0]++;
	const uint32_t* cell = ctx->g_cell.p;
	CU (ctx->g_acc.reserve (BG_NACC * 2));
	CU (ctx->g_part.reserve ((size_t)GB * BG_NPART));

	// ---- GLCM
	if (fam_mask & (1 << FAM_GLCM))
	{
		Binner Bm;
		Bm.init (prm.greydepth, r.imin, r.imax);
		if (Bm.bin (r.imin) == Bm.bin (r.imax))		// glcm.cpp:27-95
			fill (COL_GLCM, NCOL_GLCM, prm.softnan);
		else
		{
			const int Ng = Lm;
			if (Ng > 4096)
				return fail (ctx, NYXB200_EUNSUPPORTED, "ROI %u: the large-ROI GLCM path handles at most 4096 grey levels (have %d); use coarse_gray_depth binning", r.label, Ng);
			const size_t nn = (size_t)Ng * Ng;
			CU (ctx->g_mat.reserve (4 * nn));
			CU (cudaMemsetAsync (ctx->g_mat.p, 0, 4 * nn * 4, s));
			int bw = (int)std::min<size_t> ((size_t)2 * Ng - 1, BG_SMEM_WORDS / (size_t)Ng);
			if (!(bw & 1))
				bw--;
			if (bw < 1)
				bw = 1;
			if (need_smem (k_big_glcm_build, (size_t)Ng * bw * 4))
				return NYXB200_ECUDA;
			k_big_glcm_build<<<GB, BG_T, (size_t)Ng * bw * 4, s>>> (cell, r.w, r.h, Ng, prm.glcm_offset, (prm.glcm_symmetric || ibsi) ? 1 : 0, bw, ctx->g_mat.p);
			CU (cudaGetLastError());
			const size_t rw = big_glcm_reduce_words (Ng) * 4;
			if (need_smem (k_big_glcm_reduce, rw))
				return NYXB200_ECUDA;
			k_big_glcm_reduce<<<4, BG_T, rw, s>>> (ctx->g_mat.p, Ng, r, prm, orow + COL_GLCM);
			k_big_ave<<<1, 32, 0, s>>> (orow + COL_GLCM, 1);
			CU (cudaGetLastError());
			ctx->counts[0] += 3;
			if (dbg && dbg->family == FAM_GLCM)
			{
				dbg->dense->resize (4 * nn);
				dbg->dim0 = Ng;
				CU (cudaMemcpyAsync (dbg->dense->data(), ctx->g_mat.p, 4 * nn * 4, cudaMemcpyDeviceToHost, s));
				CU (cudaStreamSynchronize (s));
			}
		}
	}

	// ---- GLRLM
	if (fam_mask & (1 << FAM_GLRLM))
	{
		if (blank)	// glrlm.cpp:29-52
			fill (COL_GLRLM, NCOL_GLRLM, prm.softnan);
		else
		{
			const int Nr = std::max (r.w, r.h);
			const size_t pw = (size_t)4 * Lm * Nr;
			if (pw > GIANT_MAX_MATRIX_WORDS || (size_t)Lm + Nr > 56000)
				return fail (ctx, NYXB200_EUNSUPPORTED, "ROI %u: run-length matrix of %d levels x %d run lengths exceeds the large-ROI GLRLM limits", r.label, Lm, Nr);
			CU (ctx->g_mat.reserve (pw));
			CU (cudaMemsetAsync (ctx->g_mat.p, 0, pw * 4, s));
			CU (cudaMemsetAsync (ctx->g_acc.p, 0, BG_NACC * 2 * 8, s));
			const int SL = (size_t)4 * Lm * 8 <= BG_SMEM_WORDS ? 8 : 0;
			if (need_smem (k_big_glrlm_build, (size_t)4 * Lm * SL * 4))
				return NYXB200_ECUDA;
			k_big_glrlm_build<<<GB, BG_T, (size_t)4 * Lm * SL * 4, s>>> (cell, r.w, r.h, Lm, Nr, SL, ctx->g_mat.p, ctx->g_acc.p);
			CU (cudaGetLastError());
			if (need_smem (k_big_glrlm_reduce, ((size_t)Lm + Nr) * 4))
				return NYXB200_ECUDA;
			k_big_glrlm_reduce<<<4, BG_T, ((size_t)Lm + Nr) * 4, s>>> (ctx->g_mat.p, Lm, Nr, ctx->g_acc.p, orow + COL_GLRLM);
			k_big_ave<<<1, 32, 0, s>>> (orow + COL_GLRLM, 0);
			CU (cudaGetLastError());
			ctx->counts[0] += 3;
			if (dbg && dbg->family == FAM_GLRLM)
			{
				dbg->dense->resize (pw);
				dbg->dim0 = Lm;
				dbg->dim1 = Nr;
				CU (cudaMemcpyAsync (dbg->dense->data(), ctx->g_mat.p, pw * 4, cudaMemcpyDeviceToHost, s));
				CU (cudaStreamSynchronize (s));
			}
		}
	}

	// ---- GLSZM
	if (fam_mask & (1 << FAM_GLSZM))
	{
		if (blank)	// glszm.cpp:61-65
			fill (COL_GLSZM, NCOL_GLSZM, prm.softnan);
		else
		{
			CU (ctx->g_auxA.reserve (ncell));
			CU (ctx->g_auxB.reserve (ncell));
			uint32_t* lab = ctx->g_auxA.p;
			uint32_t* zs = ctx->g_auxB.p;
			k_big_iota<<<GB, BG_T, 0, s>>> (lab, ncell);
			CU (cudaMemsetAsync (zs, 0, ncell * 4, s));
			const int R = std::max (4, (r.h + GB - 1) / GB);
			const int nbands = (r.h + R - 1) / R;
			CU (ctx->g_flags.reserve ((size_t)2 * (nbands + 1) + 1));
			uint32_t* dirty[2] = { ctx->g_flags.p, ctx->g_flags.p + nbands + 1 };
			uint32_t* anyf = ctx->g_flags.p + 2 * (nbands + 1);
			CU (cudaMemsetAsync (dirty[0], 0xFF, (size_t)(nbands + 1) * 4, s));
			ctx->counts[0]++;
			for (int sweep = 0; ; sweep++)
			{
				uint32_t* din = dirty[sweep & 1];
				uint32_t* dout = dirty[(sweep + 1) & 1];
				CU (cudaMemsetAsync (dout, 0, (size_t)(nbands + 1) * 4, s));
				CU (cudaMemsetAsync (anyf, 0, 4, s));
				k_big_glszm_pass<<<nbands, BG_T, 0, s>>> (cell, r.w, r.h, R, lab, din, dout, anyf);
				CU (cudaGetLastError());
				ctx->counts[0]++;
				uint32_t any = 0;
				CU (cudaMemcpyAsync (&any, anyf, 4, cudaMemcpyDeviceToHost, s));
				CU (cudaStreamSynchronize (s));
				if (!any)
					break;
				if (sweep > r.h + 4)
					return fail (ctx, NYXB200_ECUDA, "GLSZM zone labelling did not converge");
			}
			CU (cudaMemsetAsync (ctx->g_acc.p, 0, BG_NACC * 2 * 8, s));
			k_big_zone_sizes<<<GB, BG_T, 0, s>>> (cell, lab, zs, ncell, 0, 0, ctx->g_acc.p + BG_NACC);
			const uint32_t slots = glszm_hash_slots (Lm, (uint32_t)std::min<unsigned long long> (ncell, 0xFFFFFFFFull));
			CU (ctx->g_hkeys.reserve (slots));
			CU (ctx->g_hcnts.reserve (slots));
			CU (cudaMemsetAsync (ctx->g_hkeys.p, 0, (size_t)slots * 8, s));
			CU (cudaMemsetAsync (ctx->g_hcnts.p, 0, (size_t)slots * 4, s));
			const size_t dw = big_glszm_dense_words (L1);
			CU (ctx->g_mat.reserve (dw));
			CU (cudaMemsetAsync (ctx->g_mat.p, 0, dw * 4, s));
			const int use_sm = dw <= BG_SMEM_WORDS;
			if (need_smem (k_big_glszm_zones, use_sm ? dw * 4 : 0))
				return NYXB200_ECUDA;
			uint32_t* dk = nullptr;
			if (dbg && dbg->family == FAM_GLSZM)
			{
				CU (ctx->g_auxC.reserve (2 * ncell + 4));
				dk = ctx->g_auxC.p;
				CU (cudaMemsetAsync (dk + 2 * ncell, 0, 4, s));
			}
			k_big_glszm_zones<<<GB, BG_T, use_sm ? dw * 4 : 0, s>>> (cell, lab, zs, ncell, L1, use_sm, ctx->g_mat.p, ctx->g_hkeys.p,
				ctx->g_hcnts.p, slots - 1, ctx->g_acc.p, ctx->g_part.p, dk, dk ? dk + 2 * ncell : nullptr);
			k_big_glszm_reduce<<<1, BG_T, 0, s>>> (ctx->g_mat.p, L1, ctx->g_hkeys.p, ctx->g_hcnts.p, slots, ctx->g_acc.p,
				ctx->g_acc.p + BG_NACC, ctx->g_part.p, GB, ncell, prm, orow + COL_GLSZM);
			CU (cudaGetLastError());
			ctx->counts[0] += 3;
			if (dk)
			{
				uint32_t nk = 0;
				CU (cudaMemcpyAsync (&nk, dk + 2 * ncell, 4, cudaMemcpyDeviceToHost, s));
				CU (cudaStreamSynchronize (s));
				dbg->zone_keys->resize ((size_t)2 * nk);
				CU (cudaMemcpy (dbg->zone_keys->data(), dk, (size_t)2 * nk * 4, cudaMemcpyDeviceToHost));
			}
		}
	}

	// ---- GLDZM
	if (fam_mask & (1 << FAM_GLDZM))
	{
		if (blank)	// gldzm.cpp:266-289
			fill (COL_GLDZM, NCOL_GLDZM, prm.softnan);
		else
		{
			const int Nd = gldzm_nd_bound ((uint32_t)std::min<unsigned long long> (ncell, 0xFFFFFFFFull), std::max (r.w, r.h));
			const size_t pw = (size_t)L1 * Nd;
			if (pw > GIANT_MAX_MATRIX_WORDS || Nd > 56000)
				return fail (ctx, NYXB200_EUNSUPPORTED, "ROI %u: distance-zone matrix of %d levels x %d distances exceeds the large-ROI GLDZM limits", r.label, L1, Nd);
			CU (ctx->g_auxA.reserve (ncell));
			CU (ctx->g_auxB.reserve (ncell));
			uint32_t* parent = ctx->g_auxA.p;
			uint32_t* zmin = ctx->g_auxB.p;
			uint32_t* dmap = nullptr;
			k_big_runheads<<<GB, BG_T, 0, s>>> (cell, r.w, r.h, ibsi, parent, zmin);
			k_big_unite_rows<<<GB, BG_T, 0, s>>> (cell, r.w, r.h, ibsi, parent);
			k_big_flatten<<<GB, BG_T, 0, s>>> (parent, ncell);
			ctx->counts[0] += 3;
			if (ibsi)
			{
				CU (ctx->g_auxC.reserve (2 * ncell + 4));
				dmap = ctx->g_auxC.p;
				k_big_dist_rows<<<GB, BG_T, 0, s>>> (cell, r.w, r.h, dmap);
				k_big_dist_cols<<<std::max (1, std::min (GB, (r.w + BG_T - 1) / BG_T)), BG_T, 0, s>>> (cell, r.w, r.h, dmap);
				ctx->counts[0] += 2;
			}
			k_big_zone_metric<<<GB, BG_T, 0, s>>> (cell, parent, dmap, r.w, r.h, ibsi, zmin);
			CU (cudaGetLastError());
			CU (ctx->g_mat.reserve (pw));
			CU (cudaMemsetAsync (ctx->g_mat.p, 0, pw * 4, s));
			CU (cudaMemsetAsync (ctx->g_acc.p, 0, BG_NACC * 2 * 8, s));
			const int use_sm = (size_t)L1 * GD_SD <= BG_SMEM_WORDS;
			if (need_smem (k_big_gldzm_zones, use_sm ? (size_t)L1 * GD_SD * 4 : 0))
				return NYXB200_ECUDA;
			uint32_t* dk = nullptr;
			if (dbg && dbg->family == FAM_GLDZM)
			{
				CU (ctx->g_cell2.reserve (2 * ncell + 4));	// (free in a debug call: NGLDM is not selected there)
				dk = ctx->g_cell2.p;
				CU (cudaMemsetAsync (dk + 2 * ncell, 0, 4, s));
			}
			k_big_gldzm_zones<<<GB, BG_T, use_sm ? (size_t)L1 * GD_SD * 4 : 0, s>>> (cell, parent, zmin, ncell, ibsi, L1, Nd, use_sm,
				ctx->g_mat.p, ctx->g_acc.p, ctx->g_part.p, dk, dk ? dk + 2 * ncell : nullptr);
			if (need_smem (k_big_gldzm_reduce, (size_t)Nd * 4))
				return NYXB200_ECUDA;
			k_big_gldzm_reduce<<<1, BG_T, (size_t)Nd * 4, s>>> (ctx->g_mat.p, L1, Nd, ctx->g_acc.p, ctx->g_part.p, GB, r, prm, orow + COL_GLDZM);
			CU (cudaGetLastError());
			ctx->counts[0] += 3;
			if (dk)
			{
				uint32_t nk = 0;
				CU (cudaMemcpyAsync (&nk, dk + 2 * ncell, 4, cudaMemcpyDeviceToHost, s));
				CU (cudaStreamSynchronize (s));
				dbg->zone_keys->resize ((size_t)2 * nk);
				CU (cudaMemcpy (dbg->zone_keys->data(), dk, (size_t)2 * nk * 4, cudaMemcpyDeviceToHost));
			}
		}
	}

	// ---- GLDM / NGLDM
	for (int fam : { (int)FAM_GLDM, (int)FAM_NGLDM })
	{
		if (!(fam_mask & (1 << fam)))
			continue;
		const int col = fam == FAM_GLDM ? COL_GLDM : COL_NGLDM, ncol = fam == FAM_GLDM ? NCOL_GLDM : NCOL_NGLDM;
		if (blank)	// gldm.cpp:22-40, ngldm.cpp:163-185
		{
			fill (col, ncol, prm.softnan);
			continue;
		}
		const int L1f = fam == FAM_NGLDM ? (ibsi ? (int)r.imax : prm.greydepth) + 1 : L1;
		const size_t pw = (size_t)L1f * 10;
		CU (ctx->g_mat.reserve (pw));
		CU (cudaMemsetAsync (ctx->g_mat.p, 0, pw * 4, s));
		const int use_sm = pw <= BG_SMEM_WORDS;
		if (need_smem (k_big_dep_build, use_sm ? pw * 4 : 0))
			return NYXB200_ECUDA;
		k_big_dep_build<<<GB, BG_T, use_sm ? pw * 4 : 0, s>>> (fam == FAM_NGLDM ? ctx->g_cell2.p : cell, r.w, r.h, L1f, fam == FAM_NGLDM, use_sm, ctx->g_mat.p);
		if (fam == FAM_GLDM)
			k_big_gldm_reduce<<<1, BG_T, 0, s>>> (ctx->g_mat.p, L1f, prm, orow + col);
		else
			k_big_ngldm_reduce<<<1, BG_T, 0, s>>> (ctx->g_mat.p, L1f, prm, orow + col);
		CU (cudaGetLastError());
		ctx->counts[0] += 2;
		if (dbg && dbg->family == fam)
		{
			dbg->dense->resize (pw);
			dbg->dim0 = L1f;
			CU (cudaMemcpyAsync (dbg->dense->data(), ctx->g_mat.p, pw * 4, cudaMemcpyDeviceToHost, s));
			CU (cudaStreamSynchronize (s));
		}
	}

	// ---- NGTDM
	if (fam_mask & (1 << FAM_NGTDM))
	{
		// g_mat: N[L1] | PR[L1] | lst[L1] | pad | SI[L1][8] (u64) | Pl[L1] | Sl[L1] (f64)
		size_t w32 = (size_t)3 * L1;
		w32 += w32 & 1;
		const size_t total = w32 + (size_t)L1 * 16 + (size_t)L1 * 4;
		CU (ctx->g_mat.reserve (total));
		CU (cudaMemsetAsync (ctx->g_mat.p, 0, total * 4, s));
		CU (cudaMemsetAsync (ctx->g_acc.p, 0, BG_NACC * 2 * 8, s));
		uint32_t* Nn = ctx->g_mat.p;
		uint32_t* PR = Nn + L1;
		uint32_t* lst = PR + L1;
		unsigned long long* SI = (unsigned long long*)(ctx->g_mat.p + w32);
		double* Pl = (double*)(SI + (size_t)L1 * 8);
		double* Sl = Pl + L1;
		const size_t smw = (size_t)L1 * 16 + (size_t)2 * L1;
		const int use_sm = smw <= BG_SMEM_WORDS;
		if (need_smem (k_big_ngtdm_build, use_sm ? smw * 4 : 0))
			return NYXB200_ECUDA;
		k_big_ngtdm_build<<<GB, BG_T, use_sm ? smw * 4 : 0, s>>> (cell, r.w, r.h, L1, use_sm, Nn, SI, PR, ctx->g_acc.p);
		k_big_ngtdm_reduce<<<1, BG_T, 0, s>>> (Nn, SI, PR, L1, ctx->g_acc.p, ibsi, prm, lst, Pl, Sl, orow + COL_NGTDM);
		CU (cudaGetLastError());
		ctx->counts[0] += 2;
		if (dbg && dbg->family == FAM_NGTDM)
		{
			dbg->dense->resize ((size_t)2 * L1);	// N[L1] then PR[L1]
			dbg->dim0 = L1;
			std::vector<unsigned long long> si ((size_t)L1 * 8);
			CU (cudaMemcpyAsync (dbg->dense->data(), Nn, (size_t)2 * L1 * 4, cudaMemcpyDeviceToHost, s));
			CU (cudaMemcpyAsync (si.data(), SI, (size_t)L1 * 64, cudaMemcpyDeviceToHost, s));
			CU (cudaStreamSynchronize (s));
			dbg->ngtdm_s->assign (L1, 0.0);
			for (int lv = 0; lv < L1; lv++)
				for (int k = 0; k < 8; k++)
					if (si[(size_t)lv * 8 + k])
						(*dbg->ngtdm_s)[lv] += (double)si[(size_t)lv * 8 + k] / (double)(k + 1);
		}
	}
	return 0;
}

int giant_intensity_roi (nyxb200_ctx* ctx, const ImageView& im, const RoiRec& r, const Params& prm, double* orow)
{
	cudaStream_t s = ctx->stream;
	const unsigned long long nbins = (unsigned long long)(r.imax - r.imin) + 1ull;
	if (nbins > ((unsigned long long)1 << 26))
		return fail (ctx, NYXB200_EUNSUPPORTED, "ROI %u: intensity range %llu exceeds the 2^26 values the large-ROI first-order path histograms exactly", r.label, nbins);
	const int GB = ctx->sm_count * 4;
	CU (ctx->g_mat.reserve (nbins + 4));
	CU (cudaMemsetAsync (ctx->g_mat.p, 0, nbins * 4, s));
	k_big_int_hist<<<GB, BG_T, 0, s>>> (im, r, ctx->g_mat.p);
	const int nb = prm.greydepth > 0 ? prm.greydepth : -prm.greydepth;
	const size_t smb = (size_t)(101 + nb + 2) * 4;
	if (set_smem (ctx, k_big_int_reduce, smb))
		return NYXB200_ECUDA;
	k_big_int_reduce<<<1, BG_T, smb, s>>> (ctx->g_mat.p, r, prm, orow + COL_INT);
	CU (cudaGetLastError());
	ctx->counts[0] += 2;
	return 0;
}

// all oversize ROIs of the slide (main stream, after the small classes were joined)
int process_giant_rois (nyxb200_ctx* ctx, const ImageView& im, const SlideInfo& info, const Params& prm, int ncol_all)
{
	std::vector<RoiRec> recs;
	std::vector<uint32_t> idx;
	const int tex_mask = ctx->fam_mask & ~(1 << FAM_INT);
	if (tex_mask && ctx->giant_tcount)
	{
		if (int rc = fetch_class_rois (ctx, ctx->d_tlist.p + ctx->giant_tbase, ctx->giant_tcount, recs, idx))
			return rc;
		for (size_t k = 0; k < recs.size(); k++)
			if (int rc = giant_texture_roi (ctx, im, recs[k], prm, tex_mask, ctx->d_full.p + (size_t)idx[k] * ncol_all, nullptr))
				return rc;
		ctx->giant_rois += (int64_t)recs.size();
	}
	if (((ctx->fam_mask & (1 << FAM_INT)) || ctx->want_hist) && ctx->giant_acount)
	{
		if (int rc = fetch_class_rois (ctx, ctx->d_alist.p + ctx->giant_abase, ctx->giant_acount, recs, idx))
			return rc;
		for (size_t k = 0; k < recs.size(); k++)
			if (int rc = giant_intensity_roi (ctx, im, recs[k], prm, ctx->d_full.p + (size_t)idx[k] * ncol_all))
				return rc;
	}
	return 0;
}

// stage st runs from this event to its stage_end (or, on the main stream, to the next stage's begin)
void stage_end (nyxb200_ctx* ctx, int st, cudaStream_t s = nullptr)
{
	const int e = ctx->cur_b * nyxb200_ctx::NEV + st;
	cudaEventRecord (ctx->ev_end[e], s ? s : ctx->stream);
	ctx->ev_end_used[e] = true;
	if (!s)
		ctx->open_stage = -1;
}
void stage_begin (nyxb200_ctx* ctx, int st, cudaStream_t s = nullptr)
{
	if (!s && ctx->open_stage >= 0)		// main-stream stages are back to back
		stage_end (ctx, ctx->open_stage);
	const int e = ctx->cur_b * nyxb200_ctx::NEV + st;
	cudaEventRecord (ctx->ev[e], s ? s : ctx->stream);
	ctx->ev_used[e] = true;
	if (!s)
		ctx->open_stage = st;
}

// fork: family streams wait for everything submitted to the main stream so far
int fork_families (nyxb200_ctx* ctx)
{
	if (!ctx->concurrent)
		return 0;
	CU (cudaEventRecord (ctx->ev_fork, ctx->stream));
	for (auto& fs : ctx->fstream)
		CU (cudaStreamWaitEvent (fs, ctx->ev_fork, 0));
	return 0;
}
// family `slot` submits to its own stream (or to the main stream when serialised)
void use_family_stream (nyxb200_ctx* ctx, int slot)
{
	ctx->lsi = slot;
	ctx->ls = ctx->concurrent ? ctx->fstream[slot] : ctx->stream;
}
int join_families (nyxb200_ctx* ctx)
{
	ctx->ls = ctx->stream;
	ctx->lsi = 0;
	if (!ctx->concurrent)
		return 0;
	for (int i = 0; i < nyxb200_ctx::NFSTREAM; i++)
	{
		CU (cudaEventRecord (ctx->ev_join[i], ctx->fstream[i]));
		CU (cudaStreamWaitEvent (ctx->stream, ctx->ev_join[i], 0));
	}
	return 0;
}

// grow-only pinned host buffer for the result table (keeps its content when it grows); buffers come from the pool
int reserve_pinned_out (nyxb200_ctx* ctx, size_t n)
{
	if (n <= ctx->h_out_cap)
		return 0;
	const size_t want = ctx->h_out ? n + n / 2 + 1024 : n + 1024;
	size_t got = 0;
	double* p = (double*)pool_acquire (want * sizeof (double), &got);
	if (!p)
		return fail (ctx, NYXB200_ENOMEM, "cannot allocate %zu bytes of pinned host memory for the result table", want * sizeof (double));
	if (ctx->h_out)
	{
		memcpy (p, ctx->h_out, ctx->h_out_used * sizeof (double));
		pool_release (ctx->h_out);
	}
	ctx->h_out = p;
	ctx->h_out_cap = got / sizeof (double);
	return 0;
}

// phase-1 scan launch: a grid of exactly the resident CTAs; rows per unit as large as keeps >= 4 units per resident warp
// (long units = more runs merged per commit by the per-lane cache)
template <bool U16, bool VEC, bool D3>
int launch_scan (nyxb200_ctx* ctx, const ImageView& im, uint32_t nrows, uint32_t nchunk, uint2* ztab)
{
	auto kern = k_scan_stats<U16, VEC, D3>;
	const size_t smem = VEC ? scan_smem_bytes<U16>() : 0;	// the cp.async row ring
	if (set_smem (ctx, kern, smem))
		return NYXB200_ECUDA;
	int per_sm = 0;
	CU (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, kern, 256, smem));
	per_sm = std::max (per_sm, 1);
	const unsigned long long resident_warps = (unsigned long long)ctx->sm_count * per_sm * 8;
	uint32_t rpu = 32;
	while (rpu > 1 && (unsigned long long)((nrows + rpu - 1) / rpu) * nchunk < 4 * resident_warps)
		rpu >>= 1;
	const unsigned long long nunits = (unsigned long long)((nrows + rpu - 1) / rpu) * nchunk;
	const int grid = (int)std::max<unsigned long long> (1, std::min<unsigned long long> ((nunits + 7) / 8, (unsigned long long)ctx->sm_count * per_sm));
	kern<<<grid, 256, smem, ctx->stream>>> (im.inten, im.label, (uint32_t)im.W, nrows, (uint32_t)im.H, nchunk, rpu, ctx->d_stats.p,
		ctx->stats_cap, ctx->d_info.p, ztab);
	CU (cudaGetLastError());
	return 0;
}

// One slide already resident on the device.  Appends its rows to the context's pinned table buffer.
int process_slide_body (nyxb200_ctx* ctx, const ImageView& im, int img_index, int64_t* nrows_out);

// Failure funnel: an error return in the middle of a slide (unsupported configuration found after the fork, a CUDA
// failure) must not leave kernels in flight on the family / copy streams or the launch-stream selection pointing at
// a family stream -- the caller's buffers may be released as soon as the call returns.
int process_slide (nyxb200_ctx* ctx, const ImageView& im, int img_index, int64_t* nrows_out)
{
	const int rc = process_slide_body (ctx, im, img_index, nrows_out);
	if (rc)
	{
		const std::string msg = ctx->err;
		for (auto& fs : ctx->fstream)
			cudaStreamSynchronize (fs);
		cudaStreamSynchronize (ctx->copy_stream);
		cudaStreamSynchronize (ctx->stream);
		cudaGetLastError();
		ctx->ls = ctx->stream;
		ctx->lsi = 0;
		ctx->open_stage = -1;
		ctx->cur_b = 0;
		ctx->err = msg;
		if (SharedOut* sh = ctx->share)
		{
			std::lock_guard<std::mutex> lk (sh->m);
			sh->failed = true;
			sh->cv.notify_all();
		}
	}
	return rc;
}

int process_slide_body (nyxb200_ctx* ctx, const ImageView& im_in, int img_index, int64_t* nrows_out)
{
	ImageView im = im_in;	// (the mask pointer is replaced when the labels have to be made dense)
	ctx->relabelled = false;
	const unsigned long long npix = (unsigned long long)im.D * (unsigned long long)im.H * (unsigned long long)im.W;
	const bool d3 = ctx->st.dim == 3;
	// 2D: the canonical row ends after the last selected family (the texture + first-order block is always there; the
	// appended morphology / IH / moment blocks cost their 1.4 KB per ROI of memset and row stride only when selected)
	if (!d3)
		ctx->row_stride = ctx->want_hist ? COL_HIST + ctx->st.coarse_gray_depth : ctx->wmom_variants ? NCOL_ALL : ctx->geom_variants ? COL_GEOM + NCOL_GEOMOM : ctx->want_ih ? COL_IH + NCOL_IH : ctx->want_morph ? COL_MORPH + NCOL_MORPH : NCOL_TEXINT;
	const int ncol_all = d3 ? NCOL3_ALL : ctx->row_stride;
	Params prm = make_params (ctx);
	cudaStream_t s = ctx->stream;
	ctx->cur_b = 0;
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
		const bool vec = (im.W % SCAN_PX == 0) && ((uintptr_t)im.label % 16 == 0) && ((uintptr_t)im.inten % 16 == 0);
		const uint32_t nrows = (uint32_t)im.D * (uint32_t)im.H, nchunk = ((uint32_t)im.W + SCAN_CHUNK - 1) / SCAN_CHUNK;
#define NYX_SCAN(U16, VEC, D3) rc_scan = launch_scan<U16, VEC, D3> (ctx, im, nrows, nchunk, ztab)
		int rc_scan = 0;
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
		if (rc_scan)
			return rc_scan;
		CU (cudaGetLastError());
		ctx->counts[0] += 2;
		stage_begin (ctx, ST_COMPACT);
		const uint32_t nblk = (ctx->stats_cap + CB - 1) / CB;
		CU (ctx->d_blkcnt.reserve (nblk));
		CU (ctx->d_blkarea.reserve (nblk));
		CU (ctx->d_blkcells.reserve (nblk));
		k_compact_count<<<nblk, CB, 0, s>>> (ctx->d_stats.p, ctx->stats_cap, ctx->d_info.p, ctx->d_blkcnt.p, ctx->d_blkarea.p,
			ctx->d_blkcells.p, d3 ? 1 : 0);
		k_compact_scan<<<1, CB, 0, s>>> (ctx->d_blkcnt.p, ctx->d_blkarea.p, ctx->d_blkcells.p, nblk, ctx->d_info.p);
		CU (cudaGetLastError());
		ctx->counts[0] += 2;
		CU (cudaMemcpyAsync (&info, ctx->d_info.p, sizeof (info), cudaMemcpyDeviceToHost, s));
		CU (cudaStreamSynchronize (s));
		if (info.max_label >= ctx->stats_cap)
		{
			if (attempt > 0)
				return fail (ctx, NYXB200_ECUDA, "label table overflow persisted");
			if (info.max_label >= RELABEL_MIN)
			{
				// large / sparse label values: ranks 1 .. K instead (ascending label order, so the rows come out in the same order)
				const size_t nwords = ((size_t)info.max_label >> 5) + 1, nblocks = (nwords + RL_BLOCK_WORDS - 1) / RL_BLOCK_WORDS;
				CU (ctx->d_rl_bitmap.reserve (nwords));
				CU (ctx->d_rl_wprefix.reserve (nwords));
				CU (ctx->d_rl_blocks.reserve (nblocks + 1));
				CU (ctx->d_rl_mask.reserve ((size_t)npix));
				CU (cudaMemsetAsync (ctx->d_rl_bitmap.p, 0, nwords * 4, s));
				const int g = ctx->sm_count * 16;
				k_rl_mark<<<g, 256, 0, s>>> (im.label, (size_t)npix, ctx->d_rl_bitmap.p);
				k_rl_block_popc<<<(unsigned)nblocks, 256, 0, s>>> (ctx->d_rl_bitmap.p, nwords, ctx->d_rl_blocks.p);
				k_rl_scan_blocks<<<1, 1024, 0, s>>> (ctx->d_rl_blocks.p, (uint32_t)nblocks, ctx->d_rl_blocks.p + nblocks);
				k_rl_word_prefix<<<(unsigned)nblocks, 32, 0, s>>> (ctx->d_rl_bitmap.p, nwords, ctx->d_rl_blocks.p, ctx->d_rl_wprefix.p);
				uint32_t K = 0;
				CU (cudaMemcpyAsync (&K, ctx->d_rl_blocks.p + nblocks, 4, cudaMemcpyDeviceToHost, s));
				CU (cudaStreamSynchronize (s));
				CU (ctx->d_rl_labmap.reserve (std::max<uint32_t> (K, 1)));
				k_rl_list<<<g, 256, 0, s>>> (ctx->d_rl_bitmap.p, nwords, ctx->d_rl_wprefix.p, ctx->d_rl_labmap.p);
				k_rl_apply<<<g, 256, 0, s>>> (im.label, (size_t)npix, ctx->d_rl_bitmap.p, ctx->d_rl_wprefix.p, ctx->d_rl_mask.p);
				CU (cudaGetLastError());
				ctx->counts[0] += 6;
				if (K >= (1u << 27))
					return fail (ctx, NYXB200_EUNSUPPORTED, "%u distinct labels exceed the label table limit (2^27)", K);
				im.label = ctx->d_rl_mask.p;
				ctx->relabelled = true;
				ctx->stats_cap = std::max<uint32_t> (ctx->stats_cap, K + 1);
				continue;
			}
			ctx->stats_cap = info.max_label + 1;
			continue;
		}
		break;
	}
	const uint32_t nroi = info.n_roi;
	*nrows_out = nroi;
	double* shared_dst = nullptr;
	if (SharedOut* sh = ctx->share)
	{
		// claim rows [row0, row0 + nroi) of the shared table, in slide order
		std::unique_lock<std::mutex> lk (sh->m);
		sh->cv.wait (lk, [&] { return sh->failed || sh->next_slide == img_index; });
		if (sh->failed)
			return fail (ctx, NYXB200_ECUDA, "another device of the multi-device call failed");
		if (!sh->base)
		{
			// capacity from the first slide's ROI count (slides of one job are alike); a slide that does not fit is
			// deferred to a second pass with the exact size
			sh->cap_rows = (size_t)nroi * sh->n_img + (size_t)nroi * sh->n_img / 4 + (size_t)ctx->multi_row_pad;
			size_t got = 0;
			sh->base = (double*)pool_acquire (sh->cap_rows * sh->rowlen * sizeof (double), &got);
			if (!sh->base)
			{
				sh->failed = true;
				sh->cv.notify_all();
				return fail (ctx, NYXB200_ENOMEM, "cannot allocate the pinned result table of the multi-device call");
			}
			sh->cap_rows = got / (sh->rowlen * sizeof (double));
		}
		sh->row0[img_index] = sh->cursor;
		sh->nrows[img_index] = nroi;
		const bool fits = (size_t)(sh->cursor + nroi) <= sh->cap_rows;
		sh->deferred[img_index] = fits ? 0 : 1;
		if (fits)
			shared_dst = sh->base + (size_t)sh->cursor * sh->rowlen;
		sh->cursor += nroi;
		sh->next_slide++;
		lk.unlock();
		sh->cv.notify_all();
		if (!fits)
		{
			stage_begin (ctx, ST_TOTAL);
			ctx->open_stage = -1;
			return 0;
		}
	}
	ctx->counts[1] += nroi;
	ctx->counts[2] += (int64_t)info.n_fg;
	ctx->counts[3] += (int64_t)npix;
	ctx->last_im = im;
	ctx->h_rois_valid = false;
	if (nroi == 0)
	{
		ctx->last_info = info;
		stage_begin (ctx, ST_TOTAL);
		ctx->open_stage = -1;
		return 0;
	}
	if (ctx->st.ibsi && info.max_inten > CELL_LEVEL)
		return fail (ctx, NYXB200_EUNSUPPORTED, "IBSI mode needs intensities < 2^24 (max %u)", info.max_inten);
	if (!d3 && !ctx->st.ibsi && ctx->st.coarse_gray_depth < 0)
	{
		// PyRadiomics bin count as the general grey depth (C ABI only): GLCM, GLRLM, GLSZM, GLDM, NGTDM and first-order
		// follow the reference.  GLDZM: the reference indexes I.end() for the level-0 zones and throws (gldzm.cpp:245-258);
		// NGLDM: to_grayscale() takes the negative depth as a huge unsigned level count (ngldm.cpp:42) -- neither is a
		// result worth reproducing
		if (ctx->fam_mask & ((1 << FAM_GLDZM) | (1 << FAM_NGLDM)))
			return fail (ctx, NYXB200_EUNSUPPORTED, "GLDZM and NGLDM are undefined for a negative (PyRadiomics) grey depth in the reference (gldzm.cpp:245-258 throws, ngldm.cpp:42)");
	}

	// ---- ROI batches: rows of batch b travel to the host while batch b+1 computes (the reference also walks the ROIs in
	// batches, phase2_2d.cpp:427-465 -- there to bound RAM, here to hide the table's D2H copy)
	int nb = ctx->roi_batches > 0 ? ctx->roi_batches : (int)(nroi / 8192);
	nb = std::max (1, std::min (nb, NB_MAX));
	const uint32_t bs = (nroi + nb - 1) / nb;		// ROIs per batch (index ranges of the label-ordered ROI array)
	nb = (int)((nroi + bs - 1) / bs);
	// The copy of the LAST batch's rows is the one transfer nothing hides.  Automatic batching tapers the end of the slide:
	// the last uniform batch [tail_start, nroi) is cut into 2q | q | rest, so the exposed copy carries a quarter of a batch
	uint32_t tail_start = 0xFFFFFFFFu, tail_q = 1;
	if (ctx->roi_batches <= 0 && nb >= 4 && nb + 2 <= NB_MAX && nroi - (uint32_t)(nb - 1) * bs >= 256)
	{
		tail_start = (uint32_t)(nb - 1) * bs;
		tail_q = (nroi - tail_start + 3) / 4;
		nb += 2;
	}
	// first ROI index of batch b (b == nb: the end)
	auto batch_lo = [&] (int b) -> uint32_t {
		if (tail_start == 0xFFFFFFFFu)
			return std::min<uint32_t> (nroi, (uint32_t)b * bs);
		const int nbm = nb - 3;		// uniform batches in front of the tapered tail
		if (b <= nbm)
			return (uint32_t)b * bs;
		return b == nbm + 1 ? std::min (nroi, tail_start + 2 * tail_q) : b == nbm + 2 ? std::min (nroi, tail_start + 3 * tail_q) : nroi;
	};

	// ---- ROI records + class lists per (batch, class)
	CU (ctx->d_rois.reserve (nroi));
	CU (ctx->d_classslot.reserve ((size_t)nroi * 4));
	CU (ctx->d_tlist.reserve (nroi));
	CU (ctx->d_alist.reserve (nroi));
	CU (ctx->d_bases.reserve (NTK + NAK));
	{
		const uint32_t nblk = (ctx->stats_cap + CB - 1) / CB;
		k_compact_write<<<nblk, CB, 0, s>>> (ctx->d_stats.p, ctx->stats_cap, ctx->d_info.p, ctx->d_blkcnt.p, ctx->d_blkarea.p,
			ctx->d_blkcells.p, ctx->d_rois.p, ctx->d_classslot.p, bs, d3 ? ctx->d_ztab.p : nullptr, tail_start, tail_q);
		CU (cudaGetLastError());
		ctx->counts[0]++;
		CU (cudaMemcpyAsync (&info, ctx->d_info.p, sizeof (info), cudaMemcpyDeviceToHost, s));
		CU (cudaStreamSynchronize (s));
		// list layout: the shared-memory classes batch-major, then the oversize class of every batch at the end (one
		// contiguous list for the giant path)
		uint32_t bases[NTK + NAK];
		uint32_t acc = 0;
		for (int bb = 0; bb < NB_MAX; bb++)
			for (int c = 0; c < NCLASS; c++)
			{
				ctx->tbase[bb * (NCLASS + 1) + c] = acc;
				acc += info.class_count[bb * (NCLASS + 1) + c];
			}
		ctx->giant_tbase = acc;
		for (int bb = 0; bb < NB_MAX; bb++)
		{
			ctx->tbase[bb * (NCLASS + 1) + NCLASS] = acc;
			acc += info.class_count[bb * (NCLASS + 1) + NCLASS];
		}
		ctx->giant_tcount = acc - ctx->giant_tbase;
		acc = 0;
		for (int bb = 0; bb < NB_MAX; bb++)
			for (int c = 0; c < NACLASS; c++)
			{
				ctx->abase[bb * (NACLASS + 1) + c] = acc;
				acc += info.aclass_count[bb * (NACLASS + 1) + c];
			}
		ctx->giant_abase = acc;
		for (int bb = 0; bb < NB_MAX; bb++)
		{
			ctx->abase[bb * (NACLASS + 1) + NACLASS] = acc;
			acc += info.aclass_count[bb * (NACLASS + 1) + NACLASS];
		}
		ctx->giant_acount = acc - ctx->giant_abase;
		memcpy (bases, ctx->tbase, sizeof (uint32_t) * NTK);
		memcpy (bases + NTK, ctx->abase, sizeof (uint32_t) * NAK);
		CU (cudaMemcpyAsync (ctx->d_bases.p, bases, sizeof (bases), cudaMemcpyHostToDevice, s));
		k_fill_class_lists<<<(nroi + 255) / 256, 256, 0, s>>> (ctx->d_classslot.p, nroi, ctx->d_bases.p, ctx->d_bases.p + NTK,
			ctx->d_tlist.p, ctx->d_alist.p);
		CU (cudaGetLastError());
		ctx->counts[0]++;
	}
	ctx->last_info = info;
	const bool any_giant = !d3 && (ctx->giant_tcount || ctx->giant_acount);
	if (any_giant && !ctx->st.ibsi && ctx->st.coarse_gray_depth < 0 && (ctx->fam_mask & ~(1 << FAM_INT)))
		return fail (ctx, NYXB200_EUNSUPPORTED, "a negative (PyRadiomics) grey depth is not built for ROIs of the large-ROI path (bounding box > 40000 pixels)");

	// ---- canonical table
	CU (ctx->d_full.reserve ((size_t)nroi * ncol_all));
	CU (cudaMemsetAsync (ctx->d_full.p, 0, (size_t)nroi * ncol_all * sizeof (double), s));
	// COVERED_IMAGE_INTENSITY_RANGE: ROI range over the slide's range.  File-based workflows prescan the slide over its
	// labelled pixels / voxels (slideprops.cpp:83-200, 305-327; 3d_intensity.cpp:56-61) -- the phase-1 scan has exactly
	// those extrema; the in-memory montage path leaves the sentinels (-> -0.0, see intensity_body.h)
	if (d3 || ctx->covered_from_scan)
		prm.covered_den = (double)info.max_inten - (double)(~info.inv_min_inten);

	const int Lmax = ctx->st.ibsi ? (int)std::max<uint32_t> (info.max_inten, 1u) : std::abs (ctx->st.coarse_gray_depth);
	DebugOut nodbg { nullptr, nullptr, nullptr };
	ImageView imf = im;		// the view the family kernels get (with the tile store when it is used)
	const bool want_cloud = (ctx->fam_mask & (1 << FAM_INT)) != 0 || ((ctx->want_ih || ctx->want_hist) && !d3);
	const bool want_tex = !d3 && (ctx->fam_mask & ~(1 << FAM_INT)) != 0;
	// 16-bit tile codes hold 14-bit levels: MATLAB binning up to 16383 levels, IBSI (identity) up to that intensity
	const bool tiles = want_tex && ctx->use_tile_store && info.n_tile_cells > 0 && (uint32_t)Lmax <= TILE16_LEVEL
		&& (uint32_t)std::abs (ctx->st.coarse_gray_depth) <= TILE16_LEVEL;
	const bool tiles_b = tiles && (ctx->fam_mask & (1 << FAM_NGLDM));
	if (want_cloud)
		CU (ctx->d_cloud.reserve ((size_t)info.n_fg + 4));
	if (tiles)
	{
		CU (ctx->d_tiles.reserve ((size_t)info.n_tile_cells * (tiles_b ? 2 : 1)));
		imf.tileA = ctx->d_tiles.p;
		imf.tileB = tiles_b ? ctx->d_tiles.p + info.n_tile_cells : nullptr;
	}
	ctx->last_im = imf;

	// ---- output buffers
	const int nsel = (int)ctx->sel_cols.size();
	if (ctx->sel_dirty)
	{
		CU (ctx->d_sel.reserve (std::max (nsel, 1)));
		if (nsel)
			CU (cudaMemcpyAsync (ctx->d_sel.p, ctx->sel_cols.data(), sizeof (int) * nsel, cudaMemcpyHostToDevice, s));
		ctx->sel_dirty = false;
	}
	const size_t rowlen = (size_t)nsel + 2, outn = (size_t)nroi * rowlen;
	CU (ctx->d_sel_out.reserve (outn));
	double* h_dst = shared_dst;
	if (!h_dst)
	{
		if (int rc = reserve_pinned_out (ctx, ctx->h_out_used + outn))
			return rc;
		h_dst = ctx->h_out + ctx->h_out_used;
		ctx->h_out_used += outn;
	}
	auto select_and_copy = [&](uint32_t r0, uint32_t r1, bool overlap) -> int
	{
		// rows [r0, r1): gather the enabled columns on the main stream, copy them on the copy stream (overlap) or in line
		const size_t n = (size_t)(r1 - r0) * rowlen;
		stage_begin (ctx, ST_SELECT);
		const int grid = (int)std::min<size_t> ((n + 255) / 256, (size_t)ctx->sm_count * 16);
		k_select_columns<<<grid, 256, 0, s>>> (ctx->d_full.p + (size_t)r0 * ncol_all, ncol_all, ctx->d_rois.p + r0, r1 - r0, ctx->d_sel.p, nsel,
			ctx->st.soft_nan, 0.0, ctx->d_sel_out.p + (size_t)r0 * rowlen, ctx->relabelled ? ctx->d_rl_labmap.p : nullptr);
		CU (cudaGetLastError());
		ctx->counts[0]++;
		stage_end (ctx, ST_SELECT);
		cudaStream_t cs = overlap ? ctx->copy_stream : s;
		if (overlap)
		{
			CU (cudaEventRecord (ctx->ev_sel, s));
			CU (cudaStreamWaitEvent (cs, ctx->ev_sel, 0));
		}
		stage_begin (ctx, ST_D2H, cs);
		CU (cudaMemcpyAsync (h_dst + (size_t)r0 * rowlen, ctx->d_sel_out.p + (size_t)r0 * rowlen, n * sizeof (double), cudaMemcpyDeviceToHost, cs));
		stage_end (ctx, ST_D2H, cs);
		return 0;
	};

	static const int fam_stage[FAM_COUNT] = { ST_INT, ST_GLCM, ST_GLRLM, ST_GLDZM, ST_GLSZM, ST_GLDM, ST_NGLDM, ST_NGTDM };

	// ---- K2: CSR cloud + binned-tile store in one pass over the AABBs of ALL ROIs (the class lists of the batches are
	// contiguous; main stream: every family depends on it)
	ctx->cur_b = 0;
	stage_begin (ctx, ST_PACK);
	{
		const uint32_t nsmall = ctx->giant_tbase - ctx->tbase[0];
		if (nsmall && (want_cloud || tiles))
		{
			if (d3)
				k_pack_cloud<<<(nsmall * 32 + 255) / 256, 256, 0, s>>> (im, ctx->d_rois.p, ctx->d_tlist.p + ctx->tbase[0], nsmall, ctx->d_cloud.p);
			else
				k_pack_tiles<<<(nsmall * 32 + 255) / 256, 256, 0, s>>> (im, ctx->d_rois.p, ctx->d_tlist.p + ctx->tbase[0], nsmall, prm,
					want_cloud ? ctx->d_cloud.p : nullptr, (uint16_t*)imf.tileA, (uint16_t*)imf.tileB);
			CU (cudaGetLastError());
			ctx->counts[0]++;
		}
		const uint32_t nlarge = ctx->giant_tcount;
		if (nlarge && want_cloud)
		{
			const int grid = (int)std::min<uint32_t> (nlarge, (uint32_t)ctx->sm_count * 4);
			const size_t stride = (size_t)im.H * (size_t)im.D + 2;
			CU (ctx->d_rowoff.reserve ((size_t)grid * stride));
			k_pack_cloud_large<<<grid, 256, 0, s>>> (im, ctx->d_rois.p, ctx->d_tlist.p + ctx->giant_tbase, nlarge, ctx->d_cloud.p,
				ctx->d_rowoff.p, stride);
			CU (cudaGetLastError());
			ctx->counts[0]++;
		}
	}
	// ---- families: ONE fork onto the family streams (slot 0 = first-order, then one slot per texture family); the
	// batches follow each other on every family stream without waiting for the main stream, which only joins them
	// batch by batch to select and ship the finished rows
	stage_begin (ctx, ST_FAMILIES);
	ctx->open_stage = -1;
	if (int rc = fork_families (ctx))
		return rc;
	for (int bb = 0; bb < nb; bb++)
	{
		ctx->cur_b = bb;
		const uint32_t* ccount = info.class_count + bb * (NCLASS + 1);
		const uint32_t* cmaxcells = info.class_maxcells + bb * (NCLASS + 1);
		const uint32_t* cmaxwh = info.class_maxwh + bb * (NCLASS + 1);
		const uint32_t* acount = info.aclass_count + bb * (NACLASS + 1);
		const uint32_t* amaxarea = info.aclass_maxarea + bb * (NACLASS + 1);
		const uint32_t* tb = ctx->tbase + bb * (NCLASS + 1);
		const uint32_t* ab = ctx->abase + bb * (NACLASS + 1);

		use_family_stream (ctx, 0);
		stage_begin (ctx, ST_INT, ctx->ls);
		if (want_cloud)
		{
			for (int c = 0; c < NACLASS; c++)
				if (int rc = launch_intensity (ctx, ctx->d_alist.p + ab[c], acount[c], prm, amaxarea[c], ctx->d_full.p, ncol_all))
					return rc;
			// clouds above 32768 pixels: 3D keeps the CTA-per-ROI kernel (global scratch); 2D goes through the giant path below
			if (d3 && acount[NACLASS])
				if (int rc = launch_intensity (ctx, ctx->d_alist.p + ab[NACLASS], acount[NACLASS], prm, amaxarea[NACLASS], ctx->d_full.p, ncol_all))
					return rc;
		}
		if (ctx->want_morph && !d3)
		{
			// basic morphology rides on the first-order stream: warp per ROI for the shared-memory size classes of this batch
			// (their lists are contiguous), CTA per ROI for the oversize class
			uint32_t nsmall = 0;
			for (int c = 0; c < NCLASS; c++)
				nsmall += ccount[c];
			if (nsmall)
			{
				k_roi_morph_w<<<(nsmall * 32 + 255) / 256, 256, 0, ctx->ls>>> (im, ctx->d_rois.p, ctx->d_tlist.p + tb[0], nsmall, prm, ctx->d_full.p, ncol_all);
				CU (cudaGetLastError());
				ctx->counts[0]++;
			}
			if (ccount[NCLASS])
			{
				const int grid = (int)std::min<uint32_t> (ccount[NCLASS], (uint32_t)ctx->sm_count * 4);
				k_roi_morph<<<grid, ROI_THREADS, RED_WORDS * 4, ctx->ls>>> (im, ctx->d_rois.p, ctx->d_tlist.p + tb[NCLASS], ccount[NCLASS], prm, ctx->d_full.p, ncol_all);
				CU (cudaGetLastError());
				ctx->counts[0]++;
			}
		}
		if (ctx->geom_variants && !d3)
		{
			uint32_t nsmall = 0;
			for (int c = 0; c < NCLASS; c++)
				nsmall += ccount[c];
			if (nsmall)
			{
				k_roi_geom_w<<<(nsmall * 32 + 255) / 256, 256, 0, ctx->ls>>> (im, ctx->d_rois.p, ctx->d_tlist.p + tb[0], nsmall, ctx->geom_variants, ctx->d_full.p, ncol_all);
				CU (cudaGetLastError());
				ctx->counts[0]++;
			}
			if (ccount[NCLASS])
			{
				const int grid = (int)std::min<uint32_t> (ccount[NCLASS], (uint32_t)ctx->sm_count * 4);
				k_roi_geom<<<grid, ROI_THREADS, RED_WORDS * 4, ctx->ls>>> (im, ctx->d_rois.p, ctx->d_tlist.p + tb[NCLASS], ccount[NCLASS], ctx->geom_variants, ctx->d_full.p, ncol_all);
				CU (cudaGetLastError());
				ctx->counts[0]++;
			}
		}
		if (ctx->wmom_variants && !d3)
		{
			// whole-slide mode weighs by the distance to the four slide edges instead of a traced contour (2d_geomoments_basic.cpp:
			// 59-96, contour.cpp:914-931); that variant is reachable through TIFF input only and has no golden here: refused
			if (ctx->single_roi)
				return fail (ctx, NYXB200_EUNSUPPORTED, "the distance-to-contour weighted moments are not built for whole-slide (single ROI) mode");
			// distance-to-contour weighted moments: warp per ROI, one launch per size class (the scratch slot of a warp is sized
			// for the class); CTA per ROI for the oversize class
			for (int c = 0; c <= NCLASS; c++)
			{
				const uint32_t n = ccount[c];
				if (!n)
					continue;
				const unsigned long long mc = cmaxcells[c];
				const int mwh = (int)cmaxwh[c];
				if (mwh > 65532)
					return fail (ctx, NYXB200_EUNSUPPORTED, "weighted geometric moments: a ROI edge of %d pixels exceeds 65532", mwh);
				const size_t slotw = (wmom_scratch_words (mc, mwh) + 3) & ~(size_t)3;
				const bool big = c == NCLASS;
				size_t freeb = 0, totalb = 0;
				CU (cudaMemGetInfo (&freeb, &totalb));
				const size_t budget = freeb / 2 + ctx->d_wmom.cap * 4;
				if (slotw * 4 > budget)
					return fail (ctx, NYXB200_ENOMEM, "weighted geometric moments: a ROI of %llu cells needs %zu MB of work space", mc, slotw * 4 >> 20);
				size_t slots = big ? std::min<size_t> (n, (size_t)ctx->sm_count) : std::min<size_t> (((size_t)n + ROI_WARPS - 1) / ROI_WARPS * ROI_WARPS, (size_t)ctx->sm_count * 16);
				slots = std::max<size_t> (big ? 1 : ROI_WARPS, std::min<size_t> (slots, budget / (slotw * 4)));
				if (!big)
					slots = std::max<size_t> (ROI_WARPS, slots / ROI_WARPS * ROI_WARPS);
				CU (ctx->d_wmom.reserve (slots * slotw));
				if (big)
					k_roi_wmom<<<(int)slots, ROI_THREADS, RED_WORDS * 4, ctx->ls>>> (im, ctx->d_rois.p, ctx->d_tlist.p + tb[c], n, ctx->wmom_variants, ctx->d_wmom.p, slotw, mc, mwh, ctx->d_full.p, ncol_all);
				else
					k_roi_wmom_w<<<(int)(slots / ROI_WARPS), ROI_THREADS, 0, ctx->ls>>> (im, ctx->d_rois.p, ctx->d_tlist.p + tb[c], n, ctx->wmom_variants, ctx->d_wmom.p, slotw, (uint32_t)mc, mwh, ctx->d_full.p, ncol_all);
				CU (cudaGetLastError());
				ctx->counts[0]++;
			}
		}
		if (ctx->want_ih && !d3)
		{
			// the IBSI intensity-histogram family reads the same CSR slices: warp per ROI for the area classes of this batch
			// (contiguous lists), CTA per ROI for clouds above 32768 pixels
			uint32_t nsmall = 0;
			for (int c = 0; c < NACLASS; c++)
				nsmall += acount[c];
			const size_t memw = (ih_smem_words (prm.greydepth) + 3) & ~(size_t)3;
			if (nsmall)
			{
				const size_t bytes = memw * 4 * ROI_WARPS;
				if (set_smem (ctx, k_roi_ih_w, bytes))
					return NYXB200_ECUDA;
				const int grid = (int)std::min<uint32_t> ((nsmall + ROI_WARPS - 1) / ROI_WARPS, (uint32_t)ctx->sm_count * 8);
				k_roi_ih_w<<<grid, ROI_THREADS, bytes, ctx->ls>>> (ctx->d_rois.p, ctx->d_alist.p + ab[0], nsmall, prm, ctx->d_cloud.p, ctx->d_full.p, ncol_all, memw);
				CU (cudaGetLastError());
				ctx->counts[0]++;
			}
			if (acount[NACLASS])
			{
				const size_t bytes = (RED_WORDS + memw) * 4;
				if (set_smem (ctx, k_roi_ih, bytes))
					return NYXB200_ECUDA;
				const int grid = (int)std::min<uint32_t> (acount[NACLASS], (uint32_t)ctx->sm_count * 4);
				k_roi_ih<<<grid, ROI_THREADS, bytes, ctx->ls>>> (ctx->d_rois.p, ctx->d_alist.p + ab[NACLASS], acount[NACLASS], prm, ctx->d_cloud.p, ctx->d_full.p, ncol_all);
				CU (cudaGetLastError());
				ctx->counts[0]++;
			}
		}
		stage_end (ctx, ST_INT, ctx->ls);
		if (d3)
		{
			const int maxI = (int)std::max<uint32_t> (info.max_inten, 1u);
			for (int fam = FAM3_GLCM; fam < FAM3_COUNT; fam++)
			{
				static const int st3[FAM3_COUNT] = { ST_INT, ST_GLCM, ST_GLSZM, ST_NGTDM, ST_GLDM, ST_GLRLM, ST_GLDZM, ST_NGLDM };
				if (!(ctx->fam_mask & (1 << fam)))
					continue;
				use_family_stream (ctx, fam);
				stage_begin (ctx, st3[fam], ctx->ls);
				// grey-binning info per family (3d_glcm.cpp:53-55, 3d_glszm.cpp:490-492, 3d_ngtdm.cpp:195-197, 3d_gldm.cpp:93-97,
				// 3d_glrlm.cpp:148-152: per-family depths, 0 = no binning; 3d_gldzm.cpp:65-72 and 3d_ngldm.cpp:204: the coarse depth)
				int g = fam == FAM3_GLCM ? prm.glcm_greydepth : fam == FAM3_GLSZM ? ctx->glszm3_greydepth : fam == FAM3_GLDM ? ctx->gldm3_greydepth
					: fam == FAM3_GLRLM ? ctx->glrlm3_greydepth : fam == FAM3_NGTDM ? ctx->ngtdm3_greydepth : prm.greydepth;
				if (ctx->st.ibsi)
					g = 0;
				int Lm = g != 0 ? std::abs (g) : maxI;	// > 0 MATLAB levels, < 0 PyRadiomics bins (texture_feature.h:52-77), 0 = intensities as they are
				if (fam == FAM3_NGLDM)
				{
					// to_grayscale over [0, aux_max]: levels 0..depth (IBSI: the intensities themselves)
					g = ctx->st.ibsi ? 0 : prm.greydepth;
					if (g < 0)
						return fail (ctx, NYXB200_EUNSUPPORTED, "3D NGLDM needs a positive grey depth");
					Lm = ctx->st.ibsi ? maxI : g;
				}
				if (fam == FAM3_GLDZM && g < 0)
					return fail (ctx, NYXB200_EUNSUPPORTED, "3D GLDZM with a negative (PyRadiomics) grey depth is undefined in the reference "
						"(level-0 zones index its matrix out of range, 3d_gldzm.cpp:294-297)");
				if (fam == FAM3_NGTDM && ctx->ngtdm3_radius == 0)
					Lm = 1;		// nothing is accumulated (see ngtdm3d_body)
				for (int c = 0; c <= NCLASS; c++)
				{
					const uint32_t n = ccount[c], mc = cmaxcells[c];
					if (!n)
						continue;
					const bool big = c == NCLASS;	// cube beyond the shared-memory classes: global-memory path
					const uint32_t* lst = ctx->d_tlist.p + tb[c];
					int rc = 0;
					if (fam == FAM3_GLCM)
					{
						if (g == 0 && Lm > 4096)
							return fail (ctx, NYXB200_EUNSUPPORTED, "3D GLCM without grey binning needs intensities <= 4096 (max %d)", Lm);
						rc = big ? launch_family3_big<FAM3_GLCM> (ctx, im, lst, n, prm, g, Lm, mc, ctx->d_full.p)
							: mc < 32768 ? launch_family3_t<FAM3_GLCM, true> (ctx, im, lst, n, prm, g, Lm, mc, ctx->d_full.p)
							: launch_family3_t<FAM3_GLCM, false> (ctx, im, lst, n, prm, g, Lm, mc, ctx->d_full.p);
					}
					else if (fam == FAM3_GLSZM)
					{
						if (Lm > 65535)
							return fail (ctx, NYXB200_EUNSUPPORTED, "3D GLSZM needs grey levels < 65536 (max %d)", Lm);
						rc = big ? launch_family3_big<FAM3_GLSZM> (ctx, im, lst, n, prm, g, Lm, mc, ctx->d_full.p)
							: launch_family3_t<FAM3_GLSZM, false> (ctx, im, lst, n, prm, g, Lm, mc, ctx->d_full.p);
					}
					else if (fam == FAM3_GLRLM || fam == FAM3_GLDZM || fam == FAM3_NGLDM)
					{
						if (Lm > 65535)
							return fail (ctx, NYXB200_EUNSUPPORTED, "3D GLRLM / GLDZM / NGLDM need grey levels < 65536 (max %d)", Lm);
						Params p3 = prm;
						p3.cube_maxdim = (int)cmaxwh[c];
						if (fam == FAM3_GLRLM)
							rc = big ? launch_family3_big<FAM3_GLRLM> (ctx, im, lst, n, p3, g, Lm, mc, ctx->d_full.p)
								: launch_family3_t<FAM3_GLRLM, false> (ctx, im, lst, n, p3, g, Lm, mc, ctx->d_full.p);
						else if (fam == FAM3_GLDZM)
							rc = big ? launch_family3_big<FAM3_GLDZM> (ctx, im, lst, n, p3, g, Lm, mc, ctx->d_full.p)
								: launch_family3_t<FAM3_GLDZM, false> (ctx, im, lst, n, p3, g, Lm, mc, ctx->d_full.p);
						else
							rc = big ? launch_family3_big<FAM3_NGLDM> (ctx, im, lst, n, p3, g, Lm, mc, ctx->d_full.p)
								: launch_family3_t<FAM3_NGLDM, false> (ctx, im, lst, n, p3, g, Lm, mc, ctx->d_full.p);
					}
					else if (fam == FAM3_GLDM)
					{
						if (Lm > 65535)
							return fail (ctx, NYXB200_EUNSUPPORTED, "3D GLDM needs grey levels < 65536 (max %d); set 3gldm/greydepth", Lm);
						rc = big ? launch_family3_big<FAM3_GLDM> (ctx, im, lst, n, prm, g, Lm, mc, ctx->d_full.p)
							: launch_family3_t<FAM3_GLDM, false> (ctx, im, lst, n, prm, g, Lm, mc, ctx->d_full.p);
					}
					else
					{
						if ((unsigned long long)(Lm + 2) * (unsigned long long)ngtdm3_ndmax (ctx->ngtdm3_radius) * (unsigned long long)mc >= 0xFFFFFFFFull || Lm + 2 > 65535)
							return fail (ctx, NYXB200_EUNSUPPORTED, "3D NGTDM: %d grey levels x %d neighbours x %u voxels overflow the exact 32-bit accumulators; lower 3ngtdm/greydepth or 3ngtdm/radius", Lm, ngtdm3_ndmax (ctx->ngtdm3_radius), mc);
						rc = big ? launch_family3_big<FAM3_NGTDM> (ctx, im, lst, n, prm, g, Lm, mc, ctx->d_full.p)
							: launch_family3_t<FAM3_NGTDM, false> (ctx, im, lst, n, prm, g, Lm, mc, ctx->d_full.p);
					}
					if (rc)
						return rc;
				}
				stage_end (ctx, st3[fam], ctx->ls);
			}
		}
		else
		for (int fam = FAM_GLCM; fam < FAM_COUNT; fam++)
		{
			if (!(ctx->fam_mask & (1 << fam)))
				continue;
			use_family_stream (ctx, fam);
			stage_begin (ctx, fam_stage[fam], ctx->ls);
			for (int c = 0; c < NCLASS; c++)
				if (int rc = launch_family (ctx, fam, imf, ctx->d_tlist.p + tb[c], ccount[c], prm, Lmax,
					cmaxcells[c], (int)cmaxwh[c], ctx->d_full.p, nodbg))
					return rc;
			stage_end (ctx, fam_stage[fam], ctx->ls);
		}
		if (int rc = join_families (ctx))
			return rc;
		if (bb == nb - 1)
		{
			ctx->cur_b = 0;
			stage_end (ctx, ST_FAMILIES);
			ctx->cur_b = bb;
		}
		// ---- rows of this batch: select + D2H, overlapped with the next batch (deferred when giant ROIs are pending)
		if (!any_giant)
			if (int rc = select_and_copy (batch_lo (bb), batch_lo (bb + 1), nb > 1))
				return rc;
	}
	// ---- ROIs too large for the shared-memory classes: one at a time, whole-GPU kernels (giant.cuh)
	if (any_giant)
	{
		stage_begin (ctx, ST_GIANT);
		if (int rc = process_giant_rois (ctx, im, info, prm, ncol_all))
			return rc;
		stage_end (ctx, ST_GIANT);
		if (int rc = select_and_copy (0, nroi, false))
			return rc;
	}
	else if (nb > 1)
	{
		CU (cudaEventRecord (ctx->ev_sel, ctx->copy_stream));
		CU (cudaStreamWaitEvent (s, ctx->ev_sel, 0));
	}
	ctx->cur_b = 0;
	stage_begin (ctx, ST_TOTAL);	// ev[ST_TOTAL] marks the end of the slide
	ctx->open_stage = -1;
	CU (cudaStreamSynchronize (s));
	(void)img_index;
	return 0;
}

void collect_timings (nyxb200_ctx* ctx)
{
	// every stage has its own begin / end event pair per ROI batch (family stages overlap when they run concurrently);
	// total = first event of the slide to the end-of-slide event on the main stream
	for (int b = 0; b < NB_MAX; b++)
		for (int i = 0; i < NYXB200_NSTAGES; i++)
		{
			const int e = b * nyxb200_ctx::NEV + i;
			float ms = 0;
			if (i != ST_TOTAL && ctx->ev_used[e] && ctx->ev_end_used[e] && cudaEventElapsedTime (&ms, ctx->ev[e], ctx->ev_end[e]) == cudaSuccess)
				ctx->stage_ms[i] += ms;
		}
	float ms = 0;
	if (ctx->ev_used[ST_H2D] && ctx->ev_used[ST_TOTAL] && cudaEventElapsedTime (&ms, ctx->ev[ST_H2D], ctx->ev[ST_TOTAL]) == cudaSuccess)
		ctx->stage_ms[ST_TOTAL] += ms;
}

bool host_pageable (const void* p)
{
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes (&a, p) != cudaSuccess)
	{
		cudaGetLastError();
		return true;
	}
	return a.type == cudaMemoryTypeUnregistered;
}

// pageable slide -> device: `staging_threads` workers, each with two pinned 8 MiB slots and its own stream; worker t takes
// the chunks c == t (mod T) of both buffers.  Returns at once; staged_upload_wait joins the workers (data is then on
// the device).
int staged_upload_start (nyxb200_ctx* ctx, uint8_t* d_int, const uint8_t* h_int, size_t ibytes, uint8_t* d_lab, const uint8_t* h_lab, size_t lbytes)
{
	const int T = ctx->staging_threads;
	for (int t = 0; t < T; t++)
	{
		if (!ctx->stg_stream[t])
			CU (cudaStreamCreateWithFlags (&ctx->stg_stream[t], cudaStreamNonBlocking));
		for (int q = 0; q < 2; q++)
		{
			if (!ctx->stg_slot[t * 2 + q])
				CU (cudaHostAlloc ((void**)&ctx->stg_slot[t * 2 + q], nyxb200_ctx::STG_CHUNK, cudaHostAllocDefault));
			if (!ctx->stg_ev[t * 2 + q])
				CU (cudaEventCreateWithFlags (&ctx->stg_ev[t * 2 + q], cudaEventDisableTiming));
		}
	}
	ctx->stg_rc = 0;
	const size_t CH = nyxb200_ctx::STG_CHUNK;
	const size_t nci = (ibytes + CH - 1) / CH, ncl = (lbytes + CH - 1) / CH;
	const int dev = ctx->st.device;
	for (int t = 0; t < T; t++)
		ctx->stg_workers.emplace_back ([=]()
		{
			if (cudaSetDevice (dev) != cudaSuccess)
			{
				ctx->stg_rc = 1;
				return;
			}
			int q = 0;
			bool used[2] = { false, false };
			for (size_t c = t; c < nci + ncl; c += T)
			{
				const bool isl = c >= nci;
				const size_t off = (isl ? c - nci : c) * CH, tot = isl ? lbytes : ibytes, n = std::min (CH, tot - off);
				const uint8_t* src = (isl ? h_lab : h_int) + off;
				uint8_t* dst = (isl ? d_lab : d_int) + off;
				uint8_t* slot = ctx->stg_slot[t * 2 + q];
				if (used[q] && cudaEventSynchronize (ctx->stg_ev[t * 2 + q]) != cudaSuccess)
					ctx->stg_rc = 1;
				memcpy (slot, src, n);
				if (cudaMemcpyAsync (dst, slot, n, cudaMemcpyHostToDevice, ctx->stg_stream[t]) != cudaSuccess
					|| cudaEventRecord (ctx->stg_ev[t * 2 + q], ctx->stg_stream[t]) != cudaSuccess)
					ctx->stg_rc = 1;
				used[q] = true;
				q ^= 1;
			}
			if (cudaStreamSynchronize (ctx->stg_stream[t]) != cudaSuccess)
				ctx->stg_rc = 1;
		});
	return 0;
}

int staged_upload_wait (nyxb200_ctx* ctx)
{
	for (auto& t : ctx->stg_workers)
		if (t.joinable())
			t.join();
	ctx->stg_workers.clear();
	if (ctx->stg_rc)
		return fail (ctx, NYXB200_ECUDA, "staged upload of a pageable host slide failed");
	return 0;
}

// `inten[k]` / `label[k]`: slide k (host or device memory).  Host slides are uploaded on the upload stream into two
// alternating device buffers, slide k+1 while slide k computes; the main stream's "h2d" stage is the part of the
// upload it had to wait for.
int featurize_slides (nyxb200_ctx* ctx, const void* const* inten, int inten_u16, const uint32_t* const* label, bool on_device,
	int n_img, int64_t D, int64_t H, int64_t W, nyxb200_table* out, const int32_t** row_image,
	const int* order = nullptr, int n_order = 0)
{
	// `order` (multi-device calls): the slides of the list this context processes, ascending; else all of them
	if (!ctx || !inten || !label || (!out && !ctx->share))
		return fail (ctx, NYXB200_EINVAL, "null argument");
	if (n_img <= 0 || H <= 0 || W <= 0 || D <= 0)
		return fail (ctx, NYXB200_EINVAL, "bad image dimensions");
	if (D * H > 0x7FFFFFFF)
		return fail (ctx, NYXB200_EUNSUPPORTED, "volume has more than 2^31-1 rows");
	if (H > 0x7FFFFFFF || W > 0x7FFFFFFF)
		return fail (ctx, NYXB200_EUNSUPPORTED, "image side exceeds 2^31-1");
	const int nmine = order ? n_order : n_img;
	auto slide_of = [&](int j) { return order ? order[j] : j; };
	for (int j = 0; j < nmine; j++)
		if (!inten[slide_of (j)] || !label[slide_of (j)])
			return fail (ctx, NYXB200_EINVAL, "null slide pointer (image %d)", slide_of (j));
	if (out)
	{
		out->data = nullptr;
		out->nrows = 0;
		out->ncols = (int)ctx->sel_cols.size() + 2;
		out->reserved = 0;
	}
	CU (cudaSetDevice (ctx->st.device));
	for (auto& m : ctx->stage_ms) m = 0;
	for (auto& v : ctx->counts) v = 0;
	ctx->giant_rois = 0;
	ctx->row_image.clear();
	const size_t npix = (size_t)D * (size_t)H * (size_t)W;
	const size_t ibytes = npix * (inten_u16 ? 2 : 4), lbytes = npix * 4;
	ctx->h_out_used = 0;
	const int nbuf = on_device ? 0 : (nmine > 1 ? 2 : 1);
	for (int b = 0; b < nbuf; b++)
	{
		CU (ctx->d_inten[b].reserve (ibytes));
		CU (ctx->d_label[b].reserve (npix));
	}
	bool staged[2] = { false, false };
	auto upload = [&](int j) -> int
	{
		const int b = j & 1, k = slide_of (j);
		staged[b] = ctx->staging_threads > 0 && ibytes + lbytes >= ((size_t)32 << 20) && host_pageable (inten[k]) && host_pageable (label[k]);
		if (staged[b])
			return staged_upload_start (ctx, ctx->d_inten[b].p, (const uint8_t*)inten[k], ibytes, (uint8_t*)ctx->d_label[b].p, (const uint8_t*)label[k], lbytes);
		CU (cudaMemcpyAsync (ctx->d_inten[b].p, inten[k], ibytes, cudaMemcpyHostToDevice, ctx->up_stream));
		CU (cudaMemcpyAsync (ctx->d_label[b].p, label[k], lbytes, cudaMemcpyHostToDevice, ctx->up_stream));
		CU (cudaEventRecord (ctx->ev_up[b], ctx->up_stream));
		return 0;
	};
	if (!on_device && nmine > 0)
		if (int rc = upload (0))
			return rc;
	for (int j = 0; j < nmine; j++)
	{
		const int k = slide_of (j);
		for (auto& u : ctx->ev_used) u = false;
		for (auto& u : ctx->ev_end_used) u = false;
		ctx->open_stage = -1;
		ImageView im;
		im.H = (int32_t)H;
		im.W = (int32_t)W;
		im.D = (int32_t)D;
		im.inten_u16 = inten_u16;
		stage_begin (ctx, ST_H2D);
		if (on_device)
		{
			im.inten = inten[k];
			im.label = label[k];
		}
		else
		{
			// buffer (j+1)&1 held the slide before this one, which has been fully processed (process_slide ends synchronised)
			if (staged[j & 1])
			{
				if (int rc = staged_upload_wait (ctx))	// slide j is on the device from here on
					return rc;
			}
			else
				CU (cudaStreamWaitEvent (ctx->stream, ctx->ev_up[j & 1], 0));
			if (j + 1 < nmine)
				if (int rc = upload (j + 1))
					return rc;
			im.inten = ctx->d_inten[j & 1].p;
			im.label = ctx->d_label[j & 1].p;
		}
		int64_t nr = 0;
		int rc = process_slide (ctx, im, k, &nr);
		if (rc)
		{
			const std::string msg = ctx->err;
			staged_upload_wait (ctx);
			cudaStreamSynchronize (ctx->up_stream);	// nothing may still be reading the caller's buffers
			ctx->err = msg;
			return rc;
		}
		collect_timings (ctx);
		if (!ctx->share)
			ctx->row_image.insert (ctx->row_image.end(), (size_t)nr, k);	// (one fill: the device is idle between two slides of a job)
	}
	if (out)
	{
		out->nrows = (int64_t)ctx->row_image.size();
		if (out->nrows)
			out->data = ctx->h_out;
	}
	if (row_image)
		*row_image = ctx->row_image.data();
	return NYXB200_OK;
}

// Longest-processing-time-first assignment of slides to devices by weight (pixel count); every device's list ascending.
// Equal weights give slide k to device k % n_dev.
void shard_slides (int n_dev, int n_img, const int64_t* weight, std::vector<int>& owner)
{
	owner.assign (n_img, 0);
	std::vector<int> idx (n_img);
	for (int i = 0; i < n_img; i++)
		idx[i] = i;
	if (weight)
		std::stable_sort (idx.begin(), idx.end(), [&](int a, int b) { return weight[a] > weight[b]; });
	std::vector<long long> load (n_dev, 0);
	for (int i : idx)
	{
		int best = 0;
		for (int d = 1; d < n_dev; d++)
			if (load[d] < load[best])
				best = d;
		owner[i] = best;
		load[best] += weight ? (long long)std::max<int64_t> (weight[i], 1) : 1;
	}
}

// contiguous stack of n_img slides
int featurize_common (nyxb200_ctx* ctx, const void* inten, int inten_u16, const uint32_t* label, bool on_device,
	int n_img, int64_t D, int64_t H, int64_t W, nyxb200_table* out, const int32_t** row_image)
{
	if (!ctx || !inten || !label || !out)
		return fail (ctx, NYXB200_EINVAL, "null argument");
	if (n_img <= 0 || H <= 0 || W <= 0 || D <= 0)
		return fail (ctx, NYXB200_EINVAL, "bad image dimensions");
	const size_t npix = (size_t)D * (size_t)H * (size_t)W;
	std::vector<const void*> ip (n_img);
	std::vector<const uint32_t*> lp (n_img);
	for (int k = 0; k < n_img; k++)
	{
		ip[k] = (const uint8_t*)inten + (size_t)k * npix * (inten_u16 ? 2 : 4);
		lp[k] = label + (size_t)k * npix;
	}
	return featurize_slides (ctx, ip.data(), inten_u16, lp.data(), on_device, n_img, D, H, W, out, row_image);
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

int nyxb200_featurize_2d_list (nyxb200_ctx* ctx, const void* const* inten, int inten_is_u16, const uint32_t* const* label,
	int on_device, int n_img, int64_t H, int64_t W, nyxb200_table* out, const int32_t** row_image)
{
	if (ctx && ctx->st.dim != 2)
		return fail (ctx, NYXB200_EINVAL, "this context was created for 3D features (dim = 3)");
	return featurize_slides (ctx, inten, inten_is_u16 ? 1 : 0, label, on_device != 0, n_img, 1, H, W, out, row_image);
}

int nyxb200_featurize_3d_list (nyxb200_ctx* ctx, const void* const* inten, int inten_is_u16, const uint32_t* const* label,
	int on_device, int n_vol, int64_t D, int64_t H, int64_t W, nyxb200_table* out, const int32_t** row_image)
{
	if (ctx && ctx->st.dim != 3)
		return fail (ctx, NYXB200_EINVAL, "this context was created for 2D features (dim = 2)");
	return featurize_slides (ctx, inten, inten_is_u16 ? 1 : 0, label, on_device != 0, n_vol, D, H, W, out, row_image);
}

int nyxb200_shard_slides (int n_dev, int n_img, const int64_t* weight, int32_t* owner)
{
	if (n_dev <= 0 || n_img < 0 || (!owner && n_img))
		return NYXB200_EINVAL;
	std::vector<int> o;
	shard_slides (n_dev, n_img, weight, o);
	for (int i = 0; i < n_img; i++)
		owner[i] = o[i];
	return NYXB200_OK;
}

int nyxb200_featurize_2d_list_multi (nyxb200_ctx* const* ctxs, int n_ctx, const void* const* inten, int inten_is_u16,
	const uint32_t* const* label, int on_device, int n_img, int64_t H, int64_t W, const int64_t* weight,
	nyxb200_table* out, const int32_t** row_image)
{
	if (!ctxs || n_ctx <= 0 || !ctxs[0])
		return NYXB200_EINVAL;
	nyxb200_ctx* c0 = ctxs[0];
	if (!inten || !label || !out)
		return fail (c0, NYXB200_EINVAL, "null argument");
	if (n_img <= 0 || H <= 0 || W <= 0)
		return fail (c0, NYXB200_EINVAL, "bad image dimensions");
	for (int d = 0; d < n_ctx; d++)
	{
		if (!ctxs[d] || ctxs[d]->st.dim != 2)
			return fail (c0, NYXB200_EINVAL, "context %d is null or not a 2D context", d);
		if (ctxs[d]->sel_cols != c0->sel_cols)
			return fail (c0, NYXB200_EINVAL, "context %d selects other features than context 0", d);
		for (int e = 0; e < d; e++)
			if (ctxs[e] == ctxs[d])
				return fail (c0, NYXB200_EINVAL, "context %d is given twice", d);
	}
	out->data = nullptr;
	out->nrows = 0;
	out->ncols = (int)c0->sel_cols.size() + 2;
	out->reserved = 0;
	std::vector<int> owner;
	shard_slides (n_ctx, n_img, weight, owner);
	std::vector<std::vector<int>> mine (n_ctx);
	for (int k = 0; k < n_img; k++)
		mine[owner[k]].push_back (k);
	// the previous result of context 0 goes back to the pool (the shared table takes its place)
	pool_release (c0->h_out);
	c0->h_out = nullptr;
	c0->h_out_cap = c0->h_out_used = 0;
	SharedOut sh;
	sh.rowlen = (size_t)out->ncols;
	sh.n_img = n_img;
	sh.row0.assign (n_img, 0);
	sh.nrows.assign (n_img, 0);
	sh.deferred.assign (n_img, 0);
	std::vector<int> rcs (n_ctx, 0);
	auto work = [&](int d)
	{
		nyxb200_ctx* c = ctxs[d];
		c->share = &sh;
		rcs[d] = mine[d].empty() ? 0 : featurize_slides (c, inten, inten_is_u16 ? 1 : 0, label, on_device != 0, n_img, 1, H, W, nullptr, nullptr,
			mine[d].data(), (int)mine[d].size());
		c->share = nullptr;
		if (rcs[d])
		{
			std::lock_guard<std::mutex> lk (sh.m);
			sh.failed = true;
			sh.cv.notify_all();
		}
	};
	{
		// one host thread per device (context 0 on the calling thread)
		std::vector<std::thread> th;
		for (int d = 1; d < n_ctx; d++)
			th.emplace_back (work, d);
		work (0);
		for (auto& t : th)
			t.join();
	}
	{
		// report the failure that started it, not the "another device failed" of the contexts that gave up after it
		int bad = -1;
		for (int d = 0; d < n_ctx; d++)
			if (rcs[d] && (bad < 0 || ctxs[bad]->err.find ("another device") != std::string::npos))
				bad = d;
		if (bad >= 0)
		{
			if (bad)
				c0->err = "context " + std::to_string (bad) + " (device " + std::to_string (ctxs[bad]->st.device) + "): " + ctxs[bad]->err;
			pool_release (sh.base);
			return rcs[bad];
		}
	}
	cudaSetDevice (c0->st.device);
	// slides that did not fit the estimate: grow to the exact size, then a second pass on context 0 (their row ranges are known)
	const size_t total = (size_t)sh.cursor;
	bool any_deferred = false;
	for (int k = 0; k < n_img; k++)
		any_deferred = any_deferred || sh.deferred[k];
	if (any_deferred)
	{
		size_t got = 0;
		double* nb = (double*)pool_acquire ((total + 1) * sh.rowlen * sizeof (double), &got);
		if (!nb)
		{
			pool_release (sh.base);
			return fail (c0, NYXB200_ENOMEM, "cannot grow the pinned result table");
		}
		for (int k = 0; k < n_img; k++)
			if (!sh.deferred[k])
				memcpy (nb + (size_t)sh.row0[k] * sh.rowlen, sh.base + (size_t)sh.row0[k] * sh.rowlen, (size_t)sh.nrows[k] * sh.rowlen * sizeof (double));
		pool_release (sh.base);
		sh.base = nb;
		sh.cap_rows = got / (sh.rowlen * sizeof (double));
		for (int k = 0; k < n_img; k++)
		{
			if (!sh.deferred[k])
				continue;
			// replay the claim for exactly this slide
			const int64_t r0 = sh.row0[k];
			SharedOut one;
			one.rowlen = sh.rowlen;
			one.n_img = 1;
			one.base = sh.base + (size_t)r0 * sh.rowlen;
			one.cap_rows = (size_t)sh.nrows[k];
			one.row0.assign (n_img, 0);
			one.nrows.assign (n_img, 0);
			one.deferred.assign (n_img, 0);
			one.next_slide = k;
			c0->share = &one;
			const int rc = featurize_slides (c0, inten, inten_is_u16 ? 1 : 0, label, on_device != 0, n_img, 1, H, W, nullptr, nullptr, &k, 1);
			c0->share = nullptr;
			if (rc || one.deferred[k])
			{
				pool_release (sh.base);
				return rc ? rc : fail (c0, NYXB200_ECUDA, "second pass of slide %d saw another ROI count", k);
			}
		}
	}
	c0->row_image.resize (total);
	for (int k = 0; k < n_img; k++)
		std::fill (c0->row_image.begin() + sh.row0[k], c0->row_image.begin() + sh.row0[k] + sh.nrows[k], k);
	// the table becomes context 0's current result (nyxb200_table_detach / the next call handle it like any other)
	c0->h_out = sh.base;
	c0->h_out_cap = sh.cap_rows * sh.rowlen;
	c0->h_out_used = total * sh.rowlen;
	out->nrows = (int64_t)total;
	out->data = total ? sh.base : nullptr;
	if (row_image)
		*row_image = c0->row_image.data();
	return NYXB200_OK;
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

int nyxb200_table_detach (nyxb200_ctx* ctx, nyxb200_table* t)
{
	if (!ctx || !t)
		return fail (ctx, NYXB200_EINVAL, "null argument");
	if (!t->data)
		return NYXB200_OK;	// empty table: nothing to hand over
	if (t->data != ctx->h_out)
		return fail (ctx, NYXB200_EINVAL, "the table is not the context's current result");
	// the caller owns the buffer from here on (nyxb200_host_release); the next featurize call takes another one
	ctx->h_out = nullptr;
	ctx->h_out_cap = 0;
	ctx->h_out_used = 0;
	return NYXB200_OK;
}

void nyxb200_host_release (void* p)
{
	pool_release (p);
}

int nyxb200_get_param (const nyxb200_ctx* ctx, const char* key, double* value)
{
	if (!ctx || !key || !value)
		return NYXB200_EINVAL;
	if (!strcmp (key, "glcm/offset")) *value = ctx->st.glcm_offset;
	else if (!strcmp (key, "coarse_gray_depth")) *value = ctx->st.coarse_gray_depth;
	else if (!strcmp (key, "glcm/greydepth") || !strcmp (key, "3glcm/greydepth"))
		*value = ctx->glcm_greydepth_set ? ctx->glcm_greydepth : ctx->st.coarse_gray_depth;
	else if (!strcmp (key, "ibsi")) *value = ctx->st.ibsi;
	else if (!strcmp (key, "glcm/symmetric")) *value = ctx->st.glcm_symmetric;
	else if (!strcmp (key, "soft_nan")) *value = ctx->st.soft_nan;
	else if (!strcmp (key, "concurrent_families")) *value = ctx->concurrent;
	else if (!strcmp (key, "roi_batches")) *value = ctx->roi_batches;
	else if (!strcmp (key, "tile_store")) *value = ctx->use_tile_store;
	else if (!strcmp (key, "staging_threads")) *value = ctx->staging_threads;
	else if (!strcmp (key, "covered_range_from_scan")) *value = ctx->covered_from_scan;
	else if (!strcmp (key, "single_roi")) *value = ctx->single_roi;
	else if (!strcmp (key, "3glszm/greydepth")) *value = ctx->glszm3_greydepth;
	else if (!strcmp (key, "3gldm/greydepth")) *value = ctx->gldm3_greydepth;
	else if (!strcmp (key, "3glrlm/greydepth")) *value = ctx->glrlm3_greydepth;
	else if (!strcmp (key, "3ngtdm/greydepth")) *value = ctx->ngtdm3_greydepth;
	else if (!strcmp (key, "3ngtdm/radius")) *value = ctx->ngtdm3_radius;
	else
		return NYXB200_EINVAL;
	return NYXB200_OK;
}

int nyxb200_device_count (void)
{
	int ndev = 0;
	if (cudaGetDeviceCount (&ndev) != cudaSuccess)
	{
		cudaGetLastError();
		return 0;
	}
	int n = 0;
	for (int d = 0; d < ndev; d++)
	{
		int major = 0;
		if (cudaDeviceGetAttribute (&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10)
			n = d + 1;	// ordinals are CUDA's: the count is "one past the last usable device"
	}
	return n;
}

int nyxb200_device_info (int device, char* name, size_t namecap, int64_t* total_mem, int* sm_count, int* cc_major, int* cc_minor)
{
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties (&prop, device) != cudaSuccess)
	{
		cudaGetLastError();
		return NYXB200_ECUDA;
	}
	if (name && namecap)
		snprintf (name, namecap, "%s", prop.name);
	if (total_mem) *total_mem = (int64_t)prop.totalGlobalMem;
	if (sm_count) *sm_count = prop.multiProcessorCount;
	if (cc_major) *cc_major = prop.major;
	if (cc_minor) *cc_minor = prop.minor;
	return NYXB200_OK;
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
	const Params prm = make_params (ctx);
	if ((unsigned long long)r.w * (unsigned long long)r.h > 40000ull)
	{
		// giant path: run this one family on the ROI and read its finished matrix back
		GiantDebug gd;
		std::vector<uint32_t> dense, zk;
		std::vector<double> sv;
		gd.family = family;
		gd.dense = &dense;
		gd.zone_keys = &zk;
		gd.ngtdm_s = &sv;
		DevBuf<double> row;
		CU (row.reserve (NCOL_ALL));
		int rc = giant_texture_roi (ctx, ctx->last_im, r, prm, 1 << family, row.p, &gd);
		cudaStreamSynchronize (ctx->stream);
		row.release();
		if (rc)
			return rc;
		int64_t cnt = 0;
		auto emit = [&](int64_t ang, int64_t a, int64_t b, int64_t c, double s) {
			if (trip && cnt < cap)
			{
				trip[cnt * 4] = ang; trip[cnt * 4 + 1] = a; trip[cnt * 4 + 2] = b; trip[cnt * 4 + 3] = c;
				if (svals)
					svals[cnt] = s;
			}
			cnt++;
		};
		if (family == FAM_GLCM && !dense.empty())
		{
			const size_t Ng = gd.dim0;
			for (size_t a = 0; a < 4; a++)
				for (size_t row_ = 0; row_ < Ng; row_++)
					for (size_t col = 0; col < Ng; col++)
						if (uint32_t v = dense[(a * Ng + row_) * Ng + col])
							emit ((int64_t)a, (int64_t)row_ + 1, (int64_t)col + 1, v, 0);
		}
		else if (family == FAM_GLRLM && !dense.empty())
		{
			const size_t Lm = gd.dim0, Nr = gd.dim1;
			for (size_t a = 0; a < 4; a++)
				for (size_t lv = 0; lv < Lm; lv++)
					for (size_t j = 0; j < Nr; j++)
						if (uint32_t v = dense[(a * Lm + lv) * Nr + j])
							emit ((int64_t)a, (int64_t)lv + 1, (int64_t)j + 1, v, 0);
		}
		else if (family == FAM_GLSZM || family == FAM_GLDZM)
		{
			std::vector<unsigned long long> keys (zk.size() / 2);
			for (size_t i = 0; i < keys.size(); i++)
				keys[i] = ((unsigned long long)zk[2 * i] << 32) | zk[2 * i + 1];
			std::sort (keys.begin(), keys.end());
			for (size_t i = 0; i < keys.size(); )
			{
				size_t j = i;
				while (j < keys.size() && keys[j] == keys[i]) j++;
				emit (0, (int64_t)(keys[i] >> 32), (int64_t)(keys[i] & 0xFFFFFFFFull), (int64_t)(j - i), 0);
				i = j;
			}
		}
		else if ((family == FAM_GLDM || family == FAM_NGLDM) && !dense.empty())
		{
			const int j0 = family == FAM_GLDM ? 1 : 0, j1 = family == FAM_GLDM ? 9 : 8;
			for (int lv = family == FAM_GLDM ? 1 : 0; lv < gd.dim0; lv++)
				for (int j = j0; j <= j1; j++)
					if (uint32_t v = dense[(size_t)lv * 10 + j])
						emit (0, lv, j, v, 0);
		}
		else if (family == FAM_NGTDM && !dense.empty())
		{
			int present = 0;	// the reference returns before building the matrix when |I| < 2 (ngtdm.cpp:70-78)
			for (int lv = 0; lv < gd.dim0; lv++)
				present += dense[(size_t)gd.dim0 + lv] != 0;
			const int isize = ctx->st.ibsi ? (present ? gd.dim0 : 0) : present;
			for (int lv = 0; lv < gd.dim0 && isize >= 2; lv++)
				if (dense[lv])
					emit (0, lv, 0, dense[lv], sv[lv]);
		}
		*n = cnt;
		return NYXB200_OK;
	}
	const int Lmax = ctx->st.ibsi ? (int)std::max<uint32_t> (r.imax, 1u) : std::abs (ctx->st.coarse_gray_depth);
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
