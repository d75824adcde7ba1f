/*
 * nyxus_b200 -- C ABI of the B200-native texture + intensity feature path.
 *
 * This is the drop-in boundary for ONE hot path of PolusAI/nyxus: everything that happens between
 * `featurize_montage_imp()` receiving two uint32 arrays (reference src/nyx/python/new_bindings_py.cpp:562-645)
 * and the numeric feature table it returns, i.e.
 *     phase 1  gatherRoisMetricsInMemory      src/nyx/phase1.cpp:373-407
 *     phase 2  processTrivialRoisInMemory     src/nyx/phase2_2d.cpp:637-770
 *     reduce   reduce_trivial_2d              src/nyx/reduce_trivial_rois.cpp:62-253   (PixelIntensity, GLCM, GLRLM,
 *                                                                                       GLDZM, GLSZM, GLDM, NGLDM, NGTDM)
 *     output   save_features_2_buffer         src/nyx/output_2_buffer.cpp:303-582
 * Plain pointers and sizes only; no C++/torch types.  Every function returns 0 on success or a negative
 * NYXB200_E* code; nyxb200_last_error() gives the message (reference convention: C++ exceptions turned into
 * Python exceptions by pybind11 -- the binding layer raises on a non-zero status).
 * There is NO CPU fallback: without a CUDA device of compute capability 10.x nyxb200_create() fails.
 */
#ifndef NYXUS_B200_H
#define NYXUS_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NYXB200_OK 0
#define NYXB200_EINVAL -1      /* bad argument / unknown feature name (reference: std::invalid_argument, env_features.cpp:496,507) */
#define NYXB200_ECUDA -2       /* CUDA runtime failure or no usable device */
#define NYXB200_EUNSUPPORTED -3 /* input outside what this build handles (message says which limit) */
#define NYXB200_ENOMEM -4

typedef struct nyxb200_ctx nyxb200_ctx;

/* Mirrors the subset of Environment / Fsettings the path reads (reference env_features.cpp:712-736,
 * feature_settings.h:19-53, nyxus.py:157-261 defaults). */
typedef struct nyxb200_settings
{
	int32_t device;             /* CUDA device ordinal (>= 0) */
	int32_t coarse_gray_depth;  /* NyxSetting::GREYDEPTH and GLCM_GREYDEPTH, default 64, must be > 0 */
	int32_t ibsi;               /* NyxSetting::IBSI */
	int32_t glcm_offset;        /* NyxSetting::GLCM_OFFSET, default 1 */
	int32_t glcm_symmetric;     /* GLCMFeature::symmetric_glcm, default 0 */
	int32_t verbose;
	double soft_nan;            /* NyxSetting::SOFTNAN (resultOptions.noval), default 0.0 */
	int32_t dim;                /* 2: Nyxus (2D feature codes, nyxb200_featurize_2d_*); 3: Nyxus3D (Feature3D codes, nyxb200_featurize_3d_*) */
	int32_t reserved;
} nyxb200_settings;

/* Numeric result table, row-major, `ncols` doubles per row: ROI_label, t_index, then the enabled feature columns
 * (reference output_2_buffer.cpp:441-578: rows by ascending label, every value through force_finite_number(v, noval)).
 * `data` is owned by the library (the context's pinned host buffer): valid until the next featurize call on the
 * same context, nyxb200_table_free() or nyxb200_destroy(). */
typedef struct nyxb200_table
{
	double* data;
	int64_t nrows;
	int32_t ncols;
	int32_t reserved;
} nyxb200_table;

void nyxb200_default_settings (nyxb200_settings* s);

/* replaces backend.initialize_environment (new_bindings_py.cpp:75-154) for this path */
int nyxb200_create (const nyxb200_settings* s, nyxb200_ctx** out);
/* replaces backend.release_environment_imp (new_bindings_py.cpp:42-46) */
void nyxb200_destroy (nyxb200_ctx* ctx);
const char* nyxb200_last_error (const nyxb200_ctx* ctx);   /* ctx may be NULL: error of the last failed create */

/* Feature selection: group names ("*ALL_INTENSITY*", "*ALL_GLCM*", ..., "*BASIC_MORPHOLOGY*", "*ALL_IH*", "*GEOMOMS*", "*SGEOMOMS*",
 * "*IGEOMOMS*", plus "*ALL_TEXTURE*"; dim == 3: "*3D_ALL_INTENSITY*", "*3D_ALL_TEXTURE*", "*3D_GLCM*" ...) and individual
 * feature names, exactly as Environment::expand_featuregroups does (env_features.cpp:456-528).  Unknown name -> EINVAL. */
int nyxb200_select_features (nyxb200_ctx* ctx, const char* const* names, int n);
int nyxb200_num_columns (const nyxb200_ctx* ctx);               /* feature columns (without ROI_label, t_index) */
const char* nyxb200_column_name (const nyxb200_ctx* ctx, int i); /* e.g. "GLCM_ASM_45" */
int nyxb200_column_code (const nyxb200_ctx* ctx, int i);         /* Feature2D code of column i */
/* replaces set_metaparam_imp("glcm/offset=..") etc. for the keys this path reads (env_metaparams.cpp:63-247):
 * "glcm/offset", "glcm/greydepth", "coarse_gray_depth", "ibsi", "glcm/symmetric", "soft_nan", and for dim == 3
 * "3glcm/greydepth", "3glszm/greydepth", "3gldm/greydepth", "3glrlm/greydepth", "3ngtdm/greydepth", "3ngtdm/radius";
 * per-call flags of the file-based workflows: "covered_range_from_scan", "single_roi" (the reference's whole-slide mode) */
int nyxb200_set_param (nyxb200_ctx* ctx, const char* key, double value);

/* THE hot path.  Host buffers (pageable or pinned), `n_img` images of H x W, row-major, C-contiguous:
 * replaces featurize_montage_imp -> processMontage (workflow_pythonapi.cpp:88-192).  Rows of all images are
 * concatenated; `row_image` (optional, may be NULL) receives a library-owned int32 array giving the image index
 * of every row (valid until the table is freed). */
int nyxb200_featurize_2d_u32 (nyxb200_ctx* ctx, const uint32_t* inten, const uint32_t* label,
	int n_img, int64_t H, int64_t W, nyxb200_table* out, const int32_t** row_image);
/* same, 16-bit intensities (avoids the uint32 widening copy of nyxus.py:485-486) */
int nyxb200_featurize_2d_u16 (nyxb200_ctx* ctx, const uint16_t* inten, const uint32_t* label,
	int n_img, int64_t H, int64_t W, nyxb200_table* out, const int32_t** row_image);
/* same, but `inten` / `label` are DEVICE pointers on the context's device (inputs already resident in HBM);
 * inten_is_u16 selects the pixel type.  The table still comes back to (pinned) host memory. */
int nyxb200_featurize_2d_device (nyxb200_ctx* ctx, const void* inten, int inten_is_u16, const uint32_t* label,
	int n_img, int64_t H, int64_t W, nyxb200_table* out, const int32_t** row_image);
/* 3D (dim == 3 contexts): `n_vol` volumes of D x H x W voxels, voxel (z,y,x) at (z*H + y)*W + x.  Replaces the compute
 * part of featurize_fname_lists_3D_imp -> processDataset_3D_segmented (new_bindings_py.cpp:753-856,
 * workflow_3d_segmented.cpp:35-260): gatherRoisMetrics_3D (phase1.cpp:248-366), scanTrivialRois_3D +
 * allocateTrivialRoisBuffers_3D (phase2_3d.cpp:24-158,816-840) and the D3_VoxelIntensityFeatures / D3_GLCM_feature /
 * D3_GLSZM_feature / D3_NGTDM_feature reducers (reduce_trivial_rois.cpp:342-413).  File decoding (NIfTI) stays with the
 * caller.  Feature names are the reference's "3COV", "3GLCM_ASM", ... and the groups *3D_ALL_INTENSITY*, *3D_GLCM*,
 * *3D_GLSZM*, *3D_NGTDM*. */
int nyxb200_featurize_3d_u32 (nyxb200_ctx* ctx, const uint32_t* inten, const uint32_t* label,
	int n_vol, int64_t D, int64_t H, int64_t W, nyxb200_table* out, const int32_t** row_image);
int nyxb200_featurize_3d_u16 (nyxb200_ctx* ctx, const uint16_t* inten, const uint32_t* label,
	int n_vol, int64_t D, int64_t H, int64_t W, nyxb200_table* out, const int32_t** row_image);
int nyxb200_featurize_3d_device (nyxb200_ctx* ctx, const void* inten, int inten_is_u16, const uint32_t* label,
	int n_vol, int64_t D, int64_t H, int64_t W, nyxb200_table* out, const int32_t** row_image);
/* A LIST of slides / volumes of one shape, each with its own pointer (host when on_device == 0, else device): replaces
 * the slide loop of processMontage (workflow_pythonapi.cpp:104-190) and of the file-list workflows
 * (featurize_fname_lists_imp / featurize_fname_lists_3D_imp, new_bindings_py.cpp:647-856) after decoding.  Host slides are
 * pipelined: slide k+1 is copied to the device (second buffer, own stream) while slide k computes and the rows of
 * slide k-1 travel back, so a multi-slide job runs at max(upload, compute) per slide instead of their sum.  The
 * contiguous-stack entry points above take the same pipeline when n_img > 1. */
int nyxb200_featurize_2d_list (nyxb200_ctx* ctx, const void* const* inten, int inten_is_u16, const uint32_t* const* label,
	int on_device, int n_img, int64_t H, int64_t W, nyxb200_table* out, const int32_t** row_image);
int nyxb200_featurize_3d_list (nyxb200_ctx* ctx, const void* const* inten, int inten_is_u16, const uint32_t* const* label,
	int on_device, int n_vol, int64_t D, int64_t H, int64_t W, nyxb200_table* out, const int32_t** row_image);
/* SEVERAL DEVICES, one call (BASELINE configs[4]: a list of slides sharded by pixel count over the GPUs of one box;
 * reference: the slide loop of processDataset_2D_segmented / processMontage, one slide after the other on the host
 * cores).  `ctxs`: one 2D context per device with the same feature selection.  Slide k goes to the device
 * nyxb200_shard_slides() names: longest-processing-time-first on `weight[k]` (pixel count or any cost estimate; NULL =
 * equal, slide k -> device k % n_ctx), every device walking its slides in ascending order with the pipeline of
 * nyxb200_featurize_2d_list (upload of its next slide || compute || D2H of finished ROI batches), one host thread per
 * device.  No inter-GPU exchange: ROIs never span slides.  Rows land in ONE pinned host table, slide-major in slide
 * order and by ascending label inside a slide -- the order a single-device call gives -- without a gather pass: a
 * slide claims its row range as soon as its ROI count is known and its rows are copied from the device straight to
 * their final place.  With on_device != 0, inten[k] / label[k] must live on the device of ctxs[owner[k]].
 * The table belongs to ctxs[0] (valid until its next call; nyxb200_table_detach works on it). */
int nyxb200_shard_slides (int n_dev, int n_img, const int64_t* weight, int32_t* owner);
int nyxb200_featurize_2d_list_multi (nyxb200_ctx* const* ctxs, int n_ctx, const void* const* inten, int inten_is_u16,
	const uint32_t* const* label, int on_device, int n_img, int64_t H, int64_t W, const int64_t* weight,
	nyxb200_table* out, const int32_t** row_image);
void nyxb200_table_free (nyxb200_ctx* ctx, nyxb200_table* t);
/* Ownership hand-over of a result table (reference: the std::vector<double> moved into a capsule-owned numpy array,
 * as_pyarray new_bindings_py.cpp:63-73,631-640): after a successful detach `t->data` belongs to the caller and stays
 * valid across later featurize calls and nyxb200_destroy; give it back with nyxb200_host_release (the pinned buffer
 * returns to a process-wide pool and is reused by the next call instead of being re-pinned). */
int nyxb200_table_detach (nyxb200_ctx* ctx, nyxb200_table* t);
void nyxb200_host_release (void* data);

/* Device-free checks for the binding layer (the reference validates feature names in initialize_environment,
 * env_features.cpp:496,507, and answers gpu_available_imp / get_gpu_props without touching a device context,
 * new_bindings_py.cpp:887-922). */
int nyxb200_validate_features (int dim, const char* const* names, int n, char* msg, size_t msgcap);
int nyxb200_get_param (const nyxb200_ctx* ctx, const char* key, double* value);   /* same keys as nyxb200_set_param */
int nyxb200_device_count (void);      /* one past the last CUDA ordinal of compute capability 10.x; 0 = none / no driver */
int nyxb200_device_info (int device, char* name, size_t namecap, int64_t* total_mem, int* sm_count, int* cc_major, int* cc_minor);

/* ---- measurement hooks ---- */
#define NYXB200_NSTAGES 18
/* per-stage device time (ms, CUDA events on the launching stream) of the last featurize call:
 * 0 h2d, 1 init, 2 scan, 3 compact, 4 pack, 5 intensity, 6 glcm, 7 glrlm, 8 gldzm, 9 glszm, 10 gldm, 11 ngldm,
 * 12 ngtdm, 13 select, 14 d2h, 15 total (first to last event of the call), 16 families (the fork/join window: the
 * family stages 4..12 run on their own streams and overlap unless nyxb200_set_param("concurrent_families", 0)),
 * 17 giant (ROIs larger than the shared-memory size classes, processed one at a time by whole-GPU kernels) */
int nyxb200_last_timings (const nyxb200_ctx* ctx, float* ms, int n);
/* counts of the last featurize call: 0 kernel launches, 1 ROIs, 2 foreground pixels, 3 all pixels */
int nyxb200_last_counts (const nyxb200_ctx* ctx, int64_t* v, int n);
const char* nyxb200_stage_name (int i);

/* ---- test hooks (matrix-exact parity) ---- */
/* ROI records of the last featurize call, 8 int64 per ROI: label, area, min, max, xmin, ymin, width, height */
int nyxb200_debug_rois (nyxb200_ctx* ctx, int64_t* out, int64_t cap, int64_t* n);
/* the CSR pixel cloud of one ROI (intensities in raster order) */
int nyxb200_debug_cloud (nyxb200_ctx* ctx, uint32_t label, uint32_t* out, int64_t cap, int64_t* n);
/* Sparse dump of one family's matrix for the ROI with `label` (image of the last call must be a 1-image call):
 * rows of 4 int64 {angle, a, b, count}; family ids as in nyxb200_family_id().  For NGTDM `svals` receives S per row. */
int nyxb200_debug_matrix (nyxb200_ctx* ctx, uint32_t label, int family, int64_t* trip, int64_t cap, int64_t* n,
	double* svals);
int nyxb200_family_id (const char* name); /* "INT" 0, "GLCM" 1, "GLRLM" 2, "GLDZM" 3, "GLSZM" 4, "GLDM" 5, "NGLDM" 6, "NGTDM" 7 */

const char* nyxb200_version (void);

#ifdef __cplusplus
}
#endif
#endif
