// nyxus_b200 -- canonical column layout of the texture + intensity path.
//
// The order is the reference's output contract: feature codes ascending in the Feature2D enum
// (reference featureset.h:10-350), GLCM / GLRLM non-_AVE codes expanded into 4 angle columns
// (reference output_2_buffer.cpp:331-358, 470-520).  337 columns in total.
#pragma once

namespace nyx
{

enum Family : int
{
	FAM_INT = 0, FAM_GLCM, FAM_GLRLM, FAM_GLDZM, FAM_GLSZM, FAM_GLDM, FAM_NGLDM, FAM_NGTDM, FAM_COUNT
};

// number of canonical columns per family and their base offsets
constexpr int NCOL_INT = 36;
constexpr int NCOL_GLCM = 30 * 4 + 29;
constexpr int NCOL_GLRLM = 16 * 4 + 16;
constexpr int NCOL_GLDZM = 18;
constexpr int NCOL_GLSZM = 16;
constexpr int NCOL_GLDM = 14;
constexpr int NCOL_NGLDM = 19;
constexpr int NCOL_NGTDM = 5;

constexpr int COL_INT = 0;
constexpr int COL_GLCM = COL_INT + NCOL_INT;
constexpr int COL_GLRLM = COL_GLCM + NCOL_GLCM;
constexpr int COL_GLDZM = COL_GLRLM + NCOL_GLRLM;
constexpr int COL_GLSZM = COL_GLDZM + NCOL_GLDZM;
constexpr int COL_GLDM = COL_GLSZM + NCOL_GLSZM;
constexpr int COL_NGLDM = COL_GLDM + NCOL_GLDM;
constexpr int COL_NGTDM = COL_NGLDM + NCOL_NGLDM;
constexpr int NCOL_TEXINT = COL_NGTDM + NCOL_NGTDM;	// 337: first-order + the seven texture families
// *BASIC_MORPHOLOGY*  is appended to the canonical table; the OUTPUT order is by Feature2D code (the feature
// table), so its columns still come out between the first-order and the GLCM columns like the reference's
constexpr int FAM_MORPH = FAM_COUNT;			// not a texture family: no size classes, no tile
constexpr int NCOL_MORPH = 15;
constexpr int COL_MORPH = NCOL_TEXINT;
// *ALL_IH* (SURVEY 8f-2): the IBSI intensity-histogram family, 46 columns, IBSI mode only
constexpr int FAM_IH = FAM_COUNT + 1;
constexpr int NCOL_IH = 46;
constexpr int COL_IH = COL_MORPH + NCOL_MORPH;
// unweighted geometric moments: shape variant (SPAT_MOMENT_* ... HU_M*) then intensity variant (IMOM_*),
// each RM[13] | CM[16] | NRM[16] | NCM[7] | HU[7]
constexpr int FAM_GEOM = FAM_COUNT + 2;
constexpr int GM_NCOL = 13 + 16 + 16 + 7 + 7;		// 59
constexpr int NCOL_GEOMOM = 2 * GM_NCOL;
constexpr int COL_GEOM = COL_IH + NCOL_IH;
// the distance-to-contour weighted moments (contour_body.h): shape variant [WRM 10 | WCM 7 | WNCM 7 | WHU 7], then intensity
constexpr int FAM_WMOM = FAM_COUNT + 3;
constexpr int WMOM_NCOL = 10 + 7 + 7 + 7;		// 31
constexpr int NCOL_WMOM = 2 * WMOM_NCOL;
constexpr int COL_WMOM = COL_GEOM + NCOL_GEOMOM;
constexpr int NCOL_ALL = COL_WMOM + NCOL_WMOM;	// 578
// HISTOGRAM (featureset.h: opt-in, multi-valued): one column per bin of the first-order family's grey-depth histogram,
// HISTOGRAM_BIN_0 .. _N-1 with N = coarse grey depth (output_2_buffer.cpp:158-168,283-293) -- the only block of the canonical
// row whose width is a run-time value; it sits after every fixed block
constexpr int FAM_HIST = FAM_COUNT + 4;
constexpr int COL_HIST = NCOL_ALL;

// ---- first-order (Feature2D::COV .. UNIFORMITY_PIU, featureset.h:13-48)
enum IntCol : int
{
	I_COV = 0, I_COVERED_IMAGE_INTENSITY_RANGE, I_ENERGY, I_ENTROPY, I_EXCESS_KURTOSIS, I_HYPERFLATNESS, I_HYPERSKEWNESS,
	I_INTEGRATED_INTENSITY, I_INTERQUARTILE_RANGE, I_KURTOSIS, I_MAX, I_MEAN, I_MEAN_ABSOLUTE_DEVIATION, I_MEDIAN,
	I_MEDIAN_ABSOLUTE_DEVIATION, I_MIN, I_MODE, I_P01, I_P10, I_P25, I_P75, I_P90, I_P99, I_QCOD, I_RANGE, I_ROBUST_MEAN,
	I_ROBUST_MEAN_ABSOLUTE_DEVIATION, I_ROOT_MEAN_SQUARED, I_SKEWNESS, I_STANDARD_DEVIATION, I_STANDARD_DEVIATION_BIASED,
	I_STANDARD_ERROR, I_VARIANCE, I_VARIANCE_BIASED, I_UNIFORMITY, I_UNIFORMITY_PIU
};

// ---- basic morphology (Feature2D::AREA_PIXELS_COUNT .. ASPECT_RATIO, featureset.h:46-60)
enum MorphCol : int
{
	MO_AREA_PIXELS_COUNT = 0, MO_AREA_UM2, MO_CENTROID_X, MO_CENTROID_Y, MO_WEIGHTED_CENTROID_Y, MO_WEIGHTED_CENTROID_X,
	MO_MASS_DISPLACEMENT, MO_COMPACTNESS, MO_BBOX_YMIN, MO_BBOX_XMIN, MO_BBOX_HEIGHT, MO_BBOX_WIDTH, MO_DIAMETER_EQUAL_AREA,
	MO_EXTENT, MO_ASPECT_RATIO
};

// ---- IBSI intensity histogram (Feature2D::IH_MEAN_VAL .. IH_BIN_SIZE, featureset.h:590-637)
enum IhCol : int
{
	IH_MEAN_VAL = 0, IH_VARIANCE_VAL, IH_SKEWNESS_VAL, IH_EXCESS_KURTOSIS_VAL, IH_MEDIAN_VAL, IH_MINIMUM_VAL, IH_P10_VAL, IH_P90_VAL,
	IH_MAXIMUM_VAL, IH_MODE_VAL, IH_INTERQUANTILE_RANGE_VAL, IH_RANGE_VAL, IH_MEAN_ABSOLUTE_DEVIATION_VAL,
	IH_ROBUST_MEAN_ABSOLUTE_DEVIATION_VAL, IH_MEDIAN_ABSOLUTE_DEVIATION_VAL, IH_COEFFICIENT_OF_VARIATION_VAL,
	IH_QUANTILE_COEFFICIENT_OF_DISPERSION_VAL, IH_ENTROPY_VAL, IH_UNIFORMITY_VAL, IH_ROBUST_MEAN_VAL, IH_MEAN_IDX, IH_VARIANCE_IDX,
	IH_SKEWNESS_IDX, IH_EXCESS_KURTOSIS_IDX, IH_MEDIAN_IDX, IH_MINIMUM_IDX, IH_P10_IDX, IH_P90_IDX, IH_MAXIMUM_IDX, IH_MODE_IDX,
	IH_INTERQUANTILE_RANGE_IDX, IH_RANGE_IDX, IH_MEAN_ABSOLUTE_DEVIATION_IDX, IH_ROBUST_MEAN_ABSOLUTE_DEVIATION_IDX,
	IH_MEDIAN_ABSOLUTE_DEVIATION_IDX, IH_COEFFICIENT_OF_VARIATION_IDX, IH_QUANTILE_COEFFICIENT_OF_DISPERSION_IDX, IH_ENTROPY_IDX,
	IH_UNIFORMITY_IDX, IH_MAX_GRADIENT, IH_MAX_GRADIENT_IDX, IH_MIN_GRADIENT, IH_MIN_GRADIENT_IDX, IH_ROBUST_MEAN_IDX, IH_NUM_BINS,
	IH_BIN_SIZE
};

// ---- GLCM angled features in enum order (featureset.h:174-203); column = COL_GLCM + 4*f + angle
enum GlcmF : int
{
	G_ASM = 0, G_ACOR, G_CLUPROM, G_CLUSHADE, G_CLUTEND, G_CONTRAST, G_CORRELATION, G_DIFAVE, G_DIFENTRO, G_DIFVAR, G_DIS,
	G_ENERGY, G_ENTROPY, G_HOM1, G_HOM2, G_ID, G_IDN, G_IDM, G_IDMN, G_INFOMEAS1, G_INFOMEAS2, G_IV, G_JAVE, G_JE, G_JMAX,
	G_JVAR, G_SUMAVERAGE, G_SUMENTROPY, G_SUMVARIANCE, G_VARIANCE, G_NANGLED
};
// the 29 _AVE columns follow in THEIR enum order (featureset.h:205-233); entry = the angled feature averaged
#define NYX_GLCM_AVE_ORDER { G_ASM, G_ACOR, G_CLUPROM, G_CLUSHADE, G_CLUTEND, G_CONTRAST, G_CORRELATION, G_DIFAVE, G_DIFENTRO, \
	G_DIFVAR, G_DIS, G_ENERGY, G_ENTROPY, G_HOM1, G_ID, G_IDN, G_IDM, G_IDMN, G_IV, G_JAVE, G_JE, G_INFOMEAS1, G_INFOMEAS2, \
	G_VARIANCE, G_JMAX, G_JVAR, G_SUMAVERAGE, G_SUMENTROPY, G_SUMVARIANCE }
constexpr int G_NAVE = 29;

// ---- GLRLM (featureset.h:236-268): 16 angled (column = COL_GLRLM + 4*f + angle) then 16 _AVE in the same order
enum GlrlmF : int
{
	R_SRE = 0, R_LRE, R_GLN, R_GLNN, R_RLN, R_RLNN, R_RP, R_GLV, R_RV, R_RE, R_LGLRE, R_HGLRE, R_SRLGLE, R_SRHGLE, R_LRLGLE,
	R_LRHGLE, R_NANGLED
};

// ---- GLDZM (featureset.h:271-288)
enum GldzmF : int
{
	Z_SDE = 0, Z_LDE, Z_LGLZE, Z_HGLZE, Z_SDLGLE, Z_SDHGLE, Z_LDLGLE, Z_LDHGLE, Z_GLNU, Z_GLNUN, Z_ZDNU, Z_ZDNUN, Z_ZP, Z_GLM,
	Z_GLV, Z_ZDM, Z_ZDV, Z_ZDE
};

// ---- GLSZM (featureset.h:291-306)
enum GlszmF : int
{
	S_SAE = 0, S_LAE, S_GLN, S_GLNN, S_SZN, S_SZNN, S_ZP, S_GLV, S_ZV, S_ZE, S_LGLZE, S_HGLZE, S_SALGLE, S_SAHGLE, S_LALGLE,
	S_LAHGLE
};

// ---- GLDM (featureset.h:309-322)
enum GldmF : int
{
	D_SDE = 0, D_LDE, D_GLN, D_DN, D_DNN, D_GLV, D_DV, D_DE, D_LGLE, D_HGLE, D_SDLGLE, D_SDHGLE, D_LDLGLE, D_LDHGLE
};

// ---- NGLDM (featureset.h:325-343)
enum NgldmF : int
{
	N_LDE = 0, N_HDE, N_LGLCE, N_HGLCE, N_LDLGLE, N_LDHGLE, N_HDLGLE, N_HDHGLE, N_GLNU, N_GLNUN, N_DCNU, N_DCNUN, N_DCP,
	N_GLM, N_GLV, N_DCM, N_DCV, N_DCENT, N_DCENE
};

// ---- NGTDM (featureset.h:346-350)
enum NgtdmF : int
{
	T_COARSENESS = 0, T_CONTRAST, T_BUSYNESS, T_COMPLEXITY, T_STRENGTH
};

} // namespace nyx
