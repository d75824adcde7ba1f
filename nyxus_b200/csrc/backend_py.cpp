// nyxus_b200 -- the pybind11 module `backend` over the C ABI (include/nyxus_b200.h).
//
// Same module name, same 23 functions, same positional signatures and the same return shapes as the reference's
// src/nyx/python/new_bindings_py.cpp:1090-1116, so the reference's own Python layer (src/nyx/python/nyxus/nyxus.py,
// `from .backend import ...`) drives the B200 path unmodified.  What is on the texture + intensity path is computed by
// libnyxus_b200.so; what is outside it (OME-Zarr / DICOM ingest, nested ROIs, Arrow writers, ROI blacklists)
// raises the way the reference raises for a build without that component.  There is no CPU fallback: everything that
// computes needs an sm_100 device, and `use_gpu_device = -1` (the reference's "CPU") means "device 0" here.
//
// State model (reference :25-46): one Env per Python object id, created on first touch, erased by
// release_environment_imp.  Like the reference, the per-family settings are COMPILED in initialize_environment
// (env_features.cpp:579-736): later set_environment_params_imp / set_if_ibsi_imp calls change what get_params_imp
// reports, not what the feature calculators use; set_metaparam_imp writes the compiled per-family values
// (env_metaparams.cpp:63-247).
#include <pybind11/pybind11.h>
#include <pybind11/numpy.h>
#include <pybind11/stl.h>
#include <zlib.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <filesystem>
#include <map>
#include <regex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <variant>
#include <vector>
#include "../../include/nyxus_b200.h"

namespace py = pybind11;
namespace fs = std::filesystem;

namespace
{

struct Env
{
	int dim = 2;
	std::vector<std::string> features;	// recognizedFeatureNames (what get_params reports)
	std::vector<std::string> compiled_features;	// what initialize_environment expanded
	int neighbor_distance = 5;
	float pixels_per_micron = 1.f;
	uint32_t coarse_gray_depth = 64;	// reported value
	uint32_t n_reduce_threads = 4;
	int device = 0;
	bool ibsi = false;		// reported value
	float dynamic_range = 10000.f, min_intensity = 0.f, max_intensity = 1.f;
	long long ram_limit_mb = -1;
	int verbosity = 0;
	bool single_roi = false;
	bool initialized = false;
	// compiled settings (fixed at initialize_environment, changed only by set_metaparam_imp)
	int c_greydepth = 64;
	bool c_ibsi = false;
	// per-family metaparameters: key of nyxb200_set_param -> value (env_metaparams.cpp)
	std::map<std::string, double> meta;
	// metaparameters the reference stores but this path does not read (reported back by get_metaparam_imp)
	std::map<std::string, double> inert;
	// Gabor raw strings (customize_gabor_feature_imp is called by every constructor, nyxus.py:246-254)
	std::string gabor[7];
	nyxb200_ctx* ctx = nullptr;
	bool ctx_stale = true;		// metaparameters changed since they were pushed to ctx
	std::string arrow_path;		// file written by the last arrowipc / parquet call (ArrowOutputStream::get_arrow_path)

	~Env()
	{
		if (ctx)
			nyxb200_destroy (ctx);
	}
};

std::unordered_map<uint64_t, Env> g_envs;

Env& findenv (uint64_t instid)
{
	return g_envs[instid];	// created if missing (reference :28-35)
}

void releaseenv (uint64_t instid)
{
	g_envs.erase (instid);
}

// context of an instance, created on first use (needs the device) and brought up to date with the metaparameters
nyxb200_ctx* ctx_for (Env& e)
{
	if (!e.initialized)
		throw std::runtime_error ("Error: the environment of this instance was not initialized");
	if (!e.ctx)
	{
		nyxb200_settings s;
		nyxb200_default_settings (&s);
		s.device = e.device;
		s.coarse_gray_depth = e.c_greydepth;
		s.ibsi = e.c_ibsi ? 1 : 0;
		s.verbose = e.verbosity;
		s.dim = e.dim;
		nyxb200_ctx* c = nullptr;
		if (nyxb200_create (&s, &c) != NYXB200_OK)
			throw std::runtime_error (std::string ("nyxus_b200 backend: ") + nyxb200_last_error (nullptr));
		std::vector<const char*> names;
		for (auto& f : e.compiled_features)
			names.push_back (f.c_str());
		const int rc = nyxb200_select_features (c, names.data(), (int)names.size());
		if (rc != NYXB200_OK)
		{
			std::string msg = nyxb200_last_error (c);
			nyxb200_destroy (c);
			if (rc == NYXB200_EINVAL)
				throw std::invalid_argument (msg);
			throw std::runtime_error (msg);
		}
		e.ctx = c;
		e.ctx_stale = true;
	}
	if (e.ctx_stale)
	{
		for (auto& [k, v] : e.meta)
			if (nyxb200_set_param (e.ctx, k.c_str(), v) != NYXB200_OK)
				throw std::runtime_error (std::string ("metaparameter ") + k + ": " + nyxb200_last_error (e.ctx));
		e.ctx_stale = false;
	}
	return e.ctx;
}

// ---- numeric table -> numpy without a copy: the pinned buffer is detached from the context and returned to the
// library's pool when the array dies (reference: as_pyarray, :63-73)
py::array table_to_numpy (nyxb200_ctx* c, nyxb200_table& t)
{
	const py::ssize_t nr = (py::ssize_t)t.nrows, nc = t.ncols;
	if (!t.data || nr == 0)
		return py::array_t<double> (std::vector<py::ssize_t> { 0, nc });
	if (nyxb200_table_detach (c, &t) != NYXB200_OK)
		throw std::runtime_error (nyxb200_last_error (c));
	py::capsule owner (t.data, [](void* p) { nyxb200_host_release (p); });
	return py::array_t<double> ({ nr, nc }, { (py::ssize_t)(nc * sizeof (double)), (py::ssize_t)sizeof (double) }, t.data, owner);
}

py::array header_array (nyxb200_ctx* c)
{
	// globals.h:38-41 + output_2_buffer.cpp:303-439
	std::vector<std::string> h { "intensity_image", "mask_image", "ROI_label", "t_index" };
	const int n = nyxb200_num_columns (c);
	for (int i = 0; i < n; i++)
		h.push_back (nyxb200_column_name (c, i));
	return py::array (py::cast (h));
}

// [nRows, 2] string columns: pure file names of the slide each row came from (output_2_buffer.cpp:455-465)
py::array string_columns (const std::vector<std::string>& inames, const std::vector<std::string>& lnames, const int32_t* row_image, int64_t nrows)
{
	std::vector<std::string> in, ln;
	for (auto& s : inames) in.push_back (fs::path (s).filename().string());
	for (auto& s : lnames) ln.push_back (fs::path (s).filename().string());
	py::module_ np = py::module_::import ("numpy");
	py::array idx = py::array_t<int32_t> ((py::ssize_t)nrows, row_image);
	py::object a = np.attr ("asarray") (py::cast (in)).attr ("__getitem__") (idx);
	py::object b = np.attr ("asarray") (py::cast (ln)).attr ("__getitem__") (idx);
	return np.attr ("stack") (py::make_tuple (a, b), py::arg ("axis") = 1);
}

[[noreturn]] void outside_path (const char* what)
{
	throw std::runtime_error (std::string (what) + " is outside the texture + intensity path of the nyxus_b200 backend");
}

// ---- Arrow IPC / Parquet output (reference: arrow_output_stream.cpp, output_writers.cpp, arrow_helpers.cpp).  The reference links
// libarrow; this module drives the interpreter's pyarrow (like its image decoding drives the interpreter's OpenCV / Pillow) and fails
// with the reference's own message where pyarrow is absent.
bool have_pyarrow()
{
	try
	{
		py::module_::import ("pyarrow");
		return true;
	}
	catch (py::error_already_set&)
	{
		return false;
	}
}

bool is_arrow_type (const std::string& output_type)
{
	return output_type == "arrowipc" || output_type == "parquet";
}

void check_output_type (const std::string& output_type)
{
	if (is_arrow_type (output_type) && !have_pyarrow())
		throw std::runtime_error ("Arrow functionality is not available. Rebuild Nyxus with Arrow enabled.");	// reference :1038,1052
}

bool ends_with (const std::string& s, const char* t)
{
	const size_t n = strlen (t);
	return s.size() >= n && s.compare (s.size() - n, n, t) == 0;
}

// arrow_helpers.cpp:9-62: directory (existing, or spelled with a trailing separator, or without extension under a missing parent)
// -> default file name inside it; the extension always follows the output type
std::string arrow_filename (const std::string& output_path, bool parquet)
{
	const char* ext = parquet ? ".parquet" : ".arrow";
	if (output_path.empty())
		return std::string ("NyxusFeatures") + ext;
	fs::path p (output_path);
	if (fs::is_directory (p) || ends_with (output_path, "/") || ends_with (output_path, "\\"))
		p = p / "NyxusFeatures";
	else if (!p.has_extension() && !fs::is_directory (p.parent_path()))
		p = p / "NyxusFeatures";
	p.replace_extension (ext);
	return p.string();
}

// header[0..1] utf8, header[2] (ROI_label) int32, the rest float64 (output_writers.cpp:21-30, 228-237); one record batch
void write_arrow (Env& e, const py::object& header, const py::object& strs, const py::object& num, const std::string& output_type,
	const std::string& output_path)
{
	const bool parquet = output_type == "parquet";
	const std::string path = arrow_filename (output_path, parquet);
	e.arrow_path = path;
	const fs::path parent = fs::path (path).parent_path();
	if (!parent.empty() && !fs::is_directory (parent))
		fs::create_directory (parent);		// arrow_output_stream.cpp:19-27 (one level, like the reference)
	py::module_ pa = py::module_::import ("pyarrow"), np = py::module_::import ("numpy");
	py::list names, arrays;
	const py::ssize_t ncol = py::len (header);
	const py::object all = py::slice (py::none(), py::none(), py::none());
	for (py::ssize_t k = 0; k < ncol; k++)
	{
		names.append (py::str (header.attr ("__getitem__") (k)));
		if (k < 2)
			arrays.append (pa.attr ("array") (strs.attr ("__getitem__") (py::make_tuple (all, k)).attr ("tolist") (), py::arg ("type") = pa.attr ("utf8") ()));
		else
		{
			py::object col = np.attr ("ascontiguousarray") (num.attr ("__getitem__") (py::make_tuple (all, k - 2)));
			if (k == 2)
				col = col.attr ("astype") (np.attr ("int32"));
			arrays.append (pa.attr ("array") (col));
		}
	}
	py::object table = pa.attr ("Table").attr ("from_arrays") (arrays, py::arg ("names") = names);
	if (parquet)
		py::module_::import ("pyarrow.parquet").attr ("write_table") (table, path);	// the Arrow schema is stored with the file (output_writers.cpp:40-43)
	else
	{
		py::object sink = pa.attr ("OSFile") (path, "wb");
		py::object writer = pa.attr ("ipc").attr ("new_file") (sink, table.attr ("schema"));
		writer.attr ("write_table") (table);
		writer.attr ("close") ();
		sink.attr ("close") ();
	}
}

// the tail of every featurize_*_imp: the buffers as they are, or an Arrow file and the reference's empty / one-element tuple
py::tuple finish_output (Env& e, const py::tuple& res, const std::string& output_type, const std::string& output_path, bool montage)
{
	if (!is_arrow_type (output_type))
		return res;
	write_arrow (e, res[0], res[1], res[2], output_type, output_path);
	if (montage)
		return py::make_tuple (std::string (""));	// new_bindings_py.cpp:643
	return py::tuple (0);				// new_bindings_py.cpp:559
}

// ---- minimal NIfTI-1 reader (single-file .nii / .nii.gz) for the Nyxus3D entry points.  The reference decodes with its
// vendored niftilib (raw_nifti.h:189-330) and hands integer voxel types over as unsigned grey values.
struct Volume
{
	int64_t W = 0, H = 0, D = 0;
	std::vector<uint32_t> vox;
};

std::vector<uint8_t> read_all (const std::string& path)
{
	std::vector<uint8_t> raw;
	gzFile f = gzopen (path.c_str(), "rb");	// transparent for files that are not gzip
	if (!f)
		throw std::runtime_error ("Error: cannot open file " + path);
	gzbuffer (f, 1 << 20);
	std::vector<uint8_t> chunk (1 << 24);
	for (;;)
	{
		int n = gzread (f, chunk.data(), (unsigned)chunk.size());
		if (n < 0)
		{
			gzclose (f);
			throw std::runtime_error ("Error: cannot read file " + path);
		}
		if (n == 0)
			break;
		raw.insert (raw.end(), chunk.begin(), chunk.begin() + n);
	}
	gzclose (f);
	return raw;
}

template <class T> T rd (const uint8_t* p, bool swap)
{
	uint8_t b[sizeof (T)];
	memcpy (b, p, sizeof (T));
	if (swap)
		std::reverse (b, b + sizeof (T));
	T v;
	memcpy (&v, b, sizeof (T));
	return v;
}

Volume read_nifti (const std::string& path, const char* what)
{
	const std::vector<uint8_t> raw = read_all (path);
	if (raw.size() < 352)
		throw std::runtime_error ("Error: " + path + " is not a NIfTI-1 file");
	bool swap = false;
	if (rd<int32_t> (raw.data(), false) != 348)
	{
		if (rd<int32_t> (raw.data(), true) != 348)
			throw std::runtime_error ("Error: " + path + ": bad NIfTI-1 header size");
		swap = true;
	}
	if (memcmp (raw.data() + 344, "n+1", 3) != 0)
		throw std::runtime_error ("Error: " + path + ": only single-file NIfTI-1 (magic n+1) is supported");
	int16_t dim[8];
	for (int i = 0; i < 8; i++)
		dim[i] = rd<int16_t> (raw.data() + 40 + 2 * i, swap);
	const int16_t datatype = rd<int16_t> (raw.data() + 70, swap);
	const size_t off = std::max<size_t> ((size_t)rd<float> (raw.data() + 108, swap), 352);
	const float slope = rd<float> (raw.data() + 112, swap), inter = rd<float> (raw.data() + 116, swap);
	const bool scaled = !(slope == 0.f || (slope == 1.f && inter == 0.f));
	Volume v;
	v.W = dim[1];
	v.H = dim[2];
	v.D = dim[0] >= 3 ? dim[3] : 1;
	if (v.W <= 0 || v.H <= 0 || v.D <= 0)
		throw std::runtime_error ("Error: " + path + ": bad volume dimensions");
	const size_t n = (size_t)v.W * v.H * v.D;
	size_t bpv = 0;
	switch (datatype)
	{
	case 2: case 256: bpv = 1; break;
	case 4: case 512: bpv = 2; break;
	case 8: case 768: case 16: bpv = 4; break;
	case 64: bpv = 8; break;
	default: throw std::runtime_error ("Error: " + path + ": unsupported NIfTI data type " + std::to_string (datatype));
	}
	if (raw.size() < off + n * bpv)
		throw std::runtime_error ("Error: " + path + ": file shorter than its header says");
	v.vox.resize (n);
	const uint8_t* p = raw.data() + off;
	auto put = [&](size_t i, double x)
	{
		if (scaled)
			x = x * slope + inter;
		if (!(x >= 0) || x != std::floor (x) || x > 4294967295.0)
			throw std::runtime_error (std::string ("Error: ") + what + " volume " + path + " holds negative or non-integer voxels: the "
				"floating-point dynamic-range remap and the Hounsfield shift of the reference's loaders are outside the nyxus_b200 path");
		v.vox[i] = (uint32_t)x;
	};
	for (size_t i = 0; i < n; i++)
		switch (datatype)
		{
		case 2: put (i, p[i]); break;
		case 256: put (i, (int8_t)p[i]); break;
		case 4: put (i, rd<int16_t> (p + 2 * i, swap)); break;
		case 512: put (i, rd<uint16_t> (p + 2 * i, swap)); break;
		case 8: put (i, rd<int32_t> (p + 4 * i, swap)); break;
		case 768: put (i, rd<uint32_t> (p + 4 * i, swap)); break;
		case 16: put (i, rd<float> (p + 4 * i, swap)); break;
		case 64: put (i, rd<double> (p + 8 * i, swap)); break;
		}
	return v;
}

// =====================================================================================================
// the 23 functions of the reference module
// =====================================================================================================

void initialize_environment (
	uint64_t instid,
	int n_dim,
	const std::vector<std::string>& features,
	int neighbor_distance,
	float pixels_per_micron,
	uint32_t coarse_gray_depth,
	uint32_t n_reduce_threads,
	int using_gpu,
	bool ibsi,
	float dynamic_range,
	float min_intensity,
	float max_intensity,
	bool is_imq,
	int ram_limit_mb,
	int verb_lvl,
	float aniso_x,
	float aniso_y,
	float aniso_z,
	bool merge_labels = false,
	bool preserve_hu = false)
{
	Env& e = findenv (instid);
	if (is_imq)
		outside_path ("image quality (ImageQuality class)");
	if (n_dim != 2 && n_dim != 3)
		throw std::invalid_argument ("n_dim must be 2 or 3");
	if (aniso_x != 1.f || aniso_y != 1.f || aniso_z != 1.f)
		outside_path ("anisotropy");
	if (merge_labels)
		outside_path ("mergerois");
	if (preserve_hu)
		outside_path ("preserve_hu");
	if (coarse_gray_depth == 0 || coarse_gray_depth > 65535)
		throw std::invalid_argument ("coarse_gray_depth must be in 1..65535");
	// Throws if an invalid feature is passed (reference :119-120, env_features.cpp:496,507) -- checked without a device
	std::vector<const char*> names;
	for (auto& f : features)
		names.push_back (f.c_str());
	char msg[1024];
	const int rc = nyxb200_validate_features (n_dim, names.data(), (int)names.size(), msg, sizeof (msg));
	if (rc == NYXB200_EINVAL)
		throw std::invalid_argument (msg);
	if (rc != NYXB200_OK)
		throw std::runtime_error (msg);
	if (e.ctx)
	{
		nyxb200_destroy (e.ctx);
		e.ctx = nullptr;
	}
	e.dim = n_dim;
	e.features = e.compiled_features = features;
	e.neighbor_distance = neighbor_distance;
	e.pixels_per_micron = pixels_per_micron;
	e.coarse_gray_depth = coarse_gray_depth;
	e.n_reduce_threads = n_reduce_threads;
	e.ibsi = ibsi;
	e.dynamic_range = dynamic_range;
	e.min_intensity = min_intensity;
	e.max_intensity = max_intensity;
	e.ram_limit_mb = ram_limit_mb;
	e.verbosity = verb_lvl;
	e.c_greydepth = (int)coarse_gray_depth;
	e.c_ibsi = ibsi;
	e.meta.clear();
	e.inert.clear();
	// -1 = the reference's "no GPU": this backend has no CPU path, it runs on device 0
	if (using_gpu >= 0)
	{
		if (using_gpu >= nyxb200_device_count())
			throw std::runtime_error ("nyxus_b200 backend: GPU device " + std::to_string (using_gpu) + " is not an available sm_100 device");
		e.device = using_gpu;
	}
	else
		e.device = 0;
	e.initialized = true;
}

void set_if_ibsi_imp (uint64_t instid, bool ibsi)
{
	findenv (instid).ibsi = ibsi;	// reference :156-160: the flag only; the compiled settings keep the constructor's value
}

void set_environment_params_imp (
	uint64_t instid,
	const std::vector<std::string>& features = {},
	int neighbor_distance = -1,
	float pixels_per_micron = -1,
	uint32_t coarse_gray_depth = 0,
	uint32_t n_reduce_threads = 0,
	int using_gpu = -2,
	float dynamic_range = -1,
	float min_intensity = -1,
	float max_intensity = -1,
	int ram_limit_mb = -1,
	int verb_level = 0)
{
	Env& e = findenv (instid);	// reference :162-222
	if (features.size() > 0) e.features = features;
	if (neighbor_distance > -1) e.neighbor_distance = neighbor_distance;
	if (pixels_per_micron > -1) e.pixels_per_micron = pixels_per_micron;
	if (coarse_gray_depth != 0) e.coarse_gray_depth = coarse_gray_depth;
	if (n_reduce_threads != 0) e.n_reduce_threads = n_reduce_threads;
	if (dynamic_range >= 0) e.dynamic_range = dynamic_range;
	if (min_intensity >= 0) e.min_intensity = min_intensity;
	if (max_intensity >= 0) e.max_intensity = max_intensity;
	if (verb_level >= 0) e.verbosity = verb_level;
	if (ram_limit_mb >= 0) e.ram_limit_mb = ram_limit_mb;
	(void)using_gpu;
}

// ---- 2D image files.  Decoding is host IO outside the hot path: the reference links libtiff / z5 / dcmtk, this module
// asks the interpreter's OpenCV (cv2.imread, IMREAD_UNCHANGED: tiled / stripped / compressed TIFF incl. OME-TIFF pages,
// PNG, ...) and falls back to Pillow.  Integer grey images only (the floating-point dynamic-range remap of the
// reference's loaders, image_loader.cpp, is outside this path).
py::array_t<uint32_t> decode_image (const std::string& path, const char* what)
{
	py::object arr = py::none();
	try
	{
		py::module_ cv2 = py::module_::import ("cv2");
		arr = cv2.attr ("imread") (path, -1);
	}
	catch (py::error_already_set&) {}
	if (arr.is_none())
	{
		try
		{
			py::module_ np = py::module_::import ("numpy");
			arr = np.attr ("asarray") (py::module_::import ("PIL.Image").attr ("open") (path));
		}
		catch (py::error_already_set&)
		{
			throw std::runtime_error (std::string ("Error: cannot decode ") + what + " image " + path + " (needs OpenCV or Pillow in the interpreter)");
		}
	}
	py::array a = py::array::ensure (arr);
	if (!a || a.ndim() != 2)
		throw std::runtime_error (std::string ("Error: ") + what + " image " + path + " is not a single-channel 2D image");
	const char kind = a.dtype().kind();
	if (kind == 'f')
		throw std::runtime_error (std::string ("Error: ") + what + " image " + path + " holds floating-point pixels: the dynamic-range remap of the reference's loaders is outside the nyxus_b200 path");
	py::module_ np = py::module_::import ("numpy");
	if (kind == 'i' && np.attr ("min") (a).cast<long long>() < 0)
		throw std::runtime_error (std::string ("Error: ") + what + " image " + path + " holds negative pixels (Hounsfield-type scale): outside the nyxus_b200 path");
	return py::array_t<uint32_t, py::array::c_style | py::array::forcecast>::ensure (a);
}

// image files of a list -> one table (reference: processDataset_2D_segmented / _wholeslide, workflow_2d_*.cpp): slides of one
// shape go through ONE pipelined list call; COVERED_IMAGE_INTENSITY_RANGE is taken against the scanned slide range like
// the file-based workflows do (slideprops.cpp:83-200), not against the sentinels of the in-memory montage path
py::tuple featurize_images (Env& e, const std::vector<std::string>& ifiles, const std::vector<std::string>& mfiles, bool single_roi)
{
	nyxb200_ctx* c = ctx_for (e);
	const size_t n = ifiles.size();
	std::vector<py::array_t<uint32_t>> I (n), M (n);
	bool same_shape = true;
	for (size_t k = 0; k < n; k++)
	{
		I[k] = decode_image (ifiles[k], "intensity");
		if (single_roi)
		{
			M[k] = py::array_t<uint32_t> ({ I[k].shape (0), I[k].shape (1) });
			std::fill (M[k].mutable_data(), M[k].mutable_data() + M[k].size(), 1u);		// whole slide = ROI 1
		}
		else
		{
			M[k] = decode_image (mfiles[k], "mask");
			if (M[k].shape (0) != I[k].shape (0) || M[k].shape (1) != I[k].shape (1))
				throw std::runtime_error ("Error: intensity / mask image size mismatch: " + ifiles[k] + " vs " + mfiles[k]);
		}
		same_shape = same_shape && I[k].shape (0) == I[0].shape (0) && I[k].shape (1) == I[0].shape (1);
	}
	struct Restore
	{
		nyxb200_ctx* c;
		~Restore() { nyxb200_set_param (c, "covered_range_from_scan", 0); nyxb200_set_param (c, "single_roi", 0); }
	} restore { c };
	nyxb200_set_param (c, "covered_range_from_scan", 1);
	nyxb200_set_param (c, "single_roi", single_roi ? 1 : 0);
	py::module_ np = py::module_::import ("numpy");
	py::array header = header_array (c);
	const std::vector<std::string> mnames = single_roi ? std::vector<std::string> (n, "") : mfiles;
	if (same_shape)
	{
		std::vector<const void*> ip (n);
		std::vector<const uint32_t*> lp (n);
		for (size_t k = 0; k < n; k++)
		{
			ip[k] = I[k].data();
			lp[k] = M[k].data();
		}
		nyxb200_table t {};
		const int32_t* row_image = nullptr;
		int rc;
		{
			py::gil_scoped_release nogil;
			rc = nyxb200_featurize_2d_list (c, ip.data(), 0, lp.data(), 0, (int)n, I[0].shape (0), I[0].shape (1), &t, &row_image);
		}
		if (rc != NYXB200_OK)
			throw std::runtime_error (std::string ("Error occurred during dataset processing: ") + nyxb200_last_error (c));
		const int64_t nrows = t.nrows;
		py::array strs = string_columns (ifiles, mnames, row_image, nrows);
		return py::make_tuple (header, strs, table_to_numpy (c, t));
	}
	py::list str_parts, num_parts;
	for (size_t k = 0; k < n; k++)
	{
		nyxb200_table t {};
		const int32_t* row_image = nullptr;
		int rc;
		{
			py::gil_scoped_release nogil;
			rc = nyxb200_featurize_2d_u32 (c, I[k].data(), M[k].data(), 1, I[k].shape (0), I[k].shape (1), &t, &row_image);
		}
		if (rc != NYXB200_OK)
			throw std::runtime_error (std::string ("Error occurred during dataset processing: ") + nyxb200_last_error (c));
		const int64_t nrows = t.nrows;
		str_parts.append (string_columns ({ ifiles[k] }, { mnames[k] }, row_image, nrows));
		num_parts.append (table_to_numpy (c, t));
	}
	return py::make_tuple (header, np.attr ("concatenate") (str_parts, py::arg ("axis") = 0), np.attr ("concatenate") (num_parts, py::arg ("axis") = 0));
}

py::tuple featurize_fname_lists_imp (uint64_t instid, const py::list& int_fnames, const py::list& seg_fnames, bool single_roi,
	const std::string& output_type, const std::string& output_path)
{
	Env& e = findenv (instid);
	if (e.dim != 2)
		throw std::runtime_error ("featurize_fname_lists_imp needs a 2D instance");
	e.single_roi = single_roi;
	check_output_type (output_type);
	std::vector<std::string> ifiles, mfiles;
	for (auto it : int_fnames) ifiles.push_back (it.cast<std::string>());
	// reference :664-676, :709-723
	if (ifiles.size() == 0)
		throw std::runtime_error ("Intensity file list is blank");
	for (auto& f : ifiles)
		if (!fs::exists (f))
			throw std::runtime_error ("File does not exist: " + f);
	if (!single_roi)
	{
		for (auto it : seg_fnames) mfiles.push_back (it.cast<std::string>());
		if (ifiles.size() != mfiles.size())
			throw std::runtime_error ("Imbalanced intensity (" + std::to_string (ifiles.size()) + " items) and segmentation mask (" + std::to_string (mfiles.size()) + " items) file lists");
		for (auto& f : mfiles)
			if (!fs::exists (f))
				throw std::runtime_error ("File does not exist: " + f);
	}
	return finish_output (e, featurize_images (e, ifiles, mfiles, single_roi), output_type, output_path, false);
}

py::tuple featurize_directory_imp (uint64_t instid, const std::string& intensity_dir, const std::string& labels_dir,
	const std::string& file_pattern, const std::string& output_type, const std::string& output_path = "")
{
	Env& e = findenv (instid);
	if (e.dim != 2)
		throw std::runtime_error ("featurize_directory_imp needs a 2D instance");
	check_output_type (output_type);
	std::regex rx;
	try
	{
		rx = std::regex (file_pattern);
	}
	catch (const std::regex_error&)
	{
		throw std::invalid_argument ("Invalid filepattern " + file_pattern);		// reference :236-237
	}
	const bool single = intensity_dir == labels_dir;	// whole-slide flag (reference :244-248)
	e.single_roi = single;
	std::vector<std::string> names;
	for (auto& ent : fs::directory_iterator (intensity_dir))
		if (ent.is_regular_file() && std::regex_match (ent.path().filename().string(), rx))
			names.push_back (ent.path().filename().string());
	std::sort (names.begin(), names.end());
	if (names.empty())
		throw std::runtime_error ("Error reading the dataset: no files in " + intensity_dir + " match " + file_pattern);
	std::vector<std::string> ifiles, mfiles;
	for (auto& nm : names)
	{
		ifiles.push_back ((fs::path (intensity_dir) / nm).string());
		if (!single)
		{
			mfiles.push_back ((fs::path (labels_dir) / nm).string());
			if (!fs::exists (mfiles.back()))
				throw std::runtime_error ("Error reading the dataset: missing mask file " + mfiles.back());
		}
	}
	return finish_output (e, featurize_images (e, ifiles, mfiles, single), output_type, output_path, false);
}

py::tuple featurize_montage_imp (
	uint64_t instid,
	const py::array_t<unsigned int, py::array::c_style | py::array::forcecast>& intensity_images,
	const py::array_t<unsigned int, py::array::c_style | py::array::forcecast>& label_images,
	const std::vector<std::string>& intensity_names,
	const std::vector<std::string>& label_names,
	const std::string output_type = "",
	const std::string output_path = "")
{
	Env& e = findenv (instid);
	if (e.dim != 2)
		throw std::runtime_error ("featurize_montage_imp needs a 2D instance");
	e.single_roi = false;
	check_output_type (output_type);
	auto ib = intensity_images.request(), lb = label_images.request();
	if (ib.ndim != 3 || lb.ndim != 3)
		throw std::invalid_argument ("Intensity and label arrays must be 3-dimensional [n, rows, cols]");
	const auto nf = ib.shape[0], d1 = ib.shape[1], d2 = ib.shape[2];
	if (nf != lb.shape[0])
		throw std::invalid_argument ("The number of intensity (" + std::to_string (nf) + ") and label (" + std::to_string (lb.shape[0]) + ") images must be the same.");
	if (d1 != lb.shape[1] || d2 != lb.shape[2])
		throw std::invalid_argument ("Intensity (width " + std::to_string (d1) + ", height " + std::to_string (d2) + ") and label (width "
			+ std::to_string (lb.shape[1]) + ", height " + std::to_string (lb.shape[2]) + ") image size mismatch");
	if ((py::ssize_t)intensity_names.size() != nf || (py::ssize_t)label_names.size() != nf)
		throw std::invalid_argument ("The number of image names must be the same as the number of images");
	nyxb200_ctx* c = ctx_for (e);
	nyxb200_table t {};
	const int32_t* row_image = nullptr;
	int rc = NYXB200_OK;
	if (nf > 0 && d1 > 0 && d2 > 0)
	{
		py::gil_scoped_release nogil;	// the reference holds the GIL for the whole call; nothing here touches Python
		rc = nyxb200_featurize_2d_u32 (c, (const uint32_t*)ib.ptr, (const uint32_t*)lb.ptr, (int)nf, (int64_t)d1, (int64_t)d2, &t, &row_image);
	}
	if (rc != NYXB200_OK)
		throw std::runtime_error (std::string ("Error occurred during dataset processing: ") + nyxb200_last_error (c));	// reference :619-620
	const int64_t nrows = t.nrows;
	py::array header = header_array (c);
	py::array strs = string_columns (intensity_names, label_names, row_image, nrows);
	py::array num = table_to_numpy (c, t);
	return finish_output (e, py::make_tuple (header, strs, num, std::string ("")), output_type, output_path, true);
}

// volumes of a file list -> one table (reference: processDataset_3D_segmented, workflow_3d_segmented.cpp:35-260)
py::tuple featurize_volumes (Env& e, const std::vector<std::string>& ifiles, const std::vector<std::string>& mfiles, bool single_roi)
{
	nyxb200_ctx* c = ctx_for (e);
	const size_t nvol = ifiles.size();
	std::vector<Volume> I (nvol), M (nvol);
	bool same_shape = true;
	for (size_t k = 0; k < nvol; k++)
	{
		I[k] = read_nifti (ifiles[k], "intensity");
		if (single_roi)
		{
			M[k].W = I[k].W; M[k].H = I[k].H; M[k].D = I[k].D;
			M[k].vox.assign (I[k].vox.size(), 1u);	// whole volume = ROI 1
		}
		else
		{
			M[k] = read_nifti (mfiles[k], "mask");
			if (M[k].W != I[k].W || M[k].H != I[k].H || M[k].D != I[k].D)
				throw std::runtime_error ("Error: intensity / mask volume size mismatch: " + ifiles[k] + " vs " + mfiles[k]);
		}
		same_shape = same_shape && I[k].W == I[0].W && I[k].H == I[0].H && I[k].D == I[0].D;
	}
	py::module_ np = py::module_::import ("numpy");
	py::array header = header_array (c);
	std::vector<std::string> mnames = single_roi ? std::vector<std::string> (nvol, "") : mfiles;
	if (same_shape)
	{
		// one pipelined call over the whole list (upload of volume k+1 while volume k computes)
		std::vector<const void*> ip (nvol);
		std::vector<const uint32_t*> lp (nvol);
		for (size_t k = 0; k < nvol; k++)
		{
			ip[k] = I[k].vox.data();
			lp[k] = M[k].vox.data();
		}
		nyxb200_table t {};
		const int32_t* row_image = nullptr;
		int rc;
		{
			py::gil_scoped_release nogil;
			rc = nyxb200_featurize_3d_list (c, ip.data(), 0, lp.data(), 0, (int)nvol, I[0].D, I[0].H, I[0].W, &t, &row_image);
		}
		if (rc != NYXB200_OK)
			throw std::runtime_error (std::string ("Error occurred during dataset processing: ") + nyxb200_last_error (c));
		const int64_t nrows = t.nrows;
		py::array strs = string_columns (ifiles, mnames, row_image, nrows);
		return py::make_tuple (header, strs, table_to_numpy (c, t));
	}
	py::list str_parts, num_parts;
	for (size_t k = 0; k < nvol; k++)
	{
		nyxb200_table t {};
		const int32_t* row_image = nullptr;
		int rc;
		{
			py::gil_scoped_release nogil;
			rc = nyxb200_featurize_3d_u32 (c, I[k].vox.data(), M[k].vox.data(), 1, I[k].D, I[k].H, I[k].W, &t, &row_image);
		}
		if (rc != NYXB200_OK)
			throw std::runtime_error (std::string ("Error occurred during dataset processing: ") + nyxb200_last_error (c));
		const int64_t nrows = t.nrows;
		str_parts.append (string_columns ({ ifiles[k] }, { mnames[k] }, row_image, nrows));
		num_parts.append (table_to_numpy (c, t));
	}
	return py::make_tuple (header, np.attr ("concatenate") (str_parts, py::arg ("axis") = 0), np.attr ("concatenate") (num_parts, py::arg ("axis") = 0));
}

py::tuple featurize_fname_lists_3D_imp (
	uint64_t instid,
	const py::list& pyside_int_fnames,
	const py::list& pyside_seg_fnames,
	bool single_roi,
	const std::string& output_type,
	const std::string& output_path)
{
	Env& e = findenv (instid);
	if (e.dim != 3)
		throw std::runtime_error ("featurize_fname_lists_3D_imp needs a 3D instance");
	e.single_roi = single_roi;
	check_output_type (output_type);
	std::vector<std::string> ifiles, mfiles;
	for (auto it : pyside_int_fnames) ifiles.push_back (it.cast<std::string>());
	for (auto it : pyside_seg_fnames) mfiles.push_back (it.cast<std::string>());
	// reference :779-801
	if (ifiles.size() == 0)
		throw std::runtime_error ("Intensity file list is blank");
	if (!single_roi)
	{
		if (mfiles.size() == 0)
			throw std::runtime_error ("Mask file list is blank");
		if (ifiles.size() != mfiles.size())
			throw std::runtime_error ("Imbalanced intensity and segmentation mask file lists");
	}
	for (size_t i = 0; i < ifiles.size(); i++)
	{
		if (!fs::exists (ifiles[i]))
			throw std::runtime_error ("File does not exist: " + ifiles[i]);
		if (!single_roi && !fs::exists (mfiles[i]))
			throw std::runtime_error ("File does not exist: " + mfiles[i]);
	}
	return finish_output (e, featurize_volumes (e, ifiles, mfiles, single_roi), output_type, output_path, false);
}

py::tuple featurize_directory_3D_imp (
	uint64_t instid,
	const std::string& intensity_dir,
	const std::string& labels_dir,
	const std::string& file_pattern,
	const std::string& output_type,
	const std::string& output_path = "")
{
	Env& e = findenv (instid);
	if (e.dim != 3)
		throw std::runtime_error ("featurize_directory_3D_imp needs a 3D instance");
	check_output_type (output_type);
	std::regex rx;
	try
	{
		rx = std::regex (file_pattern);
	}
	catch (const std::regex_error&)
	{
		throw std::invalid_argument ("Invalid filepattern " + file_pattern);
	}
	const bool single = intensity_dir == labels_dir;	// whole-volume flag (reference :244-248 convention)
	e.single_roi = single;
	std::vector<std::string> names;
	for (auto& ent : fs::directory_iterator (intensity_dir))
		if (ent.is_regular_file() && std::regex_match (ent.path().filename().string(), rx))
			names.push_back (ent.path().filename().string());
	std::sort (names.begin(), names.end());
	if (names.empty())
		throw std::runtime_error ("Error reading the dataset: no files in " + intensity_dir + " match " + file_pattern);
	std::vector<std::string> ifiles, mfiles;
	for (auto& n : names)
	{
		ifiles.push_back ((fs::path (intensity_dir) / n).string());
		if (!single)
		{
			mfiles.push_back ((fs::path (labels_dir) / n).string());
			if (!fs::exists (mfiles.back()))
				throw std::runtime_error ("Error reading the dataset: missing mask file " + mfiles.back());
		}
	}
	return finish_output (e, featurize_volumes (e, ifiles, mfiles, single), output_type, output_path, false);
}

py::tuple findrelations_imp (uint64_t, std::string&, std::string&, std::string&)
{
	outside_path ("nested ROI relations (Nested class)");
}

bool gpu_available_imp (uint64_t)
{
	return nyxb200_device_count() > 0;
}

void use_gpu (uint64_t instid, bool yes)
{
	(void)instid;
	if (!yes)
		throw std::runtime_error ("the nyxus_b200 backend has no CPU path: features are always computed on the GPU");
}

std::vector<std::map<std::string, std::string>> get_gpu_properties (uint64_t)
{
	std::vector<std::map<std::string, std::string>> out;
	const int n = nyxb200_device_count();
	for (int d = 0; d < n; d++)
	{
		char name[256];
		int64_t mem = 0;
		int sms = 0, major = 0, minor = 0;
		if (nyxb200_device_info (d, name, sizeof (name), &mem, &sms, &major, &minor) != NYXB200_OK)
			continue;
		out.push_back ({ { "Device number", std::to_string (d) }, { "CUDA version", std::to_string (major) + "." + std::to_string (minor) },
			{ "Device name", name }, { "Total global memory", std::to_string (mem) }, { "Multiprocessors", std::to_string (sms) } });
	}
	return out;
}

void blacklist_roi_imp (uint64_t, std::string)
{
	outside_path ("ROI blacklisting");
}

void clear_roi_blacklist_imp (uint64_t) {}

py::str roi_blacklist_get_summary_imp (uint64_t)
{
	return py::str ("");
}

void customize_gabor_feature_imp (uint64_t instid, const std::string& kersize, const std::string& gamma, const std::string& sig2lam,
	const std::string& f0, const std::string& theta, const std::string& thold, const std::string& freqs)
{
	// called by every Nyxus constructor (nyxus.py:246-254); Gabor features are not on this path, the strings are kept
	Env& e = findenv (instid);
	const std::string* v[7] = { &kersize, &gamma, &sig2lam, &f0, &theta, &thold, &freqs };
	for (int i = 0; i < 7; i++)
		e.gabor[i] = *v[i];
}

using ParameterTypes = std::variant<int, float, double, unsigned int, std::vector<double>, std::vector<std::string>>;

std::vector<double> parse_doubles (const std::string& s)
{
	std::vector<double> v;
	size_t i = 0;
	while (i < s.size())
	{
		size_t j = s.find (',', i);
		if (j == std::string::npos)
			j = s.size();
		try
		{
			v.push_back (std::stod (s.substr (i, j - i)));
		}
		catch (...) {}
		i = j + 1;
	}
	return v;
}

std::map<std::string, ParameterTypes> get_params_imp (uint64_t instid, const std::vector<std::string>& vars)
{
	Env& e = findenv (instid);	// reference :983-1031
	auto num = [](const std::string& s, double dflt) { try { return std::stod (s); } catch (...) { return dflt; } };
	std::map<std::string, ParameterTypes> params;
	params["features"] = e.features;
	params["neighbor_distance"] = e.neighbor_distance;
	params["pixels_per_micron"] = e.pixels_per_micron;
	params["coarse_gray_depth"] = e.coarse_gray_depth;
	params["n_feature_calc_threads"] = e.n_reduce_threads;
	params["ibsi"] = (int)e.ibsi;
	params["gabor_kersize"] = (int)num (e.gabor[0], 16);
	params["gabor_gamma"] = num (e.gabor[1], 0.1);
	params["gabor_sig2lam"] = num (e.gabor[2], 0.8);
	params["gabor_f0"] = num (e.gabor[3], 0.1);
	params["gabor_thold"] = num (e.gabor[5], 0.025);
	params["gabor_freqs"] = parse_doubles (e.gabor[6]);
	params["gabor_thetas"] = parse_doubles (e.gabor[4]);
	params["dynamic_range"] = e.dynamic_range;
	params["min_intensity"] = e.min_intensity;
	params["max_intensity"] = e.max_intensity;
	params["ram_limit"] = (int)e.ram_limit_mb;
	if (vars.empty())
		return params;
	std::map<std::string, ParameterTypes> subset;
	for (const auto& v : vars)
	{
		auto it = params.find (v);
		if (it != params.end())
			subset.insert (*it);
	}
	return subset;
}

std::string get_arrow_file_imp (uint64_t instid)
{
	if (!have_pyarrow())
		throw std::runtime_error ("Arrow functionality is not available. Rebuild Nyxus with Arrow enabled.");
	return findenv (instid).arrow_path;
}

std::string get_parquet_file_imp (uint64_t instid)
{
	return get_arrow_file_imp (instid);
}

bool arrow_is_enabled_imp (uint64_t)
{
	return have_pyarrow();
}

// ---- metaparameters (env_metaparams.cpp:63-247): "<feature>/<param>=<value>"
std::vector<std::string> split (const std::string& s, char d)
{
	std::vector<std::string> v;
	size_t i = 0;
	while (i <= s.size())
	{
		size_t j = s.find (d, i);
		if (j == std::string::npos)
			j = s.size();
		if (j > i)
			v.push_back (s.substr (i, j - i));
		i = j + 1;
	}
	return v;
}

bool parse_int (const std::string& s, int& v)
{
	try
	{
		size_t pos = 0;
		v = std::stoi (s, &pos);
		return pos == s.size();
	}
	catch (...)
	{
		return false;
	}
}

// which (feature, parameter) pairs exist in the reference, whether this path reads them (key of nyxb200_set_param) or only stores them
struct MetaDef { const char* feature; const char* param; const char* b200_key; bool boolean; bool positive; };
const MetaDef g_meta[] = {
	{ "glcm", "greydepth", "glcm/greydepth", false, false },
	{ "glcm", "offset", "glcm/offset", false, false },
	{ "3glcm", "greydepth", "3glcm/greydepth", false, false },
	{ "3glcm", "offset", nullptr, false, false },		// D3_GLCM_feature uses a static offset of 1 (3d_glcm.cpp:16-31): stored, never read
	{ "3glcm", "numang", nullptr, false, false },
	{ "3glcm", "sparseintensities", nullptr, true, false },
	{ "3gldm", "greydepth", "3gldm/greydepth", false, false },
	{ "3ngtdm", "greydepth", "3ngtdm/greydepth", false, false },
	{ "3ngtdm", "radius", "3ngtdm/radius", false, true },
	{ "3glrlm", "greydepth", "3glrlm/greydepth", false, false },
	{ "3glszm", "greydepth", "3glszm/greydepth", false, false },
};

py::tuple set_metaparam_imp (uint64_t instid, const std::string p_val)
{
	Env& e = findenv (instid);
	auto err = [](const std::string& m) { return py::make_tuple (false, m); };
	const std::vector<std::string> sides = split (p_val, '=');
	if (sides.size() != 2)
		return err ("syntax error in \"" + p_val + "\": expecting <paramName>=<paramVal>");
	const std::vector<std::string> path = split (sides[0], '/');
	if (!(path.size() == 1 || path.size() == 2))
		return err ("syntax error in <paramName>=<paramVal> of \"" + p_val + "\": expecting <paramName> to be <feature name>/<parameter name> or <common parameter name>");
	if (path.size() == 1)
		return py::make_tuple (true, "success");	// the reference accepts and ignores common parameter names (:65-247 has no branch for them)
	bool feature_known = false;
	for (const MetaDef& m : g_meta)
	{
		if (path[0] != m.feature)
			continue;
		feature_known = true;
		if (path[1] != m.param)
			continue;
		double v;
		if (m.boolean)
		{
			std::string s = sides[1];
			std::transform (s.begin(), s.end(), s.begin(), ::tolower);
			if (s != "true" && s != "false")
				return err ("error: cannot parse value \"" + sides[1] + "\" of " + sides[0] + ": expecting a boolean (\"true\" or \"false\")");
			v = s == "true";
		}
		else
		{
			int iv;
			if (!parse_int (sides[1], iv) || (m.positive && iv <= 0))
				return err ("error: cannot parse value \"" + sides[1] + "\" of " + sides[0] + ": expecting " + (m.positive ? "a positive integer" : "an integer"));
			v = iv;
		}
		const std::string key = std::string (m.feature) + "/" + m.param;
		// a 3D key on a 2D instance (or the reverse) is stored by the reference and read by nothing
		const bool applies = m.b200_key && ((m.feature[0] == '3') == (e.dim == 3));
		if (applies)
		{
			if (e.ctx && nyxb200_set_param (e.ctx, m.b200_key, v) != NYXB200_OK)
				return err (std::string ("error: ") + nyxb200_last_error (e.ctx));
			e.meta[m.b200_key] = v;
		}
		e.inert[key] = v;
		return py::make_tuple (true, "success");
	}
	if (!feature_known)
		return err ("error: unrecognized feature \"" + path[0] + "\"");
	return err ("error: unrecognized feature parameter of feature " + path[0] + ": \"" + path[1] + "\"");
}

py::tuple get_metaparam_imp (uint64_t instid, const std::string p_name)
{
	Env& e = findenv (instid);
	const std::vector<std::string> path = split (p_name, '/');
	if (!(path.size() == 1 || path.size() == 2))
		return py::make_tuple (false, "syntax error in <paramName>=<paramVal> of \"" + p_name + "\": expecting <paramName> to be <feature name>/<parameter name> or <common parameter name>");
	if (path.size() == 1)
		return py::make_tuple (0.0, "");
	bool feature_known = false;
	for (const MetaDef& m : g_meta)
	{
		if (path[0] != m.feature)
			continue;
		feature_known = true;
		if (path[1] != m.param)
			continue;
		const std::string key = std::string (m.feature) + "/" + m.param;
		auto it = e.inert.find (key);
		double v = 0;
		if (it != e.inert.end())
			v = it->second;
		else if (!strcmp (m.param, "greydepth") && strstr (m.feature, "glcm"))
			v = e.c_greydepth;	// GLCM_GREYDEPTH defaults to the coarse grey depth (env_features.cpp:730)
		else if (!strcmp (m.param, "offset"))
			v = 1;			// env_features.cpp:727
		else if (!strcmp (m.param, "numang"))
			v = e.dim == 3 ? 13 : 4;
		return py::make_tuple (v, "");
	}
	if (!feature_known)
		return py::make_tuple (false, "error: unrecognized feature \"" + path[0] + "\"");
	return py::make_tuple (false, "error: unrecognized feature parameter of feature " + path[0] + ": \"" + path[1] + "\"");
}

} // namespace

PYBIND11_MODULE (backend, m)
{
	m.doc() = "Nyxus (nyxus_b200 backend: B200-native texture + intensity path behind the reference's pybind interface)";
	m.def ("initialize_environment", &initialize_environment, "Environment initialization",
		py::arg ("instid"), py::arg ("n_dim"), py::arg ("features"), py::arg ("neighbor_distance"), py::arg ("pixels_per_micron"),
		py::arg ("coarse_gray_depth"), py::arg ("n_reduce_threads"), py::arg ("using_gpu"), py::arg ("ibsi"), py::arg ("dynamic_range"),
		py::arg ("min_intensity"), py::arg ("max_intensity"), py::arg ("is_imq"), py::arg ("ram_limit_mb"), py::arg ("verb_lvl"),
		py::arg ("aniso_x"), py::arg ("aniso_y"), py::arg ("aniso_z"), py::arg ("merge_labels") = false, py::arg ("preserve_hu") = false);
	m.def ("release_environment_imp", &releaseenv, "Forget an instance's environment (called from __del__; id(self) is reused by CPython)");
	m.def ("featurize_directory_imp", &featurize_directory_imp, "Calculate features of images defined by intensity and mask image collection directories");
	m.def ("featurize_directory_3D_imp", &featurize_directory_3D_imp, "Calculate 3D features of images defined by intensity and mask image collection directories");
	m.def ("featurize_montage_imp", &featurize_montage_imp, "Calculate features of images defined by intensity and mask image collection directories");
	m.def ("featurize_fname_lists_imp", &featurize_fname_lists_imp, "Calculate features of intensity-mask image pairs defined by lists of image file names");
	m.def ("featurize_fname_lists_3D_imp", &featurize_fname_lists_3D_imp, "Calculate 3D features of intensity-mask volume pairs defined by lists of file names");
	m.def ("findrelations_imp", &findrelations_imp, "Find relations in segmentation mask images");
	m.def ("gpu_available_imp", &gpu_available_imp, "Check if CUDA gpu is available");
	m.def ("use_gpu", &use_gpu, "Enable/disable GPU features");
	m.def ("get_gpu_props", &get_gpu_properties, "Get properties of the active CUDA gpu device");
	m.def ("blacklist_roi_imp", &blacklist_roi_imp, "Set up a global or per-mask file blacklist definition");
	m.def ("clear_roi_blacklist_imp", &clear_roi_blacklist_imp, "Clear the ROI black list");
	m.def ("roi_blacklist_get_summary_imp", &roi_blacklist_get_summary_imp, "Returns a summary of the ROI blacklist");
	m.def ("customize_gabor_feature_imp", &customize_gabor_feature_imp, "Sets custom GABOR feature's parameters");
	m.def ("set_if_ibsi_imp", &set_if_ibsi_imp, "Set if the features will be ibsi compliant");
	m.def ("set_environment_params_imp", &set_environment_params_imp, "Set the environment variables of Nyxus");
	m.def ("get_params_imp", &get_params_imp, "Get parameters of Nyxus");
	m.def ("arrow_is_enabled_imp", &arrow_is_enabled_imp, "Check if arrow is enabled.");
	m.def ("get_arrow_file_imp", &get_arrow_file_imp, "Get path to arrow file");
	m.def ("get_parquet_file_imp", &get_parquet_file_imp, "Returns path to parquet file");
	m.def ("set_metaparam_imp", &set_metaparam_imp, "Setting a common or feature-specific metaparameter");
	m.def ("get_metaparam_imp", &get_metaparam_imp, "Getting a common or feature-specific metaparameter value");
}
