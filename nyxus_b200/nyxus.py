"""Host-side mirror of the reference's Python entry points for the texture + intensity path.

`Nyxus(features, **kwargs).featurize(intensity_images, label_images, ...)` keeps the reference's signature, argument
checks, error messages, DataFrame layout and column names (reference src/nyx/python/nyxus/nyxus.py:157-261, 385-518),
but the work is done by the B200 library through the C ABI (include/nyxus_b200.h).  Feature names outside the
texture/intensity path raise ValueError exactly like an unknown feature name does in the reference
(env_features.cpp:496)."""
import os
from typing import List, Optional

import numpy as np
import pandas as pd

from .engine import Engine

_VALID_KEYS = {
    'neighbor_distance', 'pixels_per_micron', 'coarse_gray_depth',
    'n_feature_calc_threads', 'use_gpu_device', 'ibsi',
    'gabor_kersize', 'gabor_gamma', 'gabor_sig2lam', 'gabor_f0',
    'gabor_thold', 'gabor_thetas', 'gabor_freqs', 'channel_signature',
    'parent_channel', 'child_channel', 'aggregate', 'dynamic_range', 'min_intensity',
    'max_intensity', 'ram_limit', 'verbose',
    'anisotropy_x', 'anisotropy_y', 'mergerois', 'preserve_hu'
}



# ---- Arrow IPC / Parquet output (reference arrow_helpers.cpp:9-62, output_writers.cpp:21-30,228-237): written with pyarrow
def _need_pyarrow():
    try:
        import pyarrow  # noqa: F401
    except ImportError:
        raise RuntimeError("Arrow functionality is not available. Rebuild Nyxus with Arrow enabled.")


def arrow_filename(output_path: str, parquet: bool) -> str:
    """the reference's get_arrow_filename: a directory (existing, spelled with a trailing separator, or extension-less under a
    missing parent) receives NyxusFeatures.<ext>; the extension always follows the output type"""
    ext = ".parquet" if parquet else ".arrow"
    if not output_path:
        return "NyxusFeatures" + ext
    p = output_path
    if os.path.isdir(p) or p.endswith("/") or p.endswith("\\"):
        p = os.path.join(p, "NyxusFeatures")
    elif not os.path.splitext(os.path.basename(p))[1] and not os.path.isdir(os.path.dirname(p)):
        p = os.path.join(p, "NyxusFeatures")
    return os.path.splitext(p)[0] + ext if os.path.splitext(os.path.basename(p))[1] else p + ext


def _emit(obj, df, output_type, output_path):
    """the DataFrame, or the path of the Arrow IPC / Parquet file it was written to (utf8, utf8, int32, float64...)"""
    if output_type == 'pandas':
        return df
    import pyarrow as pa
    path = arrow_filename(output_path, output_type == 'parquet')
    parent = os.path.dirname(path)
    if parent and not os.path.isdir(parent):
        os.mkdir(parent)
    cols = [pa.array(df[c].astype(str).tolist(), type=pa.utf8()) for c in df.columns[:2]]
    cols.append(pa.array(df[df.columns[2]].to_numpy().astype(np.int32)))
    cols += [pa.array(np.ascontiguousarray(df[c].to_numpy(np.float64))) for c in df.columns[3:]]
    table = pa.Table.from_arrays(cols, names=[str(c) for c in df.columns])
    if output_type == 'parquet':
        import pyarrow.parquet as pq
        pq.write_table(table, path)
    else:
        with pa.OSFile(path, "wb") as sink:
            with pa.ipc.new_file(sink, table.schema) as w:
                w.write_table(table)
    obj._arrow_path = path
    return path


class Nyxus:
    """2D feature extraction (reference nyxus.py:23-518), texture + intensity path on a B200.

    Differences from the reference, all deliberate:
      * `use_gpu_device` selects the B200 the path runs on (default -1 means device 0): there is no CPU path here.
      * `*ALL_TEXTURE*` is accepted and means the union of the seven 2D texture groups.
      * `n_feature_calc_threads` is accepted and ignored (ROI parallelism is the GPU grid).
    """

    def __init__(self, features: List[str], **kwargs):
        invalid_keys = set(kwargs.keys()) - _VALID_KEYS
        if invalid_keys:
            print(f"Warning: unexpected keyword argument(s): {', '.join(invalid_keys)}")
        neighbor_distance = kwargs.get('neighbor_distance', 5)
        pixels_per_micron = kwargs.get('pixels_per_micron', 1.0)
        coarse_gray_depth = kwargs.get('coarse_gray_depth', 64)
        n_feature_calc_threads = kwargs.get('n_feature_calc_threads', 4)
        use_gpu_device = kwargs.get('use_gpu_device', -1)
        ibsi = kwargs.get('ibsi', False)
        verb_lvl = kwargs.get('verbose', 0)
        aniso_x = kwargs.get('anisotropy_x', 1.0)
        aniso_y = kwargs.get('anisotropy_y', 1.0)
        mergerois = kwargs.get('mergerois', False)

        if neighbor_distance <= 0:
            raise ValueError("Neighbor distance must be greater than zero.")
        if pixels_per_micron <= 0:
            raise ValueError("Pixels per micron must be greater than zero.")
        if coarse_gray_depth <= 0:
            raise ValueError("Custom number of grayscale levels (parameter coarse_gray_depth, default=64) must be non-negative.")
        if n_feature_calc_threads < 1:
            raise ValueError("There must be at least one feature calculation thread.")
        if verb_lvl < 0:
            raise ValueError("verbosity must be non-negative")
        if aniso_x <= 0:
            raise ValueError("anisotropy_x must be positive")
        if aniso_y <= 0:
            raise ValueError("anisotropy_y must be positive")
        if aniso_x != 1.0 or aniso_y != 1.0:
            raise ValueError("anisotropy is outside the texture/intensity path implemented by nyxus_b200")

        self._mergerois = bool(mergerois)
        self._engine = Engine(list(features), device=max(int(use_gpu_device), 0), coarse_gray_depth=coarse_gray_depth,
                              ibsi=ibsi, verbose=verb_lvl)
        self._valid_output_types = ['pandas', 'arrowipc', 'parquet']
        self.error_message = ''

    # reference nyxus.py:1067-1086 set_metaparam_imp / get_metaparam_imp for the keys this path reads
    def set_metaparam(self, expr: str):
        key, _, val = expr.partition('=')
        self._engine.set_param(key.strip(), float(val))

    def set_params(self, **params):
        for k, v in params.items():
            if k == 'coarse_gray_depth':
                self._engine.set_param('coarse_gray_depth', v)
            elif k == 'ibsi':
                self._engine.set_param('ibsi', 1.0 if v else 0.0)
            elif k == 'features':
                self._engine.select(list(v))
            else:
                raise ValueError(f"parameter {k} is outside the texture/intensity path implemented by nyxus_b200")

    @property
    def feature_columns(self):
        return self._engine.columns

    # reference nyxus.py:870-920
    def get_arrow_ipc_file(self):
        return getattr(self, "_arrow_path", "")

    def get_parquet_file(self):
        return getattr(self, "_arrow_path", "")

    def arrow_is_enabled(self):
        try:
            import pyarrow  # noqa: F401
            return True
        except ImportError:
            return False

    def featurize(self, intensity_images: np.ndarray, label_images: np.ndarray, intensity_names: list = [],
                  label_names: list = [], output_type: Optional[str] = "pandas", output_path: Optional[str] = ""):
        if (output_type != "" and output_type not in self._valid_output_types):
            raise ValueError(f'Invalid output type: {output_type}. Valid options are: {self._valid_output_types}')
        if not isinstance(intensity_images, np.ndarray):
            raise ValueError("intensity_images parameter must be numpy.ndarray")
        if not isinstance(label_images, np.ndarray):
            raise ValueError("label_images parameter must be numpy.ndarray")
        if (output_type not in self._valid_output_types):
            raise ValueError(f'Invalid output type {output_type}. Valid output types are {self._valid_output_types}.')
        if output_type != 'pandas':
            _need_pyarrow()

        if intensity_images.ndim == 2:
            if label_images.ndim != 2:
                raise ValueError("Both intensity and label arrays must be the same dimension")
            single = True
            intensity_images = intensity_images[None]      # a view: the reference's np.array([...]) copies the slide (nyxus.py:447-448)
            label_images = label_images[None]
        elif intensity_images.ndim == 3:
            if label_images.ndim != 3:
                raise ValueError("Both intensity and label arrays must be the same dimension")
            single = False
        else:
            raise ValueError("Intensity and label arrays must be 2D or 3D")
        if intensity_images.shape != label_images.shape:
            raise ValueError("Intensity and label image arrays must have the same number of images with matching dimensions")
        if intensity_names == []:
            intensity_names = ["Intensity" + str(i) for i in range(intensity_images.shape[0])]
        if label_names == []:
            label_names = ["Segmentation" + str(i) for i in range(label_images.shape[0])]
        if intensity_images.shape[0] != len(intensity_names):
            raise ValueError("Number of _intensity image names_ (" + str(len(intensity_names)) + ") must be the same as the number of intensity images (" + str(intensity_images.shape[0]) + ")")
        if label_images.shape[0] != len(label_names):
            raise ValueError("Number of segmentation names must be the same as the number of images.")

        # Hounsfield-type scale: shift into the unsigned range like the reference (nyxus.py:481-484: IN PLACE on a stack the
        # caller passed, on its own copy for a single 2D slide).  Unsigned types cannot be negative: no pass over them.
        I = intensity_images
        if not np.issubdtype(I.dtype, np.unsignedinteger):
            min_raw_I = np.min(I)
            if min_raw_I < 0:
                if single:
                    I = I - min_raw_I
                else:
                    I -= min_raw_I
        if I.dtype != np.uint16 and I.dtype != np.uint32:     # uint16 goes to the device as is (no widening copy)
            I = I.astype(np.uint32)
        M = label_images.astype(np.uint32, copy=False)        # no copy when the mask already is uint32
        if self._mergerois:                                   # phase1.cpp:392-393, phase2_2d.cpp:665-666
            M = (M != 0).astype(np.uint32)

        # the numeric table is the library's pinned buffer, handed over without a copy (released when the frame dies)
        numeric, row_img = self._engine.featurize(I, M, copy="detach")
        self.error_message = ''
        header = ["intensity_image", "mask_image", "ROI_label", "t_index"] + self._engine.columns
        df = pd.DataFrame(numeric, columns=header[2:], copy=False)
        # string columns by a take on the (few) names in pandas' own string type: what DataFrame(string_data) infers, without
        # converting 2 x nrows Python objects
        sdt = pd.Series(np.array(["x"], dtype=object)).infer_objects().dtype
        df.insert(0, "mask_image", pd.array([str(v) for v in label_names], dtype=sdt).take(row_img))
        df.insert(0, "intensity_image", pd.array([str(v) for v in intensity_names], dtype=sdt).take(row_img))
        if "ROI_label" in df.columns:
            df.ROI_label = df.ROI_label.astype(np.uint32)
        return _emit(self, df, output_type, output_path)


    # ---- file-based entry points (reference nyxus.py:303-384, 524-600; new_bindings_py.cpp:224-384, 647-752)
    def featurize_files(self, intensity_files: list, mask_files: list, single_roi: bool, output_type: Optional[str] = "pandas",
                        output_path: Optional[str] = ""):
        """Features of intensity / mask image file pairs (single_roi: every intensity image is one whole-slide ROI)."""
        import os
        from . import imageio
        if intensity_files is None:
            raise IOError("The list of intensity file paths is empty")
        if mask_files is None:
            raise IOError("The list of segment file paths is empty")
        if output_type not in self._valid_output_types:
            raise ValueError(f'Invalid output type {output_type}. Valid output types are {self._valid_output_types}.')
        if output_type != 'pandas':
            _need_pyarrow()
        if len(intensity_files) == 0:
            raise RuntimeError("Intensity file list is blank")
        if not single_roi and len(intensity_files) != len(mask_files):
            raise RuntimeError(f"Imbalanced intensity ({len(intensity_files)} items) and segmentation mask ({len(mask_files)} items) file lists")
        for f in list(intensity_files) + ([] if single_roi else list(mask_files)):
            if not os.path.exists(f):
                raise RuntimeError(f"File does not exist: {f}")
        eng = self._engine
        frames = []
        eng.set_param("covered_range_from_scan", 1)      # file-based: the slide's range is the scanned one (slideprops.cpp:83-200)
        eng.set_param("single_roi", 1 if single_roi else 0)
        try:
            for k, ifile in enumerate(intensity_files):
                I = imageio.read_image(ifile, "intensity")
                if single_roi:
                    M = np.ones(I.shape, np.uint32)       # whole slide = ROI 1
                    mname = ""
                else:
                    M = imageio.read_image(mask_files[k], "mask").astype(np.uint32, copy=False)
                    mname = os.path.basename(mask_files[k])
                    if M.shape != I.shape:
                        raise ValueError(f"intensity / mask image size mismatch: {ifile} {I.shape} vs {mask_files[k]} {M.shape}")
                if self._mergerois:
                    M = (M != 0).astype(np.uint32)
                numeric, row_img = eng.featurize(I, M, copy="detach")
                header = ["intensity_image", "mask_image", "ROI_label", "t_index"] + eng.columns
                df = pd.DataFrame(numeric, columns=header[2:], copy=False)
                df.insert(0, "mask_image", [mname] * len(df))
                df.insert(0, "intensity_image", [os.path.basename(ifile)] * len(df))
                frames.append(df)
        finally:
            eng.set_param("covered_range_from_scan", 0)
            eng.set_param("single_roi", 0)
        df = pd.concat(frames, axis=0, ignore_index=True)
        if "ROI_label" in df.columns:
            df.ROI_label = df.ROI_label.astype(np.uint32)
        self.error_message = ''
        return _emit(self, df, output_type, output_path)

    def featurize_directory(self, intensity_dir: str, label_dir: Optional[str] = None, file_pattern: Optional[str] = ".*",
                            output_type: Optional[str] = "pandas", output_path: Optional[str] = ""):
        """Features of the images of `intensity_dir` whose names match `file_pattern`, masks of the same names in `label_dir`
        (None or the same directory: every image is one whole-slide ROI)."""
        import os
        import re
        if not os.path.exists(intensity_dir):
            raise IOError(f"Provided intensity image directory '{intensity_dir}' does not exist.")
        if label_dir is not None and not os.path.exists(label_dir):
            raise IOError(f"Provided label image directory '{label_dir}' does not exist.")
        if label_dir is None:
            label_dir = intensity_dir
        single = os.path.abspath(label_dir) == os.path.abspath(intensity_dir)
        try:
            rx = re.compile(file_pattern)
        except re.error:
            raise ValueError("Invalid filepattern " + file_pattern)
        names = sorted(f for f in os.listdir(intensity_dir) if rx.fullmatch(f) and os.path.isfile(os.path.join(intensity_dir, f)))
        if not names:
            raise RuntimeError(f"Error reading the dataset: no files in {intensity_dir} match {file_pattern}")
        ifiles = [os.path.join(intensity_dir, f) for f in names]
        mfiles = [] if single else [os.path.join(label_dir, f) for f in names]
        return self.featurize_files(ifiles, mfiles, single, output_type, output_path)


_VALID_KEYS_3D = {
    'neighbor_distance', 'pixels_per_micron', 'coarse_gray_depth',
    'n_feature_calc_threads', 'use_gpu_device', 'ibsi', 'channel_signature',
    'parent_channel', 'child_channel', 'aggregate',
    'dynamic_range', 'min_intensity', 'max_intensity', 'ram_limit',
    'verbose', 'anisotropy_x', 'anisotropy_y', 'anisotropy_z', 'preserve_hu'
}


class Nyxus3D:
    """3D feature extraction (reference nyxus.py:923-1489) for the 3D families on this path: first-order
    (*3D_ALL_INTENSITY*) and *3D_ALL_TEXTURE* (= *3D_GLCM*, *3D_GLSZM*, *3D_NGTDM*, *3D_GLDM*, *3D_GLRLM*, *3D_GLDZM*, *3D_NGLDM*)
    on a B200.

    Like the reference, the class has no in-memory entry point of its own: `featurize_files` /
    `featurize_directory` read NIfTI volumes.  `featurize(volume, labels)` is an addition for callers that already hold
    the voxels.  Metaparameters follow the reference: only 3glcm/greydepth has a default (coarse_gray_depth);
    3glszm/greydepth, 3gldm/greydepth, 3glrlm/greydepth, 3ngtdm/greydepth and 3ngtdm/radius are 0 until set with `set_metaparam` (env_features.cpp:712-736)."""

    def __init__(self, features: List[str], **kwargs):
        invalid_keys = set(kwargs.keys()) - _VALID_KEYS_3D
        if invalid_keys:
            print(f"Warning: unexpected keyword argument(s): {', '.join(invalid_keys)}")
        neighbor_distance = kwargs.get('neighbor_distance', 5)
        pixels_per_micron = kwargs.get('pixels_per_micron', 1.0)
        coarse_gray_depth = kwargs.get('coarse_gray_depth', 64)
        n_feature_calc_threads = kwargs.get('n_feature_calc_threads', 4)
        use_gpu_device = kwargs.get('use_gpu_device', -1)
        ibsi = kwargs.get('ibsi', False)
        verb_lvl = kwargs.get('verbose', 0)
        aniso = [kwargs.get('anisotropy_x', 1.0), kwargs.get('anisotropy_y', 1.0), kwargs.get('anisotropy_z', 1.0)]
        if neighbor_distance <= 0:
            raise ValueError("Neighbor distance must be greater than zero.")
        if pixels_per_micron <= 0:
            raise ValueError("Pixels per micron must be greater than zero.")
        if coarse_gray_depth <= 0:
            raise ValueError("Custom number of grayscale levels (parameter coarse_gray_depth, default=64) must be non-negative.")
        if n_feature_calc_threads < 1:
            raise ValueError("There must be at least one feature calculation thread.")
        if verb_lvl < 0:
            raise ValueError("verbosity must be non-negative")
        for name, a in zip("xyz", aniso):
            if a <= 0:
                raise ValueError(f"anisotropy_{name} must be positive")
        if any(a != 1.0 for a in aniso):
            raise ValueError("anisotropy is outside the texture/intensity path implemented by nyxus_b200")
        if kwargs.get('preserve_hu', False):
            raise ValueError("preserve_hu is outside the texture/intensity path implemented by nyxus_b200")
        self._engine = Engine(list(features), device=max(int(use_gpu_device), 0), coarse_gray_depth=coarse_gray_depth,
                              ibsi=ibsi, verbose=verb_lvl, dim=3)
        self._valid_output_types = ['pandas', 'arrowipc', 'parquet']

    def set_metaparam(self, paramval: str):
        key, _, val = paramval.partition('=')
        try:
            self._engine.set_param(key.strip(), float(val))
        except Exception as e:
            raise ValueError(f'Invalid metaparameter value {paramval}: {e}')

    @property
    def feature_columns(self):
        return self._engine.columns

    # reference nyxus.py:870-920
    def get_arrow_ipc_file(self):
        return getattr(self, "_arrow_path", "")

    def get_parquet_file(self):
        return getattr(self, "_arrow_path", "")

    def arrow_is_enabled(self):
        try:
            import pyarrow  # noqa: F401
            return True
        except ImportError:
            return False

    def _frame(self, numeric, row_vol, int_names, seg_names):
        header = ["intensity_image", "mask_image", "ROI_label", "t_index"] + self._engine.columns
        string_data = np.array([[int_names[k], seg_names[k]] for k in row_vol], dtype=object).reshape(-1, 2)
        df = pd.concat([pd.DataFrame(string_data, columns=header[:2]), pd.DataFrame(numeric, columns=header[2:])], axis=1)
        if "ROI_label" in df.columns:
            df.ROI_label = df.ROI_label.astype(np.uint32)
        return df

    def featurize(self, intensity_volume: np.ndarray, label_volume: np.ndarray, intensity_name: str = "Intensity0",
                  label_name: str = "Segmentation0"):
        """In-memory volume pair [D,H,W] (extension; the reference only reads files)."""
        if not isinstance(intensity_volume, np.ndarray) or not isinstance(label_volume, np.ndarray):
            raise ValueError("intensity_volume and label_volume must be numpy.ndarray")
        if intensity_volume.ndim != 3 or intensity_volume.shape != label_volume.shape:
            raise ValueError("Intensity and label volumes must be 3D arrays of the same shape")
        I = intensity_volume
        if I.dtype not in (np.uint16, np.uint32):
            if np.min(I) < 0:
                raise ValueError("negative voxels (Hounsfield-type scale) are outside the path implemented by nyxus_b200")
            I = I.astype(np.uint32)
        numeric, row_vol = self._engine.featurize_volume(I, label_volume.astype(np.uint32))
        return self._frame(numeric, row_vol, [intensity_name], [label_name])

    def featurize_files(self, intensity_files: list, mask_files: list, single_roi: bool, output_type: Optional[str] = "pandas",
                        output_path: Optional[str] = ""):
        import os
        from . import nifti
        if intensity_files is None:
            raise IOError("The list of intensity file paths is empty")
        if mask_files is None:
            raise IOError("The list of segment file paths is empty")
        if output_type not in self._valid_output_types:
            raise ValueError(f'Invalid output type {output_type}. Valid output types are {self._valid_output_types}')
        if output_type != 'pandas':
            _need_pyarrow()
        if not single_roi and len(intensity_files) != len(mask_files):
            raise ValueError("Number of intensity files must match the number of mask files")
        frames = []
        for k, ifile in enumerate(intensity_files):
            if not os.path.exists(ifile):
                raise IOError(f"Error: file {ifile} does not exist")
            I = nifti.to_grey(nifti.read(ifile), "intensity")
            if single_roi:
                M = np.ones(I.shape, np.uint32)       # whole volume = ROI 1 (new_bindings_py.cpp:788-795)
                mname = ""
            else:
                if not os.path.exists(mask_files[k]):
                    raise IOError(f"Error: file {mask_files[k]} does not exist")
                M = nifti.to_grey(nifti.read(mask_files[k]), "mask").astype(np.uint32)
                mname = os.path.basename(mask_files[k])
                if M.shape != I.shape:
                    raise ValueError(f"intensity / mask volume size mismatch: {I.shape} vs {M.shape}")
            numeric, row_vol = self._engine.featurize_volume(I, M)
            frames.append(self._frame(numeric, row_vol, [os.path.basename(ifile)], [mname]))
        if not frames:
            return self._frame(np.zeros((0, 2 + len(self._engine.columns))), [], [], [])
        return _emit(self, pd.concat(frames, axis=0, ignore_index=True), output_type, output_path)

    def featurize_directory(self, intensity_dir: str, label_dir: str, file_pattern: str, output_type: Optional[str] = "pandas",
                            output_path: Optional[str] = ""):
        import os
        import re
        if not os.path.exists(intensity_dir):
            raise IOError(f"Provided intensity image directory '{intensity_dir}' does not exist.")
        if label_dir is not None and not os.path.exists(label_dir):
            raise IOError(f"Provided label image directory '{label_dir}' does not exist.")
        single = label_dir is None or label_dir == intensity_dir
        rx = re.compile(file_pattern)
        names = sorted(f for f in os.listdir(intensity_dir) if rx.fullmatch(f) or rx.match(f))
        ifiles = [os.path.join(intensity_dir, f) for f in names]
        mfiles = [] if single else [os.path.join(label_dir, f) for f in names]
        return self.featurize_files(ifiles, mfiles, single, output_type, output_path)
