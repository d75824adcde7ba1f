"""GPU tests (-m gpu) of the drop-in boundary as a COMPILED module: the reference's own, unmodified Python layer
(src/nyx/python/nyxus/nyxus.py, staged under oracle/_ref/pyb200 by oracle/Makefile.backend) drives
nyxus_b200/backend*.so, and its DataFrames are compared with the same Python layer over the reference's own compiled
backend (oracle/_ref/pyref, SURVEY 8c option B).  Also reproduces tests/python/test_nyxus.py:400-446 of the reference."""
import os

import numpy as np
import pytest

import nyx_testutil as tu
from oracle import pyref
from reference_fixtures import PY_INTENS, PY_SEG

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not pyref.staged_available(), reason="reference Python layer not staged (oracle/Makefile.backend)")]

TEX = ["*ALL_GLCM*", "*ALL_GLRLM*", "*ALL_GLDZM*", "*ALL_GLSZM*", "*ALL_GLDM*", "*ALL_NGLDM*", "*ALL_NGTDM*"]


def _frames_match(ref_df, got_df):
    assert list(ref_df.columns) == list(got_df.columns)
    assert ref_df.shape == got_df.shape
    for c in ("intensity_image", "mask_image"):
        assert list(ref_df[c]) == list(got_df[c]), c
    assert np.array_equal(ref_df.ROI_label.to_numpy(), got_df.ROI_label.to_numpy())
    cols = list(ref_df.columns[2:])
    bad = tu.mismatches(ref_df[cols].to_numpy(np.float64), got_df[cols].to_numpy(np.float64), cols)
    assert not bad, bad[:8]


def test_reference_pytest_in_memory_2d():
    """tests/python/test_nyxus.py:400-409"""
    nyxus = pyref.reference_on_b200()
    nyx = nyxus.Nyxus(["*ALL_GLCM*"], ibsi=True)
    names = ["test_name_1"]
    df = nyx.featurize(PY_INTENS[0], PY_SEG[0], names, names)
    assert nyx.error_message == ''
    assert df.shape[0] == 1 and list(df.intensity_image) == names and list(df.mask_image) == names


def test_reference_pytest_in_memory_3d():
    """tests/python/test_nyxus.py:411-446 (IBSI consensus values of the phantom, 1 % / 10 %)"""
    nyxus = pyref.reference_on_b200()
    nyx = nyxus.Nyxus([
        "GLCM_ASM", "GLCM_CONTRAST", "GLCM_CORRELATION", "GLCM_DIFAVE",
        "GLCM_DIFENTRO", "GLCM_DIFVAR", "GLCM_ENERGY", "GLCM_ENTROPY",
        "GLCM_HOM1", "GLCM_INFOMEAS1", "GLCM_INFOMEAS2", "GLCM_IDM",
        "GLCM_SUMAVERAGE", "GLCM_SUMENTROPY", "GLCM_SUMVARIANCE", "GLCM_VARIANCE"], ibsi=True)
    names = ["test_name_1", "test_name_2", "test_name_3", "test_name_4"]
    nyx.set_metaparam("glcm/greydepth=100")
    nyx.set_metaparam("glcm/offset=1")
    df = nyx.featurize(PY_INTENS, PY_SEG, names, names)
    assert nyx.error_message == ''
    mean_values = df.mean(numeric_only=True).tolist()
    mean_values.pop(0)
    mean_values.pop(0)
    averaged = [sum(mean_values[i:i + 4]) / 4 for i in range(0, len(mean_values), 4)]
    assert pytest.approx(averaged[0], 0.01) == 0.368
    assert pytest.approx(averaged[1], 0.01) == 5.28
    assert pytest.approx(averaged[2], 0.01) == -0.0121
    assert pytest.approx(averaged[4], 0.01) == 1.40
    assert pytest.approx(averaged[5], 0.1) == 2.90


@pytest.mark.skipif(not pyref.reference_available(), reason="oracle/_ref/pyref (the reference's compiled backend) not built")
@pytest.mark.parametrize("ibsi", [False, True])
def test_reference_python_layer_same_table_over_both_backends(ibsi):
    from nyxus_b200 import synth
    ref_nyxus, our_nyxus = pyref.reference(), pyref.reference_on_b200()
    inten, label = synth.blob_slide(192, 256, 32, 7, 12, seed=21, jitter=2)
    inten2, label2 = synth.blob_slide(192, 256, 32, 8, 13, seed=22, jitter=1)
    I = np.stack([inten, inten2]).astype(np.uint16)
    L = np.stack([label, label2])
    if ibsi:
        I = (I >> 6).astype(np.uint16)      # IBSI mode indexes matrices by raw intensity: keep the levels small
    feats = TEX + ["*ALL_INTENSITY*"]
    kw = dict(ibsi=ibsi, coarse_gray_depth=48, n_feature_calc_threads=4)
    ref_df = ref_nyxus.Nyxus(feats, **kw).featurize(I.copy(), L.copy(), ["a.tif", "b.tif"], ["ma.tif", "mb.tif"])
    our = our_nyxus.Nyxus(feats, **kw)
    got_df = our.featurize(I.copy(), L.copy(), ["a.tif", "b.tif"], ["ma.tif", "mb.tif"])
    assert our.error_message == ''
    _frames_match(ref_df, got_df)
    # metaparameters through the reference's set_metaparam: glcm/offset changes the GLCM block, glcm/greydepth only the blank test
    r2, o2 = ref_nyxus.Nyxus(["*ALL_GLCM*"], **kw), our_nyxus.Nyxus(["*ALL_GLCM*"], **kw)
    for n in (r2, o2):
        n.set_metaparam("glcm/offset=2")
        n.set_metaparam("glcm/greydepth=3")
    _frames_match(r2.featurize(I[0].copy(), L[0].copy()), o2.featurize(I[0].copy(), L[0].copy()))
    # the opt-in multi-valued HISTOGRAM feature: one column per bin of the grey depth, through both backends
    _frames_match(ref_nyxus.Nyxus(["HISTOGRAM", "MEDIAN"], **kw).featurize(I[0].copy(), L[0].copy()),
                  our_nyxus.Nyxus(["HISTOGRAM", "MEDIAN"], **kw).featurize(I[0].copy(), L[0].copy()))
    # negative intensities are shifted by the Python layer (nyxus.py:482-484) before they reach the backend
    J = I[0].astype(np.int32) - 500
    _frames_match(ref_nyxus.Nyxus(["*ALL_GLRLM*"], **kw).featurize(J.copy(), L[0].copy()),
                  our_nyxus.Nyxus(["*ALL_GLRLM*"], **kw).featurize(J.copy(), L[0].copy()))


@pytest.mark.skipif(not pyref.reference_available(), reason="oracle/_ref/pyref (the reference's compiled backend) not built")
def test_reference_nyxus3d_featurize_files_over_both_backends(tmp_path):
    """Nyxus3D.featurize_files on NIfTI volumes (new_bindings_py.cpp:753-856): file names, row order and values"""
    from nyxus_b200 import nifti
    ref_nyxus, our_nyxus = pyref.reference(), pyref.reference_on_b200()
    rng = np.random.default_rng(5)
    files = []
    for k in range(2):
        D, H, W = 20, 24, 28
        vol = rng.integers(1, 900, (D, H, W)).astype(np.uint16)
        lab = np.zeros((D, H, W), np.uint16)
        lab[2:11, 3:14, 2:12] = 1
        lab[9:19, 10:22, 14:26] = 2 + k
        lab[4:8, 16:22, 3:9] = 7
        lab[rng.random((D, H, W)) < 0.08] = 0
        fi, fm = str(tmp_path / ("v%d.nii" % k)), str(tmp_path / ("m%d.nii" % k))
        nifti.write(fi, vol)
        nifti.write(fm, lab)
        files.append((fi, fm))
    feats = ["*3D_GLCM*", "*3D_GLSZM*", "*3D_NGTDM*", "*3D_ALL_INTENSITY*"]

    def run(mod):
        n = mod.Nyxus3D(feats, coarse_gray_depth=32)
        for mp in ("3glcm/greydepth=24", "3glszm/greydepth=16", "3ngtdm/greydepth=16", "3ngtdm/radius=1"):
            n.set_metaparam(mp)
        return n.featurize_files([f[0] for f in files], [f[1] for f in files], False)
    _frames_match(run(ref_nyxus), run(our_nyxus))


def _write_tiffs(tmp_path, n=3, same_shape=True):
    import cv2
    from nyxus_b200 import synth
    idir, sdir = tmp_path / "int", tmp_path / "seg"
    idir.mkdir()
    sdir.mkdir()
    slides = []
    for k in range(n):
        H, W = (160, 192) if same_shape or k % 2 == 0 else (128, 224)
        I, L = synth.blob_slide(H, W, 32, 7, 12, seed=60 + k, jitter=1)
        name = "p0_y1_r%d_c0.ome.tif" % k
        assert cv2.imwrite(str(idir / name), I.astype(np.uint16))
        assert cv2.imwrite(str(sdir / name), L.astype(np.uint16))
        slides.append((name, I.astype(np.uint32), L.astype(np.uint32)))
    (idir / "notes.txt").write_text("not an image")
    return str(idir), str(sdir), slides


@pytest.mark.skipif(not pyref.reference_available(), reason="oracle/_ref/pyref (the reference's compiled backend) not built")
@pytest.mark.parametrize("same_shape", [True, False])
def test_file_based_entry_points_2d(tmp_path, same_shape):
    """featurize_directory / featurize_files on TIFF files (new_bindings_py.cpp:224-384, 647-752): the reference's nyxus.py over
    the product's backend module AND the product's own Python mirror against the reference's in-memory featurize() of the
    decoded arrays (the oracle build has no libtiff).  COVERED_IMAGE_INTENSITY_RANGE is the one column that differs by
    design between the in-memory and the file-based workflow: file-based, the denominator is the slide's scanned range
    over its labelled pixels (slideprops.cpp:83-200)."""
    pytest.importorskip("cv2")
    idir, sdir, slides = _write_tiffs(tmp_path, 3, same_shape)
    feats = ["*ALL_GLCM*", "*ALL_GLSZM*", "*ALL_NGTDM*", "*ALL_INTENSITY*", "*BASIC_MORPHOLOGY*"]
    ref_nyxus, our_nyxus = pyref.reference(), pyref.reference_on_b200()
    refn = ref_nyxus.Nyxus(feats)
    want = []
    for name, I, L in slides:
        df = refn.featurize(I.copy(), L.copy(), [name], [name])
        fg = I[L > 0].astype(np.float64)
        df["COVERED_IMAGE_INTENSITY_RANGE"] = (df["MAX"] - df["MIN"]) / (fg.max() - fg.min())
        want.append(df)
    import pandas as pd
    want = pd.concat(want, axis=0, ignore_index=True)
    on_b200 = our_nyxus.Nyxus(feats)
    got_dir = on_b200.featurize_directory(idir, sdir, r".*\.ome\.tif")
    _frames_match(want, got_dir)
    got_files = on_b200.featurize_files([os.path.join(idir, s[0]) for s in slides], [os.path.join(sdir, s[0]) for s in slides], False)
    _frames_match(want, got_files)
    # the in-memory path still reports the sentinel ratio (-0.0 -> 0) afterwards
    mem = on_b200.featurize(slides[0][1].copy(), slides[0][2].copy())
    assert (mem["COVERED_IMAGE_INTENSITY_RANGE"] == 0).all()
    from nyxus_b200 import Nyxus
    mirror = Nyxus(feats)
    _frames_match(want, mirror.featurize_directory(idir, sdir, r".*\.ome\.tif"))
    # whole-slide mode: one ROI per image, label 1
    ws = on_b200.featurize_files([os.path.join(idir, slides[0][0])], [], True)
    assert ws.shape[0] == 1 and int(ws.ROI_label[0]) == 1 and ws.AREA_PIXELS_COUNT[0] == slides[0][1].size
    with pytest.raises(RuntimeError, match="File does not exist"):
        on_b200.featurize_files([os.path.join(idir, "nope.tif")], [os.path.join(sdir, "nope.tif")], False)


DROPIN_SITE = os.path.join(tu.ROOT, "oracle", "_ref", "pyb200_site")
REF_PYTESTS = os.path.join(tu.ROOT, "oracle", "_ref", "pytests")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_PYTESTS, "test_nyxus.py")), reason="reference test-suite not staged (make -f oracle/Makefile.backend dropin)")
def test_reference_own_pytest_suite_runs_unmodified_on_the_dropin_deployment():
    """INTEGRATION.md section 1, literally: the reference's `nyxus` package with the product's backend*.so + libnyxus_b200.so
    in place of the reference's binary (oracle/_ref/pyb200_site), and the reference's OWN, UNMODIFIED tests/python/test_nyxus.py
    run against it with pytest -- the tests whose features lie on the texture + intensity path:
      test_in_memory_2d / _3d             IBSI phantom, IBSI consensus GLCM values
      test_3d_glcm / ngtdm / glszm / gldm / glrlm_compatibility   NIfTI MRI + liver mask against PyRadiomics ground truth
                                          (PyRadiomics binning through set_metaparam)
      TestSingleRoi (5 tests)             featurize_files / featurize_directory on the suite's OME-TIFFs, whole-slide mode
    (the rest of that file needs *ALL* = every family of the reference, Arrow writers or image quality)."""
    import shutil
    import subprocess
    import sys
    import sysconfig
    pkg = os.path.join(DROPIN_SITE, "nyxus")
    for f in ("backend" + sysconfig.get_config_var("EXT_SUFFIX"), "libnyxus_b200.so"):      # the binaries under test = the built ones
        shutil.copy2(os.path.join(tu.ROOT, "nyxus_b200", f), os.path.join(pkg, f))
    env = dict(os.environ, PYTHONPATH=DROPIN_SITE)
    sel = ("in_memory_2d or in_memory_3d or 3d_glcm_compatibility or 3d_ngtdm_compatibility or 3d_glszm_compatibility "
           "or 3d_gldm_compatibility or 3d_glrlm_compatibility or TestSingleRoi")
    r = subprocess.run([sys.executable, "-m", "pytest", "test_nyxus.py", "-q", "-p", "no:cacheprovider", "-k", sel],
                       cwd=REF_PYTESTS, env=env, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail
    import re
    assert int(re.search(r"(\d+) passed", r.stdout).group(1)) >= 12, tail


@pytest.mark.parametrize("otype,ext", [("arrowipc", ".arrow"), ("parquet", ".parquet")])
def test_arrow_and_parquet_output_through_the_reference_python_layer(tmp_path, otype, ext):
    """output_type='arrowipc' / 'parquet' (nyxus.py:486-518, arrow_helpers.cpp:9-62, output_writers.cpp:21-30): the file the
    reference's Python layer returns holds the table of the pandas call, typed utf8 / utf8 / int32 / float64..."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    from nyxus_b200 import synth
    nyxus = pyref.reference_on_b200()
    inten, label = synth.blob_slide(128, 160, 32, 7, 12, seed=5, jitter=1)
    nyx = nyxus.Nyxus(["*ALL_GLCM*", "*ALL_INTENSITY*"])
    assert nyx.arrow_is_enabled()
    df = nyx.featurize(inten.astype(np.uint16), label, ["i.tif"], ["m.tif"])
    path = nyx.featurize(inten.astype(np.uint16), label, ["i.tif"], ["m.tif"], output_type=otype, output_path=str(tmp_path / "out") + "/")
    assert path == str(tmp_path / "out" / ("NyxusFeatures" + ext)) and os.path.exists(path)
    assert (nyx.get_arrow_ipc_file() if otype == "arrowipc" else nyx.get_parquet_file()) == path
    t = pa.ipc.open_file(path).read_all() if otype == "arrowipc" else pq.read_table(path)
    assert t.column_names == list(df.columns)
    assert [str(f.type) for f in t.schema][:4] == ["string", "string", "int32", "double"]
    got = t.to_pandas()
    assert list(got.intensity_image) == list(df.intensity_image) and np.array_equal(got.ROI_label.to_numpy(), df.ROI_label.to_numpy())
    assert np.array_equal(got.iloc[:, 3:].to_numpy(np.float64), df.iloc[:, 3:].to_numpy(np.float64))
    # the host-side mirror writes the same file
    from nyxus_b200 import Nyxus
    mine = Nyxus(["*ALL_GLCM*", "*ALL_INTENSITY*"])
    p2 = mine.featurize(inten.astype(np.uint16), label, ["i.tif"], ["m.tif"], output_type=otype, output_path=str(tmp_path / ("x" + ext)))
    t2 = pa.ipc.open_file(p2).read_all() if otype == "arrowipc" else pq.read_table(p2)
    assert t2.schema.equals(t.schema) and t2.equals(t)
