"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against
  (1) golden vectors produced by the compiled reference (tests/golden/*.npz),
  (2) the compiled reference itself (oracle/_ref, when the prebuilt .so travelled) on fresh seeded inputs,
  (3) the numpy oracle,
with every matrix entry exact and every derived feature within 1e-9 relative (north_star), plus size-independent
properties at the full configs[1] size."""
import os

import numpy as np
import pytest

import nyx_testutil as tu
from oracle import nyx_oracle as no
from oracle import ref

pytestmark = pytest.mark.gpu

ALL = ["*ALL_INTENSITY*", "*ALL_TEXTURE*"]
COLS, CASES = tu.load_golden_rois()


def _engine(features=ALL, **kw):
    from nyxus_b200 import Engine
    return Engine(features, **kw)


@pytest.fixture(scope="module")
def eng64():
    e = _engine()
    yield e
    e.close()


# ------------------------------------------------------------------------------------------------ golden, per ROI
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_roi_features_and_matrices(case):
    e = _engine(coarse_gray_depth=case["g"], ibsi=case["ibsi"])
    try:
        I, L = tu.place_roi(case["tile"], case["mask"], label=5)
        tab, img = e.featurize(I, L)
        assert e.columns == COLS
        assert tab.shape == (1, 339) and tab[0, 0] == 5 and tab[0, 1] == 0
        bad = tu.mismatches(case["features"], tab[0, 2:], COLS)
        assert not bad, bad[:8]
        for fam in tu.FAMILIES:
            got, svals = e.debug_matrix(5, fam)
            assert sorted(got) == sorted(case["mats"][fam]), (fam, case["name"])
            if fam == "NGTDM" and len(case["ngtdm_s"]):
                assert np.allclose(svals, case["ngtdm_s"], rtol=1e-12, atol=1e-12)
    finally:
        e.close()


def test_golden_slide_table(eng64):
    s = tu.load_golden_slide()
    tab, img = eng64.featurize(s["inten"], s["label"])
    assert np.array_equal(tab[:, 0].astype(np.int64), s["labels"].astype(np.int64))       # rows by ascending label
    assert not tu.mismatches(s["table"], tab[:, 2:], COLS)
    # uint16 entry point gives the identical table
    tab16, _ = eng64.featurize(s["inten"].astype(np.uint16), s["label"])
    assert np.array_equal(tab, tab16)


def test_golden_slide_table_ibsi():
    s = tu.load_golden_slide()
    e = _engine(ibsi=True)
    try:
        tab, _ = e.featurize(np.minimum(s["inten"], 255), s["label"])
        assert np.array_equal(tab[:, 0].astype(np.int64), s["labels_ibsi"].astype(np.int64))
        assert not tu.mismatches(s["table_ibsi"], tab[:, 2:], COLS)
    finally:
        e.close()


# ------------------------------------------------------------------------------------------------ phase 1 / CSR pack: exact
def test_phase1_records_and_cloud_exact(eng64):
    from nyxus_b200 import synth
    I, L = synth.random_rois(300, 257, 60, seed=21)       # odd width: scalar-load path of the scan
    tab, _ = eng64.featurize(I, L)
    recs = eng64.debug_rois()
    labels = np.unique(L[L > 0])
    assert np.array_equal(recs[:, 0], labels)
    for r in recs:
        m = L == r[0]
        ys, xs = np.nonzero(m)
        assert list(r[1:]) == [m.sum(), I[m].min(), I[m].max(), xs.min(), ys.min(), xs.max() - xs.min() + 1, ys.max() - ys.min() + 1]
        assert np.array_equal(eng64.debug_cloud(int(r[0])), I[m])      # raster order
    assert eng64.counts()["fg_pixels"] == int((L > 0).sum())


def _phase1_expected(I, L):
    """per-label (label, area, min, max, xmin, ymin, w, h) by numpy (reference: init/update_label_record_3)"""
    lab = L.ravel().astype(np.int64)
    fg = lab > 0
    ys, xs = np.divmod(np.nonzero(fg)[0], L.shape[1])
    lab = lab[fg]
    val = I.ravel()[fg].astype(np.int64)
    order = np.argsort(lab, kind="stable")
    lab, ys, xs, val = lab[order], ys[order], xs[order], val[order]
    u, first = np.unique(lab, return_index=True)
    red = lambda f, a: f.reduceat(a, first)
    area = np.diff(np.append(first, len(lab)))
    x0, x1, y0, y1 = red(np.minimum, xs), red(np.maximum, xs), red(np.minimum, ys), red(np.maximum, ys)
    return np.stack([u, area, red(np.minimum, val), red(np.maximum, val), x0, y0, x1 - x0 + 1, y1 - y0 + 1], 1)


@pytest.mark.parametrize("H,W,kind", [(37, 256, "noise"), (64, 131, "noise"), (50, 1000, "runs"), (33, 515, "runs"), (200, 384, "one"),
                                      (5, 3, "noise"), (300, 1028, "blocks"), (70, 2056, "runs"), (40, 520, "noise"), (9, 8, "one")])
def test_scan_run_segmentation_stress(H, W, kind):
    """label layouts that exercise every branch of the row-run scan: 1-pixel runs (3-4 runs inside one lane), runs crossing
    lanes and 256-pixel chunks, a whole-image label (per-lane cache), widths with and without the 128-bit load path"""
    rng = np.random.default_rng(H * 1000 + W)
    if kind == "noise":
        L = rng.integers(0, 6, (H, W)).astype(np.uint32)
    elif kind == "runs":
        L = np.zeros((H, W), np.uint32)
        for y in range(H):
            x = 0
            while x < W:
                n = int(rng.integers(1, 200))
                L[y, x:x + n] = rng.integers(0, 40)
                x += n
    elif kind == "one":
        L = np.full((H, W), 7, np.uint32)
    else:
        L = (1 + (np.arange(H)[:, None] // 50) * 40 + np.arange(W)[None, :] // 30).astype(np.uint32)
        L[rng.random((H, W)) < 0.05] = 0
    I = rng.integers(0, 60000, (H, W)).astype(np.uint16)
    I32 = rng.integers(0, 2**25, (H, W)).astype(np.uint32)          # beyond 16 bits: the two-register (min, max) path (2^26 = range limit of the large-ROI first-order histogram)
    e = _engine(features=["MEAN"])
    try:
        for inten in (I, I.astype(np.uint32), I32):
            e.featurize(inten, L)
            assert np.array_equal(e.debug_rois(), _phase1_expected(inten, L))
    finally:
        e.close()


# ------------------------------------------------------------------------------------------------ vs the compiled reference
@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libnyxus_ref.so did not travel")
@pytest.mark.parametrize("seed,g,ibsi,levels", [(1, 64, False, None), (2, 16, False, None), (3, 256, False, None), (4, 64, True, 120),
                                                (5, 64, False, 5), (6, 64, True, 9)])
def test_random_slide_against_compiled_reference(seed, g, ibsi, levels):
    from nyxus_b200 import synth
    I, L = synth.random_rois(240, 260, 45, seed=100 + seed, levels=levels)
    e = _engine(coarse_gray_depth=g, ibsi=ibsi)
    try:
        tab, _ = e.featurize(I, L)
        rl, rt, _ = ref.featurize_2d(I, L, greydepth=g, ibsi=ibsi, nthreads=4)
        assert np.array_equal(tab[:, 0].astype(np.int64), rl.astype(np.int64))
        bad = tu.mismatches(rt, tab[:, 2:], COLS)
        assert not bad, bad[:8]
        recs = e.debug_rois()
        for r in recs[::9]:
            x0, y0, w, h = int(r[4]), int(r[5]), int(r[6]), int(r[7])
            mask = L[y0:y0 + h, x0:x0 + w] == r[0]
            rm = ref.roi_matrices(I[y0:y0 + h, x0:x0 + w] * mask, mask, g, ibsi)
            for fam in tu.FAMILIES:
                got, _ = e.debug_matrix(int(r[0]), fam)
                assert sorted(got) == sorted(rm[fam]), (fam, int(r[0]))
    finally:
        e.close()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libnyxus_ref.so did not travel")
def test_config1_against_compiled_reference(eng64):
    """BASELINE.json configs[0]: 1024x1024 uint16, 256 blob ROIs"""
    from nyxus_b200 import synth
    I, L = synth.config1()
    tab, _ = eng64.featurize(I, L)
    rl, rt, _ = ref.featurize_2d(I.astype(np.uint32), L, nthreads=8)
    assert tab.shape[0] == 256 and np.array_equal(tab[:, 0].astype(np.int64), rl.astype(np.int64))
    assert not tu.mismatches(rt, tab[:, 2:], COLS)


# ------------------------------------------------------------------------------------------------ vs the numpy oracle
def test_small_rois_against_numpy_oracle(eng64):
    rng = np.random.default_rng(8)
    I = rng.integers(1, 3000, (40, 48)).astype(np.uint32)
    L = np.zeros((40, 48), np.uint32)
    L[2:12, 3:14] = 4
    L[15:22, 20:33] = 9
    L[30:39, 1:9] = 2
    L[25:27, 40:47] = 11
    tab, _ = eng64.featurize(I, L)
    for i, lab in enumerate(tab[:, 0].astype(int)):
        m = L == lab
        oc, om = no.roi_features(np.where(m, I, 0), m, 64, False)
        assert not tu.mismatches(np.array([oc[c] for c in COLS]), tab[i, 2:], COLS)
        for fam in tu.FAMILIES:
            got, _ = eng64.debug_matrix(lab, fam)
            assert sorted(got) == sorted(om[fam])


# ------------------------------------------------------------------------------------------------ API behaviour / edge cases
def test_empty_mask_and_montage_and_selection():
    from nyxus_b200 import Nyxus
    nyx = Nyxus(["GLCM_ASM", "MEAN", "*ALL_NGTDM*"])
    assert nyx.feature_columns == ["MEAN", "GLCM_ASM_0", "GLCM_ASM_45", "GLCM_ASM_90", "GLCM_ASM_135", "NGTDM_COARSENESS",
                                   "NGTDM_CONTRAST", "NGTDM_BUSYNESS", "NGTDM_COMPLEXITY", "NGTDM_STRENGTH"]
    I = np.random.default_rng(1).integers(1, 500, (2, 32, 32)).astype(np.uint16)
    L = np.zeros((2, 32, 32), np.int64)
    df = nyx.featurize(I, L)                       # no ROI at all
    assert df.shape == (0, 4 + 10)
    L[0, 4:20, 4:20] = 2
    L[1, 1:9, 1:9] = 7
    L[1, 12:30, 10:28] = 3
    df = nyx.featurize(I, L, intensity_names=["a", "b"], label_names=["ma", "mb"])
    assert list(df.columns[:4]) == ["intensity_image", "mask_image", "ROI_label", "t_index"]
    assert list(df.intensity_image) == ["a", "b", "b"] and list(df.ROI_label) == [2, 3, 7]
    assert abs(df.MEAN[0] - I[0, 4:20, 4:20].mean()) < 1e-9
    with pytest.raises(ValueError, match="not a valid 2D feature name"):
        Nyxus(["NOT_A_FEATURE"])
    with pytest.raises(ValueError):
        nyx.featurize(I, L[:, :16])


def test_pipelined_slide_list_equals_single_slide_calls(eng64):
    """nyxb200_featurize_2d_list / the contiguous stack with n_img > 1: slide k+1 is uploaded into the second device buffer
    while slide k computes.  Rows must be bit-identical to one call per slide, in slide order."""
    from nyxus_b200 import synth
    slides = [synth.blob_slide(300, 420, 30, 7, 12, seed=40 + k) for k in range(5)]
    singles = [eng64.featurize(I, L)[0] for I, L in slides]
    I = np.ascontiguousarray(np.stack([s[0] for s in slides]))
    L = np.ascontiguousarray(np.stack([s[1] for s in slides]).astype(np.uint32))
    tab, img = eng64.featurize(I, L)
    assert np.array_equal(tab, np.concatenate(singles))
    assert np.array_equal(img, np.concatenate([np.full(len(t), k, np.int32) for k, t in enumerate(singles)]))
    # separate host buffers, in reverse order
    keep = [(np.ascontiguousarray(s[0]), np.ascontiguousarray(s[1].astype(np.uint32))) for s in slides[::-1]]
    tab2, img2 = eng64.featurize_list([a.ctypes.data for a, _ in keep], keep[0][0].dtype == np.uint16,
                                      [b.ctypes.data for _, b in keep], 300, 420)
    assert np.array_equal(tab2, np.concatenate(singles[::-1]))
    assert np.array_equal(img2, np.concatenate([np.full(len(t), k, np.int32) for k, t in enumerate(singles[::-1])]))


def test_negative_intensities_are_shifted_like_the_reference():
    from nyxus_b200 import Nyxus
    nyx = Nyxus(["MIN", "MAX"])
    I = np.array([[-5, 1, 2], [3, 4, 5], [6, 7, 8]], np.int32)
    L = np.ones((3, 3), np.uint32)
    df = nyx.featurize(I, L)
    # 2D input is wrapped by np.array([..]) (a copy, nyxus.py:447) so the caller's array keeps its values ...
    assert df.MIN[0] == 0 and df.MAX[0] == 13 and I.min() == -5
    # ... while a 3D stack is shifted IN PLACE (nyxus.py:481-484)
    I3 = I[None].copy()
    df = nyx.featurize(I3, L[None])
    assert df.MIN[0] == 0 and df.MAX[0] == 13 and I3.min() == 0


def test_label_table_growth_and_limit(eng64):
    I = np.full((16, 16), 10, np.uint32)
    L = np.zeros((16, 16), np.uint32)
    L[2:6, 2:6] = 3_000_000           # beyond the initial 2^20 table -> one retry
    L[8:12, 8:12] = 17
    tab, _ = eng64.featurize(I, L)
    assert list(tab[:, 0]) == [17, 3_000_000]
    L[2:6, 2:6] = (1 << 27) + 5       # beyond the dense table: the mask is re-labelled on the device, the label comes back
    tab, _ = eng64.featurize(I, L)
    assert list(tab[:, 0]) == [17, (1 << 27) + 5]


def test_large_and_sparse_label_values_are_relabelled_on_the_device():
    """the reference keys ROIs by the int label itself (pixel_feed.cpp:28): any value is fine.  Labels beyond 2^24 go through
    the device re-labelling (ranks in ascending label order): same rows, same order, the original labels in column 0 --
    checked against the same slide with small labels and against the compiled reference"""
    from nyxus_b200 import synth
    inten, label = synth.blob_slide(160, 192, 32, 7, 12, seed=31, jitter=2)
    small = np.unique(label[label != 0])
    rng = np.random.default_rng(5)
    big = np.sort(rng.choice(np.arange(1 << 24, (1 << 31) - 1, dtype=np.int64), size=len(small), replace=False)).astype(np.uint32)
    big[0], big[-1] = 50_000_000, 2_147_000_001
    big.sort()
    lut = np.zeros(int(small.max()) + 1, np.uint32)
    lut[small] = big                                        # ascending -> ascending: the row order does not change
    Lbig = lut[label]
    e = _engine()
    try:
        t_small, _ = e.featurize(inten, label)
        t_big, _ = e.featurize(inten, Lbig)
        assert np.array_equal(t_big[:, 0].astype(np.int64), big.astype(np.int64))
        assert np.array_equal(t_big[:, 1:], t_small[:, 1:])
        t_again, _ = e.featurize(inten, label)              # and a small-label slide afterwards is untouched
        assert np.array_equal(t_again, t_small)
    finally:
        e.close()
    if ref.available():
        rl, rt, _ = ref.featurize_2d(inten.astype(np.uint32), Lbig)
        assert np.array_equal(rl.astype(np.int64), big.astype(np.int64))
        bad = tu.mismatches(rt, t_big[:, 2:], COLS)
        assert not bad, bad[:6]


@pytest.mark.parametrize("ibsi", [False, True])
def test_tile_store_and_direct_binning_give_identical_tables(ibsi):
    """K2's binned-tile store (16-bit codes read back by the family kernels) vs binning straight from the slide."""
    from nyxus_b200 import synth
    I, L = synth.random_rois(300, 257, 60, seed=33)
    if ibsi:
        I = np.minimum(I, 200)
    e = _engine(ibsi=ibsi)
    try:
        a, _ = e.featurize(I, L)
        mats = {f: e.debug_matrix(int(a[3, 0]), f)[0] for f in tu.FAMILIES}
        e.set_param("tile_store", 0)
        b, _ = e.featurize(I, L)
        assert np.array_equal(a, b)
        for f in tu.FAMILIES:
            assert e.debug_matrix(int(a[3, 0]), f)[0] == mats[f], f
    finally:
        e.close()


@pytest.mark.parametrize("nb", [2, 5, 8, 16])
def test_roi_batches_give_identical_tables(nb):
    """Rows are produced batch by batch (D2H of batch b overlaps batch b+1); the table must not depend on the split."""
    from nyxus_b200 import synth
    I, L = synth.random_rois(300, 257, 60, seed=34)
    I2, L2 = synth.random_rois(300, 257, 7, seed=35)
    e = _engine()
    try:
        e.set_param("roi_batches", 1)
        a, ia = e.featurize(np.stack([I, I2]), np.stack([L, L2]))
        e.set_param("roi_batches", nb)
        b, ib = e.featurize(np.stack([I, I2]), np.stack([L, L2]))
        assert np.array_equal(a, b) and np.array_equal(ia, ib)
        assert np.all(np.diff(a[ia == 0, 0]) > 0) and np.all(np.diff(a[ia == 1, 0]) > 0)
    finally:
        e.close()


def test_runs_are_bitwise_deterministic(eng64):
    from nyxus_b200 import synth
    I, L = synth.blob_slide(512, 512, 31, 7, 12, seed=5, jitter=0)
    a, _ = eng64.featurize(I, L)
    b, _ = eng64.featurize(I, L)
    assert np.array_equal(a, b)


# ------------------------------------------------------------------------------------------------ full-size properties (configs[1])
def test_config2_full_size_properties(eng64):
    """10k x 10k, 103,684 ROIs.  Size-independent checks: (a) sum of areas == foreground pixels, (b) rows sorted by label,
    (c) translation invariance: a ROI's features do not depend on where the ROI sits, (d) a crop processed on its own gives
    bit-identical rows, (e) every value finite."""
    from nyxus_b200 import synth
    I, L = synth.config2(seed=2, size=10000)
    tab, _ = eng64.featurize(I, L)
    assert tab.shape == (103684, 339)
    assert eng64.counts()["fg_pixels"] == int(np.count_nonzero(L))
    assert np.all(np.diff(tab[:, 0]) > 0)
    assert np.isfinite(tab).all()
    # integrated intensity column sums to the sum over all foreground pixels (exact in FP64 at this size)
    j = 2 + COLS.index("INTEGRATED_INTENSITY")
    assert tab[:, j].sum() == float(I[L > 0].astype(np.int64).sum())
    # crop: the first 5 grid rows/cols (pitch 31) re-labelled identically
    c = 31 * 5
    Ic, Lc = np.ascontiguousarray(I[:c, :c]), np.ascontiguousarray(L[:c, :c])
    sub, _ = eng64.featurize(Ic, Lc)
    keep = np.isin(tab[:, 0], sub[:, 0])
    assert np.array_equal(tab[keep], sub)
    # translation: shift the crop by (7, 13) inside a larger blank slide
    I2 = np.zeros((c + 20, c + 30), I.dtype)
    L2 = np.zeros((c + 20, c + 30), L.dtype)
    I2[7:7 + c, 13:13 + c] = Ic
    L2[7:7 + c, 13:13 + c] = Lc
    sh, _ = eng64.featurize(I2, L2)
    assert np.array_equal(sh, sub)


# ------------------------------------------------------------------------------------------------ 3D (Nyxus3D path)
GV = tu.load_golden_vol()


def _engine3(case_params, features=("*3D_ALL_INTENSITY*", "*3D_GLCM*", "*3D_NGTDM*", "*3D_GLSZM*")):
    from nyxus_b200 import Engine
    g, gz, gn, rad, ibsi = [int(v) for v in case_params]
    e = Engine(list(features), coarse_gray_depth=g, ibsi=bool(ibsi), dim=3)
    e.set_param("3glszm/greydepth", gz)
    e.set_param("3ngtdm/greydepth", gn)
    e.set_param("3ngtdm/radius", rad)
    return e


@pytest.mark.parametrize("case", ["matlab64_r1", "defaults", "g16_r1", "ibsi_r1"])
def test_golden_volume_table(case):
    e = _engine3(GV["params_" + case])
    try:
        assert e.columns == GV["columns"]
        vol = GV["inten_ibsi"] if case.startswith("ibsi") else GV["inten"]
        tab, _ = e.featurize_volume(vol, GV["label"])
        assert np.array_equal(tab[:, 0].astype(np.int64), GV["labels_" + case].astype(np.int64))
        bad = tu.mismatches(GV["table_" + case], tab[:, 2:], GV["columns"])
        assert not bad, bad[:8]
        tab32, _ = e.featurize_volume(vol.astype(np.uint32), GV["label"])       # uint32 entry point: identical table
        assert np.array_equal(tab, tab32)
        recs = e.debug_rois()
        L = GV["label"]
        for r in recs:
            zs, ys, xs = np.nonzero(L == r[0])
            assert r[1] == len(zs) and (r[4], r[5], r[6], r[7]) == (xs.min(), ys.min(), xs.max() - xs.min() + 1, ys.max() - ys.min() + 1)
    finally:
        e.close()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libnyxus_ref.so did not travel")
def test_config4_style_volume_against_compiled_reference():
    """BASELINE.json configs[3] at reduced size: 78^3 uint16, 8 ellipsoid ROIs (semi-axes 6..14), metaparams of SURVEY section 8d"""
    from nyxus_b200 import synth
    I, L = synth.config4(seed=4, size=78)
    e = _engine3([64, 64, 64, 1, 0])
    try:
        tab, _ = e.featurize_volume(I, L)
        rl, rt, _ = ref.featurize_3d(I.astype(np.uint32), L, greydepth=64, glszm_greydepth=64, ngtdm_greydepth=64, ngtdm_radius=1, nthreads=8)
        assert tab.shape[0] == 8 and np.array_equal(tab[:, 0].astype(np.int64), rl.astype(np.int64))
        bad = tu.mismatches(rt, tab[:, 2:], GV["columns"])
        assert not bad, bad[:8]
    finally:
        e.close()


def test_nyxus3d_featurize_files_nifti(tmp_path):
    import os
    from nyxus_b200 import Nyxus3D, nifti
    ip, mp = os.path.join(tmp_path, "vol_int.nii.gz"), os.path.join(tmp_path, "vol_seg.nii")
    nifti.write(ip, GV["inten"])
    nifti.write(mp, GV["label"])
    nyx = Nyxus3D(["*3D_ALL_INTENSITY*", "3GLCM_ASM", "3GLCM_CONTRAST_AVE", "*3D_GLSZM*", "*3D_NGTDM*"])
    nyx.set_metaparam("3glszm/greydepth=64")
    nyx.set_metaparam("3ngtdm/greydepth=64")
    nyx.set_metaparam("3ngtdm/radius=1")
    df = nyx.featurize_files([ip], [mp], False)
    assert list(df.columns[:4]) == ["intensity_image", "mask_image", "ROI_label", "t_index"]
    assert list(df.intensity_image.unique()) == ["vol_int.nii.gz"] and list(df.mask_image.unique()) == ["vol_seg.nii"]
    assert list(df.ROI_label) == [int(v) for v in GV["labels_matlab64_r1"]]
    cols = GV["columns"]
    ref_tab = GV["table_matlab64_r1"]
    for c in df.columns[4:]:
        assert not tu.mismatches(ref_tab[:, cols.index(c)], df[c].to_numpy()), c
    with pytest.raises(NyxusError_or_ValueError()):
        Nyxus3D(["*3D_ALL*"])                            # holds 3D families outside this path (surface morphology)
    with pytest.raises(ValueError, match="not a valid 3D feature name"):
        Nyxus3D(["GLCM_ASM"])


def NyxusError_or_ValueError():
    from nyxus_b200 import NyxusB200Error
    return (NyxusB200Error, ValueError)


# ------------------------------------------------------------------------------------------------ the mid-size tile classes
@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("ibsi", [False, True])
def test_mid_size_rois_against_compiled_reference(ibsi):
    """ROIs of 1 600 ... 39 800 bounding-box cells: the CTA-per-ROI forms of every family (tiles wider than a warp: GLRLM's
    wrapped diagonals over many 32-cell steps, GLSZM rows in several chunks, GLDZM / neighbourhood kernels with block-wide
    reductions, matrices in global scratch where they exceed shared memory), next to small ROIs of the warp-per-ROI forms"""
    rng = np.random.default_rng(2024)
    H, W = 460, 520
    yy, xx = np.mgrid[0:H, 0:W]
    I = (3000 + 2000 * np.sin(xx / 23.0) * np.cos(yy / 31.0) + rng.normal(0, 500, (H, W))).clip(1, 60000).astype(np.uint32)
    L = np.zeros((H, W), np.uint32)
    boxes = [(5, 5, 40, 40), (5, 60, 64, 70), (5, 150, 120, 130), (140, 5, 199, 200), (140, 220, 33, 290), (200, 220, 250, 35),
             (5, 300, 20, 21), (30, 300, 9, 31)]
    for k, (y0, x0, h, w) in enumerate(boxes):
        cy, cx = y0 + h / 2.0, x0 + w / 2.0
        m = ((yy - cy) / (h / 2.0)) ** 2 + ((xx - cx) / (w / 2.0)) ** 2 <= 1.0 if k % 2 == 0 else \
            (yy >= y0) & (yy < y0 + h) & (xx >= x0) & (xx < x0 + w)
        L[m & (L == 0)] = 7 + 5 * k
    L[200:215, 60:90] = 0                     # a hole in the largest ROI
    I[rng.random((H, W)) < 0.01] = 0
    if ibsi:
        I = (I >> 8).astype(np.uint32)        # IBSI mode indexes the matrices by raw intensity
    e = _engine(ibsi=ibsi)
    try:
        tab, _ = e.featurize(I, L)
    finally:
        e.close()
    rl, rt, _ = ref.featurize_2d(I, L, ibsi=ibsi, nthreads=8)
    assert np.array_equal(tab[:, 0].astype(np.int64), rl.astype(np.int64))
    bad = tu.mismatches(rt, tab[:, 2:], COLS)
    assert not bad, bad[:8]


# ------------------------------------------------------------------------------------------------ large ROIs (giant.cuh)
_GIANT = tu.giant_cases()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", _GIANT, ids=[c[0] for c in _GIANT])
def test_giant_roi_table_and_matrices_match_reference(case):
    """ROIs whose AABB exceeds the shared-memory size classes (> 40000 cells) run through the whole-GPU kernels of
    giant.cuh: the full 337-column table within 1e-9 of the compiled reference and every matrix entry exact."""
    name, I, L, kw = case
    g, ibsi = kw.get("g", 64), kw.get("ibsi", False)
    e = _engine(coarse_gray_depth=g, ibsi=ibsi)
    try:
        tab, _ = e.featurize(I, L)
        rl, rt, _ = ref.featurize_2d(I, L, greydepth=g, ibsi=ibsi, nthreads=8)
        assert e.columns == ref.columns()
        assert np.array_equal(tab[:, 0].astype(np.int64), rl.astype(np.int64))
        bad = tu.mismatches(rt, tab[:, 2:], e.columns)
        assert not bad, (name, bad[:8])
        recs = e.debug_rois()
        big = [r for r in recs if r[6] * r[7] > 40000]
        assert big, "case holds no giant ROI"
        for r in big[:2]:
            lab, x0, y0, w, h = int(r[0]), int(r[4]), int(r[5]), int(r[6]), int(r[7])
            mask = L[y0:y0 + h, x0:x0 + w] == lab
            tile = I[y0:y0 + h, x0:x0 + w].astype(np.uint32) * mask
            rm = ref.roi_matrices(tile, mask, g, ibsi)
            for fam in tu.FAMILIES:
                got, svals = e.debug_matrix(lab, fam)
                assert sorted(got) == sorted(rm[fam]), (name, fam, lab)
                if fam == "NGTDM" and len(rm["NGTDM_S"]):
                    order = np.argsort([t[1] for t in got])
                    assert np.allclose(np.array(svals)[order], rm["NGTDM_S"], rtol=1e-12, atol=1e-9)
    finally:
        e.close()


def test_giant_whole_slide_properties():
    """BASELINE configs[2] shape at 4096^2 (256 grey levels, GLCM + GLRLM + GLSZM, label == 1 everywhere): size-independent
    properties -- a transposed slide swaps the 0 and 90 degree GLCM / GLRLM columns and leaves 45 degrees alone; run
    percentage and zone percentage lie in (0, 1]; repeated calls are bit-identical (integer matrices, fixed-order FP64)."""
    from nyxus_b200 import synth
    I, L = synth.config3(size=2048)
    e = _engine(["*ALL_GLCM*", "*ALL_GLRLM*", "*ALL_GLSZM*"], coarse_gray_depth=256)
    try:
        t1, _ = e.featurize(I, L)
        t2, _ = e.featurize(I, L)
        assert np.array_equal(t1, t2)
        tt, _ = e.featurize(np.ascontiguousarray(I.T), L)
        cols = e.columns
        c = {n: i + 2 for i, n in enumerate(cols)}
        for f in ("GLCM_ASM", "GLCM_CONTRAST", "GLCM_JE", "GLRLM_SRE", "GLRLM_RP", "GLRLM_RE"):
            assert np.isclose(t1[0, c[f + "_0"]], tt[0, c[f + "_90"]], rtol=1e-12), f
            assert np.isclose(t1[0, c[f + "_90"]], tt[0, c[f + "_0"]], rtol=1e-12), f
        for f in ("GLCM_ASM", "GLCM_CONTRAST", "GLCM_JE"):   # MATLAB mode: asymmetric matrix, transposition flips (dx,dy)=(1,1) onto itself
            assert np.isclose(t1[0, c[f + "_45"]], tt[0, c[f + "_45"]], rtol=1e-12), f
        assert 0 < t1[0, c["GLRLM_RP_0"]] <= 1 and 0 < t1[0, c["GLSZM_ZP"]] <= 1
        assert t1[0, c["GLSZM_SAE"]] == 0.0     # >= 65536 cells: the reference's int j*j wrap turns SAE into NaN -> noval
    finally:
        e.close()


# ------------------------------------------------------------------------------------------------ several devices, one call
def _multi_slides(counts, seed=31):
    from nyxus_b200 import synth
    out = []
    for k, n in enumerate(counts):
        I, L = synth.blob_slide(160, 192, 32, 7, 12, seed=seed + k, jitter=1)
        L = np.where(L <= n, L, 0).astype(np.uint32)        # slide k keeps its first n ROIs
        out.append((np.ascontiguousarray(I), L))
    return out


@pytest.mark.parametrize("nctx,pad", [(1, 65536), (3, 65536), (2, 0)])
def test_multi_device_list_call_gives_the_single_device_table(nctx, pad):
    """nyxb200_featurize_2d_list_multi: slides sharded over several contexts (here on ONE device: each context has its own
    host thread, streams and buffers, which is all the call relies on), rows claimed in slide order into one table.  pad = 0
    makes the first estimate too small for the later slides, which exercises the second pass."""
    import torch
    from nyxus_b200 import Engine
    from nyxus_b200.engine import featurize_list_multi
    slides = _multi_slides([3, 30, 12, 0, 25, 7, 30])
    single = _engine()
    engines = [_engine() for _ in range(nctx)]
    try:
        want = [single.featurize(I, L) for I, L in slides]
        want_tab = np.concatenate([w[0] for w in want])
        want_idx = np.concatenate([np.full(len(w[0]), k, np.int32) for k, w in enumerate(want)])
        engines[0].set_param("multi_row_pad", pad)
        for e in engines[1:]:
            e.set_param("multi_row_pad", pad)
        for on_device in (False, True):
            if on_device:
                keep = [(torch.from_numpy(I).cuda(), torch.from_numpy(L.view(np.int32)).cuda()) for I, L in slides]
                ips, lps = [t[0].data_ptr() for t in keep], [t[1].data_ptr() for t in keep]
            else:
                ips, lps = [I.ctypes.data for I, _ in slides], [L.ctypes.data for _, L in slides]
            for rep in range(2):        # the second call reuses the pooled table
                tab, idx = featurize_list_multi(engines, ips, True, lps, 160, 192, on_device=on_device,
                                                weight=[int(np.count_nonzero(L)) + 1 for _, L in slides] if rep else None)
                assert np.array_equal(idx, want_idx)
                assert np.array_equal(tab, want_tab)
    finally:
        single.close()
        for e in engines:
            e.close()


# ------------------------------------------------------------------------------------------------ 3D: large cubes, NGTDM radius > 1
@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libnyxus_ref.so did not travel")
@pytest.mark.parametrize("radius,g", [(1, 64), (2, 32)])
def test_large_3d_rois_and_ngtdm_radius_against_compiled_reference(radius, g):
    """ROIs whose bounding cube exceeds the shared-memory classes (> 40000 voxels: organ-sized masks) run through the
    global-memory path (k_roi_family3_big: 32-bit union-find, 64-bit hash keys); small ROIs of the same volume keep the
    shared-memory kernels.  3ngtdm/radius = 2 is the 124-neighbour Chebyshev cube (3d_ngtdm.cpp:69-94)."""
    rng = np.random.default_rng(40 + radius)
    D, H, W = 50, 72, 80
    zz, yy, xx = np.mgrid[0:D, 0:H, 0:W]
    L = np.zeros((D, H, W), np.uint32)
    big = (((zz - 24) / 22.0) ** 2 + ((yy - 35) / 30.0) ** 2 + ((xx - 38) / 33.0) ** 2) < 1        # cube ~ 45 x 60 x 66 = 178k voxels
    big &= ~((((zz - 20) / 8.0) ** 2 + ((yy - 30) / 9.0) ** 2 + ((xx - 40) / 10.0) ** 2) < 1)      # a hole
    L[big] = 5
    L[2:9, 60:70, 68:79] = 9                                # small ROIs next to it
    L[40:49, 2:10, 2:12] = 2
    field = (np.sin(zz / 5.0) + np.cos(yy / 7.0) * np.sin(xx / 6.0) + 2.2) * 300.0
    I = (field + rng.normal(0, 60, (D, H, W))).clip(1, 4000).astype(np.uint16)
    I[rng.random((D, H, W)) < 0.005] = 0
    e = _engine3([g, g, g, radius, 0])
    try:
        tab, _ = e.featurize_volume(I, L)
        rl, rt, _ = ref.featurize_3d(I.astype(np.uint32), L, greydepth=g, glszm_greydepth=g, ngtdm_greydepth=g, ngtdm_radius=radius, nthreads=8)
        assert np.array_equal(tab[:, 0].astype(np.int64), rl.astype(np.int64)) and list(rl) == [2, 5, 9]
        recs = e.debug_rois()
        assert max(int(r[6]) * int(r[7]) for r in recs) * 45 > 40000
        bad = tu.mismatches(rt, tab[:, 2:], GV["columns"])
        assert not bad, bad[:8]
    finally:
        e.close()


# ------------------------------------------------------------------------------------------------ *BASIC_MORPHOLOGY* 
def test_basic_morphology_against_reference_golden():
    """golden_morph.npz = the reference's own pybind module on a slide of blobs + degenerate shapes (1 and 2 pixels, 2x2
    square, line, ring, all-zero intensities, sparse ROI with a huge bounding box).  Alone, and together with every other
    family: the 15 columns come out between the first-order and the GLCM columns, like the reference's."""
    Zm = np.load(os.path.join(tu.GOLDEN, "golden_morph.npz"))
    I, L = Zm["inten"], Zm["label"]
    e = _engine(["*BASIC_MORPHOLOGY*"])
    try:
        tab, _ = e.featurize(I, L)
        assert e.columns == [str(c) for c in Zm["columns"]]
        assert np.array_equal(tab[:, 0].astype(np.int64), Zm["labels"])
        bad = tu.mismatches(Zm["table"], tab[:, 2:], e.columns)
        assert not bad, bad[:8]
    finally:
        e.close()
    e = _engine(["*ALL_INTENSITY*", "*BASIC_MORPHOLOGY*", "*ALL_TEXTURE*"])
    try:
        tab, _ = e.featurize(I.astype(np.uint16), L)
        assert e.columns == [str(c) for c in Zm["all_columns"]]
        bad = tu.mismatches(Zm["all_table"], tab[:, 2:], e.columns)
        assert not bad, bad[:8]
        e.set_param("roi_batches", 3)
        tab3, _ = e.featurize(I.astype(np.uint16), L)
        assert np.array_equal(tab3, tab)
    finally:
        e.close()
    from nyxus_b200 import Nyxus
    df = Nyxus(["*BASIC_MORPHOLOGY*", "MEAN"]).featurize(I, L)
    assert list(df.columns[4:6]) == ["MEAN", "AREA_PIXELS_COUNT"] and df.shape[1] == 4 + 16


# ------------------------------------------------------------------------------------------------ *ALL_IH* 
@pytest.mark.parametrize("g", [64, 20, 7])
def test_intensity_histogram_family_against_reference_golden(g):
    """golden_ih.npz = the reference's own pybind module (ibsi=True) on blobs + a constant ROI, a ROI with fewer distinct
    values than bins and a three-spike ROI, at three grey depths"""
    Zi = np.load(os.path.join(tu.GOLDEN, "golden_ih.npz"))
    e = _engine(["*ALL_IH*"], ibsi=True, coarse_gray_depth=g)
    try:
        tab, _ = e.featurize(Zi["inten"], Zi["label"])
        assert e.columns == [str(c) for c in Zi["columns"]]
        assert np.array_equal(tab[:, 0].astype(np.int64), Zi["labels"])
        bad = tu.mismatches(Zi["table_%d" % g], tab[:, 2:], e.columns)
        assert not bad, bad[:8]
    finally:
        e.close()


def test_intensity_histogram_family_needs_ibsi_and_orders_by_code():
    Zi = np.load(os.path.join(tu.GOLDEN, "golden_ih.npz"))
    e = _engine(["*ALL_IH*", "MEAN"], ibsi=False)           # stripped like env_features.cpp:513-527 does
    try:
        assert e.columns == [str(c) for c in Zi["columns_ibsi_off"]]
        e.set_param("ibsi", 1)
        assert len(e.columns) == 47 and e.columns[0] == "MEAN" and e.columns[1] == "IH_MEAN_VAL"
    finally:
        e.close()
    e = _engine(["*ALL_IH*", "*ALL_INTENSITY*", "*BASIC_MORPHOLOGY*", "GLCM_ASM", "NGTDM_STRENGTH"], ibsi=True)
    try:
        assert e.columns == [str(c) for c in Zi["mixed_columns"]]
    finally:
        e.close()


# ------------------------------------------------------------------------------------------------ HISTOGRAM (opt-in, multi-valued)
@pytest.mark.parametrize("g,ibsi", [(64, False), (20, False), (7, False), (20, True)])
def test_histogram_feature_against_reference_golden(g, ibsi):
    """golden_hist.npz = the reference's own pybind module: HISTOGRAM_BIN_0 .. _N-1 (raw counts of the first-order family's
    grey-depth histogram, N = coarse grey depth) next to MEAN / ENTROPY; the slide holds a blank ROI (bins padded with zeros)"""
    Zh = np.load(os.path.join(tu.GOLDEN, "golden_hist.npz"))
    Zi = np.load(os.path.join(tu.GOLDEN, "golden_ih.npz"))
    I = Zi["inten"].copy()
    I[2:6, 2:8] = 0
    tag = "%d%s" % (g, "_ibsi" if ibsi else "")
    e = _engine(["HISTOGRAM", "MEAN", "ENTROPY"], ibsi=ibsi, coarse_gray_depth=g)
    try:
        tab, _ = e.featurize(I, Zi["label"])
        assert e.columns == [str(c) for c in Zh["columns_" + tag]]
        assert np.array_equal(tab[:, 0].astype(np.int64), Zh["labels"])
        assert np.array_equal(tab[:, 4:], Zh["table_" + tag][:, 2:])                     # the bins: exact
        bad = tu.mismatches(Zh["table_" + tag][:, :2], tab[:, 2:4], e.columns[:2])
        assert not bad, bad[:8]
        t16, _ = e.featurize(I.astype(np.uint16), Zi["label"])
        assert np.array_equal(t16, tab)
        e.set_param("coarse_gray_depth", 5)                                             # the block follows the grey depth
        assert e.columns == ["ENTROPY", "MEAN"] + ["HISTOGRAM_BIN_%d" % k for k in range(5)]
        t5, _ = e.featurize(I, Zi["label"])
        assert t5.shape[1] == 2 + 7 and np.array_equal(t5[:, 4:].sum(axis=1)[1:], tab[:, 4:].sum(axis=1)[1:])
    finally:
        e.close()


def test_histogram_feature_column_order_and_selection():
    Zh = np.load(os.path.join(tu.GOLDEN, "golden_hist.npz"))
    e = _engine(["IH_MEAN_VAL", "HISTOGRAM", "IMOM_WHU7", "GLCM_ASM", "MEAN"], ibsi=True, coarse_gray_depth=5)
    try:
        assert e.columns == [str(c) for c in Zh["mixed_columns"]]
    finally:
        e.close()
    e = _engine(["*ALL_INTENSITY*"])                           # opt-in: not part of the first-order group (intensity.cpp:50-54)
    try:
        assert len(e.columns) == 36 and not any(c.startswith("HISTOGRAM") for c in e.columns)
    finally:
        e.close()


# ------------------------------------------------------------------------------------------------ geometric moments
def test_unweighted_geometric_moments_against_reference_golden():
    """golden_geomom.npz = the reference's own pybind module: the 118 unweighted shape / intensity moments (raw, central,
    normalised raw, normalised central, Hu)"""
    Zg = np.load(os.path.join(tu.GOLDEN, "golden_geomom.npz"))
    cols = [str(c) for c in Zg["columns"]]
    e = _engine(cols)
    try:
        tab, _ = e.featurize(Zg["inten"], Zg["label"])
        assert e.columns == cols and np.array_equal(tab[:, 0].astype(np.int64), Zg["labels"])
        bad = tu.mismatches(Zg["table"], tab[:, 2:], cols)
        assert not bad, bad[:8]
        e.select(["IMOM_HU1", "IMOM_RM_11"])            # intensity variant alone
        t2, _ = e.featurize(Zg["inten"].astype(np.uint16), Zg["label"])
        assert e.columns == ["IMOM_RM_11", "IMOM_HU1"]
        assert not tu.mismatches(Zg["table"][:, [cols.index(c) for c in e.columns]], t2[:, 2:], e.columns)
    finally:
        e.close()


def test_weighted_geometric_moments_against_reference_golden():
    """golden_wmom.npz = the reference's own pybind module on 51 ROIs (concave blobs, a ring, one-pixel-wide bars, a plus sign,
    speckles, two-part ROIs, a ROI at the slide corner): the 62 distance-to-contour weighted moments -- contour tracer, the
    reference's approximate ordered-contour distance, float weights -- and the three moment groups of the reference"""
    Zw = np.load(os.path.join(tu.GOLDEN, "golden_wmom.npz"))
    cols = [str(c) for c in Zw["columns"]]
    wcols = [c for c in cols if c.startswith(("WEIGHTED_", "WT_NORM_CTR_MOM_", "IMOM_W"))]
    assert len(wcols) == 62
    I, L = tu.wmom_slide()
    e = _engine(wcols)
    try:
        tab, _ = e.featurize(I, L)
        assert e.columns == wcols and np.array_equal(tab[:, 0].astype(np.int64), Zw["labels"])
        ref = tu.finite0(Zw["table"][:, [cols.index(c) for c in wcols]])
        bad = tu.mismatches(ref, tab[:, 2:], wcols)
        assert not bad, bad[:8]
        t16, _ = e.featurize(I.astype(np.uint16), L)          # uint16 intensities: the same table
        assert np.array_equal(t16, tab)
        # a ROI beyond the 40 000-cell classes (CTA per ROI) with holes, a slit and pinholes; a thin diagonal band
        Zb = np.load(os.path.join(tu.GOLDEN, "golden_wmom_big.npz"))
        Ib, Lb = tu.wmom_slide_big()
        tb, _ = e.featurize(Ib, Lb)
        assert np.array_equal(tb[:, 0].astype(np.int64), Zb["labels"])
        bcols = [str(c) for c in Zb["columns"]]
        badb = tu.mismatches(tu.finite0(Zb["table"][:, [bcols.index(c) for c in wcols]]), tb[:, 2:], wcols)
        assert not badb, badb[:8]
        e.select(["*SGEOMOMS*"])                               # env_features.cpp:329-333: 59 unweighted + 31 weighted shape moments
        assert len(e.columns) == 90 and "WEIGHTED_HU_M7" in e.columns and "SPAT_MOMENT_00" in e.columns and not any(c.startswith("IMOM") for c in e.columns)
        ts, _ = e.featurize(I, L)
        sw = [c for c in e.columns if c in wcols]
        assert not tu.mismatches(ref[:, [wcols.index(c) for c in sw]], ts[:, 2:][:, [e.columns.index(c) for c in sw]], sw)
        e.select(["*IGEOMOMS*"])
        assert len(e.columns) == 90 and all(c.startswith("IMOM") for c in e.columns)
        e.select(["*GEOMOMS*"])
        assert len(e.columns) == 180
    finally:
        e.close()


# ------------------------------------------------------------------------------------------------ more 3D families
@pytest.mark.parametrize("g", [0, 16, -12])
def test_3d_gldm_against_reference_golden(g):
    from nyxus_b200 import Engine
    Zv = np.load(os.path.join(tu.GOLDEN, "golden_vol_more.npz"))
    e = Engine(["*3D_GLDM*"], dim=3)
    try:
        e.set_param("3gldm/greydepth", g)
        tab, _ = e.featurize_volume(Zv["vol"], Zv["label"])
        assert e.columns == [str(c) for c in Zv["GLDM/columns"]]
        assert np.array_equal(tab[:, 0].astype(np.int64), Zv["GLDM/labels"])
        bad = tu.mismatches(Zv["GLDM/table_%d" % g], tab[:, 2:], e.columns)
        assert not bad, bad[:8]
    finally:
        e.close()


@pytest.mark.parametrize("g", [0, 16, -12])
def test_3d_glrlm_against_reference_golden(g):
    from nyxus_b200 import Engine
    Zv = np.load(os.path.join(tu.GOLDEN, "golden_vol_more.npz"))
    e = Engine(["*3D_GLRLM*"], dim=3)
    try:
        e.set_param("3glrlm/greydepth", g)
        tab, _ = e.featurize_volume(Zv["vol"], Zv["label"])
        assert e.columns == [str(c) for c in Zv["GLRLM/columns"]]
        bad = tu.mismatches(Zv["GLRLM/table_%d" % g], tab[:, 2:], e.columns)
        assert not bad, bad[:8]
    finally:
        e.close()


@pytest.mark.parametrize("fam", ["GLDZM", "NGLDM"])
@pytest.mark.parametrize("tag", ["64", "20", "ibsi"])
def test_3d_gldzm_ngldm_against_reference_golden(fam, tag):
    from nyxus_b200 import Engine
    Zv = np.load(os.path.join(tu.GOLDEN, "golden_vol_more.npz"))
    kw = {"ibsi": True} if tag == "ibsi" else {"coarse_gray_depth": int(tag)}
    e = Engine(["*3D_%s*" % fam], dim=3, **kw)
    try:
        tab, _ = e.featurize_volume(Zv["vol"], Zv["label"])
        assert e.columns == [str(c) for c in Zv[fam + "/columns"]]
        assert np.array_equal(tab[:, 0].astype(np.int64), Zv[fam + "/labels"])
        bad = tu.mismatches(Zv["%s/table_%s" % (fam, tag)], tab[:, 2:], e.columns)
        assert not bad, bad[:8]
    finally:
        e.close()


def test_3d_all_texture_group_and_edge_case_rois():
    """*3D_ALL_TEXTURE* = the seven 3D texture families in one call; the edge-case volume (slab, plate, blank ROI, single voxel,
    zeros) of golden_vol_more.npz against the reference's values for the four later families"""
    from nyxus_b200 import Engine
    Zv = np.load(os.path.join(tu.GOLDEN, "golden_vol_more.npz"))
    e = Engine(["*3D_ALL_TEXTURE*"], dim=3, coarse_gray_depth=5)
    try:
        e.set_param("3gldm/greydepth", 5)
        e.set_param("3glrlm/greydepth", -4)
        tab, _ = e.featurize_volume(Zv["e/vol"], Zv["e/label"])
        cols = e.columns
        assert np.array_equal(tab[:, 0].astype(np.int64), Zv["e/GLDM/labels"])
        for fam, tag in (("GLDM", "5"), ("GLRLM", "-4"), ("GLDZM", "5"), ("NGLDM", "5")):
            fc = [str(c) for c in Zv[fam + "/columns"]]
            sel = [cols.index(c) for c in fc]
            bad = tu.mismatches(tu.finite0(Zv["e/%s/table_%s" % (fam, tag)]), tu.finite0(tab[:, 2:][:, sel]), fc)
            assert not bad, (fam, bad[:8])
    finally:
        e.close()


def test_3d_later_families_large_cube_path_against_reference_golden():
    """the four later 3D families through k_roi_family3_big (a ROI whose cube exceeds the 40 000-voxel shared-memory classes)
    and the shared-memory classes in one volume, against the reference's own pybind module (golden_vol_big.npz)"""
    from nyxus_b200 import Engine
    Zb = np.load(os.path.join(tu.GOLDEN, "golden_vol_big.npz"))
    Zm = np.load(os.path.join(tu.GOLDEN, "golden_vol_more.npz"))
    vol, lab = tu.big_volume_3d()
    for ibsi, tag_dr, tag_cz in ((False, ("0", "32", "-20"), "64"), (True, (), "ibsi")):
        e = Engine(["*3D_GLDM*", "*3D_GLRLM*", "*3D_GLDZM*", "*3D_NGLDM*"], dim=3, ibsi=ibsi)
        try:
            for t in (tag_dr or ("0",)):
                e.set_param("3gldm/greydepth", int(t))
                e.set_param("3glrlm/greydepth", int(t))
                tab, _ = e.featurize_volume(vol, lab)
                cols = e.columns
                assert np.array_equal(tab[:, 0].astype(np.int64), Zb["GLDZM/labels"])
                fams = [("GLDZM", tag_cz), ("NGLDM", tag_cz)] + ([("GLDM", t), ("GLRLM", t)] if tag_dr else [])
                for fam, tag in fams:
                    fc = [str(c) for c in Zm[fam + "/columns"]]
                    sel = [cols.index(c) for c in fc]
                    bad = tu.mismatches(tu.finite0(Zb["%s/table_%s" % (fam, tag)]), tu.finite0(tab[:, 2:][:, sel]), fc)
                    assert not bad, (fam, tag, bad[:8])
        finally:
            e.close()
