"""CPU tests of the 3D (Nyxus3D) path: the product's 3D feature bodies (host-emulated, test harness only) against golden
vectors produced by the compiled reference, the NIfTI reader/writer, and the library's 3D column contract."""
import os

import numpy as np
import pytest

import nyx_testutil as tu
from oracle import ref

GV = tu.load_golden_vol()
CASES = ["matlab64_r1", "defaults", "g16_r1", "ibsi_r1"]


@pytest.mark.parametrize("case", CASES)
def test_3d_feature_bodies_match_reference_golden(case):
    L = tu.build_emu()
    g, gz, gn, rad, ibsi = [int(v) for v in GV["params_" + case]]
    vol = GV["inten_ibsi"] if ibsi else GV["inten"]
    labels = GV["labels_" + case]
    table = GV["table_" + case]
    assert list(labels) == sorted(int(v) for v in np.unique(GV["label"]) if v)
    for k, lab in enumerate(labels):
        got = tu.emu_roi_features_3d(L, vol, GV["label"], lab, g, gz, gn, rad, ibsi, GV["columns"])
        bad = tu.mismatches(table[k], got, GV["columns"])
        assert not bad, (case, int(lab), bad[:6])


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libnyxus_ref.so not built")
def test_compiled_reference_reproduces_3d_golden():
    for case in CASES:
        g, gz, gn, rad, ibsi = [int(v) for v in GV["params_" + case]]
        vol = GV["inten_ibsi"] if ibsi else GV["inten"]
        labs, tab, _ = ref.featurize_3d(vol.astype(np.uint32), GV["label"], greydepth=g, glszm_greydepth=gz, ngtdm_greydepth=gn,
                                        ngtdm_radius=rad, ibsi=bool(ibsi))
        assert np.array_equal(labs, GV["labels_" + case])
        assert np.array_equal(tab, GV["table_" + case])
    assert ref.columns_3d() == GV["columns"]


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libnyxus_ref.so not built")
def test_3d_feature_bodies_randomised_against_compiled_reference():
    L = tu.build_emu()
    rng = np.random.default_rng(77)
    cols = GV["columns"]
    for it in range(24):
        D, H, W = [int(rng.integers(2, 8)) for _ in range(3)]
        kind = it % 3
        V = rng.integers(1, 5000, (D, H, W)) if kind == 0 else rng.integers(1, 6, (D, H, W)) if kind == 1 else rng.integers(0, 200, (D, H, W))
        M = (rng.random((D, H, W)) < rng.choice([0.6, 0.9, 1.0])).astype(np.uint32) * 7
        if not M.any():
            M[0, 0, 0] = 7
        ibsi = bool(it % 4 == 0 and kind > 0)
        g, gz, gn, rad = int(rng.choice([8, 64])), int(rng.choice([16, 64])), int(rng.choice([16, 64])), int(rng.choice([0, 1, 1]))
        labs, tab, _ = ref.featurize_3d(V.astype(np.uint32), M, greydepth=g, glszm_greydepth=gz, ngtdm_greydepth=gn, ngtdm_radius=rad, ibsi=ibsi)
        got = tu.emu_roi_features_3d(L, V, M, 7, g, gz, gn, rad, ibsi, cols)
        assert not tu.mismatches(tab[0], got, cols), (it, kind, ibsi)
ZM = np.load(os.path.join(tu.GOLDEN, "golden_vol_more.npz"))


def test_nifti_round_trip(tmp_path):
    from nyxus_b200 import nifti
    rng = np.random.default_rng(3)
    for dt, name in [(np.uint8, "a.nii"), (np.uint16, "b.nii.gz"), (np.int16, "c.nii"), (np.uint32, "d.nii.gz"), (np.float32, "e.nii")]:
        v = rng.integers(0, 200, (5, 6, 7)).astype(dt)
        p = os.path.join(tmp_path, name)
        nifti.write(p, v)
        r = nifti.read(p)
        assert r.shape == (5, 6, 7) and np.array_equal(r, v)
        assert np.array_equal(nifti.to_grey(r), v.astype(np.uint32))
    with pytest.raises(ValueError):
        nifti.to_grey(np.array([[[0.5]]], np.float32))
    with pytest.raises(ValueError):
        nifti.to_grey(np.array([[[-3]]], np.int16))


def test_3d_column_contract_of_the_library():
    """the generated 3D table (nyx_feature_table3d.inc) lists the reference's 116 columns in Feature3D code order"""
    inc = os.path.join(tu.ROOT, "nyxus_b200", "csrc", "nyx_feature_table3d.inc")
    rows = [ln for ln in open(inc) if ln.startswith("{")]
    names = [ln.split('"')[1] for ln in rows]
    codes = [int(ln.split(",")[1]) for ln in rows]
    assert codes == sorted(codes)
    extra = [n for n in names if n.startswith(("3GLDM_", "3GLRLM_", "3GLDZM_", "3NGLDM_"))]          # the later 3D families
    assert [n for n in names if n not in extra] == GV["columns"] and len(names) == 116 + len(extra) and len(extra) == 14 + 32 + 18 + 19
    # the reference's own column order of every family, as its pybind module returned it (golden_vol_more.npz)
    for fam in ("GLDM", "GLRLM", "GLDZM", "NGLDM"):
        assert [n for n in names if n.startswith("3%s_" % fam)] == [str(c) for c in ZM[fam + "/columns"]]


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libnyxus_ref.so not built")
def test_3d_ngtdm_radius_above_one_and_large_cube_glszm_against_compiled_reference():
    """3ngtdm/radius = 2, 3 (3d_ngtdm.cpp:69-94: the (2R+1)^3 - 1 Chebyshev neighbourhood) and the large-cube GLSZM body
    (32-bit union-find, 64-bit hash keys) give the reference's values"""
    L = tu.build_emu()
    rng = np.random.default_rng(91)
    cols = GV["columns"]
    try:
        for it in range(12):
            D, H, W = [int(rng.integers(3, 10)) for _ in range(3)]
            V = rng.integers(1, 3000, (D, H, W)) if it % 2 == 0 else rng.integers(1, 7, (D, H, W))
            M = (rng.random((D, H, W)) < rng.choice([0.7, 1.0])).astype(np.uint32) * 7
            M[0, 0, 0] = 7
            rad = int(rng.choice([2, 3]))
            gz, gn = int(rng.choice([0, 16, 64])), int(rng.choice([16, 64]))
            if gz == 0 and it % 2 == 0:
                gz = 32
            labs, tab, _ = ref.featurize_3d(V.astype(np.uint32), M, greydepth=64, glszm_greydepth=gz, ngtdm_greydepth=gn, ngtdm_radius=rad)
            L.nyxemu_set_big3d(it % 3 != 0)
            got = tu.emu_roi_features_3d(L, V, M, 7, 64, gz, gn, rad, False, cols)
            assert not tu.mismatches(tab[0], got, cols), (it, rad, gz, gn)
    finally:
        L.nyxemu_set_big3d(0)


@pytest.mark.parametrize("g", [0, 16, -12])
def test_3d_gldm_body_matches_reference_golden(g):
    """3D GLDM: golden_vol_more.npz = the reference's own pybind module on a NIfTI volume (make_golden3d_more.py)"""
    L = tu.build_emu()
    cols = [str(c) for c in ZM["GLDM/columns"]]
    try:
        L.nyxemu_set_gldm3_greydepth(g)
        for k, lab in enumerate(ZM["GLDM/labels"]):
            got = tu.emu_roi_features_3d(L, ZM["vol"], ZM["label"], int(lab), 64, 0, 0, 0, False, cols, fam_mask=16)
            bad = tu.mismatches(ZM["GLDM/table_%d" % g][k], got, cols)
            assert not bad, (g, int(lab), bad[:4])
    finally:
        L.nyxemu_set_gldm3_greydepth(0)


@pytest.mark.parametrize("g", [0, 16, -12])
def test_3d_glrlm_body_matches_reference_golden(g):
    """3D GLRLM: direction 0 of 13 + the 13-direction averages (3d_glrlm.cpp:266-299)"""
    L = tu.build_emu()
    cols = [str(c) for c in ZM["GLRLM/columns"]]
    try:
        L.nyxemu_set_glrlm3_greydepth(g)
        for k, lab in enumerate(ZM["GLRLM/labels"]):
            got = tu.emu_roi_features_3d(L, ZM["vol"], ZM["label"], int(lab), 64, 0, 0, 0, False, cols, fam_mask=32)
            bad = tu.mismatches(ZM["GLRLM/table_%d" % g][k], got, cols)
            assert not bad, (g, int(lab), bad[:4])
    finally:
        L.nyxemu_set_glrlm3_greydepth(0)


@pytest.mark.parametrize("fam,bit", [("GLDZM", 64), ("NGLDM", 128)])
@pytest.mark.parametrize("tag", ["64", "20", "ibsi"])
def test_3d_gldzm_ngldm_bodies_match_reference_golden(fam, bit, tag):
    """3D GLDZM (6-connected zones, in-slice distances) and 3D NGLDM (24-neighbour dependence of interior voxels): both are
    binned by the coarse grey depth of the constructor; 'ibsi' = IBSI mode (no binning)"""
    L = tu.build_emu()
    cols = [str(c) for c in ZM[fam + "/columns"]]
    depth, ibsi = (64, True) if tag == "ibsi" else (int(tag), False)
    for k, lab in enumerate(ZM[fam + "/labels"]):
        got = tu.emu_roi_features_3d(L, ZM["vol"], ZM["label"], int(lab), depth, 0, 0, 0, ibsi, cols, fam_mask=bit)
        bad = tu.mismatches(ZM["%s/table_%s" % (fam, tag)][k], got, cols)
        assert not bad, (fam, tag, int(lab), bad[:4])


def _edge_rows(L, fam, tag, bit, depth, ibsi):
    cols = [str(c) for c in ZM[fam + "/columns"]]
    for k, lab in enumerate(ZM["e/%s/labels" % fam]):
        got = tu.emu_roi_features_3d(L, ZM["e/vol"], ZM["e/label"], int(lab), depth, 0, 0, 0, ibsi, cols, fam_mask=bit)
        bad = tu.mismatches(tu.finite0(ZM["e/%s/table_%s" % (fam, tag)][k]), got, cols)
        assert not bad, (fam, tag, int(lab), bad[:4])


@pytest.mark.parametrize("g", [0, 5, -4])
def test_3d_gldm_glrlm_edge_cases_against_reference_golden(g):
    """one-slice slab, two-slice plate, constant (blank) ROI, single voxel, intensity-0 voxels, long runs"""
    L = tu.build_emu()
    try:
        L.nyxemu_set_gldm3_greydepth(g)
        L.nyxemu_set_glrlm3_greydepth(g)
        _edge_rows(L, "GLDM", str(g), 16, 64, False)
        _edge_rows(L, "GLRLM", str(g), 32, 64, False)
    finally:
        L.nyxemu_set_gldm3_greydepth(0)
        L.nyxemu_set_glrlm3_greydepth(0)


@pytest.mark.parametrize("tag", ["64", "5", "ibsi"])
def test_3d_gldzm_ngldm_edge_cases_against_reference_golden(tag):
    L = tu.build_emu()
    depth, ibsi = (64, True) if tag == "ibsi" else (int(tag), False)
    _edge_rows(L, "GLDZM", tag, 64, depth, ibsi)
    _edge_rows(L, "NGLDM", tag, 128, depth, ibsi)


ZB = np.load(os.path.join(tu.GOLDEN, "golden_vol_big.npz"))


def test_3d_later_families_on_large_cubes_match_reference_golden():
    """golden_vol_big.npz: an organ-sized ROI (cube ~ 110 000 voxels), a medium and two small ones; the bodies of the four later
    3D families against the reference's own values (host emulation; the GPU test runs the same rows through the large-cube
    kernel)"""
    L = tu.build_emu()
    vol, lab = tu.big_volume_3d()
    try:
        for fam, bit, tags in (("GLDM", 16, ("0", "32", "-20")), ("GLRLM", 32, ("0", "32", "-20")), ("GLDZM", 64, ("64", "ibsi")), ("NGLDM", 128, ("64", "ibsi"))):
            cols = [str(c) for c in ZM[fam + "/columns"]]
            for tag in tags:
                if fam in ("GLDM", "GLRLM"):
                    L.nyxemu_set_gldm3_greydepth(int(tag))
                    L.nyxemu_set_glrlm3_greydepth(int(tag))
                ibsi = tag == "ibsi"
                for k, lb in enumerate(ZB[fam + "/labels"]):
                    got = tu.emu_roi_features_3d(L, vol, lab, int(lb), 64, 0, 0, 0, ibsi, cols, fam_mask=bit)
                    bad = tu.mismatches(tu.finite0(ZB["%s/table_%s" % (fam, tag)][k]), got, cols)
                    assert not bad, (fam, tag, int(lb), bad[:4])
    finally:
        L.nyxemu_set_gldm3_greydepth(0)
        L.nyxemu_set_glrlm3_greydepth(0)
