"""GPU parity tests (-m gpu) of PyRadiomics binning (negative grey depth) and of a GLCM grey depth that differs from the
general one, through the C ABI, against tests/golden/golden_radiomics.npz (compiled reference, make_golden_rad.py) and
against the compiled reference itself on fresh inputs.  Matrices exact, features within 1e-9 relative."""
import os

import numpy as np
import pytest

import nyx_testutil as tu
from oracle import ref

pytestmark = pytest.mark.gpu

Z = np.load(os.path.join(tu.GOLDEN, "golden_radiomics.npz"))
COLS = [str(c) for c in Z["columns"]]
COLS3 = [str(c) for c in Z["columns3d"]]
NAMES = [str(n) for n in Z["names"]]
RAD_FAMS = [str(f) for f in Z["rad_families"]]
RAD_FEATURES = ["*ALL_INTENSITY*", "*ALL_GLCM*", "*ALL_GLRLM*", "*ALL_GLSZM*", "*ALL_GLDM*", "*ALL_NGTDM*"]
RAD_MASK = sum(ref.FAM[f] for f in RAD_FAMS) | ref.FAM["INT"]


def _engine(features, g=64, **kw):
    from nyxus_b200 import Engine
    e = Engine(list(features), **kw)
    if g != 64:
        e.set_param("coarse_gray_depth", g)      # negative = PyRadiomics bin count (C ABI only; nyxus.py:206 refuses it)
    return e


@pytest.mark.parametrize("name", NAMES)
def test_golden_roi_radiomics_features_and_matrices(name):
    g = int(Z[name + "/g"][0])
    e = _engine(RAD_FEATURES, g)
    try:
        I, L = tu.place_roi(Z[name + "/tile"], Z[name + "/mask"], label=9)
        tab, _ = e.featurize(I, L)
        idx = [COLS.index(c) for c in e.columns]
        bad = tu.mismatches(Z[name + "/features"][idx], tab[0, 2:], e.columns)
        assert not bad, bad[:8]
        for fam in RAD_FAMS:
            got, svals = e.debug_matrix(9, fam)
            want = [tuple(int(x) for x in r) for r in Z[name + "/" + fam]]
            assert sorted(got) == sorted(want), (fam, name)
            if fam == "NGTDM" and len(Z[name + "/NGTDM_S"]):
                assert np.allclose(svals, Z[name + "/NGTDM_S"], rtol=1e-12, atol=1e-12)
    finally:
        e.close()


def test_radiomics_depth_refuses_gldzm_and_ngldm():
    from nyxus_b200.engine import NyxusB200Error
    I, L = tu.place_roi(Z["rad_00/tile"], Z["rad_00/mask"], label=3)
    for fam in ("*ALL_GLDZM*", "*ALL_NGLDM*"):
        e = _engine([fam], -8)
        try:
            with pytest.raises(NyxusB200Error, match="PyRadiomics"):
                e.featurize(I, L)
        finally:
            e.close()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libnyxus_ref.so did not travel")
@pytest.mark.parametrize("seed,g", [(31, -20), (32, -64), (33, -4)])
def test_random_slide_radiomics_against_compiled_reference(seed, g):
    from nyxus_b200 import synth
    I, L = synth.random_rois(200, 260, 40, seed=seed)
    e = _engine(RAD_FEATURES, g)
    try:
        tab, _ = e.featurize(I, L)
        rl, rt, _ = ref.featurize_2d(I.astype(np.uint32), L, greydepth=g, ibsi=False, fam_mask=RAD_MASK, nthreads=8)
        assert np.array_equal(tab[:, 0].astype(np.int64), rl.astype(np.int64))
        idx = [COLS.index(c) for c in e.columns]
        bad = tu.mismatches(rt[:, idx], tab[:, 2:], e.columns)
        assert not bad, bad[:8]
    finally:
        e.close()


def test_glcm_greydepth_feeds_only_the_blank_test():
    """set_metaparam("glcm/greydepth=N") writes GLCM_GREYDEPTH only (env_metaparams.cpp:83-91): the GLCM blank-ROI test uses
    it (glcm.cpp:22-28), the co-occurrence matrix and every other family keep the general grey depth (glcm.cpp:354)"""
    from nyxus_b200 import Engine
    inten, label = Z["gg_inten"], Z["gg_label"]
    for k in range(int(Z["gg_n"][0])):
        g, gg = [int(v) for v in Z["gg%d/params" % k]]
        e = Engine(["*ALL_INTENSITY*", "*ALL_TEXTURE*"], coarse_gray_depth=g)
        try:
            if gg != -2147483648:
                e.set_param("glcm/greydepth", gg)
            tab, _ = e.featurize(inten, label)
            assert np.array_equal(tab[:, 0].astype(np.int64), Z["gg%d/labels" % k].astype(np.int64))
            bad = tu.mismatches(Z["gg%d/table" % k], tab[:, 2:], COLS)
            assert not bad, (k, bad[:8])
            if gg != -2147483648:
                e.set_param("coarse_gray_depth", g)      # re-deriving the settings resets GLCM_GREYDEPTH (env_features.cpp:730)
                tab2, _ = e.featurize(inten, label)
                k0 = [j for j in range(int(Z["gg_n"][0])) if int(Z["gg%d/params" % j][0]) == g and int(Z["gg%d/params" % j][1]) == -2147483648]
                if k0:
                    assert not tu.mismatches(Z["gg%d/table" % k0[0]], tab2[:, 2:], COLS)
        finally:
            e.close()


@pytest.mark.parametrize("k", range(int(Z["v_n"][0])))
def test_golden_volume_signed_depths(k):
    from nyxus_b200 import Engine
    g, gg, gz, gn, rad, which = [int(v) for v in Z["v%d/params" % k]]
    vol = Z["vol_q"] if which else Z["vol"]
    e = Engine(["*3D_ALL_INTENSITY*", "*3D_GLCM*", "*3D_NGTDM*", "*3D_GLSZM*"], coarse_gray_depth=g, dim=3)
    try:
        e.set_param("3glcm/greydepth", gg)
        e.set_param("3glszm/greydepth", gz)
        e.set_param("3ngtdm/greydepth", gn)
        e.set_param("3ngtdm/radius", rad)
        assert e.columns == COLS3
        tab, _ = e.featurize_volume(vol, Z["vol_label"])
        assert np.array_equal(tab[:, 0].astype(np.int64), Z["v%d/labels" % k].astype(np.int64))
        bad = tu.mismatches(Z["v%d/table" % k], tab[:, 2:], COLS3)
        assert not bad, bad[:8]
    finally:
        e.close()
