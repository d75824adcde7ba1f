"""CPU tests of the settings the reference's own tests drive through Fsettings / set_metaparam and round 1 left out:
PyRadiomics binning (negative grey depth, texture_feature.h:106-118) and a GLCM grey depth that differs from the general
one (env_metaparams.cpp:83-91, glcm.cpp:22-28,354).  The product's feature bodies run host-emulated (tests/emu, test
harness only) against golden vectors of the compiled reference (tests/golden/golden_radiomics.npz, generator
make_golden_rad.py) and, where oracle/_ref is present, against the compiled reference itself on fresh inputs."""
import ctypes as C
import os

import numpy as np
import pytest

import nyx_testutil as tu
from oracle import nyx_oracle as no
from oracle import ref

Z = np.load(os.path.join(tu.GOLDEN, "golden_radiomics.npz"))
COLS = [str(c) for c in Z["columns"]]
COLS3 = [str(c) for c in Z["columns3d"]]
NAMES = [str(n) for n in Z["names"]]
RAD_MASK = 1 | 2 | 4 | 16 | 32 | 128          # INT GLCM GLRLM GLSZM GLDM NGTDM (emu / ref family bits)


def _emu():
    L = tu.build_emu()
    L.nyxemu_set_glcm_greydepth.argtypes = [C.c_int]
    L.nyxemu_set_glcm3_greydepth.argtypes = [C.c_int]
    return L


def test_radiomics_bin_bit_exact():
    """Binner (nyx_common.h) == to_grayscale_radiomix of the numpy oracle (and of the compiled reference when present)"""
    L = _emu()
    rng = np.random.default_rng(11)
    for _ in range(6000):
        mx = int(rng.integers(1, 70000))
        mn = int(rng.integers(0, mx + 1))
        x = int(rng.integers(mn, mx + 1)) if rng.random() < 0.9 else 0
        n = int(rng.choice([1, 2, 4, 8, 20, 64, 100, 256, 1000]))
        want = no.bin_pixel_radiomics(x, mn, mx, n)
        assert L.nyxemu_bin(x, mn, mx, -n) == want, (x, mn, mx, n)
        if ref.available():
            assert ref.lib().nyxref_bin_pixel(x, mn, mx, -n) == want


@pytest.mark.parametrize("name", NAMES)
def test_feature_bodies_match_reference_golden_radiomics(name):
    L = _emu()
    g = int(Z[name + "/g"][0])
    for p16 in (0, 1):
        got = tu.emu_roi_features(L, Z[name + "/tile"], Z[name + "/mask"], g, False, fam=RAD_MASK, p16=p16)
        bad = tu.mismatches(Z[name + "/features"], got, COLS)
        assert not bad, bad[:8]


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_feature_bodies_randomised_radiomics_against_compiled_reference():
    L = _emu()
    rng = np.random.default_rng(5)
    for it in range(40):
        h, w = int(rng.integers(1, 24)), int(rng.integers(1, 24))
        t = rng.integers(0, [5000, 6, 300, 65000][it % 4], (h, w)).astype(np.uint32)
        m = (rng.random((h, w)) < 0.85).astype(np.uint8)
        m[0, 0] = 1
        g = -int(rng.choice([1, 4, 8, 20, 64, 256]))
        want = ref.roi_features(t, m, g, False, fam_mask=RAD_MASK)
        got = tu.emu_roi_features(L, t, m, g, False, fam=RAD_MASK, p16=it % 2)
        bad = tu.mismatches(want, got, COLS)
        assert not bad, (it, g, bad[:6])


def test_glcm_greydepth_feeds_only_the_blank_test():
    """golden slides with "glcm/greydepth" != grey depth: per ROI the host-emulated bodies reproduce the reference row; the
    non-GLCM columns never move, the GLCM block of the ROIs whose blank test flips does"""
    L = _emu()
    inten, label = Z["gg_inten"], Z["gg_label"]
    base = Z["gg0/table"]
    glcm_cols = np.array([c.startswith("GLCM_") for c in COLS])
    flipped = 0
    try:
        for k in range(int(Z["gg_n"][0])):
            g, gg = [int(v) for v in Z["gg%d/params" % k]]
            tab = Z["gg%d/table" % k]
            if g == 64:
                assert np.array_equal(tab[:, ~glcm_cols], base[:, ~glcm_cols])
                flipped += int((tab[:, glcm_cols] != base[:, glcm_cols]).any(axis=1).sum())
            L.nyxemu_set_glcm_greydepth(-2147483648 if gg == -2147483648 else gg)
            for row, lab in enumerate(Z["gg%d/labels" % k]):
                m = label == lab
                ys, xs = np.nonzero(m)
                sl = (slice(ys.min(), ys.max() + 1), slice(xs.min(), xs.max() + 1))
                got = tu.emu_roi_features(L, inten[sl], m[sl], g, False)
                bad = tu.mismatches(tab[row], got, COLS)
                assert not bad, (k, int(lab), bad[:6])
    finally:
        L.nyxemu_set_glcm_greydepth(-2147483648)
    assert flipped > 0


@pytest.mark.parametrize("k", range(int(Z["v_n"][0])))
def test_3d_bodies_match_reference_golden_signed_depths(k):
    L = _emu()
    g, gg, gz, gn, rad, which = [int(v) for v in Z["v%d/params" % k]]
    vol = Z["vol_q"] if which else Z["vol"]
    try:
        L.nyxemu_set_glcm3_greydepth(gg)
        for row, lab in enumerate(Z["v%d/labels" % k]):
            got = tu.emu_roi_features_3d(L, vol, Z["vol_label"], lab, g, gz, gn, rad, False, COLS3)
            bad = tu.mismatches(Z["v%d/table" % k][row], got, COLS3)
            assert not bad, (k, int(lab), bad[:6])
    finally:
        L.nyxemu_set_glcm3_greydepth(-2147483648)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_3d_bodies_randomised_signed_depths_against_compiled_reference():
    L = _emu()
    rng = np.random.default_rng(78)
    try:
        for it in range(30):
            D, H, W = [int(rng.integers(2, 8)) for _ in range(3)]
            kind = it % 3
            V = rng.integers(1, 5000, (D, H, W)) if kind == 0 else rng.integers(1, 6, (D, H, W)) if kind == 1 else rng.integers(0, 200, (D, H, W))
            M = (rng.random((D, H, W)) < rng.choice([0.6, 0.9, 1.0])).astype(np.uint32) * 7
            M[0, 0, 0] = 7
            gg, gz, gn = [int(rng.choice([-4, -8, -20, -64, 16, 0])) for _ in range(3)]
            rad = int(rng.choice([0, 1, 1]))
            _, tab, _ = ref.featurize_3d(V.astype(np.uint32), M, greydepth=64, glcm_greydepth=gg, glszm_greydepth=gz,
                                         ngtdm_greydepth=gn, ngtdm_radius=rad, ibsi=False)
            L.nyxemu_set_glcm3_greydepth(gg)
            got = tu.emu_roi_features_3d(L, V, M, 7, 64, gz, gn, rad, False, COLS3)
            bad = tu.mismatches(tab[0], got, COLS3)
            assert not bad, (it, (gg, gz, gn, rad), bad[:6])
    finally:
        L.nyxemu_set_glcm3_greydepth(-2147483648)
