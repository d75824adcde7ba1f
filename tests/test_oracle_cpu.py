"""CPU tests (no GPU): the oracle is pinned against the reference's own known answers and against golden vectors
produced by the compiled reference; the per-ROI feature bodies of the product (host-emulated, test harness only)
are checked against the same goldens."""
import numpy as np
import pytest

import nyx_testutil as tu
import reference_fixtures as rf
from oracle import nyx_oracle as no
from oracle import ref

COLS, CASES = tu.load_golden_rois()


def _oracle_vector(case, families=None):
    kw = {} if families is None else {"families": families}
    cols, mats = no.roi_features(case["tile"], case["mask"], case["g"], case["ibsi"], **kw)
    return cols, mats


# ---------------------------------------------------------------- reference known-answer fixtures
def test_2d_glcm_ibsi_consensus_values():
    """reference tests/test_2d_glcm_ibsi.h: mean over 4 phantom slices x 4 angles within 1 % of the IBSI consensus.
    (The reference test sets GREYDEPTH = GLCM_GREYDEPTH = 0 + IBSI; with IBSI on the grey depth is ignored except in the
    blank-ROI test, which none of the slices triggers.)"""
    tot = {k: 0.0 for k in rf.GLCM_IBSI_CONSENSUS}
    for z in range(4):
        roi = no.roi_from_tile(rf.IBSI_PHANTOM_INT[z], rf.IBSI_PHANTOM_MASK[z])
        out, _ = no.glcm(roi, 64, True)
        for k in tot:
            tot[k] += sum(out[k])
    for k, golden in rf.GLCM_IBSI_CONSENSUS.items():
        assert abs(tot[k] / 16.0 - golden) <= 0.01 * abs(golden) + 5e-3, (k, tot[k] / 16.0, golden)


def test_2d_ngldm_matrix_ibsi():
    """reference tests/test_2d_ngldm_ibsi.h:58-66 -- every cell of the IBSI fig 3.19 matrix"""
    _, sp = no.ngldm(no.roi_from_tile(rf.NGLDM_IBSI_IMAGE, np.ones((4, 4))), 64, True)
    M = np.zeros((4, 4), int)
    for (lv, j), c in sp.items():
        M[lv - 1, j] = c
    assert np.array_equal(M, rf.NGLDM_IBSI_MATRIX)


def test_2d_ngldm_matrix_nonibsi():
    """reference tests/test_2d_ngldm_regression.h:73-81"""
    _, sp = no.ngldm(no.roi_from_tile(rf.NGLDM_NONIBSI_IMAGE, np.ones((5, 5))), 64, False)
    levels = sorted({k[0] for k in sp})
    M = np.zeros((len(levels), 6), int)
    for (lv, j), c in sp.items():
        M[levels.index(lv), j] = c
    assert np.array_equal(M, rf.NGLDM_NONIBSI_MATRIX)


def test_2d_gldzm_matrix_ibsi():
    """reference tests/test_2d_gldzm_ibsi.h:66-74 -- every cell of the IBSI fig 3.17 matrix"""
    _, sp = no.gldzm(no.roi_from_tile(rf.GLDZM_IBSI_IMAGE, np.ones((4, 4))), 64, True)
    M = np.zeros((4, 2), int)
    for (lv, d), c in sp.items():
        M[lv - 1, d - 1] = c
    assert np.array_equal(M, rf.GLDZM_IBSI_MATRIX)


# ---------------------------------------------------------------- golden vectors of the compiled reference
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference_golden(case):
    cols, mats = _oracle_vector(case)
    got = np.array([cols[c] for c in COLS])
    bad = tu.mismatches(case["features"], got, COLS)
    assert not bad, bad[:8]
    for fam in tu.FAMILIES:
        assert sorted(mats[fam]) == sorted(case["mats"][fam]), fam
    if len(case["ngtdm_s"]):
        assert np.allclose(mats["NGTDM_S"], case["ngtdm_s"], rtol=1e-12, atol=1e-12)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("case", CASES[::3], ids=[c["name"] for c in CASES[::3]])
def test_compiled_reference_reproduces_golden(case):
    got = ref.roi_features(case["tile"], case["mask"], case["g"], case["ibsi"])
    assert np.array_equal(tu.finite0(got), tu.finite0(case["features"]))
    m = ref.roi_matrices(case["tile"] * case["mask"], case["mask"], case["g"], case["ibsi"])
    for fam in tu.FAMILIES:
        assert sorted(m[fam]) == sorted(case["mats"][fam])


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_column_layout_is_the_generated_table():
    from nyxus_b200 import feature_table as ft
    assert ft.CANONICAL_COLUMNS[:337] == ref.columns()        # (the appended *BASIC_MORPHOLOGY* columns: tests/test_morph_cpu.py)
    assert len(ft.CANONICAL_COLUMNS) == 337 + 15 + 46 + 118 + 62


def test_golden_slide_matches_oracle():
    s = tu.load_golden_slide()
    cols = [str(c) for c in s["columns"]]
    for i, lab in enumerate(s["labels"][:4]):
        m = s["label"] == lab
        oc, _ = no.roi_features(np.where(m, s["inten"], 0), m, 64, False)
        got = np.array([no.finite(oc[c]) for c in cols])
        bad = tu.mismatches(s["table"][i], got, cols)
        assert not bad, bad[:5]


# ---------------------------------------------------------------- scalar bit-exactness
def test_fast_log10_and_binning_bit_exact_against_oracle():
    L = tu.build_emu()
    rng = np.random.default_rng(5)
    xs = np.concatenate([10.0 ** rng.uniform(-17, 1, 4000), rng.random(2000), [1e-9, 2.2e-16, 1.0, 0.75, 1.5]])
    for x in xs:
        assert L.nyxemu_fast_log10(float(x)) == no.fast_log10(float(x))
    for _ in range(4000):
        mx = int(rng.integers(1, 70000))
        x = int(rng.integers(0, mx + 1))
        g = int(rng.choice([8, 64, 100, 256]))
        assert L.nyxemu_bin(x, 0, mx, g) == no.bin_pixel_matlab(x, mx, g)
        mn = int(rng.integers(0, mx))
        x2 = int(rng.integers(mn, mx + 1))
        assert L.nyxemu_to_grayscale(x2, mn, mx - mn, g) == no.to_grayscale(x2, mn, mx - mn, g)


# ---------------------------------------------------------------- product feature bodies (host-emulated) vs golden
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_feature_bodies_match_reference_golden(case):
    L = tu.build_emu()
    for p16 in (0, 1):
        got = tu.emu_roi_features(L, case["tile"], case["mask"], case["g"], case["ibsi"], p16=p16)
        bad = tu.mismatches(case["features"], got, COLS)
        assert not bad, bad[:8]


def test_feature_bodies_randomised_against_oracle():
    L = tu.build_emu()
    rng = np.random.default_rng(99)
    for it in range(12):
        h, w = int(rng.integers(2, 12)), int(rng.integers(2, 12))
        t = rng.integers(0, [5000, 6, 300][it % 3], (h, w)).astype(np.uint32)
        m = (rng.random((h, w)) < 0.85).astype(np.uint8)
        m[0, 0] = 1
        ibsi = it % 3 == 1
        oc, _ = no.roi_features(t, m, 64, ibsi)
        ref_v = np.array([oc[c] for c in COLS])
        got = tu.emu_roi_features(L, t, m, 64, ibsi)
        bad = tu.mismatches(ref_v, got, COLS)
        assert not bad, bad[:8]
