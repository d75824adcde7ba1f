"""The opt-in multi-valued HISTOGRAM feature on the CPU: the product's intensity_body (host-emulated, test harness only)
against the reference's own analytic fixture (tests/test_2d_intensity_histogram_analytic.h:84-123: {1,1,3,5,7}, 3 bins ->
[2, 1, 2]) and against golden vectors of the reference's own pybind module (tests/golden/golden_hist.npz)."""
import ctypes as C
import os

import numpy as np
import pytest

import nyx_testutil as tu

Z = np.load(os.path.join(tu.GOLDEN, "golden_hist.npz"))
ZI = np.load(os.path.join(tu.GOLDEN, "golden_ih.npz"))


def _emu():
    E = tu.build_emu()
    E.nyxemu_histogram.argtypes = [C.c_void_p, C.c_uint, C.c_int, C.c_void_p]
    return E


def test_reference_analytic_fixture():
    E = _emu()
    vals = np.array([1, 1, 3, 5, 7], np.uint32)
    out = np.zeros(36 + 3)
    assert E.nyxemu_histogram(vals.ctypes.data, 5, 3, out.ctypes.data) == 0
    assert out[36:].tolist() == [2.0, 1.0, 2.0]


@pytest.mark.parametrize("g", [64, 20, 7])
def test_histogram_bins_match_reference_golden(g):
    E = _emu()
    I, L = ZI["inten"].copy(), ZI["label"]
    I[2:6, 2:8] = 0
    cols = [str(c) for c in Z["columns_%d" % g]]
    assert cols == ["ENTROPY", "MEAN"] + ["HISTOGRAM_BIN_%d" % k for k in range(g)]
    for k, lab in enumerate(Z["labels"]):
        vals = np.ascontiguousarray(I[L == lab], np.uint32)
        out = np.zeros(36 + g)
        assert E.nyxemu_histogram(vals.ctypes.data, len(vals), g, out.ctypes.data) == 0
        ref = Z["table_%d" % g][k][2:]
        assert np.array_equal(out[36:], ref), (g, int(lab))
        assert ref.sum() == (0 if lab == Z["blank_label"] else len(vals))       # raw counts; the blank ROI is padded with zeros


def test_histogram_code_sits_between_the_weighted_moments_and_the_ih_family():
    from nyxus_b200 import feature_table as ft
    code = {f[0]: f[1] for f in ft.FEATURES}
    assert code["IMOM_WHU7"] < code["HISTOGRAM"] < code["IH_MEAN_VAL"]
    assert [str(c) for c in Z["mixed_columns"]] == ["MEAN", "GLCM_ASM_0", "GLCM_ASM_45", "GLCM_ASM_90", "GLCM_ASM_135", "IMOM_WHU7"] + \
        ["HISTOGRAM_BIN_%d" % k for k in range(5)] + ["IH_MEAN_VAL"]
