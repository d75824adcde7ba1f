"""Distance-to-contour weighted 2D geometric moments (second half) on the CPU: the product's wmom_body -- the
restated contour tracer, the reference's approximate contour distance, float weights, two moment passes -- host-emulated (test
harness only) against golden vectors of the reference's own pybind module (tests/golden/golden_wmom.npz, make_golden_wmom.py:
concave blobs, a ring, one-pixel-wide bars, a plus sign, speckles, two-part ROIs, a ROI at the slide corner)."""
import ctypes as C
import os

import numpy as np

import nyx_testutil as tu

Z = np.load(os.path.join(tu.GOLDEN, "golden_wmom.npz"))
COLS = [str(c) for c in Z["columns"]]
PQ10 = ["00", "01", "02", "03", "10", "11", "12", "20", "21", "30"]
PQ7 = ["02", "03", "11", "12", "20", "21", "30"]
CANON = (["WEIGHTED_SPAT_MOMENT_" + s for s in PQ10] + ["WEIGHTED_CENTRAL_MOMENT_" + s for s in PQ7] + ["WT_NORM_CTR_MOM_" + s for s in PQ7]
         + ["WEIGHTED_HU_M%d" % k for k in range(1, 8)]
         + ["IMOM_WRM_" + s for s in PQ10] + ["IMOM_WCM_" + s for s in PQ7] + ["IMOM_WNCM_" + s for s in PQ7] + ["IMOM_WHU%d" % k for k in range(1, 8)])


def test_wmom_body_matches_reference_golden():
    E = tu.build_emu()
    E.nyxemu_wmom.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
    I, L = tu.wmom_slide()
    I, L = np.ascontiguousarray(I, np.uint32), np.ascontiguousarray(L, np.uint32)
    H, W = I.shape
    sel = [COLS.index(c) for c in CANON]
    worst = []
    for k, lab in enumerate(Z["labels"]):
        out = np.zeros(62)
        assert E.nyxemu_wmom(I.ctypes.data, L.ctypes.data, H, W, int(lab), out.ctypes.data, None, None) == 0
        bad = tu.mismatches(tu.finite0(Z["table"][k][sel]), tu.finite0(out), CANON)
        if bad:
            worst.append((int(lab), bad[:3]))
    assert not worst, worst[:6]


def test_wmom_body_on_a_large_ragged_roi_matches_reference_golden():
    """golden_wmom_big.npz: a ROI of ~ 60 000 bounding-box cells with holes, a slit and pinholes (many inner contours, failed
    loops, long ordered contour for the hill descent), a thin diagonal band, a small blob"""
    Zb = np.load(os.path.join(tu.GOLDEN, "golden_wmom_big.npz"))
    cols = [str(c) for c in Zb["columns"]]
    E = tu.build_emu()
    E.nyxemu_wmom.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
    I, L = tu.wmom_slide_big()
    I, L = np.ascontiguousarray(I, np.uint32), np.ascontiguousarray(L, np.uint32)
    H, W = I.shape
    sel = [cols.index(c) for c in CANON]
    for k, lab in enumerate(Zb["labels"]):
        out = np.zeros(62)
        assert E.nyxemu_wmom(I.ctypes.data, L.ctypes.data, H, W, int(lab), out.ctypes.data, None, None) == 0
        bad = tu.mismatches(tu.finite0(Zb["table"][k][sel]), tu.finite0(out), CANON)
        assert not bad, (int(lab), bad[:4])
