"""*BASIC_MORPHOLOGY*  on the CPU: the product's morph_body (host-emulated, test harness only) against golden
vectors of the reference's own pybind module (tests/golden/golden_morph.npz, generator make_golden_morph.py)."""
import ctypes as C
import os

import numpy as np

import nyx_testutil as tu

Z = np.load(os.path.join(tu.GOLDEN, "golden_morph.npz"))
COLS = [str(c) for c in Z["columns"]]


def test_morphology_body_matches_reference_golden():
    E = tu.build_emu()
    E.nyxemu_morph.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_void_p]
    I, L = np.ascontiguousarray(Z["inten"], np.uint32), np.ascontiguousarray(Z["label"], np.uint32)
    H, W = I.shape
    assert len(COLS) == 15
    for k, lab in enumerate(Z["labels"]):
        out = np.zeros(15)
        assert E.nyxemu_morph(I.ctypes.data, L.ctypes.data, H, W, int(lab), out.ctypes.data) == 0
        bad = tu.mismatches(Z["table"][k], out, COLS)
        assert not bad, (int(lab), bad[:4])


def test_morphology_columns_sit_between_first_order_and_glcm():
    """output order = Feature2D code order (output_2_buffer.cpp:441-578): the reference puts the 15 morphology columns right
    after UNIFORMITY_PIU; the generated feature table must do the same"""
    from nyxus_b200 import feature_table as ft
    names = []
    for name, code, fam, col0, ncols in ft.FEATURES:
        names += [name] if ncols <= 1 else ["%s_%d" % (name, a) for a in (0, 45, 90, 135)]
    from nyxus_b200 import feature_table as ft2
    other = {f[0] for f in ft2.FEATURES if f[2] in ("IH", "GEOM", "WMOM", "HIST")}
    assert [n for n in names if n not in other] == [str(c) for c in Z["all_columns"]]
