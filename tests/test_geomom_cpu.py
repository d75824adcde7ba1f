"""Unweighted 2D geometric moments on the CPU: the product's geomom_body (host-emulated, test harness only)
against golden vectors of the reference's own pybind module (tests/golden/golden_geomom.npz, make_golden_geomom.py)."""
import ctypes as C
import os

import numpy as np

import nyx_testutil as tu

Z = np.load(os.path.join(tu.GOLDEN, "golden_geomom.npz"))
COLS = [str(c) for c in Z["columns"]]


def test_geomom_body_matches_reference_golden():
    from nyxus_b200 import feature_table as ft
    canon = {f[0]: f[3] for f in ft.FEATURES if f[2] == "GEOM"}
    base = min(canon.values())
    assert len(COLS) == 118 and sorted(canon) == sorted(COLS)
    assert COLS == [f[0] for f in ft.FEATURES if f[2] == "GEOM"]         # output order = Feature2D code order
    E = tu.build_emu()
    E.nyxemu_geomom.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_void_p]
    I, L = np.ascontiguousarray(Z["inten"], np.uint32), np.ascontiguousarray(Z["label"], np.uint32)
    H, W = I.shape
    for k, lab in enumerate(Z["labels"]):
        out = np.zeros(118)
        assert E.nyxemu_geomom(I.ctypes.data, L.ctypes.data, H, W, int(lab), out.ctypes.data) == 0
        got = np.array([out[canon[c] - base] for c in COLS])
        bad = tu.mismatches(Z["table"][k], got, COLS)
        assert not bad, (int(lab), bad[:4])
