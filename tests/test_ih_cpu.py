"""IBSI intensity-histogram family *ALL_IH*  on the CPU: the product's ih_body (host-emulated, test harness
only) against golden vectors of the reference's own pybind module (tests/golden/golden_ih.npz, make_golden_ih.py)."""
import ctypes as C
import os

import numpy as np
import pytest

import nyx_testutil as tu

Z = np.load(os.path.join(tu.GOLDEN, "golden_ih.npz"))
COLS = [str(c) for c in Z["columns"]]


@pytest.mark.parametrize("g", [64, 20, 7])
def test_ih_body_matches_reference_golden(g):
    E = tu.build_emu()
    E.nyxemu_ih.argtypes = [C.c_void_p, C.c_uint, C.c_int, C.c_int, C.c_void_p]
    I, L = Z["inten"], Z["label"]
    assert len(COLS) == 46
    for k, lab in enumerate(Z["labels"]):
        vals = np.ascontiguousarray(I[L == lab], np.uint32)
        out = np.zeros(46)
        assert E.nyxemu_ih(vals.ctypes.data, len(vals), g, 1, out.ctypes.data) == 0
        bad = tu.mismatches(Z["table_%d" % g][k], out, COLS)
        assert not bad, (g, int(lab), bad[:4])


def test_ih_columns_follow_the_texture_columns_and_need_ibsi():
    from nyxus_b200 import feature_table as ft
    names = [f[0] for f in ft.FEATURES if f[2] == "IH"]
    assert names == COLS and min(f[1] for f in ft.FEATURES if f[2] == "IH") > max(f[1] for f in ft.FEATURES if f[2] == "NGTDM")
    assert [str(c) for c in Z["columns_ibsi_off"]] == ["MEAN"]       # IBSI off: the reference strips the family
