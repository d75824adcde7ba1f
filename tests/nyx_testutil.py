"""Shared helpers of the test-suite."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")
FAMILIES = ["GLCM", "GLRLM", "GLDZM", "GLSZM", "GLDM", "NGLDM", "NGTDM"]

RTOL = 1e-9      # north_star: derived feature values within 1e-9 relative (FP64)
ATOL = 1e-12     # absolute floor for values that are mathematically 0 / cancellation residue


def load_golden_rois():
    z = np.load(os.path.join(GOLDEN, "golden_rois.npz"))
    cols = [str(c) for c in z["columns"]]
    cases = []
    for name in [str(n) for n in z["names"]]:
        g, ibsi = [int(v) for v in z[name + "/settings"]]
        mats = {f: [tuple(int(x) for x in r) for r in z[name + "/" + f]] for f in FAMILIES}
        cases.append(dict(name=name, tile=z[name + "/tile"], mask=z[name + "/mask"], g=g, ibsi=bool(ibsi),
                          features=z[name + "/features"], mats=mats, ngtdm_s=z[name + "/NGTDM_S"]))
    return cols, cases


def load_golden_slide():
    z = np.load(os.path.join(GOLDEN, "golden_slide.npz"))
    return {k: z[k] for k in z.files}


def finite0(a):
    a = np.asarray(a, np.float64)
    return np.where(np.isfinite(a), a, 0.0)


def mismatches(ref, got, cols=None, rtol=RTOL, atol=ATOL):
    """list of (index / column, ref, got) entries outside |a-b| <= atol + rtol*max(|a|,|b|)"""
    a, b = finite0(ref), finite0(got)
    diff = np.abs(a - b)
    bad = diff > atol + rtol * np.maximum(np.abs(a), np.abs(b))
    out = []
    for idx in zip(*np.nonzero(bad)):
        name = cols[idx[-1]] if cols is not None else idx
        out.append((idx, name, float(a[idx]), float(b[idx])))
    return out


def build_emu():
    """host emulation of the per-ROI feature bodies (TEST HARNESS, see tests/emu/emu.cpp)"""
    emu_dir = os.path.join(HERE, "emu")
    so = os.path.join(emu_dir, "libnyx_emu.so")
    src = os.path.join(emu_dir, "emu.cpp")
    csrc = os.path.join(ROOT, "nyxus_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src])
    L = C.CDLL(so)
    L.nyxemu_roi_features.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_void_p, C.c_int]
    L.nyxemu_fast_log10.restype = C.c_double
    L.nyxemu_fast_log10.argtypes = [C.c_double]
    L.nyxemu_bin.restype = C.c_uint
    L.nyxemu_bin.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_int]
    L.nyxemu_to_grayscale.restype = C.c_uint
    L.nyxemu_to_grayscale.argtypes = [C.c_uint] * 4
    L.nyxemu_roi_features_3d.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_double, C.c_uint, C.c_void_p]
    return L


def emu_roi_features(L, tile, mask, g=64, ibsi=False, off=1, fam=255, p16=1):
    T = np.ascontiguousarray(tile, np.uint32)
    K = np.ascontiguousarray(mask, np.uint8)
    out = np.zeros(337)
    rc = L.nyxemu_roi_features(T.ctypes.data, K.ctypes.data, T.shape[1], T.shape[0], g, int(ibsi), off, fam, out.ctypes.data, p16)
    assert rc == 0
    return out


def place_roi(tile, mask, H=None, W=None, y=3, x=5, label=1):
    """embed one ROI tile into a slide -> (inten uint32 [H,W], label uint32 [H,W])"""
    h, w = tile.shape
    H = H or h + 8
    W = W or w + 12
    I = np.full((H, W), 1234, np.uint32)
    L = np.zeros((H, W), np.uint32)
    I[y:y + h, x:x + w] = tile
    L[y:y + h, x:x + w] = np.where(mask != 0, label, 0)
    return I, L


# ---------------------------------------------------------------------------------------------- 3D helpers
G3 = ["ASM", "ACOR", "CLUPROM", "CLUSHADE", "CLUTEND", "CONTRAST", "CORRELATION", "DIFAVE", "DIFENTRO", "DIFVAR", "DIS", "ENERGY",
      "ENTROPY", "HOM1", "HOM2", "ID", "IDN", "IDM", "IDMN", "INFOMEAS1", "INFOMEAS2", "IV", "JAVE", "JE", "JMAX", "JVAR",
      "SUMAVERAGE", "SUMENTROPY", "SUMVARIANCE", "VARIANCE"]


def load_golden_vol():
    z = np.load(os.path.join(GOLDEN, "golden_vol.npz"))
    d = {k: z[k] for k in z.files}
    d["columns"] = [str(c) for c in d["columns"]]
    return d


def canonical3d_index(cols):
    """position of every reference 3D column in the library's canonical 3D table (36 | 30x13+29 | 5 | 16)"""
    idx, nave, pi, pn, pz = [], 0, 0, 0, 0
    for c in cols:
        if c.startswith("3GLCM_"):
            s = c[6:]
            if s.endswith("_AVE"):
                idx.append(36 + 390 + nave)
                nave += 1
            else:
                idx.append(36 + G3.index(s) * 13)
        elif c.startswith("3NGTDM_"):
            idx.append(36 + 419 + pn)
            pn += 1
        elif c.startswith("3GLSZM_"):
            idx.append(36 + 419 + 5 + pz)
            pz += 1
        else:
            idx.append(pi)
            pi += 1
    return np.array(idx)


def emu_roi_features_3d(L, vol, lbl, lab, g, gz, gn, rad, ibsi, cols):
    """host-emulated 3D bodies for the ROI `lab` of a labelled volume, in the reference's column order"""
    V = np.ascontiguousarray(vol, np.uint32)
    M = np.ascontiguousarray(lbl, np.uint32)
    fg = V[M > 0]
    den = float(fg.max()) - float(fg.min())
    out = np.zeros(36 + 419 + 5 + 16)
    D, H, W = V.shape
    rc = L.nyxemu_roi_features_3d(V.ctypes.data, M.ctypes.data, D, H, W, int(lab), int(g), int(gz), int(gn), int(rad), int(bool(ibsi)),
                                  den, 15, out.ctypes.data)
    assert rc == 0
    return finite0(out[canonical3d_index(cols)])
