"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper around oracle/_ref/libnyxus_ref.so (the UNMODIFIED reference
feature calculators compiled by oracle/Makefile.ref + our C-ABI driver oracle/ref_driver.cpp).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libnyxus_ref.so")

FAM = {"INT": 1, "GLCM": 2, "GLRLM": 4, "GLDZM": 8, "GLSZM": 16, "GLDM": 32, "NGLDM": 64, "NGTDM": 128}
FAM_ALL = 255
FAM_ID2NAME = {v: k for k, v in FAM.items()}

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.nyxref_ncols.restype = C.c_int
        L.nyxref_colname.restype = C.c_char_p
        L.nyxref_colname.argtypes = [C.c_int]
        L.nyxref_colcode.restype = C.c_int
        L.nyxref_colcode.argtypes = [C.c_int]
        L.nyxref_colfamily.restype = C.c_int
        L.nyxref_colfamily.argtypes = [C.c_int]
        L.nyxref_featurize_2d.restype = C.c_int
        L.nyxref_featurize_2d.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_uint, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.nyxref_roi_features.restype = C.c_int
        L.nyxref_roi_features.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_uint, C.c_void_p]
        L.nyxref_roi_matrices.restype = C.c_int64
        L.nyxref_roi_matrices.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_uint, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        L.nyxref_fast_log10.restype = C.c_double
        L.nyxref_fast_log10.argtypes = [C.c_double]
        L.nyxref_bin_pixel.restype = C.c_uint
        L.nyxref_bin_pixel.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_int]
        L.nyxref_to_grayscale.restype = C.c_uint
        L.nyxref_to_grayscale.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_uint]
        _lib = L
    return _lib


def set_glcm_greydepth(g=None):
    """2D calls that follow behave as after set_metaparam("glcm/greydepth=g") (None: default = the grey depth)."""
    lib().nyxref_set_glcm_greydepth(C.c_int(-2147483648 if g is None else int(g)))


def columns():
    L = lib()
    n = L.nyxref_ncols()
    return [L.nyxref_colname(i).decode() for i in range(n)]


def column_codes():
    L = lib()
    return [L.nyxref_colcode(i) for i in range(L.nyxref_ncols())]


def column_families():
    L = lib()
    return [L.nyxref_colfamily(i) for i in range(L.nyxref_ncols())]


def featurize_2d(inten, label, greydepth=64, ibsi=False, glcm_offset=1, fam_mask=FAM_ALL, nthreads=1,
                 force_finite=True, cap=None):
    """Whole-image replay of phase1 -> phase2 -> reduce -> output.  Returns (labels, table[nroi, ncols], timings[11])."""
    L = lib()
    I = np.ascontiguousarray(inten, dtype=np.uint32)
    M = np.ascontiguousarray(label, dtype=np.uint32)
    H, W = I.shape
    nc = L.nyxref_ncols()
    if cap is None:
        cap = int(len(np.unique(M))) + 1
    labels = np.zeros(cap, dtype=np.int32)
    out = np.zeros((cap, nc), dtype=np.float64)
    t = np.zeros(16, dtype=np.float64)
    n = L.nyxref_featurize_2d(I.ctypes.data, M.ctypes.data, H, W, greydepth, int(ibsi), glcm_offset, fam_mask,
                              nthreads, int(force_finite), cap, labels.ctypes.data, out.ctypes.data, t.ctypes.data)
    if n < 0:
        raise RuntimeError("cap too small: need %d" % -n)
    return labels[:n].copy(), out[:n].copy(), t


def roi_features(tile, mask, greydepth=64, ibsi=False, glcm_offset=1, fam_mask=FAM_ALL):
    L = lib()
    T = np.ascontiguousarray(tile, dtype=np.uint32)
    K = np.ascontiguousarray(mask, dtype=np.uint8)
    h, w = T.shape
    out = np.zeros(L.nyxref_ncols(), dtype=np.float64)
    rc = L.nyxref_roi_features(T.ctypes.data, K.ctypes.data, w, h, greydepth, int(ibsi), glcm_offset, fam_mask,
                               out.ctypes.data)
    if rc != 0:
        raise RuntimeError("empty ROI")
    return out


def roi_matrices(tile, mask, greydepth=64, ibsi=False, glcm_offset=1, fam_mask=FAM_ALL & ~1):
    """Sparse matrices of one ROI: dict family -> sorted list of (angle, a, b, count); 'NGTDM_S' -> list of S values."""
    L = lib()
    T = np.ascontiguousarray(tile, dtype=np.uint32)
    K = np.ascontiguousarray(mask, dtype=np.uint8)
    h, w = T.shape
    cap = 1 << 16
    while True:
        trip = np.zeros((cap, 5), dtype=np.int64)
        s = np.zeros(70000, dtype=np.float64)
        n = L.nyxref_roi_matrices(T.ctypes.data, K.ctypes.data, w, h, greydepth, int(ibsi), glcm_offset, fam_mask,
                                  cap, trip.ctypes.data, s.ctypes.data, len(s))
        if n >= 0:
            break
        cap = -n + 16
    res = {k: [] for k in FAM if k != "INT"}
    ks = 0
    S = []
    for f, ang, a, b, c in trip[:n]:
        res[FAM_ID2NAME[int(f)]].append((int(ang), int(a), int(b), int(c)))
        if int(f) == FAM["NGTDM"]:
            S.append(float(s[ks]))
            ks += 1
    res["NGTDM_S"] = S
    return res


# ---------------------------------------------------------------------------------------------- 3D (Nyxus3D path)
FAM3 = {"INT": 1, "GLCM": 2, "NGTDM": 4, "GLSZM": 8}


def columns_3d():
    L = lib()
    L.nyxref3d_colname.restype = C.c_char_p
    return [L.nyxref3d_colname(i).decode() for i in range(L.nyxref3d_ncols())]


def featurize_3d(inten, label, greydepth=64, glcm_greydepth=None, glszm_greydepth=0, ngtdm_greydepth=0, ngtdm_radius=0,
                 ibsi=False, fam_mask=15, nthreads=1, force_finite=True):
    """inten, label: [D,H,W].  Returns (labels, table[n, 116], phase seconds).  Defaults follow env_features.cpp:712-736
    (only GLCM gets a grey depth; 3D GLSZM / NGTDM depths and the NGTDM radius are 0 unless set)."""
    L = lib()
    I = np.ascontiguousarray(inten, dtype=np.uint32)
    M = np.ascontiguousarray(label, dtype=np.uint32)
    D, H, W = I.shape
    nc = L.nyxref3d_ncols()
    cap = int(len(np.unique(M))) + 1
    labels = np.zeros(cap, np.int32)
    out = np.zeros((cap, nc), np.float64)
    tph = np.zeros(8, np.float64)
    L.nyxref_featurize_3d.restype = C.c_int
    L.nyxref_featurize_3d.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 9 + [C.c_uint, C.c_int, C.c_int, C.c_int,
                                                                                  C.c_void_p, C.c_void_p, C.c_void_p]
    n = L.nyxref_featurize_3d(I.ctypes.data, M.ctypes.data, D, H, W, int(greydepth),
                              int(greydepth if glcm_greydepth is None else glcm_greydepth), int(glszm_greydepth),
                              int(ngtdm_greydepth), int(ngtdm_radius), int(bool(ibsi)), int(fam_mask), int(nthreads),
                              int(bool(force_finite)), cap, labels.ctypes.data, out.ctypes.data, tph.ctypes.data)
    assert n >= 0, n
    return labels[:n].copy(), out[:n].copy(), tph
