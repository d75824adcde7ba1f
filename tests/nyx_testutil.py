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
    """list of (index / column, ref, got) entries outside |a-b| <= atol + rtol*max(|a|,|b|).

    One documented exception: GLCM_CLUSHADE* = sum (i+j-mx-my)^3 p is a signed sum that cancels to (near) zero on symmetric
    matrices -- the reference then returns round-off residue of its own summation order (seen: 3e-11 where the exact
    value is 0).  Its terms are bounded by CLUPROM^(3/4) (power-mean inequality), so for those columns the relative
    tolerance is taken against max(|a|, |b|, CLUPROM^(3/4)) of the same angle: 1e-9 of the magnitude being summed.
    Likewise GLCM_INFOMEAS2* = sqrt(1 - exp(-2 x)), x = HXY2 - HXY (glcm.cpp:888-913): x is the difference of two
    entropies of magnitude H, each a float32-log sum good to 1e-9 relative at best, and df = dx * exp(-2x) / f, so a 1e-9
    relative agreement of the entropies allows |df| <= 1e-9 * 2H / f (ill-conditioned when the matrix is nearly
    independent, f -> 0); the tolerance scale of those columns is max(|a|, |b|, min(2 H / f, 1e4))."""
    a, b = finite0(ref), finite0(got)
    diff = np.abs(a - b)
    scale = np.maximum(np.abs(a), np.abs(b))
    if cols is not None and a.ndim >= 1 and a.shape[-1] == len(cols):
        names = list(cols)
        for j, c in enumerate(names):
            if c.startswith("GLCM_CLUSHADE"):
                prom = c.replace("CLUSHADE", "CLUPROM")
                if prom in names:
                    k = names.index(prom)
                    scale[..., j] = np.maximum(scale[..., j], np.maximum(np.abs(a[..., k]), np.abs(b[..., k])) ** 0.75)
            if c.startswith("GLCM_INFOMEAS2"):
                ent = c.replace("INFOMEAS2", "ENTROPY")
                if ent in names:
                    k = names.index(ent)
                    H = np.maximum(np.abs(a[..., k]), np.abs(b[..., k]))
                    f = np.maximum(np.maximum(np.abs(a[..., j]), np.abs(b[..., j])), 1e-30)
                    scale[..., j] = np.maximum(scale[..., j], np.minimum(2.0 * H / f, 1e4))
            # first-order CENTRAL moments (CENTRAL_MOMENT_01 / _10, IMOM_CM_01 / _10) are identically zero: sum I (x - ox) with
            # ox = m10 / m00; what either side returns is the round-off residue of sums whose terms have the magnitude
            # sqrt (mu20 * mu00) resp. sqrt (mu02 * mu00) (Cauchy-Schwarz) -- the tolerance is relative to that
            for first, second in (("_01", "_02"), ("_10", "_20")):
                if (c.startswith("CENTRAL_MOMENT") or c.startswith("IMOM_CM")) and c.endswith(first):
                    c2, c0 = c[:-3] + second, c[:-3] + "_00"
                    if c2 in names and c0 in names:
                        k2, k0 = names.index(c2), names.index(c0)
                        scale[..., j] = np.maximum(scale[..., j], np.sqrt(np.abs(a[..., k2] * a[..., k0])))
            # Hu's invariants 5-7 (2d_geomoments_basic.cpp:240-254) are signed sums of degree-4 (5, 7) and degree-3 (6) products
            # of the normalised moments: with a = n30 + n12, b = n21 + n03, e = n30 - 3 n12, f = 3 n21 - n03 every term of
            # HU5 / HU7 is bounded by sqrt (HU3 HU4) HU4 (HU3 = e^2 + f^2, HU4 = a^2 + b^2) and every term of HU6 by
            # sqrt (HU2) HU4; the values themselves can cancel by orders of magnitude (seen: 1e33 out of terms of 4e38 for a
            # weighted mass near zero) -- the tolerance is relative to the magnitude being summed
            for pre in ("HU_M", "IMOM_HU", "WEIGHTED_HU_M", "IMOM_WHU"):
                if c.startswith(pre) and c[len(pre):] in ("5", "6", "7") and all(pre + d in names for d in "234"):
                    h2, h3, h4 = (np.maximum(np.abs(a[..., names.index(pre + d)]), np.abs(b[..., names.index(pre + d)])) for d in "234")
                    bound = 2.0 * np.sqrt(h2) * h4 if c.endswith("6") else 4.0 * np.sqrt(h3 * h4) * h4
                    scale[..., j] = np.maximum(scale[..., j], bound)
    bad = diff > atol + rtol * scale
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
    """position of every reference 3D column in the library's canonical 3D table (36 | 30x13+29 | 5 | 16 | 14 | 32 | 18 | 19)"""
    idx, nave, pi = [], 0, 0
    base = {"3NGTDM_": 36 + 419, "3GLSZM_": 36 + 419 + 5, "3GLDM_": 36 + 419 + 5 + 16, "3GLRLM_": 36 + 419 + 5 + 16 + 14,
            "3GLDZM_": 36 + 419 + 5 + 16 + 14 + 32, "3NGLDM_": 36 + 419 + 5 + 16 + 14 + 32 + 18}
    pos = dict.fromkeys(base, 0)
    for c in cols:
        fam = next((f for f in base if c.startswith(f)), None)
        if c.startswith("3GLCM_"):
            s = c[6:]
            if s.endswith("_AVE"):
                idx.append(36 + 390 + nave)
                nave += 1
            else:
                idx.append(36 + G3.index(s) * 13)
        elif fam == "3GLRLM_":                                       # 16 direction-0 values, then the 16 averages
            ave = c.endswith("_AVE")
            names = [x for x in cols if x.startswith("3GLRLM_") and x.endswith("_AVE") == ave]
            idx.append(base[fam] + (16 if ave else 0) + names.index(c))
        elif fam:
            idx.append(base[fam] + pos[fam])
            pos[fam] += 1
        else:
            idx.append(pi)
            pi += 1
    return np.array(idx)


def emu_roi_features_3d(L, vol, lbl, lab, g, gz, gn, rad, ibsi, cols, fam_mask=15):
    """host-emulated 3D bodies for the ROI `lab` of a labelled volume, in the reference's column order"""
    V = np.ascontiguousarray(vol, np.uint32)
    M = np.ascontiguousarray(lbl, np.uint32)
    fg = V[M > 0]
    den = float(fg.max()) - float(fg.min())
    out = np.zeros(36 + 419 + 5 + 16 + 14 + 32 + 18 + 19)
    D, H, W = V.shape
    rc = L.nyxemu_roi_features_3d(V.ctypes.data, M.ctypes.data, D, H, W, int(lab), int(g), int(gz), int(gn), int(rad), int(bool(ibsi)),
                                  den, fam_mask, out.ctypes.data)
    assert rc == 0
    return finite0(out[canonical3d_index(cols)])


# ---------------------------------------------------------------------------------------------- large-ROI (giant path) cases
def _smooth_field(rng, H, W, lo, hi, passes=3):
    f = rng.random((H, W))
    for _ in range(passes):
        f = (f + np.roll(f, 1, 0) + np.roll(f, 1, 1) + np.roll(f, -1, 0) + np.roll(f, -1, 1)) / 5.0
    f = (f - f.min()) / (f.max() - f.min())
    return (lo + f * (hi - lo)).astype(np.uint32)


def giant_cases():
    """(name, intensity, label, settings) -- every case holds at least one ROI whose AABB exceeds the 40000-cell shared-memory
    classes, so it runs through giant.cuh.  Sizes stay small enough for the compiled reference to finish in seconds."""
    rng = np.random.default_rng(77)
    out = []
    # 1. whole-image single ROI (label == 1 everywhere), MATLAB binning, >= 65536 cells (GLSZM_SAE int-wrap NaN case)
    I = _smooth_field(rng, 260, 270, 500, 9000, 2) + rng.integers(0, 300, (260, 270)).astype(np.uint32)
    out.append(("whole_260x270", I.astype(np.uint16), np.ones((260, 270), np.uint32), dict(g=64)))
    # 2. concave blob with a hole + small ROIs around it; background inside the AABB becomes level-1 zones/runs
    H, W = 330, 380
    yy, xx = np.mgrid[0:H, 0:W]
    big = (((yy - 160) / 130.0) ** 2 + ((xx - 180) / 150.0) ** 2 < 1) & ~((((yy - 150) / 40.0) ** 2 + ((xx - 200) / 60.0) ** 2) < 1)
    big &= ~((yy > 200) & (xx > 250) & (yy - 200 > (xx - 250) * 0.3))
    L = np.where(big, 7, 0).astype(np.uint32)
    for k in range(12):
        cy, cx = 12 + 25 * k, 360
        m = ((yy - cy) ** 2 + (xx - cx) ** 2 < 81) & (L == 0)
        L[m] = 20 + k
    I = _smooth_field(rng, H, W, 100, 4000, 3) + rng.integers(0, 200, (H, W)).astype(np.uint32)
    I[rng.random((H, W)) < 0.01] = 0            # intensity-0 pixels inside ROIs
    out.append(("blob_hole_mixed", I.astype(np.uint16), L, dict(g=64)))
    # 3. IBSI mode, <= 180 levels, zeros inside the ROI (GLDZM distance maps, zero-level exclusion)
    I = _smooth_field(rng, 230, 240, 1, 180, 2)
    I[rng.random((230, 240)) < 0.02] = 0
    L = np.ones((230, 240), np.uint32)
    L[:3, :5] = 0
    out.append(("ibsi_230x240", I, L, dict(ibsi=True)))
    # 4. 256 grey levels (BASELINE configs[2] settings) on a noisy 300x300 single ROI
    I = _smooth_field(rng, 300, 300, 2000, 5000, 4) + rng.normal(0, 400, (300, 300)).clip(-1500, 1500).astype(np.int64).astype(np.uint32)
    out.append(("g256_300x300", I.astype(np.uint16), np.ones((300, 300), np.uint32), dict(g=256)))
    # 5. sparse ROI: huge AABB, tiny area (giant texture path, small-cloud first-order path) + a long thin ROI
    I = _smooth_field(rng, 300, 420, 10, 3000, 1)
    L = np.zeros((300, 420), np.uint32)
    L[5:9, 4:9] = 3
    L[280:286, 400:410] = 3
    L[150, 10:400] = 5
    L[151, 10:400:2] = 5
    out.append(("sparse_and_thin", I, L, dict(g=64)))
    # 6. constant-intensity giant ROI (blank: every texture family softNaN) next to a low-contrast one
    I = np.full((250, 420), 777, np.uint32)
    L = np.zeros((250, 420), np.uint32)
    L[:, :205] = 1
    L[:, 210:] = 2
    I[:, 210:] = 40000 + rng.integers(0, 40, (250, 210)) ** 2 // 40   # skewed, narrow range on a high mean
    out.append(("blank_and_lowcontrast", I, L, dict(g=64)))
    return out


def big_volume_3d():
    """the seeded volume of tests/golden/golden_vol_big.npz: 44 x 56 x 64 uint16 (intensities <= 1500), one ROI whose bounding
    cube holds ~ 110 000 voxels (beyond the 40 000-voxel shared-memory classes), a medium one and two small ones"""
    rng = np.random.default_rng(4242)
    D, H, W = 44, 56, 64
    zz, yy, xx = np.mgrid[0:D, 0:H, 0:W]
    vol = ((np.sin(zz / 5.0) * np.cos(yy / 6.0) + np.sin(xx / 4.5) + 2.3) * 280.0 + rng.normal(0, 60, (D, H, W))).clip(1, 1500).astype(np.uint16)
    vol[rng.random((D, H, W)) < 0.004] = 0
    lab = np.zeros((D, H, W), np.uint16)
    lab[(((zz - 21) / 20.5) ** 2 + ((yy - 27) / 25.5) ** 2 + ((xx - 30) / 27.5) ** 2) < 1] = 2     # cube ~ 41 x 51 x 55
    lab[(((zz - 30) / 9.0) ** 2 + ((yy - 40) / 11.0) ** 2 + ((xx - 48) / 10.0) ** 2) < 1] = 5      # medium, carved out of the big one
    lab[0:4, 0:9, 52:64] = 7
    lab[40:44, 50:56, 0:5] = 11
    lab[rng.random((D, H, W)) < 0.02] = 0
    return vol, lab


def wmom_slide():
    """the seeded slide of tests/golden/golden_wmom.npz: concave blobs plus shapes that stress the reference's contour tracer --
    a ring (inner boundary), a one-pixel-wide bar and an L, a plus sign (X-crossing), a single pixel, a 2-pixel ROI, a
    checkerboard patch (isolated speckles), two separate parts under one label, a ROI touching the slide border, a U shape"""
    from nyxus_b200 import synth
    I, L = synth.blob_slide(192, 256, 32, 7, 13, seed=41, jitter=2)
    I = I.astype(np.uint32)
    b = int(L.max())
    yy, xx = np.mgrid[0:192, 0:256]
    free = np.zeros_like(L)
    L[160:192, :] = 0                                   # a band for the hand-made shapes
    rr = (yy - 175) ** 2 + (xx - 16) ** 2
    L[(rr <= 100) & (rr >= 25)] = b + 1                 # ring
    L[165, 40:70] = b + 2                               # bar, one pixel wide
    L[168:188, 80] = b + 3                              # L shape
    L[187, 80:95] = b + 3
    L[170:185, 107] = b + 4                             # plus sign
    L[177, 100:115] = b + 4
    L[163, 125] = b + 5                                 # single pixel
    L[170, 125:127] = b + 6                             # two pixels
    L[172:184:2, 135:147:2] = b + 7                     # speckles
    L[173:183:2, 136:146:2] = b + 7
    L[165:175, 155:165] = b + 8                         # two parts, one label
    L[178:188, 170:178] = b + 8
    L[180:192, 244:256] = b + 9                         # touches the slide corner
    L[163:186, 190:193] = b + 10                        # U shape
    L[163:186, 205:208] = b + 10
    L[183:186, 190:208] = b + 10
    L[164:170, 215:230] = b + 11                        # a solid rectangle with a one-pixel hole
    L[166, 220] = 0
    del free
    return I, L


def wmom_slide_big():
    """the seeded slide of tests/golden/golden_wmom_big.npz: 260 x 300, one ragged ROI with holes whose bounding box exceeds
    40 000 cells, a thin diagonal band and a small blob"""
    rng = np.random.default_rng(77)
    H, W = 260, 300
    yy, xx = np.mgrid[0:H, 0:W]
    I = (1500 + 900 * np.sin(xx / 17.0) * np.cos(yy / 23.0) + rng.normal(0, 200, (H, W))).clip(1, 60000).astype(np.uint32)
    ang = np.arctan2(yy - 125, xx - 140)
    rad = np.hypot((yy - 125) / 115.0, (xx - 140) / 130.0)
    L = np.zeros((H, W), np.uint32)
    L[rad < 0.93 + 0.06 * np.sin(7 * ang)] = 3
    L[(yy - 90) ** 2 + (xx - 100) ** 2 < 300] = 0          # holes
    L[(yy - 160) ** 2 + (xx - 190) ** 2 < 120] = 0
    L[120:124, 60:220] = 0                                   # a slit
    L[rng.random((H, W)) < 0.002] = 0                        # pinholes
    band = np.abs((yy - 240) - 0.06 * (xx - 10)) < 1.2
    L[band & (xx > 10) & (xx < 280) & (L == 0)] = 9
    L[(yy - 250) ** 2 + (xx - 285) ** 2 < 40] = 12
    return I, L
