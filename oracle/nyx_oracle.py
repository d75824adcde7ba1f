"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy + plain Python loops) of the reference's texture + intensity
feature path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product
(nyxus_b200/) never does.

Every function restates one reference routine (file:line under /root/reference/src/nyx) with DENSE matrices and the
reference's own loop structure -- deliberately unlike the CUDA path (sparse marginals, integer sums), so that the two
are independent witnesses.  Sizes: small ROIs only (pure Python loops).

PINNED: this restatement is checked (tests/test_oracle_cpu.py) against
  * the reference's own known-answer fixtures (tests/golden/reference_fixtures.py: IBSI phantom GLCM consensus values,
    the NGLDM / GLDZM IBSI matrices of reference tests/test_data.h:183-265), and
  * golden vectors produced by the UNMODIFIED reference compiled here (oracle/_ref, tests/golden/make_golden.py),
    features to 1e-9 relative and every matrix entry exactly.
"""
import math
import struct

import numpy as np

LOG10_2 = 0.30102999566       # glcm.h:241
GLCM_EPS = 0.000000001        # glcm.h:249
EPS = 2.2e-16                 # glrlm.h:169 glszm.h:138 gldm.h:105 gldzm.h:68

f32 = np.float32


def fast_log10(x):
    """helpers/helpers.h:283-330 -- float32 quadratic approximation of log2, scaled by log10(2)"""
    xf = f32(x)
    bits = struct.unpack("<I", struct.pack("<f", float(xf)))[0]
    e = (bits & 0x7F800000) >> 23
    a, b = f32(-.6296735), f32(1.466967)
    if bits & 0x00400000:
        sig = struct.unpack("<f", struct.pack("<I", (bits & 0x007FFFFF) | 0x3f000000))[0]
        fexp = f32(e) - f32(126.0)
    else:
        sig = struct.unpack("<f", struct.pack("<I", (bits & 0x007FFFFF) | 0x3f800000))[0]
        fexp = f32(e) - f32(127.0)
    s = f32(sig) - f32(1.0)
    lg2 = f32(f32(fexp + f32(f32(a * s) * s)) + f32(b * s))
    return float(lg2) * 0.30102999566


def bin_pixel_matlab(i, mx, g):
    """texture_feature.h:138-196"""
    if i == 0:
        return 1
    slope = float(g) / float(mx) if mx else math.inf
    v = int(math.floor(slope * float(i) + 1.0))
    return max(1, min(v, g))


def bin_pixel_radiomics(x, mn, mx, n):
    """TextureFeature::to_grayscale_radiomix (texture_feature.h:106-118): 0 stays 0; 1-based bins of width (max-min)/n, the
    last one a unit wider.  (PixIntens)(double) as x86-64 gcc does it: a 64-bit truncating conversion whose low 32 bits are
    kept, NaN (min == max: 0/0) and out-of-range values giving 0."""
    if x == 0:
        return 0
    d = float((x - mn) & 0xFFFFFFFF)
    w = float((mx - mn) & 0xFFFFFFFF) / float(n)
    if w == 0.0:
        q = math.nan if d == 0.0 else math.inf
    else:
        q = d / w + 1.0
    y = 0 if (q != q or abs(q) >= 9.223372036854775808e18) else int(q) & 0xFFFFFFFF
    return min(y, n)


def bin_tile(R, mn, mx, greyinfo):
    """TextureFeature::bin_intensities (texture_feature.h:23-50): >0 MATLAB, 0 IBSI identity, <0 PyRadiomics bin count"""
    if greyinfo == 0:
        return R.astype(np.int64).copy()
    D = np.zeros(R.shape, np.int64)
    for idx, v in np.ndenumerate(R):
        D[idx] = bin_pixel_matlab(int(v), int(mx), greyinfo) if greyinfo > 0 else bin_pixel_radiomics(int(v), int(mn), int(mx), -greyinfo)
    return D


def to_grayscale(i, mn, rng, n, disable=False):
    """helpers.h:337-345"""
    if disable:
        return int(i)
    if rng == 0:
        return 0        # NaN -> unsigned cast, x86-64 gcc yields 0
    return int(float(i - mn) / float(rng) * float(n))


def roi_from_tile(tile, mask):
    """phase 1 + phase 2 for one ROI (tests/test_main_nyxus.h:50-90): stats, cloud (column-major like the in-memory
    path, phase1.cpp:383-384), AABB tile with 0 outside the ROI (image_matrix.h:284-304)"""
    tile = np.asarray(tile, np.int64)
    mask = np.asarray(mask).astype(bool)
    ys, xs = np.nonzero(mask)
    y0, y1, x0, x1 = ys.min(), ys.max(), xs.min(), xs.max()
    R = np.where(mask, tile, 0)[y0:y1 + 1, x0:x1 + 1]
    M = mask[y0:y1 + 1, x0:x1 + 1]
    cloud = [int(tile[y, x]) for x in range(tile.shape[1]) for y in range(tile.shape[0]) if mask[y, x]]
    return dict(R=R, M=M, cloud=cloud, mn=min(cloud), mx=max(cloud), area=len(cloud))


# --------------------------------------------------------------------------------------------------
# GLCM (features/glcm.cpp)
# --------------------------------------------------------------------------------------------------
GLCM_NAMES = ["ASM", "ACOR", "CLUPROM", "CLUSHADE", "CLUTEND", "CONTRAST", "CORRELATION", "DIFAVE", "DIFENTRO", "DIFVAR", "DIS",
              "ENERGY", "ENTROPY", "HOM1", "HOM2", "ID", "IDN", "IDM", "IDMN", "INFOMEAS1", "INFOMEAS2", "IV", "JAVE", "JE", "JMAX",
              "JVAR", "SUMAVERAGE", "SUMENTROPY", "SUMVARIANCE", "VARIANCE"]
GLCM_AVE = ["ASM", "ACOR", "CLUPROM", "CLUSHADE", "CLUTEND", "CONTRAST", "CORRELATION", "DIFAVE", "DIFENTRO", "DIFVAR", "DIS",
            "ENERGY", "ENTROPY", "HOM1", "ID", "IDN", "IDM", "IDMN", "IV", "JAVE", "JE", "INFOMEAS1", "INFOMEAS2", "VARIANCE",
            "JMAX", "JVAR", "SUMAVERAGE", "SUMENTROPY", "SUMVARIANCE"]


def glcm_matrix(roi, g, ibsi, dx, dy, symmetric=False):
    """calculateCoocMatAtAngle (glcm.cpp:343-485): P[row=centre level-1][col=neighbour level-1]"""
    R = roi["R"]
    greyinfo = 0 if ibsi else g
    D = bin_tile(R, roi["mn"], roi["mx"], greyinfo)
    Ng = g if greyinfo > 0 else int(D.max())
    P = np.zeros((Ng, Ng))
    h, w = R.shape
    for row in range(h):
        for col in range(w):
            r2, c2 = row + dy, col + dx
            if 0 <= r2 < h and 0 <= c2 < w:
                if R[row, col] == 0 or R[r2, c2] == 0:      # :445
                    continue
                b, a = D[row, col] - 1, D[r2, c2] - 1
                P[b, a] += 1                                # GLCM.xy(a,b)++  (:472)
                if symmetric or greyinfo == 0:              # :475
                    P[a, b] += 1
    return P, list(range(1, Ng + 1))


def glcm_features_1angle(P, I, softnan=0.0):
    """Extract_Texture_Features2 (glcm.cpp:227-341) and the 30 reducers (:546-1202).  P[r][c] == P.yx(r,c) == P.xy(c,r)"""
    Ng = len(I)
    sum_p = P.sum()
    if sum_p == 0:
        return {n: softnan for n in GLCM_NAMES}
    Pxpy = np.zeros(2 * Ng)
    Pxmy = np.zeros(Ng)
    for x in range(Ng):                                     # :503-513
        for y in range(Ng):
            Pxpy[x + y] += P[x, y] / sum_p
            Pxmy[abs(x - y)] += P[x, y] / sum_p
    kSum = np.array([k + 2 for k in range(2 * Ng)], float)
    kDiff = np.array([k for k in range(Ng)], float)
    px_col = np.zeros(Ng)                                   # :522-525  px[i] += P.xy(i,j)  -> column marginal
    for j in range(Ng):
        for i in range(Ng):
            px_col[i] += P[j, i] / sum_p
    by_row_mean = sum(px_col[i] * I[i] for i in range(Ng))  # :531-536
    f = {}
    f["ASM"] = sum((P[j, i] / sum_p) * (P[j, i] / sum_p) for j in range(Ng) for i in range(Ng))
    f["ENERGY"] = f["ASM"]
    f["CONTRAST"] = sum(P[i, j] * (I[i] - I[j]) ** 2 for i in range(Ng) for j in range(Ng)) / sum_p
    # f_corr :593-644
    mr = sum(P[r, c] * I[r] for c in range(Ng) for r in range(Ng)) / sum_p
    mc = sum(P[r, c] * I[c] for c in range(Ng) for r in range(Ng)) / sum_p
    s2r = sum(P[r, c] / sum_p * (I[r] - mr) ** 2 for c in range(Ng) for r in range(Ng))
    s2c = sum(P[r, c] / sum_p * (I[c] - mc) ** 2 for c in range(Ng) for r in range(Ng))
    tmp1 = sum((I[r] - mr) * (I[c] - mc) * P[r, c] / sum_p for c in range(Ng) for r in range(Ng))
    den = math.sqrt(s2r) * math.sqrt(s2c)
    f["CORRELATION"] = tmp1 / den if den > 0 else softnan
    # f_var :647-676
    mean = sum(P[r, :].sum() * I[r] for r in range(Ng)) / sum_p
    f["VARIANCE"] = sum((I[r] - mean) ** 2 * P[r, c] for r in range(Ng) for c in range(Ng)) / sum_p
    f["IDM"] = sum(Pxmy[k] / (1 + k * k) for k in range(Ng))
    f["SUMAVERAGE"] = sum(kSum[i] * Pxpy[i] for i in range(2 * Ng))
    f["SUMENTROPY"] = -sum(Pxpy[k] * fast_log10(Pxpy[k] + GLCM_EPS) / LOG10_2 for k in range(2 * Ng))
    ent = 0.0
    for j in range(Ng):
        for i in range(Ng):
            p = P[j, i] / sum_p
            ent += p * fast_log10(p + GLCM_EPS) / LOG10_2
    f["ENTROPY"] = -ent
    f["JE"] = -ent
    diffavg = sum(kDiff[i] * Pxmy[i] for i in range(Ng))
    f["DIFAVE"] = diffavg
    f["DIFVAR"] = sum(Ng * ((k - diffavg) ** 2 * Pxmy[k]) for k in range(Ng)) / Ng     # :742-766
    f["DIFENTRO"] = -sum(Pxmy[k] * fast_log10(Pxmy[k] + GLCM_EPS) / LOG10_2 for k in range(Ng) if Pxmy[k] != 0)
    # info measures :846-914 (px over columns, py over rows)
    px = np.zeros(Ng)
    py = np.zeros(Ng)
    for i in range(Ng):
        for j in range(Ng):
            px[i] += P[j, i] / sum_p
            py[j] += P[j, i] / sum_p
    HXY = HXY1 = HXY2 = HX = 0.0
    for i in range(Ng):
        for j in range(Ng):
            p = P[j, i] / sum_p
            HXY += p * fast_log10(p + GLCM_EPS) / LOG10_2
            HXY1 += p * fast_log10(px[i] * py[j] + GLCM_EPS) / LOG10_2
            HXY2 += px[i] * py[j] * fast_log10(px[i] * py[j] + GLCM_EPS) / LOG10_2
    for i in range(Ng):
        HX += px[i] * fast_log10(px[i] + GLCM_EPS) / LOG10_2
    with np.errstate(all="ignore"):
        r1 = (HXY - HXY1) / HX if HX != 0 else (math.nan if HXY - HXY1 == 0 else math.inf)
    f["INFOMEAS1"] = r1 if math.isfinite(r1) else softnan
    f["INFOMEAS2"] = math.sqrt(abs(1 - math.exp(-2 * (-HXY2 + HXY))))
    f["HOM1"] = sum(P[r, c] / sum_p / (1.0 + abs(r - c)) for r in range(Ng) for c in range(Ng))
    f["ACOR"] = sum(P[y, x] * I[x] * I[y] for x in range(Ng) for y in range(Ng)) / sum_p
    f["CLUPROM"] = sum((I[r] + I[c] - 2 * by_row_mean) ** 4 * P[r, c] / sum_p for r in range(Ng) for c in range(Ng))
    f["CLUSHADE"] = sum((I[r] + I[c] - 2 * by_row_mean) ** 3 * P[r, c] / sum_p for r in range(Ng) for c in range(Ng))
    f["CLUTEND"] = sum((I[x] + I[y] - by_row_mean * 2.0) ** 2 * P[y, x] / sum_p for x in range(Ng) for y in range(Ng))
    f["SUMVARIANCE"] = f["CLUTEND"]
    f["DIS"] = sum(abs(x - y) * P[y, x] / sum_p for x in range(Ng) for y in range(Ng))
    f["HOM2"] = sum((P[y, x] / sum_p) / (1.0 + abs(x - y) ** 2) for x in range(Ng) for y in range(Ng))
    f["IDMN"] = sum(Pxmy[k] / (1.0 + (k * k) / (Ng * Ng)) for k in range(Ng))
    f["ID"] = sum(Pxmy[k] / (1.0 + k) for k in range(Ng))
    f["IDN"] = sum(Pxmy[k] / (1.0 + k / Ng) for k in range(Ng))
    f["IV"] = sum(Pxmy[k] / (kDiff[k] ** 2) for k in range(1, Ng))
    jave = sum(P[i, j] * I[i] for i in range(Ng) for j in range(Ng)) / sum_p
    f["JAVE"] = jave
    f["JMAX"] = max(P[y, x] / sum_p for x in range(Ng) for y in range(Ng))
    f["JVAR"] = sum((x + 1 - jave) ** 2 * P[y, x] / sum_p for x in range(Ng) for y in range(Ng))   # :1185-1202
    return f


def glcm(roi, g=64, ibsi=False, offset=1, softnan=0.0, symmetric=False):
    """GLCMFeature::calculate + save_value (glcm.cpp:16-100,143-208).  Returns (dict name -> [4 angles] / AVE, matrices)"""
    if bin_pixel_matlab(roi["mn"], roi["mx"], g) == bin_pixel_matlab(roi["mx"], roi["mx"], g):     # :27-29
        out = {"GLCM_" + n: [softnan] * 4 for n in GLCM_NAMES}
        out.update({"GLCM_" + n + "_AVE": softnan for n in GLCM_AVE})
        return out, None
    shifts = [(offset, 0), (offset, offset), (0, offset), (-offset, offset)]   # (dx,dy) :234-251
    per = []
    mats = []
    for dx, dy in shifts:
        P, I = glcm_matrix(roi, g, ibsi, dx, dy, symmetric)
        mats.append(P)
        per.append(glcm_features_1angle(P, I, softnan))
    out = {"GLCM_" + n: [per[a][n] for a in range(4)] for n in GLCM_NAMES}
    for n in GLCM_AVE:
        v = out["GLCM_" + n]
        out["GLCM_" + n + "_AVE"] = (((v[0] + v[1]) + v[2]) + v[3]) / 4.0
    return out, mats


# --------------------------------------------------------------------------------------------------
# GLRLM (features/glrlm.cpp)
# --------------------------------------------------------------------------------------------------
GLRLM_NAMES = ["SRE", "LRE", "GLN", "GLNN", "RLN", "RLNN", "RP", "GLV", "RV", "RE", "LGLRE", "HGLRE", "SRLGLE", "SRHGLE", "LRLGLE", "LRHGLE"]
U32 = 0xFFFFFFFF


def glrlm_runs(D, angle_idx):
    """glrlm.cpp:107-195: forward-only runs E / SE / S / SW on a scratch copy with VISITED marks"""
    D = D.copy()
    h, w = D.shape
    step = [(1, 0), (1, 1), (0, 1), (-1, 1)][angle_idx]
    Z = []
    for row in range(h):
        for col in range(w):
            pi = D[row, col]
            if pi == 0 or pi == -1:
                continue
            x, y, n = col, row, 1
            D[y, x] = -1
            while True:
                nx, ny = x + step[0], y + step[1]
                if 0 <= nx < w and 0 <= ny < h and D[ny, nx] == pi:
                    D[ny, nx] = -1
                    n += 1
                    x, y = nx, ny
                    continue
                break
            Z.append((int(pi), n))
    return Z


def glrlm(roi, g=64, ibsi=False, softnan=0.0):
    if roi["mn"] == roi["mx"]:                                  # :29-52
        out = {"GLRLM_" + n: [softnan] * 4 for n in GLRLM_NAMES}
        out.update({"GLRLM_" + n + "_AVE": softnan for n in GLRLM_NAMES})
        return out, None
    R = roi["R"]
    D = bin_tile(R, roi["mn"], roi["mx"], 0 if ibsi else g)
    Np = int((R != 0).sum())                                    # :198-204
    per = {n: [] for n in GLRLM_NAMES}
    mats = []
    for a in range(4):
        Z = glrlm_runs(D, a)
        sp = {}
        for lv, n in Z:
            sp[(lv, n)] = sp.get((lv, n), 0) + 1
        mats.append(sp)
        s = float(len(Z))
        if s == 0:
            for n in GLRLM_NAMES:
                per[n].append(0.0)
            continue
        levels = sorted({k[0] for k in sp})
        lens = sorted({k[1] for k in sp})
        ri = {lv: sum(c for (l2, _), c in sp.items() if l2 == lv) for lv in levels}
        rj = {j: sum(c for (_, j2), c in sp.items() if j2 == j) for j in lens}
        per["SRE"].append(sum(rj[j] / (j * j) for j in lens) / s)
        per["LRE"].append(sum(c * j * j for (_, j), c in sorted(sp.items())) / s)
        per["GLN"].append(sum(ri[lv] ** 2 for lv in levels) / s)
        per["GLNN"].append(sum(ri[lv] ** 2 for lv in levels) / (s * s))
        per["RLN"].append(sum(rj[j] ** 2 for j in lens) / s)
        per["RLNN"].append(sum(rj[j] ** 2 for j in lens) / (s * s))
        per["RP"].append(s / Np if Np else math.inf)
        mu = sum(c / s * lv for (lv, _), c in sorted(sp.items()))
        per["GLV"].append(sum(c / s * (lv - mu) ** 2 for (lv, _), c in sorted(sp.items())))
        muj = sum(c / s * j for (_, j), c in sorted(sp.items()))
        per["RV"].append(sum(c / s * (j - muj) ** 2 for (_, j), c in sorted(sp.items())))
        per["RE"].append(-sum(c / s * (fast_log10(c / s + EPS) / LOG10_2) for _, c in sorted(sp.items())))
        u = lambda v: v & U32                                    # 32-bit unsigned products (:717,749,781,813,844,876)
        div = lambda a_, b_: a_ / b_ if b_ else math.inf
        per["LGLRE"].append(sum(div(c, float(u(lv * lv))) for (lv, _), c in sorted(sp.items())) / s)
        per["HGLRE"].append(sum(c * float(u(lv * lv)) for (lv, _), c in sorted(sp.items())) / s)
        per["SRLGLE"].append(sum(div(c, float(u(lv * lv * j * j))) for (lv, j), c in sorted(sp.items())) / s)
        per["SRHGLE"].append(sum(c * float(u(lv * lv)) / float(j * j) for (lv, j), c in sorted(sp.items())) / s)
        per["LRLGLE"].append(sum(div(c * float(j * j), float(u(lv * lv))) for (lv, j), c in sorted(sp.items())) / s)
        per["LRHGLE"].append(sum(c * float(u(lv * lv * j * j)) for (lv, j), c in sorted(sp.items())) / s)
    out = {"GLRLM_" + n: per[n] for n in GLRLM_NAMES}
    for n in GLRLM_NAMES:
        v = per[n]
        out["GLRLM_" + n + "_AVE"] = (((v[0] + v[1]) + v[2]) + v[3]) / 4.0
    return out, mats


# --------------------------------------------------------------------------------------------------
# GLSZM (features/glszm.cpp)
# --------------------------------------------------------------------------------------------------
GLSZM_NAMES = ["SAE", "LAE", "GLN", "GLNN", "SZN", "SZNN", "ZP", "GLV", "ZV", "ZE", "LGLZE", "HGLZE", "SALGLE", "SAHGLE", "LALGLE", "LAHGLE"]


def glszm_zones(D):
    """glszm.cpp:108-194: the reference's depth-first flood with an explicit history stack, E / SE / S / SW only"""
    D = D.copy()
    h, w = D.shape
    Z = []
    for row in range(h):
        for col in range(w):
            pi = D[row, col]
            if pi == 0 or pi == -1:
                continue
            hist = []
            x, y, n = col, row, 1
            D[y, x] = -1
            while True:
                moved = False
                for dx, dy in ((1, 0), (1, 1), (0, 1), (-1, 1)):
                    nx, ny = x + dx, y + dy
                    if 0 <= nx < w and 0 <= ny < h and D[ny, nx] != -1 and D[ny, nx] == pi:
                        D[ny, nx] = -1
                        n += 1
                        hist.append((x, y))
                        x, y = nx, ny
                        moved = True
                        break
                if moved:
                    continue
                if hist:
                    x, y = hist.pop()
                    continue
                break
            Z.append((int(pi), n))
    return Z


def glszm(roi, g=64, ibsi=False, softnan=0.0):
    if roi["mn"] == roi["mx"]:
        return {"GLSZM_" + n: softnan for n in GLSZM_NAMES}, None
    D = bin_tile(roi["R"], roi["mn"], roi["mx"], 0 if ibsi else g)
    Z = glszm_zones(D)
    Np = int((D != 0).sum())                                    # :199-204 (binned tile)
    sp = {}
    for k in Z:
        sp[k] = sp.get(k, 0) + 1
    s = float(len(Z))
    if s == 0:
        return {"GLSZM_" + n: softnan for n in GLSZM_NAMES}, sp
    items = sorted(sp.items())
    levels = sorted({k[0] for k in sp})
    sizes = sorted({k[1] for k in sp})
    si = {lv: sum(c for (l2, _), c in items if l2 == lv) for lv in levels}
    sj = {j: sum(c for (_, j2), c in items if j2 == j) for j in sizes}
    f = {}
    f["SAE"] = sum(sj[j] / (j * j) for j in sizes) / s
    f["LAE"] = sum(sj[j] * (j * j) for j in sizes) / s
    f["GLN"] = sum(v * v for v in si.values()) / s
    f["GLNN"] = sum(v * v for v in si.values()) / (s * s)
    f["SZN"] = sum(v * v for v in sj.values()) / s
    f["SZNN"] = sum(v * v for v in sj.values()) / (s * s)
    f["ZP"] = s / Np
    mu_glv = sum(c / s * lv for (lv, _), c in items)
    mu_zv = sum(c / s * j for (_, j), c in items)
    f["GLV"] = sum(c / s * (lv - mu_glv) ** 2 for (lv, _), c in items)
    f["ZV"] = sum(c / s * (j - mu_zv) ** 2 for (_, j), c in items)
    f["ZE"] = -sum(c / s * (fast_log10(c / s + EPS) / LOG10_2) for _, c in items)
    f["LGLZE"] = sum(si[lv] / (lv * lv) for lv in levels) / s
    f["HGLZE"] = sum(si[lv] * (lv * lv) for lv in levels) / s
    f["SALGLE"] = sum(c / (float(lv * lv) * float(j * j)) for (lv, j), c in items) / s
    f["SAHGLE"] = sum(c * float(lv * lv) / float(j * j) for (lv, j), c in items) / s
    f["LALGLE"] = sum(c * float(j * j) / float(lv * lv) for (lv, j), c in items) / s
    f["LAHGLE"] = sum(c * float(lv * lv) * float(j * j) for (lv, j), c in items) / s
    return {"GLSZM_" + k: v for k, v in f.items()}, sp


# --------------------------------------------------------------------------------------------------
# GLDZM (features/gldzm.cpp)
# --------------------------------------------------------------------------------------------------
GLDZM_NAMES = ["SDE", "LDE", "LGLZE", "HGLZE", "SDLGLE", "SDHGLE", "LDLGLE", "LDHGLE", "GLNU", "GLNUN", "ZDNU", "ZDNUN", "ZP", "GLM",
               "GLV", "ZDM", "ZDV", "ZDE"]


def dist2border(D, x, y):
    """gldzm.cpp:306-351"""
    h, w = D.shape
    dl = dr = dt = db = 0
    for x0 in range(x - 1, -1, -1):
        if D[y, x0] == 0 or x0 == 0:
            dl = x - x0
            break
    for x0 in range(x + 1, w):
        if D[y, x0] == 0 or x0 == w - 1:
            dr = x0 - x
            break
    for y0 in range(y - 1, -1, -1):
        if D[y0, x] == 0 or y0 == 0:
            dt = y - y0
            break
    for y0 in range(y + 1, h):
        if D[y0, x] == 0 or y0 == h - 1:
            db = y0 - y
            break
    m = min(dl + 1, dr + 1, dt + 1, db + 1)
    return 1 if m == 0 else m


def gldzm(roi, g=64, ibsi=False, softnan=0.0):
    if roi["mn"] == roi["mx"]:
        return {"GLDZM_" + n: softnan for n in GLDZM_NAMES}, None
    greyinfo = 0 if ibsi else g
    D = bin_tile(roi["R"], roi["mn"], roi["mx"], greyinfo)
    h, w = D.shape
    V = -1
    Z = []
    for row in range(h):                                        # :94-237 (E, S, W, N with a parent stack)
        for col in range(w):
            inten = D[row, col]
            if inten == V:
                continue
            if greyinfo == 0 and inten == 0:
                continue
            stack = []
            x, y, size = col, row, 1
            metric = dist2border(D, x, y)
            while True:
                D[y, x] = V
                moved = False
                for dx, dy in ((1, 0), (0, 1), (-1, 0), (0, -1)):
                    nx, ny = x + dx, y + dy
                    if 0 <= nx < w and 0 <= ny < h and D[ny, nx] != V and D[ny, nx] == inten:
                        stack.append((x, y))
                        metric = min(metric, dist2border(D, nx, ny))
                        size += 1
                        x, y = nx, ny
                        moved = True
                        break
                if moved:
                    continue
                if stack:
                    x, y = stack.pop()
                else:
                    break
            Z.append((int(inten), metric))
    sp = {}
    for k in Z:
        sp[k] = sp.get(k, 0) + 1
    items = sorted(sp.items())
    levels = sorted({k[0] for k in sp})
    dists = sorted({k[1] for k in sp})
    Mx = {lv: float(sum(c for (l2, _), c in items if l2 == lv)) for lv in levels}
    Md = {d: float(sum(c for (_, d2), c in items if d2 == d)) for d in dists}
    Ns = float(sum(sp.values()))
    Nv = float(roi["area"])
    f = {}
    f["SDE"] = sum(Md[d] / d / d for d in dists) / Ns
    f["LDE"] = sum(d * d * Md[d] for d in dists) / Ns
    f["ZDNU"] = sum(Md[d] ** 2 for d in dists) / Ns
    f["ZDNUN"] = f["ZDNU"] / Ns
    f["LGLZE"] = sum(Mx[lv] / (lv * lv) for lv in levels) / Ns
    f["HGLZE"] = sum((lv * lv) * Mx[lv] for lv in levels) / Ns
    f["GLNU"] = sum(Mx[lv] ** 2 for lv in levels) / Ns
    f["GLNUN"] = f["GLNU"] / Ns
    f["SDLGLE"] = sum(c / lv / lv / d / d for (lv, d), c in items) / Ns
    f["SDHGLE"] = sum(lv * lv * c / d / d for (lv, d), c in items) / Ns
    f["LDLGLE"] = sum(d * d * c / lv / lv for (lv, d), c in items) / Ns
    f["LDHGLE"] = sum(lv * lv * d * d * c for (lv, d), c in items) / Ns
    f["GLM"] = sum(lv * c for (lv, d), c in items) / Ns
    f["ZDM"] = sum(d * c for (lv, d), c in items) / Ns
    f["ZDE"] = -sum(c / Ns * math.log2(c / Ns + EPS) for _, c in items)
    f["ZP"] = Ns / Nv
    f["GLV"] = sum((lv - f["GLM"]) ** 2 * (c / Ns) for (lv, d), c in items)
    f["ZDV"] = sum((d - f["ZDM"]) ** 2 * (c / Ns) for (lv, d), c in items)
    return {"GLDZM_" + k: v for k, v in f.items()}, sp


# --------------------------------------------------------------------------------------------------
# GLDM (features/gldm.cpp)
# --------------------------------------------------------------------------------------------------
GLDM_NAMES = ["SDE", "LDE", "GLN", "DN", "DNN", "GLV", "DV", "DE", "LGLE", "HGLE", "SDLGLE", "SDHGLE", "LDLGLE", "LDHGLE"]
NB8 = [(-1, 0), (-1, 1), (0, 1), (1, 1), (1, 0), (1, -1), (0, -1), (-1, -1)]    # (dy,dx)


def gldm(roi, g=64, ibsi=False, softnan=0.0):
    if roi["mn"] == roi["mx"]:
        return {"GLDM_" + n: softnan for n in GLDM_NAMES}, None
    R = roi["R"]
    D = bin_tile(R, roi["mn"], roi["mx"], 0 if ibsi else g)
    h, w = R.shape
    sp = {}
    for row in range(h):                                        # :74-166
        for col in range(w):
            if R[row, col] == 0:
                continue
            nd = 1
            for dy, dx in NB8:
                y, x = row + dy, col + dx
                if 0 <= y < h and 0 <= x < w and R[y, x] != 0 and D[y, x] == D[row, col]:
                    nd += 1
            k = (int(D[row, col]), nd)
            sp[k] = sp.get(k, 0) + 1
    Nz = float(sum(sp.values()))
    if Nz == 0:
        return {"GLDM_" + n: softnan for n in GLDM_NAMES}, sp
    items = sorted(sp.items())
    si = {}
    sj = {}
    for (lv, j), c in items:
        si[lv] = si.get(lv, 0) + c
        sj[j] = sj.get(j, 0) + c
    f = {}
    f["SDE"] = sum(c / (float(j) * float(j)) for (lv, j), c in items) / Nz
    f["LDE"] = sum(c * (float(j) * float(j)) for (lv, j), c in items) / Nz
    f["GLN"] = sum(v * v for v in si.values()) / Nz
    f["DN"] = sum(v * v for v in sj.values()) / Nz
    f["DNN"] = sum(v * v for v in sj.values()) / (Nz * Nz)
    mu = sum(c / Nz * lv for (lv, j), c in items)
    f["GLV"] = sum(c / Nz * (lv - mu) ** 2 for (lv, j), c in items)
    muj = sum(c / Nz * j for (lv, j), c in items)
    f["DV"] = sum(c / Nz * (j - muj) ** 2 for (lv, j), c in items)
    f["DE"] = -sum(c / Nz * (fast_log10(c / Nz + EPS) / LOG10_2) for _, c in items)
    f["LGLE"] = sum(si[lv] / float(lv * lv) for lv in sorted(si)) / Nz
    f["HGLE"] = sum(si[lv] * float(lv) * float(lv) for lv in sorted(si)) / Nz
    f["SDLGLE"] = sum(c / (float(lv) * lv * j * j) for (lv, j), c in items) / Nz
    f["SDHGLE"] = sum(c * (float(lv) * lv) / float(j * j) for (lv, j), c in items) / Nz
    f["LDLGLE"] = sum(c * float(j * j) / (float(lv) * lv) for (lv, j), c in items) / Nz
    f["LDHGLE"] = sum(c * (float(lv) * lv * j * j) for (lv, j), c in items) / Nz
    return {"GLDM_" + k: v for k, v in f.items()}, sp


# --------------------------------------------------------------------------------------------------
# NGLDM (features/ngldm.cpp)
# --------------------------------------------------------------------------------------------------
NGLDM_NAMES = ["LDE", "HDE", "LGLCE", "HGLCE", "LDLGLE", "LDHGLE", "HDLGLE", "HDHGLE", "GLNU", "GLNUN", "DCNU", "DCNUN", "DCP", "GLM",
               "GLV", "DCM", "DCV", "DCENT", "DCENE"]


def ngldm(roi, g=64, ibsi=False, softnan=0.0):
    if roi["mn"] == roi["mx"]:
        return {"NGLDM_" + n: softnan for n in NGLDM_NAMES}, None
    R, M = roi["R"], roi["M"]
    h, w = R.shape
    mx = roi["mx"]
    lev = lambda v: to_grayscale(int(v), 0, mx, g, ibsi)        # ngldm.cpp:42,126,144
    sp = {}
    for y in range(h):
        for x in range(w):
            if not M[y, x]:
                continue
            c0 = lev(R[y, x])
            n = 0
            for dy, dx in NB8:
                yy, xx = y + dy, x + dx
                if 0 <= yy < h and 0 <= xx < w and M[yy, xx] and lev(R[yy, xx]) == c0:
                    n += 1
            sp[(c0, n)] = sp.get((c0, n), 0) + 1
    items = sorted(sp.items())
    Ns = float(sum(sp.values()))
    Sg, Sr = {}, {}
    for (lv, j), c in items:
        Sg[lv] = Sg.get(lv, 0) + c
        Sr[j] = Sr.get(j, 0) + c
    f = {n: 0.0 for n in NGLDM_NAMES}
    for (lv, j), c in items:                                    # :283-305
        gl, dc, sij = float(lv), float(j + 1), float(c)
        pij = sij / Ns
        f["LDE"] += sij / dc / dc
        f["HDE"] += sij * dc * dc
        if gl != 0.0:
            f["LGLCE"] += sij / gl / gl
            f["LDLGLE"] += sij / dc / dc / gl / gl
            f["HDLGLE"] += sij * dc * dc / gl / gl
        f["HGLCE"] += sij * gl * gl
        f["LDHGLE"] += sij * gl * gl / dc / dc
        f["HDHGLE"] += sij * dc * dc * gl * gl
        f["GLM"] += gl * pij
        f["DCM"] += dc * pij
        if pij > 0:
            f["DCENT"] -= pij * math.log(pij) / math.log(2)
        f["DCENE"] += pij * pij
    f["GLNU"] = sum(v * v for v in Sg.values())
    f["GLNUN"] = f["GLNU"]
    f["DCNU"] = sum(v * v for v in Sr.values())
    f["DCNUN"] = f["DCNU"]
    for (lv, j), c in items:
        pij = c / Ns
        f["GLV"] += (lv - f["GLM"]) ** 2 * pij
        f["DCV"] += (j + 1 - f["DCM"]) ** 2 * pij
    for n in ("LDE", "HDE", "LGLCE", "HGLCE", "LDLGLE", "LDHGLE", "HDLGLE", "HDHGLE", "GLNU", "DCNU"):
        f[n] /= Ns
    f["GLNUN"] /= Ns * Ns
    f["DCNUN"] /= Ns * Ns
    f["DCP"] = 1.0
    return {"NGLDM_" + k: v for k, v in f.items()}, sp


# --------------------------------------------------------------------------------------------------
# NGTDM (features/ngtdm.cpp)
# --------------------------------------------------------------------------------------------------
NGTDM_NAMES = ["COARSENESS", "CONTRAST", "BUSYNESS", "COMPLEXITY", "STRENGTH"]


def ngtdm(roi, g=64, ibsi=False, softnan=0.0):
    D = bin_tile(roi["R"], roi["mn"], roi["mx"], 0 if ibsi else g)
    h, w = D.shape
    U = sorted(set(int(v) for v in D.ravel()) - {0})
    if not U:
        return {"NGTDM_" + n: softnan for n in NGTDM_NAMES}, None
    I = list(range(0, max(U) + 1)) if ibsi else U               # :57-67
    if len(I) < 2:
        return {"NGTDM_" + n: softnan for n in NGTDM_NAMES}, None
    N = {lv: 0 for lv in I}
    S = {lv: 0.0 for lv in I}
    Nvp = 0
    for row in range(h):                                        # :83-146, 159-183
        for col in range(w):
            pi = int(D[row, col])
            if pi == 0:
                continue
            s, nd = 0.0, 0
            for dy, dx in NB8:
                y, x = row + dy, col + dx
                if 0 <= y < h and 0 <= x < w and D[y, x] != 0:
                    s += float(D[y, x])
                    nd += 1
            if nd > 0:
                a = s / nd
                N[pi] += 1
                S[pi] += abs(pi - a)
                if a > 0:
                    Nvp += 1
    Nvc = float(sum(N.values()))
    Ngp = len(U)
    mats = {lv: (N[lv], S[lv]) for lv in I if N[lv]}
    with np.errstate(all="ignore"):
        P = {lv: (np.float64(N[lv]) / np.float64(Nvc)) for lv in I}
        f = {}
        sps = sum(P[i] * S[i] for i in I)
        f["COARSENESS"] = float(np.float64(1.0) / np.float64(sps))
        c1 = sum(P[i] * P[j] * (i - j) ** 2 for i in I for j in I)
        ngp2 = Ngp * (Ngp - 1) if Ngp > 1 else Ngp
        f["CONTRAST"] = float((c1 / ngp2) * (np.float64(sum(S.values())) / np.float64(Nvc)))
        if Ngp == 1:
            f["BUSYNESS"] = 0.0
        else:
            s2 = sum(abs(P[i] * i - P[j] * j) for i in I for j in I if P[i] != 0 and P[j] != 0)
            f["BUSYNESS"] = 0.0 if s2 == 0 else float(sps / s2)
        cx = sum(abs(i - j) * (P[i] * S[i] + P[j] * S[j]) / (P[i] + P[j]) for i in I for j in I if P[i] != 0 and P[j] != 0)
        f["COMPLEXITY"] = float(np.float64(cx) / np.float64(Nvp))
        st = sum((P[i] + P[j]) * (i - j) ** 2 for i in I for j in I if P[i] != 0 and P[j] != 0)
        f["STRENGTH"] = float(np.float64(st) / np.float64(sum(S.values())))
    return {"NGTDM_" + k: v for k, v in f.items()}, mats


# --------------------------------------------------------------------------------------------------
# first-order (features/intensity.cpp, histogram.h, moments.h)
# --------------------------------------------------------------------------------------------------
INT_NAMES = ["COV", "COVERED_IMAGE_INTENSITY_RANGE", "ENERGY", "ENTROPY", "EXCESS_KURTOSIS", "HYPERFLATNESS", "HYPERSKEWNESS",
             "INTEGRATED_INTENSITY", "INTERQUARTILE_RANGE", "KURTOSIS", "MAX", "MEAN", "MEAN_ABSOLUTE_DEVIATION", "MEDIAN",
             "MEDIAN_ABSOLUTE_DEVIATION", "MIN", "MODE", "P01", "P10", "P25", "P75", "P90", "P99", "QCOD", "RANGE", "ROBUST_MEAN",
             "ROBUST_MEAN_ABSOLUTE_DEVIATION", "ROOT_MEAN_SQUARED", "SKEWNESS", "STANDARD_DEVIATION", "STANDARD_DEVIATION_BIASED",
             "STANDARD_ERROR", "VARIANCE", "VARIANCE_BIASED", "UNIFORMITY", "UNIFORMITY_PIU"]


def intensity(roi, g=64):
    """PixelIntensityFeatures::calculate (intensity.cpp:57-192) with sequential loops in cloud order"""
    c = roi["cloud"]
    n = float(len(c))
    mn, mx = roi["mn"], roi["mx"]
    f = {k: 0.0 for k in INT_NAMES}
    f["MIN"], f["MAX"], f["RANGE"] = float(mn), float(mx), float(mx - mn)
    f["COVERED_IMAGE_INTENSITY_RANGE"] = -0.0                   # montage sentinels (slideprops.cpp:29-30,74-75)
    mean = 0.0
    energy = 0.0
    for v in c:
        mean += v
        energy += float((v * v) & U32)                          # :90
    integ = mean
    mean /= n
    f["MEAN"], f["ENERGY"], f["INTEGRATED_INTENSITY"] = mean, energy, integ
    f["ROOT_MEAN_SQUARED"] = math.sqrt(energy / n)
    mad = var = 0.0
    for v in c:
        d = v - mean
        mad += abs(d)
        var += d * d
    f["MEAN_ABSOLUTE_DEVIATION"] = mad / n
    f["VARIANCE"] = var / (n - 1) if n > 1 else 0.0
    f["VARIANCE_BIASED"] = var / n if n > 1 else 0.0
    sd = math.sqrt(f["VARIANCE"])
    f["STANDARD_DEVIATION"] = sd
    f["STANDARD_DEVIATION_BIASED"] = math.sqrt(f["VARIANCE_BIASED"])
    with np.errstate(all="ignore"):
        f["COV"] = float(np.float64(sd) / np.float64(mean))
    f["STANDARD_ERROR"] = sd / math.sqrt(n)
    if mn == 0 and mx == 0:
        return f
    # TrivialHistogram::initialize (histogram.h:28-113)
    rng = mx - mn
    b100 = [0] * 101
    binw = rng / 100.0
    for v in c:
        idx = int((v - mn) / binw) if binw != 0 else 0
        b100[idx] += 1
    b100[99] += b100[100]
    b100[100] = 0
    bc = [0] * (g + 1)
    for v in c:
        bc[to_grayscale(v, mn, rng, g)] += 1
    bc[g - 1] += bc[g]
    bc[g] = 0

    def pct(q):                                                 # :214-243
        cnt = n * q
        p = 0.0
        run = 0
        for i in range(100):
            if run <= cnt <= run + b100[i]:
                with np.errstate(all="ignore"):
                    p = float(np.float64(cnt - run) * binw / np.float64(b100[i]) + mn + binw * i)
            run += b100[i]
        return p
    p01, p10, p25, p75, p90, p99 = [pct(q) for q in (0.01, 0.1, 0.25, 0.75, 0.9, 0.99)]
    ent = uni = 0.0
    for cnt in bc:
        p = cnt / n
        ent += p * math.log2(p + 2.2e-16)
        uni += p * p
    s = sorted(c)
    m = len(s)
    median = float(s[m // 2]) if m % 2 else float((s[m // 2] + s[m // 2 - 1]) & U32) / 2.0
    freq = {}
    for v in s:
        freq[v] = freq.get(v, 0) + 1
    mode, best = None, -1
    for v in sorted(freq):
        if freq[v] > best:
            mode, best = v, freq[v]
    sel = [v for v in c if p10 <= v <= p90]
    rmean = sum(sel) / len(sel) if sel else 0.0
    rmad = sum(abs(v - rmean) for v in sel) / len(sel) if sel else 0.0
    f.update(MEDIAN=median, MODE=float(mode), P01=p01, P10=p10, P25=p25, P75=p75, P90=p90, P99=p99,
             INTERQUARTILE_RANGE=p75 - p25, ROBUST_MEAN=rmean, ROBUST_MEAN_ABSOLUTE_DEVIATION=rmad, ENTROPY=-ent, UNIFORMITY=uni)
    with np.errstate(all="ignore"):
        f["QCOD"] = float(np.float64(p75 - p25) / np.float64(p75 + p25))
    f["MEDIAN_ABSOLUTE_DEVIATION"] = sum(abs(v - median) for v in c) / n
    f["UNIFORMITY_PIU"] = (1.0 - float(mx - mn) / float(mx + mn)) * 100.0
    # Moments4 (moments.h:53-109): sequential Welford updates in cloud order
    nn, mu, M2, M3, M4 = 0, 0.0, 0.0, 0.0, 0.0
    for x in c:
        n1 = nn
        nn += 1
        delta = x - mu
        dn = delta / nn
        dn2 = dn * dn
        t1 = delta * dn * n1
        mu += dn
        M4 += t1 * dn2 * (nn * nn - 3 * nn + 3) + 6 * dn2 * M2 - 4 * dn * M3
        M3 += t1 * dn * (nn - 2) - 3 * dn * M2
        M2 += t1
    if M2 != 0.0:
        f["SKEWNESS"] = math.sqrt(nn) * M3 / M2 ** 1.5 if nn > 3 else 0.0
        f["KURTOSIS"] = nn * M4 / (M2 * M2) if nn > 4 else 0.0
        f["EXCESS_KURTOSIS"] = nn * M4 / (M2 * M2) - 3 if nn > 4 else 0.0
    p5 = sum((v - mean) ** 5 for v in c)
    p6 = sum((v - mean) ** 6 for v in c)
    d5, d6 = n * sd ** 5, n * sd ** 6
    f["HYPERSKEWNESS"] = 0.0 if d5 == 0 else p5 / d5
    f["HYPERFLATNESS"] = 0.0 if d6 == 0 else p6 / d6
    return f


# --------------------------------------------------------------------------------------------------
# whole table
# --------------------------------------------------------------------------------------------------
def roi_features(tile, mask, g=64, ibsi=False, offset=1, families=("INT", "GLCM", "GLRLM", "GLDZM", "GLSZM", "GLDM", "NGLDM", "NGTDM")):
    """All enabled families of one ROI -> (dict column name -> value, dict family -> sparse matrix)."""
    roi = roi_from_tile(tile, mask)
    cols, mats = {}, {}
    ang = [0, 45, 90, 135]
    if "INT" in families:
        cols.update(intensity(roi, g))
    if "GLCM" in families:
        o, m = glcm(roi, g, ibsi, offset)
        for k, v in o.items():
            if isinstance(v, list):
                for a in range(4):
                    cols["%s_%d" % (k, ang[a])] = v[a]
            else:
                cols[k] = v
        if m is not None:
            mats["GLCM"] = sorted((a, r + 1, c + 1, int(m[a][r, c])) for a in range(4) for r in range(m[a].shape[0])
                                  for c in range(m[a].shape[1]) if m[a][r, c])
        else:
            mats["GLCM"] = []
    if "GLRLM" in families:
        o, m = glrlm(roi, g, ibsi)
        for k, v in o.items():
            if isinstance(v, list):
                for a in range(4):
                    cols["%s_%d" % (k, ang[a])] = v[a]
            else:
                cols[k] = v
        mats["GLRLM"] = sorted((a, lv, j, c) for a in range(4) for (lv, j), c in (m[a].items() if m else [])) if m else []
    for name, fn in (("GLDZM", gldzm), ("GLSZM", glszm), ("GLDM", gldm), ("NGLDM", ngldm)):
        if name in families:
            o, m = fn(roi, g, ibsi)
            cols.update(o)
            mats[name] = sorted((0, a, b, c) for (a, b), c in m.items()) if m else []
    if "NGTDM" in families:
        o, m = ngtdm(roi, g, ibsi)
        cols.update(o)
        mats["NGTDM"] = sorted((0, lv, 0, ns[0]) for lv, ns in m.items()) if m else []
        mats["NGTDM_S"] = [m[lv][1] for lv in sorted(m)] if m else []
    return cols, mats


def finite(v, sub=0.0):
    return sub if (math.isnan(v) or math.isinf(v)) else v


def check_table(inten, label, df, g=64, ibsi=False, rtol=1e-9, atol=1e-12, max_rois=8):
    """Compare a nyxus_b200 DataFrame against this restatement on the first `max_rois` ROIs (smoke check)."""
    labels = [int(v) for v in df.ROI_label.to_numpy()[:max_rois]]
    for lab in labels:
        m = label == lab
        cols, _ = roi_features(np.where(m, inten, 0), m, g, ibsi)
        row = df[df.ROI_label == lab].iloc[0]
        for k, v in cols.items():
            if k in df.columns:
                a, b = finite(v), float(row[k])
                if abs(a - b) > atol + rtol * max(abs(a), abs(b)):
                    raise AssertionError("ROI %d column %s: oracle %r vs %r" % (lab, k, a, b))
