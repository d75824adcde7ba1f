"""Generates tests/golden/golden_radiomics.npz: golden vectors of the COMPILED REFERENCE (oracle/_ref, built from
/root/reference by oracle/Makefile.ref) for the settings round 1 did not cover:
  * PyRadiomics binning (negative grey depth, reference features/texture_feature.h:106-118) as the 2D grey depth
    (GLCM, GLRLM, GLSZM, GLDM, NGTDM, first-order; GLDZM throws and NGLDM is undefined there) and as the 3D per-family
    depths "3glcm/greydepth", "3glszm/greydepth", "3ngtdm/greydepth" (env_metaparams.cpp:109-233), incl. depth 0;
  * "glcm/greydepth" different from the general grey depth (2D: feeds only the GLCM blank-ROI test, glcm.cpp:22-28).
Run in the container that has /root/reference:  python tests/golden/make_golden_rad.py"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref
from nyxus_b200 import synth

RAD_FAMS = ["GLCM", "GLRLM", "GLSZM", "GLDM", "NGTDM"]
RAD_MASK = sum(ref.FAM[f] for f in RAD_FAMS) | ref.FAM["INT"]


def roi_cases():
    rng = np.random.default_rng(20261018)
    out = []
    for k in range(12):
        h, w = int(rng.integers(2, 26)), int(rng.integers(2, 26))
        kind = k % 4
        if kind == 0:
            t = rng.integers(1, 60000, (h, w))
        elif kind == 1:
            t = rng.integers(1, 7, (h, w))
        elif kind == 2:
            yy, xx = np.mgrid[0:h, 0:w]
            t = (2000 + 900 * np.sin(xx / 2.5) + 600 * np.cos(yy / 3.0) + rng.normal(0, 40, (h, w))).clip(1, 65535).astype(int)
        else:
            t = rng.integers(0, 200, (h, w))
        m = (rng.random((h, w)) < rng.choice([0.6, 0.85, 1.0])).astype(int)
        m[0, 0] = 1
        out.append(("rad_%02d" % k, t, m, -int([4, 8, 20, 64][k % 4])))
    out.append(("rad_constant", np.full((4, 5), 9), np.ones((4, 5), int), -8))
    out.append(("rad_single_bin", rng.integers(1, 500, (6, 6)), np.ones((6, 6), int), -1))
    return out


def main():
    cols = ref.columns()
    data = {"columns": np.array(cols), "rad_families": np.array(RAD_FAMS)}
    names = []
    for name, tile, mask, g in roi_cases():
        tile = np.asarray(tile, np.uint32)
        mask = np.asarray(mask, np.uint8)
        names.append(name)
        data[name + "/tile"] = tile
        data[name + "/mask"] = mask
        data[name + "/g"] = np.array([g])
        data[name + "/features"] = ref.roi_features(tile, mask, g, False, fam_mask=RAD_MASK)
        mats = ref.roi_matrices(tile * mask, mask, g, False, fam_mask=RAD_MASK & ~1)
        for fam in RAD_FAMS:
            data[name + "/" + fam] = np.array(sorted(mats[fam]), np.int64).reshape(-1, 4)
        data[name + "/NGTDM_S"] = np.array(mats["NGTDM_S"], np.float64)
    data["names"] = np.array(names)

    # "glcm/greydepth" != grey depth on a small slide (whole path replay): the GLCM columns of ROIs whose blank test flips
    # change, nothing else does
    rng = np.random.default_rng(4242)
    inten, label = synth.blob_slide(96, 128, 32, 7, 12, seed=5, jitter=2)
    inten = inten.astype(np.uint32)
    flat = label == np.unique(label)[2]
    inten[flat] = 500                                      # one constant ROI
    two = label == np.unique(label)[3]
    inten[two] = np.where(rng.random(int(two.sum())) < 0.5, 1000, 1001)   # a ROI whose extremes share a MATLAB bin at depth 64
    data["gg_inten"], data["gg_label"] = inten, label
    gg_cases = [(64, None), (64, 1), (64, 4), (64, 100000), (64, -1), (64, -5), (16, 64)]
    for k, (g, gg) in enumerate(gg_cases):
        ref.set_glcm_greydepth(gg)
        labs, tab, _ = ref.featurize_2d(inten, label, g, False)
        data["gg%d/params" % k] = np.array([g, -2147483648 if gg is None else gg], np.int64)
        data["gg%d/labels" % k] = labs
        data["gg%d/table" % k] = tab
    ref.set_glcm_greydepth(None)
    data["gg_n"] = np.array([len(gg_cases)])

    # 3D: per-family depths of either sign (and 0) on the golden volume of make_golden3d.py, quantised so that depth 0 is legal
    vol, vlabel = synth.ellipsoid_volume(26, 39, 52, 13, 3, 5, seed=41)
    q = np.minimum(vol // 40, 60).astype(np.uint32)
    data["vol"], data["vol_q"], data["vol_label"] = vol, q, vlabel
    data["columns3d"] = np.array(ref.columns_3d())
    v_cases = [dict(which="vol", greydepth=64, glcm_greydepth=-20, glszm_greydepth=-20, ngtdm_greydepth=-20, ngtdm_radius=1),
               dict(which="vol", greydepth=64, glcm_greydepth=-8, glszm_greydepth=16, ngtdm_greydepth=-64, ngtdm_radius=1),
               dict(which="vol_q", greydepth=64, glcm_greydepth=0, glszm_greydepth=-4, ngtdm_greydepth=0, ngtdm_radius=1),
               dict(which="vol_q", greydepth=64, glcm_greydepth=-100, glszm_greydepth=0, ngtdm_greydepth=-7, ngtdm_radius=0)]
    for k, kw in enumerate(v_cases):
        kw = dict(kw)
        which = kw.pop("which")
        labs, tab, _ = ref.featurize_3d(data[which].astype(np.uint32), vlabel, ibsi=False, **kw)
        data["v%d/params" % k] = np.array([kw["greydepth"], kw["glcm_greydepth"], kw["glszm_greydepth"], kw["ngtdm_greydepth"],
                                           kw["ngtdm_radius"], 0 if which == "vol" else 1], np.int64)
        data["v%d/labels" % k] = labs
        data["v%d/table" % k] = tab
    data["v_n"] = np.array([len(v_cases)])
    np.savez_compressed(os.path.join(HERE, "golden_radiomics.npz"), **data)
    print("wrote %d ROI cases, %d glcm/greydepth slides, %d volumes" % (len(names), len(gg_cases), len(v_cases)))


if __name__ == "__main__":
    main()
