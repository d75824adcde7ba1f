"""Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference (oracle/_ref, compiled from
/root/reference by oracle/Makefile.ref) on fixed seeded inputs.  Run in the build container:
    python tests/golden/make_golden.py
Outputs: golden_rois.npz  (per-ROI cases: tile, mask, settings -> 337 raw feature values + sparse matrices)
         golden_slide.npz (one small multi-ROI slide -> labels + output table after force_finite_number)"""
import os, sys, json
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from oracle import ref
import reference_fixtures as rf


def cases():
    rng = np.random.default_rng(20261017)
    out = []
    # the reference's own fixtures
    for z in range(4):
        out.append(("ibsi_phantom_z%d_ibsi" % (z + 1), rf.IBSI_PHANTOM_INT[z], rf.IBSI_PHANTOM_MASK[z], 64, True))
        out.append(("ibsi_phantom_z%d_matlab" % (z + 1), rf.IBSI_PHANTOM_INT[z], rf.IBSI_PHANTOM_MASK[z], 64, False))
    out.append(("ngldm_ibsi_fig", rf.NGLDM_IBSI_IMAGE, np.ones((4, 4), int), 64, True))
    out.append(("ngldm_nonibsi", rf.NGLDM_NONIBSI_IMAGE, np.ones((5, 5), int), 64, False))
    out.append(("gldzm_ibsi_fig", rf.GLDZM_IBSI_IMAGE, np.ones((4, 4), int), 64, True))
    # edge cases
    out.append(("single_pixel", np.array([[7]]), np.array([[1]]), 64, False))
    out.append(("constant_roi", np.full((5, 6), 9), np.ones((5, 6), int), 64, False))
    out.append(("two_levels", np.array([[1, 1, 2], [2, 2, 1], [1, 2, 2]]), np.ones((3, 3), int), 64, False))
    out.append(("row_vector", rng.integers(1, 300, (1, 17)), np.ones((1, 17), int), 64, False))
    out.append(("col_vector", rng.integers(1, 300, (13, 1)), np.ones((13, 1), int), 64, False))
    m = np.ones((9, 9), int); m[3:6, 3:6] = 0
    out.append(("ring_with_hole", rng.integers(1, 1000, (9, 9)), m, 64, False))
    t = rng.integers(1, 4000, (12, 10)); t[rng.random((12, 10)) < 0.15] = 0
    out.append(("zeros_inside", t, np.ones((12, 10), int), 64, False))
    out.append(("zeros_inside_ibsi", np.where(t > 0, t % 17, 0), np.ones((12, 10), int), 64, True))
    out.append(("all_zero_roi", np.zeros((4, 4), int), np.ones((4, 4), int), 64, False))
    # random ROIs, assorted statistics / grey depths / modes
    for k in range(14):
        h, w = int(rng.integers(2, 24)), int(rng.integers(2, 24))
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
        if not m.any():
            m[0, 0] = 1
        ibsi = kind in (1, 3) and k % 2 == 1
        g = [64, 16, 256, 64][k % 4] if not ibsi else 64
        out.append(("random_%02d" % k, t, m, g, ibsi))
    return out


def main():
    cols = ref.columns()
    data = {"columns": np.array(cols)}
    names = []
    for name, tile, mask, g, ibsi in cases():
        tile = np.asarray(tile, np.uint32)
        mask = np.asarray(mask, np.uint8)
        feats = ref.roi_features(tile, mask, g, ibsi)
        mats = ref.roi_matrices(tile * mask, mask, g, ibsi)
        names.append(name)
        data[name + "/tile"] = tile
        data[name + "/mask"] = mask
        data[name + "/settings"] = np.array([g, int(ibsi)])
        data[name + "/features"] = feats
        for fam in ("GLCM", "GLRLM", "GLDZM", "GLSZM", "GLDM", "NGLDM", "NGTDM"):
            data[name + "/" + fam] = np.array(sorted(mats[fam]), np.int64).reshape(-1, 4)
        data[name + "/NGTDM_S"] = np.array(mats["NGTDM_S"], np.float64)
    data["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "golden_rois.npz"), **data)

    # a small slide: 10 ROIs with sparse label values, touching borders, uint16-range intensities
    rng = np.random.default_rng(777)
    H, W = 72, 96
    inten = rng.integers(1, 9000, (H, W)).astype(np.uint32)
    yy, xx = np.mgrid[0:H, 0:W]
    inten = (inten * 0.3 + 3000 + 2000 * np.sin(xx / 5.0) * np.cos(yy / 7.0)).clip(1, 65535).astype(np.uint32)
    inten[rng.random((H, W)) < 0.01] = 0
    label = np.zeros((H, W), np.uint32)
    k = 0
    for (y, x, h, w) in [(0, 0, 10, 12), (0, 80, 15, 16), (20, 30, 21, 19), (50, 5, 22, 25), (45, 60, 27, 30), (15, 5, 3, 3),
                         (30, 70, 1, 9), (12, 55, 8, 1), (66, 40, 6, 40), (25, 52, 12, 9)]:
        k += 1
        m = rng.random((h, w)) < 0.9
        sub = label[y:y + h, x:x + w]
        sub[m & (sub == 0)] = k * 37 + 5
    labels, table, _ = ref.featurize_2d(inten, label, 64, False)
    labels_i, table_i, _ = ref.featurize_2d(np.minimum(inten, 255), label, 64, True)
    np.savez_compressed(os.path.join(HERE, "golden_slide.npz"), inten=inten, label=label, labels=labels, table=table,
                        labels_ibsi=labels_i, table_ibsi=table_i, columns=np.array(cols))
    print("wrote %d ROI cases, slide with %d ROIs" % (len(names), len(labels)))


if __name__ == "__main__":
    main()
