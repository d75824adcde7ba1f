"""Dev script (run under gpurun): full-table and matrix parity of the CUDA path against oracle/_ref on small slides."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nyxus_b200 import Engine
from oracle import ref
from nyxus_b200 import synth

ALL = ["*ALL_INTENSITY*", "*ALL_TEXTURE*"]


def compare(inten, label, g=64, ibsi=False, name=""):
    eng = Engine(ALL, coarse_gray_depth=g, ibsi=ibsi)
    t0 = time.time()
    tab, img = eng.featurize(inten, label)
    t1 = time.time()
    rl, rt, _ = ref.featurize_2d(inten, label, greydepth=g, ibsi=ibsi, nthreads=8)
    cols = eng.columns
    assert cols == ref.columns(), "column layout differs"
    assert np.array_equal(tab[:, 0].astype(np.int64), rl.astype(np.int64)), "labels differ"
    a, b = rt, tab[:, 2:]
    diff = np.abs(a - b)
    rel = diff / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-300)
    bad = (rel > 1e-9) & (diff > 1e-12)
    print("%s: %d rois, %d cols, gpu %.3fs, bad entries %d, max rel %.3g" % (name, len(rl), len(cols), t1 - t0, bad.sum(), rel[diff > 1e-12].max() if (diff > 1e-12).any() else 0))
    if bad.any():
        js = np.unique(np.nonzero(bad)[1])
        for j in js[:25]:
            i = np.nonzero(bad[:, j])[0][0]
            print("   ", cols[j], "roi", rl[i], "ref", a[i, j], "gpu", b[i, j], "nbad", bad[:, j].sum())
    print("   timings", {k: round(v, 3) for k, v in eng.timings().items() if v > 0}, eng.counts())
    # phase-1 records
    recs = eng.debug_rois()
    for lab in rl[:3]:
        m = label == lab
        ys, xs = np.nonzero(m)
        r = recs[recs[:, 0] == lab][0]
        exp = [lab, m.sum(), inten[m].min(), inten[m].max(), xs.min(), ys.min(), xs.max() - xs.min() + 1, ys.max() - ys.min() + 1]
        assert list(r) == [int(v) for v in exp], (list(r), exp)
        cloud = eng.debug_cloud(lab)
        assert np.array_equal(cloud, inten[m].astype(np.uint32)), "cloud differs"
    # matrices of a few ROIs
    nm = 0
    for lab in rl[:: max(1, len(rl) // 6)][:6]:
        r = recs[recs[:, 0] == lab][0]
        x0, y0, w, h = int(r[4]), int(r[5]), int(r[6]), int(r[7])
        tile = inten[y0:y0 + h, x0:x0 + w].astype(np.uint32)
        mask = (label[y0:y0 + h, x0:x0 + w] == lab)
        rm = ref.roi_matrices(tile * mask, mask, g, ibsi)
        for fam in ["GLCM", "GLRLM", "GLDZM", "GLSZM", "GLDM", "NGLDM", "NGTDM"]:
            gt, gs = eng.debug_matrix(lab, fam)
            if sorted(gt) != sorted(rm[fam]):
                print("   MATRIX MISMATCH", fam, "roi", lab, len(gt), len(rm[fam]))
                nm += 1
            elif fam == "NGTDM":
                s_ref = np.array(rm["NGTDM_S"]); s_gpu = np.array(gs)
                if len(s_ref) and np.abs(s_ref - s_gpu).max() > 1e-9 * max(1, np.abs(s_ref).max()):
                    print("   NGTDM S mismatch", lab)
                    nm += 1
    print("   matrix mismatches:", nm)
    eng.close()
    return int(bad.sum()) + nm


if __name__ == "__main__":
    tot = 0
    I, L = synth.random_rois(256, 300, 40, 3)
    tot += compare(I, L, name="random uint32")
    I, L = synth.random_rois(200, 200, 30, 4, levels=200)
    tot += compare(I, L, ibsi=True, name="random ibsi 200 levels")
    tot += compare(I, L, g=16, name="random g=16")
    I, L = synth.config1()
    tot += compare(I, L, name="C1 uint16")
    print("TOTAL BAD", tot)
