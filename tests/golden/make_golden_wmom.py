"""Golden vectors for the distance-to-contour WEIGHTED 2D geometric moments (WEIGHTED_SPAT_MOMENT_*,
WEIGHTED_CENTRAL_MOMENT_*, WT_NORM_CTR_MOM_*, WEIGHTED_HU_M*, IMOM_WRM_* ... IMOM_WHU*) and for the contour features that pin
the reference's tracer (PERIMETER: order of the contour; EDGE_*: its pixel set) from the reference ITSELF (oracle/_ref/pyref:
its own nyxus.py over its own compiled pybind backend).  Run where /root/reference exists: python tests/golden/make_golden_wmom.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyref                              # noqa: E402
import nyx_testutil as tu                             # noqa: E402

nyxus = pyref.reference()
PQ10 = ["00", "01", "02", "03", "10", "11", "12", "20", "21", "30"]
PQ7 = ["02", "03", "11", "12", "20", "21", "30"]
names = (["WEIGHTED_SPAT_MOMENT_" + s for s in PQ10] + ["WEIGHTED_CENTRAL_MOMENT_" + s for s in PQ7] + ["WT_NORM_CTR_MOM_" + s for s in PQ7]
         + ["WEIGHTED_HU_M%d" % k for k in range(1, 8)]
         + ["IMOM_WRM_" + s for s in PQ10] + ["IMOM_WCM_" + s for s in PQ7] + ["IMOM_WNCM_" + s for s in PQ7] + ["IMOM_WHU%d" % k for k in range(1, 8)])
diag = ["PERIMETER", "EDGE_INTEGRATED_INTENSITY", "EDGE_MEAN_INTENSITY", "EDGE_MIN_INTENSITY", "EDGE_MAX_INTENSITY"]
I, L = tu.wmom_slide()
keep = nyxus.Nyxus(names + diag)
df = keep.featurize(I.copy(), L.copy())
np.savez_compressed(os.path.join(HERE, "golden_wmom.npz"), labels=df.ROI_label.to_numpy(np.int64),
                    columns=np.array(list(df.columns[4:])), table=df.iloc[:, 4:].to_numpy(np.float64))
print(df.shape, list(df.columns[4:8]))

# ---- one ROI beyond the 40 000-cell shared-memory classes (the CTA-per-ROI launch) with holes and a ragged outline, and a
# thin diagonal band: tests/golden/golden_wmom_big.npz (the slide is regenerated from the seed: tu.wmom_slide_big)
Ib, Lb = tu.wmom_slide_big()
keepb = nyxus.Nyxus(names + diag)
dfb = keepb.featurize(Ib.copy(), Lb.copy())
np.savez_compressed(os.path.join(HERE, "golden_wmom_big.npz"), labels=dfb.ROI_label.to_numpy(np.int64),
                    columns=np.array(list(dfb.columns[4:])), table=dfb.iloc[:, 4:].to_numpy(np.float64))
print(dfb.shape)
