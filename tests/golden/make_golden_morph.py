"""Golden vectors for *BASIC_MORPHOLOGY*  from the reference ITSELF: its own nyxus.py over its own compiled
pybind backend (oracle/_ref/pyref, built by oracle/Makefile.backend from the unmodified sources).
Run where /root/reference exists:  python tests/golden/make_golden_morph.py  ->  tests/golden/golden_morph.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyref          # noqa: E402
from nyxus_b200 import synth      # noqa: E402

nyxus = pyref.reference()
rng = np.random.default_rng(2026)
I, L = synth.blob_slide(160, 224, 32, 7, 13, seed=17, jitter=2)
I = I.astype(np.uint32)
L = L.copy()
base = int(L.max())
# special shapes: single pixel, two pixels, 2x2 square (all distances equal), a line, a ring, an all-zero-intensity ROI,
# a sparse ROI scattered over a large bounding box
L[2, 2] = base + 1
L[4, 200:202] = base + 2
L[150:152, 4:6] = base + 3
L[157, 100:140] = base + 4
yy, xx = np.mgrid[0:160, 0:224]
ring = (np.hypot(yy - 144, xx - 180) < 12) & (np.hypot(yy - 144, xx - 180) > 9) & (L == 0)
L[ring] = base + 5
L[120:126, 210:220] = base + 6
I[L == base + 6] = 0
sp = (rng.random((160, 224)) < 0.004) & (L == 0)
L[sp] = base + 7
feats = ["*BASIC_MORPHOLOGY*"]
df = nyxus.Nyxus(feats).featurize(I.copy(), L.copy())
allf = ["*ALL_INTENSITY*", "*BASIC_MORPHOLOGY*", "*ALL_GLCM*", "*ALL_GLRLM*", "*ALL_GLDZM*", "*ALL_GLSZM*", "*ALL_GLDM*", "*ALL_NGLDM*", "*ALL_NGTDM*"]
df_all = nyxus.Nyxus(allf).featurize(I.copy(), L.copy())
np.savez_compressed(os.path.join(HERE, "golden_morph.npz"), inten=I, label=L, labels=df.ROI_label.to_numpy(np.int64),
                    columns=np.array(list(df.columns[4:])), table=df.iloc[:, 4:].to_numpy(np.float64),
                    all_columns=np.array(list(df_all.columns[4:])), all_table=df_all.iloc[:, 4:].to_numpy(np.float64))
print(df.shape, df_all.shape)
