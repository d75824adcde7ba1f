"""Golden vectors for the unweighted 2D geometric moments (SPAT_MOMENT_*, CENTRAL_MOMENT_*, NORM_SPAT_MOMENT_*,
NORM_CENTRAL_MOMENT_*, HU_M*, IMOM_RM_* ... IMOM_HU*) from the reference ITSELF (oracle/_ref/pyref: its own nyxus.py over
its own compiled pybind backend).  Run where /root/reference exists: python tests/golden/make_golden_geomom.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyref                              # noqa: E402
from nyxus_b200 import synth, feature_table as ft    # noqa: E402

nyxus = pyref.reference()
names = [f[0] for f in ft.FEATURES if f[2] == "GEOM"]
I, L = synth.blob_slide(160, 224, 32, 7, 13, seed=29, jitter=2)
I = I.astype(np.uint32)
base = int(L.max())
L[150:156, 4:30] = base + 1           # a bar
L[140:149, 180:189] = base + 2        # a square
I[140:149, 180:189] = 700             # ... of constant intensity
keep = nyxus.Nyxus(names)
df = keep.featurize(I.copy(), L.copy())
np.savez_compressed(os.path.join(HERE, "golden_geomom.npz"), inten=I, label=L, labels=df.ROI_label.to_numpy(np.int64),
                    columns=np.array(list(df.columns[4:])), table=df.iloc[:, 4:].to_numpy(np.float64))
print(df.shape)
