"""Golden vectors for the opt-in multi-valued HISTOGRAM feature (HISTOGRAM_BIN_0 .. _N-1, N = coarse grey depth; featureset.h,
intensity.cpp:129, output_2_buffer.cpp:158-168,283-293) from the reference ITSELF: its own nyxus.py over its own compiled
pybind backend (oracle/_ref/pyref).  The slide is the one of golden_ih.npz with its constant ROI blanked (all-zero
intensities: the reference pads the bins with zeros).  Run where /root/reference exists:
python tests/golden/make_golden_hist.py  ->  tests/golden/golden_hist.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyref          # noqa: E402

nyxus = pyref.reference()
Z = np.load(os.path.join(HERE, "golden_ih.npz"))
I, L = Z["inten"].copy(), Z["label"].copy()
I[2:6, 2:8] = 0                                      # the constant ROI becomes a blank one
out = {"blank_label": np.int64(L[3, 3])}
keep = []
for g, ibsi in ((64, False), (20, False), (7, False), (20, True)):
    keep.append(nyxus.Nyxus(["HISTOGRAM", "MEAN", "ENTROPY"], ibsi=ibsi, coarse_gray_depth=g))
    df = keep[-1].featurize(I.copy(), L.copy())
    tag = "%d%s" % (g, "_ibsi" if ibsi else "")
    out["table_" + tag] = df.iloc[:, 4:].to_numpy(np.float64)
    out["columns_" + tag] = np.array(list(df.columns[4:]))
    out["labels"] = df.ROI_label.to_numpy(np.int64)
    print(tag, df.shape, list(df.columns[4:8]), flush=True)
# column ORDER next to the neighbouring codes (weighted moments before, the IH family after)
keep.append(nyxus.Nyxus(["IH_MEAN_VAL", "HISTOGRAM", "IMOM_WHU7", "GLCM_ASM", "MEAN"], ibsi=True, coarse_gray_depth=5))
dfm = keep[-1].featurize(I.copy(), L.copy())
out["mixed_columns"] = np.array(list(dfm.columns[4:]))
print(list(dfm.columns[4:]))
np.savez_compressed(os.path.join(HERE, "golden_hist.npz"), **out)
