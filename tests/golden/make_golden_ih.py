"""Golden vectors for the IBSI intensity-histogram family *ALL_IH*  from the reference ITSELF: its own nyxus.py
over its own compiled pybind backend (oracle/_ref/pyref).  Run where /root/reference exists:
python tests/golden/make_golden_ih.py  ->  tests/golden/golden_ih.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyref          # noqa: E402
from nyxus_b200 import synth      # noqa: E402

nyxus = pyref.reference()
rng = np.random.default_rng(77)
I, L = synth.blob_slide(160, 224, 32, 7, 13, seed=23, jitter=2)
I = I.astype(np.uint32)
base = int(L.max())
L[2:6, 2:8] = base + 1
I[2:6, 2:8] = 500                                   # constant ROI: max <= min -> every IH value is `noval`
L[150:156, 4:30] = base + 2
I[150:156, 4:30] = rng.integers(10, 14, (6, 26))    # fewer distinct values than bins
L[140:148, 180:220] = base + 3
I[140:148, 180:220] = rng.integers(0, 3, (8, 40)) * 1000     # three spikes (min 0)
out = {"inten": I, "label": L}
keep = []        # (the instances are kept alive: the reference keys its environments on id(self), see nyxus.py:146-155)
for g in (64, 20, 7):
    keep.append(nyxus.Nyxus(["*ALL_IH*"], ibsi=True, coarse_gray_depth=g))
    print("g", g, flush=True)
    df = keep[-1].featurize(I.copy(), L.copy())
    out["table_%d" % g] = df.iloc[:, 4:].to_numpy(np.float64)
    out["labels"] = df.ROI_label.to_numpy(np.int64)
    out["columns"] = np.array(list(df.columns[4:]))
# IBSI off: the family is stripped (env_features.cpp:513-527)
keep.append(nyxus.Nyxus(["*ALL_IH*", "MEAN"], ibsi=False))
print("ibsi off", flush=True)
df0 = keep[-1].featurize(I.copy(), L.copy())
out["columns_ibsi_off"] = np.array(list(df0.columns[4:]))
keep.append(nyxus.Nyxus(["*ALL_IH*", "*ALL_INTENSITY*", "*BASIC_MORPHOLOGY*", "GLCM_ASM", "NGTDM_STRENGTH"], ibsi=True))
print("mixed", flush=True)
Ip, Lp = synth.blob_slide(96, 128, 32, 7, 12, seed=5, jitter=1)          # (the plain slide: the column ORDER is what is recorded)
dfa = keep[-1].featurize((Ip.astype(np.uint32) >> 6), Lp.copy())
out["mixed_columns"] = np.array(list(dfa.columns[4:]))
np.savez_compressed(os.path.join(HERE, "golden_ih.npz"), **out)
print({k: getattr(v, "shape", None) for k, v in out.items()}, list(out["columns_ibsi_off"]))
