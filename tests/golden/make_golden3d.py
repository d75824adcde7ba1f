"""Generates tests/golden/golden_vol.npz: a small labelled volume and the feature tables the COMPILED REFERENCE
(oracle/_ref, built from /root/reference by oracle/Makefile.ref) produces for it under several settings.
Run in the container that has /root/reference:  python tests/golden/make_golden3d.py"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref
from nyxus_b200 import synth

inten, label = synth.ellipsoid_volume(26, 39, 52, 13, 3, 5, seed=41)
# two hand-made ROIs: a constant one and a 1-voxel-thick slab touching the volume border
label[0:1, 0:6, 0:9] = 900
inten[label == 900] = np.random.default_rng(5).integers(1, 4000, int((label == 900).sum()))
label[20:25, 30:36, 44:50] = 901
inten[label == 901] = 777
out = {"inten": inten, "label": label, "columns": np.array(ref.columns_3d())}
cases = {
    "matlab64_r1": dict(greydepth=64, glszm_greydepth=64, ngtdm_greydepth=64, ngtdm_radius=1, ibsi=False),
    "defaults": dict(greydepth=64, glszm_greydepth=0, ngtdm_greydepth=0, ngtdm_radius=0, ibsi=False),
    "g16_r1": dict(greydepth=16, glszm_greydepth=8, ngtdm_greydepth=32, ngtdm_radius=1, ibsi=False),
}
for name, kw in cases.items():
    labs, tab, _ = ref.featurize_3d(inten.astype(np.uint32), label, **kw)
    out["labels_" + name] = labs
    out["table_" + name] = tab
    out["params_" + name] = np.array([kw["greydepth"], kw["glszm_greydepth"], kw["ngtdm_greydepth"], kw["ngtdm_radius"], int(kw["ibsi"])])
q = np.minimum(inten // 40, 60).astype(np.uint16)      # <= 61 levels for the IBSI (no binning) run
kw = dict(greydepth=64, glszm_greydepth=64, ngtdm_greydepth=64, ngtdm_radius=1, ibsi=True)
labs, tab, _ = ref.featurize_3d(q.astype(np.uint32), label, **kw)
out["inten_ibsi"] = q
out["labels_ibsi_r1"] = labs
out["table_ibsi_r1"] = tab
out["params_ibsi_r1"] = np.array([64, 64, 64, 1, 1])
np.savez_compressed(os.path.join(HERE, "golden_vol.npz"), **out)
print({k: v.shape for k, v in out.items()})
