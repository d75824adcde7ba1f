"""Golden vectors for the 3D texture families beyond GLCM / GLSZM / NGTDM from the reference ITSELF: its own
nyxus.py over its own compiled pybind backend (oracle/_ref/pyref), Nyxus3D.featurize_files on a NIfTI volume written here.
Run where /root/reference exists: python tests/golden/make_golden3d_more.py  ->  tests/golden/golden_vol_more.npz"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyref              # noqa: E402
from nyxus_b200 import nifti          # noqa: E402

nyxus = pyref.reference()
rng = np.random.default_rng(2027)
D, H, W = 16, 22, 26
zz, yy, xx = np.mgrid[0:D, 0:H, 0:W]
vol = ((np.sin(zz / 3.0) + np.cos(yy / 4.0) * np.sin(xx / 3.5) + 2.2) * 250.0 + rng.normal(0, 40, (D, H, W))).clip(1, 3000).astype(np.uint16)
vol[rng.random((D, H, W)) < 0.01] = 0
lab = np.zeros((D, H, W), np.uint16)
lab[(((zz - 7) / 6.5) ** 2 + ((yy - 10) / 9.0) ** 2 + ((xx - 12) / 10.0) ** 2) < 1] = 3
lab[12:16, 1:8, 1:9] = 8
lab[0:3, 15:22, 18:26] = 5
lab[rng.random((D, H, W)) < 0.05] = 0
tmp = tempfile.mkdtemp()
fi, fm = os.path.join(tmp, "v.nii"), os.path.join(tmp, "m.nii")
nifti.write(fi, vol)
nifti.write(fm, lab)
out = {"vol": vol, "label": lab}
keep = []
FAMS = {"GLDM": ("*3D_GLDM*", "3gldm/greydepth")}
for fam, (group, key) in FAMS.items():
    for g in (0, 16, -12):
        n = nyxus.Nyxus3D([group])
        keep.append(n)
        if g:
            n.set_metaparam("%s=%d" % (key, g))
        df = n.featurize_files([fi], [fm], False)
        out["%s/columns" % fam] = np.array(list(df.columns[4:]))
        out["%s/labels" % fam] = df.ROI_label.to_numpy(np.int64)
        out["%s/table_%d" % (fam, g)] = df.iloc[:, 4:].to_numpy(np.float64)
np.savez_compressed(os.path.join(HERE, "golden_vol_more.npz"), **out)
print({k: getattr(v, "shape", None) for k, v in out.items()})
