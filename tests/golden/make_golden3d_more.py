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
FAMS = {"GLDM": ("*3D_GLDM*", "3gldm/greydepth"), "GLRLM": ("*3D_GLRLM*", "3glrlm/greydepth")}
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
# families binned by the coarse grey depth of the constructor (3d_gldzm.cpp:65-70, 3d_ngldm.cpp:204): tag = depth, "ibsi" = IBSI mode
for fam, group in (("GLDZM", "*3D_GLDZM*"), ("NGLDM", "*3D_NGLDM*")):
    for tag, kw in (("64", {}), ("20", {"coarse_gray_depth": 20}), ("ibsi", {"ibsi": True})):
        n = nyxus.Nyxus3D([group], **kw)
        keep.append(n)
        df = n.featurize_files([fi], [fm], False)
        out["%s/columns" % fam] = np.array(list(df.columns[4:]))
        out["%s/labels" % fam] = df.ROI_label.to_numpy(np.int64)
        out["%s/table_%s" % (fam, tag)] = df.iloc[:, 4:].to_numpy(np.float64)
# ---- edge cases (keys "e/..."): a one-slice slab and a two-voxel-thick plate (no interior voxel for NGLDM), a constant ROI
# (blank: soft NaN), a single voxel, a ROI holding intensity-0 voxels, few grey levels with long runs
D2, H2, W2 = 9, 14, 17
vol2 = rng.integers(1, 9, (D2, H2, W2)).astype(np.uint16) * 37
lab2 = np.zeros((D2, H2, W2), np.uint16)
lab2[0, 1:8, 1:9] = 2                   # slab, d = 1
lab2[2:4, 8:14, 0:7] = 4                # plate, d = 2
lab2[5:9, 0:5, 10:17] = 6               # constant ROI
vol2[lab2 == 6] = 500
lab2[8, 13, 0] = 7                      # single voxel
lab2[4:9, 6:13, 1:9] = 9                # zeros inside + runs
vol2[5:7, 7:12, 2:8] = 111
vol2[7, 8:10, 3:6] = 0
fi2, fm2 = os.path.join(tmp, "v2.nii"), os.path.join(tmp, "m2.nii")
nifti.write(fi2, vol2)
nifti.write(fm2, lab2)
out["e/vol"], out["e/label"] = vol2, lab2
for fam, (group, key) in FAMS.items():
    for g in (0, 5, -4):
        n = nyxus.Nyxus3D([group])
        keep.append(n)
        if g:
            n.set_metaparam("%s=%d" % (key, g))
        df = n.featurize_files([fi2], [fm2], False)
        out["e/%s/labels" % fam] = df.ROI_label.to_numpy(np.int64)
        out["e/%s/table_%d" % (fam, g)] = df.iloc[:, 4:].to_numpy(np.float64)
for fam, group in (("GLDZM", "*3D_GLDZM*"), ("NGLDM", "*3D_NGLDM*")):
    for tag, kw in (("64", {}), ("5", {"coarse_gray_depth": 5}), ("ibsi", {"ibsi": True})):
        n = nyxus.Nyxus3D([group], **kw)
        keep.append(n)
        df = n.featurize_files([fi2], [fm2], False)
        out["e/%s/labels" % fam] = df.ROI_label.to_numpy(np.int64)
        out["e/%s/table_%s" % (fam, tag)] = df.iloc[:, 4:].to_numpy(np.float64)
np.savez_compressed(os.path.join(HERE, "golden_vol_more.npz"), **out)
print({k: getattr(v, "shape", None) for k, v in out.items()})
