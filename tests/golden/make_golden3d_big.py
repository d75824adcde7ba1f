"""Golden vectors for the 3D families of SURVEY 8f-5 (GLDM, GLRLM, GLDZM, NGLDM) at sizes that leave the shared-memory classes:
one organ-sized ROI whose bounding cube exceeds 40 000 voxels (the global-memory path, k_roi_family3_big) next to medium and
small ones, from the reference ITSELF (its own nyxus.py over its own compiled pybind backend, oracle/_ref/pyref).
Run where /root/reference exists: python tests/golden/make_golden3d_big.py  ->  tests/golden/golden_vol_big.npz
(the volume is regenerated from the seed by the tests: only the tables are stored)"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyref              # noqa: E402
from nyxus_b200 import nifti          # noqa: E402
import nyx_testutil as tu             # noqa: E402

nyxus = pyref.reference()
vol, lab = tu.big_volume_3d()
tmp = tempfile.mkdtemp()
fi, fm = os.path.join(tmp, "v.nii"), os.path.join(tmp, "m.nii")
nifti.write(fi, vol)
nifti.write(fm, lab)
out, keep = {}, []
for fam, group, key in (("GLDM", "*3D_GLDM*", "3gldm/greydepth"), ("GLRLM", "*3D_GLRLM*", "3glrlm/greydepth")):
    for g in (0, 32, -20):
        n = nyxus.Nyxus3D([group])
        keep.append(n)
        if g:
            n.set_metaparam("%s=%d" % (key, g))
        df = n.featurize_files([fi], [fm], False)
        out["%s/labels" % fam] = df.ROI_label.to_numpy(np.int64)
        out["%s/table_%d" % (fam, g)] = df.iloc[:, 4:].to_numpy(np.float64)
for fam, group in (("GLDZM", "*3D_GLDZM*"), ("NGLDM", "*3D_NGLDM*")):
    for tag, kw in (("64", {}), ("ibsi", {"ibsi": True})):
        n = nyxus.Nyxus3D([group], **kw)
        keep.append(n)
        df = n.featurize_files([fi], [fm], False)
        out["%s/labels" % fam] = df.ROI_label.to_numpy(np.int64)
        out["%s/table_%s" % (fam, tag)] = df.iloc[:, 4:].to_numpy(np.float64)
np.savez_compressed(os.path.join(HERE, "golden_vol_big.npz"), **out)
print({k: getattr(v, "shape", None) for k, v in out.items()})
