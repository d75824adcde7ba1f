"""Small 2D slide + multi-slide list + giant ROI + 3D volume through the C ABI -- the command wrapped by compute-sanitizer
(memcheck / racecheck / synccheck).  usage: compute-sanitizer --tool memcheck python tools/sanitize_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nyxus_b200 import Engine, synth

I, L = synth.blob_slide(248, 312, 31, 7, 12, seed=5)
eng = Engine(["*ALL_TEXTURE*", "*ALL_INTENSITY*"])
t1, _ = eng.featurize(I, L)
t2, _ = eng.featurize(np.stack([I, I[::-1].copy()]), np.stack([L, L[::-1].copy()]).astype(np.uint32))
g = np.zeros((260, 300), np.uint32); g[5:250, 7:290] = 3          # AABB > 40000 cells: giant path
gi = np.random.default_rng(1).integers(1, 4000, g.shape).astype(np.uint16)
t3, _ = eng.featurize(gi, g)
Io = np.random.default_rng(2).integers(0, 300, (45, 67)).astype(np.uint32)  # odd width: scalar scan path, uint32 intensities
Lo = np.random.default_rng(3).integers(0, 5, (45, 67)).astype(np.uint32)
t4, _ = eng.featurize(Io, Lo)
eng.close()
V, LV = synth.ellipsoid_volume(40, 48, 56, 20, 5, 8, seed=6)
e3 = Engine(["*3D_ALL_INTENSITY*", "*3D_GLCM*", "*3D_GLSZM*", "*3D_NGTDM*"], dim=3)
for k, v in (("3glszm/greydepth", 64), ("3ngtdm/greydepth", 64), ("3ngtdm/radius", 1)):
    e3.set_param(k, v)
t5, _ = e3.featurize_volume(V, LV)
e3.close()
# round 2: PyRadiomics binning, basic morphology, the multi-device list call (two contexts on this device), a 3D cube beyond
# the shared-memory classes (global-memory path) with the 124-neighbour NGTDM
er = Engine(["*ALL_GLCM*", "*ALL_GLRLM*", "*ALL_GLSZM*", "*ALL_GLDM*", "*ALL_NGTDM*", "*ALL_INTENSITY*", "*BASIC_MORPHOLOGY*"])
er.set_param("coarse_gray_depth", -24)
t6, _ = er.featurize(I, L)
er.close()
from nyxus_b200.engine import featurize_list_multi
ea, eb = Engine(["*ALL_TEXTURE*", "*ALL_INTENSITY*"]), Engine(["*ALL_TEXTURE*", "*ALL_INTENSITY*"])
sl = [(np.ascontiguousarray(I), L), (np.ascontiguousarray(I[::-1]), np.ascontiguousarray(L[::-1])), (np.ascontiguousarray(I[:, ::-1]), np.ascontiguousarray(L[:, ::-1]))]
t7, _ = featurize_list_multi([ea, eb], [a.ctypes.data for a, _ in sl], True, [b.ctypes.data for _, b in sl], I.shape[0], I.shape[1])
ea.close(); eb.close()
zz, yy, xx = np.mgrid[0:40, 0:52, 0:60]
LB = ((((zz - 19) / 18.0) ** 2 + ((yy - 25) / 24.0) ** 2 + ((xx - 29) / 27.0) ** 2) < 1).astype(np.uint32) * 4     # cube ~ 37 x 49 x 55 > 40000 voxels
LB[0:5, 0:6, 0:7] = 9
VB = np.random.default_rng(8).integers(1, 2000, LB.shape).astype(np.uint16)
e4 = Engine(["*3D_ALL_INTENSITY*", "*3D_GLCM*", "*3D_GLSZM*", "*3D_NGTDM*"], dim=3)
for k, v in (("3glszm/greydepth", 32), ("3ngtdm/greydepth", 32), ("3ngtdm/radius", 2)):
    e4.set_param(k, v)
t8, _ = e4.featurize_volume(VB, LB)
e4.close()
# round 2, later: the 3D families that complete *3D_ALL_TEXTURE* (GLDM, GLRLM, GLDZM, NGLDM) on the shared-memory classes and on
# the large-cube path, the IBSI intensity-histogram family and the unweighted geometric moments
e5 = Engine(["*3D_ALL_TEXTURE*"], dim=3, coarse_gray_depth=24)
for k, v in (("3glszm/greydepth", 24), ("3ngtdm/greydepth", 24), ("3ngtdm/radius", 1), ("3gldm/greydepth", 24), ("3glrlm/greydepth", -20)):
    e5.set_param(k, v)
t9, _ = e5.featurize_volume(V, LV)
t10, _ = e5.featurize_volume(VB, LB)
e5.close()
e6 = Engine(["*ALL_IH*", "SPAT_MOMENT_11", "HU_M3", "IMOM_RM_12", "IMOM_HU5"], ibsi=True)
t11, _ = e6.featurize((I >> 8).astype(np.uint16), L)
e6.close()
print("round-2b rows", t9.shape, t10.shape, t11.shape, "finite", all(np.isfinite(t).all() for t in (t9, t10, t11)))
print("round-2 rows", t6.shape, t7.shape, t8.shape, "finite", all(np.isfinite(t).all() for t in (t6, t7, t8)))
print("rows", t1.shape, t2.shape, t3.shape, t4.shape, t5.shape, "finite", all(np.isfinite(t).all() for t in (t1, t2, t3, t4, t5)))
