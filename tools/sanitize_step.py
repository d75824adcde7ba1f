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
print("rows", t1.shape, t2.shape, t3.shape, t4.shape, t5.shape, "finite", all(np.isfinite(t).all() for t in (t1, t2, t3, t4, t5)))
