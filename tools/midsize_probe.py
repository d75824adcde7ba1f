"""Times a slide of mid-size ROIs (1 600 ... 39 800 bounding-box cells, the CTA-per-ROI forms of every family) with the
CTA-per-ROI kernels at 128 threads and at 256 threads where their shared-memory footprint allows at most three CTAs per SM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nyxus_b200 import Engine

rng = np.random.default_rng(3)
H = W = 4096
yy, xx = np.mgrid[0:H, 0:W]
I = (3000 + 2000 * np.sin(xx / 23.0) * np.cos(yy / 31.0) + rng.normal(0, 500, (H, W))).clip(1, 60000).astype(np.uint16)
L = np.zeros((H, W), np.uint32)
lab = 1
for by in range(0, H - 200, 205):
    for bx in range(0, W - 200, 205):
        s = int(rng.choice([40, 64, 100, 150, 199]))
        cy, cx = by + 100, bx + 100
        L[((yy - cy) / (s / 2.0)) ** 2 + ((xx - cx) / (s / 2.0)) ** 2 <= 1.0] = lab
        lab += 1
e = Engine(["*ALL_TEXTURE*", "*ALL_INTENSITY*"])
for big in (0, 1, 0, 1):
    e.set_param("big_cta_threads", big)
    for _ in range(2):
        tab, _ = e.featurize(I, L)
    ts = []
    for _ in range(4):
        tab, _ = e.featurize(I, L)
        ts.append(e.timings())
    tm = {k: round(min(t[k] for t in ts), 3) for k in ts[0]}
    print("big_cta_threads", big, "rois", tab.shape[0], {k: v for k, v in tm.items() if v > 0.05}, flush=True)
