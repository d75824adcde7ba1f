"""The drop-in call on pageable numpy arrays (Nyxus.featurize, DataFrame included) for several numbers of staging threads.
usage: tools/api_probe.py [t1 t2 ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nyxus_b200 import Nyxus, synth

ts = [int(a) for a in sys.argv[1:]] or [4, 8]
di, dl = synth.config2_device(2, "cuda:0")
inten = di.cpu().numpy().copy()
label = dl.cpu().numpy().view(np.uint32).copy()
del di, dl
nyx = Nyxus(["*ALL_TEXTURE*", "*ALL_INTENSITY*"])
for t in ts:
    nyx._engine.set_param("staging_threads", t) if hasattr(nyx, "_engine") else None
    for _ in range(2):
        nyx.featurize(inten, label)
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter()
        df = nyx.featurize(inten, label)
        best = min(best, time.perf_counter() - t0)
    print("staging_threads", t, "featurize %.1f ms" % (best * 1e3), df.shape, flush=True)
