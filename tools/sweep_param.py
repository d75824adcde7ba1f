"""Times resident configs[1] steps for several values of one engine parameter.  usage: sweep_param.py KEY v1 v2 ..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nyxus_b200 import Engine, synth

key, vals = sys.argv[1], [float(v) for v in sys.argv[2:]]
inten, label = synth.config2(seed=2, size=10000)
d_int = torch.from_numpy(inten).cuda()
d_lab = torch.from_numpy(label.view(np.int32)).cuda()
eng = Engine(["*ALL_TEXTURE*", "*ALL_INTENSITY*"])
H, W = inten.shape
for v in vals:
    eng.set_param(key, v)
    for _ in range(3):
        eng.featurize_ptr(d_int.data_ptr(), True, d_lab.data_ptr(), 1, H, W, on_device=True, copy=False)
    tot = 0.0
    fam = 0.0
    for _ in range(8):
        eng.featurize_ptr(d_int.data_ptr(), True, d_lab.data_ptr(), 1, H, W, on_device=True, copy=False)
        t = eng.timings()
        tot += t["total"]; fam += t["families"]
    print(key, v, "total %.3f ms families %.3f ms" % (tot / 8, fam / 8), flush=True)
