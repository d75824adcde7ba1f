import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from nyxus_b200 import Engine, synth
inten, label = synth.config2(seed=2, size=10000)
d_int = torch.from_numpy(inten).cuda(); d_lab = torch.from_numpy(label.view(np.int32)).cuda()
eng = Engine(["MEAN"])
eng.set_param("concurrent_families", 0)
H, W = inten.shape
acc = []
for _ in range(8):
    try:
        eng.featurize_ptr(d_int.data_ptr(), True, d_lab.data_ptr(), 1, H, W, on_device=True, copy=False)
    except Exception as e:
        print("err", e); 
    acc.append(eng.timings()["scan"])
print("scan ms", [round(x, 4) for x in acc])
