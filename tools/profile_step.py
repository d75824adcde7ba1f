"""One resident pass of the hot path over a configs[1]-style slide -- the command wrapped by ncu (see profiles/README.md)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nyxus_b200 import Engine, synth

size = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
feats = sys.argv[3].split(",") if len(sys.argv) > 3 else ["*ALL_TEXTURE*", "*ALL_INTENSITY*"]
inten, label = synth.config2(seed=2, size=size)
d_int = torch.from_numpy(inten).cuda()
d_lab = torch.from_numpy(label.view(np.int32)).cuda()
eng = Engine(feats)
for kv in sys.argv[4:]:            # library knobs, e.g. roi_batches=1 concurrent_families=0
    k, v = kv.split("=")
    eng.set_param(k, float(v))
H, W = inten.shape
for _ in range(steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tab, _ = eng.featurize_ptr(d_int.data_ptr(), True, d_lab.data_ptr(), 1, H, W, on_device=True, copy=False)
    wall = (time.perf_counter() - t0) * 1e3
    print(tab.shape, "wall_ms %.2f" % wall, {k: round(v, 3) for k, v in eng.timings().items()})
