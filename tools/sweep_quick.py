"""Times resident configs[1] steps (device-generated slide, see quick_step.py) for several settings of the library's knobs.
usage: sweep_quick.py "k1=v1,k2=v2" "k1=v3" ...   (each argument is one setting; knobs not named keep their value)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nyxus_b200 import Engine, synth

d_int, d_lab = synth.config2_device(2, "cuda:0")
torch.cuda.synchronize()
eng = Engine(["*ALL_TEXTURE*", "*ALL_INTENSITY*"], coarse_gray_depth=64)
H, W = d_int.shape
for setting in sys.argv[1:]:
    for kv in setting.split(","):
        k, v = kv.split("=")
        eng.set_param(k, float(v))
    for _ in range(3):
        eng.featurize_ptr(d_int.data_ptr(), True, d_lab.data_ptr(), 1, H, W, on_device=True, copy=False)
    tot = fam = 0.0
    for _ in range(10):
        eng.featurize_ptr(d_int.data_ptr(), True, d_lab.data_ptr(), 1, H, W, on_device=True, copy=False)
        t = eng.timings()
        tot += t["total"] / 10
        fam += t["families"] / 10
    print(setting, "total %.3f ms families %.3f ms" % (tot, fam), flush=True)
