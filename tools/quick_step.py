"""Short timing of the resident configs[1] step on a device-generated slide (synth.config2_device: same geometry / intensity
model as bench.py's host-generated slide, seconds instead of a minute to set up): the step with the families on their
streams, then the stage times with the families serialised and one ROI batch.  Development aid -- bench.py is the record.
usage: tools/quick_step.py [steps] [features] [knob=value ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nyxus_b200 import Engine, synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
feats = sys.argv[2].split(",") if len(sys.argv) > 2 else ["*ALL_TEXTURE*", "*ALL_INTENSITY*"]
d_int, d_lab = synth.config2_device(2, "cuda:0")
torch.cuda.synchronize()
eng = Engine(feats, coarse_gray_depth=64)
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    eng.set_param(k, float(v))
H, W = d_int.shape


def run(n):
    acc = {}
    for _ in range(n):
        tab, _ = eng.featurize_ptr(d_int.data_ptr(), True, d_lab.data_ptr(), 1, H, W, on_device=True, copy=False)
        for k, v in eng.timings().items():
            acc[k] = acc.get(k, 0.0) + v / n
    return tab, acc


run(3)
tab, conc = run(steps)
print("rois", tab.shape, "step_ms %.3f" % conc["total"])
eng.set_param("concurrent_families", 0)
eng.set_param("roi_batches", 1)
run(1)
_, ser = run(3)
print("serial", {k: round(v, 3) for k, v in ser.items()})
