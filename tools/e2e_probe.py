"""Where the end-to-end list call spends its time: pipelined jobs of N pinned host slides (device-generated, copied to pinned
host memory), ms per slide for several job lengths, next to the resident step.  usage: tools/e2e_probe.py [n1 n2 ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nyxus_b200 import Engine, synth

ns = [int(a) for a in sys.argv[1:]] or [1, 2, 5, 10, 20]
pins = []
for seed in (2, 3):
    di, dl = synth.config2_device(seed, "cuda:0")
    pins.append((di.cpu().pin_memory(), dl.cpu().pin_memory()))
    del di, dl
torch.cuda.synchronize()
eng = Engine(["*ALL_TEXTURE*", "*ALL_INTENSITY*"], coarse_gray_depth=64)
H, W = pins[0][0].shape
for n in ns:
    ips = [pins[k % 2][0].data_ptr() for k in range(n)]
    lps = [pins[k % 2][1].data_ptr() for k in range(n)]
    eng.featurize_list(ips, True, lps, H, W, on_device=False, copy=False)
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.featurize_list(ips, True, lps, H, W, on_device=False, copy=False)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    tm = eng.timings()
    print("job of %2d slides: %.2f ms per slide (%.1f ms job)   last slide stages: total %.2f h2d %.2f d2h %.2f" % (n, best * 1e3 / n, best * 1e3, tm["total"], tm["h2d"], tm["d2h"]), flush=True)
