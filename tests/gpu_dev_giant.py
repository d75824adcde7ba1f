"""Dev script (run under gpurun): the large-ROI (giant) path against oracle/_ref -- full table + matrices."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_dev_check import compare
from nyx_testutil import giant_cases

if __name__ == "__main__":
    tot = 0
    only = sys.argv[1:]
    for name, I, L, kw in giant_cases():
        if only and name not in only:
            continue
        t0 = time.time()
        tot += compare(I, L, name=name, **kw)
        print("   (%.1fs)" % (time.time() - t0), flush=True)
    print("TOTAL BAD", tot)
