"""BASELINE configs[2]: whole-image single ROI, 256 grey levels, GLCM + GLRLM + GLSZM through the giant path.
usage: run_c3.py SIZE [ref]   -- `ref` also runs the compiled reference (oracle/_ref) on the same slide and compares
(GLSZM only up to 1024: the reference materialises a (w*h) x Ng matrix)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nyxus_b200 import Engine, synth

size = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
do_ref = len(sys.argv) > 2 and sys.argv[2] == "ref"
feats = ["*ALL_GLCM*", "*ALL_GLRLM*", "*ALL_GLSZM*"]
if len(sys.argv) > 3:
    feats = sys.argv[3].split(",")
inten, label = synth.config3(size=size)
d_int = torch.from_numpy(inten).cuda()
d_lab = torch.from_numpy(label.view(np.int32)).cuda()
eng = Engine(feats, coarse_gray_depth=256)
H, W = inten.shape
res = {}
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tab, _ = eng.featurize_ptr(d_int.data_ptr(), True, d_lab.data_ptr(), 1, H, W, on_device=True, copy=True)
    wall = (time.perf_counter() - t0) * 1e3
    tm = {k: round(v, 3) for k, v in eng.timings().items() if v > 0.01}
    print("size", size, "wall_ms %.2f" % wall, tm, eng.counts(), flush=True)
res = {"size": size, "features": feats, "wall_ms": wall, "stage_ms": tm, "mpix_per_s": H * W / wall / 1e3}
cols = eng.columns
print({c: tab[0, 2 + i] for i, c in enumerate(cols) if c in ("GLCM_ASM_0", "GLCM_CONTRAST_0", "GLRLM_SRE_0", "GLRLM_RP_0", "GLSZM_SAE", "GLSZM_ZP", "GLSZM_ZE", "GLSZM_LAE")})
if do_ref:
    from oracle import ref
    fam = 0
    for f in feats:
        fam |= {"*ALL_GLCM*": ref.FAM["GLCM"], "*ALL_GLRLM*": ref.FAM["GLRLM"], "*ALL_GLSZM*": ref.FAM["GLSZM"], "*ALL_GLDZM*": ref.FAM["GLDZM"],
                "*ALL_GLDM*": ref.FAM["GLDM"], "*ALL_NGLDM*": ref.FAM["NGLDM"], "*ALL_NGTDM*": ref.FAM["NGTDM"], "*ALL_INTENSITY*": ref.FAM["INT"]}[f]
    if size > 1024:
        fam &= ~ref.FAM["GLSZM"]
    t0 = time.perf_counter()
    rl, rt, _ = ref.featurize_2d(inten.astype(np.uint32), label, greydepth=256, fam_mask=fam, nthreads=os.cpu_count())
    dt = time.perf_counter() - t0
    rcols = ref.columns()
    rfam = ref.column_families()
    nbad, worst = 0, 0.0
    for i, c in enumerate(cols):
        j = rcols.index(c)
        if not (rfam[j] & fam):
            continue
        a, b = rt[0, j], tab[0, 2 + i]
        d = abs(a - b)
        rel = d / max(abs(a), abs(b), 1e-300)
        if d > 1e-12:
            worst = max(worst, rel)
        if rel > 1e-9 and d > 1e-12:
            nbad += 1
            print("  MISMATCH", c, a, b)
    print("reference %.2f s (%d threads), families mask %d: %d mismatches, worst rel %.3g" % (dt, os.cpu_count(), fam, nbad, worst))
    res["reference_s"] = dt
    res["mismatches"] = nbad
    res["worst_rel"] = worst
print("RESULT", json.dumps(res))
