"""The BASELINE.json configs other than the bench line (configs[1]), timed on one B200 through the C ABI:
  C1  1024^2, 256 blob ROIs, GLCM + *ALL_INTENSITY*                      (host arrays, Engine.featurize)
  C3  16k^2 single ROI, 256 levels, GLCM + GLRLM + GLSZM                 (resident, giant path)
  C4  512^3 volume, 2197 ROIs, 3D GLCM / GLSZM / NGTDM + first-order     (resident)
  C5  a job of N configs[1] slides from pinned host memory, *ALL_TEXTURE* + *ALL_INTENSITY* (pipelined list call) --
      what ONE GPU of the 64-slide job does
usage: run_configs.py [c1,c2ibsi,c3,c4,c5] [n_slides_c5] [cpu]   -> prints one JSON line per config (profiles/rN_configs_*.jsonl)
`cpu`: also time the reference's own CPU code (oracle/_ref, all host cores) for C1, C3 and C4 -- BASELINE.md section 2: C1 and
C4 at full size, C3 GLCM + GLRLM on an 8192^2 crop (one ROI = one host thread; x4 pixels to the full slide, flagged) and C3
GLSZM at 2048^2 (the reference materialises w*h*Ng ints, glszm.cpp:211-217: 4.3 GB there, 275 GB at 16k^2 -- it cannot run)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nyxus_b200 import Engine, synth

which = (sys.argv[1] if len(sys.argv) > 1 else "c1,c2ibsi,c3,c4,c5").split(",")
n5 = int(sys.argv[2]) if len(sys.argv) > 2 else 8
do_cpu = len(sys.argv) > 3 and sys.argv[3] == "cpu"
NT = os.cpu_count() or 1


def cpu_line(name, fn, sample):
    from oracle import ref
    if not ref.available():
        return
    t0 = time.perf_counter()
    n = fn(ref)
    dt = time.perf_counter() - t0
    print(json.dumps({"config": name, "impl": "reference (oracle/_ref, -O2)", "cores": NT, "s": dt, "rois": int(n), "sample": sample}), flush=True)


def timed(fn, reps=3):
    best, out = 1e30, None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best, out


if "c1" in which:
    I, L = synth.config1()
    eng = Engine(["*ALL_GLCM*", "*ALL_INTENSITY*"])
    dt, (tab, _) = timed(lambda: eng.featurize(I, L), 5)
    print(json.dumps({"config": "C1 1024^2 uint16, 256 ROIs, GLCM + *ALL_INTENSITY*, host arrays in, table out", "ms": dt * 1e3,
                      "rois": int(tab.shape[0]), "rois_per_s": tab.shape[0] / dt, "mpix_per_s": I.size / dt / 1e6,
                      "stage_ms": {k: round(v, 3) for k, v in eng.timings().items() if v > 0.001}}), flush=True)
    eng.close()
    if do_cpu:
        cpu_line("C1 CPU", lambda ref: len(ref.featurize_2d(I.astype(np.uint32), L, greydepth=64, fam_mask=ref.FAM["INT"] | ref.FAM["GLCM"], nthreads=NT)[0]),
                 "the full C1 slide, GLCM + first-order")

if "c2ibsi" in which:
    # second sweep with ibsi=True on configs[1] after re-quantising the intensities to <= 256 levels
    I, L = synth.config2(seed=2)
    Iq = (I >> 8).astype(np.uint16)                       # 0..255 (zeros stay zeros)
    d_i = torch.from_numpy(Iq).cuda()
    d_l = torch.from_numpy(L.view(np.int32)).cuda()
    eng = Engine(["*ALL_TEXTURE*", "*ALL_INTENSITY*"], ibsi=True)
    H, W = I.shape
    dt, (tab, _) = timed(lambda: eng.featurize_ptr(d_i.data_ptr(), True, d_l.data_ptr(), 1, H, W, on_device=True, copy=False))
    print(json.dumps({"config": "C2 with ibsi=True, intensities re-quantised to 8 bits (>> 8), *ALL_TEXTURE*+*ALL_INTENSITY*, resident",
                      "ms": dt * 1e3, "rois": int(tab.shape[0]), "rois_per_s": tab.shape[0] / dt, "mpix_per_s": I.size / dt / 1e6,
                      "stage_ms": {k: round(v, 3) for k, v in eng.timings().items() if v > 0.001}}), flush=True)
    eng.close()
    del d_i, d_l

if "c3" in which:
    I, L = synth.config3()
    d_i = torch.from_numpy(I).cuda()
    d_l = torch.from_numpy(L.view(np.int32)).cuda()
    eng = Engine(["*ALL_GLCM*", "*ALL_GLRLM*", "*ALL_GLSZM*"], coarse_gray_depth=256)
    H, W = I.shape
    dt, (tab, _) = timed(lambda: eng.featurize_ptr(d_i.data_ptr(), True, d_l.data_ptr(), 1, H, W, on_device=True))
    print(json.dumps({"config": "C3 16384^2 uint16 single ROI, 256 levels, GLCM + GLRLM + GLSZM, resident", "ms": dt * 1e3,
                      "rois": int(tab.shape[0]), "mpix_per_s": I.size / dt / 1e6,
                      "stage_ms": {k: round(v, 3) for k, v in eng.timings().items() if v > 0.001}}), flush=True)
    eng.close()
    del d_i, d_l
    if do_cpu:
        Ic = np.ascontiguousarray(I[:8192, :8192]).astype(np.uint32)
        Lc = np.ascontiguousarray(L[:8192, :8192])
        cpu_line("C3 CPU GLCM+GLRLM", lambda ref: len(ref.featurize_2d(Ic, Lc, greydepth=256, fam_mask=ref.FAM["GLCM"] | ref.FAM["GLRLM"], nthreads=NT)[0]),
                 "8192^2 crop of the C3 slide = 1/4 of its pixels (a single ROI runs on ONE host thread); linear extrapolation to 16384^2: x4")
        Is = np.ascontiguousarray(I[:2048, :2048]).astype(np.uint32)
        Ls = np.ascontiguousarray(L[:2048, :2048])
        cpu_line("C3 CPU GLSZM", lambda ref: len(ref.featurize_2d(Is, Ls, greydepth=256, fam_mask=ref.FAM["GLSZM"], nthreads=NT)[0]),
                 "2048^2 crop (1/64 of the pixels); the reference's GLSZM allocates w*h*Ng ints: it cannot run at 16384^2")

if "c4" in which:
    I, L = synth.config4()
    d_i = torch.from_numpy(I).cuda()
    d_l = torch.from_numpy(L.view(np.int32)).cuda()
    eng = Engine(["*3D_ALL_INTENSITY*", "*3D_GLCM*", "*3D_GLSZM*", "*3D_NGTDM*"], dim=3)
    for k, v in (("3glszm/greydepth", 64), ("3ngtdm/greydepth", 64), ("3ngtdm/radius", 1), ("3glcm/greydepth", 64)):
        eng.set_param(k, v)
    D, H, W = I.shape
    ip = [d_i.data_ptr()]
    lp = [d_l.data_ptr()]
    dt, (tab, _) = timed(lambda: eng.featurize_list(ip, True, lp, H, W, on_device=True, D=D))
    print(json.dumps({"config": "C4 512^3 uint16 volume, 3D GLCM/GLSZM/NGTDM + first-order (grey depth 64, radius 1), resident",
                      "ms": dt * 1e3, "rois": int(tab.shape[0]), "rois_per_s": tab.shape[0] / dt, "mvox_per_s": I.size / dt / 1e6,
                      "stage_ms": {k: round(v, 3) for k, v in eng.timings().items() if v > 0.001}}), flush=True)
    eng.close()
    del d_i, d_l
    if do_cpu:
        cpu_line("C4 CPU", lambda ref: len(ref.featurize_3d(I.astype(np.uint32), L, greydepth=64, glszm_greydepth=64, ngtdm_greydepth=64, ngtdm_radius=1, nthreads=NT)[0]),
                 "the full 512^3 volume, 3D GLCM / GLSZM / NGTDM + first-order, ROIs spread over the host threads")

if "c5" in which:
    slides = [synth.config2(seed=100 + k) for k in range(2)]
    pins = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b.view(np.int32)).pin_memory()) for a, b in slides]
    eng = Engine(["*ALL_TEXTURE*", "*ALL_INTENSITY*"])
    H, W = slides[0][0].shape
    ips = [pins[k % 2][0].data_ptr() for k in range(n5)]
    lps = [pins[k % 2][1].data_ptr() for k in range(n5)]
    eng.featurize_list(ips, True, lps, H, W, copy=False)
    dt, (tab, _) = timed(lambda: eng.featurize_list(ips, True, lps, H, W, copy=False), 2)
    print(json.dumps({"config": "C5 share of one GPU: %d configs[1] slides (pinned host, 2 distinct alternating), *ALL_TEXTURE*+*ALL_INTENSITY*, "
                                "one pipelined list call" % n5, "ms": dt * 1e3, "ms_per_slide": dt * 1e3 / n5, "rois": int(tab.shape[0]),
                      "rois_per_s": tab.shape[0] / dt, "mpix_per_s": n5 * H * W / dt / 1e6,
                      "extrapolated_64_slides_s": dt / n5 * 64}), flush=True)
    eng.close()
