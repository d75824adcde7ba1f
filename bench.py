#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native texture + intensity feature path.

A "step" = one pass of the hot path (phase-1 scan -> CSR pack -> 7 texture families + first-order -> table) over one
synthetic whole-slide tile: BASELINE.json configs[1], 10k x 10k uint16 intensity + uint32 label mask with 103,684
nucleus-sized ROIs, features *ALL_TEXTURE* + *ALL_INTENSITY* (337 columns), default settings (64 grey levels, MATLAB
binning, GLCM offset 1).  One process per GPU; at N > 1 every rank processes its OWN slide (ROIs/slides are independent:
weak scaling, no data-path collective) and the time is the max over ranks.

  value : ROIs/s with the slide already resident in HBM (C ABI nyxb200_featurize_2d_device), device time from CUDA events
          recorded on the library's launching stream, cross-checked against the host clock around the K steps.
  e2e   : same metric through the host-buffer C ABI call with PINNED host inputs: nyxb200_featurize_2d_list calls over
          K slides in jobs of at most 10 (two distinct slides alternating), every slide's H2D copy and every table's D2H
          copy inside the timed region; the library uploads slide k+1 while slide k computes.  e2e_single = K separate one-slide calls
          (nyxb200_featurize_2d_u16), nothing overlapped across calls.
  --impl reference : the reference's own CPU implementation (oracle/_ref = the unmodified reference sources compiled at
          -O2 + our replay driver) on a bounded crop of the same slide with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FEATURES = ["*ALL_TEXTURE*", "*ALL_INTENSITY*"]
METRIC = "ROIs/sec for *ALL_TEXTURE*+*ALL_INTENSITY* (Mpix/sec in config)"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample_side(ref, synth, cores, target_s):
    """side of the configs[1]-style crop the CPU reference needs about target_s seconds for (probe: a 1550^2 crop)"""
    probe = 1550
    pi, pl = synth.config2(seed=2, size=probe)
    t0 = time.perf_counter()
    ref.featurize_2d(pi.astype(np.uint32), pl, greydepth=64, ibsi=False, nthreads=cores)
    dt = max(time.perf_counter() - t0, 1e-3)
    side = int(probe * (target_s / dt) ** 0.5)
    return max(31 * 20, min(10000, side // 31 * 31))


def reference_arm(args, rank, world):
    """The reference's own CPU implementation on the host cores (rank 0 only)."""
    if rank != 0:
        return
    from oracle import ref
    from nyxus_b200 import synth
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libnyxus_ref.so not built (run oracle/Makefile.ref where /root/reference exists)"}))
        return
    cores = os.cpu_count() or 1
    crop = args.ref_crop if args.ref_crop > 0 else cpu_sample_side(ref, synth, cores, target_s=4.0)
    inten, label = synth.config2(seed=2, size=crop)
    inten32 = inten.astype(np.uint32)
    times = []
    nroi = 0
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        labels, tab, tph = ref.featurize_2d(inten32, label, greydepth=64, ibsi=False, nthreads=cores)
        dt = time.perf_counter() - t0
        nroi = len(labels)
        if it >= args.warmup:
            times.append(dt)
    T = float(np.mean(times))
    val = nroi / T
    sample = "%dx%d crop of the configs[1] slide generator (%d ROIs), all families, one pass per step" % (crop, crop, nroi)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "ROIs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": T * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1] bounded sample: " + sample, "features": FEATURES, "coarse_gray_depth": 64,
                       "mpix_per_s": crop * crop / T / 1e6},
            "cpu_baseline": {"value": val, "unit": "ROIs/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": val, "unit": "ROIs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def c5_job(args, rank, local_rank, world):
    """BASELINE configs[4]: ONE job of 64 distinct configs[1]-shaped slides (seeds 100..163), *ALL_TEXTURE* + *ALL_INTENSITY*,
    sharded by pixel count over N GPUs of one box through nyxb200_featurize_2d_list_multi: one process, one host thread
    and one context per device, no inter-GPU exchange, every row copied from its device straight into ONE pinned host
    table.  Strong scaling: the job is the same at every N.  Under torchrun the ranks only provide the barrier (NCCL):
    rank 0 drives all N devices, which is how a caller of the library would use a multi-GPU box."""
    import torch
    import torch.distributed as dist
    from nyxus_b200 import Engine, synth
    from nyxus_b200.engine import featurize_list_multi, shard_slides

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    N = args.gpus
    if torch.cuda.device_count() < N:
        raise SystemExit("bench.py: --gpus %d but only %d devices are visible" % (N, torch.cuda.device_count()))
    if world > 1:
        torch.cuda.set_device(local_rank)
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
        store = dist.distributed_c10d._get_default_store()
        if rank != 0:
            # rank 0 works; the others wait for it on the rendezvous store (host side: an NCCL barrier would keep a
            # spinning kernel on every GPU and a spinning thread on every rank's core while rank 0 uses them)
            from datetime import timedelta
            store.wait(["nyx_c5_done"], timedelta(minutes=30))
            dist.barrier()
            dist.destroy_process_group()
            return
    S = args.slides
    size = args.size
    owner = shard_slides(N, S)          # equal pixel counts: slide k -> device k % N
    # ---- the slides: generated on the device that will process them (resident leg) + pinned host copies (e2e leg)
    d_int, d_lab = [], []
    for k in range(S):
        i, l = synth.config2_device(100 + k, "cuda:%d" % owner[k], size=size)
        d_int.append(i)
        d_lab.append(l)
    for d in range(N):
        torch.cuda.synchronize(d)
    import psutil
    per_slide = size * size * 6
    nh = args.host_slides if args.host_slides > 0 else (S if psutil.virtual_memory().available > 3 * S * per_slide + (64 << 30) else min(S, 16))
    h_int = [torch.empty((size, size), dtype=torch.uint16).pin_memory() for _ in range(nh)]
    h_lab = [torch.empty((size, size), dtype=torch.int32).pin_memory() for _ in range(nh)]
    for k in range(nh):
        h_int[k].copy_(d_int[k])
        h_lab[k].copy_(d_lab[k])
    engines = [Engine(FEATURES, device=d, coarse_gray_depth=64) for d in range(N)]
    ncols = len(engines[0].columns)

    def sync_all():
        for d in range(N):
            torch.cuda.synchronize(d)

    def job(resident):
        if resident:
            ips, lps = [t.data_ptr() for t in d_int], [t.data_ptr() for t in d_lab]
        else:
            ips, lps = [h_int[k % nh].data_ptr() for k in range(S)], [h_lab[k % nh].data_ptr() for k in range(S)]
        tab, rimg = featurize_list_multi(engines, ips, True, lps, size, size, on_device=resident, copy=False)
        dev_ms = max(e.timings()["total"] for e in engines)
        launches = sum(e.counts()["launches"] for e in engines)
        return tab.shape[0], dev_ms, launches, tab, rimg

    res = {}
    samplers = [ClockSampler(d) for d in range(N)]
    for resident in (True, False):
        for _ in range(max(1, min(args.warmup, 2))):      # the first job also pins the 18 GB result table
            job(resident)
        sync_all()
        if resident:
            for sm in samplers:
                sm.start()
        t0 = time.perf_counter()
        dev_ms = 0.0
        rois = 0
        launches = 0
        for _ in range(args.steps):
            n, ms, ln, tab, rimg = job(resident)
            rois += n
            dev_ms += ms
            launches += ln
        sync_all()
        wall = time.perf_counter() - t0
        if resident:
            clocks = [sm.stop() for sm in samplers]
        # the one table: rows of slide k are contiguous, slides in order, labels ascending inside a slide
        assert tab.shape[1] == ncols + 2 and np.all(np.diff(rimg) >= 0) and int(rimg[-1]) == S - 1
        res[resident] = dict(rois=rois, dev_ms=dev_ms, wall=wall, launches=launches, rows=tab.shape[0])
    r, e = res[True], res[False]
    nroi = r["rows"]
    npix = size * size * S
    clk = {"sm_mhz": min(c["sm_mhz"] or 0 for c in clocks), "sm_max_mhz": max(c["sm_max_mhz"] or 0 for c in clocks),
           "reasons": sorted(set(x for c in clocks for x in c["reasons"])), "per_gpu_sm_mhz": [c["sm_mhz"] for c in clocks]}
    line = {
        "metric": METRIC, "value": r["rois"] / r["wall"], "unit": "ROIs/s", "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["wall"] / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[4]: ONE job of %d distinct %dx%d uint16 slides + uint32 masks (%d ROIs, seeds 100..%d, generated on the device: synth.config2_device), *ALL_TEXTURE*+*ALL_INTENSITY* (%d columns), sharded by pixel count over %d GPU(s) by nyxb200_featurize_2d_list_multi, rows gathered in ONE pinned host table inside the timed region" % (S, size, size, nroi, 100 + S - 1, ncols, N),
                   "features": FEATURES, "coarse_gray_depth": 64, "rois_per_job": nroi, "slides": S,
                   "mpix_per_s": npix * args.steps / r["wall"] / 1e6,
                   "l2": "every slide (600 MB) exceeds the 126 MB L2; no explicit flush",
                   "timing": "host clock around the K jobs, all devices synchronised on both sides (one call spans N devices, so there is no single stream to put events on); device_ms_per_step = max over devices of the summed per-slide CUDA-event times",
                   "device_ms_per_step": r["dev_ms"] / args.steps,
                   "process_model": "one process, one host thread + one context per GPU (torchrun ranks > 0 only join the barrier)"},
        "e2e": {"value": e["rois"] / e["wall"], "unit": "ROIs/s", "h2d_bytes_per_step": npix * 6, "d2h_bytes_per_step": nroi * (ncols + 2) * 8,
                "ms_per_step": e["wall"] / args.steps * 1e3, "mpix_per_s": npix * args.steps / e["wall"] / 1e6,
                "how": "the same call with %d distinct PINNED host slides%s; every slide's H2D and every row's D2H inside the timed region" % (nh, "" if nh == S else " cycled over the %d slides of the job" % S)},
        "gpu_launches": r["launches"],
        "clocks": clk,
    }
    print(json.dumps(line))
    for en in engines:
        en.close()
    if world > 1:
        store.set("nyx_c5_done", "1")
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=10000, help="slide side (configs[1] = 10000)")
    ap.add_argument("--ref-crop", type=int, default=0, help="side of the crop the CPU reference is timed on (0 = sized from a probe: "
                    "~4 s per step for --impl reference, ~15 s for the cpu_baseline of the default arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"], help="c2 = configs[1], one slide per GPU (default, weak scaling); "
                    "c5 = configs[4], ONE job of --slides distinct slides sharded by pixel count over the --gpus devices (strong scaling)")
    ap.add_argument("--slides", type=int, default=64, help="slides of the c5 job (configs[4] = 64)")
    ap.add_argument("--host-slides", type=int, default=0, help="c5: distinct slides held in pinned host memory for the e2e leg "
                    "(0 = all of them when the host has the memory, else 16 cycled)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    if args.workload == "c5":
        c5_job(args, rank, local_rank, world)
        return

    import torch
    import torch.distributed as dist
    from nyxus_b200 import Engine, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries exactly one JSON line (rank 0): NCCL's own banner ("NCCL version ...", printed by the C library on
        # the first communicator) is sent to stderr by pointing fd 1 at fd 2 while the communicator is created
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic slide of this rank (weak scaling: one configs[1] slide per GPU)
    inten, label = synth.config2(seed=2 + rank, size=args.size)
    H, W = inten.shape
    npix = H * W
    h_int = torch.from_numpy(inten).pin_memory()
    h_lab = torch.from_numpy(label.view(np.int32)).pin_memory()
    d_int = h_int.cuda()
    d_lab = h_lab.cuda()
    torch.cuda.synchronize()

    eng = Engine(FEATURES, device=local_rank, coarse_gray_depth=64)
    ncols = len(eng.columns)

    def step_resident():
        tab, _ = eng.featurize_ptr(d_int.data_ptr(), True, d_lab.data_ptr(), 1, H, W, on_device=True, copy=False)
        return tab

    def step_e2e():
        tab, _ = eng.featurize_ptr(h_int.data_ptr(), True, h_lab.data_ptr(), 1, H, W, on_device=False, copy=False)
        return tab

    # second host slide for the pipelined e2e job (the slides of a job differ; the device never sees the same buffer twice in a row)
    inten2, label2 = synth.config2(seed=1002 + rank, size=args.size)
    h_int2 = torch.from_numpy(inten2).pin_memory()
    h_lab2 = torch.from_numpy(label2.view(np.int32)).pin_memory()

    def job_e2e(k):
        ips = [(h_int if i % 2 == 0 else h_int2).data_ptr() for i in range(k)]
        lps = [(h_lab if i % 2 == 0 else h_lab2).data_ptr() for i in range(k)]
        tab, _ = eng.featurize_list(ips, True, lps, H, W, on_device=False, copy=False)
        return tab.shape[0]

    # ---- resident (value)
    for _ in range(args.warmup):
        tab = step_resident()
    nroi = tab.shape[0]
    counts = eng.counts()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    stage_sum = {}
    ev_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_resident()
        tm = eng.timings()
        ev_ms += tm["total"]
        for k, v in tm.items():
            stage_sum[k] = stage_sum.get(k, 0.0) + v
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = eng.counts()["launches"] * args.steps

    # ---- per-stage times for the roofline lines: the same step with the families serialised on one stream (in the timed
    # region above they overlap on eight streams, so their own begin/end events include time shared with other kernels)
    eng.set_param("concurrent_families", 0)
    eng.set_param("roi_batches", 1)      # one launch per kernel and size class: stage time == launch time
    serial_sum, n_serial = {}, 3
    step_resident()
    for _ in range(n_serial):
        step_resident()
        for k, v in eng.timings().items():
            serial_sum[k] = serial_sum.get(k, 0.0) + v
    eng.set_param("concurrent_families", 1)
    eng.set_param("roi_batches", 0)
    barrier()

    # ---- e2e (host pinned buffers through the C ABI): one pipelined K-slide job, and K separate calls
    # jobs of at most 10 slides (the library keeps the rows of a whole job in pinned host memory: 10 x 281 MB per rank),
    # back to back until exactly --steps slides went through
    JOB = 10
    jobs = [min(JOB, args.steps - i) for i in range(0, args.steps, JOB)]
    job_e2e(jobs[0])           # warm-up job of the same length (the pinned result buffer grows to its final size here)
    barrier()
    t1 = time.perf_counter()
    rois_e2e = sum(job_e2e(k) for k in jobs)
    barrier()
    wall_e2e = time.perf_counter() - t1
    for _ in range(2):
        step_e2e()
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    wall_e2e_single = time.perf_counter() - t1

    # ---- the drop-in call itself: Nyxus.featurize(numpy uint16, numpy uint32) on PAGEABLE arrays, DataFrame included
    from nyxus_b200 import Nyxus
    nyx = Nyxus(FEATURES, use_gpu_device=local_rank)
    api_rows = 0
    for _ in range(2):
        nyx.featurize(inten, label)
    barrier()
    n_api = min(args.steps, 5)
    t1 = time.perf_counter()
    for i in range(n_api):
        df = nyx.featurize(inten if i % 2 == 0 else inten2, label if i % 2 == 0 else label2)
        api_rows += df.shape[0]
        del df
    barrier()
    wall_api = time.perf_counter() - t1
    del nyx

    # ---- max over ranks
    ev_s = ev_ms / 1e3
    tt = torch.tensor([ev_s, wall, wall_e2e, wall_e2e_single, wall_api], dtype=torch.float64, device="cuda")
    nn = torch.tensor([float(nroi), float(rois_e2e), float(api_rows)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(nn, op=dist.ReduceOp.SUM)
    ev_s, wall, wall_e2e, wall_e2e_single, wall_api = [float(x) for x in tt.tolist()]
    total_rois, total_rois_e2e, total_rois_api = [float(x) for x in nn.tolist()]

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        per_conc = {k: v / args.steps for k, v in stage_sum.items()}
        per = {k: v / n_serial for k, v in serial_sum.items()}
        fam_stages = ["intensity", "glcm", "glrlm", "gldzm", "glszm", "gldm", "ngldm", "ngtdm", "scan", "pack", "compact", "select"]
        dom = max(fam_stages, key=lambda k: per.get(k, 0.0))
        nfg = counts["fg_pixels"]
        # algorithmic bytes per launch (DESIGN.md "Algorithmic bytes"): each element once at the width that crosses the API
        ncol_fam = {"intensity": 36, "glcm": 149, "glrlm": 80, "gldzm": 18, "glszm": 16, "gldm": 14, "ngldm": 19, "ngtdm": 5}
        alg = {"scan": npix * (2 + 4) + nroi * 32,
               "pack": nfg * (2 + 4) + nfg * 4,
               "intensity": nfg * 4 + nroi * 36 * 8}
        for f, nc in ncol_fam.items():
            if f != "intensity":
                alg[f] = nfg * 2 + nroi * nc * 8
        # DRAM bytes per step of every stage from the committed `ncu --set full` capture (dram__bytes_read.sum +
        # dram__bytes_write.sum, summed over the stage's launches of one step; tools/ncu_traffic.py)
        traffic = {}
        tps = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_traffic.json"))
        tp = os.path.join(ROOT, "profiles", tps[-1]) if tps else ""
        if tp and args.size == 10000:
            try:
                traffic = json.load(open(tp))["bytes_per_step"]
            except Exception:
                traffic = {}

        def roof(stage):
            ms = per.get(stage, 0.0)
            a = alg.get(stage, 0) / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            return {"kernel": stage, "bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak,
                    "traffic": traffic.get(stage), "ms_per_launch": ms, "algorithmic_bytes": alg.get(stage, 0), "peak_source": peak_src}
        line = {
            "metric": METRIC, "value": total_rois * args.steps / ev_s, "unit": "ROIs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ev_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: %dx%d uint16 slide + uint32 mask, %d ROIs per GPU, *ALL_TEXTURE*+*ALL_INTENSITY* (%d columns)" % (H, W, nroi, ncols),
                       "features": FEATURES, "coarse_gray_depth": 64, "rois_per_gpu": nroi, "fg_pixels_per_gpu": nfg,
                       "mpix_per_s": npix * world * args.steps / ev_s / 1e6,
                       "l2": "inputs (%.0f MB per step) exceed the 126 MB L2; no explicit flush" % (npix * 6 / 1e6),
                       "timing": "CUDA events on the library's launching stream (first to last event of every call), summed over the K steps, max over ranks; the families overlap on 8 streams and the ROI batches follow each other in the timed region; the roofline lines use stage times of 3 extra steps with the families serialised and one ROI batch (one launch per kernel and size class)",
                       "wall_ms_per_step": wall / args.steps * 1e3},
            "e2e": {"value": total_rois_e2e / wall_e2e, "unit": "ROIs/s", "h2d_bytes_per_step": npix * 6,
                    "d2h_bytes_per_step": nroi * (ncols + 2) * 8, "ms_per_step": wall_e2e / args.steps * 1e3,
                    "mpix_per_s": npix * world * args.steps / wall_e2e / 1e6,
                    "how": "nyxb200_featurize_2d_list calls over %s pinned host slides (2 distinct, alternating), back to back; host clock around them" % "+".join(str(k) for k in jobs)},
            "e2e_single": {"value": total_rois * args.steps / wall_e2e_single, "unit": "ROIs/s",
                           "ms_per_step": wall_e2e_single / args.steps * 1e3,
                           "how": "%d separate nyxb200_featurize_2d_u16 calls, one slide each (no overlap across calls)" % args.steps},
            "e2e_api": {"value": total_rois_api / wall_api, "unit": "ROIs/s", "ms_per_step": wall_api / n_api * 1e3,
                        "how": "%d calls of nyxus_b200.Nyxus.featurize(numpy uint16 [H,W], numpy uint32 [H,W]) on PAGEABLE arrays (two slides alternating), pandas DataFrame included; host clock" % n_api},
            "gpu_launches": launches,
            "clocks": clocks,
            "stage_ms": {k: round(v, 4) for k, v in per_conc.items()},
            "stage_ms_serialised": {k: round(v, 4) for k, v in per.items()},
            "roofline": roof(dom),
            "roofline_passes": [roof(s) for s in ("scan", "pack", "intensity")],
        }
        if not args.no_cpu_baseline:
            from oracle import ref
            if ref.available():
                cores = os.cpu_count() or 1
                crop = args.ref_crop if args.ref_crop > 0 else cpu_sample_side(ref, synth, cores, target_s=15.0)
                ci, cl = synth.config2(seed=2, size=crop)
                t2 = time.perf_counter()
                rl, _, tph = ref.featurize_2d(ci.astype(np.uint32), cl, greydepth=64, nthreads=cores)
                dt = time.perf_counter() - t2
                line["cpu_baseline"] = {"value": len(rl) / dt, "unit": "ROIs/s", "cores": cores, "kind": "reference",
                                        "sample": "%dx%d crop of the same slide generator (%d ROIs), one pass, %.1f s" % (crop, crop, len(rl), dt),
                                        "phase_s": [round(float(x), 4) for x in tph[:11]]}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "ROIs/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
