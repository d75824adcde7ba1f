#!/usr/bin/env python
"""Multi-GPU diagnosis (run under torchrun, one rank per GPU): what bounds the 1 -> 8 GPU curve of bench.py?

  1. pure-copy ceiling: every rank moves one configs[1] slide's worth of bytes between pinned host memory and its GPU,
     H2D only (600 MB), D2H only (281 MB) and both at once on two streams -- no kernels at all.
  2. the resident step and the pipelined end-to-end job of bench.py with the per-stage device times of EVERY rank.

Each rank writes gpurun_out/<tag>_n<world>_rank<r>.json; rank 0 prints a summary line per experiment.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "mgpu"
    do_engine = "--copy-only" not in sys.argv
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = {"rank": rank, "world": world, "cpu_affinity": sorted(os.sched_getaffinity(0))[:4] + ["..."], "ncpu": os.cpu_count()}
    H2D_B, D2H_B = 600_000_000, 281_191_008
    h_in = torch.empty(H2D_B, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(D2H_B, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    d_in = torch.empty(H2D_B, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(D2H_B, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    K = 10

    def run_copy(up, down):
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            if up:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        mine = (time.perf_counter() - t0) / K * 1e3
        barrier()
        allv = (time.perf_counter() - t0) / K * 1e3
        return mine, allv

    for name, up, down in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
        run_copy(up, down)
        mine, allv = run_copy(up, down)
        nbytes = (H2D_B if up else 0) + (D2H_B if down else 0)
        out["copy_" + name] = {"ms_this_rank": mine, "ms_all_ranks": allv, "gbs_this_rank": nbytes / mine / 1e6,
                               "gbs_aggregate": nbytes * world / allv / 1e6}
        if rank == 0:
            print("copy %-5s N=%d: %.2f ms per slide-equivalent on rank 0, %.2f ms until all ranks done, %.1f GB/s aggregate"
                  % (name, world, mine, allv, nbytes * world / allv / 1e6), flush=True)
    del h_in, h_out, d_in, d_out

    if do_engine:
        from nyxus_b200 import Engine, synth
        inten, label = synth.config2(seed=2 + rank, size=10000)
        Hh, Ww = inten.shape
        h_int = torch.from_numpy(inten).pin_memory()
        h_lab = torch.from_numpy(label.view(np.int32)).pin_memory()
        d_int, d_lab = h_int.cuda(), h_lab.cuda()
        eng = Engine(["*ALL_TEXTURE*", "*ALL_INTENSITY*"], device=local, coarse_gray_depth=64)
        for _ in range(3):
            eng.featurize_ptr(d_int.data_ptr(), True, d_lab.data_ptr(), 1, Hh, Ww, on_device=True, copy=False)
        barrier()
        stages = {}
        t0 = time.perf_counter()
        for _ in range(K):
            eng.featurize_ptr(d_int.data_ptr(), True, d_lab.data_ptr(), 1, Hh, Ww, on_device=True, copy=False)
            for k, v in eng.timings().items():
                stages[k] = stages.get(k, 0.0) + v / K
        torch.cuda.synchronize()
        mine = (time.perf_counter() - t0) / K * 1e3
        barrier()
        out["resident"] = {"wall_ms_this_rank": mine, "stage_ms": stages}
        # resident slides through the LIST entry (one call, K slides: the rows of slide k travel while slide k+1 computes)
        ips, lps = [d_int.data_ptr()] * K, [d_lab.data_ptr()] * K
        eng.featurize_list(ips, True, lps, Hh, Ww, on_device=True, copy=False)
        barrier()
        t0 = time.perf_counter()
        eng.featurize_list(ips, True, lps, Hh, Ww, on_device=True, copy=False)
        mine = (time.perf_counter() - t0) / K * 1e3
        barrier()
        out["resident_list"] = {"wall_ms_this_rank": mine}
        ips, lps = [h_int.data_ptr()] * K, [h_lab.data_ptr()] * K
        eng.featurize_list(ips, True, lps, Hh, Ww, on_device=False, copy=False)
        barrier()
        t0 = time.perf_counter()
        eng.featurize_list(ips, True, lps, Hh, Ww, on_device=False, copy=False)
        mine = (time.perf_counter() - t0) / K * 1e3
        barrier()
        allv = (time.perf_counter() - t0) / K * 1e3
        out["e2e_list"] = {"wall_ms_this_rank": mine, "wall_ms_all_ranks": allv, "stage_ms_last_slide": eng.timings()}
        if rank == 0:
            print("engine N=%d: resident %.2f ms (d2h stage %.2f, families %.2f), resident list %.2f ms, e2e list %.2f ms per slide (rank 0)"
                  % (world, out["resident"]["wall_ms_this_rank"], stages.get("d2h", 0), stages.get("families", 0),
                     out["resident_list"]["wall_ms_this_rank"], mine), flush=True)
        eng.close()

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "%s_n%d_rank%d.json" % (tag, world, rank)), "w") as f:
        json.dump(out, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
