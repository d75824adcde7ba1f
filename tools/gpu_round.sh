#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list and one full-set capture of a resident step.
# usage: tools/gpu_round.sh TAG   (outputs under gpurun_out/TAG_*; the .ncu-rep itself is dropped -- gpurun_out/ is capped at 64 MiB)
TAG=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py 10000 2 > gpurun_out/${TAG}_launches.log 2>&1
bash tools/gpu_full.sh ${TAG}
tail -3 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
