#!/bin/bash
# Round-2 last visit (one B200): complete GPU test-suite, both bench arms, ncu launch list + full-set capture of one resident
# configs[1] step (sanitizer summaries come from tools/gpu_sanitize.sh on the same build).  usage: tools/gpu_round2d.sh TAG
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py 10000 2 > gpurun_out/${TAG}_launches.log 2>&1
bash tools/gpu_full.sh ${TAG}
tail -4 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_bench.json | cut -c1-400; tail -3 gpurun_out/${TAG}_bench.err
