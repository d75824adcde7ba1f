#!/bin/bash
# Round-2 closing visit (one B200): complete GPU test-suite, both bench arms, ncu launch list + full-set capture of one resident
# configs[1] step, compute-sanitizer (memcheck / synccheck / racecheck), the other BASELINE configs.  usage: tools/gpu_round2c.sh TAG
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py 10000 2 > gpurun_out/${TAG}_launches.log 2>&1
bash tools/gpu_full.sh ${TAG}
bash tools/gpu_sanitize.sh ${TAG}
timeout 900 python tools/run_configs.py c1,c2ibsi,c3,c4 8 > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err
tail -4 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_bench.json | cut -c1-1200; tail -3 gpurun_out/${TAG}_bench.err; cut -c1-300 gpurun_out/${TAG}_configs.jsonl
