#!/bin/bash
# GPU-box visit: a pytest -k selection + one bench line.  usage: tools/gpu_bench_k.sh TAG "k-expr"
TAG=${1:-bk}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "$2" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -6 gpurun_out/${TAG}_pytest.log; python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
print("serial", d["stage_ms_serialised"])
PY
tail -3 gpurun_out/${TAG}_bench.err
