#!/bin/bash
# Short GPU-box visit: parity tests + one bench line.  usage: tools/gpu_quick.sh TAG [pytest -k expr]
TAG=${1:-q}
mkdir -p gpurun_out
if [ -n "$2" ]; then
timeout 900 python -m pytest tests -m gpu -x -q -k "$2" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
else
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -15 gpurun_out/${TAG}_pytest.log; python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
print("conc", d["stage_ms"]); print("serial", d["stage_ms_serialised"])
PY
tail -3 gpurun_out/${TAG}_bench.err
