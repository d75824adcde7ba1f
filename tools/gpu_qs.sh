#!/bin/bash
# GPU-box visit: a pytest -k selection + tools/quick_step.py.  usage: tools/gpu_qs.sh TAG "k-expr"
TAG=${1:-qs}
mkdir -p gpurun_out
if [ -n "$2" ]; then
timeout 900 python -m pytest tests -m gpu -x -q -k "$2" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
fi
timeout 300 python tools/quick_step.py > gpurun_out/${TAG}_qs.log 2>&1
cat gpurun_out/${TAG}_qs.log
