#!/bin/bash
# GPU-box visit: a pytest -k selection only.  usage: tools/gpu_k.sh TAG "k-expr"
TAG=${1:-k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "$2" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -30 gpurun_out/${TAG}_pytest.log
