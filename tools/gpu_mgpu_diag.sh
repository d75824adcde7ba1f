#!/bin/bash
# 8-GPU box visit: pure-copy ceiling at N=1,2,4,8 and the engine's per-rank stage times at N=8 (tools/mgpu_diag.py)
TAG=${1:-r2_mgpu}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
lscpu > gpurun_out/${TAG}_lscpu.txt 2>&1
numactl -H >> gpurun_out/${TAG}_lscpu.txt 2>&1
for N in 1 2 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_diag.py ${TAG} --copy-only 2>gpurun_out/${TAG}_n${N}.err | tee -a gpurun_out/${TAG}_summary.txt
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/mgpu_diag.py ${TAG} 2>gpurun_out/${TAG}_n8.err | tee -a gpurun_out/${TAG}_summary.txt
