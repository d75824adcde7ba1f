#!/bin/bash
# BASELINE configs[4]: the 64-slide job through nyxb200_featurize_2d_list_multi on N devices of the box (one process, one
# host thread per device), for every N given; "t8" = the torchrun form the driver uses, 8 ranks.
# usage (gpurun --gpus 8|2): tools/gpu_c5.sh TAG "8 4 t8"  |  tools/gpu_c5.sh TAG "2 1"
TAG=${1:-r2}
LIST=${2:-"8 4"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/${TAG}_c5_gpus_${LIST// /_}.txt 2>&1
free -g >> gpurun_out/${TAG}_c5_gpus_${LIST// /_}.txt
for N in $LIST; do
  if [ "$N" = "t8" ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --workload c5 --gpus 8 --steps 2 --warmup 1 > gpurun_out/${TAG}_c5_torchrun_n8.json 2> gpurun_out/${TAG}_c5_torchrun_n8.err
    echo "torchrun rc=$?"; cut -c1-330 gpurun_out/${TAG}_c5_torchrun_n8.json; tail -2 gpurun_out/${TAG}_c5_torchrun_n8.err
  else
    timeout 900 python bench.py --workload c5 --gpus $N --steps 2 --warmup 1 > gpurun_out/${TAG}_c5_n${N}.json 2> gpurun_out/${TAG}_c5_n${N}.err
    echo "N=$N rc=$?"; cut -c1-330 gpurun_out/${TAG}_c5_n${N}.json; tail -2 gpurun_out/${TAG}_c5_n${N}.err
  fi
done
