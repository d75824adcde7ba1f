#!/bin/bash
# ncu --set full of the launches matching a kernel-name regex in one resident configs[1] step, exported to CSV on the box.
# usage: tools/gpu_ncu_kernel.sh TAG REGEX [count]
TAG=${1:-k}; KREG=${2:-k_scan_stats}; CNT=${3:-4}
mkdir -p gpurun_out
REP=/tmp/${TAG}_k
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name "regex:${KREG}" -c ${CNT} -o ${REP} python tools/profile_step.py 10000 1 "*ALL_TEXTURE*,*ALL_INTENSITY*" roi_batches=1 concurrent_families=0 > gpurun_out/${TAG}_k.log 2>&1
ncu -i ${REP}.ncu-rep --page raw --csv > gpurun_out/${TAG}_k_raw.csv 2>/dev/null
ncu -i ${REP}.ncu-rep --page source --csv --print-source=sass 2>/dev/null | gzip -9 > gpurun_out/${TAG}_k_sass.csv.gz
tail -3 gpurun_out/${TAG}_k.log
