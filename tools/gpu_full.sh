#!/bin/bash
# ncu --set full capture of ONE resident configs[1] step (one ROI batch, families serialised: one launch per kernel and
# size class), exported to CSV on the box: raw page for every launch + SASS page of the heaviest kernels.
# usage: tools/gpu_full.sh TAG [kernel-regex for the SASS pages]
TAG=${1:-run}
KREG=${2:-"k_roi_family|k_scan_stats|k_pack_tiles|k_roi_intensity"}
mkdir -p gpurun_out
REP=/tmp/${TAG}_full
timeout 1200 ncu --set full --clock-control none --import-source on -c 40 -o ${REP} python tools/profile_step.py 10000 1 "*ALL_TEXTURE*,*ALL_INTENSITY*" roi_batches=1 concurrent_families=0 > gpurun_out/${TAG}_full.log 2>&1
ncu -i ${REP}.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ncu -i ${REP}.ncu-rep --page source --csv --print-source=sass --kernel-name "regex:${KREG}" 2>/dev/null | gzip -9 > gpurun_out/${TAG}_full_sass.csv.gz
ls -la ${REP}.ncu-rep gpurun_out/${TAG}_full* >> gpurun_out/${TAG}_full.log
SZ=$(stat -c %s ${REP}.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -lt 25000000 ] && [ "$SZ" -gt 0 ]; then cp ${REP}.ncu-rep gpurun_out/${TAG}_full.ncu-rep; fi
du -sh gpurun_out | tail -1
