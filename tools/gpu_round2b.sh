#!/bin/bash
# Round-2 visit B (one B200): compute-sanitizer (memcheck / synccheck / racecheck), the other BASELINE configs with the
# reference's CPU time beside them, and ncu full-set captures of the giant-ROI (C3) and 3D (C4) kernels.
# usage: tools/gpu_round2b.sh TAG
TAG=${1:-r2}
mkdir -p gpurun_out
bash tools/gpu_sanitize.sh ${TAG}
timeout 1500 python tools/run_configs.py c1,c2ibsi,c3,c4 8 cpu > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err
REP=/tmp/${TAG}_c3
timeout 900 ncu --set full --clock-control none --import-source on -c 80 -o ${REP} python tools/run_c3.py 4096 > gpurun_out/${TAG}_c3_full.log 2>&1
ncu -i ${REP}.ncu-rep --page raw --csv > /tmp/${TAG}_c3_raw.csv 2>/dev/null && python tools/ncu_summary.py /tmp/${TAG}_c3_raw.csv > gpurun_out/${TAG}_ncu_full_c3_summary.csv
REP=/tmp/${TAG}_c4
timeout 900 ncu --set full --clock-control none --import-source on -c 40 -o ${REP} python tools/run_configs.py c4 > gpurun_out/${TAG}_c4_full.log 2>&1
ncu -i ${REP}.ncu-rep --page raw --csv > /tmp/${TAG}_c4_raw.csv 2>/dev/null && python tools/ncu_summary.py /tmp/${TAG}_c4_raw.csv > gpurun_out/${TAG}_ncu_full_c4_summary.csv
cat gpurun_out/${TAG}_configs.jsonl | cut -c1-400; ls -la gpurun_out/${TAG}_*
