#!/bin/bash
# compute-sanitizer memcheck / racecheck / synccheck over tools/sanitize_step.py (2D, slide lists, giant ROI, PyRadiomics
# binning, morphology, multi-device call, 3D incl. the large-cube path); summaries land in gpurun_out/TAG_san_*.log.
# usage: tools/gpu_sanitize.sh TAG
TAG=${1:-san}
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 40 python tools/sanitize_step.py > gpurun_out/${TAG}_san_${tool}.full 2>&1
  echo "exit=$?" >> gpurun_out/${TAG}_san_${tool}.full
  # keep the report lines (not the per-thread detail) and the tool's own summary
  grep -E "^=========|^rows|^round-2|^exit=|Error|ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/${TAG}_san_${tool}.full | grep -vE "^=========     (at|by|in|Host Frame|Device Frame|Saved host)" | head -120 > gpurun_out/${TAG}_san_${tool}.log
  rm -f gpurun_out/${TAG}_san_${tool}.full
  tail -4 gpurun_out/${TAG}_san_${tool}.log
done
