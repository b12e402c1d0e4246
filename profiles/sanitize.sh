#!/bin/bash
# compute-sanitizer runs of the known-answer test and the small parity cases. Run on the GPU box:
#   bash profiles/sanitize.sh            -> gpurun_out/sanitizer_{memcheck,initcheck,racecheck,synccheck}.log
# The selection covers every kernel family on small shapes: the reference KAT, the tiny goldens (geometry, prepare,
# all forward paths), the ragged feature map (partial blocks, pad bins, idle warps of the scatter forward and the joint /
# column / block backward kernels), bf16 on the fused path, and the rolled-camera case (mixed bins of the column kernel).
set -u
mkdir -p gpurun_out
SEL='reference_kat or tiny_bev_z1 or tiny_occ_z16 or ragged or errors_are_loud or grid_transpose or truncation'
for tool in memcheck initcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/sanitizer_$tool.log | tr '\n' ' ')"
done
