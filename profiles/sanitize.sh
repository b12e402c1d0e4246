#!/bin/bash
# compute-sanitizer runs of the known-answer test and the small parity cases. Run on the GPU box:
#   bash profiles/sanitize.sh            -> gpurun_out/sanitizer_{memcheck,initcheck,racecheck,synccheck}.log
# The selection covers every kernel family on small shapes: the reference KAT, the tiny goldens (geometry, prepare,
# all forward paths), the ragged feature map (partial blocks, pad bins, idle warps of the scatter forward and the joint /
# column / block backward kernels), the small wide / narrow tile cases of the column backward (16-column kernel: raw ranks aliasing
# the bulk-copy ring, ragged last tile, rolled camera, bf16), and the truncation / NaN / Inf edges.
set -u
mkdir -p gpurun_out
SEL='reference_kat or tiny_bev_z1 or tiny_occ_z16 or ragged or wide_and_narrow or errors_are_loud or grid_transpose or truncation'
for tool in memcheck initcheck racecheck synccheck; do
  # synccheck tracks mbarriers in a fixed table: the 16-column backward has 32 per CTA x 148 resident CTAs
  extra=""; [ $tool = synccheck ] && extra="--num-cuda-barriers 65536"
  timeout 1500 compute-sanitizer --tool $tool $extra --error-exitcode 9 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/sanitizer_$tool.log | tr '\n' ' ')"
done
# the column backward's two tile widths forced onto every small shape (the knob is read once per process)
for w in 8 16; do
  for tool in memcheck racecheck; do
    BEVPOOL_BWD_TILE_W=$w timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
        python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "wide_and_narrow" > gpurun_out/sanitizer_${tool}_tile$w.log 2>&1
    echo "tile_w=$w $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/sanitizer_${tool}_tile$w.log | tr '\n' ' ')"
  done
done
