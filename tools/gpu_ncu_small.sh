#!/bin/bash
# source-level `--set full` captures of the small per-step kernels (one launch each, from a mid-decode step)
mkdir -p gpurun_out
for K in ${@:-k_vocab_merge k_attend k_beam_step}; do
  VSRDEC_GRAPH=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 30 -c 1 -f -o gpurun_out/src_$K python tools/perf_probe.py > gpurun_out/ncu_src_$K.log 2>&1
  echo "$K rc=$?"
done
