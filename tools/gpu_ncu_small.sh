#!/bin/bash
mkdir -p gpurun_out
for K in k_softmax_topk k_attend k_beam_step; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 30 -c 1 -f -o gpurun_out/src_$K python tools/perf_probe.py > gpurun_out/ncu_src_$K.log 2>&1
  echo "$K rc=$?"
done
