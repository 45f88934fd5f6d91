#!/bin/bash
# plain epilogue through TMA tensor stores vs LSU stores, decode level
mkdir -p gpurun_out
for cfg in "tma:VSRDEC_TMA_STORE=1" "lsu:VSRDEC_TMA_STORE=0" "tma2:VSRDEC_TMA_STORE=1" "lsu2:VSRDEC_TMA_STORE=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python tools/stack_probe.py 100,1000 1 > gpurun_out/r02w_probe_$name.jsonl 2>&1
  echo "== $name"; cut -c1-560 gpurun_out/r02w_probe_$name.jsonl
done
