#!/bin/bash
mkdir -p gpurun_out
for cfg in "base:" "apair_f8:VSRDEC_A_F8=1 VSRDEC_A_PAIR_MIN_ROWS=1" "apair_f16:VSRDEC_A_PAIR_MIN_ROWS=1" "apair129_f8:VSRDEC_A_F8=1 VSRDEC_A_PAIR_MIN_ROWS=129"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python tools/stack_probe.py 100,400,1000 1 > gpurun_out/r02p_probe_$name.jsonl 2>&1
  echo "== $name"; cut -c1-470 gpurun_out/r02p_probe_$name.jsonl
  env $envs timeout 300 python tools/fwd_probe.py 2>&1 | tail -3
done
