#!/bin/bash
# 256 x 128 pair tiles for B / D + C at a few hundred rows, and the vocabulary GEMM's own pair threshold
mkdir -p gpurun_out
for cfg in "base:" "p2:VSRDEC_PAIR2_MIN_ROWS=257" "e:VSRDEC_PAIR_MIN_ROWS_E=257" "p2e:VSRDEC_PAIR2_MIN_ROWS=257 VSRDEC_PAIR_MIN_ROWS_E=257"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python tools/stack_probe.py 100,60,130 1 > gpurun_out/r02u_probe_$name.jsonl 2>&1
  echo "== $name"; cut -c1-600 gpurun_out/r02u_probe_$name.jsonl
done
VSRDEC_PAIR2_MIN_ROWS=257 timeout 900 python -m pytest tests -m gpu -q -x -k "config2 or config3 or properties or shapes or pipeline" 2>&1 | tail -4
