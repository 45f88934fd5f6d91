#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -rA > gpurun_out/r02_parity_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_parity_pytest.log; grep -E "^FAILED" gpurun_out/r02_parity_pytest.log | cut -c1-200
for cfg in "base:" "min256:VSRDEC_PAIR_MIN_ROWS=256" "min640:VSRDEC_PAIR_MIN_ROWS=640"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python tools/stack_probe.py 100,200,300 1 > gpurun_out/r02q_probe_$name.jsonl 2>&1
  echo "== $name"; cut -c1-470 gpurun_out/r02q_probe_$name.jsonl
done
