#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02j_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02j_pytest.log; grep -E "^FAILED" gpurun_out/r02j_pytest.log | cut -c1-200
for cfg in "pair_f8:" "nopair_f8:VSRDEC_PAIR=0" "pair_f16x3:VSRDEC_GEMM=f16x3" "nopair_f16x3:VSRDEC_PAIR=0 VSRDEC_GEMM=f16x3"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python tools/stack_probe.py 100,300,400,800 1 > gpurun_out/r02j_probe_$name.jsonl 2>&1
  echo "== $name"; cut -c1-470 gpurun_out/r02j_probe_$name.jsonl
done
timeout 300 python tools/stack_probe.py 400,600 2 > gpurun_out/r02j_probe_l2.jsonl 2>&1; echo "== lanes 2"; cat gpurun_out/r02j_probe_l2.jsonl
