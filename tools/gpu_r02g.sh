#!/bin/bash
mkdir -p gpurun_out
timeout 300 vsr-guided-cic_b200/csrc/build/selftest_gemm > gpurun_out/r02g_selftest.log 2>&1
echo "selftest rc=$?"; grep -E "MISMATCH|SELFTEST|CUDA|VSR" gpurun_out/r02g_selftest.log | head; grep "pair" gpurun_out/r02g_selftest.log | cut -c1-200
timeout 600 python -m pytest tests -m gpu -q -x -k "other_launch_shapes or decode_pipeline or properties_at_full or sample_rl" > gpurun_out/r02g_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02g_pytest.log
for cfg in "pair:" "nopair:VSRDEC_PAIR=0" "pair_f16x3:VSRDEC_GEMM=f16x3" "nopair_f16x3:VSRDEC_PAIR=0 VSRDEC_GEMM=f16x3"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python tools/stack_probe.py 300,400,800 1 > gpurun_out/r02g_probe_$name.jsonl 2>&1
  echo "== $name"; cut -c1-440 gpurun_out/r02g_probe_$name.jsonl
done
