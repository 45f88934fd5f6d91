#!/bin/bash
mkdir -p gpurun_out
timeout 300 vsr-guided-cic_b200/csrc/build/selftest_gemm > gpurun_out/r02o_selftest.log 2>&1
echo "selftest rc=$?"; grep -E "MISMATCH|SELFTEST|CUDA|VSR" gpurun_out/r02o_selftest.log | head; grep "pair" gpurun_out/r02o_selftest.log | cut -c1-200
timeout 600 python -m pytest tests -m gpu -q -x -k "other_launch_shapes or properties_at_full or config2" > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02o_pytest.log
timeout 300 python tools/stack_probe.py 100,400,1000 1 > gpurun_out/r02o_probe.jsonl 2>&1; echo "== probe"; cut -c1-470 gpurun_out/r02o_probe.jsonl
VSRDEC_PAIR_KB32=0 timeout 300 python tools/stack_probe.py 1000 1 > gpurun_out/r02o_probe_kb64.jsonl 2>&1; echo "== probe A with 64-element k-blocks"; cut -c1-470 gpurun_out/r02o_probe_kb64.jsonl
