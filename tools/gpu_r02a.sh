#!/bin/bash
# round 2, first GPU pass: parity suite (with -s: the PARITY lines go to profiles/r02_parity.md), GEMM self-test,
# throughput of a single call vs the number of stacked captions
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r02a_pytest.log; grep -c PARITY gpurun_out/r02a_pytest.log
timeout 300 vsr-guided-cic_b200/csrc/build/selftest_gemm > gpurun_out/r02a_selftest.log 2>&1
echo "selftest rc=$?"; tail -3 gpurun_out/r02a_selftest.log
timeout 600 python tools/stack_probe.py 100,200,300,400,600,800 1 > gpurun_out/r02a_stack_l1.jsonl 2> gpurun_out/r02a_stack_l1.err
echo "stack l1 rc=$?"; cut -c1-110 gpurun_out/r02a_stack_l1.jsonl
timeout 600 python tools/stack_probe.py 100,200,300,400 2 > gpurun_out/r02a_stack_l2.jsonl 2> gpurun_out/r02a_stack_l2.err
echo "stack l2 rc=$?"; cat gpurun_out/r02a_stack_l2.jsonl
VSRDEC_ATTEND=row timeout 300 python tools/stack_probe.py 100,300 1 > gpurun_out/r02a_stack_rowatt.jsonl 2>&1
echo "row-attend A/B:"; cut -c1-400 gpurun_out/r02a_stack_rowatt.jsonl
