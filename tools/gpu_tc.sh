#!/bin/bash
mkdir -p gpurun_out
timeout 300 vsr-guided-cic_b200/csrc/build/selftest_gemm > gpurun_out/selftest_gemm.log 2>&1; echo "selftest rc=$?"
cat gpurun_out/selftest_gemm.log
timeout 900 python -m pytest tests -m gpu -q -s -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python tools/perf_probe.py > gpurun_out/perf_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/perf_probe.log
