#!/bin/bash
# first GPU shake-out: memcheck on the small model, the gpu test-suite, a perf probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/memcheck.log
timeout 1500 python -m pytest tests -m gpu -q -s -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 600 python tools/perf_probe.py > gpurun_out/perf_probe.log 2>&1
echo "probe rc=$?" | tee -a gpurun_out/perf_probe.log
tail -5 gpurun_out/memcheck.log; tail -30 gpurun_out/pytest_gpu.log; cat gpurun_out/perf_probe.log
