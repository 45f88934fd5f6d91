#!/bin/bash
# last pass of the round: memcheck over the eval pre-step tests (bounded), full parity log, both bench arms
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "sinkhorn or role_orderer or s_ssp" > gpurun_out/san3_memcheck.log 2>&1
echo "memcheck rc=$?: $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/san3_memcheck.log | tr '\n' ' ')"
bash tools/gpu_r02x.sh
