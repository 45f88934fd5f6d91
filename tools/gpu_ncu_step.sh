#!/bin/bash
# one decoder step's four GEMM launches (A, B, D+C, E) with the full metric set
mkdir -p gpurun_out
VSRDEC_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 250 -c 4 -f -o gpurun_out/prof_gemm_step python tools/perf_probe.py > gpurun_out/ncu_gemm_step.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_gemm_step.log
