#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none -k regex:k_gemm_tc2 -s 84 -c 2 -f -o gpurun_out/prof_tc2_D vsr-guided-cic_b200/csrc/build/selftest_gemm pair > gpurun_out/ncu_tc2.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_tc2.log
