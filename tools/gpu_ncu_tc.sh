#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -c 5 -f -o gpurun_out/prof_tc_shapes vsr-guided-cic_b200/csrc/build/selftest_gemm prof > gpurun_out/ncu_tc_shapes.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_tc_shapes.log; ls -la gpurun_out/*.ncu-rep
