#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_pair -s 5 -c 1 -f -o gpurun_out/r02h_pair_f16x3 vsr-guided-cic_b200/csrc/build/selftest_gemm > gpurun_out/r02h_a.log 2>&1
echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_pair -s 110 -c 1 -f -o gpurun_out/r02h_pair_f8 vsr-guided-cic_b200/csrc/build/selftest_gemm > gpurun_out/r02h_b.log 2>&1
echo rc=$?
ls -la gpurun_out/r02h*
