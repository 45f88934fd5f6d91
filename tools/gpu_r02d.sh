#!/bin/bash
# fp8-residual GEMM mode: self-test of the kernel in both modes, parity suite in the default (f16+f8x2) mode,
# decode timings of both modes, bench
mkdir -p gpurun_out
timeout 300 vsr-guided-cic_b200/csrc/build/selftest_gemm > gpurun_out/r02d_selftest.log 2>&1
echo "selftest rc=$?"; grep -c OK gpurun_out/r02d_selftest.log; grep -E "MISMATCH|SELFTEST|CUDA|VSR" gpurun_out/r02d_selftest.log | head; grep "f8x2" gpurun_out/r02d_selftest.log | head -12
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02d_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02d_pytest.log; grep -E "^FAILED|PARITY" gpurun_out/r02d_pytest.log | cut -c1-330
for rep in 1 2; do
  for cfg in "f8:" "f16x3:VSRDEC_GEMM=f16x3"; do
    name=${cfg%%:*}; envs=${cfg#*:}
    env $envs timeout 300 python tools/stack_probe.py 100,400 1 > gpurun_out/r02d_probe_${name}_$rep.jsonl 2>&1
    echo "== $name rep $rep"; cut -c1-430 gpurun_out/r02d_probe_${name}_$rep.jsonl
  done
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02d.json 2> gpurun_out/bench_r02d.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_r02d.json; tail -5 gpurun_out/bench_r02d.err
