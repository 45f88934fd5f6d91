#!/bin/bash
# compute-sanitizer over the round-2 kernels: GEMM self-test (single-CTA + CTA-pair kernels, both operand modes) and the
# tiny-model tests (beam search on both vocabulary-head paths, edge cases, odd dimensions, sampling, teacher forcing)
mkdir -p gpurun_out
SEL="small_model or edge or odd_dim or sample_rl_logprobs or odd_seq_len or sinkhorn"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 vsr-guided-cic_b200/csrc/build/selftest_gemm quick > gpurun_out/san_${tool}_gemm.log 2>&1
  echo "$tool gemm rc=$?: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|SELFTEST' gpurun_out/san_${tool}_gemm.log | tr '\n' ' ')"
done
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "$SEL" > gpurun_out/san_${tool}_pytest.log 2>&1
  echo "$tool pytest rc=$?: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/san_${tool}_pytest.log | tr '\n' ' ')"
done
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "small_model_beam or edge" > gpurun_out/san_synccheck_pytest.log 2>&1
echo "synccheck pytest rc=$?: $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/san_synccheck_pytest.log | tr '\n' ' ')"
