#!/bin/bash
# compute-sanitizer over the eval pre-step kernels added after tools/gpu_sanitize.sh ran: S-level sorter (csrc/sort.cu), R-level
# network (csrc/ssp.cu), RoleOrderer, and the small-magnitude-weights steps
mkdir -p gpurun_out
SEL="s_ssp or sinkhorn or role_orderer or small_magnitude"
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "$SEL" > gpurun_out/san2_${tool}_pytest.log 2>&1
  echo "$tool pytest rc=$?: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/san2_${tool}_pytest.log | tr '\n' ' ')"
done
