#!/bin/bash
# bounded racecheck + synccheck over a minimal eval pre-step workload (12 captions: every pre-step kernel, tcgen05 GEMMs included)
mkdir -p gpurun_out
for tool in racecheck synccheck; do
  timeout 200 compute-sanitizer --tool $tool --error-exitcode 9 python tools/racecheck_prestep.py > gpurun_out/san4_$tool.log 2>&1
  echo "$tool rc=$?: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|ordered' gpurun_out/san4_$tool.log | tr '\n' ' ')"
done
