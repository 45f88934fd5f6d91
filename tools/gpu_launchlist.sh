#!/bin/bash
# per-launch device times with WARM caches (no ncu cache flush): the in-situ kernel durations
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 300 -c 400 --csv --log-file gpurun_out/launches_warm.csv python tools/perf_probe.py > gpurun_out/ncu_probe.log 2>&1
echo "rc=$?"
