#!/bin/bash
R=${1:-r01c}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
echo "bench rc=$?"; cat gpurun_out/bench_$R.json; tail -5 gpurun_out/bench_$R.err
