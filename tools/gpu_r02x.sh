#!/bin/bash
# evidence pass without the ncu leg (hot-path kernels unchanged since tools/gpu_r02l.sh ran): parity log, both bench arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -rA > gpurun_out/r02_parity_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_parity_pytest.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_r02x.json 2> gpurun_out/bench_ref_r02x.err
echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref_r02x.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02x.json 2> gpurun_out/bench_r02x.err
echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_r02x.json; tail -3 gpurun_out/bench_r02x.err
