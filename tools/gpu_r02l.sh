#!/bin/bash
# evidence pass for profiles/: parity log, both bench arms with the driver's flags, ncu launch lists + full captures, SASS counts
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -rA > gpurun_out/r02_parity_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_parity_pytest.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_r02l.json 2> gpurun_out/bench_ref_r02l.err
echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref_r02l.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02l.json 2> gpurun_out/bench_r02l.err
echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_r02l.json; tail -3 gpurun_out/bench_r02l.err
bash tools/gpu_ncu_r02.sh r02l > gpurun_out/r02l_ncu.log 2>&1
tail -30 gpurun_out/r02l_ncu.log | cut -c1-200
