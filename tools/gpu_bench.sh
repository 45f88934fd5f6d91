#!/bin/bash
# bench (both arms) + ncu launch list + ncu full captures of the top kernels; results -> gpurun_out/
R=${1:-r01}
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err
echo "ref rc=$?"; cat gpurun_out/bench_ref_$R.json
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
echo "bench rc=$?"; cat gpurun_out/bench_$R.json; tail -5 gpurun_out/bench_$R.err
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 900 --csv \
    --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$R.log 2>&1
echo "ncu launches rc=$?"
for K in k_gemm_tc k_attend k_softmax_topk; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 25 -c 3 -f \
      -o gpurun_out/prof_${K}_$R python tools/perf_probe.py > gpurun_out/ncu_${K}_$R.log 2>&1
  echo "ncu $K rc=$?"
done
ls -la gpurun_out
