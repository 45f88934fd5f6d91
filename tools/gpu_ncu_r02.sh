#!/bin/bash
# round 2 ncu evidence: warm-cache launch lists (b=100 and a stacked b=1000 call) and `--set full` captures of the
# steady-state step kernels (GEMM A, B, D+C, E of decoder steps >= 1, attention, vocabulary merge, beam step); summaries -> gpurun_out/
R=${1:-r02}
mkdir -p gpurun_out
export VSRDEC_GRAPH=0
for B in 100 1000; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 300 -c 300 --csv \
      --log-file gpurun_out/${R}_launches_b$B.csv python tools/ncu_probe.py $B 5 > gpurun_out/${R}_launches_b$B.log 2>&1
  echo "launch list b=$B rc=$?"
  # per decode: 2 prologue GEMMs + 20 x (A, B, D+C, E); decode 3 starts at GEMM launch 246: its step 1 = 252..255, step 2 = 256..259
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm -s 252 -c 8 -f \
      -o gpurun_out/${R}_gemm_b$B python tools/ncu_probe.py $B 5 > gpurun_out/${R}_ncu_gemm_b$B.log 2>&1
  echo "gemm full b=$B rc=$?"
  for K in k_attend k_vocab_merge k_beam_step; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 70 -c 2 -f \
        -o gpurun_out/${R}_${K}_b$B python tools/ncu_probe.py $B 5 > gpurun_out/${R}_ncu_${K}_b$B.log 2>&1
    echo "$K full b=$B rc=$?"
  done
done
python tools/ncu_summary.py gpurun_out/${R}_launches_b100.csv gpurun_out/${R}_launches_b1000.csv gpurun_out/${R}_*.ncu-rep > gpurun_out/${R}_ncu_summary.md 2>&1
python tools/ncu_summary.py --traffic-json gpurun_out/${R}_ncu_traffic.json gpurun_out/${R}_gemm_b100.ncu-rep gpurun_out/${R}_gemm_b1000.ncu-rep gpurun_out/${R}_k_attend_b100.ncu-rep gpurun_out/${R}_k_attend_b1000.ncu-rep
rm -f gpurun_out/${R}_*.ncu-rep      # (64 MiB cap on what comes back; the summaries carry the numbers)
head -40 gpurun_out/${R}_ncu_summary.md; tail -8 gpurun_out/${R}_ncu_traffic.json
