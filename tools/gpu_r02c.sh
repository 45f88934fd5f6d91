#!/bin/bash
# isolated kernel durations (warm-cache ncu launch lists) of the new vs the round-1 small kernels, and repeated,
# interleaved decode timings of the four combinations
mkdir -p gpurun_out
export VSRDEC_GRAPH=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 300 -c 300 --csv \
    --log-file gpurun_out/r02c_launches_new.csv python tools/ncu_probe.py 100 5 > gpurun_out/r02c_l1.log 2>&1
VSRDEC_FUSE_TAIL=0 VSRDEC_ATTEND=row timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 300 -c 300 --csv \
    --log-file gpurun_out/r02c_launches_old.csv python tools/ncu_probe.py 100 5 > gpurun_out/r02c_l2.log 2>&1
python tools/ncu_summary.py gpurun_out/r02c_launches_new.csv gpurun_out/r02c_launches_old.csv
unset VSRDEC_GRAPH
for rep in 1 2 3; do
  for cfg in "default:" "tail0:VSRDEC_FUSE_TAIL=0" "rowatt:VSRDEC_ATTEND=row" "both_old:VSRDEC_FUSE_TAIL=0 VSRDEC_ATTEND=row"; do
    name=${cfg%%:*}; envs=${cfg#*:}
    echo "== $name rep $rep: $(env $envs timeout 300 python tools/stack_probe.py 100,400 1 2>&1 | cut -c1-75 | tr '\n' ' ')"
  done
done
