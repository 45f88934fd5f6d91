#!/bin/bash
# warm-cache ncu launch list of the eval pre-step (one RoleOrderer.order call on 100 captions ~ 250 launches)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 1200 -c 600 --csv \
    --log-file gpurun_out/r02_launches_prestep.csv python tools/prestep_probe.py 100 > gpurun_out/r02_launches_prestep.log 2>&1
echo "rc=$?"
python tools/ncu_summary.py gpurun_out/r02_launches_prestep.csv > gpurun_out/r02_prestep_ncu_summary.md 2>&1
head -30 gpurun_out/r02_prestep_ncu_summary.md
