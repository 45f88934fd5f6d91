#!/bin/bash
# round 2, second GPU pass: parity suite on the per-caption attention + fused tail build, A/B of the new kernels,
# the reworked bench (both arms)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02b_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02b_pytest.log
for cfg in "default:" "tail0:VSRDEC_FUSE_TAIL=0" "rowatt:VSRDEC_ATTEND=row" "both_old:VSRDEC_FUSE_TAIL=0 VSRDEC_ATTEND=row"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python tools/stack_probe.py 100,400 1 > gpurun_out/r02b_probe_$name.jsonl 2>&1
  echo "== $name"; cut -c1-420 gpurun_out/r02b_probe_$name.jsonl
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err
echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_r02b.json; tail -5 gpurun_out/bench_r02b.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02b.json 2> gpurun_out/bench_ref_r02b.err
echo "ref rc=$?"; cat gpurun_out/bench_ref_r02b.json
