#!/bin/bash
# ramp-up of DecodePipeline: pipeline tests, then the bench line (e2e_indexed is the number that moves)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "pipeline" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02r.json 2> gpurun_out/bench_r02r.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02r.json'))
print('value',round(d['value']),'one',round(d['one_at_a_time']['value']),'e2e',round(d['e2e']['value']),'idx',round(d['e2e_indexed']['value']),d['parity_check']['ok'])
PY
