#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02f_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02f_pytest.log; grep -E "^FAILED|PARITY" gpurun_out/r02f_pytest.log | cut -c1-330
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err
echo "bench rc=$?"; tail -5 gpurun_out/bench_r02f.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02f.json'))
for k in ['value','ms_per_step','one_at_a_time','p50_step_latency_ms','e2e','e2e_indexed','parity_check','forward_teacher','cpu_baseline','clocks','gpu_launches']:
    print(k, json.dumps(d.get(k))[:400])
for k in ['roofline','roofline_b100','roofline_attend']:
    r=d[k]; print(k, {x:r[x] for x in ['achieved','peak','frac','traffic','ms_per_decode','share_of_step'] if x in r})
PY
