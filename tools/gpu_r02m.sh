#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -rA > gpurun_out/r02_parity_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_parity_pytest.log; grep -E "^FAILED" gpurun_out/r02_parity_pytest.log | cut -c1-200
timeout 300 python tools/stack_probe.py 100,400,1000 1 > gpurun_out/r02m_probe.jsonl 2>&1; echo "== probe"; cut -c1-470 gpurun_out/r02m_probe.jsonl
VSRDEC_ATTEND_CAP_ROWS=100000000 timeout 300 python tools/stack_probe.py 1000 1 > gpurun_out/r02m_probe_rowatt.jsonl 2>&1; echo "== probe row-attention at b=1000"; cut -c1-470 gpurun_out/r02m_probe_rowatt.jsonl
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02m.json 2> gpurun_out/bench_r02m.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_r02m.err
python - gpurun_out/bench_r02m.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print({k:(round(d[k],1) if isinstance(d[k],float) else d[k]) for k in ['value','ms_per_step']}, 'e2e',round(d['e2e']['value']), 'e2e_idx', round(d['e2e_indexed']['value']), 'one', d['one_at_a_time'], 'parity', d['parity_check']['ok'], d['parity_check']['stacked_decode_equals_single_decode'], 'roofline', round(d['roofline']['frac'],3), d['roofline']['traffic'], 'fwd', round(d['forward_teacher']['ms_per_forward'],3), 'p50', d['p50_step_latency_ms'])
PY
