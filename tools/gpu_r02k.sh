#!/bin/bash
mkdir -p gpurun_out
timeout 300 vsr-guided-cic_b200/csrc/build/selftest_gemm > gpurun_out/r02k_selftest.log 2>&1; echo "selftest rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02k_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02k_pytest.log; grep -E "^FAILED" gpurun_out/r02k_pytest.log | cut -c1-200
timeout 300 python tools/stack_probe.py 100,400,800,1000 1 > gpurun_out/r02k_probe.jsonl 2>&1; echo "== probe"; cut -c1-470 gpurun_out/r02k_probe.jsonl
for sl in "5 2" "10 1" "10 2" "20 1"; do
  set -- $sl
  timeout 600 python bench.py --steps 20 --warmup 5 --stack $1 --lanes $2 --no-cpu-baseline > gpurun_out/bench_r02k_s$1_l$2.json 2> gpurun_out/bench_r02k_s$1_l$2.err
  echo "== stack $1 lanes $2 rc=$?"; python - gpurun_out/bench_r02k_s$1_l$2.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print({k:(round(d[k],1) if isinstance(d[k],float) else d[k]) for k in ['value','ms_per_step']}, 'e2e',round(d['e2e']['value']), 'e2e_idx', round(d['e2e_indexed']['value']), 'one', round(d['one_at_a_time']['value']), 'parity', d['parity_check']['ok'], d['parity_check']['stacked_decode_equals_single_decode'], 'roofline', round(d['roofline']['frac'],3), 'fwd', round(d['forward_teacher']['ms_per_forward'],3))
PY
done
