#!/bin/bash
# headline sweep: batches per call x calls in flight (K = 20 batches)
mkdir -p gpurun_out
for cfg in "10 2" "7 3" "5 4" "20 1"; do
  set -- $cfg
  timeout 600 python bench.py --steps 20 --warmup 5 --stack $1 --lanes $2 > gpurun_out/bench_r02v_s$1_l$2.json 2> gpurun_out/bench_r02v_s$1_l$2.err
  python - gpurun_out/bench_r02v_s$1_l$2.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], 'value',round(d['value']),'one',round(d['one_at_a_time']['value']),'e2e',round(d['e2e']['value']),'idx',round(d['e2e_indexed']['value']),'pre',d.get('eval_prestep',{}).get('ms_per_100_captions'), d['parity_check']['ok'])
PY
done
