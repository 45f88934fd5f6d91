#!/bin/bash
# end-of-round 8-GPU check (no H2D probes: profiles/r02_h2d_n*.json): the 2-GPU sharded-vs-single test and the bench line at N
N=${1:-8}; R=${2:-r02y}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "two_gpu" > gpurun_out/pytest_2gpu_${R}.log 2>&1; echo "2gpu test rc=$?"; tail -2 gpurun_out/pytest_2gpu_${R}.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_${R}_n$N.json 2> gpurun_out/scale_${R}_n$N.err
echo "bench rc=$?"; tail -3 gpurun_out/scale_${R}_n$N.err
python - gpurun_out/scale_${R}_n$N.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print({k:(round(d[k],1) if isinstance(d[k],float) else d[k]) for k in ['value','ms_per_step','n_gpus']}, 'e2e',round(d['e2e']['value']), 'h2d_gbs', round(d['e2e']['h2d_gbs'],1), 'e2e_idx', round(d['e2e_indexed']['value']), 'one', round(d['one_at_a_time']['value']), 'parity', d['parity_check'])
PY
