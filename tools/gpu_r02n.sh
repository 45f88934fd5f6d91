#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s -x -k "sinkhorn" > gpurun_out/r02n_ssp.log 2>&1; echo "ssp test rc=$?"; grep -E "PARITY|passed|failed|Error|error" gpurun_out/r02n_ssp.log | head -10
python - <<'PY'
import sys, os, time
sys.path.insert(0, 'vsr-guided-cic_b200'); sys.path.insert(0, '.')
import torch
from models import SinkhornNet
net = SinkhornNet(10, 20, 0.1).cuda().eval()
for B in (1, 16, 500, 5000):
    x = torch.relu(torch.randn(B, 10, 2352, device='cuda'))
    for _ in range(3): net.assign(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): net.assign(x)
    e1.record(); torch.cuda.synchronize()
    print(f"sinkhorn+assign B={B}: {e0.elapsed_time(e1)/10*1e3:.1f} us per call, {e0.elapsed_time(e1)/10*1e3/B:.2f} us per problem")
PY
bash tools/gpu_sanitize.sh 2>&1 | tail -8
