#!/bin/bash
# full GPU regression: self-tests, gpu test-suite on both GEMM paths, smoke, bench (both arms)
R=${1:-r01f}
mkdir -p gpurun_out
timeout 120 vsr-guided-cic_b200/csrc/build/selftest_gemm > gpurun_out/selftest_gemm.log 2>&1; echo "selftest rc=$?"
timeout 120 vsr-guided-cic_b200/csrc/build/selftest_gemm pair > gpurun_out/selftest_pair.log 2>&1; echo "selftest pair rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest(tc) rc=$?"; tail -2 gpurun_out/pytest_gpu.log
VSRDEC_GEMM=simt timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_simt.log 2>&1; echo "pytest(simt twin) rc=$?"; tail -2 gpurun_out/pytest_gpu_simt.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$R.json")); r=json.load(open("gpurun_out/bench_ref_$R.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
print("roofline", d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["issued_frac"], "attend", d["roofline_attend"]["achieved"], d["roofline_attend"]["frac"])
print("one_at_a_time", d["one_at_a_time"]["value"], d["one_at_a_time"]["ms_per_step"]); print("e2e_indexed", d["e2e_indexed"]["value"], d["e2e_indexed"]["device_resident_value"], d["e2e_indexed"]["h2d_bytes_per_step"])
print("cpu", d.get("cpu_baseline",{}).get("value"), "ref arm", r["value"], r["cpu_baseline"]["cores"], "clocks", d["clocks"])
PY
