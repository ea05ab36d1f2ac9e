#!/bin/bash
# one gpurun call: the GPU test-suite, then the default bench line (N = 1); outputs under gpurun_out/
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2b_pytest_gpu.log; tail -3 gpurun_out/r2b_pytest_gpu.log
python bench.py > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2b_bench_n1.json")); r=d["roofline"]
    print(d["value"], d["ms_per_step"], r["kernels"], r["step_algorithmic_tflops"], d["parity_check"]["ok"], d["weak_scaling"]["ms_per_step"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2b_bench_n1.err").read()[-1500:])
PY
