python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r3_pytest_a.log; tail -3 gpurun_out/r3_pytest_a.log
python bench.py > gpurun_out/r3_bench_a.json 2> gpurun_out/r3_bench_a.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r3_bench_a.json")); r=d["roofline"]
    print(d["value"], d["ms_per_step"], r["kernels"], r["step_algorithmic_tflops"], d["parity_check"]["ok"], d["weak_scaling"]["ms_per_step"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r3_bench_a.err").read()[-1500:])
PY
