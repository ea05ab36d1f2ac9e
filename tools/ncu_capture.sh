#!/bin/bash
# ncu evidence for the headline kernels at config 3 on one GPU (run under gpurun; reports land in gpurun_out/):
#   tools/ncu_capture.sh <tag>      e.g. r2b
# 1. launch list of bench.py (gpu__time_duration per kernel; cold-cache, serialised: compare SHARES)
# 2. one --set full capture per kernel (third launch of tools/run_c3.py), source-level counters included
tag=${1:-rX}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-extras --no-cpu-baseline --no-second-mode > gpurun_out/${tag}_bench_under_ncu.log 2>&1
for k in k_pair_fwd2 k_pair_bwd2; do
  ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 2 --launch-count 1 -f \
      -o gpurun_out/${tag}_c3_$k python tools/run_c3.py > gpurun_out/ncu_${tag}_$k.log 2>&1
done
# both gradient-GEMM launches of one backward
ncu --set full --clock-control none --import-source on -k regex:k_grad_gemm2 --launch-skip 4 --launch-count 2 -f \
    -o gpurun_out/${tag}_c3_k_grad_gemm2 python tools/run_c3.py > gpurun_out/ncu_${tag}_k_grad_gemm2.log 2>&1
ls -la gpurun_out/${tag}_*
