#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "k1_variants" 2>&1 | tail -8
for m in 0 2 3 5; do
  echo "== multi $m"; SES_K1_MULTI=$m python tools/k1_bench.py --reps 5
  SES_K1_MULTI=$m python tools/k1_bench.py --reps 3 --pop 8192
done
for m in 2 5; do
SES_K1_MULTI=$m ncu --set full --clock-control none --import-source on -k regex:k_rollout_slots -s 2 -c 1 -f -o gpurun_out/k1m$m python tools/k1_bench.py --regime converged --reps 1 | tail -1
done
} > gpurun_out/exp_k1_multi.log 2>&1
tail -30 gpurun_out/exp_k1_multi.log
