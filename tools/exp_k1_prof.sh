#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for s in 8 6; do
  echo "== variant 2 slots $s"; SES_K1_VARIANT=2 SES_K1_SLOTS=$s python tools/k1_bench.py --reps 5
  SES_K1_VARIANT=2 SES_K1_SLOTS=$s python tools/k1_bench.py --reps 5 --pop 8192
done
SES_K1_VARIANT=2 SES_K1_SLOTS=8 ncu --set full --clock-control none --import-source on -k regex:k_rollout_slots -s 2 -c 1 -f -o gpurun_out/k1v2s8 python tools/k1_bench.py --regime converged --reps 1 | tail -3
SES_K1_VARIANT=2 SES_K1_SLOTS=6 ncu --set full --clock-control none --import-source on -k regex:k_rollout_slots -s 2 -c 1 -f -o gpurun_out/k1v2s6 python tools/k1_bench.py --regime converged --reps 1 | tail -3
python tools/variants_bench.py
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "spread" 2>&1 | tail -5
} > gpurun_out/exp_k1_prof.log 2>&1
tail -30 gpurun_out/exp_k1_prof.log
