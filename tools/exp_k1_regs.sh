#!/bin/bash
# GPU experiment: K1 variants (0 scalar FFMA, 1 / 2 packed FFMA2 with / without the Newton step, 3 / 4 / 5 part of the slot's
# weights in registers; 4 is the default) -- parity, throughput in both regimes at P = 65536 and P = 8192, operand-form
# micro-benchmark.  Results of round 1: profiles/r01_k1_experiments.md.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o gpurun_out/ffma2_forms tools/micro/ffma2_forms.cu
{
./gpurun_out/ffma2_forms
for v in 0 1 2 3 4 5; do
  echo "== variant $v"; SES_K1_VARIANT=$v python tools/k1_bench.py --reps 5
  SES_K1_VARIANT=$v python tools/k1_bench.py --reps 5 --pop 8192
done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "k1_variants or math_contract" 2>&1 | tail -5
} > gpurun_out/exp_k1_regs.log 2>&1
tail -40 gpurun_out/exp_k1_regs.log
