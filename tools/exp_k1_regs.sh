#!/bin/bash
# GPU experiment: K1 variants with part of the slot's weights in registers (3: W2+b2, 4: W2+b2+b1, 5: b1) vs variant 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for v in 2 3 4 5; do
  echo "== variant $v"; SES_K1_VARIANT=$v python tools/k1_bench.py --reps 5
  SES_K1_VARIANT=$v python tools/k1_bench.py --reps 5 --pop 8192
done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "k1_variants" 2>&1 | tail -5
} > gpurun_out/exp_k1_regs.log 2>&1
tail -40 gpurun_out/exp_k1_regs.log
