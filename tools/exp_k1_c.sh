#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "math_contract or k1_variants or rollout_philox or verification_mode or edge_cases or trained_parent or gru" 2>&1 | tail -5
python tools/k1_bench.py --reps 5
python tools/k1_bench.py --reps 3 --pop 8192
python tools/variants_bench.py
} > gpurun_out/exp_k1_c.log 2>&1
tail -30 gpurun_out/exp_k1_c.log
