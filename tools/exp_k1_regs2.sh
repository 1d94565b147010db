#!/bin/bash
# GPU experiment: K1 register-resident variants, 1-warp CTAs, ncu of variant 4, FFMA2/DFMA operand-form micro-benchmark
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
./tools/micro/ffma2_forms.bin
for cfg in "4 4" "4 1" "7 1" "6 1"; do
  set -- $cfg
  echo "== variant $1 cta_warps $2"; SES_K1_VARIANT=$1 SES_ROLLOUT_CTA_WARPS=$2 python tools/k1_bench.py --reps 5
  SES_K1_VARIANT=$1 SES_ROLLOUT_CTA_WARPS=$2 python tools/k1_bench.py --reps 5 --pop 8192
done
for v in 4 6 7; do SES_K1_VARIANT=$v python - <<'PY'
import os, numpy as np, torch, sys
sys.path.insert(0, os.getcwd())
from oracle import twin
from simple_es_b200.engine import RolloutEngine
P, E, D = 2048, 5, 226
for sigma, seed in ((2.0, 3), (0.05, 4)):
    mu = np.zeros((1, D), np.float32)
    if sigma < 1:
        mu[0, :4] = [0.0, 0.5, 10.0, 3.0]; mu[0, 160 + 32] = 5.0; mu[0, 160] = -5.0
    eng = RolloutEngine("CartPole-v1", 4, 2, False, False, 500, E, P, P, 1, 1, seed=seed)
    fit, steps = eng.rollout(1, sigma, torch.from_numpy(mu).cuda())
    tf, ts = twin.population_cartpole(mu, sigma=sigma, seed=seed, gen=1, group=P, n_head=1, n=P, E=E, nthreads=8)
    print("variant", os.environ["SES_K1_VARIANT"], "sigma", sigma, "bit-exact", np.array_equal(steps.cpu().numpy(), ts), int(ts.sum()))
PY
done
SES_K1_VARIANT=4 ncu --set full --clock-control none --import-source on -k regex:k_rollout_slots -s 2 -c 1 -f -o gpurun_out/k1v4 python tools/k1_bench.py --regime converged --reps 1 | tail -2
} > gpurun_out/exp_k1_regs2.log 2>&1
tail -60 gpurun_out/exp_k1_regs2.log
