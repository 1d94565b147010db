#!/bin/bash
# First gpurun call of the next round (one B200, ~3 minutes): everything that was written after round 1's GPU budget ran out
# gets measured or verified in one go.  Results land in gpurun_out/r2_first/.
#   /usr/local/graft/bin/gpurun --timeout 400 -- 'bash tools/r2_first_call.sh'
cd "$(dirname "$0")/.."
O=gpurun_out/r2_first; mkdir -p $O
# 1. the whole single-GPU suite (incl. the last-sorted file with the end-of-round additions)
timeout 300 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
# 2. K1 variants: default (4), branch-free division (7), speculative physics (6); throughput regime, one rank's share at N = 8, generation 0
for v in 4 7 6; do for p in 65536 8192; do SES_K1_VARIANT=$v python tools/k1_bench.py --pop $p >> $O/k1_variants.jsonl 2>> $O/err.log; done; done
cat $O/k1_variants.jsonl | cut -c1-260
# 2b. GRU rollout: default kernel vs SES_GRU_VARIANT=1 (physics for both actions on idle lanes at the start of the step)
for v in 0 1; do echo "{\"SES_GRU_VARIANT\": $v}" >> $O/gru_variants.jsonl; SES_GRU_VARIANT=$v python tools/variants_bench.py gru_converged >> $O/gru_variants.jsonl 2>> $O/err.log; done
cat $O/gru_variants.jsonl | cut -c1-220
# 3. K2 fused vs separate kernels (repeat of round 1's A/B), then the headline bench line and the reference arm
python tools/k2_bench.py > $O/k2_ab.log 2>&1; cp gpurun_out/k2_ab.jsonl $O/ 2>/dev/null
python bench.py > $O/bench_n1.json 2>> $O/err.log; tail -c 600 $O/bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2>> $O/err.log
# 4. launch list of the bench command (kernel shares) with the fused K2 default
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench_under_ncu.log 2>&1
