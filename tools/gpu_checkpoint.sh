#!/bin/bash
# One-GPU checkpoint: full GPU test-suite, the headline bench (both arms), the ncu launch list of the bench command and full
# captures of the rollout kernels, summarised on the box (the reports themselves stay in /tmp: gpurun_out/ is capped at 64 MiB).
# Everything lands in gpurun_out/final/.
cd "$(dirname "$0")/.."
O=gpurun_out/final
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 600 $O/bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2>> $O/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 10 --warmup 3 --no-cpu --no-regimes --no-extra > $O/bench_under_ncu.log 2>&1
python tools/ncu_summary.py list $O/launches.csv > $O/launches_bench_n1.txt
ncu --set full --clock-control none --import-source on -k regex:k_rollout_slots -s 2 -c 1 -f -o /tmp/k1_conv python tools/k1_bench.py --regime converged --reps 1 > /dev/null 2>&1
python tools/ncu_summary.py raw /tmp/k1_conv.ncu-rep > $O/k1_conv.txt
ncu --set full --clock-control none --import-source on -k regex:k_rollout_slots -s 2 -c 1 -f -o /tmp/k1_gen0 python tools/k1_bench.py --regime gen0 --reps 1 > /dev/null 2>&1
python tools/ncu_summary.py raw /tmp/k1_gen0.ncu-rep > $O/k1_gen0.txt
ncu --set full --clock-control none --import-source on -k regex:k_rollout_slots -s 4 -c 1 -f -o /tmp/k1_conv8k python tools/k1_rounds.py --points 8192:0:0:7 --reps 1 > /dev/null 2>&1
python tools/ncu_summary.py raw /tmp/k1_conv8k.ncu-rep > $O/k1_conv8k.txt
python tools/variants_bench.py > $O/variants.log 2>&1; cut -c1-200 $O/variants.log
