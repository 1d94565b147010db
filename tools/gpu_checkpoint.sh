#!/bin/bash
# One-GPU checkpoint: full GPU test-suite, the headline bench, the ncu launch list of the bench command and full
# captures of the three rollout kernels.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1200 gpurun_out/bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rollout_slots -s 2 -c 1 -f -o gpurun_out/k1_final python tools/k1_bench.py --regime converged --reps 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rollout_cartpole_gru -s 1 -c 1 -f -o gpurun_out/gru_final python tools/variants_bench.py gru_converged > /dev/null 2>&1
python tools/variants_bench.py > gpurun_out/variants.log 2>&1; cat gpurun_out/variants.log | cut -c1-200
