mkdir -p gpurun_out/gru
export SES_B200_TEST_BUILD=1
for v in 1 2 3; do
SES_GRU_VARIANT=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rollout_cartpole_gru --launch-skip 1 -c 1 -o /tmp/gru_v$v -f python tools/variants_bench.py gru_converged > gpurun_out/gru/ncu_v$v.log 2>&1
ncu -i /tmp/gru_v$v.ncu-rep --page raw --csv > gpurun_out/gru/gru_v$v.raw.csv 2>/dev/null
ncu -i /tmp/gru_v$v.ncu-rep --page details > gpurun_out/gru/gru_v$v.details.txt 2>/dev/null
done
cp /tmp/gru_v1.ncu-rep gpurun_out/gru/
ls -la gpurun_out/gru
