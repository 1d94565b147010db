#!/bin/bash
# GPU experiment: K1 variants (0 scalar FFMA, 1 packed FFMA2, 2 packed without the Newton step) -- parity, peaks, throughput
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
python - <<'PY'
import ctypes as C
from simple_es_b200 import _lib
lib = _lib.load()
a = C.c_double(); b = C.c_double()
lib.ses_measure_fp32_peak(0, C.byref(a)); lib.ses_measure_fp32x2_peak(0, C.byref(b))
print("PEAK ffma TF/s", a.value, "ffma2 TF/s", b.value)
bad = C.c_uint64(0)
for newton in (1, 0):
    lib.ses_test_tanh_x2_exhaustive(newton, 0.0, 3.0e38, C.byref(bad))
    print("tanh_x2 exhaustive newton=%d mismatches=%d" % (newton, bad.value))
PY
for v in 0 1 2; do
  echo "== variant $v"; SES_K1_VARIANT=$v python tools/k1_bench.py --reps 5
  SES_K1_VARIANT=$v python tools/k1_bench.py --reps 5 --pop 8192
done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "math_contract or k1_variants or rollout_philox or verification_mode_matches or edge_cases or trained_parent" 2>&1 | tail -5
} > gpurun_out/exp_k1_variants.log 2>&1
tail -40 gpurun_out/exp_k1_variants.log
