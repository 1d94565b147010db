#!/bin/bash
# Multi-GPU checkpoint (run with `gpurun --gpus 8`): 2-rank equality tests, then the headline bench at N = 1, 2, 4, 8 with the
# peer exchange, and N = 8 with the NCCL baseline.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi.log 2>&1; tail -3 gpurun_out/pytest_multi.log
python bench.py --no-cpu > gpurun_out/scale_n1.json 2> gpurun_out/scale.err
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n > gpurun_out/scale_n$n.json 2>> gpurun_out/scale.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --exchange nccl > gpurun_out/scale_n8_nccl.json 2>> gpurun_out/scale.err
for f in gpurun_out/scale_n*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "N", d["n_gpus"], "%.2f G env-steps/s" % (d["value"] / 1e9), "%.1f gen/s" % d["generations_per_sec"], "ms/gen %.3f" % d["ms_per_step"], "e2e %.2f G" % (d["e2e"]["value"] / 1e9), d["config"]["fitness_exchange"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
