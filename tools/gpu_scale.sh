#!/bin/bash
# Multi-GPU checkpoint: `gpurun --gpus N -- bash tools/gpu_scale.sh N [nccl]` runs the 2-rank equality tests and the headline
# bench at N ranks with the peer exchange (and, with `nccl`, the NCCL baseline as well).  Everything lands in gpurun_out/scale/.
cd "$(dirname "$0")/.."
N=${1:-8}
O=gpurun_out/scale
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > $O/pytest_multi_n$N.log 2>&1; tail -2 $O/pytest_multi_n$N.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N > $O/scale_n$N.json 2>> $O/scale.err
if [ "$2" = nccl ]; then
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --exchange nccl --no-regimes --no-extra > $O/scale_n${N}_nccl.json 2>> $O/scale.err
fi
for f in $O/scale_n$N*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "N", d["n_gpus"], "%.2f G env-steps/s" % (d["value"] / 1e9), "%.1f gen/s" % d["generations_per_sec"], "ms/gen %.3f" % d["ms_per_step"], "e2e %.2f G" % (d["e2e"]["value"] / 1e9), d["config"]["fitness_exchange"], d.get("rank_consistency", {}).get("identical"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
tail -3 $O/scale.err
