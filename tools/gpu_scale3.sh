#!/bin/bash
# 8-GPU box: block-cyclic vs contiguous sharding on BASELINE configs[4] (simple_genetic, P = 2^20) and the headline bench at N = 8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/scale3.jsonl
for sh in cyclic contiguous; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) tools/scale_bench.py --conf cartpole_genetic.yaml --offspring-num 1048576 --elite-num 16 --generations 20 --shard $sh 2>> gpurun_out/scale3.err | grep '^{' >> gpurun_out/scale3.jsonl
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29811 tools/scale_bench.py --conf cartpole_genetic.yaml --offspring-num 1048576 --elite-num 16 --generations 20 2>> gpurun_out/scale3.err | grep '^{' >> gpurun_out/scale3.jsonl
cat gpurun_out/scale3.jsonl
for sh in cyclic contiguous; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29400 + RANDOM % 100)) bench.py --gpus 8 --shard $sh > gpurun_out/scale3_n8_$sh.json 2>> gpurun_out/scale3.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/scale3_n8_$sh.json').read().strip().splitlines()[-1]); print('$sh', 'N', d['n_gpus'], '%.2f G' % (d['value']/1e9), '%.1f gen/s' % d['generations_per_sec'], d['k1_ms_per_generation_over_ranks'], 'e2e %.2f G' % (d['e2e']['value']/1e9))"
done
