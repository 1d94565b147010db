#!/bin/bash
# 8-GPU box: BASELINE configs[4] (CartPole simple_genetic, 16 elites) over P = 2^16 .. 2^20 and N = 1 .. 8, plus weak scaling of
# the headline config (openai_es, 65536 offspring per GPU).  Lines land in gpurun_out/scale2.jsonl.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/scale2.jsonl
run() { n=$1; shift; if [ "$n" = 1 ]; then python tools/scale_bench.py "$@"; else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) tools/scale_bench.py "$@"; fi 2>> gpurun_out/scale2.err | grep '^{' >> gpurun_out/scale2.jsonl; }
for n in 1 2 4 8; do run $n --conf cartpole_genetic.yaml --offspring-num 1048576 --elite-num 16 --generations 20; done
for p in 65536 262144; do run 8 --conf cartpole_genetic.yaml --offspring-num $p --elite-num 16 --generations 20; done
run 8 --conf cartpole_genetic.yaml --offspring-num 1048576 --elite-num 16 --generations 20 --shard contiguous
for n in 2 8; do run $n --conf cartpole_openai.yaml --offspring-num $((65536 * n)) --generations 40; done
cat gpurun_out/scale2.jsonl
