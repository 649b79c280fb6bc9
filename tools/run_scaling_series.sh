#!/bin/bash
# Multi-GPU series of round 2 (one 8-GPU B200 box): cfg 1 weak scaling with the token-sharded check, and the
# fixed-global-batch (strong scaling) series SURVEY.md 8d asks for on cfg 4 / cfg 5 — batch sharding (global batch 8
# split over the ranks) and, for cfg 5's 65 536-token axis, token sharding (every rank streams 1/N of the tokens of
# the same 8 samples). One JSON line per run into gpurun_out/scaling_r2.jsonl.
out=gpurun_out/scaling_r2.jsonl
: > $out
port=29600
run() {  # n_gpus, extra args...
  n=$1; shift
  port=$((port + 1))
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --no-cpu "$@" 2>/dev/null | tail -1 >> $out
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --no-cpu "$@" 2>/dev/null | tail -1 >> $out
  fi
}
G=${1:-8}
for n in 2 4 8; do [ $n -le $G ] && run $n --steps 10 --warmup 3; done                          # cfg 1, weak, + token_sharded key
for w in cfg4 cfg5; do
  for n in 1 2 4 8; do [ $n -le $G ] && run $n --workload $w --batch $((8 / n)) --steps 6 --warmup 3; done
done
for n in 2 4 8; do [ $n -le $G ] && run $n --workload cfg5 --batch 8 --shard tokens --steps 6 --warmup 3; done
wc -l $out
