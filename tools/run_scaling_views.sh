#!/bin/bash
# C1 single-query latency with the views sharded over the GPUs given (default 1 2 4 8)
mkdir -p gpurun_out
: > gpurun_out/r1_c1_views_sharded.jsonl
P=29700
for n in ${@:-1 2 4 8}; do
  P=$((P+1))
  if [ $n -eq 1 ]; then python tools/bench_localize_views_sharded.py >> gpurun_out/r1_c1_views_sharded.jsonl 2> gpurun_out/r1_c1v_n$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P tools/bench_localize_views_sharded.py >> gpurun_out/r1_c1_views_sharded.jsonl 2> gpurun_out/r1_c1v_n$n.err; fi
done
grep config gpurun_out/r1_c1_views_sharded.jsonl | cut -c1-400
