#!/bin/bash
# Scaling measurements on one 8-GPU box (gpurun --gpus 8): C3 row-sharded 1/2/4/8, C5 on 8, C2 pair-sharded 1/2/4/8.
mkdir -p gpurun_out
P=29500
for n in 1 2 4 8; do
  P=$((P+1))
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1_scale_n$n.json 2> gpurun_out/r1_scale_n$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r1_scale_n$n.json 2> gpurun_out/r1_scale_n$n.err; fi
  tail -c 300 gpurun_out/r1_scale_n$n.json | cut -c1-200
done
P=$((P+1))
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 5 --warmup 3 --workload c5 > gpurun_out/r1_c5_n8.json 2> gpurun_out/r1_c5_n8.err
cut -c1-200 gpurun_out/r1_c5_n8.json
: > gpurun_out/r1_c2_sharded.jsonl
for n in 1 2 4 8; do
  P=$((P+1))
  if [ $n -eq 1 ]; then python tools/bench_pairs_sharded.py >> gpurun_out/r1_c2_sharded.jsonl 2> gpurun_out/r1_c2_n$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P tools/bench_pairs_sharded.py >> gpurun_out/r1_c2_sharded.jsonl 2> gpurun_out/r1_c2_n$n.err; fi
done
cut -c1-260 gpurun_out/r1_c2_sharded.jsonl
