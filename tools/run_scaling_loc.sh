#!/bin/bash
# C4 end-to-end localisation throughput at 1/2/4/8 GPUs (queries sharded, map replicated)
mkdir -p gpurun_out
: > gpurun_out/r1_c4_sharded.jsonl
P=29600
for n in 1 2 4 8; do
  P=$((P+1))
  if [ $n -eq 1 ]; then python tools/bench_localize_sharded.py >> gpurun_out/r1_c4_sharded.jsonl 2> gpurun_out/r1_c4_n$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P tools/bench_localize_sharded.py >> gpurun_out/r1_c4_sharded.jsonl 2> gpurun_out/r1_c4_n$n.err; fi
done
cut -c1-330 gpurun_out/r1_c4_sharded.jsonl
python -c "import __graft_entry__ as g; g.smoke()"
