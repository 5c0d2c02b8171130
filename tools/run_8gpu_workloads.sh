# C5, C2 and C4 on 8 GPUs of one box with the final engines (gpurun --gpus 8)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 8 --workload c5 --steps 5 --warmup 3 > gpurun_out/scale_c5_n8.json 2> gpurun_out/scale_c5_n8.err; echo "c5 exit $?"; grep -h '^{' gpurun_out/scale_c5_n8.json | cut -c1-260
for W in c2 c4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29660 bench.py --gpus 8 --workload $W --steps 3 --warmup 1 > gpurun_out/scale_${W}_n8.json 2> gpurun_out/scale_${W}_n8.err; echo "$W exit $?"; grep -h '^{' gpurun_out/scale_${W}_n8.json | cut -c1-200
done
