# headline bench at 1, 2, 4, 8 GPUs on one box (gpurun --gpus 8)
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    timeout 900 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/scale_c3_n$N.json 2> gpurun_out/scale_c3_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 40 --warmup 3 > gpurun_out/scale_c3_n$N.json 2> gpurun_out/scale_c3_n$N.err
  fi
  echo "c3 N=$N exit $?"; grep -h '^{' gpurun_out/scale_c3_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), d['engines'], d['parity_vs_oracle_sample'], d['clocks'])"
done
