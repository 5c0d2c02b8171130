# C1 with the views sharded over the ranks given (default 2): sharded tests first, then bench.py --workload c1
LIST=${1:-2}
[ -n "$SKIP_TESTS" ] || timeout 600 python -m pytest tests/test_sharded.py -x -q -m gpu 2>&1 | tail -3
for N in $LIST; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) bench.py --gpus $N --workload c1 --steps 30 > gpurun_out/c1_sharded_n$N.json 2> gpurun_out/c1_sharded_n$N.err; echo "N=$N exit $?"
  grep -h '^{' gpurun_out/c1_sharded_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['ms_per_step'],3), d['localize']['stage_ms'], d.get('sharded_equals_single_gpu'), d['config'].get('sharding'))"
  tail -c 300 gpurun_out/c1_sharded_n$N.err
done
