# submit / collect pipeline: single-GPU test, 2-GPU worker, headline bench at N = 1 and N = 2
timeout 600 python -m pytest tests/test_knn2_gpu.py -x -q -m gpu -k "submit_collect or device_resident" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_sharded.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/pipe_n1.json 2> gpurun_out/pipe_n1.err; echo "N=1 exit $?"; tail -c 300 gpurun_out/pipe_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/pipe_n2.json 2> gpurun_out/pipe_n2.err; echo "N=2 exit $?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/pipe_n2.err | tail -5
for N in 1 2; do grep -h '^{' gpurun_out/pipe_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'blocking', round(d['e2e']['blocking_calls_value'],1), d['e2e']['pipeline_equals_blocking_calls'], d['parity_vs_oracle_sample'], d['engines_agree'])"; done
