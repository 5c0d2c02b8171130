# two-GPU checks: the NCCL / peer-store tests that skip on one GPU, and the headline bench at N = 2
N=${1:-2}
timeout 900 python -m pytest tests/test_sharded.py -x -q -m gpu > gpurun_out/pytest_sharded.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_sharded.log; tail -5 gpurun_out/pytest_sharded.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err; echo "bench N=$N exit $?"; tail -c 300 gpurun_out/bench_c3_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref N=$N exit $?"
grep -h '^{' gpurun_out/bench_c3_n$N.json | cut -c1-400
grep -h '^{' gpurun_out/bench_ref_n$N.json | cut -c1-300
