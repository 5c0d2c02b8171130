timeout 900 python -m pytest tests/test_match_gpu.py tests/test_engine_gpu.py tests/test_knn2_gpu.py tests/test_knn2_tc_gpu.py -x -q -m gpu > gpurun_out/pytest_items.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_items.log; tail -15 gpurun_out/pytest_items.log
for W in c1 c2 c4; do
  timeout 900 python bench.py --workload $W --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "$W exit $?"; tail -c 300 gpurun_out/bench_$W.err
done
