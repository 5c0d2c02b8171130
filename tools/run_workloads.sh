# the other BASELINE.json configurations through bench.py on one B200
for W in c1 c2 c4; do
  timeout 1500 python bench.py --workload $W --steps 3 --warmup 1 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "$W exit $?"; tail -c 400 gpurun_out/bench_$W.err
done
