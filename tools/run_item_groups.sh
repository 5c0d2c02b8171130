# item-mode K1t4 with two vs three epilogue groups: tests, then C1 / C2 / C4 with either
timeout 900 python -m pytest tests/test_knn2_tc_gpu.py tests/test_match_gpu.py tests/test_engine_gpu.py -x -q -m gpu 2>&1 | tail -3
for G in ${GROUPS_LIST:-2 3}; do
  for W in c1 c2 c4; do
    HULO_TC4_ITEM_GROUPS=$G timeout 600 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/items_g${G}_$W.json 2> gpurun_out/items_g${G}_$W.err
    echo "G=$G $W exit $?: $(grep -h '^{' gpurun_out/items_g${G}_$W.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['metric'], round(d['value'],1), d['unit'], round(d['ms_per_step'],3))")"
  done
done
