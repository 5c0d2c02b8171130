# launch list and one full capture of the top kernel, from the bench command itself (1 GPU)
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_bench.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list exit $?"
timeout 600 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn2_tc4_kernel -s 3 -c 1 -o gpurun_out/r2_k1t4_c3 $CMD > gpurun_out/ncu_full.log 2>&1
echo "full exit $?"; tail -3 gpurun_out/ncu_full.log
