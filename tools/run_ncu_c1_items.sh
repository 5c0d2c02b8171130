# one full capture of the item-mode K1t4 launch of a C1 query (1 GPU)
CMD="python tools/probe_resect_trace.py"
timeout 300 $CMD > gpurun_out/plain_c1.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn2_tc4_kernel -s 30 -c 1 -o gpurun_out/r2_k1t4_c1_items $CMD > gpurun_out/ncu_c1_items.log 2>&1
echo "full exit $?"; tail -3 gpurun_out/ncu_c1_items.log
