# one full capture of K3 on the 2000 x 300 x 500 shape (1 GPU)
CMD="python tools/bench_geom.py --one 2000 300 0.6 500"
timeout 300 $CMD > gpurun_out/plain_k3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fmatrix -s 1 -c 1 -o gpurun_out/r2_k3 $CMD > gpurun_out/ncu_k3.log 2>&1
echo "full exit $?"; tail -3 gpurun_out/ncu_k3.log
