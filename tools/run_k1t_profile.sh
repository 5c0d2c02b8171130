# ncu capture of one steady-state K1t launch (4096 x 2M, cluster size $1, default 1)
CL=${1:-1}
timeout 100 tools/k1t_probe 2000000 4 $CL 1 > gpurun_out/plain.log 2>&1 && timeout 250 ncu --set full --clock-control none --import-source on -k regex:knn2_tc_kernel -s 2 -c 1 -o gpurun_out/k1t_prof tools/k1t_probe 2000000 4 $CL 1 > gpurun_out/ncu_k1t.log 2>&1
tail -3 gpurun_out/ncu_k1t.log
