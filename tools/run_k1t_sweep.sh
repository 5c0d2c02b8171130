# sustained K1t throughput per cluster size (timing only), 4096 x $1 rows, $2 launches each
for CL in 1 2 4; do
  timeout 150 tools/k1t_probe ${1:-10000000} ${2:-60} $CL 1 > gpurun_out/k1t_sweep_cl$CL.log 2>&1
  grep "Gdist" gpurun_out/k1t_sweep_cl$CL.log | awk '{print $(NF-1)}' | awk -v cl=$CL '{a[NR]=$1} END {s=0; for(i=int(NR/2)+1;i<=NR;i++) s+=a[i]; print "cluster",cl,"first",a[1],"mean of second half",s/(NR-int(NR/2)), "n",NR}'
done
