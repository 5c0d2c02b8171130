# K1t probe: correctness sections + sustained timing with clock sampling.  $1 = rows, $2 = reps, $3 = cluster (0: sweep)
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader -lms 50 > gpurun_out/clocks_k1t.csv &
SMI=$!
timeout 150 tools/k1t_probe ${1:-10000000} ${2:-60} ${3:-1} > gpurun_out/k1t_probe.log 2>&1
echo exit $? >> gpurun_out/k1t_probe.log
kill $SMI
grep -v "Gdist" gpurun_out/k1t_probe.log
grep "Gdist" gpurun_out/k1t_probe.log | awk '{print $(NF-1)}' | sort -n | awk '{a[NR]=$1} END {print "Gdist/s min",a[1],"med",a[int((NR+1)/2)],"max",a[NR], "n", NR}'
awk -F', ' '{print $1}' gpurun_out/clocks_k1t.csv | sort | uniq -c | sort -rn | head -5
